// Internal declarations shared by the sm_100a kernels of libprb200 (not part of the ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <atomic>

#include "prb200.h"

namespace prb {

// ---------------------------------------------------------------------------------------------------------------------
// Device-side kernel specification, passed BY VALUE as a kernel parameter (constant bank => uniform loads).
// ---------------------------------------------------------------------------------------------------------------------
struct DevFactor {
  int32_t kind;
  int32_t n_dims;
  int32_t power;
  int32_t pad;
  int32_t dims[PRB_MAX_DIMS];
  double inv_ls[PRB_MAX_DIMS];  // Matern: 1/l ; categorical: 1/(l * n_dims)
};
struct DevTerm {
  double outputscale;
  int32_t additive;
  int32_t n_factors;
  DevFactor f[PRB_MAX_FACTORS];
};
struct DevSpec {
  int32_t n_terms;
  int32_t d_k;
  DevTerm t[PRB_MAX_TERMS];
  // Distinct factor functions (kind, dims, lengthscales): "mixed_categorical" lists its factors in both of its terms
  // (kernels.py:85-94), so the hot kernels evaluate each distinct function once.  uid[t][f] = index of factor (t, f)'s
  // function, defined by factor (first_t[u], first_f[u]); at most PRB_MAX_UNIQUE of them, else n_unique = 0 (no sharing).
  int8_t n_unique;
  int8_t uid[PRB_MAX_TERMS][PRB_MAX_FACTORS];
  int8_t first_t[4], first_f[4];
};

struct DevSpace {
  int32_t d, n_int, n_cat, apply_numeric;
  int32_t n_disc, d_k, n_cont, raw_int;
  int32_t int_cols[PRB_MAX_DIMS];
  double int_lo[PRB_MAX_DIMS];
  double int_hi[PRB_MAX_DIMS];
  int32_t cat_start[PRB_MAX_CAT];
  int32_t cat_card[PRB_MAX_CAT];
  int32_t cont_cols[PRB_MAX_DIMS];  // theta columns that are continuous (same index in the kernel-input layout)
  double tau;
  // latent categorical embedding (latent_tab == nullptr: off): categorical j -> latent_dim[j] kernel-input columns starting
  // at cat_start[0] + latent_col[j], read from row `category` of the table at latent_tab + latent_off[j]
  int32_t latent_dim[PRB_MAX_CAT];
  int32_t latent_col[PRB_MAX_CAT];
  int32_t latent_off[PRB_MAX_CAT];
  const double* latent_tab;
};

struct DevAcq {
  int32_t kind, ei_variant;
  double param, min_variance, acq_var_floor;
};

constexpr int MODEL_ONES = 4096;
constexpr int GEMM_BM = 128;  // row tile of the triangular products (prb_gemm_tma.cu)
constexpr int GEMM_BK = 16;   // contraction tile; Kp is a multiple of it

constexpr int GRAD_JSPLIT_ROWS = 128;  // rows of W handled by one CTA of the gradient contraction

// number of kernels this library has launched (prb_launch_count(); bench.py reports it as gpu_launches)
extern std::atomic<unsigned long long> g_launches;
inline void count_launch(int k = 1) { g_launches.fetch_add((unsigned long long)k, std::memory_order_relaxed); }

// Tuning / A-B knobs from the environment, read ONCE per process (first use), never on a hot call.
struct EnvKnobs {
  bool no_pdl;          // PRB_NO_PDL: plain stream-ordered launches inside the fused steps
  bool no_fused_step;   // PRB_NO_FUSED_STEP: the unfused multi-launch sequence everywhere
  long panel_mb;        // PRB_PANEL_MB: column-panel budget override (0 = default)
  int force_bn;         // PRB_BN: column-tile width override of the separable products (0 = pick by tile count)
  int group_waves;      // PRB_GROUP_WAVES: scheduling-group size in waves of CTAs (default 4; 0 = row-tile-major)
  int mixed_order;      // PRB_MIXED_ORDER: heavy / light alternating tile order (1: all groups but the last, 2: all, 0: off)
};
const EnvKnobs& env_knobs();

// ---- optional per-kernel timing (prb_profile_enable / prb_profile_read): CUDA events around every launch ---------------
enum KernelId {
  K_SAMPLE = 0, K_KSTAR, K_TRGEMM_LOWER, K_EPILOGUE, K_TRGEMM_UPPER, K_GRAD_CONTRACT, K_SUM_SPLITS, K_MC_REDUCE,
  K_MA_UPDATE, K_ANALYTIC, K_ADAM, K_MISC, K_COUNT
};
struct ProfScope {  // no-op unless profiling is enabled; never use while the stream is being captured into a graph
  ProfScope(int id, cudaStream_t st);
  ~ProfScope();
  int id_;
  cudaStream_t st_;
  cudaEvent_t beg_, end_;
  bool on_;
};

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------
// Inside a PdlScope the launchers below attach cudaLaunchAttributeProgrammaticStreamSerialization: the next kernel of the
// fused optimisation step is scheduled while the previous one still runs, executes its prologue (barrier init, descriptor
// prefetch) and blocks in griddepcontrol.wait until the predecessor has completed and flushed.  Every kernel of the step
// calls grid_dep_launch() first and grid_dep_wait() before it touches anything a predecessor wrote.
extern thread_local bool g_pdl_on;
struct PdlScope {
  explicit PdlScope(bool on) : prev_(g_pdl_on) { g_pdl_on = on; }
  ~PdlScope() { g_pdl_on = prev_; }
  bool prev_;
};
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = g_pdl_on ? attr : nullptr;
  cfg.numAttrs = g_pdl_on ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
inline int64_t round_up64(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// opaque 128-byte TMA descriptor (CUtensorMap) so that only prb_gemm_tma.cu needs <cuda.h>
struct alignas(64) TensorMap {
  unsigned char bytes[128];
};

// per-call arguments of the separable fast path (operand generation + fused gradient epilogue)
struct SepArgs {
  const double* kc;       // [nb, Kp]
  const double* Dc;       // [nb, n_cont, Kp]
  const uint32_t* zbits;  // panel-local [ncols_pad]
  int n_cont, mc, nb;
  int64_t col0;
  // MC de-duplication: columns are the UNIQUE configurations, packed restart by restart
  const int* colb;       // [ncols_pad] restart of each column (NULL: column / mc)
  const int* ncols_dev;  // device scalar: total number of unique columns (NULL: every column of the panel is valid)
};

struct BitMap {  // integer parameter k of the PR space -> bit position in the binary-factor code (-1: not a binary dim)
  int32_t pos[PRB_MAX_DIMS];
};

struct FusedAcq {  // acquisition epilogue fused into the lower product (single row tile)
  DevAcq aq;
  double *a, *dmu, *dvar;
  int ncols;
};

struct DimMap {  // S_part slot -> kernel-input column
  int32_t n;
  int32_t dims[PRB_MAX_DIMS];
};

// Pathwise-gradient group of the generic path: the Matern factor(s) -- identical dims and lengthscales, e.g. the ARD factor
// that appears in both terms of "mixed_categorical" -- through which the continuous PR parameters enter the kernel.
//   d k*(col, j) / d x_dim = Q[col, j] * w[dim] * (x_dim - X[j, dim]),   Q = sum over member factors of
//   outputscale_t * (d term / d F_f) * G_f,   G = (dF/dr)/r.
// kstar_q_kernel writes Q next to k*; the upper product contracts it in its epilogue (MODE_UPPER_GRADQ).
constexpr int GQ_MAX_DIMS = 16;
struct GradGroup {
  int32_t n_dims;
  int32_t dims[GQ_MAX_DIMS];  // kernel-input columns that carry a gradient
  double w[GQ_MAX_DIMS];      // 1 / lengthscale^2
  int8_t member[PRB_MAX_TERMS][PRB_MAX_FACTORS];
};
bool build_grad_group(const DevSpec& spec, uint64_t diffmask, GradGroup* out);

struct Model {
  int n, d_k;
  int Kp;        // round_up(n, 16): padded contraction length / leading dimension of the eval-major panels
  int Mp1;       // round_up(n + 1, 128): rows of A1 = [Linv ; alpha^T ; 0]
  int Mp2;       // round_up(n, 128): rows of A2 = Linv^T
  void* block;   // GP models: ONE stream-ordered pool allocation that every device array below points into
  double* A1;    // [Mp1, Kp] row-major
  double* A2;    // [Mp2, Kp] row-major (upper triangular: A2[m, k] = Linv[k, m], k >= m)
  double* Xtr;   // [n, d_k]
  double* alpha; // [n]
  double mean_const;
  double kss;    // k(x, x)
  DevSpec spec;
  int device;
  // ---- second-generation engine (TMA operands) ----
  TensorMap mapA1, mapA2;    // 2-D maps (box = 4 k values x 128 rows): the single-kernel step of prb_small.cu
  TensorMap mapA1k, mapA2k;  // 3-D maps (box = 4 k values x 128 rows x 4 k-chunks): one TMA per operand and k-tile
  double* alpha_pad;  // [max(Mp1, Mp2)] alpha, zero padded
  double* ones;       // [MODEL_ONES] 1.0: the unit upstream gradient of the host-buffer entry (no fill launch per call)
  // separable structure: one term = Matern52(continuous dims) x isotropic Matern52(binary dims), binary training columns
  int sep_ok;
  int sep_cont_factor, sep_bin_factor;  // indices into spec.t[0].f
  int sep_nbits;                        // number of binary dims (<= 32)
  uint32_t* Zbits;                      // [max(Mp1, Mp2)] bit codes of the training points
  double* sep_tab;                      // [64] binary-factor value by Hamming distance
  int opt_dedup;                        // PRB_OPT_DEDUP
  int opt_unfused;                      // PRB_OPT_UNFUSED: plain calls take the multi-launch sequence (cross-check path)
  // completion counter of the fused step kernels' "last CTA" logic: zeroed once here, re-armed by the kernels themselves
  unsigned int* done_ctr;
  // cached executable graphs of the fused optimisation loops (slot 0: prb_mc_adam, 1: prb_analytic_adam).  A later call
  // re-captures its step (a few microseconds) and UPDATES the cached executable in place instead of instantiating a new one;
  // executables that could not be updated (different launch topology) are parked and destroyed with the model, so no call
  // ever has to wait for the stream to drain.
  cudaGraphExec_t graph_exec[2];
  cudaGraphExec_t graph_dead[16];
  int n_graph_dead;
  // random-Fourier-feature posterior sample (prb_rff_model_create): f(x) = rff_scale * sum_j rff_w[j] cos(x . W[:, j] + b_j);
  // rff_m > 0 marks the model as deterministic -- no GP state, every engine call goes to rff_eval_kernel
  int rff_m;
  double* rff_W;  // [d_k, m]
  double* rff_b;  // [m]
  double* rff_w;  // [m]
  double rff_scale;
};

constexpr int DEDUP_MAX_BITS = 12;  // histogram de-duplication: 2^12 bins of shared memory per restart

// prb_small.cu: one kernel per optimisation step for experiment-sized separable problems
bool small_step_eligible(const Model& m, const DevSpace& sp, const DimMap& dm, int b, int mc);
bool small_gen_eligible(const Model& m, const DevSpace& sp, DimMap* dm, int b, int mc);  // generic single-term variant
size_t small_step_scratch_bytes(int b, int mc, int d);
cudaError_t small_step_init();
cudaError_t small_step_timestamps(long long* host_out);
cudaError_t launch_small_step(const Model& m, bool gen, const DevSpace& sp, const DevAcq& aq, const BitMap& bm,
                              const DimMap& dm, const double* theta_in, const double* lower, const double* upper, double* Xc, int b, int mc,
                              const double* u_int, const double* u_cat, uint64_t seed, uint64_t offset, int estimator,
                              double* ma_state, int update_ma, bool need_grad, const double* grad_out, double* out_value,
                              double* out_grad, double* adam_theta, double* adam_m, double* adam_v, double lr,
                              int* step_ptr, void* scratch, cudaStream_t st);

bool small_adam_persistent_ok(int b, int mc);
cudaError_t launch_small_adam_persistent(const Model& m, bool gen, const DevSpace& sp, const DevAcq& aq, const BitMap& bm,
                                         const DimMap& dm, double* theta, const double* lower, const double* upper, int b,
                                         int mc, const double* u_int, const double* u_cat, uint64_t seed, uint64_t offset,
                                         int estimator, double* ma_state, int update_ma, double* step_value,
                                         double* final_value, double* adam_m, double* adam_v, double lr, int maxiter,
                                         int step0, void* scratch, cudaStream_t st);

// prb_gemm_tma.cu
cudaError_t tma_init();
cudaError_t make_operand_map(TensorMap* out, const double* base, int rows, int k_extent, int ld, int box_rows = 128);
cudaError_t make_operand_map3(TensorMap* out, const double* base, int rows, int k_extent, int ld, int box_rows = 128);
cudaError_t launch_tma_lower_gen(const Model& m, const SepArgs& sa, int ncols_pad, double* VT, double* norm_part,
                                 double* mu_dot, const FusedAcq* fa, cudaStream_t st);
cudaError_t launch_tma_lower(const Model& m, const double* panelK, int panel_rows, int ncols_pad, double* VT,
                             double* norm_part, double* mu_dot, const FusedAcq* fa, cudaStream_t st);
cudaError_t launch_tma_upper(const Model& m, const double* panelV, int panel_rows, int ncols_pad, double* W,
                             cudaStream_t st);
cudaError_t launch_tma_upper_gradq(const Model& m, const double* panelV, int panel_rows, int ncols, int ncols_pad,
                                   const GradGroup& gq, const double* Qp, const double* xk, const double* dmu,
                                   const double* dvar, double* S_part, cudaStream_t st);
cudaError_t launch_tma_upper_grad(const Model& m, const double* panelV, int panel_rows, const SepArgs& sa, int ncols,
                                  int ncols_pad, const double* dmu, const double* dvar, double* S_part,
                                  cudaStream_t st);

}  // namespace prb
