// One kernel per optimisation step for experiment-sized problems (n + 1 <= 128, mc % 32 == 0): the separable "mixed"
// kernel over binary integers (small_step_kernel<false>) and single-term Matern products over integers of any range
// (small_step_kernel<true>, see small_eval_block).
//
// At the reference's own sizes (ackley13: n = 20 .. 119, 5 restarts x 128 MC samples per step, 800 Adam steps per BO
// iteration, gen.py:320-337) a step is ~7e7 flop: microseconds of a B200, so wall time is the latency of the kernels in
// the dependency chain.  The fused four-kernel step (prb_kernels.cu) costs ~42 us; this kernel does the whole step in one
// launch.  A CTA owns 32 evaluation columns (MC samples) of one restart end to end:
//
//   TMA: the packed factor A1 = [Linv ; alpha^T] (<= 128 x 128 doubles) -> shared memory, overlapped with
//   prep:     clamp(theta), rounding probabilities, the 32 samples (bit codes), per-restart kernel factors kc / Dc
//   generate: k*[col, k] = kc[k] * tab[popc(z_col ^ Z_k)]                       -> shared (B operand)
//   lower:    v = Linv k*, mu = alpha . k*        (DMMA, zero 8 x 4 sub-blocks of the triangle skipped)
//   epilogue: ||v||^2, mu -> EI / UCB / mean and derivatives; v^T written over k* in shared memory (B operand of the next)
//   upper:    w = Linv^T v, reading A1 TRANSPOSED out of the same shared tile (conflict-free: a fragment is 2 x 128 B)
//   epilogue: S[c, col] = sum_j (dA/dmu alpha_j - 2 dA/dvar w_j) kb(z, Z_j) Dc[c, j]
//   reduce:   per-tile partial sums of value / pathwise / score-function terms -> global; the LAST tile of a restart
//             finishes the MC means and applies the Adam update; the LAST restart advances the EMA baseline and the step.
//
// Nothing but 14 doubles per tile leaves the SM.  Reductions are fixed-order (deterministic).
#include <cuda.h>

#include <cstdlib>

#include "prb_device.cuh"
#include "prb_kernels.cuh"
#include "prb_ptx.cuh"

namespace prb {

constexpr int S_THREADS = 256;
constexpr int S_BN = 32;
constexpr int S_MAXC = 8;  // continuous dims supported by this kernel

struct SmallStepParams {
  DevSpace sp;
  BitMap bm;
  DimMap dm;
  DevTerm term;  // the model's (single) product term
  int cf;        // separable kernel: index of the continuous factor in `term`
  int gf;        // generic kernel: index of the Matern factor that holds every continuous PR parameter (-1: none)
  double gw[8];  // generic kernel: 1 / lengthscale^2 of continuous parameter c inside factor gf
  uint32_t contmask;  // generic kernel: kernel-input dims that are continuous PR parameters
  DevAcq aq;
  double outputscale, mean_const, kss;
  const double* Xtr;
  int n, d_k, nchunks;  // nchunks = ceil(n / 4): k-chunks that hold non-zero columns of A1
  const double* alpha_pad;
  const uint32_t* Zbits;
  const double* tab;
  int ntab;
  const double* theta;  // [nb, d] optimiser parameters (read, clamped)
  const double* lower;
  const double* upper;
  double* Xc;  // clamped rows out (may alias theta for the in-place final clamp), or NULL
  int mc, nb, tiles;
  const double* u_int;
  const double* u_cat;
  uint64_t seed, offset;
  int estimator;
  double* ma_state;
  int update_ma, need_grad;
  const double* grad_out;
  double* out_value;
  double* out_grad;
  double* adam_theta;
  double* adam_m;
  double* adam_v;
  double lr;
  int* step_ptr;
  double* partial;           // [nb, tiles, 1 + d]
  unsigned int* tile_done;   // [nb]
  unsigned int* done;        // [1]
  // persistent mode (cooperative launch, every CTA resident): `iters` = maxiter Adam steps + the final clamp-and-evaluate,
  // separated by a grid-wide barrier on `release`; single-step mode: iters = 1, persistent = 0
  int iters, persistent, step0;
  unsigned int* release;     // [1]
  double* final_value;       // [nb] values of the final evaluation (persistent mode)
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* ptr) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}

// shared-memory carve-up (doubles unless noted)
constexpr int SM_A = 128 * 128;        // A1, [k-chunk][row][4]
constexpr int SM_B = 32 * 128;         // k* tile, then v^T tile: [k-chunk][col][4]
constexpr int SM_KC = 128;
constexpr int SM_DC = S_MAXC * 128;
constexpr int SM_AL = 128;
constexpr int SM_RED = 5 * 32 + 32;
constexpr int SM_S = S_MAXC * 32;
constexpr int SM_ACQ = 3 * 32;
constexpr int SM_VEC = 4 * 32;         // th, ps, offs, g
constexpr int SM_TAB = 64;
constexpr int SM_U = 32 * 32;          // base uniforms of this tile's 32 samples (fixed for the life of the kernel)
constexpr int G_MAXD = 16;             // generic variant: kernel-input dimension limit
constexpr int SM_X = 128 * G_MAXD;     // training inputs: separable [row][S_MAXC] (continuous factor); generic [dim][128]
constexpr int SM_XK = G_MAXD * 32;     // generic variant: kernel inputs of this tile's 32 samples, [dim][col]
constexpr int SM_PART = 8 * 33;        // staging of the per-tile partial sums in the finish chain
constexpr int SM_DOUBLES =
    SM_A + SM_B + SM_KC + SM_DC + SM_AL + SM_RED + SM_S + SM_ACQ + SM_VEC + SM_TAB + SM_U + SM_X + SM_XK + SM_PART;
constexpr size_t SMALL_SMEM_BYTES = size_t(SM_DOUBLES) * 8 + 128 * 4 + 32 * 4 + 32 * 32 + 16 + 16;

// phase timestamps of CTA 0 (clock64), read back by prb_debug_small_timestamps(): diagnostic for the latency budget
__device__ long long g_small_ts[16];
#define SMALL_TS(k)                                                               \
  do {                                                                            \
    if (blockIdx.x == 0 && threadIdx.x == 0 && ts_on) g_small_ts[k] = clock64(); \
  } while (0)

// Generic variant (GEN): a single product term of Matern-5/2 factors (power 1 or 2) over arbitrary dims -- integer parameters of
// any range inside an ARD factor (rosenbrock10_scaled, mixed_int_f1: kernels.py:52-66 with non-binary integers).  k* is
// evaluated entry by entry from the sampled kernel inputs; the pathwise gradient needs d k / d x_c only for the continuous
// parameters, which all sit in factor `gf`:  d k / d x_c = Q(row, col) * gw[c] * (x_c - X[row, c]) with
// Q = outputscale * p F_gf^(p-1) G_gf * prod_{g != gf} F_g^p_g   (WITH_G: returns Q instead of k).
// Branch-free sqrt / exp for the block evaluation below: the library versions carry slow-path branches, which keep the
// compiler from interleaving the 16 independent evaluations of a thread.  Arguments are bounded here (r2 in [1e-30, 1e8],
// exponent in [-690, 0]); both are accurate to about one ulp, far inside the 1e-9 parity tolerance.
// The evaluation is written for a block of NR training rows x NC columns per thread so that the FP64 dependency chains
// (distance accumulation, sqrt, exp) of the entries interleave: with 8 warps per SM a one-entry-at-a-time loop is bound by
// instruction latency, not throughput.  The continuous parameters are the same for every column of a restart: their share
// of factor gf's squared distance is computed once per training row (rc_s) and the dims in `contmask` are skipped here.
template <int NR, int NC>
__device__ __forceinline__ void small_eval_block(const DevTerm& T, int gf, uint32_t contmask,
                                                 const double* __restrict__ rc_s, const double* __restrict__ xk_s,
                                                 const int (&cols)[NC], const double* __restrict__ X_s,
                                                 const int (&rows)[NR], double (&kv)[NR][NC], double (&qv)[NR][NC]) {
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int c = 0; c < NC; ++c) kv[r][c] = qv[r][c] = T.outputscale;
  for (int f = 0; f < T.n_factors; ++f) {  // every factor is Matern-5/2 with power 1 or 2 (small_gen_setup)
    const DevFactor& F = T.f[f];
    const bool grad_here = f == gf;
    const bool squared = F.power == 2;
    double r2[NR][NC];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const double base = grad_here ? rc_s[rows[r]] : 0.0;
#pragma unroll
      for (int c = 0; c < NC; ++c) r2[r][c] = base;
    }
#pragma unroll 1
    for (int di = 0; di < F.n_dims; ++di) {
      const int dim = F.dims[di];
      if (grad_here && ((contmask >> dim) & 1u)) continue;
      const double ils = F.inv_ls[di];
      double xc[NC], xr[NR];
#pragma unroll
      for (int c = 0; c < NC; ++c) xc[c] = xk_s[dim * 32 + cols[c]];
#pragma unroll
      for (int r = 0; r < NR; ++r) xr[r] = X_s[dim * 128 + rows[r]];
#pragma unroll
      for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const double df = (xc[c] - xr[r]) * ils;
          r2[r][c] = fma(df, df, r2[r][c]);
        }
    }
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        double G;
        const double v = matern52_nb(r2[r][c], &G);
        kv[r][c] *= squared ? v * v : v;
        qv[r][c] *= grad_here ? (squared ? 2.0 * v * G : G) : (squared ? v * v : v);
      }
  }
}

template <bool GEN>
__global__ void __launch_bounds__(S_THREADS, 1)
    small_step_kernel(const __grid_constant__ CUtensorMap tmA, const SmallStepParams p) {
  extern __shared__ __align__(128) unsigned char smraw[];
  double* A_s = reinterpret_cast<double*>(smraw);
  double* B_s = A_s + SM_A;
  double* kc_s = B_s + SM_B;
  double* Dc_s = kc_s + SM_KC;
  double* al_s = Dc_s + SM_DC;
  double* red = al_s + SM_AL;
  double* S_s = red + SM_RED;
  double* a_s = S_s + SM_S;
  double* dmu_s = a_s + 32;
  double* dvar_s = dmu_s + 32;
  double* th_s = dvar_s + 32;
  double* ps_s = th_s + 32;
  double* offs_s = ps_s + 32;
  double* g_s = offs_s + 32;
  double* tab_s = g_s + 32;
  double* u_s = tab_s + SM_TAB;   // [32][n_int]
  double* X_s = u_s + SM_U;       // [128][S_MAXC], GEN: [d_k][128]
  double* xk_s = X_s + SM_X;      // GEN: [d_k][32]
  double* pt_s = xk_s + SM_XK;    // [8][33]
  uint32_t* Zb_s = reinterpret_cast<uint32_t*>(pt_s + SM_PART);  // [128]
  uint32_t* zb_s = Zb_s + 128;                                   // [32]
  uint8_t* z_s = reinterpret_cast<uint8_t*>(zb_s + 32);          // [32][n_disc <= 32]
  uint64_t* bar = reinterpret_cast<uint64_t*>(z_s + 32 * 32);
  int* flag_s = reinterpret_cast<int*>(bar + 1);

  const DevSpace& sp = p.sp;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;  // 4 warps along the 128 rows, 2 along the 32 columns
  const int bi = blockIdx.x / p.tiles, tile = blockIdx.x % p.tiles;
  const int i0 = tile * S_BN;  // first MC sample of this tile
  const int n = p.n;

  bool ts_on = true;  // timestamps: the launch prologue, then the second iteration (a steady-state gradient step) if any
  SMALL_TS(0);
  // ---- 0. A1 -> shared memory by TMA (only the k-chunks that hold non-zero columns) --------------------------------
  if (tid == 0) {
    tma_prefetch_desc(&tmA);
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (warp == 0) {  // one box (4 k values x 128 rows) per lane: the 32 TMA issues go out in parallel
    if (lane == 0) mbar_arrive_expect_tx(bar, uint32_t(p.nchunks) * 128 * 4 * 8);
    __syncwarp();
    if (lane < p.nchunks) tma_load_2d(A_s + lane * 512, &tmA, bar, 4 * lane, 0);
  }
  const DevFactor& fc = p.term.f[p.cf];
  if constexpr (!GEN) {
    if (tid < 64) tab_s[tid] = (tid <= p.ntab) ? p.tab[tid] : 0.0;
  }
  if (tid < 128) {
    al_s[tid] = p.alpha_pad[tid];
    if constexpr (GEN) {
      for (int dd = 0; dd < p.d_k; ++dd) X_s[dd * 128 + tid] = (tid < n) ? p.Xtr[size_t(tid) * p.d_k + dd] : 0.0;
    } else {
      Zb_s[tid] = p.Zbits[tid];
      for (int di = 0; di < fc.n_dims; ++di)
        X_s[tid * S_MAXC + di] = (tid < n) ? p.Xtr[size_t(tid) * p.d_k + fc.dims[di]] : 0.0;
    }
  }
  // The base uniforms of this tile's samples never change between steps (the reference caches its Sobol base samples, the
  // Philox stream is keyed on (sample, parameter) only): fetched / generated once, kept in shared memory.
  for (int e = tid; e < S_BN * sp.n_int; e += S_THREADS) {
    const int col = e / sp.n_int, k = e % sp.n_int;
    u_s[e] = p.u_int ? p.u_int[size_t(i0 + col) * sp.n_int + k]
                     : philox_uniform(p.seed, p.offset, uint32_t(i0 + col), uint32_t(k));
  }
  const DimMap& dmref = p.dm;
  for (int it = 0; it < p.iters; ++it) {
  ts_on = (it == (p.iters > 2 ? 1 : p.iters - 1));
  SMALL_TS(11);
  // persistent mode: the last iteration is the final clamp (in place) + value-only evaluation (gen.py:338-341)
  const bool fin = p.persistent && (it == p.iters - 1);
  const int need_grad = fin ? 0 : p.need_grad;
  const bool do_adam = (p.adam_theta != nullptr) && !fin;
  double* const Xc = fin ? p.adam_theta : p.Xc;
  double* const out_value = fin ? p.final_value : p.out_value;
  // Everything that depends on the pre-step optimiser / EMA state is read before this CTA signals completion.  The three
  // pow() calls (EMA de-biasing, Adam bias corrections) are FP64 and slow: one thread does them off the critical path.
  // State that another CTA rewrote in the previous iteration is read with ld.cg (L1 is not coherent across SMs).
  if (tid == S_THREADS - 1) {
    double baseline = 0.0;
    if (p.estimator == PRB_EST_REINFORCE_MA && p.ma_state != nullptr) {
      const double cnt = __ldcg(p.ma_state);
      if (cnt != 0.0) baseline = __ldcg(p.ma_state + 1) / (1.0 - pow(0.7, cnt));
    }
    g_s[0] = baseline;
    if (do_adam) {
      const double t_step = p.persistent ? double(p.step0 + it + 1) : (p.step_ptr ? double(*p.step_ptr + 1) : 1.0);
      g_s[1] = 1.0 - pow(0.9, t_step);
      g_s[2] = 1.0 - pow(0.999, t_step);
    }
  }

  SMALL_TS(1);
  // ---- 1. prep: clamp, rounding probabilities, this tile's 32 samples, per-restart kernel factors ----------------------
  if (tid < sp.d) {
    double v = __ldcg(p.theta + size_t(bi) * sp.d + tid);
    if (p.lower) v = fmin(fmax(v, p.lower[tid]), p.upper[tid]);
    th_s[tid] = v;
    if (Xc && tile == 0) Xc[size_t(bi) * sp.d + tid] = v;
  }
  if (tid < S_BN) zb_s[tid] = 0u;
  __syncthreads();
  if (tid < 128) {
    // warps 0-3: rounding probabilities, then the 32 samples x n_int binary parameters of this tile, one (sample, parameter)
    // pair per thread and pass; the barrier between the two is a named barrier over these four warps only
    pr_restart_probs(sp, th_s, tid, 128, ps_s, offs_s);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int col = tid & 31;
    for (int k = tid >> 5; k < sp.n_int; k += 4) {
      uint8_t zk;
      const double tn = pr_sample_int(sp, th_s, ps_s, offs_s, col, k, u_s, 0, 0, &zk);  // u = u_s[col * n_int + k]
      z_s[col * sp.n_disc + k] = zk;
      if constexpr (GEN) {
        xk_s[sp.int_cols[k] * 32 + col] = tn;
      } else {
        if (p.bm.pos[k] >= 0 && tn != 0.0) atomicOr(&zb_s[col], 1u << p.bm.pos[k]);
      }
    }
  } else if constexpr (GEN) {
    // warps 4-7: continuous kernel inputs of the 32 columns and the restart's gradient directions gw_c (x_c - X[j, c])
    const int j = tid - 128;
    for (int e = j; e < 32 * dmref.n; e += 128) {
      const int dim = dmref.dims[e >> 5];
      xk_s[dim * 32 + (e & 31)] = th_s[dim];
    }
    double rc = 0.0;
    for (int c = 0; c < dmref.n; ++c) {
      const int dim = dmref.dims[c];
      const double df = th_s[dim] - X_s[dim * 128 + j];
      Dc_s[c * 128 + j] = (j < n) ? p.gw[c] * df : 0.0;
      rc = fma(p.gw[c] * df, df, rc);
    }
    kc_s[j] = rc;  // continuous share of factor gf's squared distance
  } else {
    // warps 4-7: the restart's continuous-factor values and derivatives for the 128 (padded) training points
    const int j = tid - 128;
    double val = 0.0, G = 0.0;
    if (j < n) {
      double r2 = 0.0;
      for (int di = 0; di < fc.n_dims; ++di) {
        const int dim = fc.dims[di];
        const double df = (th_s[dim] - X_s[j * S_MAXC + di]) * fc.inv_ls[di];
        r2 = fma(df, df, r2);
      }
      val = p.outputscale * matern52(r2, &G);
    }
    kc_s[j] = val;
    for (int di = 0; di < fc.n_dims; ++di) {
      const int dim = fc.dims[di];
      double v = 0.0;
      if (j < n) v = p.outputscale * G * (fc.inv_ls[di] * fc.inv_ls[di]) * (th_s[dim] - X_s[j * S_MAXC + di]);
      Dc_s[di * 128 + j] = v;
    }
  }
  __syncthreads();

  SMALL_TS(2);
  // row ownership as in prb_gemm_tma.cu: M-warp wm owns the 8-row groups {wm, 7 - wm, 8 + wm, 15 - wm}
  const int frag_r = lane >> 2, frag_k = lane & 3;
  int grp[4];
  grp[0] = wm;
  grp[1] = 7 - wm;
  grp[2] = 8 + wm;
  grp[3] = 15 - wm;
  [[maybe_unused]] double qk[4][4];  // GEN: gradient factor Q(row group i, column q) of this thread's accumulator entries

  // ---- 2. generate the B operand: k*[col, k] for the 32 columns x 4 nchunks k values -------------------------------------
  {
    // lane <-> (k within a chunk, 8 columns): a warp store covers 32 consecutive doubles of B_s (conflict-free)
    const int k4 = tid & 3;
    const int col = (tid >> 2) & 31;
    const int chi = tid >> 7;  // chunks chi * 16 .. chi * 16 + 15
    if constexpr (GEN) {
      // Every thread evaluates exactly the (row, column) entries its accumulators own in the two products: the value goes
      // to the B operand, the gradient factor Q stays in registers for the contraction after the upper product.
      int cl[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) cl[q] = wn * 16 + 8 * (q >> 1) + 2 * frag_k + (q & 1);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int rw[2] = {8 * grp[2 * h] + frag_r, 8 * grp[2 * h + 1] + frag_r};
        double kv[2][4], qv[2][4];
        if (8 * min(grp[2 * h], grp[2 * h + 1]) < n) {  // warp-uniform: skip row groups that are all padding
          small_eval_block<2, 4>(p.term, p.gf, p.contmask, kc_s, xk_s, cl, X_s, rw, kv, qv);
        } else {
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) kv[r][q] = qv[r][q] = 0.0;
        }
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int row = rw[r];
            B_s[(row >> 2) * 128 + cl[q] * 4 + (row & 3)] = (row < n) ? kv[r][q] : 0.0;
            qk[2 * h + r][q] = qv[r][q];
          }
      }
    } else {
      const uint32_t zc = zb_s[col];
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const int c = chi * 16 + t;
        const int k = 4 * c + k4;
        B_s[c * 128 + col * 4 + k4] = kc_s[k] * tab_s[__popc(zc ^ Zb_s[k])];
      }
    }
  }
  __syncthreads();
  SMALL_TS(3);
  if (it == 0) mbar_wait(bar, 0);  // A1 stays resident for every later iteration
  SMALL_TS(4);

  const int fb_off = (wn * 16 + frag_r) * 4 + frag_k;
  const int rows_valid = n + 1;  // rows of A1 at or beyond this index are zero

  // ---- 3. lower product: V[row, col] = sum_{k <= row} A1[row, k] k*[col, k] ---------------------------------------------
  // A warp owns only 4 x 2 accumulator tiles here, and a DMMA that depends on the previous one waits ~200 cycles: the
  // contraction is split over NS independent accumulator sets (k-chunk c -> set c % NS), summed in a fixed order at the end.
  constexpr int NS = 4;
  double acc[4][2][2];
  int fa_off[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) fa_off[i] = (8 * grp[i] + frag_r) * 4 + frag_k;
  // Sub-block (8-row group g) x (k-chunk c) of the lower triangle is non-zero for c < 2 g + 2: each accumulator tile has a
  // contiguous chunk range, walked with NS chunks in flight and no branch inside the unrolled body.
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int cend = (8 * grp[i] < rows_valid) ? min(p.nchunks, 2 * grp[i] + 2) : 0;
    double as[NS][2][2];
#pragma unroll
    for (int s = 0; s < NS; ++s) as[s][0][0] = as[s][0][1] = as[s][1][0] = as[s][1][1] = 0.0;
    int c = 0;
    for (; c + NS <= cend; c += NS) {
      double a[NS], b0[NS], b1[NS];
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        a[s] = A_s[(c + s) * 512 + fa_off[i]];
        b0[s] = B_s[(c + s) * 128 + fb_off];
        b1[s] = B_s[(c + s) * 128 + fb_off + 32];
      }
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        dmma884(as[s][0][0], as[s][0][1], a[s], b0[s]);
        dmma884(as[s][1][0], as[s][1][1], a[s], b1[s]);
      }
    }
    for (; c < cend; ++c) {
      const double a = A_s[c * 512 + fa_off[i]];
      dmma884(as[0][0][0], as[0][0][1], a, B_s[c * 128 + fb_off]);
      dmma884(as[0][1][0], as[0][1][1], a, B_s[c * 128 + fb_off + 32]);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) acc[i][j][e] = (as[0][j][e] + as[1][j][e]) + (as[2][j][e] + as[3][j][e]);
  }
  __syncthreads();  // every warp is done reading k*: B_s is rewritten with v^T below
  SMALL_TS(5);

  // ---- 4. epilogue: ||v||^2, mu, acquisition; v^T -> B_s as the next B operand -------------------------------------------
  {
    double nrm[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = 8 * grp[i] + frag_r;
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = wn * 16 + 8 * j + 2 * frag_k + e;
          const double v = acc[i][j][e];
          if (row == n) {
            red[5 * 32 + col] = v;  // alpha . k*
            B_s[(row >> 2) * 128 + col * 4 + (row & 3)] = 0.0;  // the alpha row is not part of v
          } else {
            nrm[j][e] = fma(v, v, nrm[j][e]);
            B_s[(row >> 2) * 128 + col * 4 + (row & 3)] = v;
          }
        }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        double v = nrm[j][e];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (frag_r == 0) red[wm * 32 + wn * 16 + 8 * j + 2 * frag_k + e] = v;
      }
  }
  __syncthreads();
  if (tid < S_BN) {
    const double s = ((red[tid] + red[32 + tid]) + red[64 + tid]) + red[96 + tid];
    double val, d_mu, d_var, var;
    acq_eval(p.aq, p.mean_const + red[5 * 32 + tid], p.kss - s, &val, &d_mu, &d_var, &var);
    a_s[tid] = val;
    dmu_s[tid] = d_mu;
    dvar_s[tid] = d_var;
  }
  // no barrier here: a_s / dmu_s / dvar_s are first read behind the barriers of the gradient contraction / partial sums

  SMALL_TS(6);
  if (need_grad) {
    // ---- 5. upper product: W[m, col] = sum_{k >= m} Linv[k, m] V[k, col], A1 read transposed out of A_s ---------------
    int ft_off[4];  // element (m = 8 g + r, k) of Linv^T lives at A_s[(m / 4) * 512 + k * 4 + m % 4]
#pragma unroll
    for (int i = 0; i < 4; ++i) ft_off[i] = (2 * grp[i] + (frag_r >> 2)) * 512 + frag_k * 4 + (frag_r & 3);
    // sub-block (group g) x (chunk c) of the upper triangle is non-zero for c >= 2 g
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int cbeg = (8 * grp[i] < n) ? 2 * grp[i] : p.nchunks;
      double as[NS][2][2];
#pragma unroll
      for (int s = 0; s < NS; ++s) as[s][0][0] = as[s][0][1] = as[s][1][0] = as[s][1][1] = 0.0;
      int c = cbeg;
      for (; c + NS <= p.nchunks; c += NS) {
        double a[NS], b0[NS], b1[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          a[s] = A_s[ft_off[i] + (c + s) * 16];
          b0[s] = B_s[(c + s) * 128 + fb_off];
          b1[s] = B_s[(c + s) * 128 + fb_off + 32];
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          dmma884(as[s][0][0], as[s][0][1], a[s], b0[s]);
          dmma884(as[s][1][0], as[s][1][1], a[s], b1[s]);
        }
      }
      for (; c < p.nchunks; ++c) {
        const double a = A_s[ft_off[i] + c * 16];
        dmma884(as[0][0][0], as[0][0][1], a, B_s[c * 128 + fb_off]);
        dmma884(as[0][1][0], as[0][1][1], a, B_s[c * 128 + fb_off + 32]);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) acc[i][j][e] = (as[0][j][e] + as[1][j][e]) + (as[2][j][e] + as[3][j][e]);
    }
    __syncthreads();  // dmu_s / dvar_s written by warp 0 after the lower product are visible from here on
    SMALL_TS(7);
    // ---- 6. gradient contraction: S[c, col] = sum_m (dA/dmu alpha_m - 2 dA/dvar W[m, col]) kb(z_col, Z_m) Dc[c, m] -------
    int rows[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) rows[i] = 8 * grp[i] + frag_r;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = wn * 16 + 8 * j + 2 * frag_k + e;
        const double cm = dmu_s[col];
        const double cv = -2.0 * dvar_s[col];
        if constexpr (GEN) {
          if (dmref.n > 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              acc[i][j][e] = (rows[i] < n) ? fma(cv, acc[i][j][e], cm * al_s[rows[i]]) * qk[i][2 * j + e] : 0.0;
          }
        } else {
          const uint32_t zc = zb_s[col];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            acc[i][j][e] =
                (rows[i] < n) ? fma(cv, acc[i][j][e], cm * al_s[rows[i]]) * tab_s[__popc(zc ^ Zb_s[rows[i]])] : 0.0;
        }
      }
    for (int c = 0; c < dmref.n; ++c) {
      double part[2][2];
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          double s = 0.0;
#pragma unroll
          for (int i = 0; i < 4; ++i) s = fma(acc[i][j][e], Dc_s[c * 128 + rows[i]], s);
          s += __shfl_xor_sync(0xffffffffu, s, 4);
          s += __shfl_xor_sync(0xffffffffu, s, 8);
          s += __shfl_xor_sync(0xffffffffu, s, 16);
          part[j][e] = s;
        }
      __syncthreads();  // red from the previous use has been consumed
      if (frag_r == 0) {
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) red[wm * 32 + wn * 16 + 8 * j + 2 * frag_k + e] = part[j][e];
      }
      __syncthreads();
      if (tid < S_BN) S_s[c * 32 + tid] = ((red[tid] + red[32 + tid]) + red[64 + tid]) + red[96 + tid];
    }
    __syncthreads();
  }

  SMALL_TS(8);
  // ---- 7. per-tile partial sums over the 32 columns (one warp per output), then the completion chain --------------------
  const int d = sp.d;
  double* my_part = p.partial + (size_t(bi) * p.tiles + tile) * (1 + d);
  const int n_tasks = need_grad ? 1 + dmref.n + sp.n_int : 1;
  for (int task = warp; task < n_tasks; task += S_THREADS / 32) {
    if (task == 0) {
      const double t = warp_sum(a_s[lane]);
      if (lane == 0) my_part[0] = t;
    } else if (task <= dmref.n) {
      const double t = warp_sum(S_s[(task - 1) * 32 + lane]);
      if (lane == 0) my_part[1 + dmref.dims[task - 1]] = t;
    } else {
      const int kd = task - 1 - dmref.n;
      const double pk = ps_s[kd];
      const double zi = double(z_s[lane * sp.n_disc + kd]);
      const double t = warp_sum(((a_s[lane] - g_s[0]) * (zi - pk)) / sp.tau);
      if (lane == 0) my_part[1 + sp.int_cols[kd]] = t;
    }
  }
  SMALL_TS(9);
  // ---- 8. completion: ONE ticket over all CTAs of the step.  The last CTA to arrive finishes the step for every restart:
  // MC means from the per-tile partial sums (fixed order), Adam update, EMA baseline, and opens the grid barrier.
  __threadfence();
  __syncthreads();
  if (tid == 0) flag_s[0] = (atomicAdd(p.done, 1u) == unsigned(p.nb * p.tiles - 1)) ? 1 : 0;
  __syncthreads();
  SMALL_TS(10);
  if (flag_s[0]) {
    __threadfence();
    const int nslot = need_grad ? 1 + d : 1;
    for (int e = tid; e < p.nb * nslot; e += S_THREADS) {
      const int rb = e / nslot, sl = e % nslot;
      const double* src = p.partial + size_t(rb) * p.tiles * (1 + d) + sl;
      double s = 0.0;
#pragma unroll 4
      for (int t = 0; t < p.tiles; ++t) s += __ldcg(src + size_t(t) * (1 + d));
      s /= double(p.mc);
      if (sl == 0) {
        out_value[rb] = s;
        if (rb < SM_PART) pt_s[rb] = s;
      } else {
        const int c = sl - 1;
        const double go = p.grad_out ? p.grad_out[rb] : 1.0;
        const double gacq = go * s;
        if (p.out_grad) p.out_grad[size_t(rb) * d + c] = gacq;
        if (do_adam) {  // torch.optim.Adam on loss = -acq(X).sum() (as adam_step_kernel)
          const size_t idx = size_t(rb) * d + c;
          const double g = -gacq;
          const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
          const double m_old = __ldcg(p.adam_m + idx);
          const double mi = m_old + (g - m_old) * (1.0 - b1);
          const double vi = __ldcg(p.adam_v + idx) * b2 + (1.0 - b2) * g * g;
          p.adam_m[idx] = mi;
          p.adam_v[idx] = vi;
          p.adam_theta[idx] = __ldcg(p.adam_theta + idx) - (p.lr / g_s[1]) * (mi / (sqrt(vi) / sqrt(g_s[2]) + eps));
        }
      }
    }
    __syncthreads();
    if (warp == 0) {
      // EMA baseline over ALL restarts (probabilistic_reparameterization.py:499-505), fixed summation order
      double v = 0.0;
      for (int i = lane; i < p.nb; i += 32) v += (i < SM_PART) ? pt_s[i] : out_value[i];
      v = warp_sum(v);
      if (lane == 0) {
        if (p.update_ma && p.ma_state != nullptr) {
          const double mean = v / double(p.nb);
          const double hidden = __ldcg(p.ma_state + 1);
          p.ma_state[0] = __ldcg(p.ma_state) + 1.0;
          p.ma_state[1] = hidden - (hidden - mean) * (1.0 - 0.7);
        }
        if (!p.persistent && p.step_ptr && do_adam) *p.step_ptr += 1;
        *p.done = 0u;
        if (p.persistent) {  // open the grid barrier: everything this step wrote is visible before the flag is
          __threadfence();
          atomicExch(p.release, unsigned(it + 1));
        }
      }
    }
  }
  SMALL_TS(12);
  if (p.persistent && it + 1 < p.iters) {
    // grid-wide barrier between steps (all CTAs are co-resident: cooperative launch); bounded spin, trap instead of hanging
    if (tid == 0) {
      unsigned int spins = 0;
      while (ld_acquire_u32(p.release) < unsigned(it + 1)) {
        if (++spins > (1u << 28)) __trap();
      }
    }
    __syncthreads();
  }
  SMALL_TS(13);
  }  // iterations
}

cudaError_t small_step_timestamps(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g_small_ts, sizeof(long long) * 16);
}

bool small_step_eligible(const Model& m, const DevSpace& sp, const DimMap& dm, int b, int mc) {
  static const bool off = getenv("PRB_NO_SMALL_STEP") != nullptr;
  return !off && m.sep_ok && m.Mp1 == 128 && m.Kp <= 128 && mc % S_BN == 0 && mc >= S_BN && dm.n <= S_MAXC &&
         sp.n_disc <= 32 && sp.n_cat == 0 && sp.d <= 32 && b >= 1;
}

// Generic variant: one product term; every continuous PR parameter inside one Matern factor; no categoricals.
static bool small_gen_setup(const Model& m, const DevSpace& sp, DimMap* dm, int* gf_out, double* gw) {
  if (m.spec.n_terms != 1 || m.spec.t[0].additive || sp.n_cat != 0 || sp.raw_int || sp.d != m.d_k || m.d_k > G_MAXD ||
      sp.n_cont > S_MAXC || sp.n_disc > 32)
    return false;
  const DevTerm& T = m.spec.t[0];
  for (int f = 0; f < T.n_factors; ++f)
    if (T.f[f].kind != PRB_FACTOR_MATERN52 || T.f[f].power < 1 || T.f[f].power > 2) return false;
  int gf = -1;
  dm->n = sp.n_cont;
  for (int c = 0; c < sp.n_cont; ++c) {
    const int dim = sp.cont_cols[c];
    dm->dims[c] = dim;
    int hits = 0;
    for (int f = 0; f < T.n_factors; ++f)
      for (int di = 0; di < T.f[f].n_dims; ++di)
        if (T.f[f].dims[di] == dim) {
          if (gf >= 0 && gf != f) return false;
          gf = f;
          ++hits;
          if (gw) gw[c] = T.f[f].inv_ls[di] * T.f[f].inv_ls[di];
        }
    if (hits > 1) return false;              // the same dim listed twice in a factor
    if (hits == 0 && gw) gw[c] = 0.0;        // a parameter the kernel ignores: zero pathwise gradient
  }
  *gf_out = gf;
  return true;
}

bool small_gen_eligible(const Model& m, const DevSpace& sp, DimMap* dm, int b, int mc) {
  static const bool off = getenv("PRB_NO_SMALL_STEP") != nullptr || getenv("PRB_NO_SMALL_GEN") != nullptr;
  int gf;
  return !off && m.Mp1 == 128 && m.Kp <= 128 && mc % S_BN == 0 && mc >= S_BN && b >= 1 &&
         small_gen_setup(m, sp, dm, &gf, nullptr);
}

size_t small_step_scratch_bytes(int b, int mc, int d) {
  return (size_t(b) * (mc / S_BN) * (1 + d)) * sizeof(double) + size_t(b + 2) * sizeof(unsigned int) + 64;
}

cudaError_t small_step_init() {
  cudaError_t e = cudaFuncSetAttribute(small_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       int(SMALL_SMEM_BYTES));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(small_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              int(SMALL_SMEM_BYTES));
}

static void fill_small_params(SmallStepParams& p, bool gen, const Model& m, const DevSpace& sp, const DevAcq& aq,
                              const BitMap& bm, const DimMap& dm, int b, int mc, const double* u_int, const double* u_cat,
                              uint64_t seed, uint64_t offset, int estimator, double* ma_state, int update_ma,
                              void* scratch) {
  memset(&p, 0, sizeof(p));
  p.sp = sp;
  p.bm = bm;
  p.dm = dm;
  p.term = m.spec.t[0];
  if (gen) {
    DimMap dm2;
    small_gen_setup(m, sp, &dm2, &p.gf, p.gw);
    p.dm = dm2;
    for (int c = 0; c < dm2.n; ++c) p.contmask |= 1u << dm2.dims[c];
  } else {
    p.cf = m.sep_cont_factor;
  }
  p.aq = aq;
  p.outputscale = m.spec.t[0].outputscale;
  p.mean_const = m.mean_const;
  p.kss = m.kss;
  p.Xtr = m.Xtr;
  p.n = m.n;
  p.d_k = m.d_k;
  p.nchunks = (m.n + 3) / 4;
  p.alpha_pad = m.alpha_pad;
  p.Zbits = m.Zbits;
  p.tab = m.sep_tab;
  p.ntab = m.sep_nbits;
  p.mc = mc;
  p.nb = b;
  p.tiles = mc / S_BN;
  p.u_int = u_int;
  p.u_cat = u_cat;
  p.seed = seed;
  p.offset = offset;
  p.estimator = estimator;
  p.ma_state = ma_state;
  p.update_ma = update_ma;
  p.partial = static_cast<double*>(scratch);
  p.tile_done = reinterpret_cast<unsigned int*>(p.partial + size_t(b) * p.tiles * (1 + sp.d));
  p.done = p.tile_done + b;
  p.release = p.done + 1;
  p.iters = 1;
}

cudaError_t launch_small_step(const Model& m, bool gen, const DevSpace& sp, const DevAcq& aq, const BitMap& bm,
                              const DimMap& dm, const double* theta_in, const double* lower, const double* upper,
                              double* Xc, int b, int mc, const double* u_int, const double* u_cat, uint64_t seed,
                              uint64_t offset, int estimator, double* ma_state, int update_ma, bool need_grad,
                              const double* grad_out, double* out_value, double* out_grad, double* adam_theta,
                              double* adam_m, double* adam_v, double lr, int* step_ptr, void* scratch, cudaStream_t st) {
  ProfScope prof_(K_ADAM, st);
  SmallStepParams p;
  fill_small_params(p, gen, m, sp, aq, bm, dm, b, mc, u_int, u_cat, seed, offset, estimator, ma_state, update_ma, scratch);
  p.theta = theta_in;
  p.lower = lower;
  p.upper = upper;
  p.Xc = Xc;
  p.need_grad = need_grad ? 1 : 0;
  p.grad_out = grad_out;
  p.out_value = out_value;
  p.out_grad = out_grad;
  p.adam_theta = adam_theta;
  p.adam_m = adam_m;
  p.adam_v = adam_v;
  p.lr = lr;
  p.step_ptr = step_ptr;
  const CUtensorMap& tm = *reinterpret_cast<const CUtensorMap*>(&m.mapA1);
  if (gen)
    small_step_kernel<true><<<b * p.tiles, S_THREADS, SMALL_SMEM_BYTES, st>>>(tm, p);
  else
    small_step_kernel<false><<<b * p.tiles, S_THREADS, SMALL_SMEM_BYTES, st>>>(tm, p);
  count_launch();
  return cudaGetLastError();
}

// Can the whole multi-start Adam run as ONE cooperative launch?  Every CTA must be resident (one per SM: ~210 KB of
// shared memory each) because the steps are separated by a grid-wide barrier.
bool small_adam_persistent_ok(int b, int mc) {
  static int sms = -1, coop = 0;
  if (sms < 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess)
      sms = 0;
  }
  static const bool off = getenv("PRB_NO_PERSISTENT") != nullptr;
  return !off && coop && b * (mc / S_BN) <= sms;
}

// gen_candidates_torch (gen.py:304-343) as one persistent kernel: maxiter x (clamp -> value + gradient -> Adam), then the
// final clamp and value-only evaluation; the steps are separated by a grid barrier instead of kernel boundaries.
cudaError_t launch_small_adam_persistent(const Model& m, bool gen, const DevSpace& sp, const DevAcq& aq, const BitMap& bm,
                                         const DimMap& dm, double* theta, const double* lower, const double* upper, int b,
                                         int mc, const double* u_int, const double* u_cat, uint64_t seed, uint64_t offset,
                                         int estimator, double* ma_state, int update_ma, double* step_value,
                                         double* final_value, double* adam_m, double* adam_v, double lr, int maxiter,
                                         int step0, void* scratch, cudaStream_t st) {
  ProfScope prof_(K_ADAM, st);
  SmallStepParams p;
  fill_small_params(p, gen, m, sp, aq, bm, dm, b, mc, u_int, u_cat, seed, offset, estimator, ma_state, update_ma, scratch);
  p.theta = theta;
  p.lower = lower;
  p.upper = upper;
  p.Xc = nullptr;
  p.need_grad = 1;
  p.out_value = step_value;
  p.final_value = final_value;
  p.adam_theta = theta;
  p.adam_m = adam_m;
  p.adam_v = adam_v;
  p.lr = lr;
  p.iters = maxiter + 1;
  p.persistent = 1;
  p.step0 = step0;
  CUtensorMap tm = *reinterpret_cast<const CUtensorMap*>(&m.mapA1);
  void* args[] = {&tm, &p};
  const void* fn = gen ? reinterpret_cast<const void*>(&small_step_kernel<true>)
                       : reinterpret_cast<const void*>(&small_step_kernel<false>);
  cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(b * p.tiles), dim3(S_THREADS), args, SMALL_SMEM_BYTES, st);
  g_launches += 1;  // one launch; it stands for (maxiter + 1) fused steps
  return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace prb
