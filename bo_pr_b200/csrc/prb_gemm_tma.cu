// FP64 triangular panel products, second generation: TMA + mbarrier pipeline feeding the sm_100a DMMA pipe, with the
// cross-covariance generated inside the kernel and the pathwise-gradient contraction fused into the epilogue.
//
//   lower:  V^T[col, j] = sum_{k <= j} A1[j, k] * k*[col, k]      A1 = [Linv ; alpha^T]   (posterior variance + mean)
//   upper:  W[j, col]   = sum_{k >= j} A2[j, k] * V^T[col, k]     A2 = Linv^T            (K^-1 k*, for d sigma^2 / d x)
//
// Reference: the dense k* @ L^-T GEMM inside GPyTorch's exact prediction strategy ([3P], reached from
// probabilistic_reparameterization.py:491) and its autograd backward (:624-633).
//
// sm_100a has no tcgen05 kind for FP64; the FP64 tensor path is mma.sync DMMA (37.1 TFLOP/s measured), whose accumulators
// live in registers, so there is no TMEM stage.  What IS Blackwell-native here: operand tiles arrive by TMA
// (cp.async.bulk.tensor, one elected thread, no per-thread address arithmetic) into a dense [k-chunk][row][4] layout that
// makes every DMMA fragment load a fully coalesced, conflict-free 256-byte LDS.64; stage hand-off is by full/empty
// mbarriers (no __syncthreads in the main loop, producer runs two k-tiles ahead).  The first-generation kernel of this
// round (cp.async + bar.sync per k-tile, materialised operands; removed) reached 56 % of the DMMA peak: the LDGSTS address
// registers were recycled by the next LDS (long-scoreboard stall, 15 % of samples) and every k-tile ended in a CTA barrier.
//
// Tuning record (B200, n = 1000, 65 536 columns; profiles/): 4 stages / producer 2 k-tiles ahead beats 5/3, 5/2 and 4/3;
// 8 warps of 32 x 64 (246-254 registers, 1 CTA/SM) beat 16 warps of 32 x 32 (128 registers): 4.49 vs 4.79 ms per step;
// two 4-warp CTAs per SM with 64-column tiles (same 32 x 64 warp tile) beat one 8-warp CTA by 1-2 % (4.41 vs 4.46 ms);
// routing the whole last row tile through the masked (diagonal) body to skip its all-padding row groups costs more than
// the 1.5 % of DMMAs it saves (4.55 vs 4.46 ms).  Diagonal k-tiles (22 % of the k-tiles at n = 1000, 0.53 of a full
// tile's DMMAs) cost 0.63 of a full k-tile in warp time (ncu source page): predicating the DMMAs instead of branching
// around them is slower (4.73 vs 4.40 ms: every predicated DMMA gets its own WARPSYNC), and dispatching each k-tile to
// one of nine branch-free bodies specialised on the warp's activity pattern changes nothing (4.41 ms) -- what remains
// per row tile is the CTA prologue / epilogue (~7 % of warp time), not the masked body.  One warp per sub-partition
// already reaches 97 % of the DMMA peak from registers (tools/microbench/dmma_occupancy.cu), so two 4-warp CTAs per SM
// cover each other's epilogues.  A persistent variant (2 CTAs per SM striding over the heavy-first tile list, mbarrier
// pipeline kept alive across tiles, next tile's first TMA loads issued before the epilogue, separate epilogue scratch) was
// built and measured SLOWER: 4.59 ms, and 4.52 ms for the same restructured code launched one tile per CTA -- the
// hardware's own CTA dispatch de-synchronises the two CTAs of an SM, whereas persistent CTAs walk equally long tiles in
// lockstep and hit their epilogues together.  A second attempt that left the tile body untouched (stage parities offset
// by a per-stage bit mask, scratch in the last stage buffer, prefetch into stages 0-1) gave the same picture: 4.57 ms with
// one tile per CTA -- merely wrapping the tile body in a loop costs 4 % through ptxas' scheduling of the hot loop -- and
// 4.66 ms persistent.  Moving the per-dim CTA barriers of the UPPER_GRAD epilogue into one also lost (2.12 -> 2.15 ms),
// as did unrolling the k-tile loop by 2 or 4 (4.49 / 4.48 ms).
// (That reading of the dispatch was wrong -- see the per-CTA trace below: one tile per CTA runs in lockstep too.)
// The kernel as it stands is a fragile local optimum of ptxas' register allocation; the next step needs the main loop
// pinned down (hand-scheduled or CUTLASS-style explicit pipelining) before restructuring around it.
// Round 2 (tools/microbench/fp64_lds_mix.py, profiles/r02_fp64_lds_mix.txt): the SASS of the full k-tile body is already at
// the pipe's pace -- every DMMA carries a 15-cycle stall + NOP, i.e. one m8n8k4 per 16 cycles per warp, and a register-fed
// DMMA stream with this loop's 12 LDS.64 per 32 DMMAs measures 37.06 TFLOP/s against 37.07 from registers (24 LDS.64: 36.0),
// so shared-memory loads do not compete with the tensor pipe.  DFMA does (DMMA + DFMA interleaved: 36.1 TFLOP/s together),
// which is why the operand generator's 8 DMULs per thread and k-tile show up as `math` stalls.  What is left is per-CTA: a
// 32-column / 4-warp / 3-stage variant with THREE independent CTAs per SM (160-168 registers, template parameters NST / NAH /
// MINB) is SLOWER (4.71 vs 4.38 ms per step at 65 536 columns; 0.72 vs 0.60 ms at 8 192) -- twice the CTAs means twice the
// prologues, epilogues and A-tile fetches, and that costs more than a third warp per sub-partition recovers.
//
// Per-CTA clock trace (PRB_TRACE_CTAS build + tools/debug/cta_trace.py, profiles/r02_cta_trace.txt): with the plain
// heavy-first tile order the two CTAs of an SM -- and all SMs -- run in LOCKSTEP for the whole launch (equal-length tiles
// that start together share the pipe evenly, end together, and so do their successors): both are outside their k loop at
// the same time for 6.1 % (lower) / 7.6 % (upper) of the launch, against 1.2 % / 4.0 % with exactly one inside, and inside
// the loop the pair keeps the pipe 96 % busy.  A lone 4-warp CTA, however, only reaches ~62 % of the pipe (one CTA per SM:
// 6.20 vs 4.33 ms per step), so breaking the lockstep recovers less than the idle share: the MIXED tile order (heaviest and
// lightest row tile of a column tile alternate in the queue; last scheduling group stays heavy-first for the tail) takes
// the both-outside share to 1.4 % / 3.3 % and the step from 4.370 to 4.326 ms; mixing the last group as well costs more in
// the tail (4.399 ms).  A start-up stagger of the second CTA on each SM (0.5-2 tile units) decays within a few tiles
// (4.353 ms) and adds nothing on top of the mixed order.  Generating the operand by a GATHER from a per-restart table
// kc[b][j] * tab[h] (no DMUL, no table LDS in the consumer warps) changes nothing (4.370 vs 4.373 ms): the `math` stalls of
// the generator are time the pipe spends on the sibling CTA's DMMAs, not lost time.  Two 8-warp CTAs per SM with 32 x 32 warp
// tiles (128 registers, four warps per sub-partition, so that a lone CTA still has two) are slower: 4.553 vs 4.328 ms.
// What DID pay after the trace: the elected producer thread's issue work.  One 3-D TMA box per operand and k-tile instead
// of four 2-D boxes, and -- where both operands arrive by TMA -- the four M-warps taking turns as producer: upper product
// 2.111 -> 2.024 ms, generic-path step 4.98 -> 4.77 ms, bench step 4.326 -> 4.247 ms.  In LOWER_GEN (every thread generates)
// the rotation is slower (2.201 -> 2.226 ms, 2.208 -> 2.214 with 3-D boxes), and so is one full-barrier arrival per warp
// after a __syncwarp instead of one per thread (2.208 -> 2.255 ms).
// Evaluating the acquisition function of a CTA's own columns inside the UPPER_GRAD prologue (while the first TMA tiles
// are in flight; removes the separate epilogue launch from the fused step) saves 3 us at 8 192 columns but the added
// code costs the 65 536-column step 0.03-0.04 ms even when it is switched off at run time (4.247 -> 4.28-4.29 ms): reverted.
// LOWER_GEN against LOWER_TMA (2.208 vs 2.036 ms), by timing-only variants: the popc -> table LDS -> DMUL chain 0.08 ms, the
// generator's global loads 0.03 ms, the per-thread barrier protocol 0.06 ms, the stores nothing.  Forming the operand values
// in registers in the MIDDLE of the current k-tile's DMMA block (gen_compute at chunk 3; the in-order warp no longer sits
// through the chain after its last DMMA) took the lower product to 2.148 ms and the step to 4.190 ms.  Also tried, slower:
// the same at chunks 1 / 2 (2.154 / 2.158), storing inside the block as well (2.155), the generator's loads inside the
// block (2.158).  st.async stores that signal their bytes on the full barrier (no per-thread arrive) raise an illegal
// instruction outside a cluster launch; a non-blocking test of the stage's empty barrier inside the block (blocking wait
// only if it failed) is slower too (2.195 ms): a volatile barrier instruction in the block disturbs the DMMA schedule.
// UPPER_GRAD epilogue (9-12 k cycles of a ~140 k-cycle CTA: shuffle trees + two CTA barriers per gradient dim, ~70 global
// loads per thread after the k loop): partial column sums of ALL dims through the free stage buffers behind one barrier
// (upper product 2.023 -> 2.004 ms), and the epilogue's operands fetched into shared memory in the PROLOGUE, while the first
// TMA tiles are in flight (-> 1.986 ms; bench step 4.148 ms).
//
// Four modes (template):
//   LOWER_GEN   the B operand k*[col, k] = kc[b(col), k] * tab[popc(z[col] ^ Z[k])] is GENERATED into shared memory by the
//               consumer threads (separable "mixed" kernel: Matern over continuous dims x isotropic Matern over binary dims);
//               the k* panel never exists in HBM.
//   LOWER_TMA   B operand = materialised k* panel (generic kernel structures), via TMA.
//   UPPER_W     stores W (generic path; the gradient contraction kernel consumes it).
//   UPPER_GRAD  epilogue contracts W with the kernel derivative in registers:
//               S[c, col] = sum_j (dA/dmu alpha_j - 2 dA/dvar W[j, col]) * kb(z_col, Z_j) * Dc[b(col), c, j]; W never hits HBM.
//   UPPER_GRADQ the same fusion for kernel structures that do not separate (ordinals inside the ARD Matern, categorical
//               kernels): T[j, col] = (dA/dmu alpha_j - 2 dA/dvar W[j, col]) * Q[col, j] with the derivative-factor panel Q
//               written by kstar_q_kernel next to k*, then  S[dim, col] = w_dim * (x_dim[col] * sum_j T - sum_j T X[j, dim])
//               for the dims of the gradient group.  Replaces upper(W panel) + grad_contract_kernel.
#include <cuda.h>

#include <cstdlib>

#include "prb_device.cuh"
#include "prb_ptx.cuh"

namespace prb {

#ifndef PRB_T_STAGES
#define PRB_T_STAGES 4
#endif
#ifndef PRB_T_AHEAD
#define PRB_T_AHEAD 2
#endif
constexpr int T_BM = 128, T_BN = 128, T_BK = 16, T_STAGES = PRB_T_STAGES, T_AHEAD = PRB_T_AHEAD, T_THREADS = 256;
static_assert(T_AHEAD >= 1 && T_AHEAD <= T_STAGES - 1, "the producer may run at most T_STAGES - 1 k-tiles ahead");
constexpr int T_OPER_DOUBLES = T_BM * T_BK;            // 2048 doubles = 16 KB per operand per stage
constexpr int T_CHUNK_DOUBLES = T_BM * 4;              // one k-chunk (4 k values) of 128 rows
constexpr size_t tma_smem_bytes(int bn, int stages = T_STAGES) {
  return size_t(stages) * (T_OPER_DOUBLES + bn * T_BK) * 8 + 2 * stages * 8 + 64 * 8 + (3 * 128 + 6 * 128) * 8;
}
constexpr size_t T_SMEM_BYTES = tma_smem_bytes(128);

enum { MODE_LOWER_GEN = 0, MODE_LOWER_TMA = 1, MODE_UPPER_W = 2, MODE_UPPER_GRAD = 3, MODE_UPPER_GRADQ = 4 };

// PRB_TRACE_CTAS=1 builds record, per CTA, {smid, clock64 at entry, after the first full-barrier wait, after the k loop, at
// exit} into the buffer set with prb_debug_set_cta_trace (tools/debug/cta_trace.py); product builds carry none of it.
#ifndef PRB_TRACE_CTAS
#define PRB_TRACE_CTAS 0
#endif
#if PRB_TRACE_CTAS
static long long* g_cta_trace = nullptr;
static int g_cta_trace_launch = 0;
constexpr int CTA_TRACE_STRIDE = 8 * 16384;
extern "C" void prb_debug_set_cta_trace(long long* buf) {
  g_cta_trace = buf;
  g_cta_trace_launch = 0;
}
#define CTA_TRACE(slot)                                                        \
  do {                                                                         \
    if (p.trace != nullptr && tid == 0) p.trace[size_t(blockIdx.x) * 8 + (slot)] = clock64(); \
  } while (0)
#else
#define CTA_TRACE(slot) \
  do {                  \
  } while (0)
#endif

#ifndef PRB_GEN_ROTATE
#define PRB_GEN_ROTATE 0
#endif
#ifndef PRB_UPPER_EPI_SMEM
#define PRB_UPPER_EPI_SMEM 1
#endif
#ifndef PRB_UPPER_EPI_PRELOAD
#define PRB_UPPER_EPI_PRELOAD 1
#endif
struct TmaGemmParams {
  long long* trace = nullptr;
  int mixed_groups = 0;    // leading scheduling groups that use the mixed (heavy / light alternating) order
  int mtiles, ntiles, Kp, ncols_pad;
  int ngroup;  // column tiles per scheduling group (see the launch order in the kernel)
  // lower epilogue
  double* C;  // V^T [ncols_pad, ldc] (lower) or W [Mp2, ncols_pad] (UPPER_W)
  int ldc;
  int special_row;  // row of A1 holding alpha^T (its product is the posterior-mean dot), -1 for upper
  double* norm_part;
  double* mu_dot;
  // operand generation / gradient epilogue (separable kernel)
  const double* kc;       // [nb, Kp]  outputscale * Matern52 over the continuous dims, per restart (0 beyond n)
  const uint32_t* Zbits;  // [Mp]      bit codes of the training points over the binary dims (0 beyond n)
  const uint32_t* zbits;  // [ncols_pad] bit codes of this panel's evaluation points
  const double* tab;      // [ntab + 1] binary-factor value by Hamming distance
  int ntab;
  int mc;        // evaluation points per restart (b(col) = (col0 + col) / mc)
  int64_t col0;  // global index of the panel's first column
  int nb;        // number of restarts (clamp for padded columns)
  const int* colb;       // de-duplicated columns: restart of each column (panel-local), else NULL
  const int* ncols_dev;  // de-duplicated columns: device scalar with the total number of columns, else NULL
  const double* Dc;         // [nb, n_cont, Kp] outputscale * G(r) * (x_c - X_jc) / l_c^2
  int n_cont;
  const double* alpha_pad;  // [Mp2] alpha, 0 beyond n
  const double* dmu;        // [ncols] dA/dmu       (panel-local)
  const double* dvar;       // [ncols] dA/dsigma^2
  int ncols;                // valid columns in this panel
  double* S_part;           // [mtiles, n_cont, ncols_pad]
  // LOWER modes with a single row tile (n + 1 <= 128): the acquisition epilogue is fused (||v||^2 and mu are complete here)
  int fuse_acq;
  DevAcq aq;
  double mean_const, kss;
  double* acq_a;     // [ncols] acquisition value
  double* acq_dmu;   // [ncols] dA/dmu
  double* acq_dvar;  // [ncols] dA/dsigma^2
  int acq_ncols;
  // UPPER_GRADQ: gradient group, derivative-factor panel, kernel inputs of the panel's columns, training inputs
  GradGroup gq;
  const double* Qp;   // [panel rows = columns][Kp]
  const double* xk;   // [ncols, d_k]
  const double* Xtr;  // [n, d_k]
  int n, d_k;
};

// BN = evaluation columns per tile: 128 for throughput; 32 for small problems, where a 128-wide tiling would leave most
// of the 148 SMs idle and the step is bound by ONE SM's DMMA rate (n = 119, 640 columns: 5 tiles, 23 us per product).
template <int MODE, int BN, int THREADS = T_THREADS, int NST = T_STAGES, int NAH = T_AHEAD, int MINB = (THREADS == 128 ? 2 : 1)>
__global__ void __launch_bounds__(THREADS, MINB)
    trgemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const TmaGemmParams p) {
  constexpr bool UPPER = (MODE == MODE_UPPER_W || MODE == MODE_UPPER_GRAD || MODE == MODE_UPPER_GRADQ);
  constexpr bool GEN = (MODE == MODE_LOWER_GEN);
  constexpr int NWN = THREADS / 128;   // warps along N (4 along M)
  constexpr int TJ = BN / (8 * NWN);   // 8 x 8 accumulator tiles per warp along N (warp = 32 rows x 8 TJ columns)
  constexpr int GEN_VPT = BN * T_BK / THREADS;  // generated operand values per thread and k-tile (8 or 2)
  static_assert(GEN_VPT == 8 || GEN_VPT == 4 || GEN_VPT == 2, "operand generator mappings");
  constexpr int B_CHUNK = BN * 4;      // doubles per k-chunk of the B operand
  constexpr int B_OPER = BN * T_BK;    // doubles per k-tile of the B operand
  constexpr int STAGE_DOUBLES = T_OPER_DOUBLES + B_OPER;  // A tile + B tile of one k-tile
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stages = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + size_t(NST) * STAGE_DOUBLES * 8);
  uint64_t* empty = full + NST;
  double* tab_s = reinterpret_cast<double*>(empty + NST);  // [64]
  // UPPER_GRAD: the epilogue's operands, fetched in the prologue while the first TMA tiles are in flight
  double* e_cm = tab_s + 64;                            // [BN] dA/dmu of the tile's columns (0 beyond the valid ones)
  double* e_cv = e_cm + 128;                            // [BN] -2 dA/dsigma^2
  uint32_t* e_zc = reinterpret_cast<uint32_t*>(e_cv + 128);  // [BN] bit codes of the columns
  double* e_al = e_cv + 256;                            // [128] alpha of the tile's rows
  uint32_t* e_Zr = reinterpret_cast<uint32_t*>(e_al + 128);  // [128] bit codes of the rows
  double* e_D = e_al + 256;                             // [4][128] derivative rows of the first four dims (one restart per tile)

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int wm = warp & 3;   // 4 warps along M (32 rows each)
  const int wn = warp >> 2;  // 2 warps along N (64 cols each)

  // Launch order: column tiles are taken in groups of p.ngroup; inside a group every row tile of those columns is scheduled
  // back to back (heavy first: lower -> largest row block first, upper -> smallest row block first), so the B-operand
  // columns that the row tiles of one column tile re-read (the v panel in the upper product, the k* panel on the generic
  // path) are still in L2 instead of coming from HBM once per row tile.  p.ngroup = ntiles: plain row-tile-major order.
  CTA_TRACE(1);
#if PRB_TRACE_CTAS
  if (p.trace != nullptr && tid == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    p.trace[size_t(blockIdx.x) * 8] = smid;
  }
#endif
  const int t = blockIdx.x;
  const int per_group = p.ngroup * p.mtiles;
  const int grp_i = t / per_group;
  const int r_in = t - grp_i * per_group;
  const int gcols = min(p.ngroup, p.ntiles - grp_i * p.ngroup);  // column tiles in this (possibly last, shorter) group
  int mi_sched = r_in / gcols;
  int nt = grp_i * p.ngroup + (r_in - mi_sched * gcols);
  if (grp_i < p.mixed_groups) {
    // Mixed order (all groups but the last): heaviest and lightest remaining row tile of a column tile alternate in the
    // queue, so the CTAs that share an SM -- and the SMs among each other -- differ in length and drift out of phase.  In
    // the plain heavy-first order whole waves have equal length: every CTA pair starts, stores and restarts in lockstep
    // (tools/debug/cta_trace.py), the FP64 pipe idles through both epilogues and the stores of all SMs hit L2 at once.
    const int pair = r_in / (2 * gcols);
    if (pair < p.mtiles / 2) {
      const int rem = r_in - pair * 2 * gcols;
      mi_sched = (rem & 1) ? p.mtiles - 1 - pair : pair;
      nt = grp_i * p.ngroup + (rem >> 1);
    } else {  // odd number of row tiles: the middle one comes last
      mi_sched = p.mtiles / 2;
      nt = grp_i * p.ngroup + (r_in - (p.mtiles / 2) * 2 * gcols);
    }
  }
  const int mt = UPPER ? mi_sched : (p.mtiles - 1 - mi_sched);
  const int m0 = mt * T_BM;
  const int n0 = nt * BN;
  int kbeg, kend;
  if (UPPER) {
    kbeg = m0;
    kend = p.Kp;
  } else {
    kbeg = 0;
    kend = (mt == p.mtiles - 1) ? p.Kp : min(p.Kp, m0 + T_BM);
  }
  const int num_k = (kend - kbeg) / T_BK;

  grid_dep_launch();  // PDL: let the next kernel of the step start its prologue (no-op for ordinary launches)
  // de-duplicated columns: the grid covers the worst case (no duplicates); tiles beyond the unique columns do nothing
  int ncols_valid = p.ncols;
  if (p.ncols_dev != nullptr) {
    grid_dep_wait();
    const int64_t left = int64_t(*p.ncols_dev) - p.col0;
    if (left <= n0) return;
    if (left < ncols_valid || ncols_valid == 0) ncols_valid = int(left);
  }

  if (tid == 0) {
    tma_prefetch_desc(&tmA);
    if (!GEN) tma_prefetch_desc(&tmB);
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full[s], GEN ? 1 + THREADS : 1);
      mbar_init(&empty[s], THREADS / 32);
    }
    mbar_fence_init();
  }
  if ((GEN || MODE == MODE_UPPER_GRAD) && tid < 64) tab_s[tid] = (tid <= p.ntab) ? p.tab[tid] : 0.0;
  __syncthreads();
  grid_dep_wait();  // PDL: everything above is independent of the predecessor kernel; its outputs are read from here on

  // ---- B-operand generation: thread <-> (evaluation column, 8 of the 16 k values of a k-tile) ---------------------------
  // (the 32-byte column stride of the STS.128 below costs 2-way bank conflicts, 21 M per launch; a lane-pair-per-column
  //  mapping that removes them needs 8 instead of 6 global loads per thread and k-tile and measured 1.7 % SLOWER)
  // BN = 32: thread <-> (column, one pair of the 16 k values)
  const int gcol = tid & (BN - 1);
  const int ghalf = tid / BN;
  uint32_t g_zb = 0;
  const double* g_kc = nullptr;
  if (GEN) {
    g_zb = p.zbits[n0 + gcol];
    int64_t br = p.colb ? int64_t(p.colb[n0 + gcol]) : (p.col0 + n0 + gcol) / p.mc;
    if (br > p.nb - 1) br = p.nb - 1;
    g_kc = p.kc + size_t(br) * p.Kp;
  }

  // TMA side of a k-tile (one elected thread): A always, B unless it is generated
  auto produce_tma = [&](int kt) {
    // A + B by TMA: the four M-warps take turns as producer, which spreads the elected thread's issue work over the warps
    // (upper product 2.111 -> 2.075 ms).  With the B operand generated by all threads the rotation is SLOWER (lower product
    // 2.197 -> 2.226 ms), so there warp 0 stays the producer.
    if (tid != ((GEN && !PRB_GEN_ROTATE) ? 0 : 32 * (kt & 3))) return;
    const int s = kt % NST;
    const int use = kt / NST;
    const int k0 = kbeg + kt * T_BK;
    double* As = stages + s * STAGE_DOUBLES;
    double* Bs = As + T_OPER_DOUBLES;
    if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
    mbar_arrive_expect_tx(&full[s], (T_OPER_DOUBLES + (GEN ? 0 : B_OPER)) * 8);
    // one 3-D box (4 k values x rows x 4 k-chunks) per operand lands as the [k-chunk][row][4] stage layout
    tma_load_3d(As, &tmA, &full[s], 0, m0, k0 >> 2);
    if (!GEN) tma_load_3d(Bs, &tmB, &full[s], 0, n0, k0 >> 2);
  };
  // generated B operand: global loads are issued a whole k-tile of DMMA work before they are consumed
  double2 g_k01, g_k23, g_k45, g_k67;
  uint4 g_za, g_zc;
  uint2 g_zs;
  auto gen_load = [&](int kt) {
    const int k0 = kbeg + kt * T_BK;
    if constexpr (GEN_VPT == 8) {
      const double2* kc2 = reinterpret_cast<const double2*>(g_kc + k0 + ghalf * 8);
      const uint4* zz = reinterpret_cast<const uint4*>(p.Zbits + k0 + ghalf * 8);
      g_k01 = __ldg(kc2);
      g_k23 = __ldg(kc2 + 1);
      g_k45 = __ldg(kc2 + 2);
      g_k67 = __ldg(kc2 + 3);
      g_za = __ldg(zz);
      g_zc = __ldg(zz + 1);
    } else if constexpr (GEN_VPT == 4) {
      const double2* kc2 = reinterpret_cast<const double2*>(g_kc + k0 + ghalf * 4);
      g_k01 = __ldg(kc2);
      g_k23 = __ldg(kc2 + 1);
      g_za = __ldg(reinterpret_cast<const uint4*>(p.Zbits + k0 + ghalf * 4));
    } else {
      g_k01 = __ldg(reinterpret_cast<const double2*>(g_kc + k0 + ghalf * 2));
      g_zs = __ldg(reinterpret_cast<const uint2*>(p.Zbits + k0 + ghalf * 2));
    }
  };
  // GEN_VPT == 8: the operand values of the NEXT generated k-tile are formed in registers in the middle of the current
  // k-tile's DMMA block (gen_compute: popc -> table LDS -> DMUL, in the same basic block as the DMMAs, so their latencies
  // hide behind the warp's own 16-cycle DMMA cadence) and only stored at the end of the k-tile (gen_store)
  double2 g_o0, g_o1, g_o2, g_o3;
  auto gen_compute = [&]() {
    if constexpr (GEN_VPT == 8) {
      g_o0.x = g_k01.x * tab_s[__popc(g_zb ^ g_za.x)];
      g_o0.y = g_k01.y * tab_s[__popc(g_zb ^ g_za.y)];
      g_o1.x = g_k23.x * tab_s[__popc(g_zb ^ g_za.z)];
      g_o1.y = g_k23.y * tab_s[__popc(g_zb ^ g_za.w)];
      g_o2.x = g_k45.x * tab_s[__popc(g_zb ^ g_zc.x)];
      g_o2.y = g_k45.y * tab_s[__popc(g_zb ^ g_zc.y)];
      g_o3.x = g_k67.x * tab_s[__popc(g_zb ^ g_zc.z)];
      g_o3.y = g_k67.y * tab_s[__popc(g_zb ^ g_zc.w)];
    } else if constexpr (GEN_VPT == 2) {
      g_o0.x = g_k01.x * tab_s[__popc(g_zb ^ g_zs.x)];
      g_o0.y = g_k01.y * tab_s[__popc(g_zb ^ g_zs.y)];
    }
  };
  auto gen_store = [&](int kt) {
    const int s = kt % NST;
    const int use = kt / NST;
    double* Bs = stages + s * STAGE_DOUBLES + T_OPER_DOUBLES;
    if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
    if constexpr (GEN_VPT == 8) {
      const double2 o0 = g_o0, o1 = g_o1, o2 = g_o2, o3 = g_o3;
      double2* d0 = reinterpret_cast<double2*>(Bs + (2 * ghalf) * B_CHUNK + gcol * 4);
      double2* d1 = reinterpret_cast<double2*>(Bs + (2 * ghalf + 1) * B_CHUNK + gcol * 4);
      d0[0] = o0;
      d0[1] = o1;
      d1[0] = o2;
      d1[1] = o3;
    } else if constexpr (GEN_VPT == 4) {  // thread <-> (column, one k-chunk of the k-tile)
      double2 o0, o1;
      o0.x = g_k01.x * tab_s[__popc(g_zb ^ g_za.x)];
      o0.y = g_k01.y * tab_s[__popc(g_zb ^ g_za.y)];
      o1.x = g_k23.x * tab_s[__popc(g_zb ^ g_za.z)];
      o1.y = g_k23.y * tab_s[__popc(g_zb ^ g_za.w)];
      double2* d0 = reinterpret_cast<double2*>(Bs + ghalf * B_CHUNK + gcol * 4);
      d0[0] = o0;
      d0[1] = o1;
    } else {
      const double2 o = g_o0;
      *reinterpret_cast<double2*>(Bs + (ghalf >> 1) * B_CHUNK + gcol * 4 + (ghalf & 1) * 2) = o;
    }
    mbar_arrive(&full[s]);
  };

  double acc[4][TJ][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int kt = 0; kt < NAH && kt < num_k; ++kt) {
    produce_tma(kt);
    if (GEN) {
      gen_load(kt);
      gen_compute();
      gen_store(kt);
    }
  }

#if PRB_UPPER_EPI_PRELOAD
  bool e_single = false;
  int64_t e_blo = 0;
  if constexpr (MODE == MODE_UPPER_GRAD) {
    // restart index of the tile's first / last column: one restart per tile is the common case (mc % 128 == 0)
    int64_t b_hi = p.colb ? int64_t(p.colb[n0 + BN - 1]) : (p.col0 + n0 + BN - 1) / p.mc;
    e_blo = p.colb ? int64_t(p.colb[n0]) : (p.col0 + n0) / p.mc;
    if (e_blo > p.nb - 1) e_blo = p.nb - 1;
    if (b_hi > p.nb - 1) b_hi = p.nb - 1;
    e_single = (e_blo == b_hi);
    for (int t2 = tid; t2 < BN; t2 += THREADS) {
      const int col = n0 + t2;
      const bool ok = col < ncols_valid;
      e_cm[t2] = ok ? p.dmu[col] : 0.0;
      e_cv[t2] = ok ? -2.0 * p.dvar[col] : 0.0;
      e_zc[t2] = p.zbits[col];
    }
    for (int r = tid; r < T_BM; r += THREADS) {
      e_al[r] = p.alpha_pad[m0 + r];
      e_Zr[r] = p.Zbits[m0 + r];
    }
    if (e_single) {
      const int nd = p.n_cont < 4 ? p.n_cont : 4;
      for (int idx = tid; idx < nd * T_BM; idx += THREADS) {
        const int c = idx / T_BM, r = idx - c * T_BM;
        e_D[idx] = (m0 + r < p.Kp) ? p.Dc[(size_t(e_blo) * p.n_cont + c) * p.Kp + m0 + r] : 0.0;
      }
    }
  }
#endif

  // Row ownership: M-warp wm owns the 8-row groups {wm, 7 - wm, 8 + wm, 15 - wm} of the tile, so that the triangular work
  // inside a diagonal block (where whole 8 x 4 sub-blocks of A are zero and their DMMAs are skipped) is split evenly.
  const int frag_r = lane >> 2;
  const int frag_k = lane & 3;
  int grp[4];
  grp[0] = wm;
  grp[1] = 7 - wm;
  grp[2] = 8 + wm;
  grp[3] = 15 - wm;
  int fa_off[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) fa_off[i] = (8 * grp[i] + frag_r) * 4 + frag_k;
  const int fb_off = (wn * (8 * TJ) + frag_r) * 4 + frag_k;

  for (int kt = 0; kt < num_k; ++kt) {
    const int nk = kt + NAH;
    if (nk < num_k) {
      produce_tma(nk);
      if (GEN) gen_load(nk);
    }
    const int s = kt % NST;
    mbar_wait(&full[s], (kt / NST) & 1);
#if PRB_TRACE_CTAS
    if (kt == 0) CTA_TRACE(2);
#endif
    const double* pa = stages + s * STAGE_DOUBLES;
    const double* pb = pa + T_OPER_DOUBLES + fb_off;
    const int q0 = kbeg + kt * T_BK - m0;  // offset of this k-tile inside the tile's diagonal block
    if (q0 < 0 || q0 >= T_BM) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        double a[4], b[TJ];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = pa[c * T_CHUNK_DOUBLES + fa_off[i]];
#pragma unroll
        for (int j = 0; j < TJ; ++j) b[j] = pb[c * B_CHUNK + j * 32];
        if (GEN && c == 3 && nk < num_k) gen_compute();  // the loads of gen_load(nk) were issued three chunks ago
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < TJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    } else {
      bool gen_done = false;
      // diagonal block: sub-block (8 rows of group g) x (4 k values at offset q) of A is all zero when
      //   lower (k <= row):  q > 8 g + 7          upper (k >= row):  q + 3 < 8 g
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int q = q0 + 4 * c;
        bool act[4];
        bool any = false;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          act[i] = UPPER ? (q + 3 >= 8 * grp[i]) : (q <= 8 * grp[i] + 7);
          any = any || act[i];
        }
        if (!any) continue;
        double b[TJ];
#pragma unroll
        for (int j = 0; j < TJ; ++j) b[j] = pb[c * B_CHUNK + j * 32];
        if (GEN && c == 3 && nk < num_k) {  // as in the full-tile body; chunk 3 of a generating diagonal tile has active
          gen_compute();                     // rows for every warp (q <= 92 there), the flag covers the general case
          gen_done = true;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (act[i]) {
            const double a = pa[c * T_CHUNK_DOUBLES + fa_off[i]];
#pragma unroll
            for (int j = 0; j < TJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a, b[j]);
          }
        }
      }
      if (GEN && nk < num_k && !gen_done) gen_compute();
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    if (GEN && nk < num_k) gen_store(nk);
  }
  __syncthreads();  // every stage consumed, no TMA in flight: the stage buffers are free for the epilogue
  CTA_TRACE(3);

  // ---- epilogues ---------------------------------------------------------------------------------------------------
  int rows[4];  // global row of accumulator sub-tile i for this lane
#pragma unroll
  for (int i = 0; i < 4; ++i) rows[i] = m0 + 8 * grp[i] + frag_r;
  const int col_base = n0 + wn * (8 * TJ) + 2 * frag_k;  // + 8 * j + e
  double* red = stages;                            // [4][128] scratch

  if constexpr (MODE == MODE_UPPER_W) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double* crow = p.C + size_t(rows[i]) * p.ldc + col_base;
#pragma unroll
      for (int j = 0; j < TJ; ++j) *reinterpret_cast<double2*>(crow + 8 * j) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
    CTA_TRACE(4);
    return;
  }

  if constexpr (MODE == MODE_UPPER_GRAD) {
    // T[row, col] = (dA/dmu alpha_row - 2 dA/dvar W[row, col]) * kb(z_col, Z_row), in place of the accumulators
    double al[4];
    uint32_t Zr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#if PRB_UPPER_EPI_PRELOAD
      al[i] = e_al[rows[i] - m0];
      Zr[i] = e_Zr[rows[i] - m0];
#else
      al[i] = p.alpha_pad[rows[i]];
      Zr[i] = p.Zbits[rows[i]];
#endif
    }
#pragma unroll
    for (int j = 0; j < TJ; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = col_base + 8 * j + e;
#if PRB_UPPER_EPI_PRELOAD
        const double cm = e_cm[col - n0];
        const double cv = e_cv[col - n0];
        const uint32_t zc = e_zc[col - n0];
#else
        const bool cv_ok = col < ncols_valid;
        const double cm = cv_ok ? p.dmu[col] : 0.0;
        const double cv = cv_ok ? -2.0 * p.dvar[col] : 0.0;
        const uint32_t zc = p.zbits[col];
#endif
#pragma unroll
        for (int i = 0; i < 4; ++i)
          acc[i][j][e] = fma(cv, acc[i][j][e], cm * al[i]) * tab_s[__popc(zc ^ Zr[i])];
      }
    // restart index of the tile's first / last column: one restart per tile is the common case (mc % 128 == 0)
#if PRB_UPPER_EPI_PRELOAD
    const int64_t b_lo = e_blo;
    const bool single = e_single;
#else
    int64_t b_lo = p.colb ? int64_t(p.colb[n0]) : (p.col0 + n0) / p.mc;
    int64_t b_hi = p.colb ? int64_t(p.colb[n0 + BN - 1]) : (p.col0 + n0 + BN - 1) / p.mc;
    if (b_lo > p.nb - 1) b_lo = p.nb - 1;
    if (b_hi > p.nb - 1) b_hi = p.nb - 1;
    const bool single = (b_lo == b_hi);
#endif
#if PRB_UPPER_EPI_SMEM
    // Column sums of T * D over the tile's 128 rows, all dims of a batch behind ONE CTA barrier: every thread stores its
    // partial sums over its four row groups as [dim][wm * 8 + frag_r][column] (row pitch BN + 8 doubles: the STS.128 of a
    // warp fall on disjoint bank groups), then one thread per (dim, column) adds the 32 partials in a fixed order.  The
    // shuffle tree + two barriers per dim this replaces made the epilogue 9-12 k cycles per CTA.
    constexpr int EP = BN + 8;
    constexpr int ECAP = (NST * STAGE_DOUBLES) / (32 * EP);  // dims per batch that fit the (now free) stage buffers
    double* part_s = stages;
    const int slot = wm * 8 + frag_r;
    for (int c0 = 0; c0 < p.n_cont; c0 += ECAP) {
      const int nc = min(ECAP, p.n_cont - c0);
      for (int cc = 0; cc < nc; ++cc) {
        const int c = c0 + cc;
        double part[TJ][2];
        if (single) {
          const double* D = p.Dc + (size_t(b_lo) * p.n_cont + c) * p.Kp;
          double dr[4];
#if PRB_UPPER_EPI_PRELOAD
          if (c < 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) dr[i] = e_D[c * T_BM + rows[i] - m0];
          } else
#endif
          {
#pragma unroll
            for (int i = 0; i < 4; ++i) dr[i] = (rows[i] < p.Kp) ? D[rows[i]] : 0.0;
          }
#pragma unroll
          for (int j = 0; j < TJ; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              double sacc = 0.0;
#pragma unroll
              for (int i = 0; i < 4; ++i) sacc = fma(acc[i][j][e], dr[i], sacc);
              part[j][e] = sacc;
            }
        } else {
#pragma unroll
          for (int j = 0; j < TJ; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              int64_t br = p.colb ? int64_t(p.colb[col_base + 8 * j + e]) : (p.col0 + col_base + 8 * j + e) / p.mc;
              if (br > p.nb - 1) br = p.nb - 1;
              const double* D = p.Dc + (size_t(br) * p.n_cont + c) * p.Kp;
              double sacc = 0.0;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int row = rows[i];
                sacc = fma(acc[i][j][e], row < p.Kp ? D[row] : 0.0, sacc);
              }
              part[j][e] = sacc;
            }
        }
        double* prow = part_s + (size_t(cc) * 32 + slot) * EP + wn * (8 * TJ) + 2 * frag_k;
#pragma unroll
        for (int j = 0; j < TJ; ++j) *reinterpret_cast<double2*>(prow + 8 * j) = make_double2(part[j][0], part[j][1]);
      }
      __syncthreads();
      for (int idx = tid; idx < nc * BN; idx += THREADS) {
        const int cc = idx / BN, col = idx - cc * BN;
        const double* src = part_s + size_t(cc) * 32 * EP + col;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int r = 0; r < 32; r += 4) {
          s0 += src[(r + 0) * EP];
          s1 += src[(r + 1) * EP];
          s2 += src[(r + 2) * EP];
          s3 += src[(r + 3) * EP];
        }
        p.S_part[(size_t(mt) * p.n_cont + c0 + cc) * p.ncols_pad + n0 + col] = (s0 + s1) + (s2 + s3);
      }
      if (c0 + ECAP < p.n_cont) __syncthreads();
    }
#else
    for (int c = 0; c < p.n_cont; ++c) {
      double part[TJ][2];
      if (single) {
        const double* D = p.Dc + (size_t(b_lo) * p.n_cont + c) * p.Kp;
        double dr[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) dr[i] = (rows[i] < p.Kp) ? D[rows[i]] : 0.0;
#pragma unroll
        for (int j = 0; j < TJ; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) s = fma(acc[i][j][e], dr[i], s);
            part[j][e] = s;
          }
      } else {
#pragma unroll
        for (int j = 0; j < TJ; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            int64_t br = p.colb ? int64_t(p.colb[col_base + 8 * j + e]) : (p.col0 + col_base + 8 * j + e) / p.mc;
            if (br > p.nb - 1) br = p.nb - 1;
            const double* D = p.Dc + (size_t(br) * p.n_cont + c) * p.Kp;
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int row = rows[i];
              s = fma(acc[i][j][e], row < p.Kp ? D[row] : 0.0, s);
            }
            part[j][e] = s;
          }
      }
      // sum over the 8 row-lanes (lane bits 2..4), then over the 4 M-warps, in a fixed order
#pragma unroll
      for (int j = 0; j < TJ; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          double v = part[j][e];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          part[j][e] = v;
        }
      if (frag_r == 0) {
#pragma unroll
        for (int j = 0; j < TJ; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) red[wm * BN + wn * (8 * TJ) + 8 * j + 2 * frag_k + e] = part[j][e];
      }
      __syncthreads();
      if (tid < BN) {
        const double s = ((red[tid] + red[BN + tid]) + red[2 * BN + tid]) + red[3 * BN + tid];
        p.S_part[(size_t(mt) * p.n_cont + c) * p.ncols_pad + n0 + tid] = s;
      }
      __syncthreads();
    }
#endif
    CTA_TRACE(4);
    return;
  }

  if constexpr (MODE == MODE_UPPER_GRADQ) {
    // T[row, col] = (dA/dmu alpha_row - 2 dA/dvar W[row, col]) * Q[col, row], in place of the accumulators
    double al[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) al[i] = p.alpha_pad[rows[i]];
#pragma unroll
    for (int j = 0; j < TJ; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = col_base + 8 * j + e;
        const bool cv_ok = col < ncols_valid;
        const double cm = cv_ok ? p.dmu[col] : 0.0;
        const double cv = cv_ok ? -2.0 * p.dvar[col] : 0.0;
        const double* qrow = p.Qp + size_t(col) * p.Kp;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const double q = (cv_ok && rows[i] < p.Kp) ? __ldg(qrow + rows[i]) : 0.0;
          acc[i][j][e] = fma(cv, acc[i][j][e], cm * al[i]) * q;
        }
      }
    double* s0buf = red + 4 * BN;  // [BN] sum_j T[j, col] of this row tile
    for (int s = 0; s <= p.gq.n_dims; ++s) {
      // slot 0: plain sum; slot s >= 1: weighted by the training inputs' column dims[s - 1]
      double xr[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        xr[i] = (s == 0) ? 1.0 : (rows[i] < p.n ? __ldg(p.Xtr + size_t(rows[i]) * p.d_k + p.gq.dims[s - 1]) : 0.0);
      double part[TJ][2];
#pragma unroll
      for (int j = 0; j < TJ; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          double v = 0.0;
#pragma unroll
          for (int i = 0; i < 4; ++i) v = fma(acc[i][j][e], xr[i], v);
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          part[j][e] = v;
        }
      if (frag_r == 0) {
#pragma unroll
        for (int j = 0; j < TJ; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) red[wm * BN + wn * (8 * TJ) + 8 * j + 2 * frag_k + e] = part[j][e];
      }
      __syncthreads();
      if (tid < BN) {
        const double sum = ((red[tid] + red[BN + tid]) + red[2 * BN + tid]) + red[3 * BN + tid];
        if (s == 0) {
          s0buf[tid] = sum;
        } else {
          const int dim = p.gq.dims[s - 1];
          const int col = n0 + tid;
          const double xd = col < ncols_valid ? p.xk[size_t(col) * p.d_k + dim] : 0.0;
          p.S_part[(size_t(mt) * p.d_k + dim) * p.ncols_pad + col] = p.gq.w[s - 1] * fma(xd, s0buf[tid], -sum);
        }
      }
      __syncthreads();
    }
    // kernel-input columns outside the group carry no pathwise gradient: the consumers sum every slot, so write zeros
    if (tid < BN) {
      for (int dim = 0; dim < p.d_k; ++dim) {
        bool in_group = false;
        for (int s = 0; s < p.gq.n_dims; ++s) in_group = in_group || p.gq.dims[s] == dim;
        if (!in_group) p.S_part[(size_t(mt) * p.d_k + dim) * p.ncols_pad + n0 + tid] = 0.0;
      }
    }
    CTA_TRACE(4);
    return;
  }

  if constexpr (!UPPER) {
  // lower: V^T[col, row].  4-warp variants: the tile is staged in the (now free) stage buffers, one 1 KB row of V^T per
  // column at a pitch of 132 doubles (the four frag_k groups of a warp land on disjoint bank halves), and leaves as 64 bulk
  // shared->global copies that the TMA engine drains while the CTA finishes -- direct 8-byte stores (eight 32-byte sectors
  // per warp instruction, 128 instructions per thread) measured ~6 400 cycles per CTA with the FP64 pipe idle.
  // 8-warp variants: 8 lanes (frag_r = 0..7) write 64 contiguous bytes per column.
  constexpr bool BULK_EPI = (BN == 64 && THREADS == 128);
  constexpr int VT_PITCH = 132;
  double* vt = stages + 512;  // behind red[5 * BN]
  double nrm[TJ][2];
#pragma unroll
  for (int j = 0; j < TJ; ++j) nrm[j][0] = nrm[j][1] = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = rows[i];
#pragma unroll
    for (int j = 0; j < TJ; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = col_base + 8 * j + e;
        const double v = acc[i][j][e];
        if (p.C) {
          if constexpr (BULK_EPI)
            vt[(col - n0) * VT_PITCH + (row - m0)] = v;
          else
            p.C[size_t(col) * p.ldc + row] = v;
        }
        if (row == p.special_row) {
          p.mu_dot[col] = v;
          red[4 * BN + col - n0] = v;
        } else {
          nrm[j][e] = fma(v, v, nrm[j][e]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < TJ; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      double v = nrm[j][e];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      nrm[j][e] = v;
    }
  if (frag_r == 0) {
#pragma unroll
    for (int j = 0; j < TJ; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) red[wm * BN + wn * (8 * TJ) + 8 * j + 2 * frag_k + e] = nrm[j][e];
  }
  if (BULK_EPI && p.C) fence_proxy_async_smem();
  __syncthreads();
  if (BULK_EPI && p.C && tid < BN) {
    bulk_store_1d(p.C + size_t(n0 + tid) * p.ldc + m0, vt + tid * VT_PITCH, T_BM * sizeof(double));
    bulk_commit();
  }
  if (tid < BN) {
    const double s = ((red[tid] + red[BN + tid]) + red[2 * BN + tid]) + red[3 * BN + tid];
    p.norm_part[size_t(mt) * p.ncols_pad + n0 + tid] = s;
    if (p.fuse_acq && n0 + tid < p.acq_ncols) {
      double val, d_mu, d_var, var;
      acq_eval(p.aq, p.mean_const + red[4 * BN + tid], p.kss - s, &val, &d_mu, &d_var, &var);
      p.acq_a[n0 + tid] = val;
      p.acq_dmu[n0 + tid] = d_mu;
      p.acq_dvar[n0 + tid] = d_var;
    }
  }
  if (BULK_EPI && p.C && tid < BN) bulk_wait_read0();
  CTA_TRACE(4);
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

cudaError_t tma_init() {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess) return e;
    if (q != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const void* kernels[] = {
      reinterpret_cast<const void*>(&trgemm_tma_kernel<MODE_LOWER_GEN, 128>),
      reinterpret_cast<const void*>(&trgemm_tma_kernel<MODE_LOWER_GEN, 32>),
      reinterpret_cast<const void*>(&trgemm_tma_kernel<MODE_LOWER_GEN, 64, 128>),
      reinterpret_cast<const void*>(&trgemm_tma_kernel<MODE_UPPER_GRAD, 64, 128>),
      reinterpret_cast<const void*>(&trgemm_tma_kernel<MODE_LOWER_TMA, 128>),
      reinterpret_cast<const void*>(&trgemm_tma_kernel<MODE_UPPER_W, 128>),
      reinterpret_cast<const void*>(&trgemm_tma_kernel<MODE_LOWER_TMA, 32>),
      reinterpret_cast<const void*>(&trgemm_tma_kernel<MODE_UPPER_W, 32>),
      reinterpret_cast<const void*>(&trgemm_tma_kernel<MODE_UPPER_GRAD, 128>),
      reinterpret_cast<const void*>(&trgemm_tma_kernel<MODE_UPPER_GRAD, 32>),
      reinterpret_cast<const void*>(&trgemm_tma_kernel<MODE_UPPER_GRADQ, 64, 128>),
      reinterpret_cast<const void*>(&trgemm_tma_kernel<MODE_UPPER_GRADQ, 32>),
      reinterpret_cast<const void*>(&trgemm_tma_kernel<MODE_LOWER_TMA, 64, 128>),
  };
  for (const void* k : kernels) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(T_SMEM_BYTES));
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// 2-D map over a row-major [rows, ld] matrix whose contraction index is contiguous: box = 4 k values x box_rows rows
cudaError_t make_operand_map(TensorMap* out, const double* base, int rows, int k_extent, int ld, int box_rows) {
  if (!g_encode) return cudaErrorNotReady;
  cuuint64_t gdim[2] = {cuuint64_t(k_extent), cuuint64_t(rows)};
  cuuint64_t gstride[1] = {cuuint64_t(ld) * sizeof(double)};
  cuuint32_t box[2] = {4, cuuint32_t(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2,
                        const_cast<double*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// 3-D view of the same matrix: (k mod 4, row, k div 4) with strides (8, 8 ld, 32) bytes; box = 4 x box_rows x 4 -> one TMA
// instruction delivers a whole k-tile of an operand in the [k-chunk][row][4] layout (four 2-D boxes before: the elected
// thread's issue work per k-tile fell from 8 instructions to 2)
cudaError_t make_operand_map3(TensorMap* out, const double* base, int rows, int k_extent, int ld, int box_rows) {
  if (!g_encode) return cudaErrorNotReady;
  cuuint64_t gdim[3] = {4, cuuint64_t(rows), cuuint64_t(k_extent / 4)};
  cuuint64_t gstride[2] = {cuuint64_t(ld) * sizeof(double), 4 * sizeof(double)};
  cuuint32_t box[3] = {4, cuuint32_t(box_rows), 4};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3,
                        const_cast<double*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// Column-tile width of the separable-path products (PRB_BN overrides, for A/B runs):
//   64  (default) 4 warps of 32 x 64 per CTA, TWO CTAs per SM: one CTA's TMA prologue, epilogue and end-of-tile barrier
//       overlap the other's main loop -- 1-2 % faster than one 8-warp 128-column CTA per SM at every n from 200 to 4000;
//   32  when the heaviest row tile, not the total work, bounds the launch (narrow_tiles_win): few columns per GPU;
//   128 one 8-warp CTA per SM (the generic-path kernels use this shape).
// 32-column tiles (one 8-warp CTA per SM) against 64-column tiles (two 4-warp CTAs per SM) by a two-term time estimate in
// units of (one k-block of 128) x (one column): the launch takes at least its heaviest row tile -- mtiles k-blocks at the
// rate ONE CTA gets -- and at least the total triangular work spread over 148 SMs.  Rates as measured: a 4-warp CTA gets
// ~0.6 of an SM whether alone or paired, an 8-warp 32-column CTA ~0.8; full launches run at ~0.9 resp. ~0.75 of the SMs.
// The heaviest tile decides for few columns: n = 2000 with 1 024 columns (the 8-GPU share at mc = 128) is 0.345 ms with
// 32-column tiles and 0.581 ms with 64 (r02's tile-count rule picked 64 there); every other swept shape keeps its choice.
static bool narrow_tiles_win(int mtiles, int ncols_pad) {
  const double work = double(ncols_pad) * mtiles * (mtiles + 1) / 2.0;
  const double est64 = fmax(mtiles * 64.0 / 0.6, work / (148.0 * 0.9));
  const double est32 = fmax(mtiles * 32.0 / 0.8, work / (148.0 * 0.75));
  return est32 < est64;
}

static int pick_bn(const Model& m, int ncols_pad, bool upper) {
  const int force = env_knobs().force_bn;
  if (force == 32 || force == 64 || force == 128) return force;
  return narrow_tiles_win((upper ? m.Mp2 : m.Mp1) / T_BM, ncols_pad) ? 32 : 64;
}

static void fill_common(TmaGemmParams& p, const Model& m, int ncols_pad, bool upper, int bn = T_BN) {
  memset(&p, 0, sizeof(p));
#if PRB_TRACE_CTAS
  p.trace = g_cta_trace ? g_cta_trace + size_t(g_cta_trace_launch++) * CTA_TRACE_STRIDE : nullptr;
#endif
  p.mtiles = (upper ? m.Mp2 : m.Mp1) / T_BM;
  p.ntiles = ncols_pad / bn;
  p.Kp = m.Kp;
  p.ncols_pad = ncols_pad;
  p.special_row = upper ? -1 : m.n;
  // scheduling groups of ~4 waves of CTAs (PRB_GROUP_WAVES overrides; 0 = row-tile-major order).  Measured at n = 1000,
  // 65 536 columns: the upper product's DRAM reads fall from 2.34 GB to 0.57 GB per launch (algorithmic 0.54 GB; L2 hit rate
  // 64 % -> 86 %) at unchanged run time -- the kernel is tensor-pipe bound -- with groups of 1 wave 1.5 % slower.
  const int waves = env_knobs().group_waves;
  const int slots = (bn == 64 ? 2 : 1) * 148;
  int g = waves > 0 ? (waves * slots) / (p.mtiles > 0 ? p.mtiles : 1) : p.ntiles;
  if (g < 1) g = 1;
  p.ngroup = g < p.ntiles ? g : p.ntiles;
  const int ngroups = (p.ntiles + p.ngroup - 1) / p.ngroup;
  const int mixed = env_knobs().mixed_order;  // 0 off, 1 all groups but the last, 2 every group
  p.mixed_groups = (bn == 64 && mixed > 0) ? (mixed == 2 ? ngroups : ngroups - 1) : 0;
}

static void fill_sep(TmaGemmParams& p, const Model& m, const SepArgs& sa) {
  p.kc = sa.kc;
  p.Zbits = m.Zbits;
  p.zbits = sa.zbits;
  p.tab = m.sep_tab;
  p.ntab = m.sep_nbits;
  p.mc = sa.mc;
  p.col0 = sa.col0;
  p.nb = sa.nb;
  p.Dc = sa.Dc;
  p.n_cont = sa.n_cont;
  p.colb = sa.colb;
  p.ncols_dev = sa.ncols_dev;
  p.alpha_pad = m.alpha_pad;
}

cudaError_t launch_tma_lower_gen(const Model& m, const SepArgs& sa, int ncols_pad, double* VT, double* norm_part,
                                 double* mu_dot, const FusedAcq* fa, cudaStream_t st) {
  ProfScope prof_(K_TRGEMM_LOWER, st);
  TmaGemmParams p;
  const int bn = pick_bn(m, ncols_pad, false);
  fill_common(p, m, ncols_pad, false, bn);
  fill_sep(p, m, sa);
  p.C = VT;
  p.ldc = m.Mp1;
  p.norm_part = norm_part;
  p.mu_dot = mu_dot;
  if (fa != nullptr && p.mtiles == 1) {
    p.fuse_acq = 1;
    p.aq = fa->aq;
    p.mean_const = m.mean_const;
    p.kss = m.kss;
    p.acq_a = fa->a;
    p.acq_dmu = fa->dmu;
    p.acq_dvar = fa->dvar;
    p.acq_ncols = fa->ncols;
  }
  const CUtensorMap& mA = *reinterpret_cast<const CUtensorMap*>(&m.mapA1k);
  cudaError_t le;
  if (bn == 32)
    le = launch_kernel(trgemm_tma_kernel<MODE_LOWER_GEN, 32, T_THREADS>, dim3(p.mtiles * p.ntiles), dim3(T_THREADS),
                       T_SMEM_BYTES, st, mA, mA, p);
  else if (bn == 64)  // 4 warps per CTA, two CTAs per SM
    le = launch_kernel(trgemm_tma_kernel<MODE_LOWER_GEN, 64, 128>, dim3(p.mtiles * p.ntiles), dim3(128),
                       tma_smem_bytes(64), st, mA, mA, p);
  else
    le = launch_kernel(trgemm_tma_kernel<MODE_LOWER_GEN, 128, T_THREADS>, dim3(p.mtiles * p.ntiles), dim3(T_THREADS),
                       T_SMEM_BYTES, st, mA, mA, p);
  count_launch();
  return le != cudaSuccess ? le : cudaGetLastError();
}

// generic path (materialised operands): 128-column tiles, or 32 when that leaves most SMs idle
static int pick_bn_generic(const Model& m, int ncols_pad, bool upper) {
  return narrow_tiles_win((upper ? m.Mp2 : m.Mp1) / T_BM, ncols_pad) ? 32 : 128;
}

cudaError_t launch_tma_lower(const Model& m, const double* panelK, int panel_rows, int ncols_pad, double* VT,
                             double* norm_part, double* mu_dot, const FusedAcq* fa, cudaStream_t st) {
  ProfScope prof_(K_TRGEMM_LOWER, st);
  TmaGemmParams p;
  const int bn = pick_bn_generic(m, ncols_pad, false) == 32 ? 32 : 64;  // 64: two 4-warp CTAs per SM (as the separable path)
  fill_common(p, m, ncols_pad, false, bn);
  p.C = VT;
  p.ldc = m.Mp1;
  p.norm_part = norm_part;
  p.mu_dot = mu_dot;
  if (fa != nullptr && p.mtiles == 1) {
    p.fuse_acq = 1;
    p.aq = fa->aq;
    p.mean_const = m.mean_const;
    p.kss = m.kss;
    p.acq_a = fa->a;
    p.acq_dmu = fa->dmu;
    p.acq_dvar = fa->dvar;
    p.acq_ncols = fa->ncols;
  }
  TensorMap mapK;  // the k* panel [panel_rows, Kp], box = 4 k values x bn columns
  cudaError_t e = make_operand_map3(&mapK, panelK, panel_rows, m.Kp, m.Kp, bn);
  if (e != cudaSuccess) return e;
  const CUtensorMap& mA = *reinterpret_cast<const CUtensorMap*>(&m.mapA1k);
  const CUtensorMap& mB = *reinterpret_cast<const CUtensorMap*>(&mapK);
  cudaError_t le;
  if (bn == 32)
    le = launch_kernel(trgemm_tma_kernel<MODE_LOWER_TMA, 32, T_THREADS>, dim3(p.mtiles * p.ntiles), dim3(T_THREADS),
                       T_SMEM_BYTES, st, mA, mB, p);
  else
    le = launch_kernel(trgemm_tma_kernel<MODE_LOWER_TMA, 64, 128>, dim3(p.mtiles * p.ntiles), dim3(128),
                       tma_smem_bytes(64), st, mA, mB, p);
  count_launch();
  return le != cudaSuccess ? le : cudaGetLastError();
}

cudaError_t launch_tma_upper(const Model& m, const double* panelV, int panel_rows, int ncols_pad, double* W,
                             cudaStream_t st) {
  ProfScope prof_(K_TRGEMM_UPPER, st);
  TmaGemmParams p;
  const int bn = pick_bn_generic(m, ncols_pad, true);
  fill_common(p, m, ncols_pad, true, bn);
  p.C = W;
  p.ldc = ncols_pad;
  TensorMap mapV;
  cudaError_t e = make_operand_map3(&mapV, panelV, panel_rows, m.Kp, m.Mp1, bn);
  if (e != cudaSuccess) return e;
  const CUtensorMap& mA = *reinterpret_cast<const CUtensorMap*>(&m.mapA2k);
  const CUtensorMap& mB = *reinterpret_cast<const CUtensorMap*>(&mapV);
  if (bn == 32)
    trgemm_tma_kernel<MODE_UPPER_W, 32><<<p.mtiles * p.ntiles, T_THREADS, T_SMEM_BYTES, st>>>(mA, mB, p);
  else
    trgemm_tma_kernel<MODE_UPPER_W, 128><<<p.mtiles * p.ntiles, T_THREADS, T_SMEM_BYTES, st>>>(mA, mB, p);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_tma_upper_grad(const Model& m, const double* panelV, int panel_rows, const SepArgs& sa, int ncols,
                                  int ncols_pad, const double* dmu, const double* dvar, double* S_part,
                                  cudaStream_t st) {
  ProfScope prof_(K_TRGEMM_UPPER, st);
  TmaGemmParams p;
  const int bn = pick_bn(m, ncols_pad, true);
  fill_common(p, m, ncols_pad, true, bn);
  fill_sep(p, m, sa);
  p.dmu = dmu;
  p.dvar = dvar;
  p.ncols = ncols;
  p.S_part = S_part;
  TensorMap mapV;  // the v panel [panel_rows, Mp1], box = 4 k values x bn columns
  cudaError_t e = make_operand_map3(&mapV, panelV, panel_rows, m.Kp, m.Mp1, bn);
  if (e != cudaSuccess) return e;
  const CUtensorMap& mA = *reinterpret_cast<const CUtensorMap*>(&m.mapA2k);
  const CUtensorMap& mB = *reinterpret_cast<const CUtensorMap*>(&mapV);
  if (bn == 32)
    e = launch_kernel(trgemm_tma_kernel<MODE_UPPER_GRAD, 32, T_THREADS>, dim3(p.mtiles * p.ntiles), dim3(T_THREADS),
                      T_SMEM_BYTES, st, mA, mB, p);
  else if (bn == 64)
    e = launch_kernel(trgemm_tma_kernel<MODE_UPPER_GRAD, 64, 128>, dim3(p.mtiles * p.ntiles), dim3(128),
                      tma_smem_bytes(64), st, mA, mB, p);
  else
    e = launch_kernel(trgemm_tma_kernel<MODE_UPPER_GRAD, 128, T_THREADS>, dim3(p.mtiles * p.ntiles), dim3(T_THREADS),
                      T_SMEM_BYTES, st, mA, mB, p);
  count_launch();
  return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_tma_upper_gradq(const Model& m, const double* panelV, int panel_rows, int ncols, int ncols_pad,
                                   const GradGroup& gq, const double* Qp, const double* xk, const double* dmu,
                                   const double* dvar, double* S_part, cudaStream_t st) {
  ProfScope prof_(K_TRGEMM_UPPER, st);
  TmaGemmParams p;
  const int bn = pick_bn_generic(m, ncols_pad, true) == 32 ? 32 : 64;  // 64: two 4-warp CTAs per SM, as UPPER_GRAD
  fill_common(p, m, ncols_pad, true, bn);
  p.alpha_pad = m.alpha_pad;
  p.dmu = dmu;
  p.dvar = dvar;
  p.ncols = ncols;
  p.S_part = S_part;
  p.gq = gq;
  p.Qp = Qp;
  p.xk = xk;
  p.Xtr = m.Xtr;
  p.n = m.n;
  p.d_k = m.d_k;
  TensorMap mapV;  // the v panel [panel_rows, Mp1], box = 4 k values x bn columns
  cudaError_t e = make_operand_map3(&mapV, panelV, panel_rows, m.Kp, m.Mp1, bn);
  if (e != cudaSuccess) return e;
  const CUtensorMap& mA = *reinterpret_cast<const CUtensorMap*>(&m.mapA2k);
  const CUtensorMap& mB = *reinterpret_cast<const CUtensorMap*>(&mapV);
  if (bn == 32)
    e = launch_kernel(trgemm_tma_kernel<MODE_UPPER_GRADQ, 32, T_THREADS>, dim3(p.mtiles * p.ntiles), dim3(T_THREADS),
                      T_SMEM_BYTES, st, mA, mB, p);
  else
    e = launch_kernel(trgemm_tma_kernel<MODE_UPPER_GRADQ, 64, 128>, dim3(p.mtiles * p.ntiles), dim3(128),
                      tma_smem_bytes(64), st, mA, mB, p);
  count_launch();
  return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace prb
