// GP-state builder on the device: from K + diag(noise) and the targets to the packed factors the hot
// path reads -- A1 = [Linv ; alpha^T], A2 = Linv^T, alpha -- without cuSOLVER / cuBLAS.
//
// Reference: the caches GPyTorch's exact prediction strategy builds behind model.posterior ([3P]: psd_safe_cholesky, the
// explicit inverse root, mean_cache), reached from experiment_utils.py:237-351 once per BO iteration.
//
//   1. blocked right-looking Cholesky, 64 x 64 blocks.  Per block column k: one CTA factors the diagonal block in shared memory
//      and inverts the factor (64 x 64 triangular inverse); the panel below is multiplied by that inverse; the trailing matrix
//      gets its rank-64 update.  Panel and update are the same FP64 DMMA block-GEMM kernel (64 x 64 output tile per CTA, four
//      warps of 32 x 32, mma.sync.m8n8k4.f64).
//   2. Linv by divide and conquer on the lower-triangular factor:
//        inv([[A, 0], [B, C]]) = [[inv A, 0], [-inv C * B * inv A, inv C]]
//      starting from the 64 x 64 diagonal inverses of step 1 and doubling the block size; every level is two batched block
//      GEMMs (all pairs and all tiles in parallel), with the k-range of each tile clipped to the triangular operand's
//      non-zeros.  log2(n / 64) levels instead of n / 64 dependent block rows.
//   3. alpha = Linv^T (Linv (y - c)): two matrix-vector kernels; then the packing into the padded A1 / A2 layouts.
// All reductions run in a fixed order: the state is bitwise reproducible.
#include "prb_device.cuh"
#include "prb_kernels.cuh"
#include "prb_ptx.cuh"

namespace prb {

constexpr int CB = 64;  // block size of the factorisation and of the GEMM tiles

// W (lower triangle + diagonal) = K + diag(noise), identity on the padding
__global__ void chol_init_kernel(const double* __restrict__ K, const double* __restrict__ noise, int n, int Np,
                                 double* __restrict__ W) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= Np) return;
  double v = 0.0;
  if (i < n && j < n) {
    if (j <= i) v = K[size_t(i) * n + j];
    if (i == j) v += noise[i];
  } else if (i == j) {
    v = 1.0;
  }
  W[size_t(i) * Np + j] = v;
}

// Diagonal block k: L_kk = chol(A_kk) in place (lower triangle) and Dinv[k] = inv(L_kk) (full 64 x 64, zeros above the
// diagonal).  This kernel is the serial spine of the factorisation -- n / 64 of them run back to back -- so it is organised
// for latency: one CTA of 256 threads, thread (tr, tc) keeps the 4 x 4 sub-blocks [4 tr .., 4 tc ..] of A AND of X = inv(L) in
// registers for the whole factorisation.  Step c (64 of them, ONE barrier each, buffers alternate):
//   the owners publish the raw column c of A and the raw row c of X;  everyone forms 1 / sqrt(pivot) itself;
//   right-looking updates with independent FMAs only:  A[r][cc] -= L[r][c] L[cc][c]   (c < cc <= r)
//   forward substitution fused in:  X[c][:] = raw / pivot ;  X[r][:] -= L[r][c] X[c][:]   (r > c)
// `info` receives 1 + (global row) of the first non-positive pivot (the matrix is not positive definite to working precision).
__global__ void __launch_bounds__(256) chol_diag_kernel(double* __restrict__ W, int Np, int k, double* __restrict__ Dinv,
                                                         int* __restrict__ info) {
  __shared__ double colA[2][CB], rowX[2][CB];
  const int tid = threadIdx.x, tr = tid >> 4, tc = tid & 15;
  double* blk = W + (size_t(k) * CB) * Np + size_t(k) * CB;
  double a[4][4], x[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = 4 * tr + i, cc = 4 * tc + j;
      a[i][j] = (cc <= r) ? blk[size_t(r) * Np + cc] : 0.0;
      x[i][j] = (cc == r) ? 1.0 : 0.0;
    }
  for (int c = 0; c < CB; ++c) {
    const int buf = c & 1, cq = c >> 2, cm = c & 3;
    if (tc == cq) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j == cm) {
#pragma unroll
          for (int i = 0; i < 4; ++i) colA[buf][4 * tr + i] = a[i][j];
        }
    }
    if (tr == cq) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i == cm) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rowX[buf][4 * tc + j] = x[i][j];
        }
    }
    __syncthreads();
    const double piv = colA[buf][c];
    const bool ok = piv > 0.0;
    if (!ok && tid == 0) atomicCAS(info, 0, 1 + k * CB + c);
    // 1 / sqrt(pivot) straight from a single-precision seed and two Newton steps, then d = pivot * inv: one ~75-cycle seed +
    // seven 8-cycle FP64 ops on the step's critical path instead of sqrt (~79) followed by a divide (~59)
    // (profiles/r02_fp64_latency.txt)
    const double pv = ok ? piv : 1.0;
    double inv = double(rsqrtf(__double2float_rn(pv)));
    const double hp = 0.5 * pv;
    inv = inv * fma(-hp * inv, inv, 1.5);  // 2^-22 -> 2^-44
    inv = inv * fma(-hp * inv, inv, 1.5);  // -> working precision
    const double d = pv * inv;
    if (4 * tr + 3 < c) continue;  // every row of this thread is final (warp w owns rows 8 w .. 8 w + 7: whole warps retire)
    double lr[4], lc[4], xr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) lr[i] = (4 * tr + i > c) ? colA[buf][4 * tr + i] * inv : 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      lc[j] = (4 * tc + j > c) ? colA[buf][4 * tc + j] * inv : 0.0;
      xr[j] = rowX[buf][4 * tc + j] * inv;
    }
    // Unconditional updates: lr is zero for finished rows, lc for finished columns, xr beyond column c -- and entries above
    // the diagonal are never read or stored -- so no per-element predicate is needed (they were half of the instructions).
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a[i][j] = fma(-lr[i], lc[j], a[i][j]);
        x[i][j] = fma(-lr[i], xr[j], x[i][j]);
      }
    if (tc == cq) {  // column c of L is final: scaled entries below the pivot, the pivot's square root on the diagonal
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j == cm) {
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i][j] = (4 * tr + i > c) ? lr[i] : (4 * tr + i == c ? d : a[i][j]);
        }
    }
    if (tr == cq) {  // row c of X is final
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i == cm) {
#pragma unroll
          for (int j = 0; j < 4; ++j) x[i][j] = xr[j];
        }
    }
  }
  double* dv = Dinv + size_t(k) * CB * CB;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = 4 * tr + i, cc = 4 * tc + j;
      if (cc <= r) blk[size_t(r) * Np + cc] = a[i][j];
      dv[r * CB + cc] = (cc <= r) ? x[i][j] : 0.0;
    }
}

// ---- FP64 DMMA block GEMM: one 64 x 64 tile of C per CTA, C = alpha * A * op(B) (+ C) ------------------------------------
// A [64 x K] row-major (lda); B either [K x 64] row-major (NN) or [64 x K] row-major used transposed (NT).  Four warps, each a
// 32 x 32 sub-tile = 4 x 4 accumulator tiles of m8n8.  k is consumed in chunks of 16 staged through shared memory in the
// [k-chunk of 4][row][4] layout of the product kernels (every fragment load is one conflict-free LDS.64); the next chunk is
// fetched into registers while the current one multiplies.
struct BlkGemmArgs {
  int mode;  // 0 panel, 1 trailing update, 2 inverse level (B * invA), 3 inverse level (-invC * M)
  int Np, T;
  int k;     // block column (modes 0, 1)
  int s;     // half block size in units of CB (modes 2, 3)
  double* W;          // factor / working matrix [Np, Np]
  double* X;          // inverse [Np, Np]
  double* Tm;         // scratch [Np, Np]
  const double* Dinv; // [T][64][64]
};

constexpr int SLAB = 32;                      // k values per pipeline stage
constexpr int SLAB_DOUBLES = CB * SLAB;       // per operand and stage: [SLAB / 4 chunks][64 rows][4]
constexpr size_t BLK_GEMM_SMEM = size_t(2) * 2 * SLAB_DOUBLES * sizeof(double);  // 2 stages x (A + B) = 64 KB

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Operands go global -> shared with cp.async (no register staging), two slabs of 32 k values in flight: a K = 64 product (the
// panel and the trailing update: most launches of the factorisation) issues ALL its loads up front and pays one memory
// latency instead of four.
template <bool TRANSB>
__device__ __forceinline__ void blk_gemm_tile(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                                              double* __restrict__ C, int ldc, int kbeg, int kend, double alpha,
                                              bool accumulate, double* smem) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 1, wn = warp >> 1;
  const int fr = lane >> 2, fk = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  auto issue = [&](int stage, int k0) {
    double* As = smem + stage * (2 * SLAB_DOUBLES);
    double* Bs = As + SLAB_DOUBLES;
    // A (and B when NT): 64 rows x 32 k = 1024 pairs of k; thread -> pieces tid, tid + 128, ...: (row, pair)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int piece = tid + 128 * q;
      const int row = piece >> 4, kp = (piece & 15) * 2;  // k offset inside the slab (even)
      cp_async16(As + (kp >> 2) * (CB * 4) + row * 4 + (kp & 3), A + size_t(row) * lda + k0 + kp);
      if (TRANSB) cp_async16(Bs + (kp >> 2) * (CB * 4) + row * 4 + (kp & 3), B + size_t(row) * ldb + k0 + kp);
    }
    if (!TRANSB) {
      // NN: B [k][n] row-major -> Bs[k / 4][n][k % 4]: element-wise copies (consecutive n are 32 bytes apart in shared memory)
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int e = tid + 128 * q;
        const int kk = e >> 6, nn = e & 63;
        cp_async8(Bs + (kk >> 2) * (CB * 4) + nn * 4 + (kk & 3), B + size_t(k0 + kk) * ldb + nn);
      }
    }
    cp_async_commit();
  };
  const int nslab = (kend - kbeg) / SLAB;
  if (nslab > 0) issue(0, kbeg);
  if (nslab > 1) issue(1, kbeg + SLAB);
  for (int sl = 0; sl < nslab; ++sl) {
    if (sl + 1 < nslab) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();
    const double* As = smem + (sl & 1) * (2 * SLAB_DOUBLES);
    const double* Bs = As + SLAB_DOUBLES;
#pragma unroll
    for (int c = 0; c < SLAB / 4; ++c) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[c * (CB * 4) + (wm * 32 + 8 * i + fr) * 4 + fk];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[c * (CB * 4) + (wn * 32 + 8 * j + fr) * 4 + fk];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncthreads();  // the stage is free again
    if (sl + 2 < nslab) issue(sl & 1, kbeg + (sl + 2) * SLAB);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = wm * 32 + 8 * i + fr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double2* cp = reinterpret_cast<double2*>(C + size_t(row) * ldc + wn * 32 + 8 * j + 2 * fk);
      double2 v = make_double2(alpha * acc[i][j][0], alpha * acc[i][j][1]);
      if (accumulate) {
        const double2 o = *cp;
        v.x += o.x;
        v.y += o.y;
      }
      *cp = v;
    }
  }
}

__global__ void __launch_bounds__(128) blk_gemm_kernel(const BlkGemmArgs g) {
  extern __shared__ __align__(16) unsigned char gemm_smem_raw[];
  double* smem = reinterpret_cast<double*>(gemm_smem_raw);
  const int Np = g.Np;
  const int t = blockIdx.x;
  if (g.mode == 0) {  // panel: L_ik = A_ik * inv(L_kk)^T  (in place; the tile is fully read before it is written)
    const int i = g.k + 1 + t;
    double* blk = g.W + (size_t(i) * CB) * Np + size_t(g.k) * CB;
    blk_gemm_tile<true>(blk, Np, g.Dinv + size_t(g.k) * CB * CB, CB, blk, Np, 0, CB, 1.0, false, smem);
  } else if (g.mode == 1) {  // trailing update: A_ij -= L_ik L_jk^T for k < j <= i
    const int m = g.T - g.k - 1;  // blocks below the panel
    // t -> (ii, jj) in the lower triangle of an m x m block grid (row-major over rows)
    int ii = int((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while (ii * (ii + 1) / 2 > t) --ii;
    while ((ii + 1) * (ii + 2) / 2 <= t) ++ii;
    const int jj = t - ii * (ii + 1) / 2;
    if (ii >= m) return;
    const int i = g.k + 1 + ii, j = g.k + 1 + jj;
    blk_gemm_tile<true>(g.W + (size_t(i) * CB) * Np + size_t(g.k) * CB, Np, g.W + (size_t(j) * CB) * Np + size_t(g.k) * CB,
                        Np, g.W + (size_t(i) * CB) * Np + size_t(j) * CB, Np, 0, CB, -1.0, true, smem);
  } else {
    // inverse level: diagonal blocks of s * CB rows are already inverted in X; pair p = (2p, 2p + 1) of them
    const int s = g.s, tiles = s * s;
    const int p = t / tiles, r = (t % tiles) / s, c = t % s;
    const size_t lo = size_t(2 * p) * s * CB, hi = lo + size_t(s) * CB;  // first row of the A / C diagonal blocks
    if (hi + size_t(r) * CB >= size_t(g.T) * CB) return;  // row tile entirely in the identity padding: its result is zero
    if (g.mode == 2) {  // M = B * inv(A): inv(A) is lower triangular -> column block c has rows >= c
      blk_gemm_tile<false>(g.W + (hi + size_t(r) * CB) * Np + lo, Np, g.X + lo * Np + lo + size_t(c) * CB, Np,
                           g.Tm + (hi + size_t(r) * CB) * Np + lo + size_t(c) * CB, Np, c * CB, s * CB, 1.0, false, smem);
    } else {  // X21 = -inv(C) * M: row block r of inv(C) has columns <= r
      blk_gemm_tile<false>(g.X + (hi + size_t(r) * CB) * Np + hi, Np, g.Tm + hi * Np + lo + size_t(c) * CB, Np,
                           g.X + (hi + size_t(r) * CB) * Np + lo + size_t(c) * CB, Np, 0, (r + 1) * CB, -1.0, false, smem);
    }
  }
}

// X diagonal blocks <- Dinv, everything else zero
__global__ void inv_init_kernel(const double* __restrict__ Dinv, int Np, double* __restrict__ X) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= Np) return;
  const int bi = i / CB, bj = j / CB;
  X[size_t(i) * Np + j] = (bi == bj) ? Dinv[size_t(bi) * CB * CB + (i % CB) * CB + (j % CB)] : 0.0;
}

// t = Linv r (one warp per row, lane-strided + xor tree), then alpha = Linv^T t (one thread per column: coalesced rows)
__global__ void __launch_bounds__(256) lower_matvec_kernel(const double* __restrict__ X, int Np, int n,
                                                            const double* __restrict__ y, double mean_const,
                                                            double* __restrict__ out) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  double s = 0.0;
  for (int j = lane; j <= row; j += 32) s = fma(X[size_t(row) * Np + j], y[j] - mean_const, s);
  s = warp_sum(s);
  if (lane == 0) out[row] = s;
}
// alpha[m] = sum_{k >= m} X[k][m] t[k]: a CTA owns 32 columns (lane <-> column: coalesced rows), its 8 warps split the rows,
// four partial sums per lane, warp partials combined in a fixed order
__global__ void __launch_bounds__(256) lower_t_matvec_kernel(const double* __restrict__ X, int Np, int n,
                                                              const double* __restrict__ t, double* __restrict__ alpha) {
  __shared__ double part[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m0 = blockIdx.x * 32, m = m0 + lane;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (m < n) {
    int k = m0 + warp;
    for (; k + 24 < n; k += 32) {
      s0 = fma(k >= m ? X[size_t(k) * Np + m] : 0.0, t[k], s0);
      s1 = fma(k + 8 >= m ? X[size_t(k + 8) * Np + m] : 0.0, t[k + 8], s1);
      s2 = fma(k + 16 >= m ? X[size_t(k + 16) * Np + m] : 0.0, t[k + 16], s2);
      s3 = fma(k + 24 >= m ? X[size_t(k + 24) * Np + m] : 0.0, t[k + 24], s3);
    }
    for (; k < n; k += 8) s0 = fma(k >= m ? X[size_t(k) * Np + m] : 0.0, t[k], s0);
  }
  part[warp][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (warp == 0 && m < n) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += part[w][lane];
    alpha[m] = s;
  }
}

// A1[j, k] = Linv[j, k] (k <= j < n) ; A1[n, k] = alpha[k] ; 0 elsewhere.  A2[m, k] = Linv[k, m] (m <= k < n) ; 0 elsewhere.
__global__ void pack_from_inverse_kernel(const double* __restrict__ X, int ldx, const double* __restrict__ alpha, int n,
                                         int Kp, int Mp1, int Mp2, double* __restrict__ A1, double* __restrict__ A2) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (k >= Kp) return;
  if (r < Mp1) {
    double v = 0.0;
    if (r < n && k <= r) v = X[size_t(r) * ldx + k];
    if (r == n && k < n) v = alpha[k];
    A1[size_t(r) * Kp + k] = v;
  }
  if (r < Mp2) {
    double v = 0.0;
    if (r < n && k < n && k >= r) v = X[size_t(k) * ldx + r];
    A2[size_t(r) * Kp + k] = v;
  }
}

size_t chol_scratch_bytes(int n) {
  int T = (n + CB - 1) / CB, Tp = 1;
  while (Tp < T) Tp *= 2;
  const size_t Np = size_t(Tp) * CB;
  return (3 * Np * Np + size_t(Tp) * CB * CB + Np) * sizeof(double) + 256;
}

// K [n, n] row-major (symmetric; the lower triangle is read), noise [n], y [n] -> A1, A2, alpha of the model.  `scratch` holds
// chol_scratch_bytes(n); `info_dev` a device int (0 on success).  Optionally exports the dense Linv [n, n] (lower, zeros above).
cudaError_t build_gp_state(const double* K, const double* noise, const double* y, double mean_const, int n, int Kp, int Mp1,
                           int Mp2, double* A1, double* A2, double* alpha, double* Linv_out, void* scratch, int* info_dev,
                           cudaStream_t st) {
  ProfScope prof_(K_MISC, st);
  const int T = (n + CB - 1) / CB;
  int Tp = 1;
  while (Tp < T) Tp *= 2;
  const int Np = Tp * CB;
  double* W = static_cast<double*>(scratch);
  double* X = W + size_t(Np) * Np;
  double* Tm = X + size_t(Np) * Np;
  double* Dinv = Tm + size_t(Np) * Np;
  double* tvec = Dinv + size_t(Tp) * CB * CB;
  cudaError_t e = cudaFuncSetAttribute(blk_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(BLK_GEMM_SMEM));
  if (e != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(info_dev, 0, sizeof(int), st)) != cudaSuccess) return e;
  dim3 g2((Np + 127) / 128, Np);
  chol_init_kernel<<<g2, 128, 0, st>>>(K, noise, n, Np, W);
  BlkGemmArgs ga;
  memset(&ga, 0, sizeof(ga));
  ga.Np = Np;
  ga.T = T;
  ga.W = W;
  ga.X = X;
  ga.Tm = Tm;
  ga.Dinv = Dinv;
  int launches = 1;
  for (int k = 0; k < Tp; ++k) {
    chol_diag_kernel<<<1, 256, 0, st>>>(W, Np, k, Dinv, info_dev);  // identity blocks on the padding factor trivially
    ++launches;
    if (k >= T - 1) continue;
    ga.k = k;
    ga.mode = 0;
    blk_gemm_kernel<<<T - k - 1, 128, BLK_GEMM_SMEM, st>>>(ga);
    ga.mode = 1;
    const int m = T - k - 1;
    blk_gemm_kernel<<<m * (m + 1) / 2, 128, BLK_GEMM_SMEM, st>>>(ga);
    launches += 2;
  }
  inv_init_kernel<<<g2, 128, 0, st>>>(Dinv, Np, X);
  ++launches;
  for (int s = 1; s < Tp; s *= 2) {
    const int pairs = Tp / (2 * s);
    ga.s = s;
    ga.mode = 2;
    blk_gemm_kernel<<<pairs * s * s, 128, BLK_GEMM_SMEM, st>>>(ga);
    ga.mode = 3;
    blk_gemm_kernel<<<pairs * s * s, 128, BLK_GEMM_SMEM, st>>>(ga);
    launches += 2;
  }
  lower_matvec_kernel<<<(n + 7) / 8, 256, 0, st>>>(X, Np, n, y, mean_const, tvec);
  lower_t_matvec_kernel<<<(n + 31) / 32, 256, 0, st>>>(X, Np, n, tvec, alpha);
  dim3 gp((Kp + 127) / 128, Mp1 > Mp2 ? Mp1 : Mp2);
  pack_from_inverse_kernel<<<gp, 128, 0, st>>>(X, Np, alpha, n, Kp, Mp1, Mp2, A1, A2);
  launches += 3;
  if (Linv_out != nullptr) {
    e = cudaMemcpy2DAsync(Linv_out, size_t(n) * sizeof(double), X, size_t(Np) * sizeof(double), size_t(n) * sizeof(double), n,
                          cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return e;
  }
  count_launch(launches);
  return cudaGetLastError();
}

}  // namespace prb
