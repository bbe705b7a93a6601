// FP64 triangular panel products on the sm_100a DMMA pipe.
//
//   lower:  V^T[col, j] = sum_{k <= j} A1[j, k] * Kstar^T[col, k]      A1 = [Linv ; alpha^T]   (posterior variance + mean)
//   upper:  W[j, col]   = sum_{k >= j} A2[j, k] * V^T[col, k]          A2 = Linv^T            (d sigma^2 / d k*)
//
// This is the dominant cost of the PR hot path: the reference does it as a dense k* @ L^-T GEMM inside GPyTorch's exact
// prediction strategy ([3P], reached from probabilistic_reparameterization.py:491) and again in autograd's backward
// (:624-633).  Here both operands are K-contiguous ("NT" product), tiles are 128 x 128 x 16, staged through shared memory
// with a 4-deep cp.async pipeline, and multiplied with mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) -- sm_100a has no tcgen05 kind
// for FP64, DMMA is the FP64 tensor path (measured 37.1 TFLOP/s, profiles/r01_fp64_peak_microbench.txt).  Tiles that lie
// entirely in the zero triangle are never visited; heavy tiles are scheduled first.
//
// Epilogue fusion (lower): per-column partial ||v||^2 over the tile's rows (fixed-order reduction => bitwise
// reproducible), and the alpha row is peeled off as the posterior-mean dot product.
#include "prb_internal.cuh"

namespace prb {

struct TrGemmParams {
  const double* A;
  const double* Bt;
  double* C;
  double* norm_part;
  double* mu_dot;
  int lda, ldb, ldc;
  int mtiles, ntiles, Kp;
  int upper;
  int special_row;
  int ncols_pad;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) trgemm_dmma_kernel(const TrGemmParams p) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;                                          // [STAGES][BM][LDS]
  double* Bs = smem + GEMM_STAGES * GEMM_BM * GEMM_LDS;       // [STAGES][BN][LDS]

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int wm = warp & 3;   // 4 warps along M (32 rows each)
  const int wn = warp >> 2;  // 2 warps along N (64 cols each)

  // heavy tiles first: lower -> largest row block first, upper -> smallest row block first
  const int t = blockIdx.x;
  const int nt = t % p.ntiles;
  const int mi_sched = t / p.ntiles;
  const int mt = p.upper ? mi_sched : (p.mtiles - 1 - mi_sched);
  const int m0 = mt * GEMM_BM;
  const int n0 = nt * GEMM_BN;
  int kbeg, kend;
  if (p.upper) {
    kbeg = m0;
    kend = p.Kp;
  } else {
    kbeg = 0;
    kend = (mt == p.mtiles - 1) ? p.Kp : min(p.Kp, m0 + GEMM_BM);
  }
  const int num_k = (kend - kbeg) / GEMM_BK;

  // global -> shared: thread loads 8 doubles (4 x 16 B) of one row of A and of one row of B^T per stage
  const int ld_row = tid >> 1;
  const int ld_seg = (tid & 1) * 8;
  const double* gA = p.A + size_t(m0 + ld_row) * p.lda + kbeg + ld_seg;
  const double* gB = p.Bt + size_t(n0 + ld_row) * p.ldb + kbeg + ld_seg;
  double* sA = As + ld_row * GEMM_LDS + ld_seg;
  double* sB = Bs + ld_row * GEMM_LDS + ld_seg;

  auto load_stage = [&](int kt, int slot) {
    const double* a = gA + size_t(kt) * GEMM_BK;
    const double* b = gB + size_t(kt) * GEMM_BK;
    double* da = sA + slot * (GEMM_BM * GEMM_LDS);
    double* db = sB + slot * (GEMM_BN * GEMM_LDS);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      cp_async16(da + 2 * c, a + 2 * c);
      cp_async16(db + 2 * c, b + 2 * c);
    }
  };

  double acc[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < GEMM_STAGES - 1; ++s) {
    if (s < num_k) load_stage(s, s);
    cp_async_commit();
  }

  const int frag_r = lane >> 2;
  const int frag_k = lane & 3;
  const double* fa = As + (wm * 32 + frag_r) * GEMM_LDS + frag_k;
  const double* fb = Bs + (wn * 64 + frag_r) * GEMM_LDS + frag_k;

  for (int kt = 0; kt < num_k; ++kt) {
    cp_async_wait<GEMM_STAGES - 2>();
    __syncthreads();
    {
      const int nk = kt + GEMM_STAGES - 1;
      if (nk < num_k) load_stage(nk, nk % GEMM_STAGES);
      cp_async_commit();
    }
    const int slot = kt % GEMM_STAGES;
    const double* pa = fa + slot * (GEMM_BM * GEMM_LDS);
    const double* pb = fb + slot * (GEMM_BN * GEMM_LDS);
#pragma unroll
    for (int kk = 0; kk < GEMM_BK; kk += 4) {
      double a[4], b[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = pa[i * 8 * GEMM_LDS + kk];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = pb[j * 8 * GEMM_LDS + kk];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();

  // ---- epilogue -------------------------------------------------------------------------------------------------
  const int row_base = m0 + wm * 32 + frag_r;          // + 8 * i
  const int col_base = n0 + wn * 64 + 2 * frag_k;      // + 8 * j + e
  if (p.upper) {
    // W[row, col], col contiguous: 16-byte stores, 4 lanes cover 64 contiguous bytes
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double* crow = p.C + size_t(row_base + 8 * i) * p.ldc + col_base;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
        *reinterpret_cast<double2*>(crow + 8 * j) = v;
      }
    }
    return;
  }
  // lower: V^T[col, row]; 8 lanes (frag_r = 0..7) write 64 contiguous bytes per column
  double nrm[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) nrm[j][0] = nrm[j][1] = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = row_base + 8 * i;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = col_base + 8 * j + e;
        const double v = acc[i][j][e];
        p.C[size_t(col) * p.ldc + row] = v;
        if (row == p.special_row) {
          p.mu_dot[col] = v;
        } else {
          nrm[j][e] = fma(v, v, nrm[j][e]);
        }
      }
    }
  }
  // reduce over the 8 row-lanes of the warp (lane bits 2..4), then over the 4 M-warps in a fixed order
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      double v = nrm[j][e];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      nrm[j][e] = v;
    }
  __syncthreads();  // pipeline buffers are dead: reuse as red[4][128]
  double* red = smem;
  if (frag_r == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) red[wm * GEMM_BN + wn * 64 + 8 * j + 2 * frag_k + e] = nrm[j][e];
  }
  __syncthreads();
  if (tid < GEMM_BN) {
    const double s = ((red[tid] + red[GEMM_BN + tid]) + red[2 * GEMM_BN + tid]) + red[3 * GEMM_BN + tid];
    p.norm_part[size_t(mt) * p.ncols_pad + n0 + tid] = s;
  }
}

cudaError_t gemm_init() {
  return cudaFuncSetAttribute(trgemm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(GEMM_SMEM_BYTES));
}

cudaError_t launch_trgemm_lower(const Model& m, const double* KstarT, int ncols_pad, double* VT, double* norm_part,
                                double* mu_dot, cudaStream_t st) {
  ProfScope prof_(K_TRGEMM_LOWER, st);
  TrGemmParams p;
  p.A = m.A1;
  p.Bt = KstarT;
  p.C = VT;
  p.norm_part = norm_part;
  p.mu_dot = mu_dot;
  p.lda = m.Kp;
  p.ldb = m.Kp;
  p.ldc = m.Mp1;
  p.mtiles = m.Mp1 / GEMM_BM;
  p.ntiles = ncols_pad / GEMM_BN;
  p.Kp = m.Kp;
  p.upper = 0;
  p.special_row = m.n;
  p.ncols_pad = ncols_pad;
  trgemm_dmma_kernel<<<p.mtiles * p.ntiles, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(p);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_trgemm_upper(const Model& m, const double* VT, int ncols_pad, double* W, cudaStream_t st) {
  ProfScope prof_(K_TRGEMM_UPPER, st);
  TrGemmParams p;
  p.A = m.A2;
  p.Bt = VT;
  p.C = W;
  p.norm_part = nullptr;
  p.mu_dot = nullptr;
  p.lda = m.Kp;
  p.ldb = m.Mp1;
  p.ldc = ncols_pad;
  p.mtiles = m.Mp2 / GEMM_BM;
  p.ntiles = ncols_pad / GEMM_BN;
  p.Kp = m.Kp;
  p.upper = 1;
  p.special_row = -1;
  p.ncols_pad = ncols_pad;
  trgemm_dmma_kernel<<<p.mtiles * p.ntiles, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(p);
  count_launch();
  return cudaGetLastError();
}

}  // namespace prb
