// PR hot-path kernels around the two triangular panel products: sampler, cross-covariance panel, acquisition epilogue,
// pathwise-gradient contraction, MC / analytic reductions, EMA-baseline update, Adam step.
#include "prb_device.cuh"
#include "prb_kernels.cuh"
#include "prb_ptx.cuh"

namespace prb {

// =====================================================================================================================
// Philox materialisation (parity entry; the fused sampler calls philox_uniform() directly)
// =====================================================================================================================
__global__ void philox_fill_kernel(uint64_t seed, uint64_t offset, int mc, int n_cols, double* out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= mc * n_cols) return;
  const int i = idx / n_cols, c = idx % n_cols;
  out[idx] = philox_uniform(seed, offset, uint32_t(i), uint32_t(c));
}

cudaError_t launch_philox_fill(uint64_t seed, uint64_t offset, int mc, int n_cols, double* out, cudaStream_t st) {
  ProfScope prof_(K_MISC, st);
  const int total = mc * n_cols;
  if (total == 0) return cudaSuccess;
  philox_fill_kernel<<<(total + 255) / 256, 256, 0, st>>>(seed, offset, mc, n_cols, out);
  count_launch();
  return cudaGetLastError();
}

// =====================================================================================================================
// PR sampler: one thread per (restart b, MC sample i).
//   reference: Normalize(reverse) -> MCProbabilisticReparameterizationInputTransform.transform (input.py:637-746)
//              -> Normalize; rounding component z (probabilistic_reparameterization.py:454-465); prob (:470-475);
//              OneHotToNumeric (input.py:69-88) when apply_numeric.
// Base uniforms are indexed by (i, discrete parameter) only => shared by all restarts, as in the reference.
// =====================================================================================================================
// grid (ceil(mc / 128), b): a CTA samples 128 MC draws of one restart.  Rows are staged in shared memory and written out
// with fully coalesced stores (a row per thread would stride 8 d bytes: 12 % of HBM peak measured; profiles/).
constexpr int SAMPLE_ROWS = 128;
__global__ void __launch_bounds__(SAMPLE_ROWS) pr_sample_kernel(
    const DevSpace sp, const double* __restrict__ theta, int b, int mc, const double* __restrict__ u_int,
    const double* __restrict__ u_cat, uint64_t seed, uint64_t offset, double* __restrict__ xk,
    double* __restrict__ tilde, uint8_t* __restrict__ zout, double* __restrict__ prob, const BitMap bm,
    uint32_t* __restrict__ zbits, const double* __restrict__ lower, const double* __restrict__ upper,
    double* __restrict__ Xc) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ double th_s[PRB_MAX_DIMS];  // the restart's theta row, clamped to [lower, upper] if given (gen.py:304-305, 323)
  double* ps = reinterpret_cast<double*>(smraw);          // [n_disc]
  double* offs = ps + sp.n_disc;                          // [n_int]
  double* st_t = offs + sp.n_int;                         // [128][d]    if tilde
  double* st_x = st_t + (tilde ? SAMPLE_ROWS * sp.d : 0); // [128][d_k]  if xk
  uint8_t* st_z = reinterpret_cast<uint8_t*>(st_x + (xk ? SAMPLE_ROWS * sp.d_k : 0));  // [128][n_disc] if zout
  const int bi = blockIdx.y, tid = threadIdx.x;
  const int i0 = blockIdx.x * SAMPLE_ROWS;
  const int i = i0 + tid;
  const int nrows = min(SAMPLE_ROWS, mc - i0);
  if (tid < sp.d) {
    double v = theta[size_t(bi) * sp.d + tid];
    if (lower) v = fmin(fmax(v, lower[tid]), upper[tid]);
    th_s[tid] = v;
    if (Xc && blockIdx.x == 0) Xc[size_t(bi) * sp.d + tid] = v;
  }
  __syncthreads();
  const double* th = th_s;
  pr_restart_probs(sp, th, tid, SAMPLE_ROWS, ps, offs);
  __syncthreads();
  if (i < mc) {
    const uint32_t bits = pr_sample_point(sp, th, ps, offs, i, u_int, u_cat, seed, offset,
                                          xk ? st_x + tid * sp.d_k : nullptr, tilde ? st_t + tid * sp.d : nullptr,
                                          zout ? st_z + tid * sp.n_disc : nullptr, bm);
    if (zbits) zbits[size_t(bi) * mc + i] = bits;
  }
  if (prob && blockIdx.x == 0)
    for (int k = tid; k < sp.n_disc; k += SAMPLE_ROWS) prob[size_t(bi) * sp.n_disc + k] = ps[k];
  __syncthreads();
  const size_t row0 = size_t(bi) * mc + i0;
  if (tilde)
    for (int e = tid; e < nrows * sp.d; e += SAMPLE_ROWS) tilde[row0 * sp.d + e] = st_t[e];
  if (xk)
    for (int e = tid; e < nrows * sp.d_k; e += SAMPLE_ROWS) xk[row0 * sp.d_k + e] = st_x[e];
  if (zout)
    for (int e = tid; e < nrows * sp.n_disc; e += SAMPLE_ROWS) zout[row0 * sp.n_disc + e] = st_z[e];
}

cudaError_t launch_pr_sample(const DevSpace& sp, const double* theta, int b, int mc, const double* u_int,
                             const double* u_cat, uint64_t seed, uint64_t offset, double* xk, double* tilde,
                             uint8_t* z, double* prob, const BitMap* bm, uint32_t* zbits, cudaStream_t st,
                             const double* lower, const double* upper, double* Xc) {
  ProfScope prof_(K_SAMPLE, st);
  if (int64_t(b) * mc == 0) return cudaSuccess;
  BitMap none;
  for (int k = 0; k < PRB_MAX_DIMS; ++k) none.pos[k] = -1;
  const size_t smem = size_t(sp.n_disc + sp.n_int) * sizeof(double) +
                      (tilde ? size_t(SAMPLE_ROWS) * sp.d * sizeof(double) : 0) +
                      (xk ? size_t(SAMPLE_ROWS) * sp.d_k * sizeof(double) : 0) + (z ? size_t(SAMPLE_ROWS) * sp.n_disc : 0);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(pr_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
  }
  dim3 grid((mc + SAMPLE_ROWS - 1) / SAMPLE_ROWS, b);
  pr_sample_kernel<<<grid, SAMPLE_ROWS, smem, st>>>(sp, theta, b, mc, u_int, u_cat, seed, offset, xk, tilde, z, prob,
                                                   bm ? *bm : none, bm ? zbits : nullptr, lower, upper, Xc);
  count_launch();
  return cudaGetLastError();
}

// =====================================================================================================================
// Analytic PR "sampler": one thread per (restart b, discrete configuration o).  Builds the kernel input
// (input.py:466-488 + Normalize + OneHotToNumeric) and P(z | theta) (input.py:410-464).
// options [n_opt, n_int + n_cat]: un-normalised integer value / category index.
// =====================================================================================================================
// Optional fused clamp (prb_analytic_adam): theta is clamped to [lower, upper] on the fly and configuration 0 of every
// restart writes the clamped row to Xc (which may alias theta only if no other thread still reads the raw row: the fused
// step passes a separate buffer, the final evaluation clamps in a launch of its own).
__global__ void analytic_build_kernel(const DevSpace sp, const double* __restrict__ theta, int b, int n_opt,
                                      const int32_t* __restrict__ options, double* __restrict__ xk,
                                      double* __restrict__ probs, const double* __restrict__ lower,
                                      const double* __restrict__ upper, double* __restrict__ Xc) {
  const int64_t col = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (col >= int64_t(b) * n_opt) return;
  const int bi = int(col / n_opt);
  const int o = int(col % n_opt);
  auto th = [&](int c) -> double {
    const double v = theta[size_t(bi) * sp.d + c];
    return lower ? fmin(fmax(v, lower[c]), upper[c]) : v;
  };
  if (Xc && o == 0)
    for (int c = 0; c < sp.d; ++c) Xc[size_t(bi) * sp.d + c] = th(c);
  const int32_t* op = options + size_t(o) * (sp.n_int + sp.n_cat);
  double* xr = xk ? xk + size_t(col) * sp.d_k : nullptr;
  const int numeric_start = sp.n_cat > 0 ? sp.cat_start[0] : sp.d;
  if (xr)
    for (int c = 0; c < sp.d; ++c)
      if (!sp.apply_numeric || c < numeric_start) xr[c] = th(c);
  double P = 1.0;
  for (int k = 0; k < sp.n_int; ++k) {
    const int c = sp.int_cols[k];
    const double lo = sp.int_lo[k], hi = sp.int_hi[k];
    const double x_un = sp.raw_int ? th(c) : unnormalize_int(th(c), lo, hi);
    double off;
    const double p = int_round_prob(x_un, sp.tau, &off);
    const double zv = double(op[k]);
    const double diff = (off + 1.0) - zv;
    const double f = (diff == 0.0) ? p : ((diff == 1.0) ? (1.0 - p) : 0.0);
    P *= f;
    if (xr) xr[c] = sp.raw_int ? zv : __dsub_rn(zv, lo) / __dsub_rn(hi, lo);
  }
  for (int j = 0; j < sp.n_cat; ++j) {
    const int s = sp.cat_start[j], card = sp.cat_card[j];
    const int cat = op[sp.n_int + j];
    double mx = -1e300;
    for (int k = 0; k < card; ++k) mx = fmax(mx, __dsub_rn(th(s + k), 0.5) / sp.tau);
    double sum = 0.0, sel = 0.0;
    for (int k = 0; k < card; ++k) {
      const double e = exp(__dsub_rn(__dsub_rn(th(s + k), 0.5) / sp.tau, mx));
      sum += e;
      if (k == cat) sel = e;
    }
    P *= sel * (1.0 / sum);
    if (xr) {
      if (sp.apply_numeric) {
        write_categorical_input(sp, xr, numeric_start, j, cat);
      } else {
        for (int k = 0; k < card; ++k) xr[s + k] = (k == cat) ? 1.0 : 0.0;
      }
    }
  }
  if (probs) probs[col] = P;
}

cudaError_t launch_analytic_build(const DevSpace& sp, const double* theta, int b, int n_opt, const int32_t* options,
                                  double* xk, double* probs, cudaStream_t st, const double* lower, const double* upper,
                                  double* Xc) {
  ProfScope prof_(K_ANALYTIC, st);
  const int64_t total = int64_t(b) * n_opt;
  if (total == 0) return cudaSuccess;
  analytic_build_kernel<<<unsigned((total + 127) / 128), 128, 0, st>>>(sp, theta, b, n_opt, options, xk, probs, lower,
                                                                       upper, Xc);
  count_launch();
  return cudaGetLastError();
}

// =====================================================================================================================
// Cross-covariance panel  Kstar^T[col, j] = k(x_col, X_j)   (eval-major, training index contiguous, zero padded).
//   reference: covar_module(X*, X_train) inside model.posterior ([3P]; kernel structure kernels.py:18-101).
// grid (ceil(Kp/128), ncols_pad/32); thread <-> training point j, CTA loops over 32 evaluation points.
// =====================================================================================================================
constexpr int KS_ROWS = 128;
constexpr int KS_COLS = 32;

__device__ __forceinline__ double eval_kernel_value(const DevSpec& spec, const double* __restrict__ xs /*[d_k]*/,
                                                    const double* __restrict__ Xs /*[d_k][KS_ROWS] + tid*/) {
  double kval = 0.0;
  for (int t = 0; t < spec.n_terms; ++t) {
    const DevTerm& T = spec.t[t];
    double acc = T.additive ? 0.0 : 1.0;
    for (int f = 0; f < T.n_factors; ++f) {
      const DevFactor& F = T.f[f];
      double v;
      if (F.kind == PRB_FACTOR_MATERN52) {
        double r2 = 0.0;
        for (int di = 0; di < F.n_dims; ++di) {
          const int dim = F.dims[di];
          const double df = (xs[dim] - Xs[dim * KS_ROWS]) * F.inv_ls[di];
          r2 = fma(df, df, r2);
        }
        v = matern52_value(r2);
      } else {
        double s = 0.0;
        for (int di = 0; di < F.n_dims; ++di) {
          const int dim = F.dims[di];
          s += (xs[dim] != Xs[dim * KS_ROWS]) ? F.inv_ls[di] : 0.0;
        }
        v = exp(-s);
      }
      v = ipow(v, F.power);
      acc = T.additive ? acc + v : acc * v;
    }
    kval = fma(T.outputscale, acc, kval);
  }
  return kval;
}

__global__ void __launch_bounds__(KS_ROWS) kstar_panel_kernel(const DevSpec spec, const double* __restrict__ Xtr, int n,
                                                               int Kp, const double* __restrict__ xk, int ncols,
                                                               int ncols_store, double* __restrict__ KstarT,
                                                               int ks_cols) {
  extern __shared__ double sm[];
  const int d_k = spec.d_k;
  double* Xs = sm;                     // [d_k][KS_ROWS]
  double* xs = sm + d_k * KS_ROWS;     // [ks_cols][d_k]
  const int tid = threadIdx.x;
  const int j = blockIdx.x * KS_ROWS + tid;
  const int col0 = blockIdx.y * ks_cols;
  const bool valid = j < n;
  for (int dd = 0; dd < d_k; ++dd) Xs[dd * KS_ROWS + tid] = valid ? Xtr[size_t(j) * d_k + dd] : 0.0;
  for (int idx = tid; idx < ks_cols * d_k; idx += KS_ROWS) {
    const int c = idx / d_k;
    xs[idx] = (col0 + c < ncols) ? xk[size_t(col0) * d_k + idx] : 0.0;
  }
  __syncthreads();
  if (j >= Kp) return;
  for (int c = 0; c < ks_cols; ++c) {
    const int col = col0 + c;
    double v = 0.0;
    if (col >= ncols_store) break;
    if (valid && col < ncols) v = eval_kernel_value(spec, xs + c * d_k, Xs + tid);
    KstarT[size_t(col) * Kp + j] = v;
  }
}

// ---- hot-path variant: k* and, when the call needs the pathwise gradient, the derivative factor Q of the gradient group ----
// Same grid and thread mapping as kstar_panel_kernel.  Branch-free sqrt / exp (one ulp) so that the per-column evaluations of
// a thread interleave; the Matern value and its derivative factor share one sqrt and one exp, so Q costs a handful of
// multiplies on top of k* -- the separate gradient-contraction kernel this replaces evaluated every Matern a second time
// (4.1 ms against 4.4 ms of products at 65 536 x 1000).
bool build_grad_group(const DevSpec& spec, uint64_t diffmask, GradGroup* out) {
  memset(out, 0, sizeof(GradGroup));
  const DevFactor* first = nullptr;
  for (int t = 0; t < spec.n_terms; ++t) {
    const DevTerm& T = spec.t[t];
    for (int f = 0; f < T.n_factors; ++f) {
      const DevFactor& F = T.f[f];
      if (F.kind != PRB_FACTOR_MATERN52) continue;
      bool touches = false;
      for (int di = 0; di < F.n_dims; ++di) touches = touches || ((diffmask >> F.dims[di]) & 1u);
      if (!touches) continue;
      if (first == nullptr) {
        first = &F;
      } else {  // every member must be the same factor (dims and lengthscales): one Q panel serves them all
        if (F.n_dims != first->n_dims) return false;
        for (int di = 0; di < F.n_dims; ++di)
          if (F.dims[di] != first->dims[di] || F.inv_ls[di] != first->inv_ls[di]) return false;
      }
      out->member[t][f] = 1;
    }
  }
  if (first == nullptr || spec.n_unique == 0) return false;
  for (int di = 0; di < first->n_dims; ++di) {
    if (!((diffmask >> first->dims[di]) & 1u)) continue;
    if (out->n_dims == GQ_MAX_DIMS) return false;
    out->dims[out->n_dims] = first->dims[di];
    out->w[out->n_dims] = first->inv_ls[di] * first->inv_ls[di];
    ++out->n_dims;
  }
  return out->n_dims > 0;
}

static_assert(KS_COLS % 4 == 0, "double4 loads of the kernel-input tile");
constexpr int KQ_NC = 4;  // columns evaluated together by a thread: amortises the per-dim loads and loop overhead over four
                          // independent distance / sqrt / exp chains
template <bool WITH_Q>
__global__ void __launch_bounds__(KS_ROWS) kstar_q_kernel(const DevSpec spec, const GradGroup gq,
                                                           const double* __restrict__ Xtr, int n, int Kp,
                                                           const double* __restrict__ xk, int ncols, int ncols_store,
                                                           double* __restrict__ KstarT, double* __restrict__ Qp,
                                                           int ks_cols) {
  extern __shared__ double sm[];
  __shared__ int fdim_s[4][PRB_MAX_DIMS];     // dims / inverse lengthscales of the distinct factor functions: the inner loop
  __shared__ double fils_s[4][PRB_MAX_DIMS];  // reads them from shared memory (indexed constant-bank loads cost an LDC each)
  __shared__ int fnd_s[4], fkind_s[4];
  __shared__ int tf_u[PRB_MAX_TERMS][PRB_MAX_FACTORS], tf_pw[PRB_MAX_TERMS][PRB_MAX_FACTORS],
      tf_mem[PRB_MAX_TERMS][PRB_MAX_FACTORS], t_nf[PRB_MAX_TERMS], t_add[PRB_MAX_TERMS];
  __shared__ double t_sc[PRB_MAX_TERMS];
  const int d_k = spec.d_k;
  double* Xs = sm;                  // [d_k][KS_ROWS]
  double* xs = sm + d_k * KS_ROWS;  // [d_k][KS_COLS]  (dim-major: a thread's KQ_NC columns are contiguous)
  double* vs = xs + d_k * KS_COLS;  // [4 functions][KQ_NC][KS_ROWS] factor values of the thread's current columns ...
  double* gs = vs + 4 * KQ_NC * KS_ROWS;  // ... and their derivative factors (indexed by the runtime function id: shared
                                          // memory instead of select chains over registers)
  const int tid = threadIdx.x;
  const int j = blockIdx.x * KS_ROWS + tid;
  const int col0 = blockIdx.y * ks_cols;
  const bool valid = j < n;
  grid_dep_launch();
  for (int idx = tid; idx < 4 * PRB_MAX_DIMS; idx += KS_ROWS) {
    const int u = idx / PRB_MAX_DIMS, di = idx % PRB_MAX_DIMS;
    if (u < spec.n_unique) {
      const DevFactor& F = spec.t[spec.first_t[u]].f[spec.first_f[u]];
      fdim_s[u][di] = di < F.n_dims ? F.dims[di] : 0;
      fils_s[u][di] = di < F.n_dims ? F.inv_ls[di] : 0.0;
      if (di == 0) {
        fnd_s[u] = F.n_dims;
        fkind_s[u] = F.kind;
      }
    }
  }
  if (tid < PRB_MAX_TERMS * PRB_MAX_FACTORS) {
    const int t = tid / PRB_MAX_FACTORS, f = tid % PRB_MAX_FACTORS;
    const bool live = t < spec.n_terms && f < spec.t[t].n_factors;
    tf_u[t][f] = live ? spec.uid[t][f] : 0;
    tf_pw[t][f] = live ? spec.t[t].f[f].power : 1;
    tf_mem[t][f] = (WITH_Q && live) ? gq.member[t][f] : 0;
    if (f == 0) {
      t_nf[t] = t < spec.n_terms ? spec.t[t].n_factors : 0;
      t_add[t] = t < spec.n_terms ? spec.t[t].additive : 0;
      t_sc[t] = t < spec.n_terms ? spec.t[t].outputscale : 0.0;
    }
  }
  for (int dd = 0; dd < d_k; ++dd) Xs[dd * KS_ROWS + tid] = valid ? Xtr[size_t(j) * d_k + dd] : 0.0;
  grid_dep_wait();  // xk comes from the sampler kernel
  for (int idx = tid; idx < ks_cols * d_k; idx += KS_ROWS) {
    const int c = idx / d_k, dd = idx - c * d_k;
    xs[dd * KS_COLS + c] = (col0 + c < ncols) ? xk[size_t(col0) * d_k + idx] : 0.0;
  }
  __syncthreads();
  if (j >= Kp) return;
  const double* Xj = Xs + tid;
  const int n_unique = spec.n_unique;
  for (int cb = 0; cb < ks_cols; cb += KQ_NC) {
    if (col0 + cb >= ncols_store) break;
    double kval[KQ_NC], qval[KQ_NC];
#pragma unroll
    for (int c = 0; c < KQ_NC; ++c) kval[c] = qval[c] = 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (u >= n_unique) continue;
      const int nd = fnd_s[u];
      double acc[KQ_NC];
#pragma unroll
      for (int c = 0; c < KQ_NC; ++c) acc[c] = 0.0;
      if (fkind_s[u] == PRB_FACTOR_MATERN52) {
        for (int di = 0; di < nd; ++di) {
          const int dim = fdim_s[u][di];
          const double il = fils_s[u][di];
          const double xj = Xj[dim * KS_ROWS];
          const double4 xv = *reinterpret_cast<const double4*>(xs + dim * KS_COLS + cb);
          const double x4[KQ_NC] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
          for (int c = 0; c < KQ_NC; ++c) {
            const double df = (x4[c] - xj) * il;
            acc[c] = fma(df, df, acc[c]);
          }
        }
#pragma unroll
        for (int c = 0; c < KQ_NC; ++c) {
          double G;
          vs[(u * KQ_NC + c) * KS_ROWS + tid] = matern52_nb(acc[c], &G);
          if (WITH_Q) gs[(u * KQ_NC + c) * KS_ROWS + tid] = G;
        }
      } else {
        for (int di = 0; di < nd; ++di) {
          const int dim = fdim_s[u][di];
          const double il = fils_s[u][di];
          const double xj = Xj[dim * KS_ROWS];
          const double4 xv = *reinterpret_cast<const double4*>(xs + dim * KS_COLS + cb);
          const double x4[KQ_NC] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
          for (int c = 0; c < KQ_NC; ++c) acc[c] += (x4[c] != xj) ? il : 0.0;
        }
#pragma unroll
        for (int c = 0; c < KQ_NC; ++c) {
          vs[(u * KQ_NC + c) * KS_ROWS + tid] = exp_neg_nb(-acc[c]);
          if (WITH_Q) gs[(u * KQ_NC + c) * KS_ROWS + tid] = 0.0;
        }
      }
    }
    // combine per term (runtime loops over the term's factors, the thread's columns unrolled inside): product or sum of
    // F^p; Q adds outputscale * d(term)/dF_f * G_f for the gradient group's member factors
    const int n_terms = spec.n_terms;
    for (int t = 0; t < n_terms; ++t) {
      const int nf = t_nf[t];
      const bool additive = t_add[t] != 0;
      const double sc = t_sc[t];
      double acc[KQ_NC];
#pragma unroll
      for (int c = 0; c < KQ_NC; ++c) acc[c] = additive ? 0.0 : 1.0;
      for (int f = 0; f < nf; ++f) {
        const int u = tf_u[t][f], pw = tf_pw[t][f];
#pragma unroll
        for (int c = 0; c < KQ_NC; ++c) {
          const double v = vs[(u * KQ_NC + c) * KS_ROWS + tid];
          double Fp = v;
          for (int i = 1; i < pw; ++i) Fp *= v;
          acc[c] = additive ? acc[c] + Fp : acc[c] * Fp;
        }
      }
#pragma unroll
      for (int c = 0; c < KQ_NC; ++c) kval[c] = fma(sc, acc[c], kval[c]);
      if (WITH_Q) {
        for (int f = 0; f < nf; ++f) {
          if (!tf_mem[t][f]) continue;
          const int u = tf_u[t][f], pw = tf_pw[t][f];
          double q[KQ_NC];
#pragma unroll
          for (int c = 0; c < KQ_NC; ++c) {
            const double v = vs[(u * KQ_NC + c) * KS_ROWS + tid];
            double dp = double(pw);
            for (int i = 2; i < pw + 1 && pw > 1; ++i) dp *= v;  // p v^(p-1)
            q[c] = sc * gs[(u * KQ_NC + c) * KS_ROWS + tid] * (pw > 1 ? dp : 1.0);
          }
          if (!additive) {
            for (int g = 0; g < nf; ++g) {
              if (g == f) continue;
              const int u2 = tf_u[t][g], pw2 = tf_pw[t][g];
#pragma unroll
              for (int c = 0; c < KQ_NC; ++c) {
                const double v2 = vs[(u2 * KQ_NC + c) * KS_ROWS + tid];
                double Fp = v2;
                for (int i = 1; i < pw2; ++i) Fp *= v2;
                q[c] *= Fp;
              }
            }
          }
#pragma unroll
          for (int c = 0; c < KQ_NC; ++c) qval[c] += q[c];
        }
      }
    }
#pragma unroll
    for (int c = 0; c < KQ_NC; ++c) {
      const int col = col0 + cb + c;
      if (col < ncols_store) {
        const bool live = valid && col < ncols;
        KstarT[size_t(col) * Kp + j] = live ? kval[c] : 0.0;
        if (WITH_Q) Qp[size_t(col) * Kp + j] = live ? qval[c] : 0.0;
      }
    }
  }
}

cudaError_t launch_kstar_panel(const Model& m, const double* xk, int ncols, int ncols_pad, double* KstarT,
                               cudaStream_t st, const GradGroup* gq, double* Qp) {
  ProfScope prof_(K_KSTAR, st);
  // few columns: 4 per CTA instead of 32, so that a small panel still spreads over the SMs
  const int ks_cols = (m.Kp + KS_ROWS - 1) / KS_ROWS * (ncols_pad / KS_COLS) < 148 ? 4 : KS_COLS;
  dim3 grid((m.Kp + KS_ROWS - 1) / KS_ROWS, ncols_pad / ks_cols);
  const size_t smem = size_t(m.d_k) * (KS_ROWS + KS_COLS) * sizeof(double);
  const bool with_q = gq != nullptr && Qp != nullptr;
  const size_t smem_q = smem + size_t(2) * 4 * KQ_NC * KS_ROWS * sizeof(double);  // + the factor value / derivative staging
  if (m.spec.n_unique == 0) {  // more than four distinct factor functions: the general entry-by-entry kernel
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kstar_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
      if (e != cudaSuccess) return e;
    }
    kstar_panel_kernel<<<grid, KS_ROWS, smem, st>>>(m.spec, m.Xtr, m.n, m.Kp, xk, ncols, ncols_pad, KstarT, ks_cols);
    count_launch();
    return cudaGetLastError();
  }
  auto kern = with_q ? kstar_q_kernel<true> : kstar_q_kernel<false>;
  if (smem_q > 40 * 1024) {  // the 48 KB default limit covers static (3.2 KB of tables) + dynamic shared memory
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_q));
    if (e != cudaSuccess) return e;
  }
  GradGroup none;
  memset(&none, 0, sizeof(none));
  cudaError_t le = launch_kernel(kern, grid, dim3(KS_ROWS), smem_q, st, m.spec, with_q ? *gq : none, m.Xtr, m.n, m.Kp, xk,
                                 ncols, ncols_pad, KstarT, Qp, ks_cols);
  count_launch();
  return le != cudaSuccess ? le : cudaGetLastError();
}

// plain K(xk, X_train) -> out[n_points, n] (leading dimension n, nothing stored beyond n_points)
cudaError_t launch_kernel_matrix(const DevSpec& spec, const double* Xtr, int n, const double* xk, int n_points,
                                 double* out, cudaStream_t st) {
  ProfScope prof_(K_KSTAR, st);
  if (n_points == 0) return cudaSuccess;
  dim3 grid((n + KS_ROWS - 1) / KS_ROWS, (n_points + KS_COLS - 1) / KS_COLS);
  const size_t smem = size_t(spec.d_k) * (KS_ROWS + KS_COLS) * sizeof(double);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kstar_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
  }
  kstar_panel_kernel<<<grid, KS_ROWS, smem, st>>>(spec, Xtr, n, n, xk, n_points, n_points, out, KS_COLS);
  count_launch();
  return cudaGetLastError();
}

// =====================================================================================================================
// Separable structure ("mixed" kernel of kernels.py:39-63, 85-89 with binary integer dims):
//   k*(x, X_j) = [s * M52_ARD(x_c, X_jc)] * M52_iso(Hamming(z, Z_j))  -- the second factor takes nbits + 1 distinct values.
// Model side (once per GP state): bit codes of the training points, the Hamming table, and a check that the binary
// columns of X_train really are {0, 1}.  Call side (once per step): kc[b, j] and Dc[b, c, j] per restart.
// =====================================================================================================================
__global__ void sep_model_prep_kernel(const DevFactor fb, const double* __restrict__ Xtr, int n, int d_k, int n_pad,
                                      uint32_t* __restrict__ Zbits, double* __restrict__ tab, int* __restrict__ bad) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n_pad) {
    uint32_t bits = 0;
    if (j < n) {
      for (int di = 0; di < fb.n_dims; ++di) {
        const double v = Xtr[size_t(j) * d_k + fb.dims[di]];
        if (v != 0.0 && v != 1.0) atomicExch(bad, 1);
        if (v != 0.0) bits |= 1u << di;
      }
    }
    Zbits[j] = bits;
  }
  if (j < 64) {
    // same accumulation order as the generic kernel evaluation: r2 = fma(df, df, r2) with df = +-1/l per differing dim
    double r2 = 0.0;
    const double il = fb.inv_ls[0];
    for (int h = 0; h < j && h < fb.n_dims; ++h) r2 = fma(il, il, r2);
    tab[j] = (j <= fb.n_dims) ? matern52_value(r2) : 0.0;
  }
}

cudaError_t launch_sep_model_prep(const Model& m, int n_pad, int* bad_flag, cudaStream_t st) {
  ProfScope prof_(K_MISC, st);
  const DevFactor& fb = m.spec.t[0].f[m.sep_bin_factor];
  const int threads = 128;
  const int blocks = (max(n_pad, 64) + threads - 1) / threads;
  sep_model_prep_kernel<<<blocks, threads, 0, st>>>(fb, m.Xtr, m.n, m.d_k, n_pad, m.Zbits, m.sep_tab, bad_flag);
  count_launch();
  return cudaGetLastError();
}

// grid (ceil(Kp / 128), nb): thread <-> training point j, block row <-> restart
__global__ void sep_restart_prep_kernel(const DevFactor fc, double outputscale, const double* __restrict__ Xtr, int n,
                                        int d_k, int Kp, const double* __restrict__ theta, int d_theta,
                                        double* __restrict__ kc, double* __restrict__ Dc) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (j >= Kp) return;
  const double* th = theta + size_t(b) * d_theta;
  double val = 0.0, G = 0.0;
  if (j < n) {
    double r2 = 0.0;
    for (int di = 0; di < fc.n_dims; ++di) {
      const int dim = fc.dims[di];
      const double df = (th[dim] - Xtr[size_t(j) * d_k + dim]) * fc.inv_ls[di];
      r2 = fma(df, df, r2);
    }
    val = outputscale * matern52(r2, &G);
  }
  kc[size_t(b) * Kp + j] = val;
  if (Dc) {
    for (int di = 0; di < fc.n_dims; ++di) {
      const int dim = fc.dims[di];
      double v = 0.0;
      if (j < n) v = outputscale * G * (fc.inv_ls[di] * fc.inv_ls[di]) * (th[dim] - Xtr[size_t(j) * d_k + dim]);
      Dc[(size_t(b) * fc.n_dims + di) * Kp + j] = v;
    }
  }
}

cudaError_t launch_sep_restart_prep(const Model& m, const double* theta, int nb, int d_theta, double* kc, double* Dc,
                                    cudaStream_t st) {
  ProfScope prof_(K_KSTAR, st);
  if (nb == 0) return cudaSuccess;
  const DevFactor& fc = m.spec.t[0].f[m.sep_cont_factor];
  dim3 grid((m.Kp + 127) / 128, nb);
  sep_restart_prep_kernel<<<grid, 128, 0, st>>>(fc, m.spec.t[0].outputscale, m.Xtr, m.n, m.d_k, m.Kp, theta, d_theta, kc,
                                               Dc);
  count_launch();
  return cudaGetLastError();
}

// alpha zero-padded to n_pad entries
__global__ void pad_copy_kernel(const double* __restrict__ src, int n, int n_pad, double* __restrict__ dst) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n_pad) dst[j] = j < n ? src[j] : 0.0;
}
cudaError_t launch_pad_copy(const double* src, int n, int n_pad, double* dst, cudaStream_t st) {
  ProfScope prof_(K_MISC, st);
  pad_copy_kernel<<<(n_pad + 127) / 128, 128, 0, st>>>(src, n, n_pad, dst);
  count_launch();
  return cudaGetLastError();
}

// =====================================================================================================================
// Acquisition epilogue: posterior mean / variance -> EI / UCB / mean and its partial derivatives.
//   reference: [3P] gpytorch variance clamp (min_variance), botorch ExpectedImprovement / UpperConfidenceBound /
//   PosteriorMean; call sites experiment_utils.py:391-393, 476-480, 489-492.
// =====================================================================================================================
__global__ void acq_epilogue_kernel(const DevAcq aq, double mean_const, double kss, const double* __restrict__ norm_part,
                                    int mtiles, int ncols_pad, const double* __restrict__ mu_dot, int ncols,
                                    double* __restrict__ a, double* __restrict__ dmu, double* __restrict__ dvar,
                                    double* __restrict__ out_mean, double* __restrict__ out_var,
                                    const int* __restrict__ ncols_dev, int64_t col0) {
  grid_dep_launch();
  grid_dep_wait();
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  if (ncols_dev != nullptr && col0 + col >= int64_t(*ncols_dev)) return;
  double vs = 0.0;
  for (int t = 0; t < mtiles; ++t) vs += norm_part[size_t(t) * ncols_pad + col];
  double val, d_mu, d_var, mean, var;
  acq_eval(aq, mean_const + mu_dot[col], kss - vs, &val, &d_mu, &d_var, &var);
  mean = mean_const + mu_dot[col];
  a[col] = val;
  if (dmu) dmu[col] = d_mu;
  if (dvar) dvar[col] = d_var;
  if (out_mean) out_mean[col] = mean;
  if (out_var) out_var[col] = var;
}

cudaError_t launch_acq_epilogue(const Model& m, const DevAcq& aq, const double* norm_part, int ncols_pad,
                                const double* mu_dot, int ncols, double* a, double* dmu, double* dvar,
                                double* out_mean, double* out_var, const int* ncols_dev, int64_t col0,
                                cudaStream_t st) {
  ProfScope prof_(K_EPILOGUE, st);
  if (ncols == 0) return cudaSuccess;
  cudaError_t e = launch_kernel(acq_epilogue_kernel, dim3((ncols + 127) / 128), dim3(128), 0, st, aq, m.mean_const, m.kss,
                                norm_part, m.Mp1 / GEMM_BM, ncols_pad, mu_dot, ncols, a, dmu, dvar, out_mean, out_var,
                                ncols_dev, col0);
  count_launch();
  return e != cudaSuccess ? e : cudaGetLastError();
}

// =====================================================================================================================
// Random-Fourier-feature posterior sample (Thompson sampling, `pr__ts`):  f(x) = scale * sum_j w_j cos(x . W[:, j] + b_j)
//   reference: rffs.py:24-133 (get_gp_samples -> deterministic model) under PosteriorMean (experiment_utils.py:481-492);
//   the feature map is BoTorch's RandomFourierFeatures [3P]: cos((x / l) @ omega + b) * sqrt(2 outputscale / m), with the
//   lengthscales folded into W by the host.  One warp per evaluation point, lanes stride over the features (W is [d_k, m]:
//   coalesced), gradient  df/dx_c = -scale * sum_j w_j sin(.) W[c, j]  for the dims in `diffmask`, 8 dims per pass.
// =====================================================================================================================
constexpr int RFF_WARPS = 8;
constexpr int RFF_MAXG = 8;  // gradient dims per pass
// WPC warps cooperate on one evaluation point (WPC = 1 for wide panels, RFF_WARPS when there are too few points to fill the
// GPU otherwise); the feature range is split between them and the partial sums meet in shared memory (fixed order).
template <int WPC>
__global__ void __launch_bounds__(RFF_WARPS * 32) rff_eval_kernel(const double* __restrict__ W, const double* __restrict__ bias,
                                                                  const double* __restrict__ w, int m, int d_k,
                                                                  double scale, const double* __restrict__ xk,
                                                                  int64_t ncols, uint64_t diffmask, int need_grad,
                                                                  double* __restrict__ a, double* __restrict__ S_all,
                                                                  double* __restrict__ out_mean,
                                                                  double* __restrict__ out_var, int64_t dim_stride) {
  constexpr int CPB = RFF_WARPS / WPC;  // evaluation points per block
  __shared__ double xs[CPB][PRB_MAX_DIMS];
  __shared__ double part[RFF_WARPS][1 + RFF_MAXG];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cslot = warp / WPC, sub = warp % WPC;
  const int64_t col = int64_t(blockIdx.x) * CPB + cslot;
  const bool valid = col < ncols;
  if (valid && sub == 0 && lane < d_k) xs[cslot][lane] = xk[size_t(col) * d_k + lane];
  __syncthreads();
  const double* x = xs[cslot];
  int dims[PRB_MAX_DIMS];
  int nd = 0;
  if (need_grad)
    for (int c = 0; c < d_k; ++c)
      if ((diffmask >> c) & 1u) dims[nd++] = c;
  const int per = (m + WPC - 1) / WPC;
  const int j0 = sub * per, j1 = min(m, j0 + per);
  for (int pass = 0; pass == 0 || pass * RFF_MAXG < nd; ++pass) {
    double f = 0.0;
    double g[RFF_MAXG] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (valid) {
      for (int j = j0 + lane; j < j1; j += 32) {
        double ph = bias[j];
        for (int c = 0; c < d_k; ++c) ph = fma(x[c], W[size_t(c) * m + j], ph);
        double sn, cs;
        sincos(ph, &sn, &cs);
        const double wj = w[j];
        f = fma(wj, cs, f);
        const double ws = -wj * sn;
#pragma unroll
        for (int q = 0; q < RFF_MAXG; ++q)
          if (pass * RFF_MAXG + q < nd) g[q] = fma(ws, W[size_t(dims[pass * RFF_MAXG + q]) * m + j], g[q]);
      }
    }
    f = warp_sum(f);
#pragma unroll
    for (int q = 0; q < RFF_MAXG; ++q) g[q] = warp_sum(g[q]);
    if (WPC > 1) {
      __syncthreads();  // the previous pass has been read
      if (lane == 0) {
        part[warp][0] = f;
#pragma unroll
        for (int q = 0; q < RFF_MAXG; ++q) part[warp][1 + q] = g[q];
      }
      __syncthreads();
      if (sub == 0 && lane == 0) {
        for (int s2 = 1; s2 < WPC; ++s2) {
          f += part[warp + s2][0];
#pragma unroll
          for (int q = 0; q < RFF_MAXG; ++q) g[q] += part[warp + s2][1 + q];
        }
      }
    }
    if (valid && sub == 0 && lane == 0) {
      if (pass == 0) {
        const double fv = scale * f;
        if (a) a[col] = fv;
        if (out_mean) out_mean[col] = fv;
        if (out_var) out_var[col] = 0.0;
      }
#pragma unroll
      for (int q = 0; q < RFF_MAXG; ++q)
        if (pass * RFF_MAXG + q < nd) {
          // dim_stride == 0: S[col, dim] (the engine's layout); else S[dim * dim_stride + col] (the fused step's finish kernel)
          const int dd = dims[pass * RFF_MAXG + q];
          S_all[dim_stride ? size_t(dd) * dim_stride + col : size_t(col) * d_k + dd] = scale * g[q];
        }
    }
  }
}

cudaError_t launch_rff_eval(const Model& m, const double* xk, int64_t ncols, uint64_t diffmask, bool need_grad, double* a,
                            double* S_all, double* out_mean, double* out_var, cudaStream_t st, int64_t dim_stride) {
  ProfScope prof_(K_KSTAR, st);
  if (ncols == 0) return cudaSuccess;
  if (ncols < 8 * 148 * RFF_WARPS) {  // few points: all eight warps of a block share one point
    rff_eval_kernel<RFF_WARPS><<<unsigned(ncols), RFF_WARPS * 32, 0, st>>>(m.rff_W, m.rff_b, m.rff_w, m.rff_m, m.d_k, m.rff_scale,
                                                                         xk, ncols, diffmask, need_grad ? 1 : 0, a, S_all,
                                                                         out_mean, out_var, dim_stride);
  } else {
    const unsigned grid = unsigned((ncols + RFF_WARPS - 1) / RFF_WARPS);
    rff_eval_kernel<1><<<grid, RFF_WARPS * 32, 0, st>>>(m.rff_W, m.rff_b, m.rff_w, m.rff_m, m.d_k, m.rff_scale, xk, ncols,
                                                      diffmask, need_grad ? 1 : 0, a, S_all, out_mean, out_var, dim_stride);
  }
  count_launch();
  return cudaGetLastError();
}

// =====================================================================================================================
// Pathwise-gradient contraction:  S[col, dim] = sum_j C_j(col) * d k*_j(col) / d x_dim ,
//   C_j = dA/dmu * alpha_j - 2 dA/dsigma^2 * W[j, col]      (W = Linv^T Linv k* = K^-1 k*)
//   reference: torch.autograd.grad(mean_acq_values, cont_input) through the GP posterior
//   (probabilistic_reparameterization.py:624-633).
// grid (ncols_pad/128, Mp2/128): thread <-> evaluation point, CTA covers 128 training rows; partial sums per row block are
// written to S_part[jblock][dim][col] and summed in a fixed order afterwards (bitwise reproducible).
// =====================================================================================================================
constexpr int GR_COLS = 128;

__global__ void __launch_bounds__(GR_COLS) grad_contract_kernel(const DevSpec spec, uint64_t diffmask,
                                                                const double* __restrict__ Xtr,
                                                                const double* __restrict__ alpha, int n,
                                                                const double* __restrict__ xk, int ncols,
                                                                int ncols_pad, const double* __restrict__ W,
                                                                const double* __restrict__ dmu,
                                                                const double* __restrict__ dvar,
                                                                double* __restrict__ S_part, int jrows) {
  extern __shared__ double sm[];
  const int d_k = spec.d_k;
  double* xs = sm;                                   // [d_k][GR_COLS]
  double* accs = xs + d_k * GR_COLS;                 // [d_k][GR_COLS]
  double* Xs = accs + d_k * GR_COLS;                 // [jrows][d_k]
  double* als = Xs + jrows * d_k;                    // [jrows]
  const int tid = threadIdx.x;
  const int col = blockIdx.x * GR_COLS + tid;
  const int j0 = blockIdx.y * jrows;
  const int jn = max(0, min(jrows, n - j0));
  const bool cvalid = col < ncols;
  for (int dd = 0; dd < d_k; ++dd) {
    xs[dd * GR_COLS + tid] = cvalid ? xk[size_t(col) * d_k + dd] : 0.0;
    accs[dd * GR_COLS + tid] = 0.0;
  }
  for (int idx = tid; idx < jn * d_k; idx += GR_COLS) Xs[idx] = Xtr[size_t(j0) * d_k + idx];
  for (int idx = tid; idx < jn; idx += GR_COLS) als[idx] = alpha[j0 + idx];
  __syncthreads();
  if (cvalid) {
    const double cm = dmu[col];
    const double cv = -2.0 * dvar[col];
    const double* x = xs + tid;
    double* ac = accs + tid;
    for (int jj = 0; jj < jn; ++jj) {
      const double Cj = fma(cv, W[size_t(j0 + jj) * ncols_pad + col], cm * als[jj]);
      const double* Xj = Xs + jj * d_k;
      for (int t = 0; t < spec.n_terms; ++t) {
        const DevTerm& T = spec.t[t];
        double Fv[PRB_MAX_FACTORS], Gv[PRB_MAX_FACTORS];
#pragma unroll
        for (int f = 0; f < PRB_MAX_FACTORS; ++f) {
          Fv[f] = 1.0;
          Gv[f] = 0.0;
          if (f < T.n_factors) {
            const DevFactor& F = T.f[f];
            if (F.kind == PRB_FACTOR_MATERN52) {
              double r2 = 0.0;
              for (int di = 0; di < F.n_dims; ++di) {
                const int dim = F.dims[di];
                const double df = (x[dim * GR_COLS] - Xj[dim]) * F.inv_ls[di];
                r2 = fma(df, df, r2);
              }
              Fv[f] = matern52(r2, &Gv[f]);
            } else {
              double s = 0.0;
              for (int di = 0; di < F.n_dims; ++di) {
                const int dim = F.dims[di];
                s += (x[dim * GR_COLS] != Xj[dim]) ? F.inv_ls[di] : 0.0;
              }
              Fv[f] = exp(-s);
            }
          }
        }
#pragma unroll
        for (int f = 0; f < PRB_MAX_FACTORS; ++f) {
          if (f < T.n_factors && T.f[f].kind == PRB_FACTOR_MATERN52) {
            const DevFactor& F = T.f[f];
            // Q = outputscale * d(term)/dF_f * G_f * C_j
            double q = T.outputscale * Gv[f] * Cj;
            if (!T.additive) {
              if (F.power > 1) q *= double(F.power) * ipow(Fv[f], F.power - 1);
#pragma unroll
              for (int g = 0; g < PRB_MAX_FACTORS; ++g)
                if (g != f && g < T.n_factors) q *= ipow(Fv[g], T.f[g].power);
            } else if (F.power > 1) {
              q *= double(F.power) * ipow(Fv[f], F.power - 1);
            }
            for (int di = 0; di < F.n_dims; ++di) {
              const int dim = F.dims[di];
              if ((diffmask >> dim) & 1u) {
                const double w = F.inv_ls[di] * F.inv_ls[di];
                ac[dim * GR_COLS] = fma(q * w, x[dim * GR_COLS] - Xj[dim], ac[dim * GR_COLS]);
              }
            }
          }
        }
      }
    }
  }
  for (int dd = 0; dd < d_k; ++dd)
    S_part[(size_t(blockIdx.y) * d_k + dd) * ncols_pad + col] = accs[dd * GR_COLS + tid];
}

cudaError_t launch_grad_contract(const Model& m, uint64_t diffmask, const double* xk, int ncols, int ncols_pad,
                                 const double* W, const double* dmu, const double* dvar, double* S_part, int jrows,
                                 cudaStream_t st) {
  ProfScope prof_(K_GRAD_CONTRACT, st);
  dim3 grid(ncols_pad / GR_COLS, m.Mp2 / jrows);
  const size_t smem = (size_t(m.d_k) * (2 * GR_COLS + jrows) + jrows) * sizeof(double);
  if (smem > 48 * 1024) {  // d_k > 15: opt in to a larger dynamic shared-memory window (at most 99 KB at d_k = 32)
    cudaError_t e = cudaFuncSetAttribute(grad_contract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
  }
  grad_contract_kernel<<<grid, GR_COLS, smem, st>>>(m.spec, diffmask, m.Xtr, m.alpha, m.n, xk, ncols, ncols_pad, W,
                                                    dmu, dvar, S_part, jrows);
  count_launch();
  return cudaGetLastError();
}

// S_all[col_global, dim] = sum_jblock S_part[jblock][dim][col]   (fixed order)
__global__ void sum_splits_kernel(const double* __restrict__ S_part, int jblocks, const DimMap dm, int d_k, int ncols,
                                  int ncols_pad, double* __restrict__ S_all, const int* __restrict__ ncols_dev,
                                  int64_t col0) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  if (ncols_dev != nullptr && col0 + col >= int64_t(*ncols_dev)) return;
  for (int dd = 0; dd < dm.n; ++dd) {
    double s = 0.0;
    for (int jb = 0; jb < jblocks; ++jb) s += S_part[(size_t(jb) * dm.n + dd) * ncols_pad + col];
    S_all[size_t(col) * d_k + dm.dims[dd]] = s;
  }
}

cudaError_t launch_sum_splits(const Model& m, const double* S_part, int jblocks, const DimMap* dm, int ncols,
                              int ncols_pad, double* S_all, const int* ncols_dev, int64_t col0, cudaStream_t st) {
  ProfScope prof_(K_SUM_SPLITS, st);
  if (ncols == 0) return cudaSuccess;
  DimMap ident;
  ident.n = m.d_k;
  for (int i = 0; i < PRB_MAX_DIMS; ++i) ident.dims[i] = i;
  sum_splits_kernel<<<(ncols + 127) / 128, 128, 0, st>>>(S_part, jblocks, dm ? *dm : ident, m.d_k,
                                                         ncols, ncols_pad, S_all, ncols_dev, col0);
  count_launch();
  return cudaGetLastError();
}

// =====================================================================================================================
// Block reduction helper: fixed-order (lane tree, then warp 0 sums warp partials sequentially) => deterministic.
// =====================================================================================================================
__device__ __forceinline__ double block_sum(double v, double* red /*[32]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  const int nw = (blockDim.x + 31) >> 5;
  for (int w = 0; w < nw; ++w) s += red[w];
  return s;
}

// =====================================================================================================================
// MC de-duplication (separable path, PRB_OPT_DEDUP).  All MC samples of one restart share the continuous columns, so two
// samples with the same binary code are the SAME evaluation point: the base acquisition value is computed once per unique
// (restart, code) and weighted by its multiplicity in the reductions.  Exact up to the order of floating-point summation;
// bench.py reports it separately (evals/s is always counted per candidate x MC sample).
//   dedup_count   : one CTA per restart, shared-memory histogram over the 2^nbits codes, compaction in code order
//   dedup_offsets : exclusive scan of the unique counts over restarts (+ total, device-resident: no host round trip)
//   dedup_pack    : unique columns packed restart by restart: code, owning restart, weight
// =====================================================================================================================
__global__ void __launch_bounds__(256) dedup_count_kernel(const uint32_t* __restrict__ zbits, int mc, int nbits,
                                                          int* __restrict__ ucount, uint32_t* __restrict__ ucode,
                                                          int* __restrict__ uweight, int* __restrict__ inv) {
  extern __shared__ int dsm[];
  const int nbins = 1 << nbits;
  int* hist = dsm;           // [nbins]
  int* slot = dsm + nbins;   // [nbins]
  __shared__ int wsum[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const uint32_t* zb = zbits + size_t(b) * mc;
  for (int i = tid; i < nbins; i += 256) hist[i] = 0;
  __syncthreads();
  for (int i = tid; i < mc; i += 256) atomicAdd(&hist[zb[i]], 1);
  __syncthreads();
  // slots in increasing code order: each thread owns a contiguous range of bins
  const int per = (nbins + 255) / 256;
  const int lo = tid * per, hi = min(nbins, lo + per);
  int cnt = 0;
  for (int i = lo; i < hi; ++i) cnt += hist[i] > 0;
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if ((tid & 31) >= o) incl += v;
  }
  if ((tid & 31) == 31) wsum[tid >> 5] = incl;
  __syncthreads();
  int base = 0;
  for (int w = 0; w < (tid >> 5); ++w) base += wsum[w];
  int pos = base + incl - cnt;
  for (int i = lo; i < hi; ++i) {
    if (hist[i] > 0) {
      slot[i] = pos;
      ucode[size_t(b) * mc + pos] = uint32_t(i);
      uweight[size_t(b) * mc + pos] = hist[i];
      ++pos;
    }
  }
  if (tid == 255) ucount[b] = base + incl;
  __syncthreads();
  if (inv)
    for (int i = tid; i < mc; i += 256) inv[size_t(b) * mc + i] = slot[zb[i]];
}

__global__ void dedup_offsets_kernel(const int* __restrict__ ucount, int nb, int* __restrict__ ustart,
                                     int* __restrict__ total) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int acc = 0;
    for (int b = 0; b < nb; ++b) {
      ustart[b] = acc;
      acc += ucount[b];
    }
    ustart[nb] = acc;
    total[0] = acc;
    total[1] = (acc + 127) / 128 * 128;
  }
}

__global__ void __launch_bounds__(256) dedup_pack_kernel(const int* __restrict__ ucount, const int* __restrict__ ustart,
                                                         const uint32_t* __restrict__ ucode,
                                                         const int* __restrict__ uweight, int mc, int nb,
                                                         const int* __restrict__ total, uint32_t* __restrict__ zbits_u,
                                                         int* __restrict__ colb, double* __restrict__ w_u) {
  const int b = blockIdx.x;
  const int n_u = ucount[b], s0 = ustart[b];
  for (int s = threadIdx.x; s < n_u; s += 256) {
    zbits_u[s0 + s] = ucode[size_t(b) * mc + s];
    colb[s0 + s] = b;
    w_u[s0 + s] = double(uweight[size_t(b) * mc + s]);
  }
  if (b == nb - 1) {  // pad the last column tile with harmless entries
    for (int s = total[0] + threadIdx.x; s < total[1]; s += 256) {
      zbits_u[s] = 0;
      colb[s] = nb - 1;
      w_u[s] = 0.0;
    }
  }
}

cudaError_t launch_dedup(const uint32_t* zbits, int nb, int mc, int nbits, int* ucount, int* ustart, int* total,
                         uint32_t* ucode, int* uweight, int* inv, uint32_t* zbits_u, int* colb, double* w_u,
                         cudaStream_t st) {
  ProfScope prof_(K_SAMPLE, st);
  if (nb == 0) return cudaSuccess;
  const size_t smem = size_t(2) * (size_t(1) << nbits) * sizeof(int);
  dedup_count_kernel<<<nb, 256, smem, st>>>(zbits, mc, nbits, ucount, ucode, uweight, inv);
  dedup_offsets_kernel<<<1, 32, 0, st>>>(ucount, nb, ustart, total);
  dedup_pack_kernel<<<nb, 256, 0, st>>>(ucount, ustart, ucode, uweight, mc, nb, total, zbits_u, colb, w_u);
  count_launch(3);
  return cudaGetLastError();
}

// Weighted MC reduction over the unique columns of each restart (one CTA per restart); same quantities as mc_reduce_kernel.
__global__ void __launch_bounds__(128) mc_reduce_dedup_kernel(const DevSpace sp, const BitMap bm, int mc,
                                                              const int* __restrict__ ucount,
                                                              const int* __restrict__ ustart,
                                                              const double* __restrict__ w_u,
                                                              const uint32_t* __restrict__ zbits_u,
                                                              const double* __restrict__ a,
                                                              const double* __restrict__ prob,
                                                              const double* __restrict__ S_all, int estimator,
                                                              const double* __restrict__ ma_state,
                                                              const double* __restrict__ grad_out,
                                                              double* __restrict__ out_value,
                                                              double* __restrict__ out_grad) {
  __shared__ double red[32];
  const int bi = blockIdx.x, tid = threadIdx.x;
  const int n_u = ucount[bi];
  const size_t c0 = size_t(ustart[bi]);
  double s = 0.0;
  for (int i = tid; i < n_u; i += blockDim.x) s = fma(w_u[c0 + i], a[c0 + i], s);
  s = block_sum(s, red);
  if (tid == 0) out_value[bi] = s / double(mc);
  if (!grad_out) return;
  double baseline = 0.0;
  if (estimator == PRB_EST_REINFORCE_MA) {
    const double cnt = ma_state[0];
    if (cnt != 0.0) baseline = ma_state[1] / (1.0 - pow(0.7, cnt));
  }
  const double go = grad_out[bi];
  double* g = out_grad + size_t(bi) * sp.d;
  for (int ci = 0; ci < sp.n_cont; ++ci) {
    const int c = sp.cont_cols[ci];
    double t = 0.0;
    for (int i = tid; i < n_u; i += blockDim.x) t = fma(w_u[c0 + i], S_all[(c0 + i) * sp.d_k + c], t);
    t = block_sum(t, red);
    if (tid == 0) g[c] = go * (t / double(mc));
  }
  for (int kd = 0; kd < sp.n_int; ++kd) {
    const double p = prob[size_t(bi) * sp.n_disc + kd];
    const int bit = bm.pos[kd];
    double t = 0.0;
    for (int i = tid; i < n_u; i += blockDim.x) {
      const double zi = double((zbits_u[c0 + i] >> bit) & 1u);
      t += w_u[c0 + i] * (((a[c0 + i] - baseline) * (zi - p)) / sp.tau);
    }
    t = block_sum(t, red);
    if (tid == 0) g[sp.int_cols[kd]] = go * (t / double(mc));
  }
}

cudaError_t launch_mc_reduce_dedup(const DevSpace& sp, const BitMap& bm, int b, int mc, const int* ucount,
                                   const int* ustart, const double* w_u, const uint32_t* zbits_u, const double* a,
                                   const double* prob, const double* S_all, int estimator, const double* ma_state,
                                   const double* grad_out, double* out_value, double* out_grad, cudaStream_t st) {
  ProfScope prof_(K_MC_REDUCE, st);
  if (b == 0) return cudaSuccess;
  mc_reduce_dedup_kernel<<<b, 128, 0, st>>>(sp, bm, mc, ucount, ustart, w_u, zbits_u, a, prob, S_all, estimator,
                                            ma_state, grad_out, out_value, out_grad);
  count_launch();
  return cudaGetLastError();
}

// per-sample acquisition values from the unique ones: out[b, i] = a[ustart[b] + inv[b, i]]
__global__ void dedup_scatter_kernel(const double* __restrict__ a, const int* __restrict__ ustart,
                                     const int* __restrict__ inv, int mc, int64_t total, double* __restrict__ out) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  out[idx] = a[ustart[idx / mc] + inv[idx]];
}
cudaError_t launch_dedup_scatter(const double* a, const int* ustart, const int* inv, int b, int mc, double* out,
                                 cudaStream_t st) {
  ProfScope prof_(K_MC_REDUCE, st);
  const int64_t total = int64_t(b) * mc;
  if (total == 0) return cudaSuccess;
  dedup_scatter_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(a, ustart, inv, mc, total, out);
  count_launch();
  return cudaGetLastError();
}

// =====================================================================================================================
// MC reduction: one CTA per restart.
//   value_b = mean_i a_i                                                  (probabilistic_reparameterization.py:495-498)
//   discrete columns: grad_out_b * mean_i[ (a_i - baseline)(z_i - p) / tau ]        (:598-622)
//   continuous columns: grad_out_b * mean_i S[i, col]                                (:624-633)
//   baseline from the PRE-update EMA state (:600-609).
// =====================================================================================================================
constexpr int REDUCE_THREADS = 512;
__global__ void __launch_bounds__(REDUCE_THREADS) mc_reduce_kernel(const DevSpace sp, int mc, const double* __restrict__ a,
                                                                   const uint8_t* __restrict__ z,
                                                                   const double* __restrict__ prob,
                                                                   const double* __restrict__ S_all, int estimator,
                                                                   const double* __restrict__ ma_state,
                                                                   const double* __restrict__ grad_out,
                                                                   double* __restrict__ out_value,
                                                                   double* __restrict__ out_grad) {
  // one CTA per restart, one WARP per output (value, each continuous column, each discrete column): lane-strided partial
  // sums + xor tree, fixed order => deterministic; no block-wide barrier (thirteen sequential block reductions cost 17 us
  // at experiment sizes, this form ~4 us)
  const int bi = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t c0 = size_t(bi) * mc;
  double baseline = 0.0;
  if (grad_out && estimator == PRB_EST_REINFORCE_MA) {
    const double cnt = ma_state[0];
    if (cnt != 0.0) baseline = ma_state[1] / (1.0 - pow(0.7, cnt));
  }
  const double go = grad_out ? grad_out[bi] : 0.0;
  double* g = grad_out ? out_grad + size_t(bi) * sp.d : nullptr;
  const int cat0 = sp.n_cat > 0 ? sp.cat_start[0] : 0;
  const int n_tasks = grad_out ? 1 + sp.n_cont + sp.n_disc : 1;
  for (int task = warp; task < n_tasks; task += REDUCE_THREADS / 32) {
    double t = 0.0;
    if (task == 0) {
      for (int i = lane; i < mc; i += 32) t += a[c0 + i];
      t = warp_sum(t);
      if (lane == 0) out_value[bi] = t / double(mc);
    } else if (task <= sp.n_cont) {
      const int c = sp.cont_cols[task - 1];
      for (int i = lane; i < mc; i += 32) t += S_all[(c0 + i) * sp.d_k + c];
      t = warp_sum(t);
      if (lane == 0) g[c] = go * (t / double(mc));
    } else {
      // discrete columns: integers first, then the one-hot columns (the order of `discrete_indices`, input.py:574-597)
      const int kd = task - 1 - sp.n_cont;
      const double p = prob[size_t(bi) * sp.n_disc + kd];
      for (int i = lane; i < mc; i += 32) {
        const double zi = double(z[(c0 + i) * sp.n_disc + kd]);
        t += ((a[c0 + i] - baseline) * (zi - p)) / sp.tau;
      }
      t = warp_sum(t);
      if (lane == 0) {
        const int c = kd < sp.n_int ? sp.int_cols[kd] : cat0 + (kd - sp.n_int);
        g[c] = go * (t / double(mc));
      }
    }
  }
}

cudaError_t launch_mc_reduce(const DevSpace& sp, int b, int mc, const double* a, const uint8_t* z, const double* prob,
                             const double* S_all, int estimator, const double* ma_state, const double* grad_out,
                             double* out_value, double* out_grad, cudaStream_t st) {
  ProfScope prof_(K_MC_REDUCE, st);
  if (b == 0) return cudaSuccess;
  mc_reduce_kernel<<<b, REDUCE_THREADS, 0, st>>>(sp, mc, a, z, prob, S_all, estimator, ma_state, grad_out, out_value,
                                                 out_grad);
  count_launch();
  return cudaGetLastError();
}

// EMA baseline update (probabilistic_reparameterization.py:499-505): counter += 1; hidden -= (hidden - mean) * 0.3,
// mean over ALL b*mc acquisition values of the call (= mean of the per-restart means, equal mc).
__global__ void ma_update_kernel(const double* __restrict__ value, int b, double* __restrict__ ma_state) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < b; i += blockDim.x) s += value[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    const double mean = s / double(b);
    const double hidden = ma_state[1];
    ma_state[0] += 1.0;
    ma_state[1] = hidden - (hidden - mean) * (1.0 - 0.7);
  }
}

cudaError_t launch_ma_update(const double* value, int b, double* ma_state, cudaStream_t st) {
  ProfScope prof_(K_MA_UPDATE, st);
  ma_update_kernel<<<1, 128, 0, st>>>(value, b, ma_state);
  count_launch();
  return cudaGetLastError();
}

// =====================================================================================================================
// Analytic reduction: one CTA per restart.
//   value_b = sum_o P_o a_o ;  d/d theta_cont = sum_o P_o S[o, col] ;  d/d theta_disc = sum_o a_o dP_o/d theta
//   (autograd through get_probs, input.py:410-464, including the Normalize chain-rule factor `range` for integers).
// =====================================================================================================================
__global__ void __launch_bounds__(REDUCE_THREADS) analytic_reduce_kernel(
    const DevSpace sp, const double* __restrict__ theta, int n_opt, const int32_t* __restrict__ options,
    const double* __restrict__ a, const double* __restrict__ probs, const double* __restrict__ S_all,
    const double* __restrict__ grad_out, double* __restrict__ out_value, double* __restrict__ out_grad,
    const double* __restrict__ S_part, int jblocks, int ncols_pad, double* __restrict__ adam_theta,
    double* __restrict__ adam_m, double* __restrict__ adam_v, double lr, int* __restrict__ step_ptr,
    unsigned int* __restrict__ done) {
  // one CTA per restart; per-restart factor tables first, then one WARP per output (value, continuous columns, integer
  // columns, one-hot columns): lane-strided sums over the configurations + xor tree, fixed order => deterministic
  __shared__ double p_int[PRB_MAX_DIMS], dp_int[PRB_MAX_DIMS], off_int[PRB_MAX_DIMS];
  __shared__ double sm_cat[PRB_MAX_CAT * PRB_MAX_CARD];
  __shared__ double csum[PRB_MAX_DIMS][4];
  __shared__ int last_s;
  const int bi = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t c0 = size_t(bi) * n_opt;
  const double* th = theta + size_t(bi) * sp.d;
  // fused step (S_part != NULL): the row-block partial sums of the gradient contraction are added up here, split over
  // nsplit warps per continuous column; the pre-step optimiser state is read before this CTA signals completion
  const int nsplit = (S_part != nullptr && jblocks >= 8) ? 4 : 1;
  const int jb_per = (jblocks + nsplit - 1) / nsplit;
  const double t_step = step_ptr ? double(*step_ptr + 1) : 1.0;
  if (grad_out) {
    if (tid < sp.n_int) {
      const double lo = sp.int_lo[tid], hi = sp.int_hi[tid];
      const double x_un = unnormalize_int(th[sp.int_cols[tid]], lo, hi);
      double off;
      const double p = int_round_prob(x_un, sp.tau, &off);
      const double sgn = (x_un > 0.0) ? 1.0 : ((x_un < 0.0) ? -1.0 : 0.0);
      p_int[tid] = p;
      off_int[tid] = off;
      dp_int[tid] = p * (1.0 - p) / sp.tau * sgn * (hi - lo);
    }
    if (tid >= 64 && tid - 64 < sp.n_cat) {
      const int j = tid - 64;
      const int st = sp.cat_start[j], card = sp.cat_card[j];
      double mx = -1e300;
      for (int k = 0; k < card; ++k) mx = fmax(mx, __dsub_rn(th[st + k], 0.5) / sp.tau);
      double sum = 0.0;
      for (int k = 0; k < card; ++k) {
        const double e = exp(__dsub_rn(__dsub_rn(th[st + k], 0.5) / sp.tau, mx));
        sm_cat[j * PRB_MAX_CARD + k] = e;
        sum += e;
      }
      const double inv = 1.0 / sum;
      for (int k = 0; k < card; ++k) sm_cat[j * PRB_MAX_CARD + k] *= inv;
    }
  }
  __syncthreads();
  const double go = grad_out ? grad_out[bi] : 0.0;
  double* g = grad_out ? out_grad + size_t(bi) * sp.d : nullptr;
  const int nv = sp.n_int + sp.n_cat;
  int n_onehot = 0;
  for (int j = 0; j < sp.n_cat; ++j) n_onehot += sp.cat_card[j];
  const int n_cont_tasks = sp.n_cont * nsplit;
  const int n_tasks = grad_out ? 1 + n_cont_tasks + sp.n_int + n_onehot : 1;
  for (int task = warp; task < n_tasks; task += REDUCE_THREADS / 32) {
    double t = 0.0;
    if (task == 0) {
      for (int o = lane; o < n_opt; o += 32) t += a[c0 + o] * probs[c0 + o];
      t = warp_sum(t);
      if (lane == 0) out_value[bi] = t;
    } else if (task <= n_cont_tasks) {
      const int ci = (task - 1) / nsplit, part = (task - 1) % nsplit;
      const int c = sp.cont_cols[ci];
      if (S_part != nullptr) {
        const int jb0 = part * jb_per, jb1 = min(jblocks, jb0 + jb_per);
        for (int o = lane; o < n_opt; o += 32) {
          double v = 0.0;
          for (int jb = jb0; jb < jb1; ++jb) v += S_part[(size_t(jb) * sp.d_k + c) * ncols_pad + c0 + o];
          t += probs[c0 + o] * v;
        }
      } else {
        for (int o = lane; o < n_opt; o += 32) t += probs[c0 + o] * S_all[(c0 + o) * sp.d_k + c];
      }
      t = warp_sum(t);
      if (lane == 0) csum[ci][part] = t;
    } else if (task <= n_cont_tasks + sp.n_int) {
      // integer column k: d P / d theta = (+-) dp_k * product of the other factors
      const int k = task - 1 - n_cont_tasks;
      for (int o = lane; o < n_opt; o += 32) {
        const int32_t* op = options + size_t(o) * nv;
        const double diff = (off_int[k] + 1.0) - double(op[k]);
        const double df = (diff == 0.0) ? dp_int[k] : ((diff == 1.0) ? -dp_int[k] : 0.0);
        if (df != 0.0) {
          double excl = 1.0;
          for (int k2 = 0; k2 < sp.n_int; ++k2) {
            if (k2 == k) continue;
            const double d2 = (off_int[k2] + 1.0) - double(op[k2]);
            excl *= (d2 == 0.0) ? p_int[k2] : ((d2 == 1.0) ? (1.0 - p_int[k2]) : 0.0);
          }
          for (int j = 0; j < sp.n_cat; ++j) excl *= sm_cat[j * PRB_MAX_CARD + op[sp.n_int + j]];
          t += a[c0 + o] * excl * df;
        }
      }
      t = warp_sum(t);
      if (lane == 0) g[sp.int_cols[k]] = go * t;
    } else {
      // one-hot column m of categorical j: d softmax_c / d theta_m = softmax_c (delta_cm - softmax_m) / tau
      int rem = task - 1 - n_cont_tasks - sp.n_int, j = 0;
      while (rem >= sp.cat_card[j]) {
        rem -= sp.cat_card[j];
        ++j;
      }
      const int mcol = rem;
      for (int o = lane; o < n_opt; o += 32) {
        const int32_t* op = options + size_t(o) * nv;
        const int c = op[sp.n_int + j];
        double excl = 1.0;
        for (int k2 = 0; k2 < sp.n_int; ++k2) {
          const double d2 = (off_int[k2] + 1.0) - double(op[k2]);
          excl *= (d2 == 0.0) ? p_int[k2] : ((d2 == 1.0) ? (1.0 - p_int[k2]) : 0.0);
        }
        for (int j2 = 0; j2 < sp.n_cat; ++j2)
          if (j2 != j) excl *= sm_cat[j2 * PRB_MAX_CARD + op[sp.n_int + j2]];
        const double smc = sm_cat[j * PRB_MAX_CARD + c];
        const double smm = sm_cat[j * PRB_MAX_CARD + mcol];
        t += a[c0 + o] * excl * smc * ((c == mcol ? 1.0 : 0.0) - smm) / sp.tau;
      }
      t = warp_sum(t);
      if (lane == 0) g[sp.cat_start[j] + mcol] = go * t;
    }
  }
  if (grad_out == nullptr) return;
  __syncthreads();
  if (tid < sp.n_cont) {
    double t = csum[tid][0];
    for (int part = 1; part < nsplit; ++part) t += csum[tid][part];
    g[sp.cont_cols[tid]] = go * t;
  }
  if (adam_theta == nullptr) return;
  __syncthreads();
  if (tid < sp.d) {  // torch.optim.Adam on loss = -acq(X).sum() (same arithmetic as adam_step_kernel)
    const size_t idx = size_t(bi) * sp.d + tid;
    const double gr = -g[tid];
    const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
    const double mi = adam_m[idx] + (gr - adam_m[idx]) * (1.0 - b1);
    const double vi = adam_v[idx] * b2 + (1.0 - b2) * gr * gr;
    adam_m[idx] = mi;
    adam_v[idx] = vi;
    const double bc1 = 1.0 - pow(b1, t_step);
    const double bc2 = 1.0 - pow(b2, t_step);
    adam_theta[idx] = adam_theta[idx] - (lr / bc1) * (mi / (sqrt(vi) / sqrt(bc2) + eps));
  }
  // the last CTA advances the step counter and re-arms the completion counter
  __threadfence();
  __syncthreads();
  if (tid == 0) last_s = (atomicAdd(done, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (last_s && tid == 0) {
    if (step_ptr) *step_ptr += 1;
    *done = 0u;
  }
}

cudaError_t launch_analytic_reduce(const DevSpace& sp, const double* theta, int b, int n_opt, const int32_t* options,
                                   const double* a, const double* probs, const double* S_all, const double* grad_out,
                                   double* out_value, double* out_grad, cudaStream_t st, const double* S_part,
                                   int jblocks, int ncols_pad, double* adam_theta, double* adam_m, double* adam_v,
                                   double lr, int* step_ptr, unsigned int* done) {
  ProfScope prof_(K_ANALYTIC, st);
  if (b == 0) return cudaSuccess;
  analytic_reduce_kernel<<<b, REDUCE_THREADS, 0, st>>>(sp, theta, n_opt, options, a, probs, S_all, grad_out, out_value,
                                                       out_grad, S_part, jblocks, ncols_pad, adam_theta, adam_m, adam_v,
                                                       lr, step_ptr, done);
  count_launch();
  return cudaGetLastError();
}

// =====================================================================================================================
// Fused optimisation step for the separable path (prb_mc_adam): at experiment sizes (n <= 220, 640 evaluations per step)
// the step is launch-latency bound, so the eleven launches of the unfused sequence are merged into four:
//   sep_step_prep    clamp(theta) + per-restart kernel factors (kc, Dc) + sampler          (one CTA per restart)
//   lower product    with the acquisition epilogue fused when there is a single row tile   (prb_gemm_tma.cu)
//   upper product    with the gradient contraction fused                                   (prb_gemm_tma.cu)
//   sep_step_finish  split sums + MC reduction + Adam update (one CTA per restart); the last CTA to finish advances the EMA
//                    baseline and the step counter
// =====================================================================================================================
constexpr int PREP_THREADS = 128;
// grid (nb, nsplit): CTA (b, s) owns the s-th slice of the training points (kc, Dc) and of the MC samples (sampler) of
// restart b, so that a call with few restarts (8 per GPU in the 8-GPU split) still spreads over the SMs: one CTA per restart
// measured 41-47 us at n = 1000, mc = 1024 -- 7 % of a 0.6 ms step.  The rounding probabilities are recomputed per CTA (a
// dozen sigmoids).
__global__ void __launch_bounds__(PREP_THREADS) sep_step_prep_kernel(
    const DevSpace sp, const BitMap bm, const DevFactor fc, double outputscale, const double* __restrict__ Xtr, int n,
    int d_k, int Kp, const double* __restrict__ theta, const double* __restrict__ lower,
    const double* __restrict__ upper, double* __restrict__ Xc, int mc, const double* __restrict__ u_int,
    const double* __restrict__ u_cat, uint64_t seed, uint64_t offset, double* __restrict__ kc, double* __restrict__ Dc,
    uint8_t* __restrict__ z, double* __restrict__ prob, uint32_t* __restrict__ zbits) {
  __shared__ double th_s[PRB_MAX_DIMS];
  __shared__ double ps_s[PRB_MAX_DIMS], offs_s[PRB_MAX_DIMS];  // the separable path has no categoricals: n_disc = n_int
  grid_dep_launch();
  const int b = blockIdx.x, tid = threadIdx.x;
  const int split = blockIdx.y, nsplit = gridDim.y;
  if (tid < sp.d) {
    double v = theta[size_t(b) * sp.d + tid];
    if (lower) v = fmin(fmax(v, lower[tid]), upper[tid]);  // columnwise_clamp (gen.py:304-305, 323)
    th_s[tid] = v;
    if (Xc && split == 0) Xc[size_t(b) * sp.d + tid] = v;
  }
  __syncthreads();
  pr_restart_probs(sp, th_s, tid, PREP_THREADS, ps_s, offs_s);
  __syncthreads();
  if (prob && split == 0)
    for (int k = tid; k < sp.n_disc; k += PREP_THREADS) prob[size_t(b) * sp.n_disc + k] = ps_s[k];
  const int jper = (Kp + nsplit - 1) / nsplit;
  const int j_end = min(Kp, (split + 1) * jper);
  for (int j = split * jper + tid; j < j_end; j += PREP_THREADS) {
    double val = 0.0, G = 0.0;
    if (j < n) {
      double r2 = 0.0;
      for (int di = 0; di < fc.n_dims; ++di) {
        const int dim = fc.dims[di];
        const double df = (th_s[dim] - Xtr[size_t(j) * d_k + dim]) * fc.inv_ls[di];
        r2 = fma(df, df, r2);
      }
      val = outputscale * matern52(r2, &G);
    }
    kc[size_t(b) * Kp + j] = val;
    if (Dc) {
      for (int di = 0; di < fc.n_dims; ++di) {
        const int dim = fc.dims[di];
        double v = 0.0;
        if (j < n) v = outputscale * G * (fc.inv_ls[di] * fc.inv_ls[di]) * (th_s[dim] - Xtr[size_t(j) * d_k + dim]);
        Dc[(size_t(b) * fc.n_dims + di) * Kp + j] = v;
      }
    }
  }
  const int iper = (mc + nsplit - 1) / nsplit;
  const int i_end = min(mc, (split + 1) * iper);
  for (int i = split * iper + tid; i < i_end; i += PREP_THREADS) {
    const size_t col = size_t(b) * mc + i;
    zbits[col] = pr_sample_point(sp, th_s, ps_s, offs_s, i, u_int, u_cat, seed, offset, nullptr, nullptr,
                                 z ? z + col * sp.n_disc : nullptr, bm);
  }
}

cudaError_t launch_sep_step_prep(const Model& m, const DevSpace& sp, const BitMap& bm, const double* theta,
                                 const double* lower, const double* upper, double* Xc, int nb, int mc,
                                 const double* u_int, const double* u_cat, uint64_t seed, uint64_t offset, double* kc,
                                 double* Dc, uint8_t* z, double* prob, uint32_t* zbits, cudaStream_t st) {
  ProfScope prof_(K_SAMPLE, st);
  if (nb == 0) return cudaSuccess;
  // ~4 CTAs per SM in total, but no slice thinner than one item per thread
  const int work = (m.Kp > mc ? m.Kp : mc);
  int nsplit = (4 * 148 + nb - 1) / nb;
  const int max_split = (work + PREP_THREADS - 1) / PREP_THREADS;
  if (nsplit > max_split) nsplit = max_split;
  if (nsplit < 1) nsplit = 1;
  sep_step_prep_kernel<<<dim3(nb, nsplit), PREP_THREADS, 0, st>>>(sp, bm, m.spec.t[0].f[m.sep_cont_factor],
                                                                 m.spec.t[0].outputscale, m.Xtr, m.n, m.d_k, m.Kp, theta,
                                                                 lower, upper, Xc, mc, u_int, u_cat, seed, offset, kc, Dc,
                                                                 z, prob, zbits);
  count_launch();
  return cudaGetLastError();
}

constexpr int FINISH_THREADS = 1024;
constexpr int FIN_ILP = 8;  // independent partial sums per lane in the finish kernel's reductions
// One CTA per restart, one WARP per reduction (value, each continuous column, each discrete column): no block-wide
// barrier inside the reductions -- at experiment sizes the step is latency bound and thirteen sequential block reductions
// cost 17 us; this form costs ~3 us.  Fixed lane-strided order + xor-tree => deterministic.
__global__ void __launch_bounds__(FINISH_THREADS) sep_step_finish_kernel(
    const DevSpace sp, const DimMap dm, int slots, int slot_is_dim, int mc, const double* __restrict__ a,
    const uint8_t* __restrict__ z, const double* __restrict__ prob, const double* __restrict__ S_part, int mtiles,
    int ncols_pad, int estimator,
    double* __restrict__ ma_state, int update_ma, const double* __restrict__ grad_out, double* __restrict__ out_value,
    double* __restrict__ out_grad, int need_grad, double* __restrict__ theta, double* __restrict__ adam_m,
    double* __restrict__ adam_v, double lr, int* __restrict__ step_ptr, unsigned int* __restrict__ done, int nb) {
  __shared__ double red[32];
  __shared__ double g_s[PRB_MAX_DIMS];
  __shared__ int last_s;
  grid_dep_launch();
  grid_dep_wait();
  const int bi = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t c0 = size_t(bi) * mc;
  // Everything that depends on the pre-call optimiser / EMA state is read before this CTA signals completion.  The three
  // pow() calls (EMA bias correction of the baseline, Adam's two bias corrections) are long dependent chains (microseconds
  // each): ONE thread of the last warp evaluates them while the other warps run the reductions; the baseline enters the
  // score-function sums only afterwards (sum_i (a_i - beta)(z_i - p) = sum_i a_i (z_i - p) - beta sum_i (z_i - p)).
  __shared__ double pre_s[3];                                  // baseline, 1 - 0.9^t, 1 - 0.999^t
  __shared__ double s1_s[PRB_MAX_DIMS], s2_s[PRB_MAX_DIMS];    // per discrete column: sum a (z - p), sum (z - p)
  if (warp == FINISH_THREADS / 32 - 1 && lane == 0) {
    double baseline = 0.0;
    if (estimator == PRB_EST_REINFORCE_MA && ma_state != nullptr) {
      const double cnt = ma_state[0];
      if (cnt != 0.0) baseline = ma_state[1] / (1.0 - pow(0.7, cnt));
    }
    pre_s[0] = baseline;
    const double t_step = step_ptr ? double(*step_ptr + 1) : 1.0;
    pre_s[1] = theta != nullptr ? 1.0 - pow(0.9, t_step) : 1.0;
    pre_s[2] = theta != nullptr ? 1.0 - pow(0.999, t_step) : 1.0;
  }
  const double go = (need_grad && grad_out) ? grad_out[bi] : 1.0;
  // S_part is [mtiles][slots][ncols_pad]: the separable path stores one slot per continuous dim (slot = index in dm), the
  // generic gradient contraction one slot per kernel-input column (slot = the column itself).  With many row blocks (the
  // small-panel generic contraction uses 4-row blocks) each continuous column is split over NSPLIT warps by row-block range.
  __shared__ double csum[PRB_MAX_DIMS][4];
  const int nsplit = mtiles >= 8 ? 4 : 1;
  const int mt_per = (mtiles + nsplit - 1) / nsplit;
  const int cat0 = sp.n_cat > 0 ? sp.cat_start[0] : 0;
  const int n_cont_tasks = dm.n * nsplit;
  const int n_tasks = need_grad ? 1 + n_cont_tasks + sp.n_disc : 1;
  for (int task = warp; task < n_tasks; task += FINISH_THREADS / 32) {
    double t = 0.0;
    // lane-strided sums with FIN_ILP independent chains per lane (the loads of one trip are all in flight together: a
    // single dependent chain of mc / 32 global loads cost ~1 us each and made this kernel 37 us); chains are combined in a
    // fixed order, then the xor tree => deterministic
    double tc[FIN_ILP];
#pragma unroll
    for (int q = 0; q < FIN_ILP; ++q) tc[q] = 0.0;
    if (task == 0) {
      for (int i0 = lane; i0 < mc; i0 += 32 * FIN_ILP) {
#pragma unroll
        for (int q = 0; q < FIN_ILP; ++q) {
          const int i = i0 + 32 * q;
          if (i < mc) tc[q] += a[c0 + i];
        }
      }
#pragma unroll
      for (int q = 0; q < FIN_ILP; ++q) t += tc[q];
      t = warp_sum(t);
      if (lane == 0) out_value[bi] = t / double(mc);
    } else if (task <= n_cont_tasks) {
      const int ci = (task - 1) / nsplit, part = (task - 1) % nsplit;
      const int slot = slot_is_dim ? dm.dims[ci] : ci;
      const int mt0 = part * mt_per, mt1 = min(mtiles, mt0 + mt_per);
      for (int i0 = lane; i0 < mc; i0 += 32 * FIN_ILP) {
        for (int mt = mt0; mt < mt1; ++mt) {
          const double* row = S_part + (size_t(mt) * slots + slot) * ncols_pad + c0;
#pragma unroll
          for (int q = 0; q < FIN_ILP; ++q) {
            const int i = i0 + 32 * q;
            if (i < mc) tc[q] += row[i];
          }
        }
      }
#pragma unroll
      for (int q = 0; q < FIN_ILP; ++q) t += tc[q];
      t = warp_sum(t);
      if (lane == 0) csum[ci][part] = t;
    } else {
      // discrete columns: integers first, then the one-hot columns (the order of `discrete_indices`, input.py:574-597)
      const int kd = task - 1 - n_cont_tasks;
      const double p = prob[size_t(bi) * sp.n_disc + kd];
      double tz[FIN_ILP];
#pragma unroll
      for (int q = 0; q < FIN_ILP; ++q) tz[q] = 0.0;
      for (int i0 = lane; i0 < mc; i0 += 32 * FIN_ILP) {
#pragma unroll
        for (int q = 0; q < FIN_ILP; ++q) {
          const int i = i0 + 32 * q;
          if (i < mc) {
            const double zp = double(z[(c0 + i) * sp.n_disc + kd]) - p;
            tc[q] = fma(a[c0 + i], zp, tc[q]);
            tz[q] += zp;
          }
        }
      }
      double t2 = 0.0;
#pragma unroll
      for (int q = 0; q < FIN_ILP; ++q) {
        t += tc[q];
        t2 += tz[q];
      }
      t = warp_sum(t);
      t2 = warp_sum(t2);
      if (lane == 0) {
        s1_s[kd] = t;
        s2_s[kd] = t2;
      }
    }
  }
  __syncthreads();
  if (need_grad && tid < sp.n_disc) {
    const int kd = tid;
    const double sc = (s1_s[kd] - pre_s[0] * s2_s[kd]) / sp.tau;
    g_s[kd < sp.n_int ? sp.int_cols[kd] : cat0 + (kd - sp.n_int)] = go * (sc / double(mc));
  }
  if (need_grad && tid < dm.n) {
    double t = csum[tid][0];
    for (int part = 1; part < nsplit; ++part) t += csum[tid][part];
    g_s[dm.dims[tid]] = go * (t / double(mc));
  }
  __syncthreads();
  if (need_grad && tid < sp.d) {
    const double gacq = g_s[tid];
    if (out_grad) out_grad[size_t(bi) * sp.d + tid] = gacq;
    if (theta != nullptr) {  // torch.optim.Adam on loss = -acq(X).sum() (same arithmetic as adam_step_kernel)
      const size_t idx = size_t(bi) * sp.d + tid;
      const double g = -gacq;
      const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
      const double mi = adam_m[idx] + (g - adam_m[idx]) * (1.0 - b1);
      const double vi = adam_v[idx] * b2 + (1.0 - b2) * g * g;
      adam_m[idx] = mi;
      adam_v[idx] = vi;
      const double bc1 = pre_s[1], bc2 = pre_s[2];
      theta[idx] = theta[idx] - (lr / bc1) * (mi / (sqrt(vi) / sqrt(bc2) + eps));
    }
  }
  // last CTA to finish: EMA baseline update over ALL restarts of the call (probabilistic_reparameterization.py:499-505),
  // fixed summation order; step counter; re-arm the completion counter for the next launch
  __threadfence();
  __syncthreads();
  if (tid == 0) last_s = (atomicAdd(done, 1u) == unsigned(nb - 1)) ? 1 : 0;
  __syncthreads();
  if (last_s) {
    __threadfence();
    double v = 0.0;
    for (int i = tid; i < nb; i += blockDim.x) v += __ldcg(out_value + i);
    v = block_sum(v, red);
    if (tid == 0) {
      if (update_ma && ma_state != nullptr) {
        const double mean = v / double(nb);
        const double hidden = ma_state[1];
        ma_state[0] += 1.0;
        ma_state[1] = hidden - (hidden - mean) * (1.0 - 0.7);
      }
      if (step_ptr && theta != nullptr) *step_ptr += 1;
      *done = 0u;
    }
  }
}

cudaError_t launch_sep_step_finish(const DevSpace& sp, const DimMap& dm, int slots, int slot_is_dim, int nb, int mc,
                                   const double* a, const uint8_t* z,
                                   const double* prob, const double* S_part, int mtiles, int ncols_pad, int estimator,
                                   double* ma_state, int update_ma, const double* grad_out, double* out_value,
                                   double* out_grad, int need_grad, double* theta, double* adam_m, double* adam_v,
                                   double lr, int* step_ptr, unsigned int* done, cudaStream_t st) {
  ProfScope prof_(K_MC_REDUCE, st);
  if (nb == 0) return cudaSuccess;
  cudaError_t e = launch_kernel(sep_step_finish_kernel, dim3(nb), dim3(FINISH_THREADS), 0, st, sp, dm, slots, slot_is_dim,
                                mc, a, z, prob, S_part,
                                mtiles, ncols_pad, estimator, ma_state, update_ma, grad_out, out_value, out_grad, need_grad,
                                theta, adam_m, adam_v, lr, step_ptr, done, nb);
  count_launch();
  return e != cudaSuccess ? e : cudaGetLastError();
}

// =====================================================================================================================
// Multi-start Adam pieces (gen.py:304-337): X = clamp(theta) ; theta <- Adam(theta, grad = -dA/dX)
// =====================================================================================================================
__global__ void clamp_kernel(const double* __restrict__ theta, const double* __restrict__ lower,
                             const double* __restrict__ upper, int b, int d, double* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= b * d) return;
  const int c = idx % d;
  out[idx] = fmin(fmax(theta[idx], lower[c]), upper[c]);
}

cudaError_t launch_clamp(const double* theta, const double* lower, const double* upper, int b, int d, double* out,
                         cudaStream_t st) {
  ProfScope prof_(K_ADAM, st);
  clamp_kernel<<<(b * d + 255) / 256, 256, 0, st>>>(theta, lower, upper, b, d, out);
  count_launch();
  return cudaGetLastError();
}

// torch.optim.Adam (single-tensor path, no amsgrad / weight decay): exp_avg.lerp_(g, 1-b1);
// exp_avg_sq = b2*exp_avg_sq + (1-b2) g*g; denom = sqrt(exp_avg_sq)/sqrt(1-b2^t) + eps; theta -= lr/(1-b1^t) * exp_avg/denom
__global__ void adam_step_kernel(double* __restrict__ theta, const double* __restrict__ acq_grad, double* __restrict__ m,
                                 double* __restrict__ v, int total, double lr, const int* __restrict__ step_ptr) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const double t = double(*step_ptr);
  const double g = -acq_grad[idx];  // loss = -acq(X).sum()
  const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
  const double mi = m[idx] + (g - m[idx]) * (1.0 - b1);
  const double vi = v[idx] * b2 + (1.0 - b2) * g * g;
  m[idx] = mi;
  v[idx] = vi;
  const double bc1 = 1.0 - pow(b1, t);
  const double bc2 = 1.0 - pow(b2, t);
  const double step_size = lr / bc1;
  const double denom = sqrt(vi) / sqrt(bc2) + eps;
  theta[idx] = theta[idx] - step_size * (mi / denom);
}
__global__ void step_inc_kernel(int* step_ptr) { *step_ptr += 1; }

cudaError_t launch_adam_step(double* theta, const double* acq_grad, double* m, double* v, int total, double lr,
                             int* step_ptr, cudaStream_t st) {
  ProfScope prof_(K_ADAM, st);
  step_inc_kernel<<<1, 1, 0, st>>>(step_ptr);
  adam_step_kernel<<<(total + 255) / 256, 256, 0, st>>>(theta, acq_grad, m, v, total, lr, step_ptr);
  count_launch(2);
  return cudaGetLastError();
}

__global__ void fill_kernel(double* p, double v, int n) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n) p[idx] = v;
}
cudaError_t launch_fill(double* p, double v, int n, cudaStream_t st) {
  ProfScope prof_(K_MISC, st);
  if (n == 0) return cudaSuccess;
  fill_kernel<<<(n + 255) / 256, 256, 0, st>>>(p, v, n);
  count_launch();
  return cudaGetLastError();
}

// =====================================================================================================================
// FP64 pipe microbenchmark (diagnostic; bench.py runs it OUTSIDE its timed region to put the roofline denominator of the
// triangular products into the same record as the measurement).  Register-resident operands, 8 independent accumulator
// chains per warp, every SM busy with 2 CTAs of 8 warps.
//   mode 0: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4, the instruction the products use)      512 flop per warp instruction
//   mode 1: scalar DFMA                                                                   64 flop per warp instruction
//   mode 2: even warps DMMA, odd warps DFMA (do the two share one FP64 datapath?)          flops of both counted
// =====================================================================================================================
template <int MODE>
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* __restrict__ out, int iters, double a0) {
  const bool use_mma = MODE == 0 || (MODE == 2 && ((threadIdx.x >> 5) & 1) == 0);
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    c[i][0] = double(threadIdx.x);
    c[i][1] = double(i);
  }
  const double a = a0, b = 1.0 - a0;
  if (use_mma) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
    }
  } else {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        c[i][0] = fma(c[i][0], a, b);
        c[i][1] = fma(c[i][1], a, b);
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[size_t(blockIdx.x) * blockDim.x + threadIdx.x] = s;
}

// modes 3 / 4 / 5: the DMMA stream of the triangular products' inner loop -- 32 independent m8n8k4 per 4-deep k-chunk per
// warp -- with its operands fetched from shared memory the way the product kernel does: mode 3 = 12 LDS.64 per chunk (4 A + 8 B
// fragments), mode 4 = the same bytes as 6 LDS.128 (two chunks' fragments per load), mode 5 = 24 LDS.64 (a 32 x 32 warp tile's
// ratio).  Tells whether shared-memory loads take issue bandwidth away from the FP64 tensor pipe.
template <int LDS_MODE>
__global__ void __launch_bounds__(256) fp64_lds_mix_kernel(double* __restrict__ out, int iters) {
  __shared__ __align__(16) double buf[2][12][64];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 2 * 12 * 64; i += 256) (&buf[0][0][0])[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  double c[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) c[i][j][0] = c[i][j][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
    const int pb = it & 1;
    double a[4], b[8];
    if (LDS_MODE == 4) {
      if ((it & 1) == 0) {  // one LDS.128 brings the fragments of two chunks
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const double2 v = *reinterpret_cast<const double2*>(&buf[pb][q][2 * lane]);
          if (q < 2) { a[2 * q] = v.x; a[2 * q + 1] = v.y; } else { b[2 * (q - 2)] = v.x; b[2 * (q - 2) + 1] = v.y; }
        }
      } else {
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const double2 v = *reinterpret_cast<const double2*>(&buf[pb][6 + q][2 * lane]);
          if (q < 2) { a[2 * q] = v.x; a[2 * q + 1] = v.y; } else { b[2 * (q - 2)] = v.x; b[2 * (q - 2) + 1] = v.y; }
        }
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) a[q] = buf[pb][q][lane];
#pragma unroll
      for (int q = 0; q < 8; ++q) b[q] = buf[pb][4 + q][lane + (LDS_MODE == 5 ? 32 : 0)];
      if (LDS_MODE == 5) {
#pragma unroll
        for (int q = 0; q < 4; ++q) a[q] += buf[pb ^ 1][q][lane];
#pragma unroll
        for (int q = 0; q < 8; ++q) b[q] += buf[pb ^ 1][4 + q][lane];
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) s += c[i][j][0] + c[i][j][1];
  out[size_t(blockIdx.x) * blockDim.x + threadIdx.x] = s;
}

cudaError_t run_fp64_peak(int mode, int iters, double* scratch /* [>= 2 * sms * 256] */, int sms, double* tflops,
                          cudaStream_t st) {
  cudaEvent_t e0, e1;
  cudaError_t e = cudaEventCreate(&e0);
  if (e != cudaSuccess) return e;
  if ((e = cudaEventCreate(&e1)) != cudaSuccess) {
    cudaEventDestroy(e0);
    return e;
  }
  const int blocks = 2 * sms, threads = 256;
  float best = 1e30f;
  for (int rep = 0; rep < 4 && e == cudaSuccess; ++rep) {  // rep 0 warms up
    cudaEventRecord(e0, st);
    if (mode == 0) fp64_peak_kernel<0><<<blocks, threads, 0, st>>>(scratch, iters, 0.5);
    else if (mode == 1) fp64_peak_kernel<1><<<blocks, threads, 0, st>>>(scratch, iters, 0.5);
    else if (mode == 2) fp64_peak_kernel<2><<<blocks, threads, 0, st>>>(scratch, iters, 0.5);
    else if (mode == 3) fp64_lds_mix_kernel<3><<<blocks, threads, 0, st>>>(scratch, iters / 4);
    else if (mode == 4) fp64_lds_mix_kernel<4><<<blocks, threads, 0, st>>>(scratch, iters / 4);
    else fp64_lds_mix_kernel<5><<<blocks, threads, 0, st>>>(scratch, iters / 4);
    count_launch();
    cudaEventRecord(e1, st);
    e = cudaEventSynchronize(e1);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    if (e == cudaSuccess && rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (e != cudaSuccess) return e;
  const double warps = double(blocks) * threads / 32.0;
  const double per_warp_iter_mma = 8.0 * 512.0, per_warp_iter_fma = 16.0 * 64.0;
  double flop;
  if (mode == 0) flop = warps * iters * per_warp_iter_mma;
  else if (mode == 1) flop = warps * iters * per_warp_iter_fma;
  else if (mode == 2) flop = 0.5 * warps * iters * (per_warp_iter_mma + per_warp_iter_fma);
  else flop = warps * double(iters / 4) * 32.0 * 512.0;
  *tflops = flop / (double(best) * 1e-3) * 1e-12;
  return cudaGetLastError();
}

}  // namespace prb
