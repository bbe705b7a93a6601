// Candidate finalisation on the device: the steps the reference runs on the host after
// the multi-start optimisation of a PR acquisition function --
//   sample_candidates          one discrete draw per restart from p(theta*)        probabilistic_reparameterization.py:198-233
//   base-AF re-scoring + argmax over the restarts                                   run_one_replication.py:513-528
//   trust-region box around the incumbent                                           run_one_replication.py:410-440
// -- as kernels on the caller's stream, so that a BO iteration minus the objective evaluation never leaves the GPU.
#include "prb_device.cuh"
#include "prb_kernels.cuh"

namespace prb {

// One thread per restart row.  Uniforms: u[r, k] (k < n_int: integer parameter k, then one per categorical) when given,
// else Philox4x32-10 with counter (r, k, offset) -- a different `offset` than the MC sampler's keeps the streams apart.
//   integer k:     x_un = lo + x * range;  p = sigmoid((|x_un| - floor|x_un| - 0.5) / tau);  z = [u < p] (torch.bernoulli);
//                  v = clamp(floor(x_un) + z, lo, hi)      -- floor of the SIGNED value, exactly as :207-208 does
//   categorical j: p = softmax((x - 0.5) / tau) over the block;  category = first index with cumsum(p) > u (last index if
//                  rounding leaves cumsum(p)[-1] <= u);  block <- one-hot
//   out_x  [b, d]    normalised, one-hot layout (what sample_candidates returns)
//   out_xk [b, d_k]  the same rows as the base acquisition function sees them (OneHotToNumeric when apply_numeric)
__global__ void sample_candidates_kernel(const DevSpace sp, const double* __restrict__ theta, int b,
                                         const double* __restrict__ u, uint64_t seed, uint64_t offset,
                                         double* __restrict__ out_x, double* __restrict__ out_xk) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= b) return;
  const double* th = theta + size_t(r) * sp.d;
  const int nu = sp.n_int + sp.n_cat;
  const int numeric_start = sp.n_cat > 0 ? sp.cat_start[0] : sp.d;
  double* xr = out_x ? out_x + size_t(r) * sp.d : nullptr;
  double* kr = out_xk ? out_xk + size_t(r) * sp.d_k : nullptr;
  for (int c = 0; c < sp.d; ++c) {
    const double v = th[c];
    if (xr) xr[c] = v;
    if (kr && (!sp.apply_numeric || c < numeric_start)) kr[c] = v;
  }
  for (int k = 0; k < sp.n_int; ++k) {
    const int c = sp.int_cols[k];
    const double lo = sp.int_lo[k], hi = sp.int_hi[k];
    const double x_un = sp.raw_int ? th[c] : unnormalize_int(th[c], lo, hi);
    double off_abs;
    const double p = int_round_prob(x_un, sp.tau, &off_abs);
    const double uu = u ? u[size_t(r) * nu + k] : philox_uniform(seed, offset, uint32_t(r), uint32_t(k));
    const double z = (uu < p) ? 1.0 : 0.0;
    const double v = fmin(fmax(floor(x_un) + z, lo), hi);
    const double tn = sp.raw_int ? v : __dsub_rn(v, lo) / __dsub_rn(hi, lo);
    if (xr) xr[c] = tn;
    if (kr) kr[c] = tn;
  }
  for (int j = 0; j < sp.n_cat; ++j) {
    const int s = sp.cat_start[j], card = sp.cat_card[j];
    double mx = -1e300;
    for (int k = 0; k < card; ++k) mx = fmax(mx, __dsub_rn(th[s + k], 0.5) / sp.tau);
    double sum = 0.0;
    for (int k = 0; k < card; ++k) sum = __dadd_rn(sum, exp(__dsub_rn(__dsub_rn(th[s + k], 0.5) / sp.tau, mx)));
    const double inv = 1.0 / sum;
    const double uu = u ? u[size_t(r) * nu + sp.n_int + j]
                        : philox_uniform(seed, offset, uint32_t(r), uint32_t(sp.n_int + j));
    int cat = card - 1;
    double cum = 0.0;
    bool found = false;
    for (int k = 0; k < card; ++k) {
      cum = __dadd_rn(cum, __dmul_rn(exp(__dsub_rn(__dsub_rn(th[s + k], 0.5) / sp.tau, mx)), inv));
      if (!found && cum > uu) {
        cat = k;
        found = true;
      }
    }
    for (int k = 0; k < card; ++k) {
      const double oh = (k == cat) ? 1.0 : 0.0;
      if (xr) xr[s + k] = oh;
      if (kr && !sp.apply_numeric) kr[s + k] = oh;
    }
    if (kr && sp.apply_numeric) write_categorical_input(sp, kr, numeric_start, j, cat);
  }
}

cudaError_t launch_sample_candidates(const DevSpace& sp, const double* theta, int b, const double* u, uint64_t seed,
                                     uint64_t offset, double* out_x, double* out_xk, cudaStream_t st) {
  ProfScope prof_(K_SAMPLE, st);
  if (b == 0) return cudaSuccess;
  sample_candidates_kernel<<<(b + 63) / 64, 64, 0, st>>>(sp, theta, b, u, seed, offset, out_x, out_xk);
  count_launch();
  return cudaGetLastError();
}

// argmax over the restarts (lowest index wins ties; NaN never wins) + copy of the winning row: out_best = (value, row[d]).
__global__ void __launch_bounds__(256) best_row_kernel(const double* __restrict__ vals, const double* __restrict__ rows,
                                                        int b, int d, double* __restrict__ out_best,
                                                        int32_t* __restrict__ out_index) {
  __shared__ double sv[256];
  __shared__ int si[256];
  const int tid = threadIdx.x;
  double bv = -INFINITY;
  int bi = -1;
  for (int i = tid; i < b; i += 256) {
    const double v = vals[i];
    if (v > bv || (bi < 0 && v == v)) {
      bv = v;
      bi = i;
    }
  }
  sv[tid] = bv;
  si[tid] = bi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) {
      const double v2 = sv[tid + o];
      const int i2 = si[tid + o];
      const bool take = i2 >= 0 && (si[tid] < 0 || v2 > sv[tid] || (v2 == sv[tid] && i2 < si[tid]));
      if (take) {
        sv[tid] = v2;
        si[tid] = i2;
      }
    }
    __syncthreads();
  }
  const int win = si[0] < 0 ? 0 : si[0];
  if (tid == 0) {
    if (out_index) *out_index = win;
    if (out_best) out_best[0] = vals[win];
  }
  if (out_best && rows)
    for (int c = tid; c < d; c += 256) out_best[1 + c] = rows[size_t(win) * d + c];
}

cudaError_t launch_best_row(const double* vals, const double* rows, int b, int d, double* out_best, int32_t* out_index,
                            cudaStream_t st) {
  ProfScope prof_(K_MISC, st);
  if (b == 0) return cudaSuccess;
  best_row_kernel<<<1, 256, 0, st>>>(vals, rows, b, d, out_best, out_index);
  count_launch();
  return cudaGetLastError();
}

// Trust-region box (run_one_replication.py:410-440; centre selection :413-423).  Y [n, 1 + n_con]: column 0 the objective
// (maximised), columns 1.. the constraint slacks (>= 0 feasible).  Centre = best feasible point, else the least-violating
// one; plain argmax without constraints.  bounds = clamp(centre -/+ length * (ub - lb)) with the columns flagged in reset01
// (binary dims, one-hot blocks) put back to [0, 1].  One CTA.
__global__ void __launch_bounds__(256) tr_bounds_kernel(const double* __restrict__ X, const double* __restrict__ Y, int n,
                                                         int d, int n_con, double length,
                                                         const double* __restrict__ lower,
                                                         const double* __restrict__ upper,
                                                         const uint8_t* __restrict__ reset01,
                                                         double* __restrict__ out_bounds,
                                                         int32_t* __restrict__ out_center) {
  __shared__ double sv[256];
  __shared__ int si[256];
  __shared__ int any_feas;
  const int tid = threadIdx.x;
  const int ld = 1 + n_con;
  if (tid == 0) any_feas = 0;
  __syncthreads();
  if (n_con > 0) {
    for (int i = tid; i < n; i += 256) {
      bool f = true;
      for (int k = 1; k < ld; ++k) f = f && (Y[size_t(i) * ld + k] >= 0.0);
      if (f) any_feas = 1;
    }
  }
  __syncthreads();
  const bool by_violation = n_con > 0 && !any_feas;
  double bv = -INFINITY;
  int bi = -1;
  for (int i = tid; i < n; i += 256) {
    double merit;
    if (by_violation) {
      double viol = 0.0;
      for (int k = 1; k < ld; ++k) viol += fabs(fmin(Y[size_t(i) * ld + k], 0.0));
      merit = -viol;
    } else {
      merit = Y[size_t(i) * ld];
      if (n_con > 0) {
        bool f = true;
        for (int k = 1; k < ld; ++k) f = f && (Y[size_t(i) * ld + k] >= 0.0);
        if (!f) merit = -INFINITY;
      }
    }
    if (merit > bv || bi < 0) {
      bv = merit;
      bi = i;
    }
  }
  sv[tid] = bv;
  si[tid] = bi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) {
      const double v2 = sv[tid + o];
      const int i2 = si[tid + o];
      const bool take = i2 >= 0 && (si[tid] < 0 || v2 > sv[tid] || (v2 == sv[tid] && i2 < si[tid]));
      if (take) {
        sv[tid] = v2;
        si[tid] = i2;
      }
    }
    __syncthreads();
  }
  const int win = si[0] < 0 ? 0 : si[0];
  if (tid == 0 && out_center) *out_center = win;
  for (int c = tid; c < d; c += 256) {
    const double span = length * (upper[c] - lower[c]);
    const double xc = X[size_t(win) * d + c];
    double lo = fmax(xc - span, lower[c]);
    double hi = fmin(xc + span, upper[c]);
    if (reset01 && reset01[c]) {
      lo = 0.0;
      hi = 1.0;
    }
    out_bounds[c] = lo;
    out_bounds[d + c] = hi;
  }
}

cudaError_t launch_tr_bounds(const double* X, const double* Y, int n, int d, int n_con, double length, const double* lower,
                             const double* upper, const uint8_t* reset01, double* out_bounds, int32_t* out_center,
                             cudaStream_t st) {
  ProfScope prof_(K_MISC, st);
  tr_bounds_kernel<<<1, 256, 0, st>>>(X, Y, n, d, n_con, length, lower, upper, reset01, out_bounds, out_center);
  count_launch();
  return cudaGetLastError();
}

// =====================================================================================================================
// Model lists: ConstrainedExpectedImprovement over independent GPs ([3P] botorch <= 0.8; experiment_utils.py:354-392).
//   sigma_m = max(sqrt(max(var_m, min_variance)), sigma_floor);  A = EI(mu_o, sigma_o; best_f) * prod_m P_m,
//   P_m = [1 - cdf((lower_m - mu_m)/sigma_m)] and / or [cdf((upper_m - mu_m)/sigma_m)], cdf(x) = 0.5 (1 + erf(x / sqrt2)).
// One thread per evaluation point: raw variances from every model's row-tile norms, the value, and -- for the pathwise
// gradient -- dA/dmu_m and dA/dsigma_m^2 of every model (product rule; the clamps pass no gradient where they bind).
// =====================================================================================================================
//
// Expected hypervolume improvement ([3P] botorch <= 0.8 multi_objective/analytic.py; experiment_utils.py:395-402), q = 1, all
// models objectives:  sigma_m = sqrt(max(var_m, var_floor)),  EHVI = sum_cells prod_m s_m(cell),
//   s = psi(l, l) - psi(l, u) + nu(l, u)   (botorch sums the 2^M products of {psi_diff, nu} choices: the same number),
//   d s / d mu = cdf(t_u) - cdf(t_l),   d s / d sigma = pdf(t_l) - pdf(t_u),   t = (bound - mu) / sigma.
// One thread per evaluation point; the cells are read by every thread in the same order (broadcast loads).
__global__ void multi_ehvi_epilogue_kernel(const MultiAcqArgs ma, int ncols_pad, int ncols, double* __restrict__ a,
                                           double* __restrict__ out_mean, double* __restrict__ out_var) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  const int M = ma.n_models;
  double mu[PRB_MAX_MODELS], sg[PRB_MAX_MODELS], pass[PRB_MAX_MODELS];
  for (int m = 0; m < M; ++m) {
    double vs = 0.0;
    for (int t = 0; t < ma.mtiles[m]; ++t) vs += ma.norm_part[m][size_t(t) * ncols_pad + col];
    const double var_post = fmax(ma.kss[m] - vs, ma.min_variance);  // gpytorch floor, then botorch's clamp_min
    mu[m] = ma.mean_const[m] + ma.mu_dot[m][col];
    sg[m] = sqrt(fmax(var_post, ma.var_floor));
    pass[m] = (ma.kss[m] - vs >= ma.min_variance && var_post >= ma.var_floor) ? 1.0 : 0.0;
    if (out_mean) out_mean[size_t(col) * M + m] = mu[m];
    if (out_var) out_var[size_t(col) * M + m] = var_post;
  }
  const bool grad = ma.dmu[0] != nullptr;
  double total = 0.0, g_mu[PRB_MAX_MODELS], g_sg[PRB_MAX_MODELS];
  for (int m = 0; m < M; ++m) g_mu[m] = g_sg[m] = 0.0;
  const double* lo = ma.cell_bounds;
  const double* hi = ma.cell_bounds + size_t(ma.n_cells) * M;
  for (int c = 0; c < ma.n_cells; ++c) {
    double s[PRB_MAX_MODELS], ds_mu[PRB_MAX_MODELS], ds_sg[PRB_MAX_MODELS];
    double prod = 1.0;
    for (int m = 0; m < M; ++m) {
      const double l = lo[size_t(c) * M + m];
      const double u = fmin(hi[size_t(c) * M + m], 1e10);
      const double tl = (l - mu[m]) / sg[m], tu = (u - mu[m]) / sg[m];
      const double pl = exp(-(tl * tl) / 2.0 - PRB_LOG_SQRT_2PI), pu = exp(-(tu * tu) / 2.0 - PRB_LOG_SQRT_2PI);
      const double cl = 0.5 * (1.0 + erf(tl * PRB_INV_SQRT2)), cu = 0.5 * (1.0 + erf(tu * PRB_INV_SQRT2));
      const double psi_ll = sg[m] * pl + (mu[m] - l) * (1.0 - cl);
      const double psi_lu = sg[m] * pu + (mu[m] - l) * (1.0 - cu);
      const double nu = (u - l) * (1.0 - cu);
      s[m] = (psi_ll - psi_lu) + nu;
      ds_mu[m] = cu - cl;
      ds_sg[m] = pl - pu;
      prod *= s[m];
    }
    total += prod;
    if (grad) {
      for (int m = 0; m < M; ++m) {
        double others = 1.0;
        for (int k = 0; k < M; ++k)
          if (k != m) others *= s[k];
        g_mu[m] = fma(others, ds_mu[m], g_mu[m]);
        g_sg[m] = fma(others, ds_sg[m], g_sg[m]);
      }
    }
  }
  a[col] = total;
  if (!grad) return;
  for (int m = 0; m < M; ++m) {
    ma.dmu[m][col] = g_mu[m];
    ma.dvar[m][col] = g_sg[m] / (2.0 * sg[m]) * pass[m];
  }
}

__global__ void multi_acq_epilogue_kernel(const MultiAcqArgs ma, int ncols_pad, int ncols, double* __restrict__ a,
                                          double* __restrict__ out_mean, double* __restrict__ out_var) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  double mu[PRB_MAX_MODELS], sg[PRB_MAX_MODELS], pass[PRB_MAX_MODELS];
  for (int m = 0; m < ma.n_models; ++m) {
    double vs = 0.0;
    for (int t = 0; t < ma.mtiles[m]; ++t) vs += ma.norm_part[m][size_t(t) * ncols_pad + col];
    const double var_raw = ma.kss[m] - vs;
    const double var = fmax(var_raw, ma.min_variance);
    const double s_raw = sqrt(var);
    mu[m] = ma.mean_const[m] + ma.mu_dot[m][col];
    sg[m] = fmax(s_raw, ma.sigma_floor);
    pass[m] = (var_raw >= ma.min_variance && s_raw >= ma.sigma_floor) ? 1.0 : 0.0;
    if (out_mean) out_mean[size_t(col) * ma.n_models + m] = mu[m];
    if (out_var) out_var[size_t(col) * ma.n_models + m] = var;
  }
  const int o = ma.objective_index;
  const double u = (mu[o] - ma.best_f) / sg[o];
  const double pdf = exp(-(u * u) / 2.0 - PRB_LOG_SQRT_2PI);
  const double cdf = 0.5 * (1.0 + erf(u * PRB_INV_SQRT2));
  const double ei = sg[o] * (pdf + u * cdf);
  double P[PRB_MAX_MODELS], dP_dmu[PRB_MAX_MODELS], dP_dvar[PRB_MAX_MODELS];
  for (int m = 0; m < ma.n_models; ++m) {
    P[m] = 1.0;
    dP_dmu[m] = dP_dvar[m] = 0.0;
    if (m == o) continue;
    double hi_c = 1.0, lo_c = 0.0, g_mu = 0.0, g_s = 0.0;  // P = hi_c - lo_c
    if (ma.has_upper[m]) {
      const double d = (ma.upper[m] - mu[m]) / sg[m];
      const double ph = exp(-(d * d) / 2.0 - PRB_LOG_SQRT_2PI);
      hi_c = 0.5 * (1.0 + erf(d * PRB_INV_SQRT2));
      g_mu -= ph / sg[m];
      g_s -= ph * d / sg[m];
    }
    if (ma.has_lower[m]) {
      const double d = (ma.lower[m] - mu[m]) / sg[m];
      const double ph = exp(-(d * d) / 2.0 - PRB_LOG_SQRT_2PI);
      lo_c = 0.5 * (1.0 + erf(d * PRB_INV_SQRT2));
      g_mu += ph / sg[m];
      g_s += ph * d / sg[m];
    }
    if (ma.has_lower[m] && !ma.has_upper[m]) {
      P[m] = 1.0 - lo_c;  // botorch: prob_l = 1 - cdf(dist_l)
    } else if (ma.has_lower[m] || ma.has_upper[m]) {
      P[m] = hi_c - lo_c;
    }
    dP_dmu[m] = g_mu;
    dP_dvar[m] = g_s / (2.0 * sg[m]) * pass[m];  // d/d sigma^2 = (d/d sigma) / (2 sigma)
  }
  double prod = 1.0;
  for (int m = 0; m < ma.n_models; ++m) prod *= P[m];
  a[col] = ei * prod;
  if (ma.dmu[0] == nullptr) return;
  for (int m = 0; m < ma.n_models; ++m) {
    if (m == o) {
      ma.dmu[m][col] = cdf * prod;
      ma.dvar[m][col] = pdf / (2.0 * sg[o]) * pass[o] * prod;
    } else {
      double others = ei;
      for (int k = 0; k < ma.n_models; ++k)
        if (k != m) others *= P[k];
      ma.dmu[m][col] = others * dP_dmu[m];
      ma.dvar[m][col] = others * dP_dvar[m];
    }
  }
}

cudaError_t launch_multi_acq_epilogue(const MultiAcqArgs& ma, int ncols_pad, int ncols, double* a, double* out_mean,
                                      double* out_var, cudaStream_t st) {
  ProfScope prof_(K_EPILOGUE, st);
  if (ncols == 0) return cudaSuccess;
  if (ma.kind == PRB_MACQ_EHVI)
    multi_ehvi_epilogue_kernel<<<(ncols + 127) / 128, 128, 0, st>>>(ma, ncols_pad, ncols, a, out_mean, out_var);
  else
    multi_acq_epilogue_kernel<<<(ncols + 127) / 128, 128, 0, st>>>(ma, ncols_pad, ncols, a, out_mean, out_var);
  count_launch();
  return cudaGetLastError();
}

__global__ void add_inplace_kernel(double* __restrict__ dst, const double* __restrict__ src, size_t count) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < count) dst[i] += src[i];
}
cudaError_t launch_add_inplace(double* dst, const double* src, size_t count, cudaStream_t st) {
  ProfScope prof_(K_SUM_SPLITS, st);
  if (count == 0) return cudaSuccess;
  add_inplace_kernel<<<unsigned((count + 255) / 256), 256, 0, st>>>(dst, src, count);
  count_launch();
  return cudaGetLastError();
}

}  // namespace prb
