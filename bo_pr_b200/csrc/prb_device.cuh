// Device-side math shared by the PR kernels: Philox4x32-10, rounding probabilities, Matern-5/2 / categorical factors.
#pragma once

#include "prb_internal.cuh"

namespace prb {

// ---------------------------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11; same constants as Random123 / cuRAND).  counter = (i, col, off_lo, off_hi),
// key = (seed_lo, seed_hi).  u = (53 high bits of (r0:r1)) * 2^-53 in [0, 1).  oracle/philox.py restates it in numpy.
// ---------------------------------------------------------------------------------------------------------------------
__host__ __device__ inline void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0,
                                             uint32_t k1) {
  const uint64_t p0 = uint64_t(0xD2511F53u) * c0;
  const uint64_t p1 = uint64_t(0xCD9E8D57u) * c2;
  const uint32_t hi0 = uint32_t(p0 >> 32), lo0 = uint32_t(p0);
  const uint32_t hi1 = uint32_t(p1 >> 32), lo1 = uint32_t(p1);
  const uint32_t n0 = hi1 ^ c1 ^ k0;
  const uint32_t n1 = lo1;
  const uint32_t n2 = hi0 ^ c3 ^ k1;
  const uint32_t n3 = lo0;
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}

__host__ __device__ inline double philox_uniform(uint64_t seed, uint64_t offset, uint32_t i, uint32_t col) {
  uint32_t c0 = i, c1 = col, c2 = uint32_t(offset), c3 = uint32_t(offset >> 32);
  uint32_t k0 = uint32_t(seed), k1 = uint32_t(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  const uint64_t bits = ((uint64_t(c0) << 32) | uint64_t(c1)) >> 11;  // 53 bits
  return double(bits) * (1.0 / 9007199254740992.0);
}

// ---------------------------------------------------------------------------------------------------------------------
// Rounding probability of one integer parameter (input.py:613-635): sigmoid((|x| - floor|x| - 0.5) / tau) on the
// UN-normalised value.  x_un = lo + x * range with a separately rounded multiply and add (torch does not fuse them).
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double unnormalize_int(double x, double lo, double hi) {
  return __dadd_rn(lo, __dmul_rn(x, __dsub_rn(hi, lo)));
}
__device__ __forceinline__ double sigmoid_f64(double y) { return 1.0 / (1.0 + exp(-y)); }
__device__ __forceinline__ double int_round_prob(double x_un, double tau, double* floor_abs) {
  const double a = fabs(x_un);
  const double off = floor(a);
  *floor_abs = off;
  return sigmoid_f64(__dsub_rn(__dsub_rn(a, off), 0.5) / tau);
}

// ---------------------------------------------------------------------------------------------------------------------
// Kernel factors ([3P] gpytorch MaternKernel nu=2.5, botorch CategoricalKernel).
// ---------------------------------------------------------------------------------------------------------------------
#define PRB_SQRT5 2.2360679774997896964
// value F(r) = (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r),  r = sqrt(max(r2, 1e-30));
// G = (dF/dr)/r = -(5/3)(1 + sqrt5 r) exp(-sqrt5 r)   so that dF/dx_k = G * (x_k - X_k) / l_k^2 (no division by r)
__device__ __forceinline__ double matern52(double r2, double* G) {
  const double r = sqrt(fmax(r2, 1e-30));
  const double s5r = PRB_SQRT5 * r;
  const double e = exp(-s5r);
  *G = -(5.0 / 3.0) * (1.0 + s5r) * e;
  return (1.0 + s5r + (5.0 / 3.0) * (r * r)) * e;
}
__device__ __forceinline__ double matern52_value(double r2) {
  const double r = sqrt(fmax(r2, 1e-30));
  const double s5r = PRB_SQRT5 * r;
  return (1.0 + s5r + (5.0 / 3.0) * (r * r)) * exp(-s5r);
}

// Branch-free sqrt / exp for kernels that evaluate many independent Matern entries per thread: the library versions carry
// slow-path branches, which keep the compiler from interleaving the evaluations.  Arguments are bounded (r2 in [1e-30, 1e8],
// exponent in [-690, 0]); both are accurate to about one ulp, far inside the 1e-9 parity tolerance.
__device__ __forceinline__ double sqrt_nb(double x) {
  double y = double(rsqrtf(__double2float_rn(x)));  // ~2^-22 relative
  const double hx = 0.5 * x;
  y = y * fma(-hx * y, y, 1.5);
  y = y * fma(-hx * y, y, 1.5);
  const double sq = x * y;
  return fma(0.5 * y, fma(-sq, sq, x), sq);
}
__device__ __forceinline__ double exp_neg_nb(double x) {  // x <= 0
  x = fmax(x, -690.0);
  const double kf = rint(x * 1.4426950408889634074);
  double r = fma(kf, -6.93147180369123816490e-01, x);
  r = fma(kf, -1.90821492927058770002e-10, r);
  double q = 1.6059043836821613e-10;  // 1/13!
  q = fma(q, r, 2.08767569878681e-09);
  q = fma(q, r, 2.505210838544172e-08);
  q = fma(q, r, 2.755731922398589e-07);
  q = fma(q, r, 2.755731922398589e-06);
  q = fma(q, r, 2.48015873015873e-05);
  q = fma(q, r, 1.984126984126984e-04);
  q = fma(q, r, 1.388888888888889e-03);
  q = fma(q, r, 8.333333333333333e-03);
  q = fma(q, r, 4.1666666666666664e-02);
  q = fma(q, r, 1.6666666666666666e-01);
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  q = fma(q, r, 1.0);
  const int k = __double2int_rn(kf);
  return __hiloint2double(__double2hiint(q) + (k << 20), __double2loint(q));
}
__device__ __forceinline__ double matern52_nb(double r2, double* G) {
  const double r = sqrt_nb(fmax(r2, 1e-30));
  const double s5r = PRB_SQRT5 * r;
  const double e = exp_neg_nb(-s5r);
  *G = -(5.0 / 3.0) * (1.0 + s5r) * e;
  return (1.0 + s5r + (5.0 / 3.0) * (r * r)) * e;
}

// ---------------------------------------------------------------------------------------------------------------------
// Base acquisition function from the posterior mean and the raw variance k** - ||v||^2, with its partial derivatives.
//   [3P] gpytorch variance clamp (min_variance), botorch ExpectedImprovement / UpperConfidenceBound / PosteriorMean
//   (SURVEY.md App. B.3-B.4); call sites experiment_utils.py:391-393, 476-480, 489-492.
// ---------------------------------------------------------------------------------------------------------------------
#define PRB_LOG_SQRT_2PI 0.91893853320467274178
#define PRB_INV_SQRT2 0.70710678118654752440
__device__ __forceinline__ void acq_eval(const DevAcq& aq, double mean, double var_raw, double* val, double* d_mu,
                                         double* d_var, double* var_out) {
  const double var = fmax(var_raw, aq.min_variance);
  const double pass1 = (var_raw >= aq.min_variance) ? 1.0 : 0.0;
  if (aq.kind == PRB_ACQ_EI) {
    const double v2 = fmax(var, aq.acq_var_floor);
    const double pass2 = (var >= aq.acq_var_floor) ? 1.0 : 0.0;
    const double sigma = sqrt(v2);
    const double u = (mean - aq.param) / sigma;
    const double pdf = exp(-(u * u) / 2.0 - PRB_LOG_SQRT_2PI);
    const double cdf = aq.ei_variant == 0 ? 0.5 * (1.0 + erf(u * PRB_INV_SQRT2)) : 0.5 * erfc(-u * PRB_INV_SQRT2);
    *val = sigma * (pdf + u * cdf);
    *d_mu = cdf;
    *d_var = pdf / (2.0 * sigma) * pass1 * pass2;
  } else if (aq.kind == PRB_ACQ_UCB) {
    const double bv = aq.param * var;
    const double sd = sqrt(bv);
    *val = mean + sd;
    *d_mu = 1.0;
    *d_var = (sd > 0.0 ? aq.param / (2.0 * sd) : 0.0) * pass1;
  } else {
    *val = mean;
    *d_mu = 1.0;
    *d_var = 0.0;
  }
  *var_out = var;
}

__device__ __forceinline__ double ipow(double v, int p) {
  double r = v;
  for (int i = 1; i < p; ++i) r *= v;
  return r;
}

// ---------------------------------------------------------------------------------------------------------------------
// PR sampler building blocks (shared by pr_sample_kernel, the fused step kernels and the small-problem step kernel)
// ---------------------------------------------------------------------------------------------------------------------
// Rounding probabilities of one restart (input.py:613-635).  They do not depend on the MC sample, so they are computed
// once per restart by `nworkers` cooperating threads (worker id t): ps[0 .. n_int) integer probabilities, offs[k] =
// floor|x_un|, ps[n_int ..) the normalised softmax entries of every one-hot block, in `discrete_indices` order.
__device__ __forceinline__ void pr_restart_probs(const DevSpace& sp, const double* __restrict__ th, int t, int nworkers,
                                                 double* __restrict__ ps, double* __restrict__ offs) {
  for (int k = t; k < sp.n_int; k += nworkers) {
    const double x = th[sp.int_cols[k]];
    const double x_un = sp.raw_int ? x : unnormalize_int(x, sp.int_lo[k], sp.int_hi[k]);
    double off;
    ps[k] = int_round_prob(x_un, sp.tau, &off);
    offs[k] = off;
  }
  for (int j = t; j < sp.n_cat; j += nworkers) {
    int zpos = sp.n_int;
    for (int jj = 0; jj < j; ++jj) zpos += sp.cat_card[jj];
    const int s = sp.cat_start[j], card = sp.cat_card[j];
    double* pk = ps + zpos;
    double mx = -1e300;
    for (int k = 0; k < card; ++k) {
      pk[k] = __dsub_rn(th[s + k], 0.5) / sp.tau;
      mx = fmax(mx, pk[k]);
    }
    double sum = 0.0;
    for (int k = 0; k < card; ++k) {
      pk[k] = exp(__dsub_rn(pk[k], mx));
      sum = __dadd_rn(sum, pk[k]);
    }
    const double inv = 1.0 / sum;
    for (int k = 0; k < card; ++k) pk[k] = __dmul_rn(pk[k], inv);
  }
}

// Integer parameter k of MC sample i (input.py:674-692; probabilistic_reparameterization.py:454-465): returns the sampled,
// clamped, re-normalised value and the rounding component z used by the score-function estimator.
__device__ __forceinline__ double pr_sample_int(const DevSpace& sp, const double* __restrict__ th,
                                                const double* __restrict__ ps, const double* __restrict__ offs, int i,
                                                int k, const double* __restrict__ u_int, uint64_t seed, uint64_t offset,
                                                uint8_t* z_out) {
  const double x = th[sp.int_cols[k]];
  const double lo = sp.int_lo[k], hi = sp.int_hi[k];
  const double range = __dsub_rn(hi, lo);
  const double x_un = sp.raw_int ? x : unnormalize_int(x, lo, hi);
  const double u = u_int ? u_int[size_t(i) * sp.n_int + k] : philox_uniform(seed, offset, uint32_t(i), uint32_t(k));
  const double zb = (ps[k] >= u) ? 1.0 : 0.0;
  const double xa = offs[k] + zb;
  const double xs = (x_un < 0.0) ? -xa : xa;
  const double xc = fmin(fmax(xs, lo), hi);
  // x / 1.0 == x exactly: binary parameters (range 1) skip the FP64 divide without changing a bit
  const double tn = sp.raw_int ? xc : (range == 1.0 ? __dsub_rn(xc, lo) : __dsub_rn(xc, lo) / range);
  *z_out = ((__dsub_rn(tn, x) > 0.0) || (x == 1.0)) ? 1 : 0;
  return tn;
}

// Kernel-input columns of categorical j on the numeric layout: the category index (OneHotToNumeric, input.py:69-88) or,
// with a latent embedding, row `cat` of the categorical's table (EmbeddingTransform.transform, input.py:784-802).
__device__ __forceinline__ void write_categorical_input(const DevSpace& sp, double* __restrict__ xr, int numeric_start,
                                                        int j, int cat) {
  if (sp.latent_tab == nullptr) {
    xr[numeric_start + j] = double(cat);
    return;
  }
  const int e = sp.latent_dim[j];
  const double* row = sp.latent_tab + sp.latent_off[j] + size_t(cat) * e;
  double* dst = xr + numeric_start + sp.latent_col[j];
  for (int t = 0; t < e; ++t) dst[t] = row[t];
}

// One (restart, MC sample i) of the sampler.  `th` = the restart's theta row, ps / offs from pr_restart_probs.
// Row outputs (any may be NULL): xr [d_k] kernel input, tr [d] tilde_x, zr [n_disc] rounding components; returns the bit
// code over the binary dims selected by `bm`.
__device__ __forceinline__ uint32_t pr_sample_point(const DevSpace& sp, const double* __restrict__ th,
                                                    const double* __restrict__ ps, const double* __restrict__ offs,
                                                    int i, const double* __restrict__ u_int,
                                                    const double* __restrict__ u_cat, uint64_t seed, uint64_t offset,
                                                    double* __restrict__ xr, double* __restrict__ tr,
                                                    uint8_t* __restrict__ zr, const BitMap& bm) {
  const int numeric_start = sp.n_cat > 0 ? sp.cat_start[0] : sp.d;
  for (int c = 0; c < sp.d; ++c) {
    const double v = th[c];
    if (tr) tr[c] = v;
    if (xr && (!sp.apply_numeric || c < numeric_start)) xr[c] = v;
  }
  uint32_t bits = 0;
  // ---- integers (input.py:674-692)
  for (int k = 0; k < sp.n_int; ++k) {
    uint8_t zk;
    const double tn = pr_sample_int(sp, th, ps, offs, i, k, u_int, seed, offset, &zk);
    const int c = sp.int_cols[k];
    if (tr) tr[c] = tn;
    if (xr) xr[c] = tn;
    if (zr) zr[k] = zk;
    if (bm.pos[k] >= 0 && tn != 0.0) bits |= 1u << bm.pos[k];
  }
  // ---- categoricals (input.py:715-744): sequential cumsum of the softmax, first index with cum > u (else 0)
  int zpos = sp.n_int;
  for (int j = 0; j < sp.n_cat; ++j) {
    const int s = sp.cat_start[j], card = sp.cat_card[j];
    const double* pk = ps + zpos;
    const double u = u_cat ? u_cat[size_t(i) * sp.n_cat + j]
                           : philox_uniform(seed, offset, uint32_t(i), uint32_t(sp.n_int + j));
    int cat = 0;
    bool found = false;
    double cum = 0.0;
    for (int k = 0; k < card; ++k) {
      cum = __dadd_rn(cum, pk[k]);
      if (!found && cum > u) {
        cat = k;
        found = true;
      }
    }
    for (int k = 0; k < card; ++k) {
      const double oh = (k == cat) ? 1.0 : 0.0;
      if (tr) tr[s + k] = oh;
      if (xr && !sp.apply_numeric) xr[s + k] = oh;
      if (zr) zr[zpos + k] = (k == cat) ? 1 : 0;
    }
    if (xr && sp.apply_numeric) write_categorical_input(sp, xr, numeric_start, j, cat);
    zpos += card;
  }
  return bits;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace prb
