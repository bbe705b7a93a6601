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

}  // namespace prb
