// Launcher declarations for prb_kernels.cu (internal).
#pragma once

#include "prb_internal.cuh"

namespace prb {

cudaError_t launch_philox_fill(uint64_t seed, uint64_t offset, int mc, int n_cols, double* out, cudaStream_t st);
cudaError_t launch_pr_sample(const DevSpace& sp, const double* theta, int b, int mc, const double* u_int,
                             const double* u_cat, uint64_t seed, uint64_t offset, double* xk, double* tilde,
                             uint8_t* z, double* prob, const BitMap* bm, uint32_t* zbits, cudaStream_t st,
                             const double* lower = nullptr, const double* upper = nullptr, double* Xc = nullptr);
cudaError_t launch_sep_model_prep(const Model& m, int n_pad, int* bad_flag, cudaStream_t st);
cudaError_t launch_sep_restart_prep(const Model& m, const double* theta, int nb, int d_theta, double* kc, double* Dc,
                                    cudaStream_t st);
cudaError_t launch_sep_step_prep(const Model& m, const DevSpace& sp, const BitMap& bm, const double* theta,
                                 const double* lower, const double* upper, double* Xc, int nb, int mc,
                                 const double* u_int, const double* u_cat, uint64_t seed, uint64_t offset, double* kc,
                                 double* Dc, uint8_t* z, double* prob, uint32_t* zbits, cudaStream_t st);
cudaError_t launch_sep_step_finish(const DevSpace& sp, const DimMap& dm, int slots, int slot_is_dim, int nb, int mc,
                                   const double* a, const uint8_t* z,
                                   const double* prob, const double* S_part, int mtiles, int ncols_pad, int estimator,
                                   double* ma_state, int update_ma, const double* grad_out, double* out_value,
                                   double* out_grad, int need_grad, double* theta, double* adam_m, double* adam_v,
                                   double lr, int* step_ptr, unsigned int* done, cudaStream_t st);
cudaError_t launch_pad_copy(const double* src, int n, int n_pad, double* dst, cudaStream_t st);
cudaError_t launch_analytic_build(const DevSpace& sp, const double* theta, int b, int n_opt, const int32_t* options,
                                  double* xk, double* probs, cudaStream_t st, const double* lower = nullptr,
                                  const double* upper = nullptr, double* Xc = nullptr);
cudaError_t launch_kstar_panel(const Model& m, const double* xk, int ncols, int ncols_pad, double* KstarT,
                               cudaStream_t st, const GradGroup* gq = nullptr, double* Qp = nullptr);
cudaError_t launch_kernel_matrix(const DevSpec& spec, const double* Xtr, int n, const double* xk, int n_points,
                                 double* out, cudaStream_t st);
cudaError_t launch_rff_eval(const Model& m, const double* xk, int64_t ncols, uint64_t diffmask, bool need_grad, double* a,
                            double* S_all, double* out_mean, double* out_var, cudaStream_t st, int64_t dim_stride = 0);
cudaError_t launch_acq_epilogue(const Model& m, const DevAcq& aq, const double* norm_part, int ncols_pad,
                                const double* mu_dot, int ncols, double* a, double* dmu, double* dvar,
                                double* out_mean, double* out_var, const int* ncols_dev, int64_t col0,
                                cudaStream_t st);
cudaError_t launch_dedup(const uint32_t* zbits, int nb, int mc, int nbits, int* ucount, int* ustart, int* total,
                         uint32_t* ucode, int* uweight, int* inv, uint32_t* zbits_u, int* colb, double* w_u,
                         cudaStream_t st);
cudaError_t launch_mc_reduce_dedup(const DevSpace& sp, const BitMap& bm, int b, int mc, const int* ucount,
                                   const int* ustart, const double* w_u, const uint32_t* zbits_u, const double* a,
                                   const double* prob, const double* S_all, int estimator, const double* ma_state,
                                   const double* grad_out, double* out_value, double* out_grad, cudaStream_t st);
cudaError_t launch_dedup_scatter(const double* a, const int* ustart, const int* inv, int b, int mc, double* out,
                                 cudaStream_t st);
cudaError_t launch_grad_contract(const Model& m, uint64_t diffmask, const double* xk, int ncols, int ncols_pad,
                                 const double* W, const double* dmu, const double* dvar, double* S_part, int jrows,
                                 cudaStream_t st);
cudaError_t launch_sum_splits(const Model& m, const double* S_part, int jblocks, const DimMap* dm, int ncols,
                              int ncols_pad, double* S_all, const int* ncols_dev, int64_t col0, cudaStream_t st);
cudaError_t launch_mc_reduce(const DevSpace& sp, int b, int mc, const double* a, const uint8_t* z, const double* prob,
                             const double* S_all, int estimator, const double* ma_state, const double* grad_out,
                             double* out_value, double* out_grad, cudaStream_t st);
cudaError_t launch_ma_update(const double* value, int b, double* ma_state, cudaStream_t st);
cudaError_t launch_analytic_reduce(const DevSpace& sp, const double* theta, int b, int n_opt, const int32_t* options,
                                   const double* a, const double* probs, const double* S_all, const double* grad_out,
                                   double* out_value, double* out_grad, cudaStream_t st, const double* S_part = nullptr,
                                   int jblocks = 0, int ncols_pad = 0, double* adam_theta = nullptr,
                                   double* adam_m = nullptr, double* adam_v = nullptr, double lr = 0.0,
                                   int* step_ptr = nullptr, unsigned int* done = nullptr);
cudaError_t launch_clamp(const double* theta, const double* lower, const double* upper, int b, int d, double* out,
                         cudaStream_t st);
cudaError_t launch_adam_step(double* theta, const double* acq_grad, double* m, double* v, int total, double lr,
                             int* step_ptr, cudaStream_t st);
cudaError_t launch_fill(double* p, double v, int n, cudaStream_t st);
// model lists (prb_finalize.cu): combined acquisition epilogue and the accumulation of the per-model gradient partials
struct MultiAcqArgs {
  int kind;  // prb_multi_acq_kind
  int n_models, objective_index;
  double best_f, min_variance, sigma_floor;
  double var_floor;           // EHVI
  int n_cells;                // EHVI
  const double* cell_bounds;  // EHVI: [2, n_cells, n_models]
  int has_lower[PRB_MAX_MODELS], has_upper[PRB_MAX_MODELS];
  double lower[PRB_MAX_MODELS], upper[PRB_MAX_MODELS];
  double mean_const[PRB_MAX_MODELS], kss[PRB_MAX_MODELS];
  const double* norm_part[PRB_MAX_MODELS];  // [mtiles, ncols_pad] per model
  const double* mu_dot[PRB_MAX_MODELS];     // [ncols] per model
  double* dmu[PRB_MAX_MODELS];              // [ncols] dA/dmu_m   (NULL: value only)
  double* dvar[PRB_MAX_MODELS];             // [ncols] dA/dsigma_m^2
  int mtiles[PRB_MAX_MODELS];
};
cudaError_t launch_multi_acq_epilogue(const MultiAcqArgs& ma, int ncols_pad, int ncols, double* a, double* out_mean,
                                      double* out_var, cudaStream_t st);
cudaError_t launch_add_inplace(double* dst, const double* src, size_t count, cudaStream_t st);
// prb_chol.cu: GP-state builder (blocked Cholesky, divide-and-conquer triangular inverse, alpha, packing)
size_t chol_scratch_bytes(int n);
cudaError_t build_gp_state(const double* K, const double* noise, const double* y, double mean_const, int n, int Kp, int Mp1,
                           int Mp2, double* A1, double* A2, double* alpha, double* Linv_out, void* scratch, int* info_dev,
                           cudaStream_t st);
// prb_finalize.cu
cudaError_t launch_sample_candidates(const DevSpace& sp, const double* theta, int b, const double* u, uint64_t seed,
                                     uint64_t offset, double* out_x, double* out_xk, cudaStream_t st);
cudaError_t launch_best_row(const double* vals, const double* rows, int b, int d, double* out_best, int32_t* out_index,
                            cudaStream_t st);
cudaError_t launch_tr_bounds(const double* X, const double* Y, int n, int d, int n_con, double length, const double* lower,
                             const double* upper, const uint8_t* reset01, double* out_bounds, int32_t* out_center,
                             cudaStream_t st);
cudaError_t run_fp64_peak(int mode, int iters, double* scratch, int sms, double* tflops, cudaStream_t st);

}  // namespace prb
