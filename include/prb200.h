/*
 * prb200.h -- C ABI of the B200-native probabilistic-reparameterization (PR) acquisition hot path.
 *
 * The reference (facebookresearch/bo_pr) is pure Python on BoTorch/GPyTorch; it has no FFI.  The boundary this
 * library sits behind is BoTorch's object protocol.  Each entry point below names the
 * reference interface it replaces (paths relative to /root/reference/discrete_mixed_bo).  INTEGRATION.md shows
 * the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *  - every function returns 0 (PRB_OK) or a negative prb_status; nothing throws across the ABI
 *  - all tensor pointers are CALLER-OWNED DEVICE memory, contiguous, float64 unless stated, on the current device
 *  - the hot entry points never allocate device memory and never synchronise; work is enqueued on `stream` (a
 *    cudaStream_t passed as void*; NULL = legacy default stream); scratch comes from a caller-provided workspace whose
 *    size prb_workspace_bytes() reports.  prb_mc_adam / prb_analytic_adam replay one captured optimisation step as a CUDA
 *    graph: the FIRST call on a model (or the first with a different launch topology) instantiates the executable graph
 *    and caches it on the model; later calls update it in place (no instantiation, no synchronisation)
 *  - a prb_model is bound to the device that was current at creation; it is not thread-safe
 *  - there is NO CPU implementation behind this ABI
 */
#ifndef PRB200_H_
#define PRB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PRB_VERSION 200 /* 200: r02 layouts (prb_space_desc latent fields, prb_multi_acq_desc EHVI fields) */

#define PRB_MAX_DIMS 64     /* kernel-input columns / theta columns (svm: 50 binary + 3 continuous = 53) */
#define PRB_MAX_FACTORS 4   /* multiplicative (or additive) factors per term */
#define PRB_MAX_TERMS 2     /* additive ScaleKernel terms */
#define PRB_MAX_CAT 8       /* categorical parameters */
#define PRB_MAX_CARD 32     /* categories per categorical parameter */

typedef enum {
  PRB_OK = 0,
  PRB_ERR_INVALID = -1,      /* bad argument (NULL pointer, negative size, unsupported combination) */
  PRB_ERR_UNSUPPORTED = -2,  /* exceeds a compile-time limit above */
  PRB_ERR_WORKSPACE = -3,    /* workspace too small */
  PRB_ERR_CUDA = -4,         /* a CUDA runtime call failed; see prb_last_cuda_error() */
  PRB_ERR_NO_DEVICE = -5,    /* no sm_100 device / wrong device */
  PRB_ERR_NOT_PD = -6        /* K + diag(noise) is not positive definite to working precision; see prb_last_cuda_error() */
} prb_status;

/* ---- kernel structure: what get_kernel() builds (kernels.py:18-101) -------------------------------------------
 * K(x, x') = sum_t outputscale_t * combine_f( factor_f(x, x') ^ power_f ),  combine = product, or sum if additive.
 * factor kinds: gpytorch MaternKernel(nu=2.5) over `dims` with one lengthscale per dim (isotropic = repeated),
 *               botorch CategoricalKernel: exp(-mean_j( [x_j != x'_j] / lengthscale_j )).                          */
typedef enum { PRB_FACTOR_MATERN52 = 0, PRB_FACTOR_CATEGORICAL = 1 } prb_factor_kind;

typedef struct {
  int32_t kind;  /* prb_factor_kind */
  int32_t n_dims;
  int32_t power; /* 1, or 2 for the mixed_categorical double-multiply (kernels.py:90-94) */
  int32_t reserved;
  int32_t dims[PRB_MAX_DIMS];
  double lengthscale[PRB_MAX_DIMS];
} prb_factor;

typedef struct {
  double outputscale;
  int32_t additive;
  int32_t n_factors;
  prb_factor factors[PRB_MAX_FACTORS];
} prb_term;

/* ---- GP state: the caches GPyTorch's exact prediction strategy holds ----------------------- */
typedef struct {
  int32_t n;             /* training points */
  int32_t d_k;           /* kernel-input dimension (numeric layout if the model sees OneHotToNumeric inputs) */
  const double* X_train; /* [n, d_k] row-major */
  const double* Linv;    /* [n, n] row-major, lower triangle of inv(chol(K + diag(noise))); upper part ignored */
  const double* alpha;   /* [n] (K + diag(noise))^-1 (y - mean_const) */
  double mean_const;     /* ConstantMean */
  int32_t n_terms;
  int32_t reserved;
  prb_term terms[PRB_MAX_TERMS];
} prb_model_desc;

typedef struct prb_model prb_model;

/* ---- mixed search space: constructor arguments of the PR wrappers (probabilistic_reparameterization.py:102-112) */
typedef struct {
  int32_t d;      /* theta columns: [continuous..., integers..., one-hot blocks...] (problems/base.py:64-78) */
  int32_t n_int;
  int32_t n_cat;
  int32_t apply_numeric; /* 1: base AF sees OneHotToNumeric(tilde_x) (input.py:69-88) */
  int32_t int_cols[PRB_MAX_DIMS];
  double int_lo[PRB_MAX_DIMS]; /* integer_bounds[0] */
  double int_hi[PRB_MAX_DIMS]; /* integer_bounds[1] */
  int32_t cat_start[PRB_MAX_CAT]; /* first one-hot column of each categorical */
  int32_t cat_card[PRB_MAX_CAT];
  double tau;
  int32_t raw_integer_space; /* 1: theta's integer columns are already UN-normalised and outputs stay so (the bare
                                "round" member of the transform chain, input.py:637-746, without the Normalize pair) */
  int32_t reserved;
  /* Latent categorical embedding (kernel_type "mixed_latent": the model's LatentCategoricalEmbedding input transform,
   * input.py:777-802, 833-908, applied after OneHotToNumeric).  latent_tables == NULL: off.  Otherwise apply_numeric must
   * be 1 and categorical j's numeric column is replaced by row `category` of its table: the kernel input is
   * [non-categorical columns..., table_0 row, table_1 row, ...], d_k = cat_start[0] + sum_j latent_dim[j].
   * latent_tables: DEVICE pointer, the tables back to back in categorical order, table j [cat_card[j], latent_dim[j]]
   * row-major; it must stay valid for as long as calls use this descriptor.                                          */
  int32_t latent_dim[PRB_MAX_CAT];
  const double* latent_tables;
} prb_space_desc;

/* ---- base acquisition function (botorch analytic AFs; experiment_utils.py:391-393, 476-480, 489-492) ----------- */
typedef enum { PRB_ACQ_EI = 0, PRB_ACQ_UCB = 1, PRB_ACQ_MEAN = 2 } prb_acq_kind;

typedef struct {
  int32_t kind;          /* prb_acq_kind */
  int32_t ei_variant;    /* 0: sigma*(exp(logpdf(u)) + u*0.5*(1+erf(u/sqrt2))) (BoTorch <= 0.8); 1: erfc form */
  double param;          /* EI: best_f ; UCB: beta */
  double min_variance;   /* gpytorch posterior variance floor (1e-10 in float64) */
  double acq_var_floor;  /* EI: clamp_min on the variance before sqrt (1e-9 in BoTorch <= 0.8); others: 0 */
} prb_acq_desc;

typedef enum { PRB_EST_REINFORCE = 0, PRB_EST_REINFORCE_MA = 1 } prb_estimator;

/* ---- acquisition function over a LIST of independent single-output GPs (botorch ModelListGP, experiment_utils.py:337-346).
 * PRB_MACQ_CEI: botorch ConstrainedExpectedImprovement as get_EI builds it (experiment_utils.py:354-392):
 *   EI(objective output; best_f) * prod_m P(lower_m <= Y_m <= upper_m),  sigma_m = max(sqrt(var_m), sigma_floor),
 *   P from Normal.cdf(x) = 0.5 (1 + erf(x / sqrt2)); a missing bound drops its factor.
 * PRB_MACQ_EHVI: botorch's analytic ExpectedHypervolumeImprovement (q = 1) as get_EHVI builds it (experiment_utils.py:395-402)
 *   over a box decomposition of the non-dominated region (cell_bounds, the partitioning's get_hypercell_bounds()):
 *   sigma_m = sqrt(max(var_m, var_floor));  EHVI = sum_cells prod_m [psi(l, l) - psi(l, u) + nu(l, u)]_m,
 *   psi(l, u) = sigma pdf((u - mu)/sigma) + (mu - l) (1 - cdf((u - mu)/sigma)),  nu(l, u) = (u - l)(1 - cdf((u - mu)/sigma)),
 *   upper bounds clamped to 1e10 (the +inf cells).  Every model of the list is an objective.                              */
#define PRB_MAX_MODELS 8
typedef enum { PRB_MACQ_CEI = 0, PRB_MACQ_EHVI = 1 } prb_multi_acq_kind;
typedef struct {
  int32_t kind;            /* prb_multi_acq_kind */
  int32_t n_models;        /* outputs = models, in list order */
  int32_t objective_index; /* which model is the objective */
  int32_t reserved;
  double best_f;
  double min_variance;     /* gpytorch posterior variance floor (1e-10 in float64) */
  double sigma_floor;      /* botorch: sqrt(var).clamp_min(1e-9) */
  int32_t has_lower[PRB_MAX_MODELS];
  int32_t has_upper[PRB_MAX_MODELS];
  double lower[PRB_MAX_MODELS]; /* bounds on model m's output; ignored for the objective */
  double upper[PRB_MAX_MODELS];
  /* PRB_MACQ_EHVI only */
  double var_floor;          /* botorch: posterior.variance.clamp_min(1e-9) */
  int32_t n_cells;
  int32_t reserved2;
  const double* cell_bounds; /* DEVICE [2, n_cells, n_models] row-major: lower corners, then upper corners (+inf allowed) */
} prb_multi_acq_desc;

/* ---- library ---------------------------------------------------------------------------------------------------- */
int prb_version(void);
/* sizeof() of the descriptor structs as compiled into the library (0 prb_factor, 1 prb_term, 2 prb_model_desc,
 * 3 prb_space_desc, 4 prb_acq_desc, 5 prb_multi_acq_desc): lets a foreign-language binding verify its mirror of the layout.               */
int prb_abi_sizeof(int which);
const char* prb_status_string(int status);
/* Number of CUDA kernels this library has launched in this process (graph replays included).  Diagnostic: bench.py
 * reports the difference over its timed region as "gpu_launches".                                                   */
uint64_t prb_launch_count(void);

/* Optional per-kernel device timing (diagnostic; bench.py's roofline leg).  While enabled, every kernel launch of the
 * hot entry points is bracketed by CUDA events on the caller's stream (do not enable around prb_mc_adam: its step is
 * captured into a CUDA graph).  prb_profile_read synchronises on the recorded events, adds up milliseconds and launch
 * counts per slot and clears the record.  Slots: 0 sampler, 1 cross-covariance panel, 2 triangular product (lower:
 * Linv k*), 3 acquisition epilogue, 4 triangular product (upper: Linv^T v), 5 gradient contraction, 6 split sums,
 * 7 MC reduction, 8 EMA update, 9 analytic build/reduce, 10 Adam/clamp, 11 other.                                     */
/* Diagnostic: clock64() timestamps of CTA 0 at the phase boundaries of the last small-problem step kernel
 * (prb_small.cu): 0 start, 1 TMA issued, 2 prep done, 3 operand generated, 4 A1 landed, 5 lower product, 6 acquisition
 * epilogue, 7 upper product, 8 gradient contraction, 9 partial sums, 10 completion counter.  out16: host long long[16]. */
int prb_debug_small_timestamps(long long* out16);
/* Diagnostic: FP64 pipe microbenchmark on the current device (allocates, synchronises; never on a hot path).  mode 0:
 * mma.sync.m8n8k4.f64 from registers (the roofline denominator of the triangular products), 1: scalar DFMA, 2: both
 * interleaved warp by warp.  bench.py calls it outside its timed region and prints the result in `roofline.peak_live`.  */
int prb_debug_fp64_peak(int mode, double* out_tflops, void* stream);
#define PRB_PROFILE_SLOTS 12
int prb_profile_enable(int on);
int prb_profile_read(double* ms_by_kernel, uint64_t* launches_by_kernel, int32_t n_slots);
const char* prb_last_cuda_error(void);

/* Builds the device-side GP state (padded copies of Linv / Linv^T with alpha appended).  Allocates; call once per
 * BO iteration.  Replaces: the caches behind model.posterior() that acq_function(X) reads
 * (probabilistic_reparameterization.py:314, 491; [3P] gpytorch exact prediction strategy).                       */
int prb_model_create(prb_model** out, const prb_model_desc* desc, void* stream);
/* The same, starting from the DATA: K(X, X) + diag(noise) is assembled, factored (blocked Cholesky), inverted (triangular,
 * divide and conquer) and packed by the library's own kernels (FP64 DMMA block GEMMs; no cuSOLVER / cuBLAS).  desc->Linv and
 * desc->alpha are ignored.  y [n] targets, noise [n] observation variances (device).  out_Linv [n, n] (lower, zeros above) and
 * out_alpha [n] receive copies when not NULL.  Returns PRB_ERR_NOT_PD if the matrix is not positive definite to working
 * precision (prb_last_cuda_error() names the pivot).  Allocates scratch (3 n^2 doubles, released before returning) and
 * synchronises `stream` once.  Replaces: the prediction-strategy caches behind model.posterior, built from
 * experiment_utils.py:237-351's FixedNoiseGP ([3P] psd_safe_cholesky, explicit inverse root, mean_cache).                   */
int prb_model_create_from_data(prb_model** out, const prb_model_desc* desc, const double* y, const double* noise,
                               double* out_Linv, double* out_alpha, void* stream);
/* A deterministic posterior SAMPLE as the model (Thompson sampling, label `pr__ts`):
 *   f(x) = scale * sum_j weights[j] * cos(x . W[:, j] + bias[j]),   W [d_k, n_features] row-major (lengthscales folded in),
 * used with PRB_ACQ_MEAN in every entry point that takes a prb_model (other acquisition kinds return PRB_ERR_INVALID).
 * Replaces: the GenericDeterministicModel built by get_gp_samples / get_gp_sample_w_transforms (rffs.py:24-133) from
 * BoTorch's RandomFourierFeatures ([3P]) and evaluated under PosteriorMean (experiment_utils.py:481-492).  All pointers
 * are device memory and are copied.                                                                                   */
int prb_rff_model_create(prb_model** out, int32_t d_k, int32_t n_features, const double* W, const double* bias,
                         const double* weights, double scale, void* stream);
int prb_model_destroy(prb_model* model);
/* Per-model options.
 * PRB_OPT_DEDUP (default 0): in prb_mc_value_grad / prb_mc_adam on the separable kernel structure (Matern over continuous
 *   dims x isotropic Matern over <= 12 binary dims), evaluate the base acquisition function once per UNIQUE discrete
 *   configuration of each restart and weight it by its multiplicity.  Same result up to floating-point summation order;
 *   the work saved grows as the PR distribution concentrates (p -> 0/1).  Throughput with this option on must be reported
 *   separately from the per (candidate x MC sample) figure.                                                            */
#define PRB_OPT_DEDUP 1
/* PRB_OPT_UNFUSED (default 0): prb_mc_value_grad / prb_mc_screen take the stream-ordered multi-launch sequence (one kernel
 *   per stage) instead of the fused single-panel step.  Same arithmetic; kept as the cross-check path of the test-suite. */
#define PRB_OPT_UNFUSED 2
int prb_model_set_option(prb_model* model, int32_t option, int64_t value);
int prb_model_n(const prb_model* model);
int prb_model_dk(const prb_model* model);

/* Bytes of scratch needed to evaluate `n_points` (= b*mc, or b*n_discrete) base-AF points with this model, coming from
 * `n_restarts` = b rows of theta (0 for prb_acq_points).  `chunk_cols` = 0 picks the default L2-resident panel width. */
size_t prb_workspace_bytes(const prb_model* model, int64_t n_points, int32_t n_restarts, int32_t d_theta,
                           int32_t chunk_cols);

/* Counter-based Philox4x32-10 uniforms in [0,1): out[i, c] for i < mc, c < n_cols, 53-bit mantissa,
 * counter = (i, c, offset_lo, offset_hi), key = (seed_lo, seed_hi).  The fused sampler draws the same stream
 * on the fly.  Replaces: draw_sobol_samples(...) base samples (input.py:655-673, 696-713).                         */
int prb_philox_uniform(uint64_t seed, uint64_t offset, int32_t mc, int32_t n_cols, double* out, void* stream);

/* PR sampler, materialised (parity / debugging entry).  theta [b, d] normalised.  u_int [mc, n_int], u_cat [mc, n_cat]
 * shared by all b (NULL => Philox(seed, offset); integer columns use stream columns 0..n_int-1, categoricals
 * n_int..n_int+n_cat-1).
 * out_tilde_x [b, mc, d]   sampled configurations, normalised, one-hot layout       (input.py:637-746 + Normalize)
 * out_z       [b, mc, n_disc] uint8 rounding components used by the estimator        (probabilistic_reparameterization.py:454-465)
 * out_prob    [b, n_disc]  rounding probabilities                                    (input.py:613-635)
 * any output may be NULL.                                                                                            */
int prb_pr_sample(const prb_space_desc* space, const double* theta, int32_t b, int32_t mc, const double* u_int,
                  const double* u_cat, uint64_t seed, uint64_t offset, double* out_tilde_x, uint8_t* out_z,
                  double* out_prob, void* stream);

/* Final discrete draw per restart from p(theta*), on the device.
 * Replaces: AbstractProbabilisticReparameterization.sample_candidates (probabilistic_reparameterization.py:198-233).
 * theta [b, d] normalised.  u [b, n_int + n_cat] uniforms in [0, 1) (integer parameters first, then one per categorical),
 * or NULL => Philox(seed, offset) with counter (restart, parameter).  Integer k: z = [u < p_k] (torch.bernoulli's rule),
 * value = clamp(floor(x_un) + z) -- floor of the SIGNED un-normalised value, as :207-208 --, re-normalised; categorical:
 * inverse CDF of softmax((x - 0.5)/tau) (first index with cumsum > u; the last one if rounding leaves none), one-hot.
 * out_x [b, d] one-hot layout; out_xk [b, d_k] the rows as the base AF sees them (numeric layout if apply_numeric);
 * either may be NULL.                                                                                                 */
int prb_sample_candidates(const prb_space_desc* space, const double* theta, int32_t b, const double* u, uint64_t seed,
                          uint64_t offset, double* out_x, double* out_xk, void* stream);

/* sample_candidates + base-AF re-scoring of the sampled rows + argmax over the restarts, one enqueue, no host round trip.
 * Replaces: run_one_replication.py:500-528 (sample_candidates, one_hot_to_numeric, true_acq_func(...).max(dim=0),
 * candidates[best_idx]).  out_x [b, d]; out_acq [b] or NULL; out_best [1 + d] = (best value, best row) or NULL;
 * out_best_index device int32 or NULL (lowest index wins ties).  Workspace: prb_workspace_bytes(model, b, 0, d_k, 0).     */
int prb_finalize_candidates(prb_model* model, const prb_space_desc* space, const prb_acq_desc* acq, const double* theta,
                            int32_t b, const double* u, uint64_t seed, uint64_t offset, double* out_x, double* out_acq,
                            double* out_best, int32_t* out_best_index, void* workspace, size_t workspace_bytes,
                            void* stream);

/* Trust-region box around the incumbent, on the device.
 * Replaces: run_one_replication.py:410-440.  X [n, d], Y [n, 1 + n_constraints] (column 0 the objective, maximised; the
 * rest constraint slacks, >= 0 feasible).  Centre: best feasible row, else the least-violating one (:413-423).
 * out_bounds [2, d] = clamp(centre -/+ length * (upper - lower), lower, upper), columns with reset01[c] != 0 (binary dims,
 * one-hot blocks) reset to [0, 1]; out_center_index device int32 or NULL.                                              */
int prb_trust_region_bounds(const double* X, const double* Y, int32_t n, int32_t d, int32_t n_constraints, double length,
                            const double* lower, const double* upper, const uint8_t* reset01, double* out_bounds,
                            int32_t* out_center_index, void* stream);

/* Analytic-PR enumeration, materialised: x_all [b, n_options, d_k'] = every discrete configuration pasted next to the
 * continuous columns (normalised; one-hot layout when space->apply_numeric == 0, numeric layout otherwise) and
 * probs [b, n_options] = P(z | theta).  Replaces: AnalyticProbabilisticReparameterizationInputTransform.transform /
 * .get_probs (input.py:410-488).  Either output may be NULL.                                                         */
int prb_analytic_enumerate(const prb_space_desc* space, const double* theta, int32_t b, const int32_t* options,
                           int32_t n_options, double* out_x_all, double* out_probs, void* stream);

/* Raw cross-covariance K(xk, X_train) -> out [n_points, n] row-major, using only desc->{n, d_k, X_train, terms}.
 * Replaces: covar_module(X1, X2) ([3P] gpytorch kernels built by get_kernel, kernels.py:18-101); used by the GP-state
 * builder for K(X_train, X_train).                                                                                  */
int prb_kernel_matrix(const prb_model_desc* desc, const double* xk, int64_t n_points, double* out, void* stream);

/* Base acquisition function at explicit kernel-input points xk [n_points, d_k].
 * Replaces: acq_function(X) for ExpectedImprovement / UpperConfidenceBound / PosteriorMean on a FixedNoiseGP
 * (run_one_replication.py:513-528 re-scoring; initializers.py:186 screening of non-PR AFs).
 * out_acq / out_mean / out_var [n_points] (NULL to skip); out_dacq_dx [n_points, d_k] or NULL: gradient of the AF
 * value w.r.t. every kernel-input column (columns that only feed a categorical factor get 0).                       */
int prb_acq_points(prb_model* model, const prb_acq_desc* acq, const double* xk, int64_t n_points, double* out_acq,
                   double* out_mean, double* out_var, double* out_dacq_dx, void* workspace, size_t workspace_bytes,
                   int32_t chunk_cols, void* stream);

/* MC probabilistic reparameterization, value (+ gradient).
 * Replaces: _MCProbabilisticReparameterization.forward/backward (probabilistic_reparameterization.py:402-648).
 * theta [b, d]; u_int/u_cat as in prb_pr_sample.
 * ma_state: device double[2] = (ma_counter, ma_hidden), the EMA-baseline buffers of the wrapper (:185-196).
 *           The baseline uses the PRE-call state (:600-609); if update_ma != 0 the state is advanced in place
 *           exactly as :499-505 does (counter += 1; hidden -= 0.3 (hidden - mean(all acq values))).
 *           May be NULL for PRB_EST_REINFORCE with update_ma == 0.
 * grad_out [b] upstream gradient, or NULL for a value-only call (skips the second triangular product).
 * out_value [b]; out_grad [b, d] (required iff grad_out != NULL); out_acq_values [b, mc] or NULL.                   */
int prb_mc_value_grad(prb_model* model, const prb_space_desc* space, const prb_acq_desc* acq, const double* theta,
                      int32_t b, int32_t mc, const double* u_int, const double* u_cat, uint64_t seed,
                      uint64_t offset, int32_t estimator, double* ma_state, int32_t update_ma,
                      const double* grad_out, double* out_value, double* out_grad, double* out_acq_values,
                      void* workspace, size_t workspace_bytes, int32_t chunk_cols, void* stream);

/* The same call with HOST buffers: theta_host [b, d] is copied in, out_value_host [b] and out_grad_host [b, d] (unit upstream
 * gradient; required iff want_grad) are copied out, all enqueued on `stream` around the kernels -- one call, no device tensors
 * on the caller's side.  Pinned host memory makes the copies asynchronous; pageable memory works (the runtime stages it).
 * u_int / u_cat / ma_state stay DEVICE pointers (they are the wrapper's buffers).  synchronize != 0: the call returns after
 * the results have landed in the host buffers (the one entry point that synchronises, by request).  The device staging
 * lives in the caller's workspace (same prb_workspace_bytes).  If out_value_host == out_grad_host + b * d (gradient rows
 * and values are two views of one host block) both come back with a single device->host copy.                            */
int prb_mc_value_grad_host(prb_model* model, const prb_space_desc* space, const prb_acq_desc* acq,
                           const double* theta_host, int32_t b, int32_t mc, const double* u_int, const double* u_cat,
                           uint64_t seed, uint64_t offset, int32_t estimator, double* ma_state, int32_t update_ma,
                           int32_t want_grad, double* out_value_host, double* out_grad_host, void* workspace,
                           size_t workspace_bytes, int32_t synchronize, void* stream);

/* The same call over a model list: every model contributes its own posterior (its own k*, Cholesky factor and -- for the
 * pathwise gradient -- its own second product); the acquisition epilogue combines them by the product rule.
 * Replaces: _MCProbabilisticReparameterization.forward/backward around ConstrainedExpectedImprovement on a ModelListGP
 * (experiment_utils.py:354-392; the welded-beam / pressure-vessel style configs).  All models must share d_k and n.
 * Workspace: prb_workspace_bytes_multi.  One column panel only (PRB_ERR_UNSUPPORTED beyond ~10^5 evaluation points).        */
size_t prb_workspace_bytes_multi(prb_model* const* models, int32_t n_models, int64_t n_points, int32_t n_restarts,
                                 int32_t d_theta);
int prb_mc_value_grad_multi(prb_model* const* models, int32_t n_models, const prb_space_desc* space,
                            const prb_multi_acq_desc* acq, const double* theta, int32_t b, int32_t mc,
                            const double* u_int, const double* u_cat, uint64_t seed, uint64_t offset, int32_t estimator,
                            double* ma_state, int32_t update_ma, const double* grad_out, double* out_value,
                            double* out_grad, double* out_acq_values, void* workspace, size_t workspace_bytes,
                            void* stream);
/* Model-list acquisition value at explicit kernel-input points xk [n_points, d_k] (re-scoring of sampled candidates,
 * run_one_replication.py:513-528).  out_acq [n_points]; out_mean / out_var [n_points, n_models] or NULL.                    */
int prb_acq_points_multi(prb_model* const* models, int32_t n_models, const prb_multi_acq_desc* acq, const double* xk,
                         int64_t n_points, double* out_acq, double* out_mean, double* out_var, void* workspace,
                         size_t workspace_bytes, void* stream);

/* Raw-sample screening: the value-only MC-PR forward of prb_mc_value_grad applied to theta [n_rows, d] in consecutive chunks
 * of `chunk_rows` rows, one enqueue after the other with no host work in between.  Each chunk counts as ONE forward call for
 * the EMA-baseline state, exactly as the chunked loop it replaces.
 * Replaces: the init_batch_limit loop of gen_batch_initial_conditions (initializers.py:179-191).
 * out_value [n_rows].  The workspace must cover prb_workspace_bytes(model, chunk_rows * mc, chunk_rows, d, chunk_cols).   */
int prb_mc_screen(prb_model* model, const prb_space_desc* space, const prb_acq_desc* acq, const double* theta,
                  int32_t n_rows, int32_t chunk_rows, int32_t mc, const double* u_int, const double* u_cat, uint64_t seed,
                  uint64_t offset, int32_t estimator, double* ma_state, double* out_value, void* workspace,
                  size_t workspace_bytes, int32_t chunk_cols, void* stream);

/* Analytic probabilistic reparameterization: sum over all discrete configurations.
 * Replaces: AnalyticProbabilisticReparameterization.forward + autograd (probabilistic_reparameterization.py:285-320,
 * input.py:410-488).  options [n_options, n_int + n_cat] int32: un-normalised integer value per integer parameter,
 * category index per categorical (the rows of all_discrete_options, input.py:345-383).
 * out_value [b]; out_grad [b, d] if grad_out != NULL; out_probs [b, n_options] or NULL.                             */
int prb_analytic_value_grad(prb_model* model, const prb_space_desc* space, const prb_acq_desc* acq,
                            const double* theta, int32_t b, const int32_t* options, int32_t n_options,
                            const double* grad_out, double* out_value, double* out_grad, double* out_probs,
                            double* out_acq_values, void* workspace, size_t workspace_bytes, int32_t chunk_cols,
                            void* stream);

/* Multi-start Adam on the MC-PR objective, entirely on device.
 * Replaces: gen_candidates_torch (gen.py:264-343) for a PR acquisition function: clamp -> value+grad -> Adam(lr,
 * betas (0.9, 0.999), eps 1e-8) repeated `maxiter` times, final clamp and value-only evaluation.
 * theta [b, d] in/out; lower/upper [d]; adam_state: device double[2*b*d] zero-initialised by the caller (exp_avg,
 * exp_avg_sq); `step0` = number of Adam steps already taken (bias correction).  out_value [b] final values.          */
int prb_mc_adam(prb_model* model, const prb_space_desc* space, const prb_acq_desc* acq, double* theta, int32_t b,
                int32_t mc, const double* u_int, const double* u_cat, uint64_t seed, uint64_t offset,
                int32_t estimator, double* ma_state, const double* lower, const double* upper, double lr,
                int32_t maxiter, int32_t step0, double* adam_state, double* out_value, void* workspace,
                size_t workspace_bytes, int32_t chunk_cols, void* stream);

/* Multi-start Adam on the ANALYTIC PR objective, on device (same loop as prb_mc_adam around prb_analytic_value_grad).
 * Replaces: gen_candidates_torch (gen.py:264-343) for an AnalyticProbabilisticReparameterization, which every pr__* label
 * optimises with Adam too (run_one_replication.py:176-183).                                                          */
int prb_analytic_adam(prb_model* model, const prb_space_desc* space, const prb_acq_desc* acq, double* theta, int32_t b,
                      const int32_t* options, int32_t n_options, const double* lower, const double* upper, double lr,
                      int32_t maxiter, int32_t step0, double* adam_state, double* out_value, void* workspace,
                      size_t workspace_bytes, int32_t chunk_cols, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PRB200_H_ */
