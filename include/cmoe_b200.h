/* cornell-moe_b200 — C ABI of the B200-native GP-posterior + Monte-Carlo acquisition path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / CUDA types.  Every entry point
 * states which interface of the reference (wujian16/Cornell-MOE, paths relative to
 * moe/optimal_learning/cpp/) it replaces.  The pybind11 module `GPP` in cornell-moe_b200/csrc/gpp_module.cpp
 * re-exports the reference's `moe.build.GPP` names on top of exactly these calls; INTEGRATION.md shows the
 * binding a reference maintainer would add.
 *
 * Conventions are the reference's own (gpp_common.hpp:29,390; gpp_math.hpp:301-305):
 *   - all reals are IEEE double, all matrices column-major, point sets are [num_points][dim];
 *   - observations are interleaved per point: (value, d/dx_{derivatives[0]}, ...);
 *   - noise_variance has 1+num_derivatives entries and is indexed by observation TYPE, not by point
 *     (gpp_math.cpp:447-449);
 *   - host pointers in, host pointers out, no ownership transfer (gpp_python_common.cpp:37, 79-128).
 *
 * Errors: every function returns a status code; CMOE_ERR_* map 1:1 onto the reference's exception
 * classes (gpp_exception.hpp:170-509, gpp_python.cpp:189-206).  `info` receives the payload
 * (leading-minor index k+1 for CMOE_ERR_SINGULAR, exactly as gpp_linear_algebra.cpp:141-142 returns it).
 * There is NO CPU fallback: without a CUDA device every compute call returns CMOE_ERR_NO_DEVICE.
 *
 * Threading: like the reference's Python boundary (GIL held for the whole call) a handle is used by one host thread at a
 * time; calls on the same cmoe_gp are issued on that handle's CUDA stream and must be serialised by the caller.
 * Repeated cmoe_kg_eval calls with an unchanged configuration reuse a device workspace cached in the handle.
 */
#ifndef CMOE_B200_H_
#define CMOE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMOE_OK 0
#define CMOE_ERR_SINGULAR 1      /* SingularMatrixException   (gpp_exception.hpp:465) */
#define CMOE_ERR_BOUNDS 2        /* BoundsException           (gpp_exception.hpp:243) */
#define CMOE_ERR_INVALID_VALUE 3 /* InvalidValueException     (gpp_exception.hpp:382) */
#define CMOE_ERR_RUNTIME 4       /* OptimalLearningException  (gpp_exception.hpp:170): CUDA / internal failure */
#define CMOE_ERR_NO_DEVICE 5     /* OptimalLearningException: no CUDA device; there is no CPU path */

#define CMOE_KERNEL_SQUARE_EXPONENTIAL 0 /* gpp_covariance.hpp:195 */
#define CMOE_KERNEL_MATERN_NU_2P5 1      /* gpp_covariance.hpp:313 */

#define CMOE_MAX_DIM 32

typedef struct cmoe_gp cmoe_gp; /* opaque; replaces optimal_learning::GaussianProcess (gpp_math.hpp:275) */

/* GradientDescentParameters, gpp_optimizer_parameters.hpp:81-135 (same field order as its constructor) */
typedef struct cmoe_gd_params {
  int num_multistarts;
  int max_num_steps;
  int max_num_restarts;
  int num_steps_averaged;
  double gamma;
  double pre_mult;
  double max_relative_change;
  double tolerance;
} cmoe_gd_params;

/* Work counters reported by the fused q-KG kernel (it counts what it actually executed). */
typedef struct cmoe_kg_stats {
  uint64_t mc_samples;         /* candidates * num_mc */
  uint64_t posterior_evals;    /* posterior-mean evaluations the reference's line search would have made for the
                                  same trajectory (trial points + domain-limited points + start points) */
  uint64_t line_search_steps;  /* accepted inner gradient steps */
  uint64_t point_evals;        /* executed: value+gradient evaluations at one query point */
  uint64_t line_batches;       /* executed: batched backtracking passes (SquareExponential fast path; all trial step
                                  sizes of one step share one pass over the training points) */
} cmoe_kg_stats;

/* Human-readable description of the last error on the calling thread. */
const char* cmoe_last_error(void);
const char* cmoe_version(void);
/* Number of visible CUDA devices (0 if none / driver missing). */
int cmoe_device_count(void);

/* ---- GaussianProcess --------------------------------------------------------------------------------------
 * cmoe_gp_create  replaces GaussianProcess::GaussianProcess (gpp_math.cpp:553-573) == covariance build
 *   (gpp_math.cpp:426-455) + Cholesky (gpp_linear_algebra.cpp:109-148) + K^-1 (y - mean) (gpp_math.cpp:481-511);
 *   Python boundary: make_gaussian_process, gpp_python_gaussian_process.cpp:42-62.
 *   hyperparameters = (alpha, lengths[dim]).  device = CUDA ordinal (the reference's dead `which_gpu`).
 *   On a singular K: returns CMOE_ERR_SINGULAR, *info = leading minor index, *gp_out = NULL. */
int cmoe_gp_create(int kernel, double alpha, const double* lengths, const double* points_sampled,
                   const double* points_sampled_value, const double* noise_variance, const int* derivatives,
                   int num_derivatives, int dim, int num_sampled, int device, cmoe_gp** gp_out, int* info);
/* log p(y | X, theta) = -1/2 (y-m)^T K^-1 (y-m) - sum_i log L_ii - n/2 log(2 pi) with K = K(X,X) + diag(noise by type)
 * + 1e-6 I: LogMarginalLikelihoodEvaluator::ComputeLogLikelihood with FillLogLikelihoodState
 * (gpp_model_selection.cpp:540-612); Python boundary compute_log_likelihood (gpp_python_model_selection.cpp:43-87) —
 * the call the hyper-parameter MCMC of the front end makes thousands of times per refit.  One GP fit on the device
 * (covariance build, Cholesky, K^-1 y) plus two reductions over n numbers.  A singular K yields -inf (the reference
 * ignores the failed factorisation, :551-553).  SURVEY.md 8f rank 2, value only (no hyper-parameter gradient yet). */
int cmoe_log_marginal_likelihood(int kernel, double alpha, const double* lengths, const double* points_sampled,
                                 const double* points_sampled_value, const double* noise_variance,
                                 const int* derivatives, int num_derivatives, int dim, int num_sampled, int device,
                                 double* log_likelihood, int* info);
/* d log p / d (alpha, l_1..l_dim, noise_0..noise_g): LogMarginalLikelihoodEvaluator::ComputeGradLogLikelihood
 * (gpp_model_selection.cpp:629-690) with HyperparameterGradCovariance (gpp_covariance.cpp:245-317, 461-489); Python
 * boundary compute_hyperparameter_grad_log_likelihood (gpp_python_model_selection.cpp:89-140).  grad[dim + 2 + g].
 * One device fit, K^-1 by the blocked solves, one fused 1/2 tr[(a a^T - K^-1) dK/d theta] contraction kernel.  Keeps the
 * reference's Matern quirk (only the value-value entry of a derivative block contributes).  Singular K: zeros. */
int cmoe_grad_log_marginal_likelihood(int kernel, double alpha, const double* lengths, const double* points_sampled,
                                      const double* points_sampled_value, const double* noise_variance,
                                      const int* derivatives, int num_derivatives, int dim, int num_sampled, int device,
                                      double* grad, int* info);
/* Run-time switches between kernel generations (also read once from the environment: CMOE_LEGACY_LINALG, CMOE_COV_TMA):
 *   "legacy_linalg" 1 = round-1 launch-per-step Cholesky / chained trsv for large n (default 0: cooperative kernels);
 *   "cov_tma"       1 = TMA / DMMA covariance build with tensor-map tile stores (default 0: LDGSTS kernel, faster at d ~ 10). */
int cmoe_set_option(const char* name, int value);
void cmoe_gp_destroy(cmoe_gp* gp);
int cmoe_gp_dim(const cmoe_gp* gp);
int cmoe_gp_device(const cmoe_gp* gp);
int cmoe_gp_num_sampled(const cmoe_gp* gp);
int cmoe_gp_num_derivatives(const cmoe_gp* gp);
/* Copies out K_chol_ (n*n, lower triangle valid), K_inv_y_ (n) and mean_ (gpp_math.hpp:838-867); any may be NULL. */
int cmoe_gp_get_state(const cmoe_gp* gp, double* K_chol, double* K_inv_y, double* mean);
/* GaussianProcess::AddPointsToGP (gpp_math.cpp:1699-1718): append and refit. */
int cmoe_gp_add_sampled_points(cmoe_gp* gp, const double* new_points, const double* new_points_value,
                               int num_new_points, int* info);
/* Timings of the last fit, microseconds of device time: {covariance build, Cholesky, K^-1 y solve}. */
int cmoe_gp_fit_timings(const cmoe_gp* gp, double* usec3);

/* ---- posterior queries, batched over point sets -------------------------------------------------------------
 * Replaces ComputeMeanOfPoints (gpp_math.cpp:662), ComputeGradMeanOfPoints (:721), ComputeVarianceOfPoints (:924),
 * ComputeGradVarianceOfPoints (:1366), ComputeGradCholeskyVarianceOfPoints (:1466) and the q*q Cholesky; Python
 * boundary gpp_python_gaussian_process.cpp:64-236.  `sets` = [num_sets][num_pts][dim]; each point carries the
 * derivative rows derivs_s[g_s] (Q = num_pts*(1+g_s)).  Outputs per set, reference layouts:
 *   mean[Q]; grad_mean[Q][dim] (d fastest); var[Q*Q] col-major (full symmetric); chol_var[Q*Q] (lower, upper zeroed);
 *   grad_var / grad_chol [num_pts][Q][Q][dim] (d fastest; grad_chol in the reference's transposed storage,
 *   gpp_math.cpp:1416-1417).  Any output may be NULL.  chol failure -> CMOE_ERR_SINGULAR, *info = index. */
int cmoe_gp_posterior(const cmoe_gp* gp, const double* sets, int num_sets, int num_pts, const int* derivs_s, int g_s,
                      double* mean, double* grad_mean, double* var, double* chol_var, double* grad_var,
                      double* grad_chol, int* info);

/* ---- q-EI Monte Carlo, batched over candidates ----------------------------------------------------------------
 * Replaces ExpectedImprovementEvaluator::ComputeExpectedImprovement / ComputeGradExpectedImprovement
 * (gpp_math.cpp:1991-2126); Python boundary compute_expected_improvement / compute_grad_expected_improvement /
 * evaluate_EI_at_point_list (gpp_python_expected_improvement.cpp:44-109, 401-441).
 * candidates = [num_candidates][q][dim]; points_being_sampled = [p][dim] shared by all candidates.
 * Normals: Philox4x32-10 + Box-Muller keyed by `seed` (same stream for every candidate = common random numbers,
 * the device analogue of ResetToMostRecentSeed, gpp_random.cpp:130-135), or, if normals_table != NULL, the table
 * is replayed exactly like NormalRNGSimulator (gpp_random.hpp:314): (q+p) normals per iteration.
 * ei[num_candidates]; grad_ei[num_candidates][q][dim] or NULL. */
int cmoe_ei_eval(const cmoe_gp* gp, const double* candidates, int num_candidates, int q,
                 const double* points_being_sampled, int p, int num_mc, double best_so_far, uint64_t seed,
                 const double* normals_table, double* ei, double* grad_ei, int* info);

/* ---- q-KG / d-KG Monte Carlo, batched over candidates -----------------------------------------------------------
 * Replaces KnowledgeGradientEvaluator::ComputeKnowledgeGradient / ComputeGradKnowledgeGradient
 * (gpp_knowledge_gradient_optimization.cpp:69-227) including KnowledgeGradientState::PreCompute (:292-317),
 * ComputeOptimalPosteriorMean (:420-472), the line-search gradient descent (gpp_optimization.hpp:708-828) and
 * TensorProductDomain::LimitUpdate (gpp_domain.cpp:64-104); Python boundary compute_knowledge_gradient /
 * compute_grad_knowledge_gradient / evaluate_KG_at_point_list (gpp_python_knowledge_gradient.cpp:77-154, 352-397).
 * inner_bounds = [dim-num_fidelity][2]; discrete_pts = [num_pts][dim-num_fidelity].
 * Normals as for EI but (q+p)(1+g) per EVEN iteration; odd iterations are antithetic (:88-97).
 * kg[num_candidates]; grad_kg[num_candidates][q][dim] or NULL; stats may be NULL. */
int cmoe_kg_eval(const cmoe_gp* gp, int num_fidelity, const cmoe_gd_params* inner, const double* inner_bounds,
                 const double* discrete_pts, int num_pts, const double* candidates, int num_candidates, int q,
                 const double* points_being_sampled, int p, int num_mc, double best_so_far, uint64_t seed,
                 const double* normals_table, double* kg, double* grad_kg, cmoe_kg_stats* stats, int* info);

/* ---- device-resident plan for the q-KG path (what bench.py times as `value`) ------------------------------------
 * Same computation as cmoe_kg_eval, split so that inputs can be made resident in HBM before the timed region:
 *   create -> upload (H2D of candidates) -> run (kernels only, asynchronous on the plan's stream) -> download (D2H).
 * cmoe_kg_plan_elapsed_ms returns CUDA-event time of the last run; cmoe_kg_plan_kernel_ms the time spent in the
 * fused MC kernel alone (events recorded around it on the launching stream); *_launches the kernel launches issued. */
typedef struct cmoe_kg_plan cmoe_kg_plan;
int cmoe_kg_plan_create(const cmoe_gp* gp, int num_fidelity, const cmoe_gd_params* inner, const double* inner_bounds,
                        const double* discrete_pts, int num_pts, int max_candidates, int q,
                        const double* points_being_sampled, int p, int num_mc, double best_so_far, uint64_t seed,
                        int want_grad, cmoe_kg_plan** plan_out);
void cmoe_kg_plan_destroy(cmoe_kg_plan* plan);
/* Optional table replay (NormalRNGSimulator semantics): ((num_mc+1)/2)*(q+p) normals; replaces the Philox stream. */
int cmoe_kg_plan_set_table(cmoe_kg_plan* plan, const double* normals_table, int table_len);
/* Reference-driver compatibility: the multistart drivers of the reference evaluate every start and every descent step
 * through ONE KnowledgeGradientState per thread, constructed with the first start; SetCurrentPoint refreshes the GP
 * quantities but NOT discretized_set (gpp_knowledge_gradient_optimization.cpp:233-243 vs :259-261), so the set the
 * per-sample inner optimiser starts from keeps the FIRST start's q points instead of the current ones.  Passing those
 * q points here ([q][dim]; NULL clears) makes a plan evaluate exactly that; cmoe_multistart_kg(_ex) does it itself. */
int cmoe_kg_plan_set_stale_union(cmoe_kg_plan* plan, const double* points_to_sample);
int cmoe_kg_plan_upload(cmoe_kg_plan* plan, const double* candidates, int num_candidates);
int cmoe_kg_plan_run(cmoe_kg_plan* plan);
int cmoe_kg_plan_sync(cmoe_kg_plan* plan, int* info);
int cmoe_kg_plan_download(cmoe_kg_plan* plan, double* kg, double* grad_kg, cmoe_kg_stats* stats);
int cmoe_kg_plan_timings(const cmoe_kg_plan* plan, double* total_ms, double* mc_kernel_ms, int* launches);

/* ---- multistart optimisation (the data-parallel axis) -------------------------------------------------------------
 * Replaces ComputeKGOptimalPointsToSampleViaMultistartGradientDescent (gpp_knowledge_gradient_optimization.hpp:859-935):
 * KG at every start -> keep the best 20 (hard-coded `k = 20`, :901) -> restarted gradient descent with LimitUpdate on
 * each (gpp_optimization.hpp:620-705, 1144-1185) -> strict-> argmax (gpp_optimization.hpp:1511, 1540).
 * Python boundary multistart_knowledge_gradient_optimization (gpp_python_knowledge_gradient.cpp:243-304).
 * This entry point runs the whole pipeline on one GPU.  With one process per GPU the same pipeline is sharded over
 * the starts by cornell_moe_b200/multigpu.py: cmoe_kg_eval on this rank's starts -> one all-gather of the values ->
 * identical top-20 on every rank -> cmoe_kg_gradient_descent on this rank's share -> one all-gather -> arg-max.
 * start_values (may be NULL) receives KG at every start.
 * best_value / best_point[q*dim] / found_flag follow OptimizationIOContainer (gpp_optimization.hpp:511). */
int cmoe_multistart_kg(const cmoe_gp* gp, int num_fidelity, const cmoe_gd_params* outer, const cmoe_gd_params* inner,
                       const double* domain_bounds, const double* inner_bounds, const double* discrete_pts, int num_pts,
                       const double* starts, int num_starts, int q, const double* points_being_sampled, int p,
                       int num_mc, double best_so_far, uint64_t seed, double* start_values, double* best_point,
                       double* best_value, int* found_flag, int* info);
/* EI twin: ComputeOptimalPointsToSampleViaMultistartGradientDescent (gpp_math.hpp:1683-1802);
 * Python boundary multistart_expected_improvement_optimization (gpp_python_expected_improvement.cpp:221-276). */
int cmoe_multistart_ei(const cmoe_gp* gp, const cmoe_gd_params* outer, const double* domain_bounds,
                       const double* starts, int num_starts, int q, const double* points_being_sampled, int p,
                       int num_mc, double best_so_far, uint64_t seed, double* start_values, double* best_point,
                       double* best_value, int* found_flag, int* info);

/* Extended forms of the two drivers above (opts may be NULL = plain call):
 *   normals_table / table_len — standard normals replayed for every evaluation exactly like the reference's
 *     NormalRNG after ResetToMostRecentSeed (every q-KG / q-EI evaluation of a driver call rewinds its generator,
 *     gpp_knowledge_gradient_optimization.cpp:81,164, gpp_math.cpp:2011,2076): ((num_mc+1)/2)*(q+p)*(1+g) values for
 *     q-KG, num_mc*(q+p) for q-EI.  With the table drawn from a NormalRNG of the same seed the device path and
 *     ComputeKGOptimalPointsToSampleViaMultistartGradientDescent consume identical draws.
 *   devices / num_devices — the reference's parallel axis (OpenMP threads over starts, gpp_optimization.hpp:1472-1546)
 *     inside the call: the GP is replicated on every listed GPU (one host thread and stream set per device), starts
 *     are strided over the devices for the screening pass and for the gradient descent of the kept 20, values are
 *     gathered on the host.  Every start is evaluated by the same kernels with the same seed wherever it runs, so the
 *     result is bit-identical to the single-device call. */
typedef struct cmoe_multistart_opts {
  const double* normals_table;
  size_t table_len;
  const int* devices;
  int num_devices;
  int domain_type; /* CMOE_DOMAIN_TENSOR_PRODUCT (0) or CMOE_DOMAIN_SIMPLEX (1): the outer optimiser's domain */
  /* q-KG only.  0 (default): bug-compatible with the reference driver — the inner optimiser's discretisation set keeps
   * the q points of `stale_union` (NULL = the first start), see cmoe_kg_plan_set_stale_union.  1: every evaluation uses
   * its own current points (what a freshly constructed state, e.g. compute_knowledge_gradient, does). */
  int fresh_discretisation;
  const double* stale_union; /* [q][dim] or NULL */
} cmoe_multistart_opts;
/* DomainTypes, gpp_python_common.cpp:201-240.  CMOE_DOMAIN_SIMPLEX = unit simplex intersected with the box
 * (SimplexIntersectTensorProductDomain, gpp_domain.cpp:107-289): implemented for the q-EI drivers; q-KG needs the same
 * domain for its per-sample inner optimiser, which the fused kernel only has for the tensor product — it reports
 * CMOE_ERR_INVALID_VALUE. */
#define CMOE_DOMAIN_TENSOR_PRODUCT 0
#define CMOE_DOMAIN_SIMPLEX 1
/* One LimitUpdate of the outer optimiser's domain (host arithmetic, no device needed): update[dim] is limited in place.
 * TensorProductDomain::LimitUpdate gpp_domain.cpp:64-104 / SimplexIntersectTensorProductDomain::LimitUpdate :234-289. */
int cmoe_limit_update(int domain_type, const double* domain_bounds, int dim, double max_relative_change,
                      const double* current_point, double* update);
int cmoe_multistart_kg_ex(const cmoe_gp* gp, int num_fidelity, const cmoe_gd_params* outer, const cmoe_gd_params* inner,
                          const double* domain_bounds, const double* inner_bounds, const double* discrete_pts,
                          int num_pts, const double* starts, int num_starts, int q, const double* points_being_sampled,
                          int p, int num_mc, double best_so_far, uint64_t seed, const cmoe_multistart_opts* opts,
                          double* start_values, double* best_point, double* best_value, int* found_flag, int* info);
int cmoe_multistart_ei_ex(const cmoe_gp* gp, const cmoe_gd_params* outer, const double* domain_bounds,
                          const double* starts, int num_starts, int q, const double* points_being_sampled, int p,
                          int num_mc, double best_so_far, uint64_t seed, const cmoe_multistart_opts* opts,
                          double* start_values, double* best_point, double* best_value, int* found_flag, int* info);

/* Restarted gradient descent from given starts only (the second half of the multistart drivers); used by the
 * multi-GPU host layer after the global top-20 has been agreed on.  values_out[num_starts], points_out[num_starts][q*dim]. */
/* (cmoe_kg_gradient_descent_ex: same, `stale_union` ([q][dim] or NULL) as in cmoe_multistart_opts — the sharded host
 * layer passes the global first start so that every rank reproduces the one-process driver.) */
int cmoe_kg_gradient_descent(const cmoe_gp* gp, int num_fidelity, const cmoe_gd_params* outer,
                             const cmoe_gd_params* inner, const double* domain_bounds, const double* inner_bounds,
                             const double* discrete_pts, int num_pts, const double* starts, int num_starts, int q,
                             const double* points_being_sampled, int p, int num_mc, double best_so_far, uint64_t seed,
                             double* values_out, double* points_out, int* info);
int cmoe_kg_gradient_descent_ex(const cmoe_gp* gp, int num_fidelity, const cmoe_gd_params* outer,
                                const cmoe_gd_params* inner, const double* domain_bounds, const double* inner_bounds,
                                const double* discrete_pts, int num_pts, const double* starts, int num_starts, int q,
                                const double* points_being_sampled, int p, int num_mc, double best_so_far,
                                uint64_t seed, const double* stale_union, double* values_out, double* points_out,
                                int* info);
int cmoe_ei_gradient_descent(const cmoe_gp* gp, const cmoe_gd_params* outer, const double* domain_bounds,
                             const double* starts, int num_starts, int q, const double* points_being_sampled, int p,
                             int num_mc, double best_so_far, uint64_t seed, double* values_out, double* points_out,
                             int* info);

/* Closed-form one-point EI and its gradient at num_points points (OnePotentialSampleExpectedImprovementEvaluator,
 * gpp_math.cpp:2196-2253) — what EvaluateEIAtPointList and the multistart driver use when q = 1, p = 0
 * (gpp_math.cpp:2317, gpp_math.hpp:1703-1749).  values[num_points]; grads[num_points][dim] or NULL. */
int cmoe_ei_analytic(const cmoe_gp* gp, const double* points, int num_points, double best_so_far, double* values,
                     double* grads, int* info);

/* ---- ensembles of GPs: "MCMC-averaged" acquisition, one GP per hyper-parameter sample ----------------------------
 * The reference's GaussianProcessMCMC is a list of GPs over the same data (gpp_knowledge_gradient_mcmc_optimization.cpp:
 * 24-48); here it is an array of cmoe_gp handles (same dim / derivative observations / device).
 * discrete_pts[num_gp][num_pts][dim-num_fidelity] and best_so_far[num_gp] are per member, as in the reference.
 *
 * cmoe_kg_eval_mcmc      KnowledgeGradientMCMCEvaluator::Compute{,Grad}KnowledgeGradient (..mcmc_optimization.cpp:137-180)
 *                        = mean_m KG_m / cost,  cost = max_i prod_{j>=dim-nf} x_ij (1 when num_fidelity = 0, :87-104);
 *                        gradient by the quotient rule (:162-180).  Python: compute_knowledge_gradient_mcmc,
 *                        compute_grad_knowledge_gradient_mcmc, evaluate_KG_mcmc_at_point_list
 *                        (gpp_python_knowledge_gradient_mcmc.cpp:80-198, 326-384).
 * cmoe_ei_eval_mcmc      ExpectedImprovementMCMCEvaluator (gpp_expected_improvement_mcmc_optimization.cpp:47-85):
 *                        mean_m of the Monte-Carlo q-EI; with analytic_single != 0 and q = 1, p = 0 the closed-form
 *                        1-EI per member instead, as EvaluateEIMCMCAtPointList does (..mcmc_optimization.cpp:251).  Python: compute_expected_improvement_mcmc, ..grad.., evaluate_EI_mcmc_at_point_list.
 * cmoe_multistart_*_mcmc ComputeKGMCMCOptimalPointsToSampleViaMultistartGradientDescent (..mcmc_optimization.hpp:665-745)
 *                        and ComputeEIMCMC... (gpp_expected_improvement_mcmc_optimization.hpp:860-980): screen, top-20,
 *                        restarted gradient descent, strict-> arg-max (initial best -inf for KG, 0.0 for EI; the EI
 *                        driver uses analytic 1-EI per member when q = 1, p = 0).
 * normals_table: as in cmoe_kg_eval / cmoe_ei_eval (NULL = Philox stream keyed by seed); shared by every member. */
int cmoe_kg_eval_mcmc(const cmoe_gp* const* gps, int num_gp, int num_fidelity, const cmoe_gd_params* inner,
                      const double* inner_bounds, const double* discrete_pts, int num_pts, const double* candidates,
                      int num_candidates, int q, const double* points_being_sampled, int p, int num_mc,
                      const double* best_so_far, uint64_t seed, const double* normals_table, double* values,
                      double* grads, int* info);
int cmoe_ei_eval_mcmc(const cmoe_gp* const* gps, int num_gp, const double* candidates, int num_candidates, int q,
                      const double* points_being_sampled, int p, int num_mc, const double* best_so_far, uint64_t seed,
                      const double* normals_table, int analytic_single, double* values, double* grads, int* info);
int cmoe_multistart_kg_mcmc(const cmoe_gp* const* gps, int num_gp, int num_fidelity, const cmoe_gd_params* outer,
                            const cmoe_gd_params* inner, const double* domain_bounds, const double* inner_bounds,
                            const double* discrete_pts, int num_pts, const double* starts, int num_starts, int q,
                            const double* points_being_sampled, int p, int num_mc, const double* best_so_far,
                            uint64_t seed, double* start_values, double* best_point, double* best_value,
                            int* found_flag, int* info);
int cmoe_multistart_ei_mcmc(const cmoe_gp* const* gps, int num_gp, const cmoe_gd_params* outer,
                            const double* domain_bounds, const double* starts, int num_starts, int q,
                            const double* points_being_sampled, int p, int num_mc, const double* best_so_far,
                            uint64_t seed, double* start_values, double* best_point, double* best_value,
                            int* found_flag, int* info);

/* The restarted gradient descent of the MCMC drivers alone (RestartedGradientDescentKGMCMCOptimization,
 * gpp_knowledge_gradient_mcmc_optimization.hpp:576-600, and its EI twin): every start is optimised, no top-20 selection.
 * values_out[num_starts], points_out[num_starts][q][dim].  Used by the multi-GPU layer, which shards the kept starts. */
int cmoe_kg_gradient_descent_mcmc(const cmoe_gp* const* gps, int num_gp, int num_fidelity, const cmoe_gd_params* outer,
                                  const cmoe_gd_params* inner, const double* domain_bounds, const double* inner_bounds,
                                  const double* discrete_pts, int num_pts, const double* starts, int num_starts, int q,
                                  const double* points_being_sampled, int p, int num_mc, const double* best_so_far,
                                  uint64_t seed, double* values_out, double* points_out, int* info);
int cmoe_ei_gradient_descent_mcmc(const cmoe_gp* const* gps, int num_gp, const cmoe_gd_params* outer,
                                  const double* domain_bounds, const double* starts, int num_starts, int q,
                                  const double* points_being_sampled, int p, int num_mc, const double* best_so_far,
                                  uint64_t seed, double* values_out, double* points_out, int* info);

/* posterior_mean_optimization (gpp_python_knowledge_gradient.cpp:306-350): ComputeOptimalPosteriorMean
 * (gpp_knowledge_gradient_optimization.cpp:420-472) from ONE start on the un-fantasised GP — line-search gradient
 * descent on -mu(x) over the dim-num_fidelity free coordinates (fidelity coordinates pinned to 1.0).
 * best_point[dim-num_fidelity]; best_value = -min mu found (the reference's best_function_value).
 * With max_num_restarts <= 0 nothing is written and *found_flag = 0, as in the reference (:424-426). */
int cmoe_posterior_mean_optimization(const cmoe_gp* gp, int num_fidelity, const cmoe_gd_params* params,
                                     const double* domain_bounds, const double* initial_guess, double* best_point,
                                     double* best_value, int* found_flag);

/* ---- building blocks exposed for parity tests and micro-benchmarks ------------------------------------------------ */
/* Covariance build alone on device-resident inputs: returns device time (usec) of `repeats` builds of the n*n matrix. */
int cmoe_bench_cov_build(const cmoe_gp* gp, int repeats, double* usec_per_build);
/* Blocked Cholesky alone (re-factors a fresh copy of K each repeat). */
int cmoe_bench_cholesky(const cmoe_gp* gp, int repeats, double* usec_per_factor);
/* Measured FP64 peaks of this GPU: tflops[0] = DFMA vector pipe, tflops[1] = DMMA tensor pipe (m8n8k4). */
int cmoe_bench_fp64_peaks(int device, double* tflops);
/* tflops[0] = total FP64 TFLOP/s when every warp interleaves DFMA and DMMA 1:1 by flops (do the vector pipe and the
 * tensor sub-pipe overlap?). */
int cmoe_bench_fp64_mixed(int device, double* tflops);
/* latency microbenchmarks behind the Cholesky pivot chain: out[11] (cycles), see microbench_chain.cu */
int cmoe_bench_chain_latencies(int device, double* out);
/* In-place lower Cholesky of a host matrix through the device path (ComputeCholeskyFactorL, gpp_linear_algebra.cpp:109). */
int cmoe_cholesky(int n, double* a, int device, int* info);
/* Solve (L L^T) X = B for nrhs right-hand sides (CholeskyFactorLMatrixMatrixSolve, gpp_linear_algebra.hpp:247). */
int cmoe_potrs(int n, int nrhs, const double* chol, double* x, int device);
/* Philox4x32-10 + Box-Muller normals exactly as the kernels draw them: out[num_draws][per_draw]. */
int cmoe_philox_normals(uint64_t seed, uint64_t first_draw, int num_draws, int per_draw, double* out, int device);

#ifdef __cplusplus
}
#endif
#endif /* CMOE_B200_H_ */
