/* TEST INFRASTRUCTURE — the oracle is a checker, never the product.
 *
 * Plain-C restatement of the Cornell-MOE GP-posterior + Monte-Carlo acquisition hot path
 * (reference files under moe/optimal_learning/cpp/, cited per function in moe_oracle.c).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  Parity status: PINNED — tests/test_oracle_vs_reference.py checks every entry point
 * against the unmodified reference compiled into oracle/_ref/libmoe_ref.so, and
 * tests/golden/ holds reference-generated vectors that are re-checked where the reference is absent.
 *
 * Conventions (identical to the reference): matrices column-major; point sets [num][dim];
 * observation vector interleaved (value, derivs...) per point; derivative-index lists select
 * which partial derivatives are observed; everything is IEEE double.
 */
#ifndef MOE_ORACLE_H_
#define MOE_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_gp oracle_gp;

/* kernel: 0 = SquareExponential, 1 = MaternNu2p5 */
void oracle_covariance(int kernel, int dim, double alpha, const double* lengths, const double* p1, const int* d1,
                       int g1, const double* p2, const int* d2, int g2, double* cov);
void oracle_grad_covariance(int kernel, int dim, double alpha, const double* lengths, const double* p1, const int* d1,
                            int g1, const double* p2, const int* d2, int g2, double* grad_cov);

int oracle_cholesky(int n, double* a);
void oracle_trsv(const double* a, int trans, int n, int lda, double* x);
void oracle_potrs(const double* a, int n, int nrhs, double* x);

oracle_gp* oracle_gp_create(int kernel, double alpha, const double* lengths, const double* X, const double* y,
                            const double* noise, const int* derivs, int g, int dim, int N, int* leading_minor);
void oracle_gp_destroy(oracle_gp* gp);
void oracle_gp_get_state(const oracle_gp* gp, double* K_chol, double* K_inv_y, double* mean);
int oracle_gp_posterior(const oracle_gp* gp, const double* pts, int num, const int* derivs_s, int g_s, double* mean,
                        double* grad_mean, double* var, double* chol_var, double* grad_var, double* grad_chol);
void oracle_gp_mean_additional(const oracle_gp* gp, const double* pts, int num, double* mean);

double oracle_ei(const oracle_gp* gp, const double* Xq, const double* Xp, int q, int p, int num_mc,
                 double best_so_far, const double* table, int table_len, double* grad);
double oracle_kg(const oracle_gp* gp, int num_fidelity, const double* gd, const double* inner_bounds,
                 const double* discrete_pts, int num_pts, const double* Xq, const double* Xp, int q, int p,
                 int num_mc, double best_so_far, const double* table, int table_len, double* grad,
                 double* best_points);

/* MCMC-averaged q-KG / q-EI over an ensemble (discrete_pts[num_gp][num_pts][dim-nf], best_so_far[num_gp]) */
double oracle_kg_mcmc(const oracle_gp* const* gps, int num_gp, int num_fidelity, const double* gd,
                      const double* inner_bounds, const double* discrete_pts, int num_pts, const double* Xq,
                      const double* Xp, int q, int p, int num_mc, const double* best_so_far, const double* table,
                      int table_len, double* grad);
double oracle_ei_mcmc(const oracle_gp* const* gps, int num_gp, const double* Xq, const double* Xp, int q, int p,
                      int num_mc, const double* best_so_far, const double* table, int table_len, double* grad);

double oracle_posterior_mean_optimization(const oracle_gp* gp, int num_fidelity, const double* gd,
                                          const double* bounds, const double* initial_guess, double* best_point);

/* log p(y | X, theta) of gpp_model_selection.cpp:540-612 (1e-6 jitter on top of the noise, centred y) */
double oracle_log_marginal_likelihood(int kernel, double alpha, const double* lengths, const double* X, const double* y,
                                      const double* noise, const int* derivs, int g, int dim, int N);

/* d log p / d (alpha, l_1..l_dim, noise_0..noise_g), gpp_model_selection.cpp:629-677 */
void oracle_grad_log_marginal_likelihood(int kernel, double alpha, const double* lengths, const double* X,
                                         const double* y, const double* noise, const int* derivs, int g, int dim, int N,
                                         double* grad);

void oracle_limit_update(const double* bounds, int dim, double max_relative_change, const double* current_point,
                         double* update);

/* Philox4x32-10 + Box-Muller, the stream the CUDA path draws from.  Writes `per_draw` normals for each of
 * `num_draws` consecutive draw indices starting at first_draw: out[(i*per_draw)+k]. */
void oracle_philox_normals(uint64_t seed, uint64_t first_draw, int num_draws, int per_draw, double* out);

/* Candidate-list evaluators used as the "port" CPU baseline: OpenMP over candidates, Philox normals. */
void oracle_kg_at_point_list(const oracle_gp* gp, int num_fidelity, const double* gd, const double* inner_bounds,
                             const double* discrete_pts, int num_pts, const double* candidates, const double* Xp,
                             int num_candidates, int q, int p, int num_mc, double best_so_far, int num_threads,
                             uint64_t seed, double* values, double* grads);
void oracle_ei_at_point_list(const oracle_gp* gp, const double* candidates, const double* Xp, int num_candidates,
                             int q, int p, int num_mc, double best_so_far, int num_threads, uint64_t seed,
                             double* values, double* grads);
int oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif  /* MOE_ORACLE_H_ */
