// TEST INFRASTRUCTURE — not product code.
//
// Thin extern "C" driver around the UNMODIFIED Cornell-MOE C++ core, compiled where it lies under
// /root/reference by oracle/Makefile into oracle/_ref/libmoe_ref.so.  It exists so that
//   (1) the plain-C restatement in oracle/moe_oracle.c can be pinned against the real reference, and
//   (2) tests / bench.py --impl reference can time the reference's own OpenMP path on host cores.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// Every entry point here just marshals plain pointers into the reference's classes:
//   GaussianProcess                      moe/optimal_learning/cpp/gpp_math.hpp:275
//   ExpectedImprovementEvaluator/State   gpp_math.hpp:1001 / :1149
//   KnowledgeGradientEvaluator/State     gpp_knowledge_gradient_optimization.hpp:152 / :310
//   NormalRNGSimulator (table replay)    gpp_random.hpp:314
//   EvaluateKGAtPointList                gpp_knowledge_gradient_optimization.hpp:972
//   EvaluateEIAtPointList                gpp_math.cpp:2305
#include <omp.h>

#include <algorithm>
#include <cstring>
#include <memory>
#include <vector>

// K_chol_ has no public accessor in the reference; the parity tests need the factor itself.
#define private public
#include "gpp_math.hpp"
#undef private
#include "gpp_model_selection.hpp"
#include "gpp_covariance.hpp"
#include "gpp_domain.hpp"
#include "gpp_exception.hpp"
#include "gpp_expected_improvement_mcmc_optimization.hpp"
#include "gpp_knowledge_gradient_mcmc_optimization.hpp"
#include "gpp_knowledge_gradient_optimization.hpp"
#include "gpp_linear_algebra.hpp"
#include "gpp_optimization.hpp"
#include "gpp_optimizer_parameters.hpp"
#include "gpp_random.hpp"

using namespace optimal_learning;  // NOLINT

namespace {

std::unique_ptr<CovarianceInterface> MakeCovariance(int kernel, int dim, double alpha, const double* lengths) {
  if (kernel == 0) {
    return std::unique_ptr<CovarianceInterface>(new SquareExponential(dim, alpha, lengths));
  }
  return std::unique_ptr<CovarianceInterface>(new MaternNu2p5(dim, alpha, lengths));
}

TensorProductDomain MakeDomain(const double* bounds, int dim) {
  std::vector<ClosedInterval> iv(dim);
  for (int i = 0; i < dim; ++i) {
    iv[i] = ClosedInterval(bounds[2 * i], bounds[2 * i + 1]);
  }
  return TensorProductDomain(iv.data(), dim);
}

GradientDescentParameters MakeGD(const double* p) {
  // p = {num_multistarts, max_num_steps, max_num_restarts, num_steps_averaged, gamma, pre_mult, max_relative_change, tolerance}
  return GradientDescentParameters(static_cast<int>(p[0]), static_cast<int>(p[1]), static_cast<int>(p[2]),
                                   static_cast<int>(p[3]), p[4], p[5], p[6], p[7]);
}

}  // namespace

extern "C" {

// ---- covariance (gpp_covariance.cpp:121-234, 339-459) ----
void ref_covariance(int kernel, int dim, double alpha, const double* lengths, const double* p1, const int* d1, int g1,
                    const double* p2, const int* d2, int g2, double* cov) {
  auto c = MakeCovariance(kernel, dim, alpha, lengths);
  c->Covariance(p1, d1, g1, p2, d2, g2, cov);
}

void ref_grad_covariance(int kernel, int dim, double alpha, const double* lengths, const double* p1, const int* d1,
                         int g1, const double* p2, const int* d2, int g2, double* grad_cov) {
  auto c = MakeCovariance(kernel, dim, alpha, lengths);
  c->GradCovariance(p1, d1, g1, p2, d2, g2, grad_cov);
}

// ---- linear algebra (gpp_linear_algebra.cpp:109-208) ----
int ref_cholesky(int n, double* a) { return ComputeCholeskyFactorL(n, a); }
void ref_trsv(const double* a, int trans, int n, int lda, double* x) {
  TriangularMatrixVectorSolve(a, trans ? 'T' : 'N', n, lda, x);
}
void ref_potrs(const double* a, int n, int nrhs, double* x) { CholeskyFactorLMatrixMatrixSolve(a, n, nrhs, x); }

// ---- GP (gpp_math.cpp:553-573) ----
void* ref_gp_create(int kernel, double alpha, const double* lengths, const double* X, const double* y,
                    const double* noise, const int* derivs, int g, int dim, int N, int* leading_minor) {
  *leading_minor = 0;
  auto c = MakeCovariance(kernel, dim, alpha, lengths);
  try {
    return new GaussianProcess(*c, X, y, noise, derivs, g, dim, N);
  } catch (const SingularMatrixException& e) {
    *leading_minor = e.leading_minor_index();
    return nullptr;
  }
}

void ref_gp_destroy(void* h) { delete static_cast<GaussianProcess*>(h); }

void ref_gp_get_state(void* h, double* K_chol, double* K_inv_y, double* mean) {
  auto* gp = static_cast<GaussianProcess*>(h);
  if (K_chol) std::copy(gp->K_chol_.begin(), gp->K_chol_.end(), K_chol);
  if (K_inv_y) std::copy(gp->K_inv_y_.begin(), gp->K_inv_y_.end(), K_inv_y);
  if (mean) *mean = gp->mean_;
}

// Posterior queries at `num` points carrying the derivative rows `derivs_s[g_s]`
// (gpp_math.cpp:662,721,924,1366,1466; Python boundary gpp_python_gaussian_process.cpp:64-236).
// Any output pointer may be NULL.  Returns the leading-minor index if chol(var) fails, else 0.
int ref_gp_posterior(void* h, const double* pts, int num, const int* derivs_s, int g_s, double* mean, double* grad_mean,
                     double* var, double* chol_var, double* grad_var, double* grad_chol) {
  auto* gp = static_cast<GaussianProcess*>(h);
  const int Q = num * (1 + g_s);
  const bool need_grad = (grad_mean || grad_var || grad_chol);
  GaussianProcess::StateType st(*gp, pts, num, derivs_s, g_s, need_grad ? num : 0);
  if (mean) gp->ComputeMeanOfPoints(st, mean);
  if (grad_mean) gp->ComputeGradMeanOfPoints(st, grad_mean);
  std::vector<double> v(static_cast<size_t>(Q) * Q);
  gp->ComputeVarianceOfPoints(&st, derivs_s, g_s, v.data());
  if (var) std::copy(v.begin(), v.end(), var);
  if (grad_var) gp->ComputeGradVarianceOfPoints(&st, grad_var);
  if (chol_var || grad_chol) {
    int lm = ComputeCholeskyFactorL(Q, v.data());
    if (lm != 0) return lm;
    if (chol_var) std::copy(v.begin(), v.end(), chol_var);
    if (grad_chol) gp->ComputeGradCholeskyVarianceOfPoints(&st, v.data(), grad_chol);
  }
  return 0;
}

void ref_gp_mean_additional(void* h, const double* pts, int num, double* mean) {
  static_cast<GaussianProcess*>(h)->ComputeMeanOfAdditionalPoints(pts, num, nullptr, 0, mean);
}

// ---- q-EI (gpp_math.cpp:1991-2126) with table-fed normals ----
// table holds (q+p) normals per MC iteration, iteration-major.  grad may be NULL.
double ref_ei(void* h, const double* Xq, const double* Xp, int q, int p, int num_mc, double best_so_far,
              const double* table, int table_len, double* grad) {
  auto* gp = static_cast<GaussianProcess*>(h);
  std::vector<double> tab(table, table + table_len);
  NormalRNGSimulator rng(tab);
  ExpectedImprovementEvaluator ev(*gp, num_mc, best_so_far);
  ExpectedImprovementEvaluator::StateType st(ev, Xq, Xp, q, p, grad != nullptr, &rng);
  double v = ev.ComputeExpectedImprovement(&st);
  if (grad) ev.ComputeGradExpectedImprovement(&st, grad);
  return v;
}

// ---- q-KG / d-KG (gpp_knowledge_gradient_optimization.cpp:69-227) with table-fed normals ----
// table holds (q+p)(1+g) normals per EVEN iteration (odd iterations are antithetic).
// gd = 8 inner GradientDescentParameters; inner_bounds[2*(dim-num_fidelity)].
// best_points (may be NULL) receives the per-sample minimisers x*_i [num_mc][dim].
double ref_kg(void* h, int num_fidelity, const double* gd, const double* inner_bounds, const double* discrete_pts,
              int num_pts, const double* Xq, const double* Xp, int q, int p, int num_mc, double best_so_far,
              const double* table, int table_len, double* grad, double* best_points) {
  auto* gp = static_cast<GaussianProcess*>(h);
  std::vector<double> tab(table, table + table_len);
  NormalRNGSimulator rng(tab);
  TensorProductDomain dom = MakeDomain(inner_bounds, gp->dim() - num_fidelity);
  GradientDescentParameters inner = MakeGD(gd);
  KnowledgeGradientEvaluator<TensorProductDomain> ev(*gp, num_fidelity, discrete_pts, num_pts, num_mc, dom, inner,
                                                     best_so_far);
  std::vector<int> derivs(gp->derivatives());
  KnowledgeGradientEvaluator<TensorProductDomain>::StateType st(ev, Xq, Xp, q, p, num_mc, derivs.data(),
                                                                gp->num_derivatives(), grad != nullptr, &rng);
  double v;
  if (grad) {
    v = ev.ComputeGradKnowledgeGradient(&st, grad);
  } else {
    v = ev.ComputeKnowledgeGradient(&st);
  }
  if (best_points) std::copy(st.best_point.begin(), st.best_point.end(), best_points);
  return v;
}

// q-KG through a REUSED state, as the multistart drivers evaluate it: the state is constructed with `Xfirst` and then
// moved to `Xq` by SetCurrentPoint — which refreshes the GP quantities but not discretized_set
// (gpp_knowledge_gradient_optimization.cpp:233-243 vs :259-261), so the inner optimiser's start set keeps Xfirst.
double ref_kg_reused_state(void* h, int num_fidelity, const double* gd, const double* inner_bounds,
                           const double* discrete_pts, int num_pts, const double* Xfirst, const double* Xq,
                           const double* Xp, int q, int p, int num_mc, double best_so_far, const double* table,
                           int table_len, double* grad) {
  auto* gp = static_cast<GaussianProcess*>(h);
  std::vector<double> tab(table, table + table_len);
  NormalRNGSimulator rng(tab);
  TensorProductDomain dom = MakeDomain(inner_bounds, gp->dim() - num_fidelity);
  GradientDescentParameters inner = MakeGD(gd);
  KnowledgeGradientEvaluator<TensorProductDomain> ev(*gp, num_fidelity, discrete_pts, num_pts, num_mc, dom, inner,
                                                     best_so_far);
  std::vector<int> derivs(gp->derivatives());
  KnowledgeGradientEvaluator<TensorProductDomain>::StateType st(ev, Xfirst, Xp, q, p, num_mc, derivs.data(),
                                                                gp->num_derivatives(), true, &rng);
  st.SetCurrentPoint(ev, Xq);
  return grad ? ev.ComputeGradKnowledgeGradient(&st, grad) : ev.ComputeKnowledgeGradient(&st);
}

// ---- the reference's own parallel path, for CPU baselines ----
// EvaluateKGAtPointList (NullOptimizer + OpenMP static schedule), one NormalRNG per thread seeded seed+t.
void ref_evaluate_kg_at_point_list(void* h, int num_fidelity, const double* gd, const double* bounds,
                                   const double* inner_bounds, const double* discrete_pts, int num_pts,
                                   const double* candidates, const double* Xp, int num_candidates, int q, int p,
                                   int num_mc, double best_so_far, int num_threads, unsigned seed, double* values,
                                   double* best_point) {
  auto* gp = static_cast<GaussianProcess*>(h);
  TensorProductDomain dom = MakeDomain(bounds, gp->dim());
  TensorProductDomain inner_dom = MakeDomain(inner_bounds, gp->dim() - num_fidelity);
  GradientDescentParameters inner = MakeGD(gd);
  std::vector<NormalRNG> rngs(num_threads);
  for (int t = 0; t < num_threads; ++t) rngs[t].SetExplicitSeed(seed + t);
  ThreadSchedule sched(num_threads, omp_sched_static);
  bool found = false;
  std::vector<double> bp(static_cast<size_t>(q) * gp->dim());
  EvaluateKGAtPointList(*gp, num_fidelity, inner, dom, inner_dom, sched, candidates, Xp, discrete_pts,
                        num_candidates, q, p, num_pts, best_so_far, num_mc, &found, rngs.data(), values, bp.data());
  if (best_point) std::copy(bp.begin(), bp.end(), best_point);
}

// Value AND gradient for a list of candidates, OpenMP over candidates (the work one outer-GD step does).
void ref_kg_grad_at_point_list(void* h, int num_fidelity, const double* gd, const double* inner_bounds,
                               const double* discrete_pts, int num_pts, const double* candidates, const double* Xp,
                               int num_candidates, int q, int p, int num_mc, double best_so_far, int num_threads,
                               unsigned seed, double* values, double* grads) {
  auto* gp = static_cast<GaussianProcess*>(h);
  const int dim = gp->dim();
  TensorProductDomain inner_dom = MakeDomain(inner_bounds, dim - num_fidelity);
  GradientDescentParameters inner = MakeGD(gd);
  KnowledgeGradientEvaluator<TensorProductDomain> ev(*gp, num_fidelity, discrete_pts, num_pts, num_mc, inner_dom,
                                                     inner, best_so_far);
  std::vector<int> derivs(gp->derivatives());
#pragma omp parallel num_threads(num_threads)
  {
    NormalRNG rng(seed + omp_get_thread_num());
#pragma omp for schedule(static)
    for (int c = 0; c < num_candidates; ++c) {
      KnowledgeGradientEvaluator<TensorProductDomain>::StateType st(
          ev, candidates + static_cast<size_t>(c) * q * dim, Xp, q, p, num_mc, derivs.data(), gp->num_derivatives(),
          true, &rng);
      values[c] = ev.ComputeGradKnowledgeGradient(&st, grads + static_cast<size_t>(c) * q * dim);
    }
  }
}

void ref_evaluate_ei_at_point_list(void* h, const double* candidates, const double* Xp, int num_candidates, int q,
                                   int p, int num_mc, double best_so_far, int num_threads, unsigned seed,
                                   double* values) {
  auto* gp = static_cast<GaussianProcess*>(h);
  std::vector<NormalRNG> rngs(num_threads);
  for (int t = 0; t < num_threads; ++t) rngs[t].SetExplicitSeed(seed + t);
  ThreadSchedule sched(num_threads, omp_sched_static);
  bool found = false;
  std::vector<double> bp(static_cast<size_t>(q) * gp->dim());
  EvaluateEIAtPointList(*gp, sched, candidates, Xp, num_candidates, q, p, best_so_far, num_mc, &found, rngs.data(),
                        values, bp.data());
}

void ref_ei_grad_at_point_list(void* h, const double* candidates, const double* Xp, int num_candidates, int q, int p,
                               int num_mc, double best_so_far, int num_threads, unsigned seed, double* values,
                               double* grads) {
  auto* gp = static_cast<GaussianProcess*>(h);
  const int dim = gp->dim();
  ExpectedImprovementEvaluator ev(*gp, num_mc, best_so_far);
#pragma omp parallel num_threads(num_threads)
  {
    NormalRNG rng(seed + omp_get_thread_num());
#pragma omp for schedule(static)
    for (int c = 0; c < num_candidates; ++c) {
      ExpectedImprovementEvaluator::StateType st(ev, candidates + static_cast<size_t>(c) * q * dim, Xp, q, p, true,
                                                 &rng);
      values[c] = ev.ComputeExpectedImprovement(&st);
      ev.ComputeGradExpectedImprovement(&st, grads + static_cast<size_t>(c) * q * dim);
    }
  }
}

// ComputeOptimalPosteriorMean from one start on the un-fantasised GP (gpp_knowledge_gradient_optimization.cpp:420-472);
// the body of the Python boundary's posterior_mean_optimization (gpp_python_knowledge_gradient.cpp:306-350).
double ref_posterior_mean_optimization(void* h, int num_fidelity, const double* gd, const double* bounds,
                                       const double* initial_guess, double* best_point) {
  auto* gp = static_cast<GaussianProcess*>(h);
  TensorProductDomain dom = MakeDomain(bounds, gp->dim() - num_fidelity);
  GradientDescentParameters params = MakeGD(gd);
  bool found = false;
  double best_value = 0.0;
  ComputeOptimalPosteriorMean(*gp, num_fidelity, params, dom, initial_guess, 1, &found, best_point, &best_value);
  return best_value;
}

// ---- MCMC-averaged acquisition over an ensemble of Matern-5/2 GPs (one per hyper-parameter sample) ----
// GaussianProcessMCMC (gpp_knowledge_gradient_mcmc_optimization.cpp:24-48): hypers[num_mcmc][1+dim] = (alpha, lengths),
// noises[num_mcmc][1+g].  Table-fed normals, shared by every member (each evaluation rewinds the table).
double ref_kg_mcmc(const double* hypers, const double* noises, int num_mcmc, const double* X, const double* y,
                   const int* derivs, int g, int dim, int N, int num_fidelity, const double* gd,
                   const double* inner_bounds, const double* discrete_pts, int num_pts, const double* Xq,
                   const double* Xp, int q, int p, int num_mc, const double* best_so_far, const double* table,
                   int table_len, double* grad) {
  GaussianProcessMCMC gpm(hypers, noises, num_mcmc, X, y, derivs, g, dim, N);
  std::vector<double> tab(table, table + table_len);
  NormalRNGSimulator rng(tab);
  TensorProductDomain dom = MakeDomain(inner_bounds, dim - num_fidelity);
  GradientDescentParameters inner = MakeGD(gd);
  std::vector<KnowledgeGradientState<TensorProductDomain>::EvaluatorType> evs;
  KnowledgeGradientMCMCEvaluator<TensorProductDomain> ev(gpm, num_fidelity, discrete_pts, num_pts, num_mc, dom, inner,
                                                         best_so_far, &evs);
  std::vector<KnowledgeGradientEvaluator<TensorProductDomain>::StateType> states;
  KnowledgeGradientMCMCEvaluator<TensorProductDomain>::StateType st(ev, Xq, Xp, q, p, num_pts, derivs, g,
                                                                    grad != nullptr, &rng, &states);
  const double v = ev.ComputeKnowledgeGradient(&st);
  if (grad) {
    std::fill(grad, grad + static_cast<size_t>(q) * dim, 0.0);  // the evaluator accumulates with += (:166)
    ev.ComputeGradKnowledgeGradient(&st, grad);
  }
  return v;
}

double ref_ei_mcmc(const double* hypers, const double* noises, int num_mcmc, const double* X, const double* y,
                   const int* derivs, int g, int dim, int N, const double* Xq, const double* Xp, int q, int p,
                   int num_mc, const double* best_so_far, const double* table, int table_len, double* grad) {
  GaussianProcessMCMC gpm(hypers, noises, num_mcmc, X, y, derivs, g, dim, N);
  std::vector<double> tab(table, table + table_len);
  NormalRNGSimulator rng(tab);
  std::vector<ExpectedImprovementState::EvaluatorType> evs;
  ExpectedImprovementMCMCEvaluator ev(gpm, num_mc, best_so_far, &evs);
  std::vector<ExpectedImprovementEvaluator::StateType> states;
  ExpectedImprovementMCMCEvaluator::StateType st(ev, Xq, Xp, q, p, derivs, g, grad != nullptr, &rng, &states);
  const double v = ev.ComputeExpectedImprovement(&st);
  if (grad) {
    std::fill(grad, grad + static_cast<size_t>(q) * dim, 0.0);
    ev.ComputeGradExpectedImprovement(&st, grad);
  }
  return v;
}

// ---- log marginal likelihood (gpp_model_selection.cpp:540-612; Python boundary gpp_python_model_selection.cpp:43-87) ----
// noise[1+g]; the state adds 1e-6 to the diagonal on top of the noise and ignores a failed factorisation (:546-553).
double ref_log_marginal_likelihood(int kernel, double alpha, const double* lengths, const double* X, const double* y,
                                   const double* noise, const int* derivs, int g, int dim, int N) {
  auto c = MakeCovariance(kernel, dim, alpha, lengths);
  LogMarginalLikelihoodEvaluator ev(X, y, derivs, g, dim, N);
  std::vector<double> nz(noise, noise + 1 + g);
  LogMarginalLikelihoodState st(ev, *c, nz);
  return ev.ComputeLogLikelihood(st);
}

// gradient wrt (alpha, lengths[dim], noise[1+g]) (gpp_model_selection.cpp:629-677)
void ref_grad_log_marginal_likelihood(int kernel, double alpha, const double* lengths, const double* X, const double* y,
                                      const double* noise, const int* derivs, int g, int dim, int N, double* grad) {
  auto c = MakeCovariance(kernel, dim, alpha, lengths);
  LogMarginalLikelihoodEvaluator ev(X, y, derivs, g, dim, N);
  std::vector<double> nz(noise, noise + 1 + g);
  LogMarginalLikelihoodState st(ev, *c, nz);
  ev.ComputeGradLogLikelihood(&st, grad);
}

// LimitUpdate (gpp_domain.cpp:64-104) for pinning the restatement.
void ref_limit_update(const double* bounds, int dim, double max_relative_change, const double* current_point,
                      double* update) {
  TensorProductDomain dom = MakeDomain(bounds, dim);
  dom.LimitUpdate(max_relative_change, current_point, update);
}

// SimplexIntersectTensorProductDomain::LimitUpdate (gpp_domain.cpp:234-289)
void ref_limit_update_simplex(const double* bounds, int dim, double max_relative_change, const double* current_point,
                              double* update) {
  std::vector<ClosedInterval> iv(dim);
  for (int i = 0; i < dim; ++i) iv[i] = ClosedInterval(bounds[2 * i], bounds[2 * i + 1]);
  SimplexIntersectTensorProductDomain dom(iv.data(), dim);
  dom.LimitUpdate(max_relative_change, current_point, update);
}

// the q-EI multistart driver over the simplex-intersect-box domain
void ref_multistart_ei_simplex(void* h, const double* gd_outer, const double* bounds, const double* starts,
                               int num_starts, int q, const double* Xp, int p, int num_mc, double best_so_far,
                               unsigned seed, double* best_point) {
  auto* gp = static_cast<GaussianProcess*>(h);
  std::vector<ClosedInterval> iv(gp->dim());
  for (int i = 0; i < gp->dim(); ++i) iv[i] = ClosedInterval(bounds[2 * i], bounds[2 * i + 1]);
  SimplexIntersectTensorProductDomain dom(iv.data(), gp->dim());
  GradientDescentParameters outer = MakeGD(gd_outer);
  NormalRNG rng(seed);
  ThreadSchedule sched(1, omp_sched_static);
  bool found = false;
  ComputeOptimalPointsToSampleViaMultistartGradientDescent(*gp, outer, dom, sched, starts, Xp, num_starts, q, p,
                                                           best_so_far, num_mc, &rng, &found, best_point);
}

// ---- the multistart drivers themselves, UNMODIFIED, single-threaded --------------------------------------------------
// Every q-KG / q-EI evaluation rewinds its NormalRNG (ResetToMostRecentSeed, gpp_knowledge_gradient_optimization.cpp:81,
// 164; gpp_math.cpp:2011, 2076), so a whole driver call consumes the same first draws of NormalRNG(seed) over and over:
// ref_normal_draws() returns that table, and feeding it to the device path as `normals_table` makes both sides use
// identical normals without touching the reference.
void ref_normal_draws(unsigned seed, int count, double* out) {
  NormalRNG rng(seed);
  for (int i = 0; i < count; ++i) out[i] = rng();
}

// ComputeKGOptimalPointsToSampleViaMultistartGradientDescent (gpp_knowledge_gradient_optimization.hpp:859-935)
int ref_multistart_kg(void* h, int num_fidelity, const double* gd_outer, const double* gd_inner, const double* bounds,
                      const double* inner_bounds, const double* discrete_pts, int num_pts, const double* starts,
                      int num_starts, int q, const double* Xp, int p, int num_mc, double best_so_far, unsigned seed,
                      double* best_point) {
  auto* gp = static_cast<GaussianProcess*>(h);
  TensorProductDomain dom = MakeDomain(bounds, gp->dim());
  TensorProductDomain inner_dom = MakeDomain(inner_bounds, gp->dim() - num_fidelity);
  GradientDescentParameters outer = MakeGD(gd_outer), inner = MakeGD(gd_inner);
  NormalRNG rng(seed);
  ThreadSchedule sched(1, omp_sched_static);
  bool found = false;
  ComputeKGOptimalPointsToSampleViaMultistartGradientDescent(*gp, num_fidelity, outer, inner, dom, inner_dom, sched,
                                                             starts, Xp, discrete_pts, num_starts, q, p, num_pts,
                                                             best_so_far, num_mc, &rng, &found, best_point);
  return found ? 1 : 0;
}

// ComputeOptimalPointsToSampleViaMultistartGradientDescent (gpp_math.hpp:1683-1802); the reference never sets its
// found_flag there, so only the point is returned
void ref_multistart_ei(void* h, const double* gd_outer, const double* bounds, const double* starts, int num_starts,
                       int q, const double* Xp, int p, int num_mc, double best_so_far, unsigned seed,
                       double* best_point) {
  auto* gp = static_cast<GaussianProcess*>(h);
  TensorProductDomain dom = MakeDomain(bounds, gp->dim());
  GradientDescentParameters outer = MakeGD(gd_outer);
  NormalRNG rng(seed);
  ThreadSchedule sched(1, omp_sched_static);
  bool found = false;
  ComputeOptimalPointsToSampleViaMultistartGradientDescent(*gp, outer, dom, sched, starts, Xp, num_starts, q, p,
                                                           best_so_far, num_mc, &rng, &found, best_point);
}

int ref_max_threads() { return omp_get_max_threads(); }

}  // extern "C"
