// Multistart optimisation drivers: screen every start, keep the best 20, run restarted gradient descent on those
// (all starts advance together, one batched device evaluation per step), strict-> arg-max.
//
// Replaces (reference, moe/optimal_learning/cpp/):
//   ComputeKGOptimalPointsToSampleViaMultistartGradientDescent   gpp_knowledge_gradient_optimization.hpp:859-935
//   ComputeOptimalPointsToSampleViaMultistartGradientDescent     gpp_math.hpp:1683-1802
//   GradientDescentOptimization / GradientDescentOptimizer::Optimize   gpp_optimization.hpp:620-705, 1144-1185
//   MultistartOptimizer::MultistartOptimize (arg-max part)        gpp_optimization.hpp:1452-1564
//   RepeatedDomain::LimitUpdate / TensorProductDomain::LimitUpdate gpp_domain.hpp:536-540, gpp_domain.cpp:64-104
//   OnePotentialSampleExpectedImprovementEvaluator                gpp_math.cpp:2196-2253
// The step bookkeeping is O(q*dim) per start and stays on the host; every objective / gradient evaluation is a
// batched device call (kg.cu / ei.cu / posterior.cu).
#include <algorithm>
#include <cmath>
#include <functional>
#include <limits>
#include <queue>

#include "internal.cuh"

using namespace cmoe;  // NOLINT

namespace {

constexpr int kTopK = 20;  // hard-coded in the reference (gpp_knowledge_gradient_optimization.hpp:901, gpp_math.hpp:1765)

// TensorProductDomain::LimitUpdate for one coordinate (gpp_domain.cpp:64-104)
double limit_step_host(double step, double x, double lo, double hi, double mrc) {
  double dist = std::fmin(x - lo, hi - x);
  if (std::fabs(step) > mrc * dist) step = std::copysign(mrc * dist, step);
  const double next = x + step;
  if (next < lo || next > hi) {
    if (next < lo) {
      dist = lo - x;
      step = (x + step * 0.5 < lo) ? dist * 0.5 : step * 0.5;
    } else {
      dist = hi - x;
      step = (x + step * 0.5 > hi) ? dist * 0.5 : step * 0.5;
    }
  }
  return step;
}

using BatchEval = std::function<void(const double* pts, int nc, double* values, double* grads)>;

// Restarted gradient descent on `ns` starts at once.  eval(pts, nc, values, grads): grads may be NULL.
void gradient_descent_batch(const BatchEval& eval, const cmoe_gd_params& gd, const double* domain_bounds, int q,
                            int dim, const double* starts, int ns, double* values_out, double* points_out) {
  const int ps = q * dim;
  std::vector<double> x(starts, starts + static_cast<size_t>(ns) * ps);
  if (gd.max_num_restarts > 0) {
    std::vector<char> converged(ns, 0);
    std::vector<double> x0(x.size()), pts, vals, grads;
    const double step_tol = gd.tolerance / static_cast<double>(gd.max_num_steps);
    for (int r = 0; r < gd.max_num_restarts; ++r) {
      std::vector<int> act;
      for (int s = 0; s < ns; ++s)
        if (!converged[s]) act.push_back(s);
      if (act.empty()) break;
      x0 = x;
      std::vector<char> run_done(ns, 0);
      for (int i = 0; i < gd.max_num_steps; ++i) {
        std::vector<int> cur;
        for (int s : act)
          if (!run_done[s]) cur.push_back(s);
        if (cur.empty()) break;
        const int nc = static_cast<int>(cur.size());
        pts.resize(static_cast<size_t>(nc) * ps);
        vals.resize(nc);
        grads.resize(static_cast<size_t>(nc) * ps);
        for (int k = 0; k < nc; ++k)
          std::copy(x.begin() + static_cast<size_t>(cur[k]) * ps, x.begin() + static_cast<size_t>(cur[k] + 1) * ps,
                    pts.begin() + static_cast<size_t>(k) * ps);
        eval(pts.data(), nc, vals.data(), grads.data());
        const double alpha = gd.pre_mult * std::pow(static_cast<double>(i + 1), -gd.gamma);
        for (int k = 0; k < nc; ++k) {
          double* xs = x.data() + static_cast<size_t>(cur[k]) * ps;
          const double* g = grads.data() + static_cast<size_t>(k) * ps;
          double ns2 = 0.0;
          for (int j = 0; j < ps; ++j) {
            const int d = j % dim;
            const double step = limit_step_host(alpha * g[j], xs[j], domain_bounds[2 * d], domain_bounds[2 * d + 1],
                                                gd.max_relative_change);
            xs[j] += step;
            ns2 += step * step;
          }
          if (std::sqrt(ns2) < step_tol) run_done[cur[k]] = 1;
        }
      }
      for (int s : act) {
        double nd = 0.0;
        for (int j = 0; j < ps; ++j) {
          const double df = x0[static_cast<size_t>(s) * ps + j] - x[static_cast<size_t>(s) * ps + j];
          nd += df * df;
        }
        if (std::sqrt(nd) <= gd.tolerance) converged[s] = 1;
      }
    }
  }
  eval(x.data(), ns, values_out, nullptr);
  std::copy(x.begin(), x.end(), points_out);
}

// indices of the (up to) 20 largest values, in the order the reference feeds them to the optimiser
std::vector<int> top_k_indices(const double* values, int n) {
  std::priority_queue<std::pair<double, int>> pq;
  const int k = std::min(kTopK, n);
  for (int i = 0; i < n; ++i) {
    if (i < k) {
      pq.push({-values[i], i});
    } else if (pq.top().first > -values[i]) {
      pq.pop();
      pq.push({-values[i], i});
    }
  }
  std::vector<int> out;
  while (!pq.empty()) {
    out.push_back(pq.top().second);
    pq.pop();
  }
  return out;
}

void multistart_common(const BatchEval& eval, const cmoe_gd_params& outer, const double* domain_bounds, int q, int dim,
                       const double* starts, int num_starts, double init_best, double* start_values,
                       double* best_point, double* best_value, int* found_flag) {
  const int ps = q * dim;
  std::vector<double> vals(num_starts);
  eval(starts, num_starts, vals.data(), nullptr);
  if (start_values) std::copy(vals.begin(), vals.end(), start_values);
  const std::vector<int> top = top_k_indices(vals.data(), num_starts);
  const int k = static_cast<int>(top.size());
  std::vector<double> tk(static_cast<size_t>(k) * ps), fin_v(k), fin_p(static_cast<size_t>(k) * ps);
  for (int i = 0; i < k; ++i)
    std::copy(starts + static_cast<size_t>(top[i]) * ps, starts + static_cast<size_t>(top[i] + 1) * ps,
              tk.begin() + static_cast<size_t>(i) * ps);
  gradient_descent_batch(eval, outer, domain_bounds, q, dim, tk.data(), k, fin_v.data(), fin_p.data());
  // OptimizationIOContainer: best point initialised to the first start, strict `<` update
  double best = init_best;
  int found = 0;
  std::copy(tk.begin(), tk.begin() + ps, best_point);
  for (int i = 0; i < k; ++i) {
    if (best < fin_v[i]) {
      best = fin_v[i];
      found = 1;
      std::copy(fin_p.begin() + static_cast<size_t>(i) * ps, fin_p.begin() + static_cast<size_t>(i + 1) * ps,
                best_point);
    }
  }
  if (best_value) *best_value = best;
  if (found_flag) *found_flag = found;
}

struct KgEvaluator {
  cmoe_kg_plan* vplan = nullptr;
  cmoe_kg_plan* gplan = nullptr;
  ~KgEvaluator() {
    cmoe_kg_plan_destroy(vplan);
    cmoe_kg_plan_destroy(gplan);
  }
  static void check(int rc, int info = 0) {
    if (rc != CMOE_OK) throw Error(rc, cmoe_last_error(), info);
  }
  void operator()(const double* pts, int nc, double* values, double* grads) {
    cmoe_kg_plan* pl = grads ? gplan : vplan;
    int info = 0;
    check(cmoe_kg_plan_upload(pl, pts, nc));
    check(cmoe_kg_plan_run(pl));
    check(cmoe_kg_plan_sync(pl, &info), info);
    check(cmoe_kg_plan_download(pl, values, grads, nullptr));
  }
};

void make_kg_evaluator(KgEvaluator& ev, const cmoe_gp* gp, int nf, const cmoe_gd_params* inner,
                       const double* inner_bounds, const double* discrete_pts, int num_pts, int max_value_cands,
                       int max_grad_cands, int q, const double* Xp, int p, int num_mc, double best_so_far,
                       uint64_t seed) {
  KgEvaluator::check(cmoe_kg_plan_create(gp, nf, inner, inner_bounds, discrete_pts, num_pts, max_value_cands, q, Xp, p,
                                         num_mc, best_so_far, seed, 0, &ev.vplan));
  KgEvaluator::check(cmoe_kg_plan_create(gp, nf, inner, inner_bounds, discrete_pts, num_pts, max_grad_cands, q, Xp, p,
                                         num_mc, best_so_far, seed, 1, &ev.gplan));
}

// analytic one-point EI and gradient (gpp_math.cpp:2196-2253) from device posterior quantities
void analytic_ei(const cmoe_gp* gp, const double* pts, int nc, double best_so_far, double* values, double* grads) {
  const int dim = gp->spec.dim;
  std::vector<double> mu(nc), var(nc), gmu, gvar;
  int info = 0;
  if (grads) {
    gmu.resize(static_cast<size_t>(nc) * dim);
    gvar.resize(static_cast<size_t>(nc) * dim);
  }
  const int rc = cmoe_gp_posterior(gp, pts, nc, 1, nullptr, 0, mu.data(), grads ? gmu.data() : nullptr, var.data(),
                                   nullptr, grads ? gvar.data() : nullptr, nullptr, &info);
  if (rc != CMOE_OK) throw Error(rc, cmoe_last_error(), info);
  const double kMinVarEI = std::numeric_limits<double>::min();
  const double eps = std::numeric_limits<double>::epsilon();
  const double kMinVarGradEI = 150.0 * eps * eps;
  const double inv_sqrt_2pi = 0.39894228040143267793994605993438;
  auto pdf = [&](double z) { return std::exp(-0.5 * z * z) * inv_sqrt_2pi; };
  auto cdf = [&](double z) { return 0.5 * std::erfc(-z * 0.70710678118654752440084436210485); };
  for (int c = 0; c < nc; ++c) {
    {
      const double sigma = std::sqrt(std::fmax(kMinVarEI, var[c]));
      const double t = best_so_far - mu[c];
      const double ei = t * cdf(t / sigma) + sigma * pdf(t / sigma);
      values[c] = std::fmax(0.0, ei);
    }
    if (grads) {
      const double v = std::fmax(kMinVarGradEI, var[c]);
      const double sigma = std::sqrt(v);
      const double mu_diff = best_so_far - mu[c];
      const double C = mu_diff / sigma;
      const double pdf_C = pdf(C), cdf_C = cdf(C);
      for (int d = 0; d < dim; ++d) {
        const double gm = gmu[static_cast<size_t>(c) * dim + d];
        const double gchol = 0.5 * gvar[static_cast<size_t>(c) * dim + d] / sigma;
        const double d_C = (-sigma * gm - gchol * mu_diff) / v;
        const double d_A = -gm * cdf_C + mu_diff * pdf_C * d_C;
        const double d_B = gchol * pdf_C + sigma * (-C) * pdf_C * d_C;
        grads[static_cast<size_t>(c) * dim + d] = d_A + d_B;
      }
    }
  }
}

BatchEval make_ei_eval(const cmoe_gp* gp, int q, const double* Xp, int p, int num_mc, double best_so_far,
                       uint64_t seed) {
  if (q == 1 && p == 0) {
    // special analytic case (gpp_math.hpp:1703-1749)
    return [=](const double* pts, int nc, double* values, double* grads) {
      analytic_ei(gp, pts, nc, best_so_far, values, grads);
    };
  }
  return [=](const double* pts, int nc, double* values, double* grads) {
    ei_eval_batch(*gp, pts, nc, q, Xp, p, num_mc, best_so_far, seed, nullptr, values, grads);
  };
}

void validate_bounds(const double* b, int dim) {
  for (int d = 0; d < dim; ++d)
    CMOE_REQUIRE(b[2 * d] <= b[2 * d + 1], CMOE_ERR_BOUNDS, "Tensor product region is EMPTY.");
}

}  // namespace

extern "C" {

int cmoe_multistart_kg(const cmoe_gp* gp, int num_fidelity, const cmoe_gd_params* outer, const cmoe_gd_params* inner,
                       const double* domain_bounds, const double* inner_bounds, const double* discrete_pts, int num_pts,
                       const double* starts, int num_starts, int q, const double* points_being_sampled, int p,
                       int num_mc, double best_so_far, uint64_t seed, double* start_values, double* best_point,
                       double* best_value, int* found_flag, int* info) {
  return guarded(info, [&] {
    CMOE_REQUIRE(num_starts >= 1, CMOE_ERR_BOUNDS, "num_multistarts must be > 1");
    require_device(gp->device);
    validate_bounds(domain_bounds, gp->spec.dim);
    KgEvaluator ev;
    make_kg_evaluator(ev, gp, num_fidelity, inner, inner_bounds, discrete_pts, num_pts, num_starts,
                      std::min(kTopK, num_starts), q, points_being_sampled, p, num_mc, best_so_far, seed);
    BatchEval f = std::ref(ev);
    multistart_common(f, *outer, domain_bounds, q, gp->spec.dim, starts, num_starts,
                      -std::numeric_limits<double>::infinity(), start_values, best_point, best_value, found_flag);
  });
}

int cmoe_multistart_ei(const cmoe_gp* gp, const cmoe_gd_params* outer, const double* domain_bounds,
                       const double* starts, int num_starts, int q, const double* points_being_sampled, int p,
                       int num_mc, double best_so_far, uint64_t seed, double* start_values, double* best_point,
                       double* best_value, int* found_flag, int* info) {
  return guarded(info, [&] {
    CMOE_REQUIRE(num_starts >= 1, CMOE_ERR_BOUNDS, "num_multistarts must be > 1");
    require_device(gp->device);
    validate_bounds(domain_bounds, gp->spec.dim);
    BatchEval f = make_ei_eval(gp, q, points_being_sampled, p, num_mc, best_so_far, seed);
    multistart_common(f, *outer, domain_bounds, q, gp->spec.dim, starts, num_starts, -1.0, start_values, best_point,
                      best_value, found_flag);
  });
}

int cmoe_kg_gradient_descent(const cmoe_gp* gp, int num_fidelity, const cmoe_gd_params* outer,
                             const cmoe_gd_params* inner, const double* domain_bounds, const double* inner_bounds,
                             const double* discrete_pts, int num_pts, const double* starts, int num_starts, int q,
                             const double* points_being_sampled, int p, int num_mc, double best_so_far, uint64_t seed,
                             double* values_out, double* points_out, int* info) {
  return guarded(info, [&] {
    CMOE_REQUIRE(num_starts >= 1, CMOE_ERR_BOUNDS, "num_multistarts must be > 1");
    require_device(gp->device);
    validate_bounds(domain_bounds, gp->spec.dim);
    KgEvaluator ev;
    make_kg_evaluator(ev, gp, num_fidelity, inner, inner_bounds, discrete_pts, num_pts, num_starts, num_starts, q,
                      points_being_sampled, p, num_mc, best_so_far, seed);
    BatchEval f = std::ref(ev);
    gradient_descent_batch(f, *outer, domain_bounds, q, gp->spec.dim, starts, num_starts, values_out, points_out);
  });
}

int cmoe_ei_gradient_descent(const cmoe_gp* gp, const cmoe_gd_params* outer, const double* domain_bounds,
                             const double* starts, int num_starts, int q, const double* points_being_sampled, int p,
                             int num_mc, double best_so_far, uint64_t seed, double* values_out, double* points_out,
                             int* info) {
  return guarded(info, [&] {
    CMOE_REQUIRE(num_starts >= 1, CMOE_ERR_BOUNDS, "num_multistarts must be > 1");
    require_device(gp->device);
    validate_bounds(domain_bounds, gp->spec.dim);
    BatchEval f = make_ei_eval(gp, q, points_being_sampled, p, num_mc, best_so_far, seed);
    gradient_descent_batch(f, *outer, domain_bounds, q, gp->spec.dim, starts, num_starts, values_out, points_out);
  });
}

}  // extern "C"
