// `GPP`: drop-in for the reference's Boost.Python module `moe.build.GPP` (gpp_python.cpp:453-600) for the GP-posterior
// + Monte-Carlo acquisition hot path, built with pybind11 (Boost.Python is not available) on top of the C ABI in
// include/cmoe_b200.h.  Same names, positional signatures, flat-list conventions, status-dict keys and exception
// classes as the reference, so `moe.optimal_learning.python.cpp_wrappers.*` can `import ... GPP as C_GP` unchanged:
//
//   GaussianProcess(hyperparameters, points_sampled, points_sampled_value, noise_variance, derivatives,
//                   num_derivatives, dim, num_sampled)                     gpp_python_gaussian_process.cpp:42-62, 295
//     .compute_mean_of_points / compute_mean_of_additional_points / compute_grad_mean_of_points /
//     .compute_variance_of_points / compute_cholesky_variance_of_points / compute_grad_variance_of_points /
//     .compute_grad_cholesky_variance_of_points / add_sampled_points / ...   :64-253, 456-463
//   compute_expected_improvement, compute_grad_expected_improvement, multistart_expected_improvement_optimization,
//   evaluate_EI_at_point_list                                             gpp_python_expected_improvement.cpp:44-441
//   compute_posterior_mean, compute_grad_posterior_mean, compute_knowledge_gradient, compute_grad_knowledge_gradient,
//   multistart_knowledge_gradient_optimization, posterior_mean_optimization, evaluate_KG_at_point_list
//                                                                         gpp_python_knowledge_gradient.cpp:44-397
//   GaussianProcessMCMC, compute_{,grad_}knowledge_gradient_mcmc, multistart_knowledge_gradient_mcmc_optimization,
//   evaluate_KG_mcmc_at_point_list                                        gpp_python_knowledge_gradient_mcmc.cpp:49-384
//   compute_{,grad_}expected_improvement_mcmc, multistart_expected_improvement_mcmc_optimization,
//   evaluate_EI_mcmc_at_point_list                                        gpp_python_expected_improvement_mcmc.cpp:46-300
//   GradientDescentParameters, NewtonParameters, RandomnessSourceContainer, OptimizerTypes, DomainTypes,
//   LogLikelihoodTypes                                                    gpp_python_common.cpp:201-370
//   OptimalLearningException, BoundsException, InvalidValueException, SingularMatrixException   gpp_python.cpp:189-206
//
// Differences that are inherent to the device path and documented in INTEGRATION.md:
//   * normals come from Philox4x32-10 keyed by the RandomnessSourceContainer's seed (not boost mt19937);
//   * `max_num_threads` is accepted and validated against the container but the parallel axis is the GPU;
//   * `use_gpu` / `which_gpu` (dead arguments in the reference) select the device.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "cmoe_b200.h"

namespace py = pybind11;

namespace {

PyObject* g_exc_base = nullptr;
PyObject* g_exc_bounds = nullptr;
PyObject* g_exc_invalid = nullptr;
PyObject* g_exc_singular = nullptr;

[[noreturn]] void raise_status(int rc, int info) {
  PyObject* type = g_exc_base;
  if (rc == CMOE_ERR_SINGULAR) type = g_exc_singular;
  if (rc == CMOE_ERR_BOUNDS) type = g_exc_bounds;
  if (rc == CMOE_ERR_INVALID_VALUE) type = g_exc_invalid;
  std::string msg = cmoe_last_error();
  if (rc == CMOE_ERR_SINGULAR) msg += " (leading minor index " + std::to_string(info) + ")";
  PyErr_SetString(type, msg.c_str());
  throw py::error_already_set();
}

void check(int rc, int info = 0) {
  if (rc != CMOE_OK) raise_status(rc, info);
}

std::vector<double> to_vec(const py::list& l, size_t n) {
  if (static_cast<size_t>(py::len(l)) < n) {
    PyErr_SetString(g_exc_bounds, "input list shorter than the size implied by the dimension arguments");
    throw py::error_already_set();
  }
  std::vector<double> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = l[i].cast<double>();
  return v;
}

std::vector<int> to_ivec(const py::list& l, size_t n) {
  std::vector<int> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = l[i].cast<int>();
  return v;
}

py::list to_list(const std::vector<double>& v) {
  py::list out;
  for (double x : v) out.append(x);
  return out;
}

// ---- parameter structs (gpp_optimizer_parameters.hpp:46-200) -------------------------------------------------------
struct GradientDescentParameters {
  cmoe_gd_params p;
  GradientDescentParameters(int num_multistarts, int max_num_steps, int max_num_restarts, int num_steps_averaged,
                            double gamma, double pre_mult, double max_relative_change, double tolerance)
      : p{num_multistarts, max_num_steps, max_num_restarts, num_steps_averaged, gamma, pre_mult, max_relative_change,
          tolerance} {}
};

struct NewtonParameters {
  int num_multistarts, max_num_steps;
  double gamma, time_factor, max_relative_change, tolerance;
  NewtonParameters(int a, int b, double c, double d, double e, double f)
      : num_multistarts(a), max_num_steps(b), gamma(c), time_factor(d), max_relative_change(e), tolerance(f) {}
};

enum class OptimizerTypes { kNull = 0, kGradientDescent = 1, kNewton = 2 };
enum class DomainTypes { kTensorProduct = 0, kSimplex = 1 };
enum class LogLikelihoodTypes { kLogMarginalLikelihood = 0, kLeaveOneOutLogLikelihood = 1 };

// ---- RandomnessSourceContainer (gpp_python_common.hpp:146-230) ------------------------------------------------------
struct RandomnessSourceContainer {
  static constexpr uint32_t kUniformDefaultSeed = 314, kNormalDefaultSeed = 314;
  std::mt19937 uniform_engine;
  uint32_t uniform_seed = kUniformDefaultSeed;
  std::vector<uint64_t> normal_seeds;  // one per "thread", as in the reference
  explicit RandomnessSourceContainer(int num_threads)
      : uniform_engine(kUniformDefaultSeed), normal_seeds(std::max(0, num_threads)) {
    SetExplicitNormalRNGSeed(kNormalDefaultSeed);
  }
  void SetExplicitUniformGeneratorSeed(uint32_t seed) { uniform_seed = seed; uniform_engine.seed(seed); }
  void SetRandomizedUniformGeneratorSeed(uint32_t seed) { SetExplicitUniformGeneratorSeed(mix(seed, 0)); }
  void ResetUniformGeneratorState() { uniform_engine.seed(uniform_seed); }
  void SetExplicitNormalRNGSeed(uint32_t seed) {
    for (size_t i = 0; i < normal_seeds.size(); ++i) normal_seeds[i] = static_cast<uint64_t>(seed) + i;
  }
  void SetRandomizedNormalRNGSeed(uint32_t seed) {
    for (size_t i = 0; i < normal_seeds.size(); ++i) normal_seeds[i] = mix(seed, static_cast<int>(i));
  }
  bool SetNormalRNGSeedPythonList(const py::list& seeds, const py::list& flags) {
    if (py::len(seeds) != py::len(flags) || py::len(seeds) != normal_seeds.size()) return false;
    for (size_t i = 0; i < normal_seeds.size(); ++i)
      if (flags[i].cast<int>()) normal_seeds[i] = static_cast<uint64_t>(seeds[i].cast<int64_t>());
    return true;
  }
  void ResetNormalRNGState() {}  // the Philox stream is a pure function of (seed, sample index): nothing to rewind
  void PrintState() const {
    std::printf("Uniform seed: %u\n", uniform_seed);
    for (size_t i = 0; i < normal_seeds.size(); ++i)
      std::printf("NormalRNG %zu: Philox4x32-10 key %llu\n", i, static_cast<unsigned long long>(normal_seeds[i]));
  }
  uint64_t seed0() const { return normal_seeds.empty() ? kNormalDefaultSeed : normal_seeds[0]; }
  static uint32_t mix(uint32_t seed, int thread_id) {
    std::random_device rd;
    return seed ^ (rd() + 0x9e3779b9u + (seed << 6) + (seed >> 2)) ^ static_cast<uint32_t>(thread_id * 2654435761u);
  }
};

void require_threads(int max_num_threads, const RandomnessSourceContainer& r) {
  if (max_num_threads > static_cast<int>(r.normal_seeds.size())) {
    PyErr_SetString(g_exc_bounds, "Fewer randomness_sources than max_num_threads.");
    throw py::error_already_set();
  }
}

// ---- GaussianProcess ------------------------------------------------------------------------------------------------
struct GaussianProcess {
  cmoe_gp* h = nullptr;
  int dim_ = 0, num_derivatives_ = 0;
  std::vector<int> derivatives_;
  GaussianProcess(const py::list& hyperparameters, const py::list& points_sampled, const py::list& points_sampled_value,
                  const py::list& noise_variance, const py::list& derivatives, int num_derivatives, int dim,
                  int num_sampled, const std::string& kernel, int which_gpu) {
    const double alpha = hyperparameters[0].cast<double>();
    const std::vector<double> lengths = to_vec(hyperparameters[1].cast<py::list>(), dim);
    const std::vector<double> X = to_vec(points_sampled, static_cast<size_t>(dim) * num_sampled);
    const std::vector<double> y = to_vec(points_sampled_value, static_cast<size_t>(num_sampled) * (1 + num_derivatives));
    const std::vector<double> noise = to_vec(noise_variance, 1 + num_derivatives);
    derivatives_ = to_ivec(derivatives, num_derivatives);
    dim_ = dim;
    num_derivatives_ = num_derivatives;
    // The reference's Python boundary hard-wires MaternNu2p5 (gpp_python_gaussian_process.cpp:53); SE is opt-in.
    const int kid = (kernel == "square_exponential") ? CMOE_KERNEL_SQUARE_EXPONENTIAL : CMOE_KERNEL_MATERN_NU_2P5;
    int info = 0;
    check(cmoe_gp_create(kid, alpha, lengths.data(), X.data(), y.data(), noise.data(), derivatives_.data(),
                         num_derivatives, dim, num_sampled, which_gpu, &h, &info),
          info);
  }
  ~GaussianProcess() { cmoe_gp_destroy(h); }
  GaussianProcess(const GaussianProcess&) = delete;
  GaussianProcess& operator=(const GaussianProcess&) = delete;
  int dim() const { return dim_; }
  int num_sampled() const { return cmoe_gp_num_sampled(h); }
  int Qs(int n) const { return n * (1 + num_derivatives_); }

  py::list mean(const py::list& pts, int n) const {  // GetMeanWrapper :64-79 (no derivative rows)
    const auto P = to_vec(pts, static_cast<size_t>(n) * dim_);
    std::vector<double> out(n);
    int info = 0;
    check(cmoe_gp_posterior(h, P.data(), 1, n, nullptr, 0, out.data(), nullptr, nullptr, nullptr, nullptr, nullptr, &info), info);
    return to_list(out);
  }
  py::list grad_mean(const py::list& pts, int n) const {  // GetGradMeanWrapper :98-116 (GP's derivative rows)
    const auto P = to_vec(pts, static_cast<size_t>(n) * dim_);
    std::vector<double> out(static_cast<size_t>(dim_) * Qs(n));
    int info = 0;
    check(cmoe_gp_posterior(h, P.data(), 1, n, derivatives_.data(), num_derivatives_, nullptr, out.data(), nullptr,
                            nullptr, nullptr, nullptr, &info), info);
    return to_list(out);
  }
  py::list variance(const py::list& pts, int n) const {  // GetVarWrapper :118-154: full symmetric, row by row
    const auto P = to_vec(pts, static_cast<size_t>(n) * dim_);
    const int Q = Qs(n);
    std::vector<double> v(static_cast<size_t>(Q) * Q);
    int info = 0;
    check(cmoe_gp_posterior(h, P.data(), 1, n, derivatives_.data(), num_derivatives_, nullptr, nullptr, v.data(),
                            nullptr, nullptr, nullptr, &info), info);
    py::list out;
    for (int i = 0; i < Q; ++i)
      for (int j = 0; j < Q; ++j) out.append(v[static_cast<size_t>(j) * Q + i]);
    return out;
  }
  py::list chol_variance(const py::list& pts, int n) const {  // GetCholVarWrapper :156-187
    const auto P = to_vec(pts, static_cast<size_t>(n) * dim_);
    const int Q = Qs(n);
    std::vector<double> v(static_cast<size_t>(Q) * Q);
    int info = 0;
    check(cmoe_gp_posterior(h, P.data(), 1, n, derivatives_.data(), num_derivatives_, nullptr, nullptr, nullptr,
                            v.data(), nullptr, nullptr, &info), info);
    py::list out;
    for (int i = 0; i < Q; ++i)
      for (int j = 0; j < Q; ++j) out.append(v[static_cast<size_t>(j) * Q + i]);
    return out;
  }
  py::list grad_variance(const py::list& pts, int n, int num_derivatives, bool chol) const {  // :189-236
    const auto P = to_vec(pts, static_cast<size_t>(n) * dim_);
    const int Q = Qs(n);
    std::vector<double> g(static_cast<size_t>(dim_) * Q * Q * n);
    int info = 0;
    check(cmoe_gp_posterior(h, P.data(), 1, n, derivatives_.data(), num_derivatives_, nullptr, nullptr, nullptr, nullptr,
                            chol ? nullptr : g.data(), chol ? g.data() : nullptr, &info), info);
    g.resize(static_cast<size_t>(dim_) * Q * Q * std::min(n, std::max(0, num_derivatives)));
    return to_list(g);
  }
  void add_sampled_points(const py::list& pts, const py::list& vals, int n) {  // AddPointsToGPWrapper :238-253
    const auto P = to_vec(pts, static_cast<size_t>(n) * dim_);
    const auto V = to_vec(vals, static_cast<size_t>(n) * (1 + num_derivatives_));
    int info = 0;
    check(cmoe_gp_add_sampled_points(h, P.data(), V.data(), n, &info), info);
  }
};

cmoe_gd_params gd_of(const py::object& optimizer_parameters) {
  return optimizer_parameters.attr("optimizer_parameters").cast<const GradientDescentParameters&>().p;
}

std::vector<double> full_bounds(const py::list& domain_bounds, int dim) { return to_vec(domain_bounds, 2 * static_cast<size_t>(dim)); }

// ---- EI ---------------------------------------------------------------------------------------------------------------
double compute_expected_improvement(const GaussianProcess& gp, const py::list& pts, const py::list& being, int q, int p,
                                    int max_int_steps, double best_so_far, bool /*force_monte_carlo*/,
                                    RandomnessSourceContainer& rnd) {
  const auto X = to_vec(pts, static_cast<size_t>(q) * gp.dim_);
  const auto Xp = to_vec(being, static_cast<size_t>(p) * gp.dim_);
  double ei = 0.0;
  int info = 0;
  check(cmoe_ei_eval(gp.h, X.data(), 1, q, Xp.data(), p, max_int_steps, best_so_far, rnd.seed0(), nullptr, &ei, nullptr,
                     &info), info);
  return ei;
}

py::list compute_grad_expected_improvement(const GaussianProcess& gp, const py::list& pts, const py::list& being, int q,
                                           int p, int max_int_steps, double best_so_far, bool /*force_monte_carlo*/,
                                           RandomnessSourceContainer& rnd) {
  const auto X = to_vec(pts, static_cast<size_t>(q) * gp.dim_);
  const auto Xp = to_vec(being, static_cast<size_t>(p) * gp.dim_);
  double ei = 0.0;
  std::vector<double> g(static_cast<size_t>(q) * gp.dim_);
  int info = 0;
  check(cmoe_ei_eval(gp.h, X.data(), 1, q, Xp.data(), p, max_int_steps, best_so_far, rnd.seed0(), nullptr, &ei, g.data(),
                     &info), info);
  return to_list(g);
}

// Latin hypercube starts in a repeated tensor-product domain (gpp_random.cpp:173-197, gpp_domain.hpp:490-515)
std::vector<double> lhc_starts(const std::vector<double>& bounds, int dim, int q, int num, std::mt19937& eng) {
  std::vector<double> out(static_cast<size_t>(num) * q * dim);
  std::vector<int> idx(num);
  for (int rep = 0; rep < q; ++rep) {
    for (int d = 0; d < dim; ++d) {
      const double edge = (bounds[2 * d + 1] - bounds[2 * d]) / static_cast<double>(num);
      for (int j = 0; j < num; ++j) idx[j] = j;
      std::shuffle(idx.begin(), idx.end(), eng);
      std::uniform_real_distribution<double> u(0.0, edge);
      for (int j = 0; j < num; ++j)
        out[(static_cast<size_t>(j) * q + rep) * dim + d] = bounds[2 * d] + edge * idx[j] + u(eng);
    }
  }
  return out;
}

// Starts inside the unit simplex intersected with the box, by rejection from the (clipped) box as
// SimplexIntersectTensorProductDomain::GeneratePointInDomain does (gpp_domain.cpp:153-163): every one of the q points of a
// start is drawn until it lies in the simplex (at most 10000 attempts each, then the last draw is kept).
std::vector<double> simplex_starts(const std::vector<double>& bounds, int dim, int q, int num, std::mt19937& eng) {
  std::vector<double> out(static_cast<size_t>(num) * q * dim);
  for (size_t pt = 0; pt < static_cast<size_t>(num) * q; ++pt) {
    double* x = out.data() + pt * dim;
    for (int attempt = 0; attempt < 10000; ++attempt) {
      double sum = 0.0;
      for (int d = 0; d < dim; ++d) {
        const double lo = std::max(bounds[2 * d], 0.0), hi = std::min(bounds[2 * d + 1], 1.0);
        x[d] = std::uniform_real_distribution<double>(lo, hi)(eng);
        sum += x[d];
      }
      if (sum <= 1.0) break;
    }
  }
  return out;
}

// The reference's parallel axis is `max_num_threads` OpenMP threads over the multistart starts; here that knob selects
// how many GPUs share the starts inside ONE call: min(max_num_threads, visible devices, $CMOE_MAX_DEVICES), beginning
// with the GP's own device.  One device (the default when CMOE_MAX_DEVICES is unset) keeps the single-GPU path.
std::vector<int> devices_for(int home_device, int max_num_threads) {
  int cap = 1;
  if (const char* e = std::getenv("CMOE_MAX_DEVICES")) cap = std::max(1, std::atoi(e));
  const int count = cmoe_device_count();
  const int n = std::max(1, std::min(std::min(max_num_threads, cap), count));
  std::vector<int> devs{home_device};
  for (int d = 0; d < count && static_cast<int>(devs.size()) < n; ++d)
    if (d != home_device) devs.push_back(d);
  return devs;
}

py::list multistart_expected_improvement_optimization(const py::object& optimizer_parameters, const GaussianProcess& gp,
                                                      const py::list& domain_bounds, const py::list& being, int q, int p,
                                                      double best_so_far, int max_int_steps, int max_num_threads,
                                                      bool /*use_gpu*/, int /*which_gpu*/, RandomnessSourceContainer& rnd,
                                                      py::dict& status) {
  require_threads(max_num_threads, rnd);
  const int dim = gp.dim_;
  const auto bounds = full_bounds(domain_bounds, dim);
  const auto Xp = to_vec(being, static_cast<size_t>(p) * dim);
  const auto domain_type = optimizer_parameters.attr("domain_type").cast<DomainTypes>();
  const auto opt_type = optimizer_parameters.attr("optimizer_type").cast<OptimizerTypes>();
  const bool simplex = domain_type == DomainTypes::kSimplex;
  auto make_starts = [&](int n) {
    return simplex ? simplex_starts(bounds, dim, q, n, rnd.uniform_engine) : lhc_starts(bounds, dim, q, n, rnd.uniform_engine);
  };
  std::vector<double> best(static_cast<size_t>(q) * dim, 0.0);
  double best_value = 0.0;
  int found = 0, info = 0;
  if (opt_type == OptimizerTypes::kNull) {
    const int n = optimizer_parameters.attr("num_random_samples").cast<int>();
    const auto starts = make_starts(n);
    std::vector<double> vals(n);
    check(cmoe_ei_eval(gp.h, starts.data(), n, q, Xp.data(), p, max_int_steps, best_so_far, rnd.seed0(), nullptr,
                       vals.data(), nullptr, &info), info);
    double bv = -1.0;
    std::copy(starts.begin(), starts.begin() + static_cast<size_t>(q) * dim, best.begin());
    for (int i = 0; i < n; ++i)
      if (bv < vals[i]) {
        bv = vals[i];
        found = 1;
        std::copy(starts.begin() + static_cast<size_t>(i) * q * dim, starts.begin() + static_cast<size_t>(i + 1) * q * dim, best.begin());
      }
    status[simplex ? "lhc_simplex_domain_found_update" : "lhc_tensor_product_domain_found_update"] = static_cast<bool>(found);
  } else if (opt_type == OptimizerTypes::kGradientDescent) {
    const cmoe_gd_params gd = gd_of(optimizer_parameters);
    const auto starts = make_starts(gd.num_multistarts);
    const std::vector<int> devs = devices_for(cmoe_gp_device(gp.h), max_num_threads);
    const cmoe_multistart_opts opts{nullptr, 0, devs.data(), static_cast<int>(devs.size()),
                                    simplex ? CMOE_DOMAIN_SIMPLEX : CMOE_DOMAIN_TENSOR_PRODUCT};
    check(cmoe_multistart_ei_ex(gp.h, &gd, bounds.data(), starts.data(), gd.num_multistarts, q, Xp.data(), p,
                                max_int_steps, best_so_far, rnd.seed0(), &opts, nullptr, best.data(), &best_value, &found,
                                &info), info);
    status[simplex ? "gradient_descent_simplex_domain_found_update" : "gradient_descent_tensor_product_domain_found_update"] = static_cast<bool>(found);
  } else {
    PyErr_SetString(g_exc_base, "ERROR: invalid optimizer choice. Setting all coordinates to 0.0.");
    throw py::error_already_set();
  }
  return to_list(best);
}

py::list evaluate_EI_at_point_list(const GaussianProcess& gp, const py::list& initial_guesses, const py::list& being,
                                   int num_multistarts, int q, int p, double best_so_far, int max_int_steps,
                                   int max_num_threads, RandomnessSourceContainer& rnd, py::dict& status) {
  require_threads(max_num_threads, rnd);
  const auto starts = to_vec(initial_guesses, static_cast<size_t>(num_multistarts) * q * gp.dim_);
  const auto Xp = to_vec(being, static_cast<size_t>(p) * gp.dim_);
  std::vector<double> vals(num_multistarts);
  int info = 0;
  if (q == 1 && p == 0)  // closed-form 1-EI, as EvaluateEIAtPointList does (gpp_math.cpp:2317)
    check(cmoe_ei_analytic(gp.h, starts.data(), num_multistarts, best_so_far, vals.data(), nullptr, &info), info);
  else
    check(cmoe_ei_eval(gp.h, starts.data(), num_multistarts, q, Xp.data(), p, max_int_steps, best_so_far, rnd.seed0(),
                       nullptr, vals.data(), nullptr, &info), info);
  bool found = false;
  for (double v : vals) found = found || (v > -1.0);
  status["evaluate_EI_at_point_list"] = found;
  return to_list(vals);
}

// ---- posterior mean / KG ----------------------------------------------------------------------------------------------
std::vector<double> pad_fidelity(const std::vector<double>& pt, int dim) {
  std::vector<double> full(dim, 1.0);  // fidelity coordinates pinned to 1.0 (...knowledge_gradient_optimization.cpp:365)
  std::copy(pt.begin(), pt.end(), full.begin());
  return full;
}

double compute_posterior_mean(const GaussianProcess& gp, int num_fidelity, const py::list& pt) {
  const auto x = pad_fidelity(to_vec(pt, gp.dim_ - num_fidelity), gp.dim_);
  double m = 0.0;
  int info = 0;
  check(cmoe_gp_posterior(gp.h, x.data(), 1, 1, nullptr, 0, &m, nullptr, nullptr, nullptr, nullptr, nullptr, &info), info);
  return -m;  // the evaluator maximises -mu (...cpp:334-340)
}

// Batched form of compute_posterior_mean (not in the reference module): the examples screen 1e3-1e4 points with a
// Python loop of single-point calls (examples/main.py:147-152, 180-186); this evaluates all of them in ONE device call.
py::list compute_posterior_mean_of_points(const GaussianProcess& gp, int num_fidelity, const py::list& pts,
                                          int num_points) {
  const int ps = gp.dim_ - num_fidelity;
  const auto flat = to_vec(pts, static_cast<size_t>(num_points) * ps);
  std::vector<double> full(static_cast<size_t>(num_points) * gp.dim_, 1.0);
  for (int i = 0; i < num_points; ++i)
    std::copy(flat.begin() + static_cast<size_t>(i) * ps, flat.begin() + static_cast<size_t>(i + 1) * ps,
              full.begin() + static_cast<size_t>(i) * gp.dim_);
  std::vector<double> m(num_points);
  int info = 0;
  check(cmoe_gp_posterior(gp.h, full.data(), num_points, 1, nullptr, 0, m.data(), nullptr, nullptr, nullptr, nullptr,
                          nullptr, &info), info);
  for (double& v : m) v = -v;
  return to_list(m);
}

py::list compute_grad_posterior_mean(const GaussianProcess& gp, int num_fidelity, const py::list& pt) {
  const auto x = pad_fidelity(to_vec(pt, gp.dim_ - num_fidelity), gp.dim_);
  std::vector<double> g(gp.dim_);
  int info = 0;
  check(cmoe_gp_posterior(gp.h, x.data(), 1, 1, nullptr, 0, nullptr, g.data(), nullptr, nullptr, nullptr, nullptr, &info), info);
  std::vector<double> out(gp.dim_ - num_fidelity);
  for (size_t i = 0; i < out.size(); ++i) out[i] = -g[i];
  return to_list(out);
}

double compute_knowledge_gradient(const GaussianProcess& gp, int num_fidelity, const py::object& optimizer_parameters,
                                  const py::list& domain_bounds, const py::list& discrete_pts, const py::list& pts,
                                  const py::list& being, int num_pts, int q, int p, int max_int_steps, double best_so_far,
                                  RandomnessSourceContainer& rnd) {
  const int dim = gp.dim_, ps = dim - num_fidelity;
  const auto inner_bounds = to_vec(domain_bounds, 2 * static_cast<size_t>(ps));
  const auto D = to_vec(discrete_pts, static_cast<size_t>(num_pts) * ps);
  const auto X = to_vec(pts, static_cast<size_t>(q) * dim);
  const auto Xp = to_vec(being, static_cast<size_t>(p) * dim);
  const cmoe_gd_params inner = gd_of(optimizer_parameters);
  double kg = 0.0;
  int info = 0;
  check(cmoe_kg_eval(gp.h, num_fidelity, &inner, inner_bounds.data(), D.data(), num_pts, X.data(), 1, q, Xp.data(), p,
                     max_int_steps, best_so_far, rnd.seed0(), nullptr, &kg, nullptr, nullptr, &info), info);
  return kg;
}

py::list compute_grad_knowledge_gradient(const GaussianProcess& gp, int num_fidelity,
                                         const py::object& optimizer_parameters, const py::list& domain_bounds,
                                         const py::list& discrete_pts, const py::list& pts, const py::list& being,
                                         int num_pts, int q, int p, int max_int_steps, double best_so_far,
                                         RandomnessSourceContainer& rnd) {
  const int dim = gp.dim_, ps = dim - num_fidelity;
  const auto inner_bounds = to_vec(domain_bounds, 2 * static_cast<size_t>(ps));
  const auto D = to_vec(discrete_pts, static_cast<size_t>(num_pts) * ps);
  const auto X = to_vec(pts, static_cast<size_t>(q) * dim);
  const auto Xp = to_vec(being, static_cast<size_t>(p) * dim);
  const cmoe_gd_params inner = gd_of(optimizer_parameters);
  double kg = 0.0;
  std::vector<double> g(static_cast<size_t>(q) * dim);
  int info = 0;
  check(cmoe_kg_eval(gp.h, num_fidelity, &inner, inner_bounds.data(), D.data(), num_pts, X.data(), 1, q, Xp.data(), p,
                     max_int_steps, best_so_far, rnd.seed0(), nullptr, &kg, g.data(), nullptr, &info), info);
  return to_list(g);
}

py::list multistart_knowledge_gradient_optimization(const py::object& optimizer_parameters,
                                                    const py::object& optimizer_parameters_inner,
                                                    const GaussianProcess& gp, int num_fidelity,
                                                    const py::list& domain_bounds, const py::list& discrete_pts,
                                                    const py::list& being, int num_pts, int q, int p, double best_so_far,
                                                    int max_int_steps, int max_num_threads,
                                                    RandomnessSourceContainer& rnd, py::dict& status) {
  require_threads(max_num_threads, rnd);
  const int dim = gp.dim_, ps = dim - num_fidelity;
  const auto bounds = full_bounds(domain_bounds, dim);
  const std::vector<double> inner_bounds(bounds.begin(), bounds.begin() + 2 * ps);
  const auto D = to_vec(discrete_pts, static_cast<size_t>(num_pts) * ps);
  const auto Xp = to_vec(being, static_cast<size_t>(p) * dim);
  const auto domain_type = optimizer_parameters.attr("domain_type").cast<DomainTypes>();
  const auto opt_type = optimizer_parameters.attr("optimizer_type").cast<OptimizerTypes>();
  if (domain_type != DomainTypes::kTensorProduct) {
    PyErr_SetString(g_exc_base, "only the tensor_product domain is implemented on the B200 path");
    throw py::error_already_set();
  }
  const cmoe_gd_params inner = gd_of(optimizer_parameters_inner);
  std::vector<double> best(static_cast<size_t>(q) * dim, 0.0);
  double best_value = 0.0;
  int found = 0, info = 0;
  if (opt_type == OptimizerTypes::kNull) {
    const int n = optimizer_parameters.attr("num_random_samples").cast<int>();
    const auto starts = lhc_starts(bounds, dim, q, n, rnd.uniform_engine);
    std::vector<double> vals(n);
    check(cmoe_kg_eval(gp.h, num_fidelity, &inner, inner_bounds.data(), D.data(), num_pts, starts.data(), n, q, Xp.data(),
                       p, max_int_steps, best_so_far, rnd.seed0(), nullptr, vals.data(), nullptr, nullptr, &info), info);
    double bv = -INFINITY;
    std::copy(starts.begin(), starts.begin() + static_cast<size_t>(q) * dim, best.begin());
    for (int i = 0; i < n; ++i)
      if (bv < vals[i]) {
        bv = vals[i];
        found = 1;
        std::copy(starts.begin() + static_cast<size_t>(i) * q * dim, starts.begin() + static_cast<size_t>(i + 1) * q * dim, best.begin());
      }
    status["lhc_tensor_product_domain_found_update"] = static_cast<bool>(found);
  } else if (opt_type == OptimizerTypes::kGradientDescent) {
    const cmoe_gd_params gd = gd_of(optimizer_parameters);
    const auto starts = lhc_starts(bounds, dim, q, gd.num_multistarts, rnd.uniform_engine);
    const std::vector<int> devs = devices_for(cmoe_gp_device(gp.h), max_num_threads);
    const cmoe_multistart_opts opts{nullptr, 0, devs.data(), static_cast<int>(devs.size())};
    check(cmoe_multistart_kg_ex(gp.h, num_fidelity, &gd, &inner, bounds.data(), inner_bounds.data(), D.data(), num_pts,
                                starts.data(), gd.num_multistarts, q, Xp.data(), p, max_int_steps, best_so_far,
                                rnd.seed0(), &opts, nullptr, best.data(), &best_value, &found, &info), info);
    status["gradient_descent_tensor_product_domain_found_update"] = static_cast<bool>(found);
  } else {
    PyErr_SetString(g_exc_base, "ERROR: invalid optimizer choice. Setting all coordinates to 0.0.");
    throw py::error_already_set();
  }
  return to_list(best);
}

py::list evaluate_KG_at_point_list(const GaussianProcess& gp, int num_fidelity, const py::object& optimizer_parameters,
                                   const py::list& domain_bounds, const py::list& discrete_being_sampled,
                                   const py::list& initial_guesses, int num_multistarts, int num_pts, int q, int p,
                                   double best_so_far, int max_int_steps, int max_num_threads,
                                   RandomnessSourceContainer& rnd, py::dict& status) {
  require_threads(max_num_threads, rnd);
  const int dim = gp.dim_, ps = dim - num_fidelity;
  const auto bounds = full_bounds(domain_bounds, dim);
  const std::vector<double> inner_bounds(bounds.begin(), bounds.begin() + 2 * ps);
  // [discrete_pts (num_pts x dim) ; points_being_sampled (p x dim)], as the reference wrapper slices it (:376-390)
  const auto both = to_vec(discrete_being_sampled, static_cast<size_t>(num_pts + p) * dim);
  std::vector<double> D(static_cast<size_t>(num_pts) * ps);
  for (int j = 0; j < num_pts; ++j)
    for (int d = 0; d < ps; ++d) D[static_cast<size_t>(j) * ps + d] = both[static_cast<size_t>(j) * dim + d];
  const std::vector<double> Xp(both.begin() + static_cast<size_t>(num_pts) * dim, both.end());
  const auto starts = to_vec(initial_guesses, static_cast<size_t>(num_multistarts) * q * dim);
  const cmoe_gd_params inner = gd_of(optimizer_parameters);
  std::vector<double> vals(num_multistarts);
  int info = 0;
  check(cmoe_kg_eval(gp.h, num_fidelity, &inner, inner_bounds.data(), D.data(), num_pts, starts.data(), num_multistarts, q,
                     Xp.data(), p, max_int_steps, best_so_far, rnd.seed0(), nullptr, vals.data(), nullptr, nullptr, &info),
        info);
  bool found = false;
  for (double v : vals) found = found || std::isfinite(v);
  status["evaluate_KG_at_point_list"] = found;
  return to_list(vals);
}

// ---- GaussianProcessMCMC + MCMC-averaged acquisition ---------------------------------------------------------------------
// (gpp_python_knowledge_gradient_mcmc.cpp:49-384, gpp_python_expected_improvement_mcmc.cpp:46-300)
struct GaussianProcessMCMC {
  std::vector<cmoe_gp*> members;
  int dim_ = 0, num_derivatives_ = 0, num_sampled_ = 0;
  std::vector<int> derivatives_;
  GaussianProcessMCMC(const py::list& hyperparameters_list, const py::list& noise_variance_list,
                      const py::list& points_sampled, const py::list& points_sampled_value, const py::list& derivatives,
                      int num_mcmc, int num_derivatives, int dim, int num_sampled, const std::string& kernel,
                      int which_gpu) {
    // hyperparameters_list: num_mcmc x (alpha, length_1..length_dim) flat; noise_variance_list: num_mcmc x (1+g) flat
    const auto hyp = to_vec(hyperparameters_list, static_cast<size_t>(num_mcmc) * (dim + 1));
    const auto noise = to_vec(noise_variance_list, static_cast<size_t>(num_mcmc) * (1 + num_derivatives));
    const auto X = to_vec(points_sampled, static_cast<size_t>(dim) * num_sampled);
    const auto y = to_vec(points_sampled_value, static_cast<size_t>(num_sampled) * (1 + num_derivatives));
    derivatives_ = to_ivec(derivatives, num_derivatives);
    dim_ = dim;
    num_derivatives_ = num_derivatives;
    num_sampled_ = num_sampled;
    const int kid = (kernel == "square_exponential") ? CMOE_KERNEL_SQUARE_EXPONENTIAL : CMOE_KERNEL_MATERN_NU_2P5;
    for (int m = 0; m < num_mcmc; ++m) {
      cmoe_gp* h = nullptr;
      int info = 0;
      const int rc = cmoe_gp_create(kid, hyp[static_cast<size_t>(m) * (dim + 1)], hyp.data() + static_cast<size_t>(m) * (dim + 1) + 1,
                                    X.data(), y.data(), noise.data() + static_cast<size_t>(m) * (1 + num_derivatives),
                                    derivatives_.data(), num_derivatives, dim, num_sampled, which_gpu, &h, &info);
      if (rc != CMOE_OK) {
        for (cmoe_gp* g : members) cmoe_gp_destroy(g);
        members.clear();
        check(rc, info);
      }
      members.push_back(h);
    }
  }
  ~GaussianProcessMCMC() {
    for (cmoe_gp* g : members) cmoe_gp_destroy(g);
  }
  GaussianProcessMCMC(const GaussianProcessMCMC&) = delete;
  GaussianProcessMCMC& operator=(const GaussianProcessMCMC&) = delete;
  int num_mcmc() const { return static_cast<int>(members.size()); }
  const cmoe_gp* const* handles() const { return members.data(); }
};

double compute_knowledge_gradient_mcmc(const GaussianProcessMCMC& gp, int num_fidelity, const py::object& optimizer_parameters,
                                       const py::list& domain_bounds, const py::list& discrete_pts, const py::list& pts,
                                       const py::list& being, int num_pts, int q, int p, int max_int_steps,
                                       const py::list& best_so_far, RandomnessSourceContainer& rnd) {
  const int dim = gp.dim_, ps = dim - num_fidelity, M = gp.num_mcmc();
  const auto inner_bounds = to_vec(domain_bounds, 2 * static_cast<size_t>(ps));
  const auto D = to_vec(discrete_pts, static_cast<size_t>(M) * num_pts * ps);
  const auto X = to_vec(pts, static_cast<size_t>(q) * dim);
  const auto Xp = to_vec(being, static_cast<size_t>(p) * dim);
  const auto best = to_vec(best_so_far, M);
  const cmoe_gd_params inner = gd_of(optimizer_parameters);
  double kg = 0.0;
  int info = 0;
  check(cmoe_kg_eval_mcmc(gp.handles(), M, num_fidelity, &inner, inner_bounds.data(), D.data(), num_pts, X.data(), 1, q,
                          Xp.data(), p, max_int_steps, best.data(), rnd.seed0(), nullptr, &kg, nullptr, &info), info);
  return kg;
}

py::list compute_grad_knowledge_gradient_mcmc(const GaussianProcessMCMC& gp, int num_fidelity,
                                              const py::object& optimizer_parameters, const py::list& domain_bounds,
                                              const py::list& discrete_pts, const py::list& pts, const py::list& being,
                                              int num_pts, int q, int p, int max_int_steps, const py::list& best_so_far,
                                              RandomnessSourceContainer& rnd) {
  const int dim = gp.dim_, ps = dim - num_fidelity, M = gp.num_mcmc();
  const auto inner_bounds = to_vec(domain_bounds, 2 * static_cast<size_t>(ps));
  const auto D = to_vec(discrete_pts, static_cast<size_t>(M) * num_pts * ps);
  const auto X = to_vec(pts, static_cast<size_t>(q) * dim);
  const auto Xp = to_vec(being, static_cast<size_t>(p) * dim);
  const auto best = to_vec(best_so_far, M);
  const cmoe_gd_params inner = gd_of(optimizer_parameters);
  double kg = 0.0;
  std::vector<double> g(static_cast<size_t>(q) * dim);
  int info = 0;
  check(cmoe_kg_eval_mcmc(gp.handles(), M, num_fidelity, &inner, inner_bounds.data(), D.data(), num_pts, X.data(), 1, q,
                          Xp.data(), p, max_int_steps, best.data(), rnd.seed0(), nullptr, &kg, g.data(), &info), info);
  return to_list(g);
}

template <typename EvalAll, typename Multistart>
py::list mcmc_optimization_common(const py::object& optimizer_parameters, const std::vector<double>& bounds, int dim,
                                  int q, double init_best, RandomnessSourceContainer& rnd, py::dict& status,
                                  EvalAll&& eval_all, Multistart&& multistart) {
  const auto domain_type = optimizer_parameters.attr("domain_type").cast<DomainTypes>();
  const auto opt_type = optimizer_parameters.attr("optimizer_type").cast<OptimizerTypes>();
  if (domain_type != DomainTypes::kTensorProduct) {
    PyErr_SetString(g_exc_base, "only the tensor_product domain is implemented on the B200 path");
    throw py::error_already_set();
  }
  std::vector<double> best(static_cast<size_t>(q) * dim, 0.0);
  int found = 0;
  if (opt_type == OptimizerTypes::kNull) {
    const int n = optimizer_parameters.attr("num_random_samples").cast<int>();
    const auto starts = lhc_starts(bounds, dim, q, n, rnd.uniform_engine);
    std::vector<double> vals(n);
    eval_all(starts, n, vals);
    double bv = init_best;
    std::copy(starts.begin(), starts.begin() + static_cast<size_t>(q) * dim, best.begin());
    for (int i = 0; i < n; ++i)
      if (bv < vals[i]) {
        bv = vals[i];
        found = 1;
        std::copy(starts.begin() + static_cast<size_t>(i) * q * dim, starts.begin() + static_cast<size_t>(i + 1) * q * dim, best.begin());
      }
    status["lhc_tensor_product_domain_found_update"] = static_cast<bool>(found);
  } else if (opt_type == OptimizerTypes::kGradientDescent) {
    const cmoe_gd_params gd = gd_of(optimizer_parameters);
    const auto starts = lhc_starts(bounds, dim, q, gd.num_multistarts, rnd.uniform_engine);
    multistart(gd, starts, best, found);
    status["gradient_descent_tensor_product_domain_found_update"] = static_cast<bool>(found);
  } else {
    PyErr_SetString(g_exc_base, "ERROR: invalid optimizer choice. Setting all coordinates to 0.0.");
    throw py::error_already_set();
  }
  return to_list(best);
}

py::list multistart_knowledge_gradient_mcmc_optimization(const py::object& optimizer_parameters,
                                                         const py::object& optimizer_parameters_inner,
                                                         const GaussianProcessMCMC& gp, int num_fidelity,
                                                         const py::list& domain_bounds, const py::list& discrete_pts,
                                                         const py::list& being, int num_pts, int q, int p,
                                                         const py::list& best_so_far, int max_int_steps,
                                                         int max_num_threads, RandomnessSourceContainer& rnd,
                                                         py::dict& status) {
  require_threads(max_num_threads, rnd);
  const int dim = gp.dim_, ps = dim - num_fidelity, M = gp.num_mcmc();
  const auto bounds = full_bounds(domain_bounds, dim);
  const std::vector<double> inner_bounds(bounds.begin(), bounds.begin() + 2 * ps);
  const auto D = to_vec(discrete_pts, static_cast<size_t>(M) * num_pts * ps);
  const auto Xp = to_vec(being, static_cast<size_t>(p) * dim);
  const auto best = to_vec(best_so_far, M);
  const cmoe_gd_params inner = gd_of(optimizer_parameters_inner);
  return mcmc_optimization_common(
      optimizer_parameters, bounds, dim, q, -INFINITY, rnd, status,
      [&](const std::vector<double>& starts, int n, std::vector<double>& vals) {
        int info = 0;
        check(cmoe_kg_eval_mcmc(gp.handles(), M, num_fidelity, &inner, inner_bounds.data(), D.data(), num_pts,
                                starts.data(), n, q, Xp.data(), p, max_int_steps, best.data(), rnd.seed0(), nullptr,
                                vals.data(), nullptr, &info), info);
      },
      [&](const cmoe_gd_params& gd, const std::vector<double>& starts, std::vector<double>& out, int& found) {
        int info = 0;
        double bv = 0.0;
        check(cmoe_multistart_kg_mcmc(gp.handles(), M, num_fidelity, &gd, &inner, bounds.data(), inner_bounds.data(),
                                      D.data(), num_pts, starts.data(), gd.num_multistarts, q, Xp.data(), p,
                                      max_int_steps, best.data(), rnd.seed0(), nullptr, out.data(), &bv, &found, &info),
              info);
      });
}

py::list evaluate_KG_mcmc_at_point_list(const GaussianProcessMCMC& gp, int num_fidelity,
                                        const py::object& optimizer_parameters, const py::list& domain_bounds,
                                        const py::list& initial_guesses, const py::list& discrete_being_sampled,
                                        int num_multistarts, int num_pts, int q, int p, const py::list& best_so_far,
                                        int max_int_steps, int max_num_threads, RandomnessSourceContainer& rnd,
                                        py::dict& status) {
  require_threads(max_num_threads, rnd);
  const int dim = gp.dim_, ps = dim - num_fidelity, M = gp.num_mcmc();
  const auto bounds = full_bounds(domain_bounds, dim);
  const std::vector<double> inner_bounds(bounds.begin(), bounds.begin() + 2 * ps);
  // flat [num_mcmc x num_pts x (dim - nf) discrete points ; p x dim points being sampled], as the reference slices it (:349-374)
  const size_t nd = static_cast<size_t>(M) * num_pts * ps;
  const auto both = to_vec(discrete_being_sampled, nd + static_cast<size_t>(p) * dim);
  const std::vector<double> D(both.begin(), both.begin() + nd);
  const std::vector<double> Xp(both.begin() + nd, both.end());
  const auto starts = to_vec(initial_guesses, static_cast<size_t>(num_multistarts) * q * dim);
  const auto best = to_vec(best_so_far, M);
  const cmoe_gd_params inner = gd_of(optimizer_parameters);
  std::vector<double> vals(num_multistarts);
  int info = 0;
  check(cmoe_kg_eval_mcmc(gp.handles(), M, num_fidelity, &inner, inner_bounds.data(), D.data(), num_pts, starts.data(),
                          num_multistarts, q, Xp.data(), p, max_int_steps, best.data(), rnd.seed0(), nullptr, vals.data(),
                          nullptr, &info), info);
  bool found = false;
  for (double v : vals) found = found || (v > -INFINITY);
  status["evaluate_KG_at_point_list"] = found;
  return to_list(vals);
}

double compute_expected_improvement_mcmc(const GaussianProcessMCMC& gp, const py::list& pts, const py::list& being, int q,
                                         int p, int max_int_steps, const py::list& best_so_far,
                                         RandomnessSourceContainer& rnd) {
  const auto X = to_vec(pts, static_cast<size_t>(q) * gp.dim_);
  const auto Xp = to_vec(being, static_cast<size_t>(p) * gp.dim_);
  const auto best = to_vec(best_so_far, gp.num_mcmc());
  double ei = 0.0;
  int info = 0;
  check(cmoe_ei_eval_mcmc(gp.handles(), gp.num_mcmc(), X.data(), 1, q, Xp.data(), p, max_int_steps, best.data(),
                          rnd.seed0(), nullptr, 0, &ei, nullptr, &info), info);
  return ei;
}

py::list compute_grad_expected_improvement_mcmc(const GaussianProcessMCMC& gp, const py::list& pts, const py::list& being,
                                                int q, int p, int max_int_steps, const py::list& best_so_far,
                                                RandomnessSourceContainer& rnd) {
  const auto X = to_vec(pts, static_cast<size_t>(q) * gp.dim_);
  const auto Xp = to_vec(being, static_cast<size_t>(p) * gp.dim_);
  const auto best = to_vec(best_so_far, gp.num_mcmc());
  double ei = 0.0;
  std::vector<double> g(static_cast<size_t>(q) * gp.dim_);
  int info = 0;
  check(cmoe_ei_eval_mcmc(gp.handles(), gp.num_mcmc(), X.data(), 1, q, Xp.data(), p, max_int_steps, best.data(),
                          rnd.seed0(), nullptr, 0, &ei, g.data(), &info), info);
  return to_list(g);
}

py::list multistart_expected_improvement_mcmc_optimization(const py::object& optimizer_parameters,
                                                           const GaussianProcessMCMC& gp, const py::list& domain_bounds,
                                                           const py::list& being, int q, int p,
                                                           const py::list& best_so_far, int max_int_steps,
                                                           int max_num_threads, RandomnessSourceContainer& rnd,
                                                           py::dict& status) {
  require_threads(max_num_threads, rnd);
  const int dim = gp.dim_, M = gp.num_mcmc();
  const auto bounds = full_bounds(domain_bounds, dim);
  const auto Xp = to_vec(being, static_cast<size_t>(p) * dim);
  const auto best = to_vec(best_so_far, M);
  return mcmc_optimization_common(
      optimizer_parameters, bounds, dim, q, 0.0, rnd, status,
      [&](const std::vector<double>& starts, int n, std::vector<double>& vals) {
        int info = 0;
        check(cmoe_ei_eval_mcmc(gp.handles(), M, starts.data(), n, q, Xp.data(), p, max_int_steps, best.data(),
                                rnd.seed0(), nullptr, 1, vals.data(), nullptr, &info), info);
      },
      [&](const cmoe_gd_params& gd, const std::vector<double>& starts, std::vector<double>& out, int& found) {
        int info = 0;
        double bv = 0.0;
        check(cmoe_multistart_ei_mcmc(gp.handles(), M, &gd, bounds.data(), starts.data(), gd.num_multistarts, q,
                                      Xp.data(), p, max_int_steps, best.data(), rnd.seed0(), nullptr, out.data(), &bv,
                                      &found, &info), info);
      });
}

py::list evaluate_EI_mcmc_at_point_list(const GaussianProcessMCMC& gp, const py::list& initial_guesses,
                                        const py::list& being, int num_multistarts, int q, int p,
                                        const py::list& best_so_far, int max_int_steps, int max_num_threads,
                                        RandomnessSourceContainer& rnd, py::dict& status) {
  require_threads(max_num_threads, rnd);
  const auto starts = to_vec(initial_guesses, static_cast<size_t>(num_multistarts) * q * gp.dim_);
  const auto Xp = to_vec(being, static_cast<size_t>(p) * gp.dim_);
  const auto best = to_vec(best_so_far, gp.num_mcmc());
  std::vector<double> vals(num_multistarts);
  int info = 0;
  check(cmoe_ei_eval_mcmc(gp.handles(), gp.num_mcmc(), starts.data(), num_multistarts, q, Xp.data(), p, max_int_steps,
                          best.data(), rnd.seed0(), nullptr, 1, vals.data(), nullptr, &info), info);
  bool found = false;
  for (double v : vals) found = found || (v > 0.0);
  status["evaluate_EI_at_point_list"] = found;
  return to_list(vals);
}

[[noreturn]] void not_on_path(const char* name) {
  PyErr_SetString(g_exc_base, (std::string(name) + " is outside the B200 hot path (SURVEY.md 8f) and is not provided by this build").c_str());
  throw py::error_already_set();
}

}  // namespace

PYBIND11_MODULE(GPP, m) {
  m.doc() = "B200-native drop-in for moe.build.GPP (GP posterior + MC acquisition hot path)";
  g_exc_base = PyErr_NewExceptionWithDoc("GPP.OptimalLearningException",
                                         "Base exception class for errors raised from the optimal_learning library.",
                                         PyExc_Exception, nullptr);
  g_exc_bounds = PyErr_NewExceptionWithDoc("GPP.BoundsException", "value not in range [min, max].", g_exc_base, nullptr);
  g_exc_invalid = PyErr_NewExceptionWithDoc("GPP.InvalidValueException", "value != truth (+/- tolerance)", g_exc_base, nullptr);
  g_exc_singular = PyErr_NewExceptionWithDoc("GPP.SingularMatrixException", "num_rows X num_cols matrix is singular", g_exc_base, nullptr);
  m.attr("OptimalLearningException") = py::handle(g_exc_base);
  m.attr("BoundsException") = py::handle(g_exc_bounds);
  m.attr("InvalidValueException") = py::handle(g_exc_invalid);
  m.attr("SingularMatrixException") = py::handle(g_exc_singular);

  py::enum_<OptimizerTypes>(m, "OptimizerTypes")
      .value("null", OptimizerTypes::kNull)
      .value("gradient_descent", OptimizerTypes::kGradientDescent)
      .value("newton", OptimizerTypes::kNewton);
  py::enum_<DomainTypes>(m, "DomainTypes")
      .value("tensor_product", DomainTypes::kTensorProduct)
      .value("simplex", DomainTypes::kSimplex);
  py::enum_<LogLikelihoodTypes>(m, "LogLikelihoodTypes")
      .value("log_marginal_likelihood", LogLikelihoodTypes::kLogMarginalLikelihood)
      .value("leave_one_out_log_likelihood", LogLikelihoodTypes::kLeaveOneOutLogLikelihood);

  py::class_<GradientDescentParameters>(m, "GradientDescentParameters")
      .def(py::init<int, int, int, int, double, double, double, double>())
      .def_property("num_multistarts", [](const GradientDescentParameters& s) { return s.p.num_multistarts; },
                    [](GradientDescentParameters& s, int v) { s.p.num_multistarts = v; })
      .def_property("max_num_steps", [](const GradientDescentParameters& s) { return s.p.max_num_steps; },
                    [](GradientDescentParameters& s, int v) { s.p.max_num_steps = v; })
      .def_property("max_num_restarts", [](const GradientDescentParameters& s) { return s.p.max_num_restarts; },
                    [](GradientDescentParameters& s, int v) { s.p.max_num_restarts = v; })
      .def_property("num_steps_averaged", [](const GradientDescentParameters& s) { return s.p.num_steps_averaged; },
                    [](GradientDescentParameters& s, int v) { s.p.num_steps_averaged = v; })
      .def_property("gamma", [](const GradientDescentParameters& s) { return s.p.gamma; },
                    [](GradientDescentParameters& s, double v) { s.p.gamma = v; })
      .def_property("pre_mult", [](const GradientDescentParameters& s) { return s.p.pre_mult; },
                    [](GradientDescentParameters& s, double v) { s.p.pre_mult = v; })
      .def_property("max_relative_change", [](const GradientDescentParameters& s) { return s.p.max_relative_change; },
                    [](GradientDescentParameters& s, double v) { s.p.max_relative_change = v; })
      .def_property("tolerance", [](const GradientDescentParameters& s) { return s.p.tolerance; },
                    [](GradientDescentParameters& s, double v) { s.p.tolerance = v; });
  py::class_<NewtonParameters>(m, "NewtonParameters")
      .def(py::init<int, int, double, double, double, double>())
      .def_readwrite("num_multistarts", &NewtonParameters::num_multistarts)
      .def_readwrite("max_num_steps", &NewtonParameters::max_num_steps)
      .def_readwrite("gamma", &NewtonParameters::gamma)
      .def_readwrite("time_factor", &NewtonParameters::time_factor)
      .def_readwrite("max_relative_change", &NewtonParameters::max_relative_change)
      .def_readwrite("tolerance", &NewtonParameters::tolerance);
  py::class_<RandomnessSourceContainer>(m, "RandomnessSourceContainer")
      .def(py::init<int>())
      .def("SetExplicitUniformGeneratorSeed", &RandomnessSourceContainer::SetExplicitUniformGeneratorSeed)
      .def("SetRandomizedUniformGeneratorSeed", &RandomnessSourceContainer::SetRandomizedUniformGeneratorSeed)
      .def("ResetUniformRNGSeed", &RandomnessSourceContainer::ResetUniformGeneratorState)
      .def("SetExplicitNormalRNGSeed", &RandomnessSourceContainer::SetExplicitNormalRNGSeed)
      .def("SetRandomizedNormalRNGSeed", &RandomnessSourceContainer::SetRandomizedNormalRNGSeed)
      .def("SetNormalRNGSeedPythonList", &RandomnessSourceContainer::SetNormalRNGSeedPythonList)
      .def("ResetNormalRNGSeed", &RandomnessSourceContainer::ResetNormalRNGState)
      .def("PrintState", &RandomnessSourceContainer::PrintState);

  py::class_<GaussianProcess>(m, "GaussianProcess")
      .def(py::init<const py::list&, const py::list&, const py::list&, const py::list&, const py::list&, int, int, int,
                    const std::string&, int>(),
           py::arg("hyperparameters"), py::arg("points_sampled"), py::arg("points_sampled_value"),
           py::arg("noise_variance"), py::arg("derivatives"), py::arg("num_derivatives"), py::arg("dim"),
           py::arg("num_sampled"), py::arg("kernel") = "matern52", py::arg("which_gpu") = 0)
      .def_property_readonly("dim", &GaussianProcess::dim)
      .def_property_readonly("num_sampled", &GaussianProcess::num_sampled)
      .def("compute_mean_of_points", &GaussianProcess::mean)
      .def("compute_mean_of_additional_points", &GaussianProcess::mean)
      .def("compute_grad_mean_of_points", &GaussianProcess::grad_mean)
      .def("compute_variance_of_points", &GaussianProcess::variance)
      .def("compute_cholesky_variance_of_points", &GaussianProcess::chol_variance)
      .def("compute_grad_variance_of_points",
           [](const GaussianProcess& g, const py::list& pts, int n, int nd) { return g.grad_variance(pts, n, nd, false); })
      .def("compute_grad_cholesky_variance_of_points",
           [](const GaussianProcess& g, const py::list& pts, int n, int nd) { return g.grad_variance(pts, n, nd, true); })
      .def("add_sampled_points", &GaussianProcess::add_sampled_points)
      .def("sample_point_from_gp", [](GaussianProcess&, const py::list&) -> py::list { not_on_path("sample_point_from_gp"); })
      .def("sample_global_optima", [](GaussianProcess&, int, int, const py::list&) -> py::list { not_on_path("sample_global_optima"); })
      .def("set_explicit_seed", [](GaussianProcess&, uint32_t) {})
      .def("set_randomized_seed", [](GaussianProcess&, uint32_t) {})
      .def("reset_to_most_recent_seed", [](GaussianProcess&) {})
      .def("print_historical_data", [](GaussianProcess&) {});

  m.def("compute_expected_improvement", &compute_expected_improvement);
  m.def("compute_grad_expected_improvement", &compute_grad_expected_improvement);
  m.def("multistart_expected_improvement_optimization", &multistart_expected_improvement_optimization);
  m.def("evaluate_EI_at_point_list", &evaluate_EI_at_point_list);
  m.def("compute_posterior_mean", &compute_posterior_mean);
  m.def("compute_grad_posterior_mean", &compute_grad_posterior_mean);
  m.def("compute_posterior_mean_of_points", &compute_posterior_mean_of_points);
  m.def("compute_knowledge_gradient", &compute_knowledge_gradient);
  m.def("compute_grad_knowledge_gradient", &compute_grad_knowledge_gradient);
  m.def("multistart_knowledge_gradient_optimization", &multistart_knowledge_gradient_optimization);
  m.def("evaluate_KG_at_point_list", &evaluate_KG_at_point_list);
  m.def("posterior_mean_optimization",
        [](const GaussianProcess& gp, int num_fidelity, const py::object& optimizer_parameters,
           const py::list& domain_bounds, const py::list& initial_guess, py::dict& status) -> py::list {
          const int ps = gp.dim_ - num_fidelity;
          const auto bounds = to_vec(domain_bounds, 2 * static_cast<size_t>(ps));
          const auto x0 = to_vec(initial_guess, ps);
          const cmoe_gd_params gd = gd_of(optimizer_parameters);
          std::vector<double> best(ps, 0.0);
          double value = 0.0;
          int found = 0;
          check(cmoe_posterior_mean_optimization(gp.h, num_fidelity, &gd, bounds.data(), x0.data(), best.data(), &value,
                                                 &found));
          (void)status;  // the reference leaves `status` untouched for this entry point (:306-350)
          return to_list(best);
        });
  py::class_<GaussianProcessMCMC>(m, "GaussianProcessMCMC")
      .def(py::init<const py::list&, const py::list&, const py::list&, const py::list&, const py::list&, int, int, int,
                    int, const std::string&, int>(),
           py::arg("hyperparameters_list"), py::arg("noise_variance_list"), py::arg("points_sampled"),
           py::arg("points_sampled_value"), py::arg("derivatives"), py::arg("num_mcmc"), py::arg("num_derivatives"),
           py::arg("dim"), py::arg("num_sampled"), py::arg("kernel") = "matern52", py::arg("which_gpu") = 0)
      .def_property_readonly("dim", [](const GaussianProcessMCMC& g) { return g.dim_; })
      .def_property_readonly("num_mcmc", &GaussianProcessMCMC::num_mcmc)
      .def_property_readonly("num_sampled", [](const GaussianProcessMCMC& g) { return g.num_sampled_; });
  m.def("compute_knowledge_gradient_mcmc", &compute_knowledge_gradient_mcmc);
  m.def("compute_grad_knowledge_gradient_mcmc", &compute_grad_knowledge_gradient_mcmc);
  m.def("multistart_knowledge_gradient_mcmc_optimization", &multistart_knowledge_gradient_mcmc_optimization);
  m.def("evaluate_KG_mcmc_at_point_list", &evaluate_KG_mcmc_at_point_list);
  m.def("compute_expected_improvement_mcmc", &compute_expected_improvement_mcmc);
  m.def("compute_grad_expected_improvement_mcmc", &compute_grad_expected_improvement_mcmc);
  m.def("multistart_expected_improvement_mcmc_optimization", &multistart_expected_improvement_mcmc_optimization);
  m.def("evaluate_EI_mcmc_at_point_list", &evaluate_EI_mcmc_at_point_list);
  // names of the reference module that are outside the hot path (model selection, heuristic EI): present, so that an
  // unmodified front end fails with the library's own exception class and a clear message rather than AttributeError
  // compute_log_likelihood(points_sampled, points_sampled_value, dim, num_sampled, objective_type, hyperparameters,
  //                        derivatives, num_derivatives, noise_variance)   gpp_python_model_selection.cpp:43-87
  m.def("compute_log_likelihood",
        [](const py::list& points_sampled, const py::list& points_sampled_value, int dim, int num_sampled,
           LogLikelihoodTypes objective_type, const py::list& hyperparameters, const py::list& derivatives,
           int num_derivatives, const py::list& noise_variance) -> double {
          if (objective_type != LogLikelihoodTypes::kLogMarginalLikelihood) {
            PyErr_SetString(g_exc_base, "ERROR: invalid objective mode choice. Setting log likelihood to -DBL_MAX.");
            throw py::error_already_set();
          }
          const double alpha = hyperparameters[0].cast<double>();
          const auto lengths = to_vec(hyperparameters[1].cast<py::list>(), dim);
          const auto X = to_vec(points_sampled, static_cast<size_t>(dim) * num_sampled);
          const auto y = to_vec(points_sampled_value, static_cast<size_t>(num_sampled) * (1 + num_derivatives));
          const auto noise = to_vec(noise_variance, 1 + num_derivatives);
          const auto derivs = to_ivec(derivatives, num_derivatives);
          double value = 0.0;
          int info = 0;
          // the reference's Python boundary hard-wires MaternNu2p5 here as well (:57)
          check(cmoe_log_marginal_likelihood(CMOE_KERNEL_MATERN_NU_2P5, alpha, lengths.data(), X.data(), y.data(),
                                             noise.data(), derivs.data(), num_derivatives, dim, num_sampled, 0, &value,
                                             &info), info);
          return value;
        });
  // compute_hyperparameter_grad_log_likelihood(... same arguments ...)   gpp_python_model_selection.cpp:89-140
  m.def("compute_hyperparameter_grad_log_likelihood",
        [](const py::list& points_sampled, const py::list& points_sampled_value, int dim, int num_sampled,
           LogLikelihoodTypes objective_type, const py::list& hyperparameters, const py::list& derivatives,
           int num_derivatives, const py::list& noise_variance) -> py::list {
          if (objective_type != LogLikelihoodTypes::kLogMarginalLikelihood) {
            PyErr_SetString(g_exc_base, "ERROR: invalid objective mode choice. Setting all gradients to DBL_MAX.");
            throw py::error_already_set();
          }
          const double alpha = hyperparameters[0].cast<double>();
          const auto lengths = to_vec(hyperparameters[1].cast<py::list>(), dim);
          const auto X = to_vec(points_sampled, static_cast<size_t>(dim) * num_sampled);
          const auto y = to_vec(points_sampled_value, static_cast<size_t>(num_sampled) * (1 + num_derivatives));
          const auto noise = to_vec(noise_variance, 1 + num_derivatives);
          const auto derivs = to_ivec(derivatives, num_derivatives);
          std::vector<double> grad(static_cast<size_t>(dim) + 2 + num_derivatives);
          int info = 0;
          check(cmoe_grad_log_marginal_likelihood(CMOE_KERNEL_MATERN_NU_2P5, alpha, lengths.data(), X.data(), y.data(),
                                                  noise.data(), derivs.data(), num_derivatives, dim, num_sampled, 0,
                                                  grad.data(), &info), info);
          return to_list(grad);
        });
  for (const char* name : {"multistart_hyperparameter_optimization", "restarted_hyperparameter_optimization",
                           "evaluate_log_likelihood_at_hyperparameter_list",
                           "heuristic_expected_improvement_optimization"}) {
    const std::string n(name);
    m.def(name, [n](const py::args&, const py::kwargs&) -> py::object { not_on_path(n.c_str()); });
  }
  m.def("run_cpp_tests", []() -> int { not_on_path("run_cpp_tests"); });
  m.def("device_count", []() { return cmoe_device_count(); });
  m.def("version", []() { return std::string(cmoe_version()); });
}
