// posterior_mean_optimization: minimise the GP posterior mean from one start point with the reference's line-search
// gradient descent (the recommendation step of the BayesOpt loop).
//
// Replaces ComputeOptimalPosteriorMean with num_starts = 1 on the un-fantasised GP
// (gpp_knowledge_gradient_optimization.cpp:420-472; PosteriorMeanEvaluator :334-359; line search
// gpp_optimization.hpp:708-828, 1242-1283; LimitUpdate gpp_domain.cpp:64-104); Python boundary
// posterior_mean_optimization (gpp_python_knowledge_gradient.cpp:306-350).
//
// The problem is inherently sequential (one trajectory), so it runs as ONE warp: lanes split the n rows of
// k(x, X) . K^-1 y, shuffle-reduce value and gradient, and every lane carries the (uniform) optimiser state.
#include <cmath>

#include "device_math.cuh"
#include "internal.cuh"

namespace cmoe {
namespace {

struct PmOptParams {
  int N, ps, max_steps, max_restarts;
  double mean, mrc, tol;
  double lo[CMOE_MAX_DIM], hi[CMOE_MAX_DIM];
};

__device__ __forceinline__ double limit_step_pm(double step, double x, double lo, double hi, double mrc) {
  double dist = fmin(x - lo, hi - x);
  if (fabs(step) > mrc * dist) step = copysign(mrc * dist, step);
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

// f(x) = -mu(x), grad f = -dmu/dx on the free coordinates; all lanes return the same values
__device__ double pm_eval(const KernelSpec& spec, const PmOptParams& prm, const double* __restrict__ X,
                          const double* __restrict__ beta, const double* xq, double* grad /* [dim] or null */) {
  const int dim = spec.dim, b1 = 1 + spec.g, n = prm.N * b1, lane = threadIdx.x & 31;
  double val = 0.0;
  double g[CMOE_MAX_DIM];
  for (int d = 0; d < dim; ++d) g[d] = 0.0;
  for (int row = lane; row < n; row += 32) {
    const int j = row / b1, ty = (row % b1) ? spec.derivs[row % b1 - 1] : -1;
    const double* xj = X + static_cast<size_t>(j) * dim;
    const KParts kp = kernel_parts(spec, weighted_sqdist(spec, xj, xq));
    val += cov_entry(spec, kp, xj, xq, ty, -1) * beta[row];
    if (grad) {
      const KParts kq = kernel_parts(spec, weighted_sqdist(spec, xq, xj));
      for (int d = 0; d < prm.ps; ++d) g[d] += grad_cov_entry(spec, kq, xq, xj, -1, ty, d) * beta[row];
    }
  }
  val = warp_sum(val);
  if (grad)
    for (int d = 0; d < dim; ++d) grad[d] = (d < prm.ps) ? -warp_sum(g[d]) : 0.0;
  return -(prm.mean + val);
}

__global__ void __launch_bounds__(32) pm_opt_kernel(const __grid_constant__ KernelSpec spec,
                                                    const __grid_constant__ PmOptParams prm,
                                                    const double* __restrict__ X, const double* __restrict__ beta,
                                                    const double* __restrict__ alpha0, const double* __restrict__ x0,
                                                    double* __restrict__ out) {
  const int dim = spec.dim, ps = prm.ps;
  double x[CMOE_MAX_DIM], gb[CMOE_MAX_DIM], xt[CMOE_MAX_DIM], step[CMOE_MAX_DIM], run0[CMOE_MAX_DIM];
  for (int d = 0; d < dim; ++d) x[d] = x0[d];
  double fb = pm_eval(spec, prm, X, beta, x, gb);
  if (prm.max_restarts > 0) {
    const double step_tol = prm.max_steps > 0 ? prm.tol / static_cast<double>(prm.max_steps) : 0.0;
    for (int r = 0; r < prm.max_restarts; ++r) {
      for (int d = 0; d < dim; ++d) run0[d] = x[d];
      for (int i = 0; i < prm.max_steps; ++i) {
        double alpha = alpha0[i];
        double nrm = 0.0;
        for (int d = 0; d < ps; ++d) nrm += gb[d] * gb[d];
        int search = 0;
        double obj = 0.0;
        while (search < 30) {
          for (int d = 0; d < dim; ++d) xt[d] = x[d] + ((d < ps) ? alpha * gb[d] : 0.0);
          obj = pm_eval(spec, prm, X, beta, xt, nullptr);
          if (obj - fb > 0.5 * alpha * nrm) break;
          alpha *= 0.5;
          search += 1;
        }
        for (int d = 0; d < dim; ++d)
          step[d] = (d < ps) ? limit_step_pm(alpha * gb[d], x[d], prm.lo[d], prm.hi[d], prm.mrc) : 0.0;
        for (int d = 0; d < dim; ++d) xt[d] = x[d] + step[d];
        obj = pm_eval(spec, prm, X, beta, xt, nullptr);
        if (obj <= fb || search == 30) break;
        double ns = 0.0;
        for (int d = 0; d < dim; ++d) {
          x[d] += step[d];
          ns += step[d] * step[d];
        }
        fb = pm_eval(spec, prm, X, beta, x, gb);
        if (sqrt(ns) < step_tol) break;
      }
      double nd = 0.0;
      for (int d = 0; d < ps; ++d) nd += (run0[d] - x[d]) * (run0[d] - x[d]);
      if (sqrt(nd) <= prm.tol) break;
    }
  }
  if (threadIdx.x == 0) {
    for (int d = 0; d < ps; ++d) out[d] = x[d];
    out[ps] = fb;
  }
}

}  // namespace
}  // namespace cmoe

using namespace cmoe;  // NOLINT

extern "C" int cmoe_posterior_mean_optimization(const cmoe_gp* gp, int num_fidelity, const cmoe_gd_params* params,
                                                const double* domain_bounds, const double* initial_guess,
                                                double* best_point, double* best_value, int* found_flag) {
  return guarded(nullptr, [&] {
    const KernelSpec& spec = gp->spec;
    const int dim = spec.dim, ps = dim - num_fidelity;
    CMOE_REQUIRE(num_fidelity >= 0 && ps >= 1, CMOE_ERR_BOUNDS, "num_fidelity out of range");
    CMOE_REQUIRE(params->max_num_steps <= 1 << 20, CMOE_ERR_BOUNDS, "max_num_steps too large");
    for (int d = 0; d < ps; ++d)
      CMOE_REQUIRE(domain_bounds[2 * d] <= domain_bounds[2 * d + 1], CMOE_ERR_BOUNDS, "Tensor product region is EMPTY.");
    require_device(gp->device);
    if (found_flag) *found_flag = 0;
    if (params->max_num_restarts <= 0) return;  // the reference returns without touching its outputs (:424-426)
    cudaStream_t s = gp->stream;
    PmOptParams prm{};
    prm.N = gp->N;
    prm.ps = ps;
    prm.max_steps = params->max_num_steps;
    prm.max_restarts = params->max_num_restarts;
    prm.mean = gp->mean;
    prm.mrc = params->max_relative_change;
    prm.tol = params->tolerance;
    for (int d = 0; d < CMOE_MAX_DIM; ++d) {
      prm.lo[d] = (d < ps) ? domain_bounds[2 * d] : -1e300;
      prm.hi[d] = (d < ps) ? domain_bounds[2 * d + 1] : 1e300;
    }
    std::vector<double> a0(std::max(1, params->max_num_steps)), x0(dim, 1.0), res(ps + 1);
    for (int i = 0; i < params->max_num_steps; ++i)
      a0[i] = params->pre_mult * std::pow(static_cast<double>(i + 1), -params->gamma);
    for (int d = 0; d < ps; ++d) x0[d] = initial_guess[d];
    DevBuf<double> dA0, dX0, dOut(ps + 1);
    dA0.upload(a0.data(), a0.size(), s);
    dX0.upload(x0.data(), x0.size(), s);
    pm_opt_kernel<<<1, 32, 0, s>>>(spec, prm, gp->dX.p, gp->dKinvY.p, dA0.p, dX0.p, dOut.p);
    count_launch();
    CMOE_CUDA(cudaGetLastError());
    dOut.download(res.data(), ps + 1, s);
    CMOE_CUDA(cudaStreamSynchronize(s));
    for (int d = 0; d < ps; ++d) best_point[d] = res[d];
    if (best_value) *best_value = res[ps];
    if (found_flag) *found_flag = 1;
  });
}
