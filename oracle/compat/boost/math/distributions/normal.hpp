// Test-infrastructure shim (oracle/_ref build only). Only pdf/cdf of N(mu, sigma) are used
// (analytic one-point EI). Not product code.
#pragma once
#include <cmath>
namespace boost { namespace math {
template <class T = double>
class normal_distribution {
 public:
  explicit normal_distribution(T mean = 0, T sd = 1) : mean_(mean), sd_(sd) {}
  T mean() const { return mean_; }
  T standard_deviation() const { return sd_; }
 private:
  T mean_, sd_;
};
using normal = normal_distribution<double>;
template <class T>
inline T pdf(const normal_distribution<T>& d, T x) {
  const T z = (x - d.mean()) / d.standard_deviation();
  return std::exp(-0.5 * z * z) / (d.standard_deviation() * 2.5066282746310005024157652848110452530069867406099);
}
template <class T>
inline T cdf(const normal_distribution<T>& d, T x) {
  const T z = (x - d.mean()) / d.standard_deviation();
  return 0.5 * std::erfc(-z * 0.70710678118654752440084436210484903928483593768847);
}
}}  // namespace boost::math
