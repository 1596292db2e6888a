// Test-infrastructure shim (oracle/_ref build only). Not product code.
// boost::uniform_real::operator() is const (the reference calls it through const&); std's is not.
#pragma once
#include <random>
namespace boost {
template <class T = double>
class uniform_real {
 public:
  uniform_real(T lo, T hi) : lo_(lo), hi_(hi) {}
  template <class Engine>
  T operator()(Engine& eng) const {
    std::uniform_real_distribution<T> dist(lo_, hi_);
    return dist(eng);
  }
  T min() const { return lo_; }
  T max() const { return hi_; }
 private:
  T lo_, hi_;
};
}  // namespace boost
