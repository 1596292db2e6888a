// Test-infrastructure shim (oracle/_ref build only). Not product code.
#pragma once
#include <random>
namespace boost {
template <class T = int>
class uniform_int {
 public:
  uniform_int(T lo, T hi) : lo_(lo), hi_(hi) {}
  template <class Engine>
  T operator()(Engine& eng) const {
    std::uniform_int_distribution<T> dist(lo_, hi_);
    return dist(eng);
  }
 private:
  T lo_, hi_;
};
}  // namespace boost
