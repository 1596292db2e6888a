// Test-infrastructure shim (oracle/_ref build only). Not product code.
#pragma once
#include <type_traits>
namespace boost {
template <class Engine, class Distribution>
class variate_generator {
 public:
  using engine_value_type = typename std::remove_reference<Engine>::type;
  using result_type = typename Distribution::result_type;
  variate_generator(engine_value_type& eng, Distribution dist) : eng_(&eng), dist_(dist) {}
  result_type operator()() { return dist_(*eng_); }
  Distribution& distribution() { return dist_; }
  engine_value_type& engine() { return *eng_; }
 private:
  engine_value_type* eng_;
  Distribution dist_;
};
}  // namespace boost
