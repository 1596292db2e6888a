// Test-infrastructure shim (oracle/_ref build only). Not product code.
#pragma once
#include <sstream>
#include <string>
namespace boost {
template <class Target, class Source>
inline Target lexical_cast(const Source& s) {
  std::ostringstream os;
  os << s;
  return Target(os.str());
}
}  // namespace boost
