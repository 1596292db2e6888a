// Test-infrastructure shim (oracle/_ref build only): maps the few Boost.Random names the
// reference core uses onto the C++ standard library. Not product code.
#pragma once
#include <random>
namespace boost { using mt19937 = std::mt19937; }
