// oracle/shim/boost/random/mersenne_twister.hpp -- TEST INFRASTRUCTURE ONLY.
// boost::mt19937 is the 32-bit Mersenne Twister MT19937 of Matsumoto & Nishimura with the standard parameters,
// i.e. the same generator as std::mt19937 (identical output for an identical 32-bit seed). include/mvn.h:21,30,47
// uses: default construction, seed(integer), and being passed to a distribution.
#pragma once
#include <random>
namespace boost {
typedef std::mt19937 mt19937;
namespace random {
typedef std::mt19937 mt19937;
}
}  // namespace boost
