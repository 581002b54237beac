// oracle/shim/boost/random/normal_distribution.hpp -- TEST INFRASTRUCTURE ONLY.
// boost::normal_distribution<T>(mean = 0, sigma = 1), operator()(engine). Boost's own sampling algorithm is
// version-dependent (Box-Muller with a cached second variate before 1.56, ziggurat since) and the reference pins
// no Boost version, so the stand-in is libstdc++'s std::normal_distribution (Marsaglia polar, second variate
// cached) -- the generator + distribution pair the reference's four other Monte Carlo programs use
// (src/mc_eur.cpp:16-20). The replay stream of the basket path is therefore `std::mt19937(seed)` +
// `std::normal_distribution<>{0,1}`, drawn d*N times in column-major order of the d x N sample matrix.
#pragma once
#include <random>
namespace boost {
template <class RealType = double>
using normal_distribution = std::normal_distribution<RealType>;
namespace random {
template <class RealType = double>
using normal_distribution = std::normal_distribution<RealType>;
}
}  // namespace boost
