// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for Boost.Math's bessel.hpp.
//
// The reference links Boost.Math through the CRAN package BH, version unpinned
// (DESCRIPTION:25 `LinkingTo: Rcpp, BH`; include at src/cocons_types.h:10; call
// sites src/cocons_full.cpp:294,450,573).  BH is absent from this image, so
// cyl_bessel_k is forwarded to libstdc++'s ISO 29124 std::cyl_bessel_k
// (Temme series for x<2, Steed's CF2 above - the same algorithm family as
// Boost's bessel_ik.hpp).  Boost's default policy promotes double to long
// double for the evaluation; the forwarder does the same so the stand-in
// rounds once, like Boost does.
#ifndef COCONS_ORACLE_BOOST_BESSEL_SHIM_HPP
#define COCONS_ORACLE_BOOST_BESSEL_SHIM_HPP

#include <cmath>

namespace boost {
namespace math {

inline double cyl_bessel_k(double v, double x) {
  return static_cast<double>(std::cyl_bessel_k(static_cast<long double>(v), static_cast<long double>(x)));
}

}  // namespace math
}  // namespace boost

#endif
