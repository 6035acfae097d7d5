// Stand-in for <boost/math/special_functions/digamma.hpp> (Boost is not installed in this image).
// The reference only calls boost::math::digamma(float) from src/cuda/device.hpp:76-80.  Boost's
// default policy promotes float to double, evaluates, and rounds back; orc_digamma (oracle.c) is
// the textbook recurrence + asymptotic series in double.  TEST INFRASTRUCTURE ONLY.
#ifndef RGBID_ORACLE_BOOST_DIGAMMA_SHIM_HPP_
#define RGBID_ORACLE_BOOST_DIGAMMA_SHIM_HPP_
extern "C" double orc_digamma(double x);
namespace boost { namespace math {
template <class T> inline T digamma(T x) { return static_cast<T>(orc_digamma(static_cast<double>(x))); }
} }
#endif
