// oracle/shim/boost/math/distributions/normal.hpp -- TEST INFRASTRUCTURE ONLY.
// [EXT] Boost.Math normal_distribution<RealType> / pdf(): published algorithm restated
// (exponent = -(x-mean)^2 / (2 sd^2); result = exp(exponent) / (sd * sqrt(2 pi)), all in RealType).
#pragma once
#include <cmath>
namespace boost { namespace math {
template <class Real = double>
class normal_distribution {
 public:
  normal_distribution(Real mean = 0, Real sd = 1) : mean_(mean), sd_(sd) {}
  Real mean() const { return mean_; }
  Real standard_deviation() const { return sd_; }
 private:
  Real mean_, sd_;
};
template <class Real>
inline Real pdf(const normal_distribution<Real>& d, const Real& x) {
  const Real sd = d.standard_deviation(), mean = d.mean();
  if (std::isinf(x)) return 0;
  Real exponent = x - mean;
  exponent *= -exponent;
  exponent /= 2 * sd * sd;
  Real result = std::exp(exponent);
  result /= sd * std::sqrt(2 * Real(3.141592653589793238462643383279502884L));
  return result;
}
}}  // namespace boost::math
