// oracle_detmath.hpp -- CPU ORACLE (test infrastructure; see oracle.h).
//
// Portable transcendental functions built ONLY from IEEE-754 basic operations (+ - * / sqrt,
// frexp/ldexp), so that a result does not depend on which libm the build links.  The reference's
// arithmetic (OpenMVG's SolveCubicPolynomial uses acos/cos/pow; ACRANSAC uses log10) is
// libm-dependent at the ulp level and is additionally compiled with -ffast-math
// (src/CMakeLists.txt:578-579), so no libm is "the" reference; this file pins one evaluation.
// Accuracy: a few ulp.  Compile with -ffp-contract=off.
#pragma once
#include <cmath>

namespace orc {
namespace det {

static const double kPi = 3.14159265358979323846;
static const double kLn2 = 0.69314718055994530942;
static const double kLog10e = 0.43429448190325182765;
static const double kLog10_2 = 0.30102999566398119521;

// natural log of m in [sqrt(1/2), sqrt(2)) by the atanh series: ln m = 2 s (1 + s^2/3 + s^4/5 ...)
inline double ln_reduced(double m) {
  const double s = (m - 1.0) / (m + 1.0);
  const double s2 = s * s;
  double p = 1.0 / 27.0;
  for (int k = 12; k >= 0; --k) p = p * s2 + 1.0 / (double)(2 * k + 1);
  return 2.0 * s * p;
}

inline double log10(double x) {  // x > 0, finite
  int e;
  double m = std::frexp(x, &e);  // x = m 2^e, m in [0.5, 1)
  if (m < 0.70710678118654752440) {
    m = m * 2.0;
    e = e - 1;
  }
  return (double)e * kLog10_2 + ln_reduced(m) * kLog10e;
}

inline double cbrt(double a) {  // a >= 0
  if (!(a > 0.0)) return a;     // 0 or NaN
  int e;
  double m = std::frexp(a, &e);  // m in [0.5,1)
  int r = e % 3;
  if (r < 0) r += 3;
  if (r != 0) {  // make the exponent a multiple of 3; m in [0.125, 1)
    m = std::ldexp(m, r - 3);
    e = e + (3 - r);
  }
  double y = 0.4285714285714286 + 0.5714285714285714 * m;  // chord of cbrt on [0.125, 1]
  for (int it = 0; it < 7; ++it) y = y - (y * y * y - m) / (3.0 * y * y);
  return std::ldexp(y, e / 3);
}

inline double sin_small(double x) {  // |x| <= pi/4
  const double x2 = x * x;
  double c = 1.0;
  for (int k = 10; k >= 1; --k) c = 1.0 - x2 / (double)((2 * k) * (2 * k + 1)) * c;
  return x * c;
}
inline double cos_small(double x) {  // |x| <= pi/4
  const double x2 = x * x;
  double c = 1.0;
  for (int k = 10; k >= 1; --k) c = 1.0 - x2 / (double)((2 * k - 1) * (2 * k)) * c;
  return c;
}

inline double cos(double t) {  // |t| <= ~pi (the cubic solver's range); NaN propagates
  double y = t < 0.0 ? -t : t;
  double sign = 1.0;
  if (y > 0.5 * kPi) {
    y = kPi - y;
    sign = -1.0;
  }
  double r;
  if (y > 0.25 * kPi)
    r = sin_small(0.5 * kPi - y);
  else
    r = cos_small(y);
  return sign * r;
}

inline double atan_pos(double x) {  // x >= 0
  // three argument halvings: atan x = 2 atan( x / (1 + sqrt(1 + x^2)) )
  double z = x;
  for (int i = 0; i < 3; ++i) z = z / (1.0 + std::sqrt(1.0 + z * z));
  // z <= tan(pi/16) ~ 0.1989 ; Taylor to z^27
  const double z2 = z * z;
  double p = 1.0 / 27.0;
  for (int k = 12; k >= 0; --k) p = 1.0 / (double)(2 * k + 1) - z2 * p;
  return 8.0 * (z * p);
}

inline double acos(double u) {  // u in [-1,1]; outside -> NaN (like libm)
  if (u >= 1.0) return (u == 1.0) ? 0.0 : std::sqrt(-1.0);
  if (u <= -1.0) return (u == -1.0) ? kPi : std::sqrt(-1.0);
  return 2.0 * atan_pos(std::sqrt((1.0 - u) / (1.0 + u)));
}

}  // namespace det
}  // namespace orc
