// detmath.cuh -- transcendental functions from IEEE-754 basic operations only (+ - * / sqrt, frexp,
// ldexp), usable on host and device.  The a-contrario RANSAC makes discrete decisions (sort order,
// argmin of the NFA, "better model?") on values that pass through acos/cos/cbrt/log10; a libm on the
// CPU and libdevice on the GPU differ in the last bits, and one flipped decision changes every
// later sample.  With these definitions -- and FMA contraction disabled for the translation units
// that use them -- a result is a pure function of its inputs on both processors.
// (Accuracy: a few ulp; the reference's own results are libm- and -ffast-math-dependent,
// src/CMakeLists.txt:578-579.)
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define R3D_HD __host__ __device__ __forceinline__
#else
#define R3D_HD inline
#endif

namespace r3d {
namespace dm {

#define R3D_PI 3.14159265358979323846
#define R3D_LOG10E 0.43429448190325182765
#define R3D_LOG10_2 0.30102999566398119521

R3D_HD double ln_reduced(double m) {  // m in [sqrt(1/2), sqrt(2)): ln m = 2 s (1 + s^2/3 + s^4/5 + ...)
  const double s = (m - 1.0) / (m + 1.0);
  const double s2 = s * s;
  double p = 1.0 / 27.0;
  for (int k = 12; k >= 0; --k) p = p * s2 + 1.0 / (double)(2 * k + 1);
  return 2.0 * s * p;
}

R3D_HD double log10_det(double x) {  // x > 0, finite
  int e;
  double m = frexp(x, &e);
  if (m < 0.70710678118654752440) {
    m = m * 2.0;
    e = e - 1;
  }
  return (double)e * R3D_LOG10_2 + ln_reduced(m) * R3D_LOG10E;
}

R3D_HD double cbrt_det(double a) {  // a >= 0
  if (!(a > 0.0)) return a;
  int e;
  double m = frexp(a, &e);
  int r = e % 3;
  if (r < 0) r += 3;
  if (r != 0) {
    m = ldexp(m, r - 3);
    e = e + (3 - r);
  }
  double y = 0.4285714285714286 + 0.5714285714285714 * m;
  for (int it = 0; it < 7; ++it) y = y - (y * y * y - m) / (3.0 * y * y);
  return ldexp(y, e / 3);
}

R3D_HD double sin_small(double x) {
  const double x2 = x * x;
  double c = 1.0;
  for (int k = 10; k >= 1; --k) c = 1.0 - x2 / (double)((2 * k) * (2 * k + 1)) * c;
  return x * c;
}
R3D_HD double cos_small(double x) {
  const double x2 = x * x;
  double c = 1.0;
  for (int k = 10; k >= 1; --k) c = 1.0 - x2 / (double)((2 * k - 1) * (2 * k)) * c;
  return c;
}

R3D_HD double cos_det(double t) {  // |t| <= ~pi
  double y = t < 0.0 ? -t : t;
  double sign = 1.0;
  if (y > 0.5 * R3D_PI) {
    y = R3D_PI - y;
    sign = -1.0;
  }
  double r;
  if (y > 0.25 * R3D_PI)
    r = sin_small(0.5 * R3D_PI - y);
  else
    r = cos_small(y);
  return sign * r;
}

R3D_HD double atan_pos(double x) {  // x >= 0
  double z = x;
  for (int i = 0; i < 3; ++i) z = z / (1.0 + sqrt(1.0 + z * z));
  const double z2 = z * z;
  double p = 1.0 / 27.0;
  for (int k = 12; k >= 0; --k) p = 1.0 / (double)(2 * k + 1) - z2 * p;
  return 8.0 * (z * p);
}

R3D_HD double acos_det(double u) {
  if (u >= 1.0) return (u == 1.0) ? 0.0 : sqrt(-1.0);
  if (u <= -1.0) return (u == -1.0) ? R3D_PI : sqrt(-1.0);
  return 2.0 * atan_pos(sqrt((1.0 - u) / (1.0 + u)));
}

}  // namespace dm
}  // namespace r3d
