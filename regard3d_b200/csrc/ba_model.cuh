// ba_model.cuh -- pinhole radial-K3 reprojection residual and its ANALYTIC Jacobian (host + device).
//
// Residual of OpenMVG's ResidualErrorFunctor_Pinhole_Intrinsic_Radial_K3 (the camera model the
// reference selects at src/threads/R3DTriangulationThread.cpp:398 and builds at
// src/R3DProject.cpp:1177-1180):
//   p = R(aa) X + t ;  (xu, yu) = p.xy / p.z ;  r2 = xu^2 + yu^2 ;  c = 1 + k1 r2 + k2 r2^2 + k3 r2^3
//   res = (ppx + f xu c - ox,  ppy + f yu c - oy)
// intrinsics = [f, ppx, ppy, k1, k2, k3] (order: src/utils/OpenMVGHelper.cpp:2691-2702),
// pose = [angle-axis(3), t(3)].  Ceres differentiates this functor automatically; the derivative
// below is the closed form of the same function (d(RX)/d(aa) = -R [X]x Jr(aa), Jr = right Jacobian
// of SO(3)); tests pin it against the oracle's forward-mode autodiff.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define R3D_BA_HD __host__ __device__ __forceinline__
#else
#define R3D_BA_HD inline
#endif

namespace r3d {
namespace ba {

// R = exp([aa]x) (Rodrigues), row-major; Jr = right Jacobian of SO(3) at aa.
R3D_BA_HD void rotation_and_right_jacobian(const double* aa, double* R, double* Jr) {
  const double x = aa[0], y = aa[1], z = aa[2];
  const double th2 = x * x + y * y + z * z;
  double A, B, C;  // sin(th)/th, (1-cos th)/th^2, (th - sin th)/th^3
  if (th2 > 1e-8) {
    const double th = sqrt(th2);
    const double s = sin(th), c = cos(th);
    A = s / th;
    B = (1.0 - c) / th2;
    C = (th - s) / (th2 * th);
  } else {  // series (ceres switches to p + aa x p below DBL_EPSILON; the two agree to O(th2))
    A = 1.0 - th2 / 6.0;
    B = 0.5 - th2 / 24.0;
    C = 1.0 / 6.0 - th2 / 120.0;
  }
  // K = [aa]x ; K2 = K*K = aa aa^T - th2 I
  const double K[9] = {0, -z, y, z, 0, -x, -y, x, 0};
  const double K2[9] = {x * x - th2, x * y, x * z, x * y, y * y - th2, y * z, x * z, y * z, z * z - th2};
  for (int i = 0; i < 9; ++i) {
    const double I = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
    R[i] = I + A * K[i] + B * K2[i];
    Jr[i] = I - B * K[i] + C * K2[i];
  }
}

// r[2]; Ji[2][6] (d/d intrinsics), Jc[2][6] (d/d pose), Jp[2][3] (d/d point), row-major.
R3D_BA_HD void residual_jacobian(const double* intr, const double* pose, const double* X, double ox, double oy,
                                 double* r, double* Ji, double* Jc, double* Jp) {
  double R[9], Jr[9];
  rotation_and_right_jacobian(pose, R, Jr);
  const double RX[3] = {R[0] * X[0] + R[1] * X[1] + R[2] * X[2], R[3] * X[0] + R[4] * X[1] + R[5] * X[2],
                        R[6] * X[0] + R[7] * X[1] + R[8] * X[2]};
  const double px = RX[0] + pose[3], py = RX[1] + pose[4], pz = RX[2] + pose[5];
  const double iz = 1.0 / pz;
  const double xu = px * iz, yu = py * iz;
  const double r2 = xu * xu + yu * yu;
  const double r4 = r2 * r2, r6 = r4 * r2;
  const double f = intr[0], k1 = intr[3], k2 = intr[4], k3 = intr[5];
  const double c = 1.0 + k1 * r2 + k2 * r4 + k3 * r6;
  const double dc = k1 + 2.0 * k2 * r2 + 3.0 * k3 * r4;  // dc / d r2
  r[0] = intr[1] + f * xu * c - ox;
  r[1] = intr[2] + f * yu * c - oy;
  // intrinsics
  Ji[0] = xu * c; Ji[1] = 1.0; Ji[2] = 0.0; Ji[3] = f * xu * r2; Ji[4] = f * xu * r4; Ji[5] = f * xu * r6;
  Ji[6] = yu * c; Ji[7] = 0.0; Ji[8] = 1.0; Ji[9] = f * yu * r2; Ji[10] = f * yu * r4; Ji[11] = f * yu * r6;
  // d res / d (xu, yu)
  const double a00 = f * (c + 2.0 * xu * xu * dc), a01 = f * (2.0 * xu * yu * dc);
  const double a10 = a01, a11 = f * (c + 2.0 * yu * yu * dc);
  // d (xu, yu) / d p = [iz 0 -xu iz ; 0 iz -yu iz]  ->  G = d res / d p (2x3)
  const double G[6] = {a00 * iz, a01 * iz, -(a00 * xu + a01 * yu) * iz, a10 * iz, a11 * iz, -(a10 * xu + a11 * yu) * iz};
  // d p / d t = I ; d p / d X = R ; d p / d aa = -R [X]x Jr = -[RX]x R Jr
  for (int a = 0; a < 2; ++a) {
    Jc[6 * a + 3] = G[3 * a];
    Jc[6 * a + 4] = G[3 * a + 1];
    Jc[6 * a + 5] = G[3 * a + 2];
    for (int j = 0; j < 3; ++j) Jp[3 * a + j] = G[3 * a] * R[j] + G[3 * a + 1] * R[3 + j] + G[3 * a + 2] * R[6 + j];
  }
  // M = R Jr ; dp/daa = -[RX]x M
  double M[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) M[3 * i + j] = R[3 * i] * Jr[j] + R[3 * i + 1] * Jr[3 + j] + R[3 * i + 2] * Jr[6 + j];
  double Dp[9];  // -[RX]x M : row i = -(RX x M_col)...  ([v]x M)_{i,j} = (v x M_{:,j})_i
  for (int j = 0; j < 3; ++j) {
    const double m0 = M[j], m1 = M[3 + j], m2 = M[6 + j];
    Dp[j] = -(RX[1] * m2 - RX[2] * m1);
    Dp[3 + j] = -(RX[2] * m0 - RX[0] * m2);
    Dp[6 + j] = -(RX[0] * m1 - RX[1] * m0);
  }
  for (int a = 0; a < 2; ++a)
    for (int j = 0; j < 3; ++j) Jc[6 * a + j] = G[3 * a] * Dp[j] + G[3 * a + 1] * Dp[3 + j] + G[3 * a + 2] * Dp[6 + j];
}

R3D_BA_HD void residual_only(const double* intr, const double* pose, const double* X, double ox, double oy, double* r) {
  double R[9], Jr[9];
  rotation_and_right_jacobian(pose, R, Jr);
  const double px = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + pose[3];
  const double py = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + pose[4];
  const double pz = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + pose[5];
  const double xu = px / pz, yu = py / pz;
  const double r2 = xu * xu + yu * yu;
  const double c = 1.0 + intr[3] * r2 + intr[4] * r2 * r2 + intr[5] * r2 * r2 * r2;
  r[0] = intr[1] + intr[0] * xu * c - ox;
  r[1] = intr[2] + intr[0] * yu * c - oy;
}

// ceres::HuberLoss(a): rho(s) and rho'(s) with s = ||r||^2 ; a <= 0 -> trivial loss
R3D_BA_HD double huber_rho(double s, double a, double* rho1) {
  if (a <= 0.0) { *rho1 = 1.0; return s; }
  const double b = a * a;
  if (s > b) {
    const double rr = sqrt(s);
    double d = a / rr;
    if (d < 2.2250738585072014e-308) d = 2.2250738585072014e-308;
    *rho1 = d;
    return 2.0 * a * rr - b;
  }
  *rho1 = 1.0;
  return s;
}

}  // namespace ba
}  // namespace r3d
