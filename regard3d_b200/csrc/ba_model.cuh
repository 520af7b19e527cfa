// ba_model.cuh -- reprojection residuals of the five OpenMVG pinhole camera models and their ANALYTIC Jacobians (host + device).
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

// ---- the five camera models Regard3D can store in sfm_data (src/R3DProject.cpp:1167-1191; openMVG::cameras::
// EINTRINSIC 1..5) and OpenMVG's residual functors for them (sfm_data_BA_ceres_camera_functor.hpp).  A group's
// parameters: intr[6] = f, ppx, ppy, then the model's first three distortion coefficients; ext[2] = coefficients 4, 5:
//   1 pinhole          -                      x_d = x_u
//   2 radial K1        k1                     x_d = x_u (1 + k1 r2)
//   3 radial K3        k1 k2 k3               x_d = x_u (1 + k1 r2 + k2 r4 + k3 r6)
//   4 Brown T2         k1 k2 k3 | t1 t2       x_d = x_u c + t2 (r2 + 2 x_u^2) + 2 t1 x_u y_u ; y_d = y_u c + t1 (r2 + 2 y_u^2) + 2 t2 x_u y_u
//   5 fisheye          k1 k2 k3 | k4          theta = atan r ; x_d = x_u (theta + k1 theta^3 + ... + k4 theta^9) / r
// res = (ppx + f x_d - ox, ppy + f y_d - oy).  The number of parameters of a model that live in intr[]:
R3D_BA_HD int model_params6(int model) { return model == 1 ? 3 : (model == 2 ? 4 : 6); }

// distortion D(x_u, y_u) -> (x_d, y_d); optionally its 2x2 Jacobian A (row-major) and d D / d k for the slots 3..5 of
// intr[] (dk[2][3])
R3D_BA_HD void distort(int model, const double* intr, const double* ext, double xu, double yu, double* xd, double* yd,
                       double* A, double* dk) {
  const double r2 = xu * xu + yu * yu;
  const double k1 = model >= 2 ? intr[3] : 0.0, k2 = model >= 3 ? intr[4] : 0.0, k3 = model >= 3 ? intr[5] : 0.0;
  if (model == 5) {
    const double r = sqrt(r2);
    const double k4 = ext ? ext[0] : 0.0;
    const double th = atan(r);
    const double th2 = th * th, th3 = th2 * th, th4 = th2 * th2, th5 = th4 * th, th7 = th3 * th3 * th, th8 = th4 * th4, th9 = th8 * th;
    const double thd = th + k1 * th3 + k2 * th5 + k3 * th7 + k4 * th9;
    const bool big = r > 1e-8;
    const double inv_r = big ? 1.0 / r : 1.0;
    const double cd = big ? thd * inv_r : 1.0;
    *xd = xu * cd;
    *yd = yu * cd;
    if (A) {
      // d cd / d r = (thd' th' r - thd) / r^2 ; th' = 1 / (1 + r2)
      const double thdp = 1.0 + 3.0 * k1 * th2 + 5.0 * k2 * th4 + 7.0 * k3 * th3 * th3 + 9.0 * k4 * th8;
      const double dcd = big ? (thdp / (1.0 + r2) * r - thd) * inv_r * inv_r : 0.0;
      const double gx = big ? dcd * xu * inv_r : 0.0, gy = big ? dcd * yu * inv_r : 0.0;  // grad cd
      A[0] = cd + xu * gx; A[1] = xu * gy; A[2] = yu * gx; A[3] = cd + yu * gy;
      const double q3 = big ? th3 * inv_r : 0.0, q5 = big ? th5 * inv_r : 0.0, q7 = big ? th7 * inv_r : 0.0;
      dk[0] = xu * q3; dk[1] = xu * q5; dk[2] = xu * q7;
      dk[3] = yu * q3; dk[4] = yu * q5; dk[5] = yu * q7;
    }
    return;
  }
  const double r4 = r2 * r2, r6 = r4 * r2;
  const double c = 1.0 + k1 * r2 + k2 * r4 + k3 * r6;
  const double dc = k1 + 2.0 * k2 * r2 + 3.0 * k3 * r4;  // dc / d r2
  double x = xu * c, y = yu * c;
  double a00 = c + 2.0 * xu * xu * dc, a01 = 2.0 * xu * yu * dc, a10 = a01, a11 = c + 2.0 * yu * yu * dc;
  if (model == 4) {
    const double t1 = ext ? ext[0] : 0.0, t2 = ext ? ext[1] : 0.0;
    x += t2 * (r2 + 2.0 * xu * xu) + 2.0 * t1 * xu * yu;
    y += t1 * (r2 + 2.0 * yu * yu) + 2.0 * t2 * xu * yu;
    a00 += 6.0 * t2 * xu + 2.0 * t1 * yu;
    a01 += 2.0 * t2 * yu + 2.0 * t1 * xu;
    a10 += 2.0 * t1 * xu + 2.0 * t2 * yu;
    a11 += 6.0 * t1 * yu + 2.0 * t2 * xu;
  }
  *xd = x;
  *yd = y;
  if (A) {
    A[0] = a00; A[1] = a01; A[2] = a10; A[3] = a11;
    dk[0] = xu * r2; dk[1] = xu * r4; dk[2] = xu * r6;
    dk[3] = yu * r2; dk[4] = yu * r4; dk[5] = yu * r6;
  }
}

// r[2]; Ji[2][6] (d/d intr[], columns beyond the model's parameters are zero), Jc[2][6] (d/d pose), Jp[2][3] (d/d point)
R3D_BA_HD void residual_jacobian(int model, const double* intr, const double* ext, const double* pose, const double* X,
                                 double ox, double oy, double* r, double* Ji, double* Jc, double* Jp) {
  double R[9], Jr[9];
  rotation_and_right_jacobian(pose, R, Jr);
  const double RX[3] = {R[0] * X[0] + R[1] * X[1] + R[2] * X[2], R[3] * X[0] + R[4] * X[1] + R[5] * X[2],
                        R[6] * X[0] + R[7] * X[1] + R[8] * X[2]};
  const double px = RX[0] + pose[3], py = RX[1] + pose[4], pz = RX[2] + pose[5];
  const double iz = 1.0 / pz;
  const double xu = px * iz, yu = py * iz;
  const double f = intr[0];
  double xd, yd, A[4], dk[6];
  distort(model, intr, ext, xu, yu, &xd, &yd, A, dk);
  r[0] = intr[1] + f * xd - ox;
  r[1] = intr[2] + f * yd - oy;
  // intrinsics
  const int np = model_params6(model);
  Ji[0] = xd; Ji[1] = 1.0; Ji[2] = 0.0;
  Ji[6] = yd; Ji[7] = 0.0; Ji[8] = 1.0;
  for (int k = 0; k < 3; ++k) {
    Ji[3 + k] = (3 + k < np) ? f * dk[k] : 0.0;
    Ji[9 + k] = (3 + k < np) ? f * dk[3 + k] : 0.0;
  }
  // d res / d (xu, yu) = f A
  const double a00 = f * A[0], a01 = f * A[1], a10 = f * A[2], a11 = f * A[3];
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
  double Dp[9];  // -[RX]x M : ([v]x M)_{i,j} = (v x M_{:,j})_i
  for (int j = 0; j < 3; ++j) {
    const double m0 = M[j], m1 = M[3 + j], m2 = M[6 + j];
    Dp[j] = -(RX[1] * m2 - RX[2] * m1);
    Dp[3 + j] = -(RX[2] * m0 - RX[0] * m2);
    Dp[6 + j] = -(RX[0] * m1 - RX[1] * m0);
  }
  for (int a = 0; a < 2; ++a)
    for (int j = 0; j < 3; ++j) Jc[6 * a + j] = G[3 * a] * Dp[j] + G[3 * a + 1] * Dp[3 + j] + G[3 * a + 2] * Dp[6 + j];
}

R3D_BA_HD void residual_only(int model, const double* intr, const double* ext, const double* pose, const double* X, double ox,
                             double oy, double* r) {
  double R[9], Jr[9];
  rotation_and_right_jacobian(pose, R, Jr);
  const double px = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + pose[3];
  const double py = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + pose[4];
  const double pz = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + pose[5];
  const double xu = px / pz, yu = py / pz;
  double xd, yd;
  distort(model, intr, ext, xu, yu, &xd, &yd, nullptr, nullptr);
  r[0] = intr[1] + intr[0] * xd - ox;
  r[1] = intr[2] + intr[0] * yd - oy;
}

// openMVG PoseCenterConstraintCostFunction (sfm_data_BA_ceres.cpp; ViewPriors / GPS, src/R3DProject.cpp:1194-1220):
// residual = weight .* (C(pose) - prior), C = -R^T t.  Jc: 3 x 6 (d/d angle-axis, d/d t), row-major.
R3D_BA_HD void prior_residual_jacobian(const double* pose, const double* center, const double* weight, double* r, double* Jc) {
  const double maa[3] = {-pose[0], -pose[1], -pose[2]};
  double Rt[9], Jrm[9];
  rotation_and_right_jacobian(maa, Rt, Jrm);  // Rt = R(-aa) = R^T
  const double* t = pose + 3;
  double C[3];
  for (int i = 0; i < 3; ++i) C[i] = -(Rt[3 * i] * t[0] + Rt[3 * i + 1] * t[1] + Rt[3 * i + 2] * t[2]);
  for (int i = 0; i < 3; ++i) r[i] = weight[i] * (C[i] - center[i]);
  if (!Jc) return;
  // d(R(w) t)/dw = -R(w) [t]x Jr(w) at w = -aa, d w / d aa = -I  ->  dC/daa = -(R^T [t]x Jr(-aa)) ; dC/dt = -R^T
  const double tx[9] = {0, -t[2], t[1], t[2], 0, -t[0], -t[1], t[0], 0};
  double T1[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T1[3 * i + j] = tx[3 * i] * Jrm[j] + tx[3 * i + 1] * Jrm[3 + j] + tx[3 * i + 2] * Jrm[6 + j];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      const double v = Rt[3 * i] * T1[j] + Rt[3 * i + 1] * T1[3 + j] + Rt[3 * i + 2] * T1[6 + j];
      Jc[6 * i + j] = -weight[i] * v;
      Jc[6 * i + 3 + j] = -weight[i] * Rt[3 * i + j];
    }
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
