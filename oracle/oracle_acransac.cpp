// oracle_acransac.cpp -- CPU ORACLE (test infrastructure; see oracle.h).
//
// Restates the fundamental-matrix geometric filter that the reference invokes as
//   collectionGeomFilter.Robust_model_estimation(GeometricFilter_FMatrix_AC(4.0, 2048), putatives, false)
// (src/R3DComputeMatches.cpp:2086-2115).  The arithmetic is un-vendored OpenMVG 1.4
// (SURVEY.md Appendix A.4-A.6):
//   matching_image_collection/F_ACRobust.hpp            -> filter_pair_F
//   robust_estimation/robust_estimator_ACRansac.hpp     -> acransac
//   robust_estimation/robust_estimator_ACRansacKernelAdaptator.hpp -> Kernel (normalisation, logalpha0)
//   robust_estimation/rand_sampling.hpp (UniformSample) -> uniform_sample
//   multiview/solver_fundamental_kernel.cpp (SevenPointSolver, SymmetricEpipolarDistanceError)
//   multiview/conditioning.cpp (PreconditionerFromPoints)
//   numeric/poly.h (SolveCubicPolynomial)
// Deliberate, documented deviations (DESIGN.md "oracle fidelity"):
//   * the 2-D nullspace of the 7x9 system is obtained by Gaussian elimination with complete
//     pivoting + Gram-Schmidt instead of Eigen::JacobiSVD (Eigen is not available; the pencil
//     F1 + x F2 and hence the <=3 solutions are the same up to scale and rounding);
//   * acos/cos/pow(.,1/3)/log10 are evaluated by oracle_detmath.hpp instead of libm.
// PARITY UNPINNED (no reference tests / golden vectors).
#include "oracle.h"
#include "oracle_detmath.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <numeric>
#include <random>
#include <utility>
#include <vector>
#include <omp.h>

namespace orc {

// ---- numeric/poly.h : SolveCubicPolynomial (libmv / GSL closed form) ----
static int solve_cubic_monic(double a, double b, double c, double* x0, double* x1, double* x2) {
  const double q = a * a - 3 * b;
  const double r = 2 * a * a * a - 9 * a * b + 27 * c;
  const double Q = q / 9;
  const double R = r / 54;
  const double Q3 = Q * Q * Q;
  const double R2 = R * R;
  const double CR2 = 729 * r * r;
  const double CQ3 = 2916 * q * q * q;
  if (R == 0 && Q == 0) {
    *x0 = *x1 = *x2 = -a / 3;
    return 3;
  } else if (CR2 == CQ3) {
    const double sqrtQ = std::sqrt(Q);
    if (R > 0) {
      *x0 = -2 * sqrtQ - a / 3;
      *x1 = sqrtQ - a / 3;
      *x2 = sqrtQ - a / 3;
    } else {
      *x0 = -sqrtQ - a / 3;
      *x1 = -sqrtQ - a / 3;
      *x2 = 2 * sqrtQ - a / 3;
    }
    return 3;
  } else if (CR2 < CQ3) {
    const double sqrtQ = std::sqrt(Q);
    const double sqrtQ3 = sqrtQ * sqrtQ * sqrtQ;
    const double theta = det::acos(R / sqrtQ3);
    const double norm = -2 * sqrtQ;
    *x0 = norm * det::cos(theta / 3) - a / 3;
    *x1 = norm * det::cos((theta + 2.0 * det::kPi) / 3) - a / 3;
    *x2 = norm * det::cos((theta - 2.0 * det::kPi) / 3) - a / 3;
    if (*x0 > *x1) std::swap(*x0, *x1);
    if (*x1 > *x2) {
      std::swap(*x1, *x2);
      if (*x0 > *x1) std::swap(*x0, *x1);
    }
    return 3;
  }
  const double sgnR = (R >= 0 ? 1 : -1);
  const double A = -sgnR * det::cbrt(std::fabs(R) + std::sqrt(R2 - Q3));
  const double B = Q / A;
  *x0 = A + B - a / 3;
  return 1;
}

static int solve_cubic(const double* coeffs, double* solutions) {
  if (coeffs[0] == 0.0) return 0;  // upstream TODO: quadratic not handled
  const double a = coeffs[2] / coeffs[3];
  const double b = coeffs[1] / coeffs[3];
  const double c = coeffs[0] / coeffs[3];
  return solve_cubic_monic(a, b, c, solutions + 0, solutions + 1, solutions + 2);
}

// 2-D nullspace of the 7x9 epipolar system (replaces Nullspace2/JacobiSVD; see header).
// returns false when the system is rank deficient (< 7).
static bool nullspace_7x9(double A[7][9], double f1[9], double f2[9]) {
  int colperm[9];
  for (int j = 0; j < 9; ++j) colperm[j] = j;
  for (int r = 0; r < 7; ++r) {
    int pi = r, pj = r;
    double best = std::fabs(A[r][r]);
    for (int i = r; i < 7; ++i)
      for (int j = r; j < 9; ++j) {
        const double v = std::fabs(A[i][j]);
        if (v > best) { best = v; pi = i; pj = j; }
      }
    if (!(best > 0.0)) return false;
    if (pi != r)
      for (int j = 0; j < 9; ++j) std::swap(A[r][j], A[pi][j]);
    if (pj != r) {
      for (int i = 0; i < 7; ++i) std::swap(A[i][r], A[i][pj]);
      std::swap(colperm[r], colperm[pj]);
    }
    for (int i = r + 1; i < 7; ++i) {
      const double f = A[i][r] / A[r][r];
      for (int j = r + 1; j < 9; ++j) A[i][j] = A[i][j] - f * A[r][j];
      A[i][r] = 0.0;
    }
  }
  double n[2][9];
  for (int t = 0; t < 2; ++t) {
    double z[9];
    z[7] = (t == 1) ? 1.0 : 0.0;
    z[8] = (t == 0) ? 1.0 : 0.0;
    for (int r = 6; r >= 0; --r) {
      double s = 0.0;
      for (int j = r + 1; j < 9; ++j) s = s + A[r][j] * z[j];
      z[r] = -s / A[r][r];
    }
    for (int k = 0; k < 9; ++k) n[t][colperm[k]] = z[k];
  }
  double nn = 0.0;
  for (int k = 0; k < 9; ++k) nn = nn + n[0][k] * n[0][k];
  nn = std::sqrt(nn);
  for (int k = 0; k < 9; ++k) f1[k] = n[0][k] / nn;
  double dp = 0.0;
  for (int k = 0; k < 9; ++k) dp = dp + n[1][k] * f1[k];
  double g[9];
  for (int k = 0; k < 9; ++k) g[k] = n[1][k] - dp * f1[k];
  double gn = 0.0;
  for (int k = 0; k < 9; ++k) gn = gn + g[k] * g[k];
  gn = std::sqrt(gn);
  for (int k = 0; k < 9; ++k) f2[k] = g[k] / gn;
  return true;
}

// ---- SevenPointSolver::Solve (minimal case, 7 correspondences) ----
// x1, x2: 7x2 (row k = point k), already normalised.  F: up to 3 row-major 3x3.
int seven_point(const double* x1, const double* x2, double* Fout) {
  double A[7][9];
  for (int i = 0; i < 7; ++i) {  // EncodeEpipolarEquation
    const double x1x = x1[2 * i], x1y = x1[2 * i + 1], x2x = x2[2 * i], x2y = x2[2 * i + 1];
    A[i][0] = x2x * x1x;
    A[i][1] = x2x * x1y;
    A[i][2] = x2x;
    A[i][3] = x2y * x1x;
    A[i][4] = x2y * x1y;
    A[i][5] = x2y;
    A[i][6] = x1x;
    A[i][7] = x1y;
    A[i][8] = 1.0;
  }
  double F1[9], F2[9];
  if (!nullspace_7x9(A, F1, F2)) return 0;
  const double a = F1[0], j = F2[0], b = F1[1], k = F2[1], c = F1[2], l = F2[2], d = F1[3], m = F2[3],
               e = F1[4], n = F2[4], f = F1[5], o = F2[5], g = F1[6], p = F2[6], h = F1[7], q = F2[7],
               i = F1[8], r = F2[8];
  const double P[4] = {
      a * e * i + b * f * g + c * d * h - a * f * h - b * d * i - c * e * g,
      a * e * r + a * i * n + b * f * p + b * g * o + c * d * q + c * h * m + d * h * l + e * i * j +
          f * g * k - a * f * q - a * h * o - b * d * r - b * i * m - c * e * p - c * g * n - d * i * k -
          e * g * l - f * h * j,
      a * n * r + b * o * p + c * m * q + d * l * q + e * j * r + f * k * p + g * k * o + h * l * m +
          i * j * n - a * o * q - b * m * r - c * n * p - d * k * r - e * l * p - f * j * q - g * l * n -
          h * j * o - i * k * m,
      j * n * r + k * o * p + l * m * q - j * o * q - k * m * r - l * n * p,
  };
  double roots[3];
  const int num_roots = solve_cubic(P, roots);
  for (int kk = 0; kk < num_roots; ++kk)
    for (int t = 0; t < 9; ++t) Fout[9 * kk + t] = F1[t] + roots[kk] * F2[t];
  return num_roots;
}

// SymmetricEpipolarDistanceError::Error(F, x1, x2)
static inline double sym_epi_error(const double* F, double x1x, double x1y, double x2x, double x2y) {
  const double Fx0 = F[0] * x1x + F[1] * x1y + F[2];
  const double Fx1 = F[3] * x1x + F[4] * x1y + F[5];
  const double Fx2 = F[6] * x1x + F[7] * x1y + F[8];
  const double Fty0 = F[0] * x2x + F[3] * x2y + F[6];
  const double Fty1 = F[1] * x2x + F[4] * x2y + F[7];
  const double yFx = x2x * Fx0 + x2y * Fx1 + Fx2;
  return (yFx * yFx) * (1.0 / (Fx0 * Fx0 + Fx1 * Fx1) + 1.0 / (Fty0 * Fty0 + Fty1 * Fty1)) / 4.0;
}

// ---- homography: FourPointSolver (multiview/solver_homography_kernel.cpp) + AsymmetricError ----
// 1-D nullspace of the 8x9 DLT system (upstream: 16x9 zero-padded, Eigen JacobiSVD; here complete
// pivoting elimination, same null vector up to scale/sign and rounding).
static bool nullspace_8x9(double A[8][9], double h[9]) {
  int colperm[9];
  for (int j = 0; j < 9; ++j) colperm[j] = j;
  for (int r = 0; r < 8; ++r) {
    int pi = r, pj = r;
    double best = std::fabs(A[r][r]);
    for (int i = r; i < 8; ++i)
      for (int j = r; j < 9; ++j) {
        const double v = std::fabs(A[i][j]);
        if (v > best) { best = v; pi = i; pj = j; }
      }
    if (!(best > 0.0)) return false;
    if (pi != r)
      for (int j = 0; j < 9; ++j) std::swap(A[r][j], A[pi][j]);
    if (pj != r) {
      for (int i = 0; i < 8; ++i) std::swap(A[i][r], A[i][pj]);
      std::swap(colperm[r], colperm[pj]);
    }
    for (int i = r + 1; i < 8; ++i) {
      const double f = A[i][r] / A[r][r];
      for (int j = r + 1; j < 9; ++j) A[i][j] = A[i][j] - f * A[r][j];
      A[i][r] = 0.0;
    }
  }
  double z[9];
  z[8] = 1.0;
  for (int r = 7; r >= 0; --r) {
    double s = 0.0;
    for (int j = r + 1; j < 9; ++j) s = s + A[r][j] * z[j];
    z[r] = -s / A[r][r];
  }
  double nn = 0.0;
  for (int k = 0; k < 9; ++k) nn = nn + z[k] * z[k];
  nn = std::sqrt(nn);
  for (int k = 0; k < 9; ++k) h[colperm[k]] = z[k] / nn;
  return true;
}

// x, y: 4x2 normalised points; H row-major, y ~ H x
int four_point(const double* x, const double* y, double* Hout) {
  double L[8][9];
  for (int i = 0; i < 4; ++i) {  // BuildActionMatrix
    const double xx = x[2 * i], xy = x[2 * i + 1], yx = y[2 * i], yy = y[2 * i + 1];
    double* a = L[2 * i];
    double* b = L[2 * i + 1];
    a[0] = xx; a[1] = xy; a[2] = 1.0; a[3] = 0.0; a[4] = 0.0; a[5] = 0.0; a[6] = -yx * xx; a[7] = -yx * xy; a[8] = -yx;
    b[0] = 0.0; b[1] = 0.0; b[2] = 0.0; b[3] = xx; b[4] = xy; b[5] = 1.0; b[6] = -yy * xx; b[7] = -yy * xy; b[8] = -yy;
  }
  return nullspace_8x9(L, Hout) ? 1 : 0;
}

// homography::kernel::AsymmetricError::Error(H, x1, x2)
static inline double asym_error(const double* H, double x1x, double x1y, double x2x, double x2y) {
  const double hx = H[0] * x1x + H[1] * x1y + H[2];
  const double hy = H[3] * x1x + H[4] * x1y + H[5];
  const double hw = H[6] * x1x + H[7] * x1y + H[8];
  const double ex = x2x - hx / hw;
  const double ey = x2y - hy / hw;
  return ex * ex + ey * ey;
}

// ---- essential: FivePointSolver (oracle_fivepoint.cpp) + EpipolarDistanceError on pixel coordinates ----
int five_point(const double* b1, const double* b2, double* Eout);

// bearing vector of a pixel: (K^-1 [x y 1]^T).normalized(), K = [f 0 ppx; 0 f ppy; 0 0 1]
// (Pinhole_Intrinsic::operator(): Kinv * homogeneous, column-normalised)
static inline void bearing(const double* K, double x, double y, double* b) {
  const double kinv00 = 1.0 / K[0], kinv02 = -K[1] / K[0], kinv12 = -K[2] / K[0];
  const double bx = kinv00 * x + kinv02, by = kinv00 * y + kinv12, bz = 1.0;
  const double n = std::sqrt((bx * bx + by * by) + bz * bz);
  b[0] = bx / n; b[1] = by / n; b[2] = bz / n;
}

// FundamentalFromEssential: F = K2^-T E K1^-1
static inline void fundamental_from_essential(const double* E, const double* K1, const double* K2, double* F) {
  const double k1[9] = {1.0 / K1[0], 0.0, -K1[1] / K1[0], 0.0, 1.0 / K1[0], -K1[2] / K1[0], 0.0, 0.0, 1.0};
  const double k2[9] = {1.0 / K2[0], 0.0, -K2[1] / K2[0], 0.0, 1.0 / K2[0], -K2[2] / K2[0], 0.0, 0.0, 1.0};
  double T[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      double a = 0.0;
      for (int k = 0; k < 3; ++k) a = a + k2[3 * k + r] * E[3 * k + c];   // K2^-T E
      T[3 * r + c] = a;
    }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      double a = 0.0;
      for (int k = 0; k < 3; ++k) a = a + T[3 * r + k] * k1[3 * k + c];
      F[3 * r + c] = a;
    }
}

// fundamental::kernel::EpipolarDistanceError::Error(F, x1, x2): squared distance of x2 to the line F x1
static inline double epi_dist_error(const double* F, double x1x, double x1y, double x2x, double x2y) {
  const double Fx0 = F[0] * x1x + F[1] * x1y + F[2];
  const double Fx1 = F[3] * x1x + F[4] * x1y + F[5];
  const double Fx2 = F[6] * x1x + F[7] * x1y + F[8];
  const double yFx = x2x * Fx0 + x2y * Fx1 + Fx2;
  return (yFx * yFx) / (Fx0 * Fx0 + Fx1 * Fx1);
}

// ---- logcombi tables (robust_estimator_ACRansac.hpp) : float, as upstream ----
static float logcombi(uint32_t k, uint32_t n, const std::vector<float>& vec_log10) {
  if (k >= n || k <= 0) return 0.0f;
  if (n - k < k) k = n - k;
  float r = 0.0f;
  for (uint32_t i = 1; i <= k; ++i) r += vec_log10[n - i + 1] - vec_log10[i];
  return r;
}
static void makelogcombi(uint32_t sizeSample, uint32_t n, std::vector<float>& logc_n,
                         std::vector<float>& logc_k) {
  std::vector<float> vec_log10(n + 1);
  for (uint32_t k = 0; k <= n; ++k) vec_log10[k] = std::log10((float)k);
  logc_n.resize(n + 1);
  for (uint32_t k = 0; k <= n; ++k) logc_n[k] = logcombi(k, n, vec_log10);
  logc_k.resize(n + 1);
  for (uint32_t k = 0; k <= n; ++k) logc_k[k] = logcombi(sizeSample, k, vec_log10);
}

// rand_sampling.hpp : UniformSample(num_samples, rng, &vec_index, &sample) (partial Fisher-Yates)
static void uniform_sample(uint32_t num_samples, std::mt19937& rng, std::vector<uint32_t>& vec_index,
                           std::vector<uint32_t>& sample) {
  const uint32_t last_idx = (uint32_t)vec_index.size() - 1;
  for (uint32_t i = 0; i < num_samples; ++i) {
    std::uniform_int_distribution<uint32_t> distribution(i, last_idx);
    const uint32_t sample_idx = distribution(rng);
    std::swap(vec_index[i], vec_index[sample_idx]);
  }
  sample.resize(num_samples);
  for (uint32_t i = 0; i < num_samples; ++i) sample[i] = vec_index[i];
}

// rand_sampling.hpp UniformSample for any sample size
static void uniform_sample_n(uint32_t num_samples, std::mt19937& rng, std::vector<uint32_t>& vec_index,
                             std::vector<uint32_t>& sample) {
  uniform_sample(num_samples, rng, vec_index, sample);
}

// ---- ACRANSAC with the ACKernelAdaptor: MODEL 0 = <SevenPointSolver, SymmetricEpipolarDistanceError>
//      (point-to-line), MODEL 1 = <FourPointSolver, AsymmetricError> (point-to-point),
//      MODEL 2 = ACKernelAdaptorEssential<FivePointSolver, EpipolarDistanceError> (bearing vectors for
//      the fit, pixel residuals, no normalisation; Kpair = f1,ppx1,ppy1,f2,ppx2,ppy2) ----
template <int MODEL>
int64_t acransac(const double* xI, const double* xJ, uint32_t M, uint32_t wI, uint32_t hI, uint32_t wJ, uint32_t hJ,
                 double precision_px, uint32_t max_iter, std::vector<uint32_t>& vec_inliers, double* M_out, double* info,
                 const double* Kpair = nullptr) {
  const uint32_t sizeSample = MODEL == 0 ? 7 : (MODEL == 1 ? 4 : 5);   // Kernel::MINIMUM_SAMPLES
  const uint32_t MAX_MODELS = MODEL == 0 ? 3 : (MODEL == 1 ? 1 : 10);  // Kernel::MAX_MODELS
  vec_inliers.clear();
  if (info) { info[0] = std::numeric_limits<double>::infinity(); info[1] = 0; info[2] = 0; }
  const uint32_t nData = M;
  if (nData <= sizeSample) return 0;

  // kernel adaptor: PreconditionerFromPoints(w,h) ; ApplyTransformationToPoints
  const double s1 = MODEL == 2 ? 1.0 : 1.0 / std::sqrt((double)((int)wI * (int)hI));
  const double s2 = MODEL == 2 ? 1.0 : 1.0 / std::sqrt((double)((int)wJ * (int)hJ));
  const double c1x = MODEL == 2 ? 0.0 : (double)(-.5f * (int)wI) * s1, c1y = MODEL == 2 ? 0.0 : -.5 * (int)hI * s1;
  const double c2x = MODEL == 2 ? 0.0 : (double)(-.5f * (int)wJ) * s2, c2y = MODEL == 2 ? 0.0 : -.5 * (int)hJ * s2;
  std::vector<double> x1k(2 * (size_t)M), x2k(2 * (size_t)M);
  std::vector<double> b1, b2;  // MODEL 2: bearing vectors
  if (MODEL == 2) {
    b1.resize(3 * (size_t)M);
    b2.resize(3 * (size_t)M);
    for (uint32_t i = 0; i < M; ++i) {
      bearing(Kpair, xI[2 * i], xI[2 * i + 1], &b1[3 * (size_t)i]);
      bearing(Kpair + 3, xJ[2 * i], xJ[2 * i + 1], &b2[3 * (size_t)i]);
    }
  }
  for (uint32_t i = 0; i < M; ++i) {
    if (MODEL == 2) {  // the essential adaptor keeps pixel coordinates (normalizer = identity)
      x1k[2 * i] = xI[2 * i]; x1k[2 * i + 1] = xI[2 * i + 1];
      x2k[2 * i] = xJ[2 * i]; x2k[2 * i + 1] = xJ[2 * i + 1];
      continue;
    }
    x1k[2 * i] = s1 * xI[2 * i] + c1x;
    x1k[2 * i + 1] = s1 * xI[2 * i + 1] + c1y;
    x2k[2 * i] = s2 * xJ[2 * i] + c2x;
    x2k[2 * i + 1] = s2 * xJ[2 * i + 1] + c2y;
  }
  double logalpha0, multError;
  if (MODEL == 0) {  // point-to-line: logalpha0 = log10(2 D / A / N2(0,0)), D = diag, A = area of image 2
    const double D = std::sqrt((double)wJ * (double)wJ + (double)hJ * (double)hJ);
    const double Aarea = (double)wJ * (double)hJ;
    logalpha0 = det::log10(2.0 * D / Aarea / s2);
    multError = 0.5;
  } else if (MODEL == 2) {  // essential adaptor: log10(2 D / A * .5), pixel units
    const double D = std::sqrt((double)wJ * (double)wJ + (double)hJ * (double)hJ);
    const double Aarea = (double)wJ * (double)hJ;
    logalpha0 = det::log10(2.0 * D / Aarea * .5);
    multError = 0.5;
  } else {           // point-to-point: logalpha0 = log10(pi / (w h) / N2(0,0)^2)
    logalpha0 = det::log10(det::kPi / ((double)wJ * (double)hJ) / (s2 * s2));
    multError = 1.0;
  }

  const double precision = precision_px * precision_px;  // upper_bound_precision = Square(dPrecision)
  const double maxThreshold = precision * s2 * s2;

  std::vector<uint32_t> vec_index(nData);
  std::iota(vec_index.begin(), vec_index.end(), 0);
  std::vector<uint32_t> vec_sample(sizeSample);
  std::vector<std::pair<double, uint32_t>> sorted(nData);

  const double loge0 = det::log10((double)MAX_MODELS * (double)(nData - sizeSample));
  std::vector<float> logc_n, logc_k;
  makelogcombi(sizeSample, nData, logc_n, logc_k);

  double minNFA = std::numeric_limits<double>::infinity();
  double errorMax = std::numeric_limits<double>::infinity();
  double bestM[9] = {0};

  uint32_t nIterReserve = max_iter / 10;
  uint32_t nIter = max_iter - nIterReserve;
  bool bACRansacMode = (precision == std::numeric_limits<double>::infinity());
  std::mt19937 random_generator(std::mt19937::default_seed);

  uint32_t iter = 0;
  for (iter = 0; iter < nIter; ++iter) {
    uniform_sample_n(sizeSample, random_generator, vec_index, vec_sample);
    double sx1[15], sx2[15], models[90];
    int nmodels;
    if (MODEL == 2) {
      for (uint32_t t = 0; t < sizeSample; ++t)
        for (int c = 0; c < 3; ++c) {
          sx1[3 * t + c] = b1[3 * (size_t)vec_sample[t] + c];
          sx2[3 * t + c] = b2[3 * (size_t)vec_sample[t] + c];
        }
      double Es[90];
      nmodels = five_point(sx1, sx2, Es);
      for (int mi = 0; mi < nmodels; ++mi) fundamental_from_essential(Es + 9 * mi, Kpair, Kpair + 3, models + 9 * mi);
    } else {
      for (uint32_t t = 0; t < sizeSample; ++t) {
        sx1[2 * t] = x1k[2 * vec_sample[t]];
        sx1[2 * t + 1] = x1k[2 * vec_sample[t] + 1];
        sx2[2 * t] = x2k[2 * vec_sample[t]];
        sx2[2 * t + 1] = x2k[2 * vec_sample[t] + 1];
      }
      nmodels = MODEL == 0 ? seven_point(sx1, sx2, models) : four_point(sx1, sx2, models);
    }
    bool better = false;
    for (int mi = 0; mi < nmodels; ++mi) {
      const double* Mm = models + 9 * mi;
      for (uint32_t i = 0; i < nData; ++i) {
        double e = MODEL == 0   ? sym_epi_error(Mm, x1k[2 * i], x1k[2 * i + 1], x2k[2 * i], x2k[2 * i + 1])
                   : MODEL == 1 ? asym_error(Mm, x1k[2 * i], x1k[2 * i + 1], x2k[2 * i], x2k[2 * i + 1])
                                : epi_dist_error(Mm, x1k[2 * i], x1k[2 * i + 1], x2k[2 * i], x2k[2 * i + 1]);
        if (!(e == e)) e = std::numeric_limits<double>::infinity();  // NaN never is an inlier
        sorted[i] = {e, i};
      }
      if (!bACRansacMode) {
        uint32_t nInlier = 0;
        for (uint32_t i = 0; i < nData; ++i)
          if (sorted[i].first <= maxThreshold) ++nInlier;
        if (nInlier > 2.5 * sizeSample) bACRansacMode = true;
      }
      if (bACRansacMode) {
        std::sort(sorted.begin(), sorted.end());
        // bestNFA
        double best_nfa = std::numeric_limits<double>::infinity();
        uint32_t best_k = sizeSample;
        for (uint32_t k = sizeSample + 1; k <= nData && sorted[k - 1].first <= maxThreshold; ++k) {
          const double logalpha =
              logalpha0 + multError * det::log10(sorted[k - 1].first + (double)FLT_EPSILON);
          const double nfa = loge0 + logalpha * (double)(k - sizeSample) + (double)logc_n[k] +
                             (double)logc_k[k];
          if (nfa < best_nfa) { best_nfa = nfa; best_k = k; }
        }
        if (best_nfa < minNFA) {
          better = true;
          minNFA = best_nfa;
          errorMax = sorted[best_k - 1].first;
          vec_inliers.resize(best_k);
          for (uint32_t i = 0; i < best_k; ++i) vec_inliers[i] = sorted[i].second;
          std::memcpy(bestM, Mm, sizeof(bestM));
        }
      }
    }
    if ((better && minNFA < 0) || (iter + 1 == nIter && nIterReserve)) {
      if (vec_inliers.empty()) {
        ++nIter;
        --nIterReserve;
      } else {
        vec_index = vec_inliers;
        if (nIterReserve) {
          nIter = iter + 1 + nIterReserve;
          nIterReserve = 0;
        }
      }
    }
  }
  if (minNFA >= 0) vec_inliers.clear();
  if (info) { info[0] = minNFA; info[2] = (double)iter; }
  if (!vec_inliers.empty()) {
    if (M_out) {
      const double N1[9] = {s1, 0, c1x, 0, s1, c1y, 0, 0, 1};
      const double N2[9] = {s2, 0, c2x, 0, s2, c2y, 0, 0, 1};
      if (MODEL == 2) {  // residual model of the essential adaptor: F = K2^-T E K1^-1 in pixel coordinates
        std::memcpy(M_out, bestM, sizeof(bestM));
      } else if (MODEL == 0) {  // UnnormalizerT: F = N2^T * F * N1
        double T[9];
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += N2[3 * k + r] * bestM[3 * k + c];
            T[3 * r + c] = s;
          }
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += T[3 * r + k] * N1[3 * k + c];
            M_out[3 * r + c] = s;
          }
      } else {           // UnnormalizerI: H = N2^-1 * H * N1
        const double N2i[9] = {1.0 / s2, 0, -c2x / s2, 0, 1.0 / s2, -c2y / s2, 0, 0, 1};
        double T[9];
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += N2i[3 * r + k] * bestM[3 * k + c];
            T[3 * r + c] = s;
          }
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += T[3 * r + k] * N1[3 * k + c];
            M_out[3 * r + c] = s;
          }
      }
    }
    if (info) info[1] = std::sqrt(errorMax) / s2;  // unormalizeError (s2 = 1 for the essential adaptor)
  }
  return (int64_t)vec_inliers.size();
}

int64_t acransac_F(const double* xI, const double* xJ, uint32_t M, uint32_t wI, uint32_t hI,
                   uint32_t wJ, uint32_t hJ, double precision_px, uint32_t max_iter,
                   std::vector<uint32_t>& vec_inliers, double* F_out, double* info) {
  return acransac<0>(xI, xJ, M, wI, hI, wJ, hJ, precision_px, max_iter, vec_inliers, F_out, info);
}
int64_t acransac_H(const double* xI, const double* xJ, uint32_t M, uint32_t wI, uint32_t hI,
                   uint32_t wJ, uint32_t hJ, double precision_px, uint32_t max_iter,
                   std::vector<uint32_t>& vec_inliers, double* H_out, double* info) {
  return acransac<1>(xI, xJ, M, wI, hI, wJ, hJ, precision_px, max_iter, vec_inliers, H_out, info);
}
int64_t acransac_E(const double* xI, const double* xJ, uint32_t M, uint32_t wI, uint32_t hI,
                   uint32_t wJ, uint32_t hJ, const double* Kpair, double precision_px, uint32_t max_iter,
                   std::vector<uint32_t>& vec_inliers, double* F_out, double* info) {
  return acransac<2>(xI, xJ, M, wI, hI, wJ, hJ, precision_px, max_iter, vec_inliers, F_out, info, Kpair);
}

}  // namespace orc

extern "C" {

int orc_seven_point(const double* x1, const double* x2, double* F) {
  return orc::seven_point(x1, x2, F);
}

int64_t orc_acransac_F(const double* xI, const double* xJ, uint32_t M, uint32_t wI, uint32_t hI,
                       uint32_t wJ, uint32_t hJ, double precision_px, uint32_t max_iter,
                       uint32_t* inliers, double* F_out, double* info) {
  std::vector<uint32_t> v;
  orc::acransac_F(xI, xJ, M, wI, hI, wJ, hJ, precision_px, max_iter, v, F_out, info);
  // GeometricFilter_FMatrix_AC::Robust_estimation: keep iff #inliers > MINIMUM_SAMPLES * 2.5
  if (!(v.size() > 7 * 2.5)) v.clear();
  std::memcpy(inliers, v.data(), v.size() * sizeof(uint32_t));
  return (int64_t)v.size();
}

// ImageCollectionGeometricFilter::Robust_model_estimation (omp parallel for schedule(dynamic) over
// the putative map; pairs returning false disappear).  Positions are float in the regions
// (SIOPointFeature) and promoted to double by MatchesPairToMat; intrinsics carry no distortion
// in this path (src/R3DProject.cpp:1177-1180 starts radial-K3 at k=0), so no undistortion.
static int64_t filter_pairs_model(int model, const float* const* xys, const uint32_t* widths, const uint32_t* heights,
                                  uint32_t n_views, const uint32_t* pairs, uint64_t P,
                                  const uint64_t* put_ofs, const orc_indmatch* put, double precision_px,
                                  uint32_t max_iter, uint64_t* out_ofs, orc_indmatch* out, int n_threads,
                                  const double* Ks = nullptr /* n_views x 3: f, ppx, ppy (model 2) */) {
  (void)n_views;
  const double min_samples = model == 0 ? 7.0 : (model == 1 ? 4.0 : 5.0);
  if (n_threads <= 0) n_threads = omp_get_max_threads();
  std::vector<std::vector<orc_indmatch>> res(P);
#pragma omp parallel for schedule(dynamic) num_threads(n_threads)
  for (int64_t p = 0; p < (int64_t)P; ++p) {
    const uint32_t I = pairs[2 * p], J = pairs[2 * p + 1];
    const uint64_t b = put_ofs[p], e = put_ofs[p + 1];
    const uint32_t M = (uint32_t)(e - b);
    if (M == 0) continue;
    std::vector<double> xI(2 * (size_t)M), xJ(2 * (size_t)M);
    for (uint32_t k = 0; k < M; ++k) {
      xI[2 * k] = (double)xys[I][2 * (size_t)put[b + k].i];
      xI[2 * k + 1] = (double)xys[I][2 * (size_t)put[b + k].i + 1];
      xJ[2 * k] = (double)xys[J][2 * (size_t)put[b + k].j];
      xJ[2 * k + 1] = (double)xys[J][2 * (size_t)put[b + k].j + 1];
    }
    std::vector<uint32_t> inl;
    if (model == 2) {
      // GeometricFilter_EMatrix_AC::Robust_estimation: both views need a valid pinhole intrinsic
      if (!(Ks[3 * I] > 0.0) || !(Ks[3 * J] > 0.0)) continue;
      const double Kpair[6] = {Ks[3 * I], Ks[3 * I + 1], Ks[3 * I + 2], Ks[3 * J], Ks[3 * J + 1], Ks[3 * J + 2]};
      orc::acransac_E(xI.data(), xJ.data(), M, widths[I], heights[I], widths[J], heights[J], Kpair,
                      precision_px, max_iter, inl, nullptr, nullptr);
    } else if (model == 0)
      orc::acransac_F(xI.data(), xJ.data(), M, widths[I], heights[I], widths[J], heights[J],
                      precision_px, max_iter, inl, nullptr, nullptr);
    else
      orc::acransac_H(xI.data(), xJ.data(), M, widths[I], heights[I], widths[J], heights[J],
                      precision_px, max_iter, inl, nullptr, nullptr);
    if (inl.size() > min_samples * 2.5) {
      res[p].reserve(inl.size());
      for (uint32_t idx : inl) res[p].push_back(put[b + idx]);
    }
  }
  uint64_t ofs = 0;
  for (uint64_t p = 0; p < P; ++p) {
    out_ofs[p] = ofs;
    std::memcpy(out + ofs, res[p].data(), res[p].size() * sizeof(orc_indmatch));
    ofs += res[p].size();
  }
  out_ofs[P] = ofs;
  return (int64_t)ofs;
}

int64_t orc_filter_pairs_F(const float* const* xys, const uint32_t* widths, const uint32_t* heights,
                           uint32_t n_views, const uint32_t* pairs, uint64_t P,
                           const uint64_t* put_ofs, const orc_indmatch* put, double precision_px,
                           uint32_t max_iter, uint64_t* out_ofs, orc_indmatch* out, int n_threads) {
  return filter_pairs_model(0, xys, widths, heights, n_views, pairs, P, put_ofs, put, precision_px, max_iter, out_ofs, out,
                            n_threads);
}
// GeometricFilter_HMatrix_AC(4.0, 2048) (src/R3DComputeMatches.cpp:2215-2219)
int64_t orc_filter_pairs_H(const float* const* xys, const uint32_t* widths, const uint32_t* heights,
                           uint32_t n_views, const uint32_t* pairs, uint64_t P,
                           const uint64_t* put_ofs, const orc_indmatch* put, double precision_px,
                           uint32_t max_iter, uint64_t* out_ofs, orc_indmatch* out, int n_threads) {
  return filter_pairs_model(1, xys, widths, heights, n_views, pairs, P, put_ofs, put, precision_px, max_iter, out_ofs, out,
                            n_threads);
}

int orc_four_point(const double* x1, const double* x2, double* H) { return orc::four_point(x1, x2, H); }

// GeometricFilter_EMatrix_AC(4.0, 2048) (src/R3DComputeMatches.cpp:2169-2171); Ks = n_views x (f, ppx, ppy),
// f <= 0 marks a view without a valid pinhole intrinsic (its pairs are dropped)
int64_t orc_filter_pairs_E(const float* const* xys, const uint32_t* widths, const uint32_t* heights, const double* Ks,
                           uint32_t n_views, const uint32_t* pairs, uint64_t P,
                           const uint64_t* put_ofs, const orc_indmatch* put, double precision_px,
                           uint32_t max_iter, uint64_t* out_ofs, orc_indmatch* out, int n_threads) {
  return filter_pairs_model(2, xys, widths, heights, n_views, pairs, P, put_ofs, put, precision_px, max_iter, out_ofs, out,
                            n_threads, Ks);
}

int64_t orc_acransac_E(const double* xI, const double* xJ, uint32_t M, uint32_t wI, uint32_t hI,
                       uint32_t wJ, uint32_t hJ, const double* Kpair, double precision_px, uint32_t max_iter,
                       uint32_t* inliers, double* F_out, double* info) {
  std::vector<uint32_t> v;
  orc::acransac_E(xI, xJ, M, wI, hI, wJ, hJ, Kpair, precision_px, max_iter, v, F_out, info);
  if (!(v.size() > 5 * 2.5)) v.clear();
  std::memcpy(inliers, v.data(), v.size() * sizeof(uint32_t));
  return (int64_t)v.size();
}

int64_t orc_acransac_H(const double* xI, const double* xJ, uint32_t M, uint32_t wI, uint32_t hI,
                       uint32_t wJ, uint32_t hJ, double precision_px, uint32_t max_iter,
                       uint32_t* inliers, double* H_out, double* info) {
  std::vector<uint32_t> v;
  orc::acransac_H(xI, xJ, M, wI, hI, wJ, hJ, precision_px, max_iter, v, H_out, info);
  if (!(v.size() > 4 * 2.5)) v.clear();
  std::memcpy(inliers, v.data(), v.size() * sizeof(uint32_t));
  return (int64_t)v.size();
}
}
