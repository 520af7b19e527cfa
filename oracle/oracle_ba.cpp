// oracle_ba.cpp -- CPU ORACLE (test infrastructure; see oracle.h).
//
// Restates the bundle adjustment the reference reaches through the OpenMVG SfM engines' Process()
// (src/threads/R3DTriangulationThread.cpp:441, :512, :250) -> Bundle_Adjustment_Ceres::Adjust
// (un-vendored OpenMVG 1.4 + Ceres):
//   * residual: pinhole radial-K3 functor (sfm_data_BA_ceres_camera_functor.hpp), intrinsics
//     [f, ppx, ppy, k1, k2, k3] (order confirmed by the reference's reader,
//     src/utils/OpenMVGHelper.cpp:2691-2702), pose = [angle-axis, t], X_cam = R X + t;
//   * Jacobians by forward-mode automatic differentiation (as Ceres' AutoDiffCostFunction does);
//   * HuberLoss(a) with Ceres' Corrector (rho'' <= 0 branch: residual and Jacobian scaled by sqrt(rho'));
//   * trust-region Levenberg-Marquardt as in Ceres' TrustRegionMinimizer / LevenbergMarquardtStrategy:
//     Jacobi column scaling 1/(1+||col||) from the initial Jacobian, diagonal clamp [1e-6, 1e32],
//     step quality rho > 1e-3, radius /= max(1/3, 1-(2 rho-1)^3) or radius /= nu, nu *= 2,
//     function / gradient / parameter tolerances;
//   * SPARSE_SCHUR: points eliminated, dense reduced camera (+ intrinsics) system, Cholesky.
//     model_cost_change uses the identity 1/2 delta^T (D^2 delta - g) valid for the LM step.
// PARITY UNPINNED (no reference tests / golden vectors; Ceres version unpinned).
#include "oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>
#include <omp.h>

namespace orc {
namespace ba {

// ---- tiny forward-mode jet (N partials) ---------------------------------------------------------
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0) { for (int i = 0; i < N; ++i) v[i] = 0; }
  explicit Jet(double x) : a(x) { for (int i = 0; i < N; ++i) v[i] = 0; }
  Jet(double x, int k) : a(x) { for (int i = 0; i < N; ++i) v[i] = 0; v[k] = 1; }
};
template <int N> Jet<N> operator+(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a + y.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] + y.v[i]; return r; }
template <int N> Jet<N> operator-(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a - y.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] - y.v[i]; return r; }
template <int N> Jet<N> operator-(const Jet<N>& x) { Jet<N> r; r.a = -x.a; for (int i = 0; i < N; ++i) r.v[i] = -x.v[i]; return r; }
template <int N> Jet<N> operator*(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a * y.a; for (int i = 0; i < N; ++i) r.v[i] = x.a * y.v[i] + x.v[i] * y.a; return r; }
template <int N> Jet<N> operator/(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; const double inv = 1.0 / y.a; r.a = x.a * inv; for (int i = 0; i < N; ++i) r.v[i] = (x.v[i] - r.a * y.v[i]) * inv; return r; }
template <int N> Jet<N> operator+(const Jet<N>& x, double s) { Jet<N> r = x; r.a += s; return r; }
template <int N> Jet<N> operator-(const Jet<N>& x, double s) { Jet<N> r = x; r.a -= s; return r; }
template <int N> Jet<N> operator*(const Jet<N>& x, double s) { Jet<N> r; r.a = x.a * s; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * s; return r; }
template <int N> Jet<N> operator*(double s, const Jet<N>& x) { return x * s; }
template <int N> Jet<N> operator+(double s, const Jet<N>& x) { return x + s; }
template <int N> Jet<N> operator-(double s, const Jet<N>& x) { return (-x) + s; }
template <int N> Jet<N> sqrt(const Jet<N>& x) { Jet<N> r; r.a = std::sqrt(x.a); const double d = 0.5 / r.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * d; return r; }
template <int N> Jet<N> sin(const Jet<N>& x) { Jet<N> r; r.a = std::sin(x.a); const double c = std::cos(x.a); for (int i = 0; i < N; ++i) r.v[i] = c * x.v[i]; return r; }
template <int N> Jet<N> cos(const Jet<N>& x) { Jet<N> r; r.a = std::cos(x.a); const double s = -std::sin(x.a); for (int i = 0; i < N; ++i) r.v[i] = s * x.v[i]; return r; }
template <int N> Jet<N> atan(const Jet<N>& x) { Jet<N> r; r.a = std::atan(x.a); const double d = 1.0 / (1.0 + x.a * x.a); for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * d; return r; }
inline double atan(double x) { return std::atan(x); }
inline double sqrt(double x) { return std::sqrt(x); }
inline double sin(double x) { return std::sin(x); }
inline double cos(double x) { return std::cos(x); }
template <int N> double value(const Jet<N>& x) { return x.a; }
inline double value(double x) { return x; }

// ceres::AngleAxisRotatePoint
template <typename T>
void angle_axis_rotate_point(const T aa[3], const T pt[3], T result[3]) {
  const T theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (value(theta2) > std::numeric_limits<double>::epsilon()) {
    const T theta = sqrt(theta2);
    const T costheta = cos(theta);
    const T sintheta = sin(theta);
    const T theta_inverse = T(1.0) / theta;
    const T w[3] = {aa[0] * theta_inverse, aa[1] * theta_inverse, aa[2] * theta_inverse};
    const T w_cross_pt[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2], w[0] * pt[1] - w[1] * pt[0]};
    const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (T(1.0) - costheta);
    result[0] = pt[0] * costheta + w_cross_pt[0] * sintheta + w[0] * tmp;
    result[1] = pt[1] * costheta + w_cross_pt[1] * sintheta + w[1] * tmp;
    result[2] = pt[2] * costheta + w_cross_pt[2] * sintheta + w[2] * tmp;
  } else {
    const T w_cross_pt[3] = {aa[1] * pt[2] - aa[2] * pt[1], aa[2] * pt[0] - aa[0] * pt[2], aa[0] * pt[1] - aa[1] * pt[0]};
    result[0] = pt[0] + w_cross_pt[0];
    result[1] = pt[1] + w_cross_pt[1];
    result[2] = pt[2] + w_cross_pt[2];
  }
}

// ResidualErrorFunctor_Pinhole_Intrinsic_Radial_K3::operator()
// OpenMVG's residual functors of the five camera models Regard3D can store (sfm_data_BA_ceres_camera_functor.hpp;
// model chosen per intrinsic group: src/R3DProject.cpp:1167-1191): 1 pinhole, 2 radial K1, 3 radial K3, 4 Brown T2
// (k1 k2 k3 in intr[3..5], t1 t2 in ext), 5 fisheye (k1 k2 k3 in intr[3..5], k4 in ext).  ext is constant.
template <typename T>
void residual_model(int model, const T intr[6], const double* ext, const T pose[6], const T X[3], const double obs[2], T out[2]) {
  T p[3];
  angle_axis_rotate_point(pose, X, p);
  p[0] = p[0] + pose[3];
  p[1] = p[1] + pose[4];
  p[2] = p[2] + pose[5];
  const T x_u = p[0] / p[2];
  const T y_u = p[1] / p[2];
  const T r2 = x_u * x_u + y_u * y_u;
  T x_d = x_u, y_d = y_u;
  if (model == 2) {
    const T r_coeff = T(1.0) + intr[3] * r2;
    x_d = x_u * r_coeff;
    y_d = y_u * r_coeff;
  } else if (model == 3 || model == 4) {
    const T r4 = r2 * r2;
    const T r6 = r4 * r2;
    const T r_coeff = T(1.0) + intr[3] * r2 + intr[4] * r4 + intr[5] * r6;
    x_d = x_u * r_coeff;
    y_d = y_u * r_coeff;
    if (model == 4) {
      const double t1 = ext ? ext[0] : 0.0, t2 = ext ? ext[1] : 0.0;
      const T t_x = t2 * (r2 + 2.0 * (x_u * x_u)) + 2.0 * t1 * (x_u * y_u);
      const T t_y = t1 * (r2 + 2.0 * (y_u * y_u)) + 2.0 * t2 * (x_u * y_u);
      x_d = x_d + t_x;
      y_d = y_d + t_y;
    }
  } else if (model == 5) {
    const double k4 = ext ? ext[0] : 0.0;
    if (value(r2) > 1e-16) {
      const T r = sqrt(r2);
      const T theta = atan(r);
      const T theta2 = theta * theta, theta3 = theta2 * theta, theta4 = theta2 * theta2, theta5 = theta4 * theta,
              theta7 = theta3 * theta3 * theta, theta8 = theta4 * theta4, theta9 = theta8 * theta;
      const T theta_dist = theta + intr[3] * theta3 + intr[4] * theta5 + intr[5] * theta7 + k4 * theta9;
      const T cdist = theta_dist / r;
      x_d = x_u * cdist;
      y_d = y_u * cdist;
    }
  }
  out[0] = intr[1] + intr[0] * x_d - obs[0];
  out[1] = intr[2] + intr[0] * y_d - obs[1];
}
inline int model_params6(int model) { return model == 1 ? 3 : (model == 2 ? 4 : 6); }

inline void residual_only(int model, const double* intr, const double* ext, const double* pose, const double* X, const double* obs,
                          double* r) {
  residual_model<double>(model, intr, ext, pose, X, obs, r);
}

// residual + Jacobian (2 x 15: intr 0..5, pose 6..11, point 12..14); intrinsic slots the model does not own: zero
inline void residual_jacobian(int model, const double* intr, const double* ext, const double* pose, const double* X,
                              const double* obs, double* r, double J[2][15]) {
  typedef Jet<15> J15;
  J15 ji[6], jp[6], jx[3], out[2];
  for (int k = 0; k < 6; ++k) ji[k] = J15(intr[k], k);
  for (int k = 0; k < 6; ++k) jp[k] = J15(pose[k], 6 + k);
  for (int k = 0; k < 3; ++k) jx[k] = J15(X[k], 12 + k);
  residual_model<J15>(model, ji, ext, jp, jx, obs, out);
  const int np = model_params6(model);
  for (int c = 0; c < 2; ++c) {
    r[c] = out[c].a;
    for (int k = 0; k < 15; ++k) J[c][k] = (k < 6 && k >= np) ? 0.0 : out[c].v[k];
  }
}

// openMVG PoseCenterConstraintCostFunction: residual = weight .* (C - prior), C = -R(aa)^T t
template <typename T>
void prior_residual(const T pose[6], const double* center, const double* weight, T out[3]) {
  const T maa[3] = {-pose[0], -pose[1], -pose[2]};
  const T t[3] = {pose[3], pose[4], pose[5]};
  T c[3];
  angle_axis_rotate_point(maa, t, c);
  for (int i = 0; i < 3; ++i) out[i] = weight[i] * ((-c[i]) - center[i]);
}
inline void prior_residual_jacobian(const double* pose, const double* center, const double* weight, double* r, double J[3][6]) {
  typedef Jet<6> J6;
  J6 jp[6], out[3];
  for (int k = 0; k < 6; ++k) jp[k] = J6(pose[k], k);
  prior_residual<J6>(jp, center, weight, out);
  for (int i = 0; i < 3; ++i) {
    r[i] = out[i].a;
    for (int k = 0; k < 6; ++k) J[i][k] = out[i].v[k];
  }
}

inline double huber_rho(double s, double a, double* rho1) {
  // ceres::HuberLoss(a): b = a^2 ; rho(s) = s (s <= b) ; 2 a sqrt(s) - b otherwise
  if (a <= 0) { *rho1 = 1.0; return s; }
  const double b = a * a;
  if (s > b) {
    const double rr = std::sqrt(s);
    *rho1 = std::max(std::numeric_limits<double>::min(), a / rr);
    return 2.0 * a * rr - b;
  }
  *rho1 = 1.0;
  return s;
}

// dense symmetric positive definite solve (lower Cholesky, blocked, OpenMP); returns false if not PD
static bool cholesky_solve(std::vector<double>& A, int n, std::vector<double>& b, int n_threads) {
  const int NB = 64;
  for (int k0 = 0; k0 < n; k0 += NB) {
    const int kb = std::min(NB, n - k0);
    for (int j = k0; j < k0 + kb; ++j) {  // diagonal block
      double d = A[(size_t)j * n + j];
      for (int t = k0; t < j; ++t) d -= A[(size_t)j * n + t] * A[(size_t)j * n + t];
      if (!(d > 0.0)) return false;
      d = std::sqrt(d);
      A[(size_t)j * n + j] = d;
      for (int i = j + 1; i < k0 + kb; ++i) {
        double s = A[(size_t)i * n + j];
        for (int t = k0; t < j; ++t) s -= A[(size_t)i * n + t] * A[(size_t)j * n + t];
        A[(size_t)i * n + j] = s / d;
      }
    }
    const int r0 = k0 + kb;
#pragma omp parallel for schedule(static) num_threads(n_threads)
    for (int i = r0; i < n; ++i) {  // panel solve
      for (int j = k0; j < k0 + kb; ++j) {
        double s = A[(size_t)i * n + j];
        for (int t = k0; t < j; ++t) s -= A[(size_t)i * n + t] * A[(size_t)j * n + t];
        A[(size_t)i * n + j] = s / A[(size_t)j * n + j];
      }
    }
#pragma omp parallel for schedule(dynamic, 8) num_threads(n_threads)
    for (int i = r0; i < n; ++i) {  // trailing update (lower triangle)
      for (int j = r0; j <= i; ++j) {
        double s = 0;
        const double* ai = &A[(size_t)i * n + k0];
        const double* aj = &A[(size_t)j * n + k0];
        for (int t = 0; t < kb; ++t) s += ai[t] * aj[t];
        A[(size_t)i * n + j] -= s;
      }
    }
  }
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int t = 0; t < i; ++t) s -= A[(size_t)i * n + t] * b[t];
    b[i] = s / A[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int t = i + 1; t < n; ++t) s -= A[(size_t)t * n + i] * b[t];
    b[i] = s / A[(size_t)i * n + i];
  }
  return true;
}

struct Problem {
  orc_ba_problem* p;
  int n_threads;
  bool refine_intr;
  double huber_a;
  // point CSR
  std::vector<uint64_t> pt_ofs;
  std::vector<uint32_t> pt_obs;
  int nB() const { return 6 * (int)p->n_cams + (refine_intr ? 6 * (int)p->n_intr : 0); }
  int intr_col(uint32_t g) const { return 6 * (int)p->n_cams + 6 * (int)g; }
  int model(uint32_t g) const { return p->intr_model ? (int)p->intr_model[g] : 3; }
  const double* ext(uint32_t g) const { return p->intrinsics_ext ? p->intrinsics_ext + 2 * (size_t)g : nullptr; }
  double prior_huber_a = 0.0;
};

static double total_cost(const Problem& P, const double* poses, const double* intr, const double* pts) {
  const orc_ba_problem& p = *P.p;
  double cost = 0;
#pragma omp parallel for reduction(+ : cost) schedule(static) num_threads(P.n_threads)
  for (int64_t o = 0; o < (int64_t)p.n_obs; ++o) {
    const uint32_t c = p.obs_cam[o], pt = p.obs_pt[o];
    double r[2];
    const uint32_t gi = p.cam_intr[c];
    residual_only(P.model(gi), intr + 6 * (size_t)gi, P.ext(gi), poses + 6 * (size_t)c, pts + 3 * (size_t)pt, p.obs_xy + 2 * o, r);
    double rho1;
    cost += 0.5 * huber_rho(r[0] * r[0] + r[1] * r[1], P.huber_a, &rho1);
  }
  for (uint32_t k = 0; k < p.n_priors; ++k) {  // pose-centre priors
    double r[3];
    prior_residual<double>(poses + 6 * (size_t)p.prior_cam[k], p.prior_center + 3 * (size_t)k, p.prior_weight + 3 * (size_t)k, r);
    double rho1;
    cost += 0.5 * huber_rho(r[0] * r[0] + r[1] * r[1] + r[2] * r[2], P.prior_huber_a, &rho1);
  }
  return cost;
}

}  // namespace ba
}  // namespace orc

using namespace orc::ba;

extern "C" {

void orc_ba_residuals(const orc_ba_problem* p, double* res) {
#pragma omp parallel for schedule(static)
  for (int64_t o = 0; o < (int64_t)p->n_obs; ++o) {
    const uint32_t c = p->obs_cam[o], pt = p->obs_pt[o];
    double r[2];
    const uint32_t gi = p->cam_intr[c];
    residual_only(p->intr_model ? (int)p->intr_model[gi] : 3, p->intrinsics + 6 * (size_t)gi,
                  p->intrinsics_ext ? p->intrinsics_ext + 2 * (size_t)gi : nullptr, p->poses + 6 * (size_t)c,
                  p->points + 3 * (size_t)pt, p->obs_xy + 2 * o, r);
    res[2 * o] = std::fabs(r[0]);  // OpenMVGHelper::calculateResiduals: abs per coordinate
    res[2 * o + 1] = std::fabs(r[1]);
  }
}

// residual + Jacobian of one observation (test hook: pins the GPU's analytic derivatives)
void orc_ba_prior(const double* pose, const double* center, const double* weight, double* r, double* J) {
  double Jm[3][6];
  prior_residual_jacobian(pose, center, weight, r, Jm);
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 6; ++k) J[6 * i + k] = Jm[i][k];
}

void orc_ba_jacobian(const double* intr, const double* pose, const double* X, const double* obs, double* r, double* J) {
  orc_ba_jacobian_model(3, intr, nullptr, pose, X, obs, r, J);
}

void orc_ba_jacobian_model(int model, const double* intr, const double* ext, const double* pose, const double* X,
                           const double* obs, double* r, double* J) {
  double Jm[2][15];
  residual_jacobian(model, intr, ext, pose, X, obs, r, Jm);
  for (int c = 0; c < 2; ++c)
    for (int k = 0; k < 15; ++k) J[c * 15 + k] = Jm[c][k];
}

int orc_bundle_adjust(orc_ba_problem* pp, const orc_ba_options* opt, orc_ba_summary* sum, double* cost_trace) {
  const auto t_begin = std::chrono::steady_clock::now();
  Problem P;
  P.p = pp;
  P.n_threads = opt->n_threads > 0 ? opt->n_threads : omp_get_max_threads();
  P.refine_intr = opt->refine_intrinsics != 0;
  P.huber_a = opt->huber_a;
  P.prior_huber_a = opt->prior_huber_a;
  const orc_ba_problem& p = *pp;
  const int nB = P.nB();
  const size_t npt = p.n_pts;
  // point -> observations
  P.pt_ofs.assign(npt + 1, 0);
  for (uint64_t o = 0; o < p.n_obs; ++o) P.pt_ofs[p.obs_pt[o] + 1]++;
  for (size_t i = 0; i < npt; ++i) P.pt_ofs[i + 1] += P.pt_ofs[i];
  P.pt_obs.resize(p.n_obs);
  {
    std::vector<uint64_t> pos(P.pt_ofs.begin(), P.pt_ofs.end() - 1);
    for (uint64_t o = 0; o < p.n_obs; ++o) P.pt_obs[pos[p.obs_pt[o]]++] = (uint32_t)o;
  }
  const size_t nparam = (size_t)nB + 3 * npt;
  std::vector<double> scale(nparam, 1.0), g(nparam), diag(nparam), delta(nparam);
  std::vector<double> S((size_t)nB * nB), rhs(nB);
  std::vector<double> Vinv(9 * npt), gp(3 * npt);
  std::vector<double> poses_new(6 * (size_t)p.n_cams), intr_new(6 * (size_t)p.n_intr), pts_new(3 * npt);

  double cost = total_cost(P, p.poses, p.intrinsics, p.points);
  sum->initial_cost = cost;
  sum->iterations = 0;
  sum->successful_steps = 0;
  sum->termination = 0;
  sum->seconds_linear = 0;
  if (cost_trace) cost_trace[0] = cost;
  double radius = opt->initial_radius, decrease_factor = 2.0;
  bool have_scale = false;

  // per-observation scaled Jacobian blocks, recomputed whenever x changes
  std::vector<double> Jc(12 * p.n_obs), Jg(12 * p.n_obs), Jp(6 * p.n_obs), rr(2 * p.n_obs);
  std::vector<double> Jpr(18 * (size_t)p.n_priors), rpr(3 * (size_t)p.n_priors);  // pose-centre prior blocks (3 x 6)
  auto evaluate = [&]() {
#pragma omp parallel for schedule(static) num_threads(P.n_threads)
    for (int64_t o = 0; o < (int64_t)p.n_obs; ++o) {
      const uint32_t c = p.obs_cam[o], pt = p.obs_pt[o], gi = p.cam_intr[c];
      double r[2], J[2][15];
      residual_jacobian(P.model(gi), p.intrinsics + 6 * (size_t)gi, P.ext(gi), p.poses + 6 * (size_t)c, p.points + 3 * (size_t)pt,
                        p.obs_xy + 2 * o, r, J);
      double rho1;
      huber_rho(r[0] * r[0] + r[1] * r[1], P.huber_a, &rho1);
      const double sq = std::sqrt(rho1);  // Corrector, rho'' <= 0 branch
      for (int a = 0; a < 2; ++a) {
        rr[2 * o + a] = r[a] * sq;
        for (int k = 0; k < 6; ++k) Jg[12 * o + 6 * a + k] = J[a][k] * sq;
        for (int k = 0; k < 6; ++k) Jc[12 * o + 6 * a + k] = J[a][6 + k] * sq;
        for (int k = 0; k < 3; ++k) Jp[6 * o + 3 * a + k] = J[a][12 + k] * sq;
      }
    }
    for (uint32_t k = 0; k < p.n_priors; ++k) {
      double r[3], J[3][6];
      prior_residual_jacobian(p.poses + 6 * (size_t)p.prior_cam[k], p.prior_center + 3 * (size_t)k, p.prior_weight + 3 * (size_t)k, r, J);
      double rho1;
      huber_rho(r[0] * r[0] + r[1] * r[1] + r[2] * r[2], P.prior_huber_a, &rho1);
      const double sq = std::sqrt(rho1);
      for (int a = 0; a < 3; ++a) {
        rpr[3 * (size_t)k + a] = r[a] * sq;
        for (int q = 0; q < 6; ++q) Jpr[18 * (size_t)k + 6 * a + q] = J[a][q] * sq;
      }
    }
    if (!have_scale) {  // Jacobi scaling from the initial Jacobian: 1 / (1 + ||column||)
      std::vector<double> n2(nparam, 0.0);
      for (uint32_t k = 0; k < p.n_priors; ++k)
        for (int a = 0; a < 3; ++a)
          for (int q = 0; q < 6; ++q) n2[6 * (size_t)p.prior_cam[k] + q] += Jpr[18 * (size_t)k + 6 * a + q] * Jpr[18 * (size_t)k + 6 * a + q];
      for (uint64_t o = 0; o < p.n_obs; ++o) {
        const uint32_t c = p.obs_cam[o], pt = p.obs_pt[o], gi = p.cam_intr[c];
        for (int a = 0; a < 2; ++a) {
          for (int k = 0; k < 6; ++k) n2[6 * (size_t)c + k] += Jc[12 * o + 6 * a + k] * Jc[12 * o + 6 * a + k];
          if (P.refine_intr)
            for (int k = 0; k < 6; ++k) n2[P.intr_col(gi) + k] += Jg[12 * o + 6 * a + k] * Jg[12 * o + 6 * a + k];
          for (int k = 0; k < 3; ++k) n2[(size_t)nB + 3 * (size_t)pt + k] += Jp[6 * o + 3 * a + k] * Jp[6 * o + 3 * a + k];
        }
      }
      for (size_t j = 0; j < nparam; ++j) scale[j] = 1.0 / (1.0 + std::sqrt(n2[j]));
      have_scale = true;
    }
#pragma omp parallel for schedule(static) num_threads(P.n_threads)
    for (int64_t o = 0; o < (int64_t)p.n_obs; ++o) {  // apply the column scaling
      const uint32_t c = p.obs_cam[o], pt = p.obs_pt[o], gi = p.cam_intr[c];
      for (int a = 0; a < 2; ++a) {
        for (int k = 0; k < 6; ++k) Jc[12 * o + 6 * a + k] *= scale[6 * (size_t)c + k];
        for (int k = 0; k < 6; ++k) Jg[12 * o + 6 * a + k] *= P.refine_intr ? scale[P.intr_col(gi) + k] : 0.0;
        for (int k = 0; k < 3; ++k) Jp[6 * o + 3 * a + k] *= scale[(size_t)nB + 3 * (size_t)pt + k];
      }
    }
    for (uint32_t k = 0; k < p.n_priors; ++k)
      for (int a = 0; a < 3; ++a)
        for (int q = 0; q < 6; ++q) Jpr[18 * (size_t)k + 6 * a + q] *= scale[6 * (size_t)p.prior_cam[k] + q];
    // gradient g = J^T r and diag(J^T J)
    std::fill(g.begin(), g.end(), 0.0);
    std::fill(diag.begin(), diag.end(), 0.0);
    for (uint32_t k = 0; k < p.n_priors; ++k)
      for (int a = 0; a < 3; ++a)
        for (int q = 0; q < 6; ++q) {
          g[6 * (size_t)p.prior_cam[k] + q] += Jpr[18 * (size_t)k + 6 * a + q] * rpr[3 * (size_t)k + a];
          diag[6 * (size_t)p.prior_cam[k] + q] += Jpr[18 * (size_t)k + 6 * a + q] * Jpr[18 * (size_t)k + 6 * a + q];
        }
    for (uint64_t o = 0; o < p.n_obs; ++o) {
      const uint32_t c = p.obs_cam[o], pt = p.obs_pt[o], gi = p.cam_intr[c];
      for (int a = 0; a < 2; ++a) {
        const double ra = rr[2 * o + a];
        for (int k = 0; k < 6; ++k) {
          g[6 * (size_t)c + k] += Jc[12 * o + 6 * a + k] * ra;
          diag[6 * (size_t)c + k] += Jc[12 * o + 6 * a + k] * Jc[12 * o + 6 * a + k];
        }
        if (P.refine_intr)
          for (int k = 0; k < 6; ++k) {
            g[P.intr_col(gi) + k] += Jg[12 * o + 6 * a + k] * ra;
            diag[P.intr_col(gi) + k] += Jg[12 * o + 6 * a + k] * Jg[12 * o + 6 * a + k];
          }
        for (int k = 0; k < 3; ++k) {
          g[(size_t)nB + 3 * (size_t)pt + k] += Jp[6 * o + 3 * a + k] * ra;
          diag[(size_t)nB + 3 * (size_t)pt + k] += Jp[6 * o + 3 * a + k] * Jp[6 * o + 3 * a + k];
        }
      }
    }
  };
  evaluate();
  auto grad_max = [&]() {
    double m = 0;
    for (size_t j = 0; j < nparam; ++j) m = std::max(m, std::fabs(g[j] / scale[j]));  // unscaled gradient
    return m;
  };
  if (grad_max() <= opt->gradient_tolerance) { sum->termination = 2; sum->final_cost = cost; return 0; }

  for (uint32_t iter = 1; iter <= opt->max_iterations; ++iter) {
    sum->iterations = iter;
    const auto t_lin = std::chrono::steady_clock::now();
    // LevenbergMarquardtStrategy: D^2 = clamp(diag(J^T J), 1e-6, 1e32) / radius
    std::vector<double> D2(nparam);
    for (size_t j = 0; j < nparam; ++j) D2[j] = std::min(std::max(diag[j], 1e-6), 1e32) / radius;
    // ---- Schur complement on the points ----
    std::fill(S.begin(), S.end(), 0.0);
    for (int j = 0; j < nB; ++j) rhs[j] = -g[j];
#pragma omp parallel for schedule(dynamic, 256) num_threads(P.n_threads)
    for (int64_t ip = 0; ip < (int64_t)npt; ++ip) {
      const uint64_t b = P.pt_ofs[ip], e = P.pt_ofs[ip + 1];
      double V[9] = {0};
      for (uint64_t t = b; t < e; ++t) {
        const uint32_t o = P.pt_obs[t];
        for (int a = 0; a < 2; ++a)
          for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) V[3 * i + j] += Jp[6 * (size_t)o + 3 * a + i] * Jp[6 * (size_t)o + 3 * a + j];
      }
      for (int i = 0; i < 3; ++i) V[4 * i] += D2[(size_t)nB + 3 * ip + i];
      // inverse of the symmetric 3x3 (adjugate)
      const double c00 = V[4] * V[8] - V[5] * V[7], c01 = V[5] * V[6] - V[3] * V[8], c02 = V[3] * V[7] - V[4] * V[6];
      const double det = V[0] * c00 + V[1] * c01 + V[2] * c02;
      double Vi[9];
      Vi[0] = c00 / det; Vi[1] = (V[2] * V[7] - V[1] * V[8]) / det; Vi[2] = (V[1] * V[5] - V[2] * V[4]) / det;
      Vi[3] = c01 / det; Vi[4] = (V[0] * V[8] - V[2] * V[6]) / det; Vi[5] = (V[2] * V[3] - V[0] * V[5]) / det;
      Vi[6] = c02 / det; Vi[7] = (V[1] * V[6] - V[0] * V[7]) / det; Vi[8] = (V[0] * V[4] - V[1] * V[3]) / det;
      std::memcpy(&Vinv[9 * ip], Vi, sizeof(Vi));
      const double* gpt = &g[(size_t)nB + 3 * ip];
      const double Vg[3] = {Vi[0] * gpt[0] + Vi[1] * gpt[1] + Vi[2] * gpt[2], Vi[3] * gpt[0] + Vi[4] * gpt[1] + Vi[5] * gpt[2],
                            Vi[6] * gpt[0] + Vi[7] * gpt[1] + Vi[8] * gpt[2]};
      // blocks touched by this point: one per observation (camera) [+ its intrinsic group]
      const int nobs = (int)(e - b);
      std::vector<double> W((size_t)nobs * 2 * 18);  // per obs: Wc (6x3), Wg (6x3)
      std::vector<int> col((size_t)nobs * 2);
      for (int t = 0; t < nobs; ++t) {
        const uint32_t o = P.pt_obs[b + t];
        const uint32_t c = p.obs_cam[o];
        col[2 * t] = 6 * (int)c;
        col[2 * t + 1] = P.refine_intr ? P.intr_col(p.cam_intr[c]) : -1;
        double* Wc = &W[(size_t)t * 36];
        double* Wg = Wc + 18;
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 3; ++j) {
            Wc[3 * i + j] = Jc[12 * (size_t)o + i] * Jp[6 * (size_t)o + j] + Jc[12 * (size_t)o + 6 + i] * Jp[6 * (size_t)o + 3 + j];
            Wg[3 * i + j] = Jg[12 * (size_t)o + i] * Jp[6 * (size_t)o + j] + Jg[12 * (size_t)o + 6 + i] * Jp[6 * (size_t)o + 3 + j];
          }
        // B part: J_B^T J_B of this observation (camera-camera, camera-intrinsic, intrinsic-intrinsic)
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 6; ++j) {
            const double cc = Jc[12 * (size_t)o + i] * Jc[12 * (size_t)o + j] + Jc[12 * (size_t)o + 6 + i] * Jc[12 * (size_t)o + 6 + j];
#pragma omp atomic
            S[(size_t)(col[2 * t] + i) * nB + col[2 * t] + j] += cc;
            if (col[2 * t + 1] >= 0) {
              const double cg = Jc[12 * (size_t)o + i] * Jg[12 * (size_t)o + j] + Jc[12 * (size_t)o + 6 + i] * Jg[12 * (size_t)o + 6 + j];
              const double gg = Jg[12 * (size_t)o + i] * Jg[12 * (size_t)o + j] + Jg[12 * (size_t)o + 6 + i] * Jg[12 * (size_t)o + 6 + j];
#pragma omp atomic
              S[(size_t)(col[2 * t] + i) * nB + col[2 * t + 1] + j] += cg;
#pragma omp atomic
              S[(size_t)(col[2 * t + 1] + j) * nB + col[2 * t] + i] += cg;
#pragma omp atomic
              S[(size_t)(col[2 * t + 1] + i) * nB + col[2 * t + 1] + j] += gg;
            }
          }
      }
      // Schur part: S[a,b] -= W_a Vinv W_b^T ; rhs[a] += W_a Vinv g_p
      const int nblk = 2 * nobs;
      for (int a = 0; a < nblk; ++a) {
        if (col[a] < 0) continue;
        const double* Wa = &W[(size_t)a * 18];
        double WV[18];
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 3; ++j) WV[3 * i + j] = Wa[3 * i] * Vi[j] + Wa[3 * i + 1] * Vi[3 + j] + Wa[3 * i + 2] * Vi[6 + j];
        for (int i = 0; i < 6; ++i) {
          const double v = Wa[3 * i] * Vg[0] + Wa[3 * i + 1] * Vg[1] + Wa[3 * i + 2] * Vg[2];
#pragma omp atomic
          rhs[col[a] + i] += v;
        }
        for (int bb = 0; bb < nblk; ++bb) {
          if (col[bb] < 0) continue;
          const double* Wb = &W[(size_t)bb * 18];
          for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) {
              const double v = WV[3 * i] * Wb[3 * j] + WV[3 * i + 1] * Wb[3 * j + 1] + WV[3 * i + 2] * Wb[3 * j + 2];
#pragma omp atomic
              S[(size_t)(col[a] + i) * nB + col[bb] + j] -= v;
            }
        }
      }
    }
    for (uint32_t k = 0; k < p.n_priors; ++k) {  // U blocks of the pose-centre priors
      const size_t c0 = 6 * (size_t)p.prior_cam[k];
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
          double v = 0;
          for (int a = 0; a < 3; ++a) v += Jpr[18 * (size_t)k + 6 * a + i] * Jpr[18 * (size_t)k + 6 * a + j];
          S[(c0 + i) * nB + c0 + j] += v;
        }
    }
    for (int j = 0; j < nB; ++j) S[(size_t)j * nB + j] += D2[j];
    std::vector<double> dB(rhs);
    const bool pd = cholesky_solve(S, nB, dB, P.n_threads);
    bool step_ok = pd;
    double model_cost_change = 0;
    if (pd) {
      for (int j = 0; j < nB; ++j) delta[j] = dB[j];
      // back substitution: delta_p = Vinv (-g_p - W^T delta_B)
#pragma omp parallel for schedule(static) num_threads(P.n_threads)
      for (int64_t ip = 0; ip < (int64_t)npt; ++ip) {
        double t3[3] = {-g[(size_t)nB + 3 * ip], -g[(size_t)nB + 3 * ip + 1], -g[(size_t)nB + 3 * ip + 2]};
        for (uint64_t t = P.pt_ofs[ip]; t < P.pt_ofs[ip + 1]; ++t) {
          const uint32_t o = P.pt_obs[t];
          const uint32_t c = p.obs_cam[o];
          double m[2] = {0, 0};  // J_B delta_B of this observation
          for (int a = 0; a < 2; ++a) {
            for (int k = 0; k < 6; ++k) m[a] += Jc[12 * (size_t)o + 6 * a + k] * delta[6 * (size_t)c + k];
            if (P.refine_intr)
              for (int k = 0; k < 6; ++k) m[a] += Jg[12 * (size_t)o + 6 * a + k] * delta[P.intr_col(p.cam_intr[c]) + k];
          }
          for (int k = 0; k < 3; ++k) t3[k] -= Jp[6 * (size_t)o + k] * m[0] + Jp[6 * (size_t)o + 3 + k] * m[1];
        }
        const double* Vi = &Vinv[9 * ip];
        for (int i = 0; i < 3; ++i) delta[(size_t)nB + 3 * ip + i] = Vi[3 * i] * t3[0] + Vi[3 * i + 1] * t3[1] + Vi[3 * i + 2] * t3[2];
      }
      double acc = 0;
      for (size_t j = 0; j < nparam; ++j) acc += delta[j] * (D2[j] * delta[j] - g[j]);
      model_cost_change = 0.5 * acc;
      step_ok = model_cost_change > 0.0;
    }
    sum->seconds_linear += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_lin).count();
    bool accepted = false;
    if (step_ok) {
      // parameter tolerance (Ceres checks it on the unscaled step before evaluating it)
      double dn = 0, xn = 0;
      for (uint32_t c = 0; c < p.n_cams; ++c)
        for (int k = 0; k < 6; ++k) {
          const double d = delta[6 * (size_t)c + k] * scale[6 * (size_t)c + k];
          poses_new[6 * (size_t)c + k] = p.poses[6 * (size_t)c + k] + d;
          dn += d * d; xn += p.poses[6 * (size_t)c + k] * p.poses[6 * (size_t)c + k];
        }
      for (uint32_t gi = 0; gi < p.n_intr; ++gi)
        for (int k = 0; k < 6; ++k) {
          const double d = P.refine_intr ? delta[P.intr_col(gi) + k] * scale[P.intr_col(gi) + k] : 0.0;
          intr_new[6 * (size_t)gi + k] = p.intrinsics[6 * (size_t)gi + k] + d;
          if (P.refine_intr) { dn += d * d; xn += p.intrinsics[6 * (size_t)gi + k] * p.intrinsics[6 * (size_t)gi + k]; }
        }
      for (size_t j = 0; j < 3 * npt; ++j) {
        const double d = delta[(size_t)nB + j] * scale[(size_t)nB + j];
        pts_new[j] = p.points[j] + d;
        dn += d * d; xn += p.points[j] * p.points[j];
      }
      if (std::sqrt(dn) <= opt->parameter_tolerance * (std::sqrt(xn) + opt->parameter_tolerance)) {
        sum->termination = 3;
        if (cost_trace) cost_trace[iter] = cost;
        break;
      }
      const double new_cost = total_cost(P, poses_new.data(), intr_new.data(), pts_new.data());
      const double relative_decrease = (cost - new_cost) / model_cost_change;
      if (relative_decrease > 1e-3) {
        accepted = true;
        std::memcpy(pp->poses, poses_new.data(), poses_new.size() * sizeof(double));
        std::memcpy(pp->intrinsics, intr_new.data(), intr_new.size() * sizeof(double));
        std::memcpy(pp->points, pts_new.data(), pts_new.size() * sizeof(double));
        const double cost_change = cost - new_cost;
        const double t = 2.0 * relative_decrease - 1.0;
        radius = radius / std::max(1.0 / 3.0, 1.0 - t * t * t);
        radius = std::min(1e16, radius);
        decrease_factor = 2.0;
        sum->successful_steps++;
        const bool ftol = std::fabs(cost_change) < opt->function_tolerance * cost;
        cost = new_cost;
        if (cost_trace) cost_trace[iter] = cost;
        evaluate();
        if (ftol) { sum->termination = 1; break; }
        if (grad_max() <= opt->gradient_tolerance) { sum->termination = 2; break; }
      }
    }
    if (!accepted) {
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      if (cost_trace) cost_trace[iter] = cost;
      if (radius < 1e-32) { sum->termination = 4; break; }
    }
  }
  sum->final_cost = cost;
  sum->seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
  return 0;
}
}
