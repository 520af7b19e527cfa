// ba_debug.cu -- host-side evaluation of the analytic BA model (diagnostics / CPU-only tests).
#include "ba_model.cuh"
#include "../../include/r3dgpu.h"

extern "C" int r3d_debug_ba_jacobian(const double* intr, const double* pose, const double* X, const double* obs,
                                     double* r, double* J /* 2 x 15: intrinsics, pose, point */) {
  return r3d_debug_ba_jacobian_model(3, intr, nullptr, pose, X, obs, r, J);
}

// residual + prior Jacobian of a pose-centre prior block: r[3], J[3 x 6]
extern "C" int r3d_debug_ba_prior(const double* pose, const double* center, const double* weight, double* r, double* J) {
  r3d::ba::prior_residual_jacobian(pose, center, weight, r, J);
  return 0;
}

extern "C" int r3d_debug_ba_jacobian_model(int model, const double* intr, const double* ext, const double* pose, const double* X,
                                           const double* obs, double* r, double* J /* 2 x 15 */) {
  if (model < 1 || model > 5) return R3D_ERR_INVALID;
  double Ji[12], Jc[12], Jp[6];
  r3d::ba::residual_jacobian(model, intr, ext, pose, X, obs[0], obs[1], r, Ji, Jc, Jp);
  for (int a = 0; a < 2; ++a) {
    for (int k = 0; k < 6; ++k) J[15 * a + k] = Ji[6 * a + k];
    for (int k = 0; k < 6; ++k) J[15 * a + 6 + k] = Jc[6 * a + k];
    for (int k = 0; k < 3; ++k) J[15 * a + 12 + k] = Jp[3 * a + k];
  }
  return 0;
}
