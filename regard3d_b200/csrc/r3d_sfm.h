// r3d_sfm.h -- openMVG::sfm::SfM_Data as the library holds it (shared by sfm_data_io.cpp and sfm_ba.cpp).
#pragma once
#include <cmath>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "../../include/r3dgpu.h"

struct r3d_sfm_data {
  std::string root_path;
  struct View {
    std::string local_path, filename;
    uint32_t width = 0, height = 0, id_view = 0, id_intrinsic = 0, id_pose = 0;
    bool priors = false;            // openMVG::sfm::ViewPriors (GPS pose-centre prior)
    bool use_pose_center = false;
    double center_weight[3] = {1.0, 1.0, 1.0}, pose_center[3] = {0.0, 0.0, 0.0};
  };
  struct Intrinsic {
    int model = R3D_CAM_PINHOLE_RADIAL3;
    uint32_t width = 0, height = 0;
    double focal = 0, ppx = 0, ppy = 0;
    std::vector<double> disto;      // K1: 1, K3: 3, Brown T2: 5 (k1 k2 k3 t1 t2), fisheye: 4
  };
  struct Pose { double R[9], C[3]; };
  struct Obs { uint32_t id_feat; double x[2]; };
  struct Landmark { double X[3]; std::map<uint32_t, Obs> obs; };
  std::map<uint32_t, View> views;
  std::map<uint32_t, Intrinsic> intrinsics;
  std::map<uint32_t, Pose> poses;
  std::map<uint32_t, Landmark> structure, control_points;
};


namespace r3d_sfm {

// ceres::RotationMatrixToAngleAxis (via the quaternion, robust near pi)
inline void rotation_to_angle_axis(const double* R, double* aa) {
  double q[4];
  const double tr = R[0] + R[4] + R[8];
  if (tr >= 0.0) {
    double t = std::sqrt(tr + 1.0);
    q[0] = 0.5 * t;
    t = 0.5 / t;
    q[1] = (R[7] - R[5]) * t;
    q[2] = (R[2] - R[6]) * t;
    q[3] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double t = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    q[i + 1] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[3 * k + j] - R[3 * j + k]) * t;
    q[j + 1] = (R[3 * j + i] + R[3 * i + j]) * t;
    q[k + 1] = (R[3 * k + i] + R[3 * i + k]) * t;
  }
  const double s2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (s2 > 0.0) {
    const double s = std::sqrt(s2), c = q[0];
    const double two_theta = 2.0 * (c < 0.0 ? std::atan2(-s, -c) : std::atan2(s, c));
    const double k = two_theta / s;
    aa[0] = q[1] * k; aa[1] = q[2] * k; aa[2] = q[3] * k;
  } else {
    aa[0] = q[1] * 2.0; aa[1] = q[2] * 2.0; aa[2] = q[3] * 2.0;
  }
}

inline void angle_axis_to_rotation(const double* aa, double* R) {  // Rodrigues, row-major
  const double th2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  double A, B;
  if (th2 > 1e-16) {
    const double th = std::sqrt(th2);
    A = std::sin(th) / th;
    B = (1.0 - std::cos(th)) / th2;
  } else {
    A = 1.0 - th2 / 6.0;
    B = 0.5 - th2 / 24.0;
  }
  const double x = aa[0], y = aa[1], z = aa[2];
  const double K[9] = {0, -z, y, z, 0, -x, -y, x, 0};
  const double K2[9] = {x * x - th2, x * y, x * z, x * y, y * y - th2, y * z, x * z, y * z, z * z - th2};
  for (int i = 0; i < 9; ++i) R[i] = ((i == 0 || i == 4 || i == 8) ? 1.0 : 0.0) + A * K[i] + B * K2[i];
}

}  // namespace r3d_sfm

