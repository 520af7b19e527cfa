// r3d_sfm.h -- openMVG::sfm::SfM_Data as the library holds it (shared by sfm_data_io.cpp and sfm_ba.cpp).
#pragma once
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

