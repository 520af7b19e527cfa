// sfm_ba.cpp -- openMVG::sfm::Bundle_Adjustment_Ceres::Adjust(SfM_Data&, Optimize_Options) on the library's SfM_Data
// container: flatten -> r3d_bundle_adjust (ba.cu) -> write back.  This is what the SfM engines call between their
// resection / triangulation rounds (src/threads/R3DTriangulationThread.cpp:441, :512, :250 -> engine.Process()); the
// C++ adaptor with the reference's class name is regard3d_b200/csrc/Bundle_Adjustment_b200.h.
//
// Parameterisation as in OpenMVG's BA: pose = angle-axis(R) | t with t = -R C (Pose3 stores R and the
// centre C); one parameter block per intrinsic id (params in getParams() order), one per landmark.  Options mapped:
// intrinsics ADJUST_ALL / NONE (the two the reference uses, R3DTriangulationThread.cpp:429-432), extrinsics and
// structure ADJUST_ALL, control points unused, use_motion_priors -> one pose-centre block per ViewPriors view.  NOT
// restated: the robust similarity registration of the scene to the GPS frame that OpenMVG runs before adding the prior
// blocks (it needs >= 3 priors and changes the gauge, not the reprojection cost); pass prior_huber_a = the fitting error
// of your own registration, or 0 for a quadratic prior.
#include <cmath>
#include <cstring>
#include <map>
#include <new>
#include <vector>

#include "r3d_sfm.h"

namespace {

// ceres::RotationMatrixToAngleAxis (via the quaternion, robust near pi)
void rotation_to_angle_axis(const double* R, double* aa) {
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

void angle_axis_to_rotation(const double* aa, double* R) {  // Rodrigues, row-major
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

}  // namespace

extern "C" void r3d_sfm_ba_default_options(r3d_sfm_ba_options* o) {
  if (!o) return;
  r3d_ba_default_options(&o->solver);
  o->use_motion_priors = 0;
}

extern "C" int r3d_sfm_bundle_adjust(r3d_ctx* ctx, r3d_sfm_data* sd, const r3d_sfm_ba_options* opt, r3d_ba_summary* summary) try {
  if (!ctx || !sd || !opt || !summary) return R3D_ERR_INVALID;
  // ---- flatten -------------------------------------------------------------------------------------------------
  std::map<uint32_t, uint32_t> pose_index, intr_index;
  std::vector<double> poses, intr, ext, points, obs_xy;
  std::vector<uint8_t> intr_model;
  std::vector<uint32_t> obs_cam, obs_pt, cam_intr;
  for (const auto& kv : sd->poses) {
    pose_index[kv.first] = (uint32_t)pose_index.size();
    double aa[3];
    rotation_to_angle_axis(kv.second.R, aa);
    const double* R = kv.second.R;
    const double* C = kv.second.C;
    poses.insert(poses.end(), {aa[0], aa[1], aa[2], -(R[0] * C[0] + R[1] * C[1] + R[2] * C[2]),
                               -(R[3] * C[0] + R[4] * C[1] + R[5] * C[2]), -(R[6] * C[0] + R[7] * C[1] + R[8] * C[2])});
  }
  for (const auto& kv : sd->intrinsics) {
    intr_index[kv.first] = (uint32_t)intr_index.size();
    const r3d_sfm_data::Intrinsic& in = kv.second;
    double p6[6] = {in.focal, in.ppx, in.ppy, 0, 0, 0}, e2[2] = {0, 0};
    for (size_t k = 0; k < in.disto.size(); ++k) {
      if (k < 3) p6[3 + k] = in.disto[k];
      else e2[k - 3] = in.disto[k];
    }
    intr.insert(intr.end(), p6, p6 + 6);
    ext.insert(ext.end(), e2, e2 + 2);
    intr_model.push_back((uint8_t)in.model);
  }
  cam_intr.assign(pose_index.size(), 0xffffffffu);
  std::vector<uint32_t> lm_ids;
  for (const auto& kv : sd->structure) {
    const uint32_t ip = (uint32_t)lm_ids.size();
    lm_ids.push_back(kv.first);
    points.insert(points.end(), kv.second.X, kv.second.X + 3);
    for (const auto& ob : kv.second.obs) {
      auto vit = sd->views.find(ob.first);
      if (vit == sd->views.end()) return R3D_ERR_INVALID;  // an observation of an unknown view (map::at would throw upstream)
      auto pit = pose_index.find(vit->second.id_pose);
      auto iit = intr_index.find(vit->second.id_intrinsic);
      if (pit == pose_index.end() || iit == intr_index.end()) return R3D_ERR_INVALID;
      // the solver keeps one intrinsic group per pose (id_pose = id_view in every sfm_data the reference writes)
      if (cam_intr[pit->second] == 0xffffffffu) cam_intr[pit->second] = iit->second;
      else if (cam_intr[pit->second] != iit->second) return R3D_ERR_UNSUPPORTED;
      obs_cam.push_back(pit->second);
      obs_pt.push_back(ip);
      obs_xy.push_back(ob.second.x[0]);
      obs_xy.push_back(ob.second.x[1]);
    }
  }
  for (uint32_t& g : cam_intr)
    if (g == 0xffffffffu) g = 0;  // a pose nobody observes: any group, it takes no part
  if (intr_index.empty() || pose_index.empty() || lm_ids.empty()) return R3D_ERR_INVALID;
  std::vector<uint32_t> prior_cam;
  std::vector<double> prior_center, prior_weight;
  if (opt->use_motion_priors)
    for (const auto& kv : sd->views) {
      const r3d_sfm_data::View& v = kv.second;
      if (!(v.priors && v.use_pose_center)) continue;
      auto pit = pose_index.find(v.id_pose);
      if (pit == pose_index.end() || intr_index.find(v.id_intrinsic) == intr_index.end()) continue;  // IsPoseAndIntrinsicDefined
      prior_cam.push_back(pit->second);
      prior_center.insert(prior_center.end(), v.pose_center, v.pose_center + 3);
      prior_weight.insert(prior_weight.end(), v.center_weight, v.center_weight + 3);
    }
  r3d_ba_problem p;
  std::memset(&p, 0, sizeof(p));
  p.n_cams = (uint32_t)pose_index.size();
  p.n_pts = (uint32_t)lm_ids.size();
  p.n_intr = (uint32_t)intr_index.size();
  p.n_obs = obs_cam.size();
  p.poses = poses.data(); p.intrinsics = intr.data(); p.points = points.data();
  p.obs_cam = obs_cam.data(); p.obs_pt = obs_pt.data(); p.cam_intr = cam_intr.data(); p.obs_xy = obs_xy.data();
  p.intr_model = intr_model.data();
  p.intrinsics_ext = ext.data();
  p.n_priors = (uint32_t)prior_cam.size();
  p.prior_cam = prior_cam.data(); p.prior_center = prior_center.data(); p.prior_weight = prior_weight.data();
  const int rc = r3d_bundle_adjust(ctx, &p, &opt->solver, summary, nullptr);
  if (rc) return rc;
  // ---- write back (Adjust updates the camera poses, intrinsics and structure with the refined values) ----------------
  for (auto& kv : sd->poses) {
    const double* ps = poses.data() + 6 * (size_t)pose_index[kv.first];
    angle_axis_to_rotation(ps, kv.second.R);
    const double* R = kv.second.R;
    for (int i = 0; i < 3; ++i) kv.second.C[i] = -(R[i] * ps[3] + R[3 + i] * ps[4] + R[6 + i] * ps[5]);  // C = -R^T t
  }
  if (opt->solver.refine_intrinsics)
    for (auto& kv : sd->intrinsics) {
      const double* q = intr.data() + 6 * (size_t)intr_index[kv.first];
      kv.second.focal = q[0]; kv.second.ppx = q[1]; kv.second.ppy = q[2];
      for (size_t k = 0; k < kv.second.disto.size() && k < 3; ++k) kv.second.disto[k] = q[3 + k];
    }
  {
    size_t k = 0;
    for (auto& kv : sd->structure) {
      std::memcpy(kv.second.X, points.data() + 3 * k, 3 * sizeof(double));
      ++k;
    }
  }
  return R3D_OK;
} catch (const std::bad_alloc&) { return R3D_ERR_NOMEM; } catch (...) { return R3D_ERR_INVALID; }
