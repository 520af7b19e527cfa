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

using r3d_sfm::rotation_to_angle_axis;
using r3d_sfm::angle_axis_to_rotation;

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
