// Bundle_Adjustment_b200.h -- header-only adaptor with the shape of openMVG::sfm::Bundle_Adjustment (sfm_data_BA.hpp):
//     bool Adjust(SfM_Data& sfm_data, const Optimize_Options& options)
// so that the SfM engines Regard3D drives (src/threads/R3DTriangulationThread.cpp:418-441, :492-512, :227-250) can be
// handed the B200 solver instead of Bundle_Adjustment_Ceres.  It copies the container into the library's r3d_sfm_data
// (flatten), calls r3d_sfm_bundle_adjust and copies poses / intrinsics / structure back (unflatten).
//
// The template parameters are the OpenMVG types it is instantiated with (so this header compiles without OpenMVG):
//   SfMData    : views (map id -> shared_ptr<View>), intrinsics (map id -> shared_ptr<IntrinsicBase>),
//                poses (map id -> Pose3), structure (map id -> Landmark{X, obs: map view -> Observation{x, id_feat}})
//   Options    : intrinsics_opt (Intrinsic_Parameter_Type: NONE = 0 -> fixed, anything else -> ADJUST_ALL),
//                use_motion_priors_opt
//   Traits     : small accessor shims, see DefaultTraits below (OpenMVG: rotation() / center() of Pose3, getParams() /
//                updateFromParams() / getType() of IntrinsicBase, ViewPriors members)
// tests/test_cpp_adaptors.py instantiates it against stand-ins of those types.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/r3dgpu.h"

namespace r3d_shim {

template <typename SfMData, typename Options, typename Traits>
class Bundle_Adjustment_b200 {
 public:
  explicit Bundle_Adjustment_b200(int device = 0) : device_(device) {}
  ~Bundle_Adjustment_b200() { if (ctx_) r3d_destroy(ctx_); }
  Bundle_Adjustment_b200(const Bundle_Adjustment_b200&) = delete;
  Bundle_Adjustment_b200& operator=(const Bundle_Adjustment_b200&) = delete;

  const std::string& lastError() const { return err_; }
  const r3d_ba_summary& summary() const { return summary_; }
  r3d_sfm_ba_options& options() { return opt_; }  // solver tolerances, Huber parameter, ... (defaults = the reference's)

  // openMVG::sfm::Bundle_Adjustment::Adjust
  bool Adjust(SfMData& sfm_data, const Options& options) {
    if (!ctx_) {
      r3d_sfm_ba_default_options(&opt_);
      const int ids[1] = {device_};
      if (r3d_create(ids, 1, &ctx_) != R3D_OK) { err_ = r3d_last_error(nullptr); ctx_ = nullptr; return false; }
    }
    r3d_sfm_data* sd = nullptr;
    if (r3d_sfm_data_create(&sd) != R3D_OK) { err_ = "r3d_sfm_data_create failed"; return false; }
    struct Free { r3d_sfm_data* p; ~Free() { r3d_sfm_data_free(p); } } guard{sd};
    // ---- flatten ----
    for (const auto& kv : sfm_data.views) {
      r3d_sfm_view v{};
      Traits::view(*kv.second, &v);
      v.local_path = "";
      v.filename = "";
      if (r3d_sfm_add_view(sd, &v) != R3D_OK) { err_ = "r3d_sfm_add_view failed"; return false; }
    }
    for (const auto& kv : sfm_data.intrinsics) {
      r3d_sfm_intrinsic in{};
      in.id = (uint32_t)kv.first;
      if (!Traits::intrinsic(*kv.second, &in)) { err_ = "unsupported camera model"; return false; }
      if (r3d_sfm_add_intrinsic(sd, &in) != R3D_OK) { err_ = "r3d_sfm_add_intrinsic failed"; return false; }
    }
    for (const auto& kv : sfm_data.poses) {
      r3d_sfm_pose ps{};
      ps.id = (uint32_t)kv.first;
      Traits::pose(kv.second, ps.rotation, ps.center);
      r3d_sfm_add_pose(sd, &ps);
    }
    std::vector<r3d_sfm_observation> obs;
    for (const auto& kv : sfm_data.structure) {
      obs.clear();
      for (const auto& ob : kv.second.obs) {
        r3d_sfm_observation o{};
        o.id_view = (uint32_t)ob.first;
        o.id_feat = (uint32_t)ob.second.id_feat;
        o.x[0] = ob.second.x[0];
        o.x[1] = ob.second.x[1];
        obs.push_back(o);
      }
      const double X[3] = {kv.second.X[0], kv.second.X[1], kv.second.X[2]};
      r3d_sfm_add_landmark(sd, 0, (uint32_t)kv.first, X, obs.data(), (uint32_t)obs.size());
    }
    opt_.solver.refine_intrinsics = Traits::refine_intrinsics(options) ? 1 : 0;
    opt_.use_motion_priors = Traits::use_motion_priors(options) ? 1 : 0;
    const int rc = r3d_sfm_bundle_adjust(ctx_, sd, &opt_, &summary_);
    if (rc != R3D_OK) { err_ = r3d_last_error(ctx_); return false; }
    // ---- unflatten: "Update camera poses / intrinsics / structure with refined data" ----
    for (uint32_t k = 0; k < r3d_sfm_num_poses(sd); ++k) {
      r3d_sfm_pose ps{};
      r3d_sfm_get_pose(sd, k, &ps);
      Traits::set_pose(sfm_data.poses.at(ps.id), ps.rotation, ps.center);
    }
    if (opt_.solver.refine_intrinsics)
      for (uint32_t k = 0; k < r3d_sfm_num_intrinsics(sd); ++k) {
        r3d_sfm_intrinsic in{};
        r3d_sfm_get_intrinsic(sd, k, &in);
        Traits::set_intrinsic(*sfm_data.intrinsics.at(in.id), in);
      }
    for (uint32_t k = 0; k < r3d_sfm_num_landmarks(sd, 0); ++k) {
      uint32_t id = 0;
      double X[3];
      r3d_sfm_get_landmark(sd, 0, k, &id, X, nullptr, 0, nullptr);
      auto& lm = sfm_data.structure.at(id);
      lm.X[0] = X[0]; lm.X[1] = X[1]; lm.X[2] = X[2];
    }
    return summary_.termination != 4;  // Ceres: summary.IsSolutionUsable()
  }

 private:
  int device_;
  r3d_ctx* ctx_ = nullptr;
  r3d_sfm_ba_options opt_{};
  r3d_ba_summary summary_{};
  std::string err_;
};

}  // namespace r3d_shim
