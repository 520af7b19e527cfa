// compute_matches.cu -- placeholder until the file-level twin of computeMatches() lands (next commit)
#include "r3d_internal.cuh"
extern "C" int r3d_compute_matches(r3d_ctx* ctx, const r3d_cm_params*, const r3d_cm_paths*, r3d_progress_cb, void*, r3d_cm_stats*) {
  return r3d::fail(ctx, R3D_ERR_UNSUPPORTED, "r3d_compute_matches: not built yet");
}
