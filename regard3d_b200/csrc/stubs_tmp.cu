// temporary stubs (replaced as the rows of SURVEY.md 8a land)
#include "r3d_internal.cuh"
extern "C" {
void r3d_ba_default_options(r3d_ba_options* o) { if (!o) return; o->max_iterations = 500; o->huber_a = 16.0; o->refine_intrinsics = 1; o->function_tolerance = 1e-6; o->gradient_tolerance = 1e-10; o->parameter_tolerance = 1e-8; o->initial_radius = 1e4; }
int r3d_bundle_adjust(r3d_ctx* ctx, r3d_ba_problem*, const r3d_ba_options*, r3d_ba_summary*, double*) { return r3d::fail(ctx, R3D_ERR_UNSUPPORTED, "r3d_bundle_adjust: not built yet"); }
int r3d_ba_residuals(r3d_ctx* ctx, const r3d_ba_problem*, double*) { return r3d::fail(ctx, R3D_ERR_UNSUPPORTED, "r3d_ba_residuals: not built yet"); }
int r3d_compute_matches(r3d_ctx* ctx, const r3d_cm_params*, const r3d_cm_paths*, r3d_progress_cb, void*, r3d_cm_stats*) { return r3d::fail(ctx, R3D_ERR_UNSUPPORTED, "r3d_compute_matches: not built yet"); }
}
