"""-m gpu: bundle adjustment through the C ABI vs the CPU oracle.  Bar (north_star): per-observation
reprojection residuals within 1e-5 relative after the same number of LM iterations."""
import numpy as np
import pytest

from regard3d_b200 import synth

pytestmark = pytest.mark.gpu

REL = 1e-5


def _prep(oracle, prob):
    return oracle.ba_prepare(prob["poses"], prob["intrinsics"], prob["points"], prob["obs_cam"], prob["obs_pt"],
                             prob["cam_intr"], prob["obs_xy"])


def _compare(gpu_ctx, oracle, prob, iters, huber_a=16.0, refine=1):
    a = _prep(oracle, prob)
    b = _prep(oracle, prob)
    opts = oracle.default_ba_options(max_iterations=iters, huber_a=huber_a, refine_intrinsics=refine)
    so, to = oracle.bundle_adjust(a, opts)
    sg, tg = gpu_ctx.bundle_adjust(b, max_iterations=iters, huber_a=huber_a, refine_intrinsics=refine)
    assert sg["iterations"] == so["iterations"]
    assert sg["successful_steps"] == so["successful_steps"]
    assert sg["termination"] == so["termination"]
    assert np.allclose(tg, to, rtol=1e-8, atol=0)
    ro = oracle.ba_residuals(a)
    rg = gpu_ctx.ba_residuals(b)
    # residual parity on the parameters each side ended with
    scale = np.maximum(np.abs(ro), 1e-3 * np.median(np.abs(ro)))
    assert (np.abs(rg - ro) / scale).max() < REL
    # and the GPU residual kernel equals the oracle's on identical parameters
    assert np.abs(gpu_ctx.ba_residuals(a) - ro).max() < 1e-9 * max(1.0, ro.max())
    return so, sg


def test_ba_small_equals_oracle(gpu_ctx, oracle):
    prob = synth.make_ba_problem(n_cams=8, n_pts=300, obs_per_pt=4, seed=3, outlier_frac=0.02)
    so, sg = _compare(gpu_ctx, oracle, prob, iters=25)
    assert sg["final_cost"] < 0.05 * sg["initial_cost"]


def test_ba_fixed_intrinsics_and_trivial_loss(gpu_ctx, oracle):
    prob = synth.make_ba_problem(n_cams=6, n_pts=200, obs_per_pt=3, seed=7, outlier_frac=0.0)
    _compare(gpu_ctx, oracle, prob, iters=15, huber_a=0.0, refine=0)
    _compare(gpu_ctx, oracle, prob, iters=15, huber_a=16.0, refine=0)


def test_ba_medium_equals_oracle(gpu_ctx, oracle):
    prob = synth.make_ba_problem(n_cams=40, n_pts=20000, obs_per_pt=5, seed=11)
    _compare(gpu_ctx, oracle, prob, iters=12)


def test_ba_two_intrinsic_groups(gpu_ctx, oracle):
    prob = synth.make_ba_problem(n_cams=10, n_pts=500, obs_per_pt=4, seed=13)
    prob["intrinsics"] = np.repeat(prob["intrinsics"], 2, 0).copy()
    prob["intrinsics"][1, 0] *= 1.01
    prob["cam_intr"] = (np.arange(10) % 2).astype(np.uint32)
    _compare(gpu_ctx, oracle, prob, iters=15)
