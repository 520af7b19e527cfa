"""Oracle self-checks for bundle adjustment + the product's analytic Jacobian (host)."""
import numpy as np
import pytest

from regard3d_b200 import synth


def _small(seed=3, **kw):
    args = dict(n_cams=8, n_pts=300, obs_per_pt=4, seed=seed, outlier_frac=0.02)
    args.update(kw)
    return synth.make_ba_problem(**args)


def _prep(oracle, prob):
    return oracle.ba_prepare(prob["poses"], prob["intrinsics"], prob["points"], prob["obs_cam"], prob["obs_pt"],
                             prob["cam_intr"], prob["obs_xy"])


def test_autodiff_jacobian_vs_finite_differences(oracle):
    rng = np.random.default_rng(0)
    for _ in range(20):
        intr = np.array([2000 + rng.normal() * 50, 960, 540, rng.normal() * 0.05, rng.normal() * 0.01, rng.normal() * 0.002])
        pose = np.concatenate([rng.normal(size=3) * 0.5, rng.normal(size=3)])
        X = rng.normal(size=3) + np.array([0, 0, 6.0])
        obs = rng.uniform(0, 1000, 2)
        r, J = oracle.ba_jacobian(intr, pose, X, obs)
        x0 = np.concatenate([intr, pose, X])
        Jn = np.zeros((2, 15))
        for k in range(15):
            h = 1e-6 * max(1, abs(x0[k]))
            xp, xm = x0.copy(), x0.copy()
            xp[k] += h
            xm[k] -= h
            Jn[:, k] = (oracle.ba_jacobian(xp[:6], xp[6:12], xp[12:], obs)[0] - oracle.ba_jacobian(xm[:6], xm[6:12], xm[12:], obs)[0]) / (2 * h)
        assert np.abs(J - Jn).max() < 1e-6 * np.abs(J).max()


def test_product_analytic_jacobian_equals_oracle_autodiff(oracle, r3dlib):
    rng = np.random.default_rng(1)
    worst = 0.0
    for _ in range(500):
        intr = np.array([2000 + rng.normal() * 50, 960 + rng.normal() * 5, 540 + rng.normal() * 5, rng.normal() * 0.05,
                         rng.normal() * 0.01, rng.normal() * 0.002])
        aa = rng.normal(size=3) * rng.choice([1e-12, 1e-5, 0.3, 2.5])     # small-angle branch included
        pose = np.concatenate([aa, rng.normal(size=3)])
        X = rng.normal(size=3) + np.array([0, 0, 6.0])
        obs = rng.uniform(0, 1000, 2)
        r0, J0 = oracle.ba_jacobian(intr, pose, X, obs)
        r1, J1 = r3dlib.debug_ba_jacobian(intr, pose, X, obs)
        worst = max(worst, np.abs(r0 - r1).max() / max(1.0, np.abs(r0).max()), np.abs(J0 - J1).max() / np.abs(J0).max())
    assert worst < 1e-9


def test_oracle_ba_converges_and_cost_is_consistent(oracle):
    prob = _small()
    p = _prep(oracle, prob)
    r0 = oracle.ba_residuals(p)
    summ, trace = oracle.bundle_adjust(p, oracle.default_ba_options(max_iterations=60))
    r1 = oracle.ba_residuals(p)
    assert summ["final_cost"] < 0.02 * summ["initial_cost"]
    assert np.median(r1) < 0.8 and np.median(r0) > 3.0
    assert (np.diff(trace) <= 1e-9 * trace[0]).all()                    # monotone (rejected steps repeat the cost)
    # independent evaluation of the robustified cost at the solution
    s = (r1 ** 2).sum(1)
    rho = np.where(s <= 256.0, s, 2 * 16.0 * np.sqrt(s) - 256.0)
    assert abs(0.5 * rho.sum() - summ["final_cost"]) < 1e-9 * summ["final_cost"]


def test_oracle_ba_matches_scipy_on_tiny_problem(oracle):
    scipy_opt = pytest.importorskip("scipy.optimize")
    prob = synth.make_ba_problem(n_cams=5, n_pts=40, obs_per_pt=4, seed=5, outlier_frac=0.0, noise_px=0.3)
    p = _prep(oracle, prob)
    q = _prep(oracle, prob)
    opts = oracle.default_ba_options(max_iterations=200, huber_a=0.0)
    opts.function_tolerance = 1e-14
    summ, _ = oracle.bundle_adjust(p, opts)
    nc, npt = len(q["poses"]), len(q["points"])

    def fun(x):
        P, I, X = x[:6 * nc].reshape(nc, 6), x[6 * nc:6 * nc + 6], x[6 * nc + 6:].reshape(npt, 3)
        return np.concatenate([oracle.ba_jacobian(I, P[q["obs_cam"][o]], X[q["obs_pt"][o]], q["obs_xy"][o])[0]
                               for o in range(len(q["obs_xy"]))])

    x0 = np.concatenate([q["poses"].ravel(), q["intrinsics"].ravel(), q["points"].ravel()])
    sol = scipy_opt.least_squares(fun, x0, method="lm", xtol=1e-14, ftol=1e-14, gtol=1e-14, max_nfev=4000)
    assert abs(sol.cost - summ["final_cost"]) < 1e-5 * max(1.0, sol.cost)      # same minimum (gauge-free quantity)


def test_fixed_intrinsics_stay_fixed(oracle):
    prob = _small(seed=4)
    p = _prep(oracle, prob)
    i0 = p["intrinsics"].copy()
    oracle.bundle_adjust(p, oracle.default_ba_options(max_iterations=10, refine_intrinsics=0))
    assert np.array_equal(p["intrinsics"], i0)
