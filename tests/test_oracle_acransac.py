"""Oracle self-checks for the AC-RANSAC fundamental filter."""
import math

import numpy as np
import pytest

from regard3d_b200 import synth


def _two_view(seed, n=600, outlier_frac=0.3, noise=0.5):
    rng = np.random.default_rng(seed)
    sc = synth.make_scene(2, 1, 8, "msurf", seed=seed, n_points=n)  # cameras only
    X = rng.uniform([2, 2, 0], [8, 8, 4], size=(n, 3))
    poses = sc["poses"]
    xs = []
    for c in range(2):
        R = synth._rodrigues(poses[c, :3])
        uv, _ = synth.project(R, poses[c, 3:], X, sc["f"], 1920, 1080)
        xs.append(uv + noise * rng.standard_normal(uv.shape))
    n_out = int(outlier_frac * n)
    xs[1][:n_out] = rng.uniform([0, 0], [1920, 1080], size=(n_out, 2))
    good = np.ones(n, bool)
    good[:n_out] = False
    return synth.round_sig(xs[0]).astype(np.float32).astype(np.float64), \
        synth.round_sig(xs[1]).astype(np.float32).astype(np.float64), good


def test_seven_point_nullspace_and_rank(oracle):
    rng = np.random.default_rng(0)
    for _ in range(20):
        x1 = rng.uniform(-0.5, 0.5, (7, 2))
        x2 = rng.uniform(-0.5, 0.5, (7, 2))
        Fs = oracle.seven_point(x1, x2)
        assert 1 <= len(Fs) <= 3
        for F in Fs:
            r = [np.array([*x2[i], 1]) @ F @ np.array([*x1[i], 1]) for i in range(7)]
            assert np.abs(r).max() < 1e-10 * max(1.0, np.abs(F).max())
            assert abs(np.linalg.det(F)) < 1e-10 * max(1.0, np.abs(F).max() ** 3)
        # the pencil agrees with numpy's SVD nullspace: every solution lies in span(V[-1], V[-2])
        A = np.array([[x2[i, 0] * x1[i, 0], x2[i, 0] * x1[i, 1], x2[i, 0], x2[i, 1] * x1[i, 0], x2[i, 1] * x1[i, 1],
                       x2[i, 1], x1[i, 0], x1[i, 1], 1.0] for i in range(7)])
        V = np.linalg.svd(A)[2][-2:]
        for F in Fs:
            f = F.ravel()
            resid = f - V.T @ (V @ f)
            assert np.linalg.norm(resid) < 1e-9 * np.linalg.norm(f)


def test_acransac_recovers_inliers(oracle):
    xI, xJ, good = _two_view(3)
    inl, F, info = oracle.acransac_F(xI, xJ, 1920, 1080, 1920, 1080)
    assert info[0] < 0                      # meaningful (NFA < 1)
    assert good[inl].mean() > 0.97          # precision
    assert good[inl].sum() > 0.9 * good.sum()
    assert 0.2 < info[1] < 4.0              # a-contrario threshold found below the 4 px bound
    # epipolar residual of the un-normalised F on the inliers
    h1 = np.c_[xI[inl], np.ones(len(inl))]
    h2 = np.c_[xJ[inl], np.ones(len(inl))]
    l = h1 @ F.T
    d = np.abs((h2 * l).sum(1)) / np.hypot(l[:, 0], l[:, 1])
    assert np.median(d) < 1.5
    cv2 = pytest.importorskip("cv2")
    _, mask = cv2.findFundamentalMat(xI, xJ, cv2.FM_RANSAC, 2.0, 0.999)
    ov = (mask.ravel()[inl] > 0).mean()
    assert ov > 0.85                        # overlap with an independent estimator (not equality)


def test_acransac_rejects_pure_noise_and_small_sets(oracle):
    rng = np.random.default_rng(5)
    xI = rng.uniform([0, 0], [1920, 1080], (300, 2))
    xJ = rng.uniform([0, 0], [1920, 1080], (300, 2))
    inl, _, info = oracle.acransac_F(xI, xJ, 1920, 1080, 1920, 1080)
    assert len(inl) == 0 and info[2] == 2048       # ran all iterations, nothing meaningful
    inl, _, _ = oracle.acransac_F(xI[:7], xJ[:7], 1920, 1080, 1920, 1080)
    assert len(inl) == 0                            # nData <= MINIMUM_SAMPLES


def test_acransac_is_deterministic(oracle):
    xI, xJ, _ = _two_view(8, n=400)
    a = oracle.acransac_F(xI, xJ, 1920, 1080, 1920, 1080)[0]
    b = oracle.acransac_F(xI, xJ, 1920, 1080, 1920, 1080)[0]
    assert np.array_equal(a, b)                     # std::mt19937 default seed per call


def test_filter_pairs_drops_failed_pairs(oracle):
    sc = synth.make_scene(3, 1500, 64, "msurf", seed=21)
    pairs = synth.exhaustive_pairs(3)
    ofs, m = oracle.match_pairs(sc["descs"], sc["xys"], pairs, 0.6)
    # corrupt pair (0,2): shuffle its second indices -> no consistent geometry
    m2 = m.copy()
    seg = slice(int(ofs[1]), int(ofs[2]))
    rng = np.random.default_rng(0)
    m2["j"][seg] = rng.permutation(m2["j"][seg])
    fo, fm = oracle.filter_pairs_F(sc["xys"], sc["widths"], sc["heights"], pairs, ofs, m2)
    assert fo[1] - fo[0] > 0.8 * (ofs[1] - ofs[0])
    assert fo[2] - fo[1] == 0                        # failed estimation -> pair disappears
    assert fo[3] - fo[2] > 0.8 * (ofs[3] - ofs[2])


def test_four_point_and_homography_acransac(oracle):
    rng = np.random.default_rng(0)
    Ht = np.array([[1.02, 0.03, 8.0], [-0.02, 0.99, -5.0], [4e-5, -3e-5, 1.0]])
    n = 500
    x1 = rng.uniform([0, 0], [640, 480], (n, 2))
    q = np.c_[x1, np.ones(n)] @ Ht.T
    x2 = q[:, :2] / q[:, 2:]
    S = np.diag([1 / 640, 1 / 640, 1])
    Hn = S @ Ht @ np.linalg.inv(S)
    H = oracle.four_point(x1[200:204] / 640, x2[200:204] / 640)
    assert np.abs(H / H[2, 2] - Hn / Hn[2, 2]).max() < 1e-9
    x2n = x2 + rng.normal(0, 0.5, (n, 2))
    x2n[:150] = rng.uniform([0, 0], [640, 480], (150, 2))
    inl, He, info = oracle.acransac_H(x1, x2n, 640, 480, 640, 480)
    assert info[0] < 0 and (inl >= 150).mean() > 0.98 and len(inl) > 300
    h = np.c_[x1[inl], np.ones(len(inl))] @ He.T
    assert np.median(np.linalg.norm(h[:, :2] / h[:, 2:] - x2n[inl], axis=1)) < 1.5


# ---- essential matrix: 5-point solver + ACKernelAdaptorEssential (GeometricFilter_EMatrix_AC) ----
def _rot(w):
    th = np.linalg.norm(w)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def test_five_point_solutions_satisfy_the_constraints_and_contain_the_truth(oracle):
    rng = np.random.default_rng(1)
    found = 0
    for _ in range(60):
        R = _rot(rng.normal(size=3) * 0.3)
        t = rng.normal(size=3)
        t /= np.linalg.norm(t)
        X = rng.uniform(-1, 1, (5, 3)) + np.array([0, 0, 5])
        b1 = X / np.linalg.norm(X, axis=1, keepdims=True)
        X2 = X @ R.T + t
        b2 = X2 / np.linalg.norm(X2, axis=1, keepdims=True)
        tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
        Et = tx @ R
        Et /= np.linalg.norm(Et)
        Es = oracle.five_point(b1, b2)
        assert 1 <= len(Es) <= 10 and len(Es) % 2 == 0      # real roots of the degree-10 polynomial come in pairs
        best = np.inf
        for E in Es:
            n = np.linalg.norm(E)
            assert np.abs(np.einsum("ni,ij,nj->n", b2, E, b1)).max() < 1e-9 * n          # epipolar constraints
            assert np.abs(2 * E @ E.T @ E - np.trace(E @ E.T) * E).max() < 1e-6 * n ** 3  # trace constraint
            assert abs(np.linalg.det(E)) < 1e-6 * n ** 3
            best = min(best, np.linalg.norm(E / n - Et), np.linalg.norm(E / n + Et))
        found += best < 1e-6
    assert found == 60


def test_acransac_E_recovers_inliers(oracle):
    xI, xJ, good = _two_view(3)
    f = 1.1 * 1920
    K = np.array([f, 960.0, 540.0, f, 960.0, 540.0])
    inl, F, info = oracle.acransac_E(xI, xJ, 1920, 1080, 1920, 1080, K)
    assert info[0] < 0
    assert good[inl].mean() > 0.97 and good[inl].sum() > 0.9 * good.sum()
    h1 = np.c_[xI[inl], np.ones(len(inl))]
    h2 = np.c_[xJ[inl], np.ones(len(inl))]
    l = h1 @ F.T
    d = np.abs((h2 * l).sum(1)) / np.hypot(l[:, 0], l[:, 1])
    assert d.max() <= 4.0 + 1e-9 and np.median(d) < 1.5      # squared residual bound = Square(4.0) px^2
    # F = K2^-T E K1^-1 has the two equal singular values of an essential matrix once the Ks are removed
    Kmat = np.array([[f, 0, 960.0], [0, f, 540.0], [0, 0, 1]])
    sv = np.linalg.svd(Kmat.T @ F @ Kmat)[1]
    assert abs(sv[0] - sv[1]) < 1e-6 * sv[0] and sv[2] < 1e-6 * sv[0]
    # deterministic
    assert np.array_equal(inl, oracle.acransac_E(xI, xJ, 1920, 1080, 1920, 1080, K)[0])


def test_filter_pairs_E_skips_views_without_intrinsics(oracle):
    sc = synth.make_scene(3, 500, 32, "msurf", seed=9)
    pairs = synth.exhaustive_pairs(3)
    ofs, m = oracle.match_pairs(sc["descs"], sc["xys"], pairs, 0.8)
    f = sc["f"]
    Ks = np.array([[f, 960.0, 540.0], [f, 960.0, 540.0], [0.0, 0.0, 0.0]])
    eo, em = oracle.filter_pairs_E(sc["xys"], sc["widths"], sc["heights"], Ks, pairs, ofs, m)
    assert eo[1] - eo[0] > 5 * 2.5            # pair (0,1) kept
    assert eo[2] == eo[1] and eo[3] == eo[2]  # pairs with view 2 dropped


def test_acransac_E_overlaps_an_independent_estimator(oracle):
    """cv2.findEssentialMat (RANSAC, its own 5-point solver) on the same two-view data: inlier-set overlap, not equality."""
    cv2 = pytest.importorskip("cv2")
    xI, xJ, good = _two_view(11, n=700, outlier_frac=0.25)
    f = 1.1 * 1920
    K = np.array([f, 960.0, 540.0, f, 960.0, 540.0])
    inl, F, info = oracle.acransac_E(xI, xJ, 1920, 1080, 1920, 1080, K)
    assert info[0] < 0 and good[inl].mean() > 0.97
    Kmat = np.array([[f, 0, 960.0], [0, f, 540.0], [0, 0, 1]])
    E, mask = cv2.findEssentialMat(xI, xJ, Kmat, method=cv2.RANSAC, prob=0.999, threshold=2.0)
    mask = mask.ravel() > 0
    assert mask[inl].mean() > 0.85                      # our inliers are (mostly) cv2's inliers
    assert (good & mask).sum() > 0.8 * good.sum()       # and cv2 found the same structure
    # the oracle's F equals cv2's E up to the intrinsics, scale and sign: compare the epipolar geometry on the inliers
    E_ours = Kmat.T @ F @ Kmat
    def sampson(Em):
        x1 = np.linalg.solve(Kmat, np.c_[xI[inl], np.ones(len(inl))].T).T
        x2 = np.linalg.solve(Kmat, np.c_[xJ[inl], np.ones(len(inl))].T).T
        Ex1 = x1 @ Em.T
        Etx2 = x2 @ Em
        num = (x2 * Ex1).sum(1) ** 2
        return num / (Ex1[:, 0] ** 2 + Ex1[:, 1] ** 2 + Etx2[:, 0] ** 2 + Etx2[:, 1] ** 2)
    assert np.median(sampson(E_ours)) < 4 * max(np.median(sampson(E[:3])), 1e-9) + 1e-6


def test_acransac_H_overlaps_cv2_findHomography(oracle):
    """cv2.findHomography (RANSAC) as the independent estimator for the homography adaptor: inlier overlap and the
    same transfer error on the inliers."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(6)
    Ht = np.array([[0.97, -0.05, 30.0], [0.04, 1.03, -12.0], [2e-5, 5e-5, 1.0]])
    n = 600
    x1 = rng.uniform([0, 0], [1920, 1080], (n, 2))
    q = np.c_[x1, np.ones(n)] @ Ht.T
    x2 = q[:, :2] / q[:, 2:] + rng.normal(0, 0.6, (n, 2))
    x2[:180] = rng.uniform([0, 0], [1920, 1080], (180, 2))
    inl, He, info = oracle.acransac_H(x1, x2, 1920, 1080, 1920, 1080)
    assert info[0] < 0 and (inl >= 180).mean() > 0.98
    Hc, mask = cv2.findHomography(x1, x2, cv2.RANSAC, 3.0)
    mask = mask.ravel() > 0
    assert mask[inl].mean() > 0.9 and (mask[180:].sum() > 0.85 * (n - 180))
    def err(H):
        h = np.c_[x1[inl], np.ones(len(inl))] @ H.T
        return np.median(np.linalg.norm(h[:, :2] / h[:, 2:] - x2[inl], axis=1))
    assert err(He) < 1.5 * err(Hc) + 0.2
