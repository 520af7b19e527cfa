"""-m gpu: AC-RANSAC fundamental filter through the C ABI vs the CPU oracle: identical inlier
sequences (same std::mt19937 stream, same decisions)."""
import numpy as np
import pytest

from regard3d_b200 import synth

pytestmark = pytest.mark.gpu


def _setup(ctx, sc):
    ctx.clear_regions()
    for v, (d, x) in enumerate(zip(sc["descs"], sc["xys"])):
        ctx.upload_regions(v, d, x)


def _default_Ks(sc):
    # R3DProject's approximation (src/R3DProject.cpp:1149-1159): f = 1.1 max(w,h), pp at the image centre
    return np.array([[1.1 * max(int(w), int(h)), w / 2.0, h / 2.0] for w, h in zip(sc["widths"], sc["heights"])])


def _check(ctx, oracle, r3dlib, sc, pairs, ofs, m, max_iter=2048, model="F", Ks=None):
    put = r3dlib.Matches.from_csr(pairs, ofs, m)
    mid = {"F": r3dlib.MODEL_F, "H": r3dlib.MODEL_H, "E": r3dlib.MODEL_E}[model]
    if model == "E" and Ks is None:
        Ks = _default_Ks(sc)
    got = ctx.filter_pairs(put, sc["widths"], sc["heights"], max_iter=max_iter, model=mid, Ks=Ks).to_dict()
    if model == "E":
        fo, fm = oracle.filter_pairs_E(sc["xys"], sc["widths"], sc["heights"], Ks, pairs, ofs, m, max_iter=max_iter)
    else:
        fo, fm = oracle.filter_pairs_F(sc["xys"], sc["widths"], sc["heights"], pairs, ofs, m, max_iter=max_iter, model=model)
    n_pairs_exp = 0
    for k, (I, J) in enumerate(pairs):
        e = fm[int(fo[k]):int(fo[k + 1])]
        g = got.get((int(I), int(J)))
        if len(e) == 0:
            assert g is None
            continue
        n_pairs_exp += 1
        assert g is not None, (I, J)
        assert np.array_equal(g, e), "pair %s: inlier sequence differs" % ((I, J),)
    assert len(got) == n_pairs_exp
    return got


def test_filter_equals_oracle_clean_scene(gpu_ctx, oracle, r3dlib):
    sc = synth.make_scene(4, 2000, 64, "msurf", seed=31)
    pairs = synth.exhaustive_pairs(4)
    _setup(gpu_ctx, sc)
    ofs, m = oracle.match_pairs(sc["descs"], sc["xys"], pairs, 0.6)
    _check(gpu_ctx, oracle, r3dlib, sc, pairs, ofs, m)
    t = gpu_ctx.filter_timing()
    assert t["kernel_launches"] >= 2 and t["hypotheses"] > 0


def test_filter_equals_oracle_with_outliers_and_failures(gpu_ctx, oracle, r3dlib):
    sc = synth.make_scene(4, 1500, 64, "msurf", seed=32)
    pairs = synth.exhaustive_pairs(4)
    _setup(gpu_ctx, sc)
    ofs, m = oracle.match_pairs(sc["descs"], sc["xys"], pairs, 0.8)
    rng = np.random.default_rng(1)
    m2 = m.copy()
    # pair 0: 40 % gross outliers; pair 1: all shuffled (must fail); pair 2: only 12 matches
    s0 = slice(int(ofs[0]), int(ofs[1]))
    n0 = int(ofs[1] - ofs[0])
    bad = rng.random(n0) < 0.4
    j0 = m2["j"][s0].copy()
    j0[bad] = rng.integers(0, 1500, bad.sum())
    m2["j"][s0] = j0
    s1 = slice(int(ofs[1]), int(ofs[2]))
    m2["j"][s1] = rng.permutation(m2["j"][s1])
    keep = np.ones(len(m2), bool)
    keep[int(ofs[2]) + 12:int(ofs[3])] = False
    new_ofs = np.zeros_like(ofs)
    for k in range(len(pairs)):
        new_ofs[k + 1] = new_ofs[k] + keep[int(ofs[k]):int(ofs[k + 1])].sum()
    m2 = m2[keep]
    got = _check(gpu_ctx, oracle, r3dlib, sc, pairs, new_ofs, m2)
    assert (0, 2) not in got


def test_filter_small_iteration_budget_and_tiny_pairs(gpu_ctx, oracle, r3dlib):
    sc = synth.make_scene(3, 400, 32, "msurf", seed=33)
    pairs = synth.exhaustive_pairs(3)
    _setup(gpu_ctx, sc)
    ofs, m = oracle.match_pairs(sc["descs"], sc["xys"], pairs, 0.9)
    _check(gpu_ctx, oracle, r3dlib, sc, pairs, ofs, m, max_iter=64)
    # pairs with <= 7 putatives are skipped by ACRANSAC
    small_ofs = np.array([0, 7, 7, 15], np.uint64)
    small_m = np.concatenate([m[int(ofs[0]):int(ofs[0]) + 7], m[int(ofs[2]):int(ofs[2]) + 8]])
    _check(gpu_ctx, oracle, r3dlib, sc, pairs, small_ofs, small_m)


def _planar_scene(seed, n_img=3, n_feat=1200):
    """Views of one plane: features related by homographies (H is the right model, F is degenerate)."""
    rng = np.random.default_rng(seed)
    base = rng.uniform([100, 100], [1800, 980], (n_feat, 2))
    xys = []
    for v in range(n_img):
        H = np.array([[1 + 0.03 * v, 0.02 * v, 12.0 * v], [-0.015 * v, 1 - 0.02 * v, -9.0 * v], [2e-5 * v, -1e-5 * v, 1.0]])
        q = np.c_[base, np.ones(n_feat)] @ H.T
        xy = q[:, :2] / q[:, 2:] + 0.5 * rng.standard_normal((n_feat, 2))
        xys.append(synth.round_sig(xy).astype(np.float32))
    return xys


def test_homography_filter_equals_oracle(gpu_ctx, oracle, r3dlib):
    xys = _planar_scene(51)
    n = len(xys[0])
    rng = np.random.default_rng(2)
    descs = [np.zeros((n, 16), np.float32) for _ in xys]          # descriptors are irrelevant for the filter
    gpu_ctx.clear_regions()
    for v in range(3):
        gpu_ctx.upload_regions(v, descs[v], xys[v])
    pairs = synth.exhaustive_pairs(3)
    ofs = np.array([0, 900, 1800, 2700], np.uint64)
    chunks = []
    for _ in range(3):
        i = rng.permutation(n)[:900].astype(np.uint32)
        j = i.copy()
        bad = rng.random(900) < 0.3
        j[bad] = rng.integers(0, n, bad.sum())
        chunks.append(np.array(list(zip(i.tolist(), j.tolist())), r3dlib.indmatch_dtype))
    m = np.concatenate(chunks)
    sc = {"xys": xys, "widths": np.full(3, 1920, np.uint32), "heights": np.full(3, 1080, np.uint32)}
    got = _check(gpu_ctx, oracle, r3dlib, sc, pairs, ofs, m, model="H")
    assert len(got) == 3
    for k, v in got.items():
        assert 550 < len(v) < 700


def test_essential_filter_equals_oracle(gpu_ctx, oracle, r3dlib):
    """GeometricFilter_EMatrix_AC: 5-point solver on bearing vectors, <= 10 models per sample -- identical inlier
    sequences, including a pair with gross outliers, a hopeless pair and a view without intrinsics."""
    sc = synth.make_scene(4, 1500, 64, "msurf", seed=33)
    pairs = synth.exhaustive_pairs(4)
    _setup(gpu_ctx, sc)
    ofs, m = oracle.match_pairs(sc["descs"], sc["xys"], pairs, 0.8)
    rng = np.random.default_rng(4)
    m2 = m.copy()
    s0 = slice(int(ofs[0]), int(ofs[1]))
    bad = rng.random(int(ofs[1] - ofs[0])) < 0.35
    j0 = m2["j"][s0].copy()
    j0[bad] = rng.integers(0, 1500, bad.sum())
    m2["j"][s0] = j0
    s1 = slice(int(ofs[1]), int(ofs[2]))
    m2["j"][s1] = rng.permutation(1500)[: int(ofs[2] - ofs[1])]      # all wrong: must fail
    Ks = _default_Ks(sc)
    got = _check(gpu_ctx, oracle, r3dlib, sc, pairs, ofs, m2, model="E", Ks=Ks)
    assert (0, 1) in got and (0, 2) not in got
    truth = sc["truth"]
    g = got[(0, 1)]
    assert (truth[0][g["i"]] == truth[1][g["j"]]).mean() > 0.97       # the kept matches are true correspondences
    Ks[3, 0] = 0.0                                                    # view 3 has no pinhole intrinsic
    got = _check(gpu_ctx, oracle, r3dlib, sc, pairs, ofs, m2, model="E", Ks=Ks)
    assert all(3 not in k for k in got)
    assert gpu_ctx.filter_timing()["hypotheses"] > 0
