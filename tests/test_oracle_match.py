"""Oracle self-checks for the putative-matching stage (the reference ships no golden
vectors, so the restatement is pinned by hand-checkable cases and independent implementations)."""
import numpy as np
import pytest

from regard3d_b200 import synth


def test_l2_hand_case(oracle):
    a = np.array([1, 2, 3, 4, 5], np.float32)
    b = np.array([0, 0, 0, 0, 0], np.float32)
    assert oracle.l2(a, b) == 55.0
    assert oracle.l2(np.array([10, 250, 3], np.uint8), np.array([20, 5, 3], np.uint8)) == 100.0 + 245.0 ** 2


def test_l2_accumulation_order(oracle):
    # 4-way unrolled float accumulation (openMVG::matching::L2<float>): reproduce it in numpy
    rng = np.random.default_rng(0)
    a = rng.standard_normal(147).astype(np.float32)
    b = rng.standard_normal(147).astype(np.float32)
    r = np.float32(0)
    k = 0
    while k + 3 < 147:
        d = (a[k:k + 4] - b[k:k + 4]).astype(np.float32)
        s = np.float32(np.float32(np.float32(d[0] * d[0]) + np.float32(d[1] * d[1])) + np.float32(d[2] * d[2]))
        s = np.float32(s + np.float32(d[3] * d[3]))
        r = np.float32(r + s)
        k += 4
    while k < 147:
        d = np.float32(a[k] - b[k])
        r = np.float32(r + np.float32(d * d))
        k += 1
    assert np.float32(oracle.l2(a, b)) == r


def test_search_neighbours_vs_numpy_and_cv2(oracle):
    sc = synth.make_scene(2, 800, 64, "msurf", seed=2)
    idx, dist = oracle.search_neighbours(sc["descs"][0], sc["descs"][1])
    A = sc["descs"][0].astype(np.float64)
    B = sc["descs"][1].astype(np.float64)
    D = ((B[:, None, :] - A[None, :, :]) ** 2).sum(-1)
    o = np.argsort(D, 1)[:, :2]
    assert (o == idx).all()
    assert np.allclose(np.take_along_axis(D, o, 1), dist, rtol=1e-5)
    cv2 = pytest.importorskip("cv2")
    knn = cv2.BFMatcher(cv2.NORM_L2).knnMatch(sc["descs"][1], sc["descs"][0], k=2)
    cv_idx = np.array([[m[0].trainIdx, m[1].trainIdx] for m in knn])
    assert (cv_idx == idx).mean() > 0.999


def test_search_neighbours_false_when_db_too_small(oracle):
    db = np.ones((1, 8), np.float32)
    q = np.ones((3, 8), np.float32)
    assert oracle.search_neighbours(db, q) is None


def test_ratio_is_strict_and_squared(oracle):
    # database: two points at squared distances 1 and 4 from the query; ratio test d1 < r^2 d2
    db = np.array([[1, 0], [2, 0]], np.float32)
    q = np.array([[0, 0]], np.float32)
    xy = np.zeros((2, 2), np.float32)
    xy[1] = 1
    xq = np.zeros((1, 2), np.float32)
    assert len(oracle.match_distance_ratio(db, xy, q, xq, 0.5)) == 0    # 1 < 0.25*4 is false (strict)
    assert len(oracle.match_distance_ratio(db, xy, q, xq, 0.51)) == 1
    m = oracle.match_distance_ratio(db, xy, q, xq, 0.9)
    assert (m["i"][0], m["j"][0]) == (0, 0)     # i_ indexes the database image, j_ the query image


def test_tie_between_best_and_second_yields_no_match(oracle):
    db = np.array([[1, 0], [-1, 0], [5, 5]], np.float32)
    q = np.array([[0, 0]], np.float32)
    assert len(oracle.match_distance_ratio(db, np.zeros((3, 2), np.float32), q, np.zeros((1, 2), np.float32), 0.99)) == 0


def test_no_mutual_check_many_to_one(oracle):
    # two queries close to the same database row both match it (there is no cross-check upstream)
    db = np.array([[0, 0], [10, 10], [20, 20]], np.float32)
    q = np.array([[0.1, 0], [0, 0.1]], np.float32)
    xyI = np.array([[0, 0], [1, 1], [2, 2]], np.float32)
    xyJ = np.array([[5, 5], [6, 7]], np.float32)
    m = oracle.match_distance_ratio(db, xyI, q, xyJ, 0.6)
    assert sorted(zip(m["i"].tolist(), m["j"].tolist())) == [(0, 0)] or len(m) == 2


def test_coord_dedup_exact_duplicates_and_quirk(oracle):
    from oracle.pyoracle import indmatch_dtype
    xyI = np.array([[1, 1], [1, 1], [3, 9], [4, 9]], np.float32)
    xyJ = np.array([[7, 7], [7, 7], [8, 8], [9, 9]], np.float32)
    m = np.array([(0, 0), (1, 1), (2, 2), (3, 3)], indmatch_dtype)
    out = oracle.coord_dedup(m, xyI, xyJ)
    s = set(zip(out["i"].tolist(), out["j"].tolist()))
    # (0,0) and (1,1) have identical coordinates -> one survives
    assert len(s & {(0, 0), (1, 1)}) == 1
    # upstream comparator quirk: equal y1 with different x1 compares "equivalent" -> one of (2,2),(3,3) is dropped
    assert len(s & {(2, 2), (3, 3)}) == 1


def test_match_pairs_skips_empty_and_keeps_map_semantics(oracle):
    sc = synth.make_scene(3, 300, 32, "msurf", seed=4)
    descs = list(sc["descs"])
    xys = list(sc["xys"])
    descs[1] = np.zeros((0, 32), np.float32)
    xys[1] = np.zeros((0, 2), np.float32)
    pairs = synth.exhaustive_pairs(3)
    ofs, m = oracle.match_pairs(descs, xys, pairs, 0.8)
    assert ofs[1] - ofs[0] == 0 and ofs[3] - ofs[2] == 0   # pairs touching the empty view
    assert ofs[2] - ofs[1] > 0


def test_u8_and_float_integer_descriptors_agree(oracle):
    sc8 = synth.make_scene(2, 500, 128, "sift", seed=6, as_u8=True)
    scf = synth.make_scene(2, 500, 128, "sift", seed=6, as_u8=False)
    assert (sc8["descs"][0].astype(np.float32) == scf["descs"][0]).all()
    p = synth.exhaustive_pairs(2)
    o1, m1 = oracle.match_pairs(sc8["descs"], sc8["xys"], p, 0.7)
    o2, m2 = oracle.match_pairs(scf["descs"], scf["xys"], p, 0.7)
    assert (m1 == m2).all()
