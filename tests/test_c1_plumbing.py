"""BASELINE configs[0] (C1): 2 synthetic 640x480 images, AKAZE keypoints + float descriptors, brute-force
L2 -- the reference's own CPU-runnable case, used here as plumbing: rendered images -> cv2.AKAZE ->
.feat/.desc files -> matching (+ F filter).  CPU part: the oracle on real detector output and the file
round trip; GPU part: the product equals the oracle on the same files."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def _render_pair(seed=20260924 + 1):
    rng = np.random.default_rng(seed)
    img = np.full((480, 640), 127, np.uint8)
    for _ in range(160):
        pts = (rng.uniform([0, 0], [640, 480], (rng.integers(3, 7), 2))).astype(np.int32)
        cv2.fillPoly(img, [pts], int(rng.integers(0, 256)))
    for _ in range(200):
        c = tuple(int(v) for v in rng.uniform([0, 0], [640, 480]))
        cv2.circle(img, c, int(rng.integers(2, 12)), int(rng.integers(0, 256)), -1)
    img = cv2.GaussianBlur(img, (0, 0), 1.0)
    H = np.array([[1.02, 0.03, 8.0], [-0.02, 0.99, -5.0], [4e-5, -3e-5, 1.0]])
    img2 = cv2.warpPerspective(img, H, (640, 480), borderValue=127)
    noise = rng.normal(0, 2.0, img2.shape)
    img2 = np.clip(img2.astype(np.float64) + noise, 0, 255).astype(np.uint8)
    return img, img2, H


def _features(img):
    ak = cv2.AKAZE_create(descriptor_type=cv2.AKAZE_DESCRIPTOR_KAZE, threshold=0.001)
    kp, desc = ak.detectAndCompute(img, None)
    xyso = np.array([[k.pt[0], k.pt[1], k.size / 2.0, k.angle] for k in kp], np.float32)   # Regard3DFeatures.cpp:831-835
    return xyso, np.ascontiguousarray(desc, np.float32)


@pytest.fixture(scope="module")
def c1(oracle, tmp_path_factory):
    d = tmp_path_factory.mktemp("c1")
    a, b, H = _render_pair()
    fa, da = _features(a)
    fb, db = _features(b)
    assert len(fa) > 200 and len(fb) > 200 and da.shape[1] == 64
    names = ["image000000", "image000001"]
    for nm, f, de in zip(names, (fa, fb), (da, db)):
        assert oracle.save_feat(str(d / (nm + ".feat")), f) == 0
        assert oracle.save_desc(str(d / (nm + ".desc")), de) == 0
    # what the matching stage sees is what the files hold (.feat is text with 6 significant digits)
    feats = [oracle.load_feat(str(d / (nm + ".feat"))) for nm in names]
    descs = [oracle.load_desc(str(d / (nm + ".desc")), 64) for nm in names]
    return {"dir": d, "names": names, "feats": feats, "descs": descs, "H": H}


def test_c1_oracle_plumbing(oracle, c1):
    pairs = np.array([[0, 1]], np.uint32)
    xys = [f[:, :2].copy() for f in c1["feats"]]
    ofs, m = oracle.match_pairs(c1["descs"], xys, pairs, 0.8)
    assert len(m) > 50
    # matches follow the homography the second image was rendered with
    p0 = xys[0][m["i"]].astype(np.float64)
    p1 = xys[1][m["j"]].astype(np.float64)
    q = np.c_[p0, np.ones(len(p0))] @ c1["H"].T
    q = q[:, :2] / q[:, 2:]
    assert np.median(np.linalg.norm(q - p1, axis=1)) < 2.0
    # independent matcher on the same descriptors
    knn = cv2.BFMatcher(cv2.NORM_L2).knnMatch(c1["descs"][1], c1["descs"][0], k=2)
    cvm = {(a.trainIdx, a.queryIdx) for a, b in knn if a.distance ** 2 < 0.8 ** 2 * b.distance ** 2}
    ours = set(zip(m["i"].tolist(), m["j"].tolist()))
    assert len(ours - cvm) <= 0.02 * len(ours) + 2       # coordinate dedup can only remove matches
    fo, fm = oracle.filter_pairs_F(xys, [640, 640], [480, 480], pairs, ofs, m)
    assert len(fm) > 0.6 * len(m)


@pytest.mark.gpu
def test_c1_gpu_equals_oracle(gpu_ctx, oracle, c1):
    pairs = np.array([[0, 1]], np.uint32)
    xys = [f[:, :2].copy() for f in c1["feats"]]
    ofs, m = oracle.match_pairs(c1["descs"], xys, pairs, 0.8)
    fo, fm = oracle.filter_pairs_F(xys, [640, 640], [480, 480], pairs, ofs, m)
    stats = gpu_ctx.compute_matches(str(c1["dir"]), c1["names"], [640, 640], [480, 480], dist_ratio=0.8, dim=64)
    assert stats["number_of_keypoints"] == [len(c1["feats"][0]), len(c1["feats"][1])]
    exp_put = str(c1["dir"] / "exp.putative.txt")
    exp_f = str(c1["dir"] / "exp.f.txt")
    oracle.save_matches_txt(exp_put, pairs, ofs, m)
    oracle.save_matches_txt(exp_f, pairs, fo, fm)
    assert open(c1["dir"] / "matches.putative.txt").read() == open(exp_put).read()
    assert open(c1["dir"] / "matches.f.txt").read() == open(exp_f).read()
