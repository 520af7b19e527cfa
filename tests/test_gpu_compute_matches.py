"""-m gpu: the file-level twin of R3DComputeMatches::computeMatches() and the C++ shim class:
.feat/.desc in, matches.putative.txt / matches.f.txt out, byte-identical to the oracle's files."""
import ctypes as C
import os

import numpy as np
import pytest

from regard3d_b200 import synth

pytestmark = pytest.mark.gpu


def _write_project(oracle, tmp_path, n_img=4, n_feat=1500):
    sc = synth.make_scene(n_img, n_feat, 144, "liop", seed=41)
    names = ["image%06d" % v for v in range(n_img)]
    for v in range(n_img):
        assert oracle.save_feat(str(tmp_path / (names[v] + ".feat")), sc["feats"][v]) == 0
        assert oracle.save_desc(str(tmp_path / (names[v] + ".desc")), sc["descs"][v]) == 0
    return sc, names


def _oracle_files(oracle, sc, tmp_path):
    pairs = synth.exhaustive_pairs(len(sc["descs"]))
    ofs, m = oracle.match_pairs(sc["descs"], sc["xys"], pairs, 0.6)
    fo, fm = oracle.filter_pairs_F(sc["xys"], sc["widths"], sc["heights"], pairs, ofs, m)
    po = str(tmp_path / "oracle.putative.txt")
    pf = str(tmp_path / "oracle.f.txt")
    oracle.save_matches_txt(po, pairs, ofs, m)
    oracle.save_matches_txt(pf, pairs, fo, fm)
    return open(po).read(), open(pf).read()


def test_compute_matches_files_equal_oracle(gpu_ctx, oracle, tmp_path):
    sc, names = _write_project(oracle, tmp_path)
    seen = []
    stats = gpu_ctx.compute_matches(str(tmp_path), names, sc["widths"], sc["heights"], dist_ratio=0.6, dim=144,
                                    progress=lambda f, msg, user: seen.append(round(f, 2)))
    exp_put, exp_f = _oracle_files(oracle, sc, tmp_path)
    assert open(tmp_path / "matches.putative.txt").read() == exp_put
    assert open(tmp_path / "matches.f.txt").read() == exp_f
    assert stats["number_of_keypoints"] == [1500] * 4
    assert stats["putative_pairs"] == 6 and stats["f_pairs"] == 6
    assert seen[:2] == [0.7, 0.8]                       # the reference's progress fractions


def test_compute_matches_essential_and_homography_files(gpu_ctx, oracle, tmp_path):
    """computeEssentialMatrix_ / computeHomographyMatrix_: matches.e.txt (after the reference's poor-overlap
    removal, src/R3DComputeMatches.cpp:2173-2191) and matches.h.txt equal the oracle's."""
    sc, names = _write_project(oracle, tmp_path)
    seen = []
    stats = gpu_ctx.compute_matches(str(tmp_path), names, sc["widths"], sc["heights"], dist_ratio=0.6, dim=144,
                                    compute_essential=True, compute_homography=True,
                                    progress=lambda f, msg, user: seen.append(round(f, 2)))
    pairs = synth.exhaustive_pairs(len(names))
    ofs, m = oracle.match_pairs(sc["descs"], sc["xys"], pairs, 0.6)
    Ks = np.array([[1.1 * 1920, 960.0, 540.0]] * len(names))
    eo, em = oracle.filter_pairs_E(sc["xys"], sc["widths"], sc["heights"], Ks, pairs, ofs, m)
    keep_ofs, keep_m = [0], []
    for k in range(len(pairs)):
        ne, npu = int(eo[k + 1] - eo[k]), int(ofs[k + 1] - ofs[k])
        if ne == 0 or ne < 50 or np.float32(ne) / np.float32(npu) < np.float32(0.3):
            keep_ofs.append(keep_ofs[-1])
            continue
        keep_m.append(em[int(eo[k]):int(eo[k + 1])])
        keep_ofs.append(keep_ofs[-1] + ne)
    pe = str(tmp_path / "oracle.e.txt")
    oracle.save_matches_txt(pe, pairs, np.array(keep_ofs, np.uint64), np.concatenate(keep_m))
    assert open(tmp_path / "matches.e.txt").read() == open(pe).read()
    ho, hm = oracle.filter_pairs_H(sc["xys"], sc["widths"], sc["heights"], pairs, ofs, m)
    ph = str(tmp_path / "oracle.h.txt")
    oracle.save_matches_txt(ph, pairs, ho, hm)
    assert open(tmp_path / "matches.h.txt").read() == open(ph).read()
    assert stats["e_pairs"] == sum(1 for k in range(len(pairs)) if keep_ofs[k + 1] > keep_ofs[k]) and stats["e_pairs"] > 0
    assert seen[:4] == [0.7, 0.8, 0.9, 0.95]


def test_cpp_shim_class(r3dlib, oracle, tmp_path):
    sc, names = _write_project(oracle, tmp_path, n_img=3, n_feat=1000)
    lib = r3dlib.lib()
    n = len(names)
    files = (C.c_char_p * n)(*[(nm + ".jpg").encode() for nm in names])
    w = (C.c_uint32 * n)(*[1920] * n)
    h = (C.c_uint32 * n)(*[1080] * n)
    kp = (C.c_uint32 * n)()
    pp, fp = C.c_uint64(), C.c_uint64()
    last = C.c_float()
    rc = lib.r3d_shim_compute_matches(str(tmp_path).encode(), files, w, h, n, C.c_float(0.6), 4, kp, C.byref(pp),
                                      C.byref(fp), C.byref(last))
    assert rc == 0
    assert list(kp) == [1000] * 3 and pp.value == 3 and fp.value == 3 and last.value == 1.0
    exp_put, exp_f = _oracle_files(oracle, sc, tmp_path)
    assert open(tmp_path / "matches.putative.txt").read() == exp_put
    assert open(tmp_path / "matches.f.txt").read() == exp_f


def test_missing_region_file_is_an_error(gpu_ctx, r3dlib, tmp_path):
    with pytest.raises(r3dlib.R3DError) as e:
        gpu_ctx.compute_matches(str(tmp_path), ["image000000", "image000001"], [640, 640], [480, 480])
    assert e.value.code == -4


def test_compute_matches_svg_output(gpu_ctx, oracle, tmp_path):
    """computeMatches(..., svgOutput = true, ...): the two adjacency-matrix SVGs (src/R3DComputeMatches.cpp:2074, :2238)."""
    sc, names = _write_project(oracle, tmp_path, n_img=3, n_feat=800)
    gpu_ctx.compute_matches(str(tmp_path), names, sc["widths"], sc["heights"], dist_ratio=0.6, dim=144, svg_output=True)
    for name in ("PutativeAdjacencyMatrix.svg", "GeometricAdjacencyMatrix.svg"):
        txt = (tmp_path / name).read_text()
        assert txt.startswith("<?xml") and txt.rstrip().endswith("</svg>") and txt.count("<rect") == 3   # pairs (0,1) (0,2) (1,2)


def test_compute_matches_with_the_cascade_hashing_matcher(gpu_ctx, oracle, r3dlib, tmp_path):
    """matching_algorithm = R3D_MATCHING_CASCADE_HASHING (an extension: the reference's switch has no such entry): the
    putative and F files equal the oracle's cascade matcher + F filter."""
    sc, names = _write_project(oracle, tmp_path)
    gpu_ctx.compute_matches(str(tmp_path), names, sc["widths"], sc["heights"], dist_ratio=0.8, dim=144,
                            matching_algorithm=r3dlib.MATCHING_CASCADE_HASHING)
    pairs = synth.exhaustive_pairs(len(names))
    ofs, m = oracle.cascade_match_pairs(sc["descs"], sc["xys"], pairs, 0.8)
    fo, fm = oracle.filter_pairs_F(sc["xys"], sc["widths"], sc["heights"], pairs, ofs, m)
    po, pf = str(tmp_path / "oracle.putative.txt"), str(tmp_path / "oracle.f.txt")
    oracle.save_matches_txt(po, pairs, ofs, m)
    oracle.save_matches_txt(pf, pairs, fo, fm)
    assert len(m) > 500
    assert open(tmp_path / "matches.putative.txt").read() == open(po).read()
    assert open(tmp_path / "matches.f.txt").read() == open(pf).read()
