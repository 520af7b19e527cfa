"""File formats and the C-ABI surface; no GPU compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

from regard3d_b200 import synth

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_header_symbols_are_exported(r3dlib):
    hdr = open(os.path.join(ROOT, "include", "r3dgpu.h")).read()
    names = set(re.findall(r"\b(r3d_[a-z0-9_]+)\s*\(", hdr))
    names -= {"r3d_progress_cb"}
    lib = r3dlib.lib()
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    assert set(r3dlib.EXPORTS) <= names
    assert lib.r3d_abi_version() == 3


def test_create_fails_loudly_without_gpu(r3dlib):
    import torch
    if torch.cuda.is_available():
        return
    try:
        r3dlib.Context((0,))
    except r3dlib.R3DError as e:
        assert e.code == -6 and "no CPU fallback" in str(e)
    else:
        raise AssertionError("Context() must fail without a GPU")


def test_product_never_imports_oracle():
    import ast
    pkg = os.path.join(ROOT, "regard3d_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            p = os.path.join(dp, f)
            if f.endswith(".py"):
                for node in ast.walk(ast.parse(open(p).read())):
                    if isinstance(node, (ast.Import, ast.ImportFrom)):
                        mod = getattr(node, "module", None) or ""
                        names = [a.name for a in node.names]
                        assert not mod.startswith("oracle") and not any(n.startswith("oracle") for n in names), p
            elif f.endswith((".cu", ".cpp", ".cuh", ".h")):
                src = open(p).read()
                assert "oracle/" not in src.replace("// oracle", "") or "#include" not in "".join(
                    l for l in src.splitlines() if "oracle" in l), p


def test_feat_desc_roundtrip(oracle, tmp_path):
    sc = synth.make_scene(1, 200, 144, "liop", seed=1)
    fp = str(tmp_path / "image000000.feat")
    dp = str(tmp_path / "image000000.desc")
    assert oracle.save_feat(fp, sc["feats"][0]) == 0
    assert oracle.save_desc(dp, sc["descs"][0]) == 0
    # .feat is text with default ostream precision (6 significant digits); synthetic coordinates are
    # pre-rounded to 6 digits so the round trip is exact
    first = open(fp).readline().split()
    assert len(first) == 4
    f2 = oracle.load_feat(fp)
    assert np.array_equal(f2, sc["feats"][0])
    # .desc: size_t count + raw float32 rows
    raw = open(dp, "rb").read()
    assert int.from_bytes(raw[:8], "little") == 200 and len(raw) == 8 + 200 * 144 * 4
    d2 = oracle.load_desc(dp, 144)
    assert np.array_equal(d2, sc["descs"][0])


def test_matches_txt_format_and_roundtrip(oracle, r3dlib, tmp_path):
    pairs = np.array([[0, 2], [0, 1], [1, 2]], np.uint32)          # deliberately unsorted
    ofs = np.array([0, 2, 2, 5], np.uint64)                        # pair (0,1) is empty
    m = np.array([(5, 6), (7, 8), (1, 2), (3, 4), (9, 9)], r3dlib.indmatch_dtype)
    p_or = str(tmp_path / "o.txt")
    p_r3 = str(tmp_path / "r.txt")
    assert oracle.save_matches_txt(p_or, pairs, ofs, m) == 0
    mm = r3dlib.Matches.from_csr(pairs, ofs, m)
    assert mm.num_pairs == 2 and mm.total == 5
    mm.save_txt(p_r3)
    txt = open(p_r3).read()
    assert txt == open(p_or).read()
    assert txt == "0 2\n2\n5 6\n7 8\n1 2\n3\n1 2\n3 4\n9 9\n"      # std::map order, empty pair absent
    back = r3dlib.Matches.load_txt(p_r3).to_dict()
    assert sorted(back.keys()) == [(0, 2), (1, 2)]
    assert back[(1, 2)]["j"].tolist() == [2, 4, 9]


def test_coordinate_dedup_replay_equals_std_set(r3dlib, oracle):
    """The host tail replays libstdc++'s std::set range insertion on a compact node array (match_post.cpp).  The
    comparator is not a strict weak ordering, so the outcome depends on the exact tree procedure: pin the replay
    against the real std::set (the oracle's coord_dedup) on adversarial inputs -- few distinct x / y values, shared
    keypoints, exact duplicates -- and on plain random ones."""
    import ctypes as C
    lib = r3dlib.lib()
    rng = np.random.default_rng(12)
    for trial in range(300):
        n_feat = int(rng.integers(2, 400))
        levels = int(rng.choice([2, 3, 5, 17, 1000]))
        xyI = (rng.integers(0, levels, (n_feat, 2)) * 1.5).astype(np.float32)
        xyJ = (rng.integers(0, levels, (n_feat, 2)) * 0.75).astype(np.float32)
        n = int(rng.integers(1, 1200))
        m = np.zeros(n, r3dlib.indmatch_dtype)
        m["i"] = rng.integers(0, n_feat, n)
        m["j"] = rng.integers(0, n_feat, n)
        # the reference's order of operations: (i,j) sort + unique, then the coordinate set
        exp = np.unique(np.stack([m["i"], m["j"]], 1), axis=0)
        e = np.zeros(len(exp), r3dlib.indmatch_dtype)
        e["i"], e["j"] = exp[:, 0], exp[:, 1]
        want = oracle.coord_dedup(e, xyI, xyJ)
        got = m.copy()
        k = lib.r3d_debug_post_process(got.ctypes.data_as(C.c_void_p), C.c_int64(n), xyI.ctypes.data_as(C.c_void_p),
                                       xyJ.ctypes.data_as(C.c_void_p), 1)
        assert k == len(want), (trial, k, len(want))
        assert np.array_equal(got[:k], want), trial


def test_lockstep_replay_of_several_pairs_equals_std_set(r3dlib, oracle):
    """The batch tails advance up to 4 pairs per host thread in lockstep: same results as pair by pair."""
    import ctypes as C
    lib = r3dlib.lib()
    rng = np.random.default_rng(21)
    for trial in range(60):
        lanes = int(rng.integers(1, 5))
        ms, xyIs, xyJs, wants = [], [], [], []
        for t in range(lanes):
            n_feat = int(rng.integers(2, 300))
            levels = int(rng.choice([2, 4, 9, 1000]))
            xyI = (rng.integers(0, levels, (n_feat, 2)) * 1.25).astype(np.float32)
            xyJ = (rng.integers(0, levels, (n_feat, 2)) * 0.5).astype(np.float32)
            n = int(rng.integers(0, 900))
            m = np.zeros(n, r3dlib.indmatch_dtype)
            m["i"] = rng.integers(0, n_feat, n)
            m["j"] = rng.integers(0, n_feat, n)
            exp = np.unique(np.stack([m["i"], m["j"]], 1), axis=0) if n else np.zeros((0, 2), np.uint32)
            e = np.zeros(len(exp), r3dlib.indmatch_dtype)
            e["i"], e["j"] = exp[:, 0], exp[:, 1]
            wants.append(oracle.coord_dedup(e, xyI, xyJ) if len(e) else e)
            ms.append(m.copy()); xyIs.append(xyI); xyJs.append(xyJ)
        mp = (C.c_void_p * lanes)(*[m.ctypes.data for m in ms])
        ip = (C.c_void_p * lanes)(*[x.ctypes.data for x in xyIs])
        jp = (C.c_void_p * lanes)(*[x.ctypes.data for x in xyJs])
        counts = (C.c_uint64 * lanes)(*[len(m) for m in ms])
        assert lib.r3d_debug_post_process_many(lanes, mp, counts, ip, jp, 1) == 0
        for t in range(lanes):
            assert counts[t] == len(wants[t]), (trial, t)
            assert np.array_equal(ms[t][:counts[t]], wants[t]), (trial, t)


def test_bench_reference_arm_prints_one_json_line(oracle):
    """`bench.py --impl reference` (the driver's reference arm) needs no GPU: one JSON line on stdout with the contract's
    keys; a 2-pair sample keeps it short."""
    import json
    import subprocess
    import sys
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    env = dict(os.environ, R3D_REF_SAMPLE_PAIRS="2")
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "matched_image_pairs_per_sec_exhaustive" and d["unit"] == "pairs/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_ranked_descent_free_replay_equals_std_set(r3dlib, oracle):
    """The batch tails replace the tree descents by a bitset over the view's y ranks (match_post.cpp::ranked_replay) and
    keep the classic descent only for keypoints whose x is shared.  Same adversarial inputs as above (few distinct x / y,
    shared keypoints, exact duplicates, negative and zero coordinates) plus realistic sparse ones, against std::set."""
    import ctypes as C
    lib = r3dlib.lib()
    rng = np.random.default_rng(33)
    for trial in range(400):
        n_feat = int(rng.integers(2, 500))
        levels = int(rng.choice([2, 3, 5, 17, 200, 100000]))
        xyI = ((rng.integers(0, levels, (n_feat, 2)) - levels // 3) * 1.5).astype(np.float32)
        if trial % 7 == 0:
            xyI[rng.integers(0, n_feat)] = (-0.0, 0.0)      # signed zeros compare equal
        xyJ = (rng.integers(0, levels, (n_feat, 2)) * 0.75).astype(np.float32)
        n = int(rng.integers(1, 1500))
        m = np.zeros(n, r3dlib.indmatch_dtype)
        m["i"] = rng.integers(0, n_feat, n)
        m["j"] = rng.integers(0, n_feat, n)
        exp = np.unique(np.stack([m["i"], m["j"]], 1), axis=0)
        e = np.zeros(len(exp), r3dlib.indmatch_dtype)
        e["i"], e["j"] = exp[:, 0], exp[:, 1]
        want = oracle.coord_dedup(e, xyI, xyJ)
        got = m.copy()
        k = lib.r3d_debug_post_process_ranked(got.ctypes.data_as(C.c_void_p), C.c_int64(n), xyI.ctypes.data_as(C.c_void_p),
                                              C.c_uint32(n_feat), xyJ.ctypes.data_as(C.c_void_p))
        assert k == len(want), (trial, levels, k, len(want))
        assert np.array_equal(got[:k], want), (trial, levels)


def test_matches_container_spans_and_edge_cases(r3dlib, tmp_path):
    """r3d_matches holds spans into slabs: duplicate pairs keep the first (map::insert), empty lists vanish, a loaded
    file with a zero-count pair keeps it (matching::Load), and handles stay valid while any view of them is alive."""
    pairs = np.array([[3, 4], [0, 1], [3, 4], [2, 5]], np.uint32)
    ofs = np.array([0, 2, 3, 5, 5], np.uint64)                      # (3,4) twice, (2,5) empty
    m = np.array([(1, 1), (2, 2), (7, 7), (8, 8), (9, 9)], r3dlib.indmatch_dtype)
    mm = r3dlib.Matches.from_csr(pairs, ofs, m)
    d = mm.to_dict()
    assert sorted(d) == [(0, 1), (3, 4)] and d[(3, 4)]["i"].tolist() == [1, 2] and mm.total == 3
    p = tmp_path / "z.txt"
    p.write_text("5 6\n0\n1 2\n1\n4 4\n")
    back = r3dlib.Matches.load_txt(str(p))
    assert back.num_pairs == 2 and back.total == 1
    assert back.to_dict()[(1, 2)]["j"].tolist() == [4]
    with pytest.raises(r3dlib.R3DError):
        r3dlib.Matches.load_txt(str(tmp_path / "missing.txt"))


def test_ranked_replay_falls_back_for_non_finite_positions(r3dlib, oracle):
    """NaN / inf positions: no rank tables, the classic replay evaluates the comparator exactly like std::set
    (a NaN x1 makes the upstream comparator return false: neither x1 < x1' nor x1 > x1').  Many seeds and all
    variants: one seed alone can hide a wrong comparator (round-1 advisor finding)."""
    import ctypes as C
    lib = r3dlib.lib()
    variants = {"nan_x": [(7, (np.nan, 3.0))], "nan_y": [(7, (3.0, np.nan))], "inf_y": [(11, (2.0, np.inf))],
                "neg_inf_x": [(5, (-np.inf, 1.0))], "nan_both": [(3, (np.nan, np.nan))],
                "mixed": [(7, (np.nan, 3.0)), (11, (2.0, np.inf)), (13, (np.nan, 5.0)), (2, (4.0, np.nan))]}
    for name, edits in variants.items():
        for seed in range(40):
            rng = np.random.default_rng(seed)
            n_feat, n = 60, 300
            xyI = rng.integers(0, 9, (n_feat, 2)).astype(np.float32)
            for idx, val in edits:
                xyI[idx] = val
            xyJ = rng.integers(0, 9, (n_feat, 2)).astype(np.float32)
            m = np.zeros(n, r3dlib.indmatch_dtype)
            m["i"] = rng.integers(0, n_feat, n)
            m["j"] = rng.integers(0, n_feat, n)
            exp = np.unique(np.stack([m["i"], m["j"]], 1), axis=0)
            e = np.zeros(len(exp), r3dlib.indmatch_dtype)
            e["i"], e["j"] = exp[:, 0], exp[:, 1]
            want = oracle.coord_dedup(e, xyI, xyJ)
            got = m.copy()
            k = lib.r3d_debug_post_process_ranked(got.ctypes.data_as(C.c_void_p), C.c_int64(n),
                                                  xyI.ctypes.data_as(C.c_void_p), C.c_uint32(n_feat),
                                                  xyJ.ctypes.data_as(C.c_void_p))
            assert k == len(want) and np.array_equal(got[:k], want), (name, seed, "ranked")
            got = m.copy()
            k = lib.r3d_debug_post_process(got.ctypes.data_as(C.c_void_p), C.c_int64(n), xyI.ctypes.data_as(C.c_void_p),
                                           xyJ.ctypes.data_as(C.c_void_p), C.c_int(1))
            assert k == len(want) and np.array_equal(got[:k], want), (name, seed, "classic")


def test_device_sample_stream_matches_host_random(r3dlib):
    """The ACRANSAC sample stream drawn on the device (mt19937 + libstdc++'s uniform_int_distribution restated in
    acransac_rng.cuh) is checked against this process's <random>: 400 rounds over pool sizes 8 ... 2^32."""
    assert r3dlib.lib().r3d_debug_rng_selftest() == 1
