"""File formats and the C-ABI surface; no GPU compute calls."""
import ctypes
import os
import re

import numpy as np

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
    assert lib.r3d_abi_version() == 2


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
