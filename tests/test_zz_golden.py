"""Committed golden fixtures (tests/golden/oracle_v1.json, made by tests/golden/make_golden.py).  The reference has no
golden vectors, so these are oracle outputs: the CPU test pins the oracle against drift, the GPU test
holds the C-ABI path to the same committed hashes (putative matches and F / E / H inlier sequences, BA cost trace)."""
import importlib.util
import json
import os

import numpy as np
import pytest

from regard3d_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "oracle_v1.json")))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)


def test_oracle_reproduces_the_golden_fixtures(oracle):
    fresh = make_golden.build_all()
    assert fresh["scenes"] == GOLD["scenes"]                      # every pair: count + SHA-1 of the (i, j) sequence
    for k in ("iterations", "successful_steps", "termination"):
        assert fresh["ba"][k] == GOLD["ba"][k]
    assert np.allclose(fresh["ba"]["cost_trace"], GOLD["ba"]["cost_trace"], rtol=1e-9, atol=0)


def _rows(d, pairs):
    """[I, J, count, sha1] rows, like make_golden.per_pair, from a Matches.to_dict() result."""
    out = []
    for I, J in pairs:
        m = d.get((int(I), int(J)))
        if m is None:
            m = np.zeros(0, dtype=[("i", np.uint32), ("j", np.uint32)])
        out.append([int(I), int(J), int(len(m)), make_golden.seq_hash(m)])
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("scene", GOLD["scenes"], ids=[s["def"]["name"] for s in GOLD["scenes"]])
def test_gpu_path_reproduces_the_golden_fixtures(scene, gpu_ctx, r3dlib):
    d = scene["def"]
    sc = synth.make_scene(d["n_img"], d["n_feat"], d["dim"], d["kind"], seed=d["seed"], as_u8=d["u8"])
    pairs = synth.exhaustive_pairs(d["n_img"])
    gpu_ctx.clear_regions()
    for v in range(d["n_img"]):
        gpu_ctx.upload_regions(v, sc["descs"][v], sc["xys"][v])
    put = gpu_ctx.match_pairs(pairs, d["ratio"])
    assert _rows(put.to_dict(), pairs) == scene["putative"]
    Ks = np.array([[1.1 * max(int(w), int(h)), w / 2.0, h / 2.0] for w, h in zip(sc["widths"], sc["heights"])])
    for name, model in (("F", r3dlib.MODEL_F), ("E", r3dlib.MODEL_E), ("H", r3dlib.MODEL_H)):
        got = gpu_ctx.filter_pairs(put, sc["widths"], sc["heights"], model=model, Ks=Ks)
        assert _rows(got.to_dict(), pairs) == scene[name], name
    cas = gpu_ctx.match_pairs(pairs, 0.8, r3dlib.MATCH_CASCADE_HASHING)
    assert _rows(cas.to_dict(), pairs) == scene["cascade_r0.8"]


@pytest.mark.gpu
def test_gpu_ba_reproduces_the_golden_cost_trace(gpu_ctx, oracle):
    b = GOLD["ba"]["def"]
    prob = synth.make_ba_problem(n_cams=b["n_cams"], n_pts=b["n_pts"], obs_per_pt=b["obs_per_pt"], seed=b["seed"], outlier_frac=0.02)
    keys = ("poses", "intrinsics", "points", "obs_cam", "obs_pt", "cam_intr", "obs_xy")
    p = oracle.ba_prepare(*[prob[k] for k in keys])
    s, trace = gpu_ctx.bundle_adjust(p, max_iterations=b["iters"])
    assert int(s["iterations"]) == GOLD["ba"]["iterations"] and int(s["successful_steps"]) == GOLD["ba"]["successful_steps"]
    assert int(s["termination"]) == GOLD["ba"]["termination"]
    assert np.allclose(trace, GOLD["ba"]["cost_trace"], rtol=1e-8, atol=0)
