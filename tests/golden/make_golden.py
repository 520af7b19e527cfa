#!/usr/bin/env python
"""Generates tests/golden/oracle_v1.json from the CPU oracle (the reference ships no golden vectors and cannot be built
here -- SURVEY.md 8c -- so these pin the ORACLE itself against drift and give the GPU path a committed target).

    python tests/golden/make_golden.py          # rewrites oracle_v1.json

Content: for a few seeded synthetic scenes (regard3d_b200/synth.py) the SHA-1 of every pair's (i, j) sequence after
putative matching, after the F / E / H a-contrario filters and of the cascade-hashing matcher (ratio 0.8), and the cost
trace of a seeded bundle adjustment."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

SCENES = [
    dict(name="liop144", n_img=4, n_feat=1200, dim=144, kind="liop", seed=101, ratio=0.6, u8=False),
    dict(name="msurf64", n_img=3, n_feat=1500, dim=64, kind="msurf", seed=102, ratio=0.8, u8=False),
    dict(name="sift128_u8", n_img=3, n_feat=1000, dim=128, kind="sift", seed=103, ratio=0.8, u8=True),
]
BA = dict(n_cams=8, n_pts=400, obs_per_pt=4, seed=104, iters=12)


def seq_hash(m):
    a = np.stack([np.asarray(m["i"], np.uint32), np.asarray(m["j"], np.uint32)], 1) if len(m) else np.zeros((0, 2), np.uint32)
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def per_pair(pairs, ofs, m):
    return [[int(I), int(J), int(ofs[k + 1] - ofs[k]), seq_hash(m[int(ofs[k]):int(ofs[k + 1])])] for k, (I, J) in enumerate(pairs)]


def scene_record(po, synth, sc_def):
    sc = synth.make_scene(sc_def["n_img"], sc_def["n_feat"], sc_def["dim"], sc_def["kind"], seed=sc_def["seed"], as_u8=sc_def["u8"])
    pairs = synth.exhaustive_pairs(sc_def["n_img"])
    ofs, m = po.match_pairs(sc["descs"], sc["xys"], pairs, sc_def["ratio"])
    Ks = np.array([[1.1 * max(int(w), int(h)), w / 2.0, h / 2.0] for w, h in zip(sc["widths"], sc["heights"])])
    fo, fm = po.filter_pairs_F(sc["xys"], sc["widths"], sc["heights"], pairs, ofs, m)
    eo, em = po.filter_pairs_E(sc["xys"], sc["widths"], sc["heights"], Ks, pairs, ofs, m)
    ho, hm = po.filter_pairs_H(sc["xys"], sc["widths"], sc["heights"], pairs, ofs, m)
    co, cm = po.cascade_match_pairs(sc["descs"], sc["xys"], pairs, 0.8)   # cascade hashing (BASELINE config 4's matcher)
    return {"def": sc_def, "putative": per_pair(pairs, ofs, m), "F": per_pair(pairs, fo, fm), "E": per_pair(pairs, eo, em),
            "H": per_pair(pairs, ho, hm), "cascade_r0.8": per_pair(pairs, co, cm)}


def ba_record(po, synth):
    prob = synth.make_ba_problem(n_cams=BA["n_cams"], n_pts=BA["n_pts"], obs_per_pt=BA["obs_per_pt"], seed=BA["seed"], outlier_frac=0.02)
    keys = ("poses", "intrinsics", "points", "obs_cam", "obs_pt", "cam_intr", "obs_xy")
    a = po.ba_prepare(*[prob[k] for k in keys])
    s, trace = po.bundle_adjust(a, po.default_ba_options(max_iterations=BA["iters"]))
    return {"def": BA, "iterations": int(s["iterations"]), "successful_steps": int(s["successful_steps"]),
            "termination": int(s["termination"]), "cost_trace": [float(x) for x in trace]}


def build_all():
    from oracle import pyoracle as po
    from regard3d_b200 import synth
    po.build()
    return {"version": 1, "scenes": [scene_record(po, synth, d) for d in SCENES], "ba": ba_record(po, synth)}


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_v1.json")
    json.dump(build_all(), open(out, "w"), indent=1)
    print("wrote", out)
