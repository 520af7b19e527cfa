"""Stand-alone BA run on BASELINE C5 for profiling (ncu launch list / --set full)."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np

from regard3d_b200 import capi, synth

n_cams = int(os.environ.get("BA_CAMS", 200))
n_pts = int(os.environ.get("BA_PTS", 200000))
iters = int(os.environ.get("BA_ITERS", 4))
prob = synth.make_ba_problem(n_cams=n_cams, n_pts=n_pts, obs_per_pt=5, seed=20260924 + 5)
arrs = {}
for k in ("poses", "intrinsics", "points", "obs_xy"):
    arrs[k] = np.ascontiguousarray(prob[k], np.float64)
for k in ("obs_cam", "obs_pt", "cam_intr"):
    arrs[k] = np.ascontiguousarray(prob[k], np.uint32)
ctx = capi.Context((0,))
t0 = time.perf_counter()
s, trace = ctx.bundle_adjust(arrs, max_iterations=iters, function_tolerance=0.0)
print("BA", s, "wall", time.perf_counter() - t0, "trace", trace)
