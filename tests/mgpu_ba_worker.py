"""Worker of tests/test_gpu_ba_multi.py (one process per GPU, launched with torch.distributed.run):
point-partitioned bundle adjustment with the in-library ncclAllReduce vs the single-GPU solve and the
CPU oracle on the same problem."""
import json
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from regard3d_b200 import capi, sharding, synth
    out_path, n_cams, n_pts, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = capi.Context((local,))
    ids = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(world, rank, ids[0])
    assert ctx.comm_world == world
    prob = synth.make_ba_problem(n_cams=n_cams, n_pts=n_pts, obs_per_pt=4, seed=21, outlier_frac=0.02)
    loc, (p0, p1) = sharding.partition_ba(prob, rank, world)
    s, trace = ctx.bundle_adjust(loc, max_iterations=iters)
    # every rank must hold the same cameras; gather the point slices on rank 0
    cams = torch.from_numpy(np.concatenate([loc["poses"].ravel(), loc["intrinsics"].ravel()])).cuda()
    lo, hi = cams.clone(), cams.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    cams_identical = bool(torch.equal(lo, hi))
    gathered = [None] * world
    dist.all_gather_object(gathered, (p0, p1, loc["points"]))
    ctx.comm_destroy()
    if rank == 0:
        from oracle import pyoracle as po
        po.build()
        pts = np.zeros_like(np.asarray(prob["points"], np.float64))
        for a, b, x in gathered:
            pts[a:b] = x
        keys = ("poses", "intrinsics", "points", "obs_cam", "obs_pt", "cam_intr", "obs_xy")
        single = po.ba_prepare(*[prob[k] for k in keys])
        ss, ts = ctx.bundle_adjust(single, max_iterations=iters)             # same context, no communicator
        ref = po.ba_prepare(*[prob[k] for k in keys])
        so, to = po.bundle_adjust(ref, po.default_ba_options(max_iterations=iters))
        multi = po.ba_prepare(loc["poses"], loc["intrinsics"], pts, prob["obs_cam"], prob["obs_pt"], prob["cam_intr"], prob["obs_xy"])
        r_multi, r_ref = ctx.ba_residuals(multi), po.ba_residuals(ref)
        scale = np.maximum(np.abs(r_ref), 1e-3 * np.median(np.abs(r_ref)))
        json.dump({
            "world": world, "cams_identical": cams_identical,
            "iterations": [int(s["iterations"]), int(ss["iterations"]), int(so["iterations"])],
            "successful": [int(s["successful_steps"]), int(ss["successful_steps"]), int(so["successful_steps"])],
            "trace_vs_single": float(np.max(np.abs(trace - ts) / ts)),
            "trace_vs_oracle": float(np.max(np.abs(trace - to) / to)),
            "residual_rel_vs_oracle": float((np.abs(r_multi - r_ref) / scale).max()),
            "final_cost": float(s["final_cost"]), "initial_cost": float(s["initial_cost"]),
        }, open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
