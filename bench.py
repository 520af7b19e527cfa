#!/usr/bin/env python
"""bench.py -- matched image-pairs / second on the compute-matches hot path (BASELINE.json metric).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload at every N: BASELINE.json configs[1] (C2) PER GPU -- 50 synthetic 1080p images x 10 000
float descriptors (D = 144, Regard3D's native R3D_AKAZE_LIOP_Regions layout), exhaustive pairs
(1 225), brute-force L2 2-NN + ratio 0.6 + de-duplications.  One step = one pass over all pairs.
N > 1: one process per GPU, every rank owns an independent 50-image set (weak scaling; image pairs
shard with no data-path collective); torch.distributed (NCCL) is only used for the
barrier and the max-over-ranks of the step time.

value : pairs/s with descriptors already resident in HBM (r3d_match_pairs only; results land in
        host memory, host de-duplication included).
e2e   : pairs/s through the C ABI from pinned HOST buffers: r3d_clear_regions + r3d_upload_regions
        of every view + r3d_match_pairs, every step.
roofline : the tcgen05 candidate kernel, algorithmic 2*N_I*N_J*D flop per pair over
        its CUDA-event time on its own stream, against MEASURED_PEAKS.json bf16 TFLOP/s.
cpu_baseline : the oracle port (same serial-I / omp-J structure as the reference) on a bounded
        sample of the same pairs, all host threads it can use.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_IMAGES, N_FEATS, DIM, KIND, RATIO = 50, 10000, 144, "liop", 0.6
WORKLOAD, AS_U8 = "C2", False
# BASELINE.json configs; C2 is the bench line (the metric's single-GPU configuration), the others are
# selectable for the record: `--workload c3|c4` (C4 = exact GPU matcher in place of CPU cascade hashing, + F filter)
WORKLOADS = {
    "c2": dict(images=50, feats=10000, dim=144, kind="liop", u8=False, name="C2"),
    "c2-msurf64": dict(images=50, feats=10000, dim=64, kind="msurf", u8=False, name="C2 (MSURF-64)"),
    "c3": dict(images=200, feats=20000, dim=128, kind="sift", u8=True, name="C3"),
    "c4": dict(images=500, feats=10000, dim=128, kind="sift", u8=True, name="C4 (exact matcher + F filter)"),
}
METRIC = "matched_image_pairs_per_sec_exhaustive"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        # "under load": samples in the upper half of what was seen
        if not sm:  # the region was shorter than one sampling period: one synchronous reading
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,power.draw",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
                f = [x.strip() for x in out.strip().split(",")]
                sm.append(float(f[0])); smax.append(float(f[1])); power.append(float(f[2]))
            except Exception:
                pass
        if sm:
            hi = [x for x in sm if x >= 0.5 * max(sm)]
            med = float(np.median(hi))
        else:
            med = None
        return {"sm_mhz": med, "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


def effective_cpus():
    """Host CPUs this process may actually use: the affinity mask, cut by the cgroup CPU-time quota (the GPU boxes
    give a container ~16 CPUs of quota per GPU although 128 hardware threads are visible)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            n = min(n, max(1, int(round(int(q) / int(per)))))
    except Exception:
        pass
    return n


def measured_traffic_per_pair():
    """DRAM bytes per image pair of the candidate kernel, from the committed `ncu --set full` capture of one
    128-pair launch (profiles/r01_k_l2_candidates_2sm_keymetrics.csv); None when the file is absent."""
    p = os.path.join(ROOT, "profiles", "r01_k_l2_candidates_2sm_keymetrics.csv")
    if not os.path.exists(p):
        return None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for ln in open(p):
        f = ln.strip().split(",")
        if len(f) >= 4 and f[1] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(f[2]) * unit.get(f[3], 1.0)
    return tot / 128.0 if tot else None


def make_workload(seed):
    from regard3d_b200 import synth
    sc = synth.make_scene(N_IMAGES, N_FEATS, DIM, KIND, seed=seed, as_u8=AS_U8)
    pairs = synth.exhaustive_pairs(N_IMAGES)
    return sc, pairs


def run_reference(args, rank, world, emit):
    """--impl reference: the reference's CPU path = the oracle port (the reference's own code cannot
    be built here: OpenMVG/Ceres/Eigen/wx are neither vendored nor installed)."""
    if rank != 0:
        return
    from oracle import pyoracle as po
    sc, pairs = make_workload(20260924 + 2)
    nthreads = po.num_threads()
    # bounded sample: all pairs that share the first image(s) -- the reference parallelises over J
    # for a fixed I (src/R3DComputeMatches.cpp:465), so one I gives up to 49 concurrent J's.
    n_sample = int(os.environ.get("R3D_REF_SAMPLE_PAIRS", "49"))
    sample = pairs[:n_sample]
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        po.match_pairs(sc["descs"], sc["xys"], sample, RATIO, n_threads=nthreads)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = len(sample) * len(times) / total
    cores = min(nthreads, len(sample), effective_cpus())
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": "%d pairs (I=0, J=1..%d) of the set per step, omp over J, %d threads on %d usable CPUs "
                                   "(cgroup quota)" % (len(sample), len(sample), nthreads, effective_cpus())},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config():
    return {"workload": "%s: %d images x %d feats, D=%d %s (%s-like), exhaustive %d pairs, ratio %.1f"
                        % (WORKLOAD, N_IMAGES, N_FEATS, DIM, "uint8" if AS_U8 else "float32", KIND,
                           N_IMAGES * (N_IMAGES - 1) // 2, RATIO),
            "images": N_IMAGES, "feats_per_image": N_FEATS, "dim": DIM, "pairs": N_IMAGES * (N_IMAGES - 1) // 2,
            "parallelism": "pairs sharded per GPU, no collective",
            "l2_policy": "inputs (descriptors + fp16 operands, %.1f GB) exceed the 126 MB L2"
                         % (N_IMAGES * N_FEATS * (DIM * (1 if AS_U8 else 4) + 2 * 2 * (DIM + 48)) / 1e9)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ba", action="store_true", help="skip the bundle-adjustment leg")
    ap.add_argument("--no-filter", action="store_true", help="skip the F-filter leg")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="BASELINE config (default c2 = the bench line)")
    ap.add_argument("--dim", type=int, default=0, help="experiment only: 64 -> MSURF-like D=64 set")
    ap.add_argument("--feats", type=int, default=0, help="experiment only")
    ap.add_argument("--images", type=int, default=0, help="experiment only")
    args = ap.parse_args()
    # stdout carries exactly ONE line (the JSON): anything a library prints there (NCCL's version banner
    # at communicator creation, ...) is sent to stderr instead
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())
    global DIM, KIND, N_FEATS, N_IMAGES, WORKLOAD, AS_U8
    wl = WORKLOADS[args.workload]
    N_IMAGES, N_FEATS, DIM, KIND, AS_U8, WORKLOAD = wl["images"], wl["feats"], wl["dim"], wl["kind"], wl["u8"], wl["name"]
    if args.workload != "c2":
        os.environ.setdefault("R3D_REF_SAMPLE_PAIRS", "16")
    if args.dim == 64:
        DIM, KIND = 64, "msurf"
    elif args.dim == 128:
        DIM, KIND = 128, "sift"
    if args.feats:
        N_FEATS = args.feats
    if args.images:
        N_IMAGES = args.images
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world, emit)
        return 0

    import torch
    import torch.distributed as dist
    from regard3d_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")     # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = capi.Context((local_rank,))
    sc, pairs = make_workload(20260924 + 2 + 1000 * rank)
    n_pairs = len(pairs)
    # pinned host staging (the e2e leg copies from here every step)
    pinned_desc, pinned_xy = [], []
    for v in range(N_IMAGES):
        d = torch.from_numpy(sc["descs"][v]).pin_memory()
        x = torch.from_numpy(sc["xys"][v]).pin_memory()
        pinned_desc.append(d)
        pinned_xy.append(x)

    def upload_all():
        for v in range(N_IMAGES):
            ctx.upload_regions(v, pinned_desc[v].numpy(), pinned_xy[v].numpy())

    # ---------------- resident leg: `value` ----------------
    upload_all()
    for _ in range(max(args.warmup, 3)):
        m = ctx.match_pairs(pairs, RATIO)
    sampler = ClockSampler(local_rank)
    cand_ms, rerank_ms, fb_ms, dev_ms, host_ms, launches = [], [], [], [], [], 0
    fbq = q = rejq = 0
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m = ctx.match_pairs(pairs, RATIO)
        t = ctx.match_timing()
        cand_ms.append(t["ms_candidates"]); rerank_ms.append(t["ms_rerank"]); fb_ms.append(t["ms_fallback"])
        dev_ms.append(t["ms_device_total"]); host_ms.append(t["ms_host_post"])
        launches += t["kernel_launches"]
        fbq += t["fallback_queries"]; q += t["queries"]; rejq += t["rejected_queries"]
        d2h_step = t["d2h_bytes"]
    barrier()
    t_res = time.perf_counter() - t0
    n_matches = m.total
    n_match_pairs = m.num_pairs

    # ---------------- end-to-end leg: `e2e` ----------------
    for _ in range(2):
        ctx.clear_regions(); upload_all(); ctx.match_pairs(pairs, RATIO)
    barrier()
    t0 = time.perf_counter()
    h2d_step = 0
    for _ in range(args.steps):
        ctx.clear_regions()
        upload_all()
        m2 = ctx.match_pairs(pairs, RATIO)
        t = ctx.match_timing()
        h2d_step = t["h2d_bytes"]
        d2h_e2e = t["d2h_bytes"]
        launches_e2e = t["kernel_launches"]
    barrier()
    t_e2e = time.perf_counter() - t0
    clocks = sampler.stop()   # sampled over both timed regions (resident + end-to-end), all of it under load

    # ---------------- geometric filter leg (reported alongside; BASELINE C2 itself stops at putatives) ----------------
    filt = None
    if not args.no_filter:
        ctx.filter_pairs(m2, sc["widths"], sc["heights"])            # warm-up
        torch.cuda.synchronize()
        tf0 = time.perf_counter()
        fm = ctx.filter_pairs(m2, sc["widths"], sc["heights"])
        tf = time.perf_counter() - tf0
        ft = ctx.filter_timing()
        filt = {"pairs_per_s": m2.num_pairs / tf, "ms": 1e3 * tf, "pairs_in": m2.num_pairs, "pairs_kept": fm.num_pairs,
                "inliers": fm.total, "hypotheses": int(ft["hypotheses"]), "rounds": int(ft["rounds"]),
                "ms_solve": ft["ms_solve"], "ms_score": ft["ms_score"], "ms_host": ft["ms_host"],
                "kernel_launches": int(ft["kernel_launches"]),
                "what": "AC-RANSAC fundamental filter (4 px, 2048 it.) over all putative pairs of the step"}

    # ---------------- bundle-adjustment leg (BASELINE C5, reported alongside) ----------------
    # N > 1: STRONG scaling of the one C5 problem -- points (+ their observations) partitioned over the
    # ranks, cameras replicated, one in-library ncclAllReduce of the reduced camera system per LM iteration.
    ba = None
    if not args.no_ba:
        from regard3d_b200 import sharding, synth
        prob = synth.make_ba_problem(n_cams=200, n_pts=200000, obs_per_pt=5, seed=20260924 + 5)
        arrs = {k: np.ascontiguousarray(v) for k, v in prob.items() if k != "truth"}
        for k in ("poses", "intrinsics", "points", "obs_xy"):
            arrs[k] = np.ascontiguousarray(arrs[k], np.float64)
        for k in ("obs_cam", "obs_pt", "cam_intr"):
            arrs[k] = np.ascontiguousarray(arrs[k], np.uint32)
        if world > 1:
            ids = [ctx.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            ctx.comm_init(world, rank, ids[0])
        local, _ = sharding.partition_ba(arrs, rank, world)
        ctx.bundle_adjust({k: v.copy() for k, v in local.items()}, max_iterations=2)     # warm-up
        n_it = 10
        g = {k: v.copy() for k, v in local.items()}
        barrier()
        tb0 = time.perf_counter()
        sg, tg = ctx.bundle_adjust(g, max_iterations=n_it, function_tolerance=0.0)
        tb = time.perf_counter() - tb0
        t_loop = max(sg["seconds_total"] - sg["seconds_setup"], 1e-9)
        if world > 1:
            tt = torch.tensor([t_loop, tb], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_loop, tb = float(tt[0]), float(tt[1])
            ctx.comm_destroy()
        n_obs = int(len(arrs["obs_xy"]))
        nB = 6 * 200 + 6
        bytes_iter = 3 * (n_obs * 24 + len(arrs["points"]) * 24) + 2 * nB * nB * 8
        ba = {"iters_per_s": sg["iterations"] / t_loop, "e2e_iters_per_s": sg["iterations"] / tb,
              "iterations": int(sg["iterations"]), "seconds_lm_loop": t_loop, "seconds_call": tb,
              "seconds_setup": sg["seconds_setup"], "seconds_linear": sg["seconds_linear"],
              "initial_cost": sg["initial_cost"], "final_cost": sg["final_cost"], "n_gpus": world,
              "scaling": "strong", "exchange": "none" if world == 1 else
              "ncclAllReduce(f64) of S|rhs = %d doubles per LM iteration + 5 scalar/vector reductions" % (nB * nB + nB),
              "config": "C5: 200 cams / %d pts / %d obs, 1 shared radial-K3 intrinsic, Huber(16)" % (len(arrs["points"]), n_obs),
              "roofline": {"bound": "hbm", "achieved": sg["iterations"] / t_loop * bytes_iter / 1e9,
                           "peak": float(load_peaks()[0].get("hbm_gbs", 6650.0)), "unit": "GB/s",
                           "frac": sg["iterations"] / t_loop * bytes_iter / 1e9 / float(load_peaks()[0].get("hbm_gbs", 6650.0)),
                           "bytes_per_iter": bytes_iter}}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import pyoracle as po
            c = po.ba_prepare(arrs["poses"], arrs["intrinsics"], arrs["points"], arrs["obs_cam"], arrs["obs_pt"],
                              arrs["cam_intr"], arrs["obs_xy"])
            o = po.default_ba_options(max_iterations=3)
            o.function_tolerance = 0.0
            tc0 = time.perf_counter()
            so, to = po.bundle_adjust(c, o)
            tcb = time.perf_counter() - tc0
            ba["cpu_baseline"] = {"iters_per_s": so["iterations"] / tcb, "iterations": int(so["iterations"]),
                                  "cores": min(po.num_threads(), effective_cpus()), "kind": "port",
                                  "cost_trace_rel_diff": float(np.max(np.abs(tg[:len(to)] - to) / to))}

    if world > 1:
        tt = torch.tensor([t_res, t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_res, t_e2e = float(tt[0]), float(tt[1])

    if rank == 0:
        peaks, peak_src = load_peaks()
        total_pairs = n_pairs * world * args.steps
        value = total_pairs / t_res
        e2e = total_pairs / t_e2e
        flop_per_launch = 2.0 * N_FEATS * N_FEATS * DIM * n_pairs          # one launch = all pairs of the step
        ms_c = float(np.mean(cand_ms))
        achieved = flop_per_launch / (ms_c * 1e-3) / 1e12
        peak = float(peaks.get("bf16_tflops", 1590.0))
        tpp = measured_traffic_per_pair()
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * t_res / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(), "clocks": clocks,
            "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d_step),
                    "d2h_bytes_per_step": int(d2h_e2e), "ms_per_step": 1e3 * t_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "k_l2_candidates_2sm (tcgen05 kind::f16 cta_group::2, f16 candidates; "
                                                      "every reported distance is re-computed in f32)",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_source": "%s bf16_tflops (burst; kernel timed alone with CUDA events)" % peak_src,
                         "ms_per_launch": ms_c, "flop_per_launch": flop_per_launch,
                         "launch": "one step = all %d pairs (13 batch launches of <= 128 pairs, summed)" % n_pairs,
                         "traffic": (tpp * n_pairs if tpp and DIM == 144 and N_FEATS == 10000 else None),
                         "traffic_unit": "DRAM bytes per step (ncu dram__bytes_read+write of one 128-pair launch, "
                                         "scaled to the step's pairs; profiles/r01_k_l2_candidates_2sm_keymetrics.csv)"},
            "breakdown_ms": {"candidates": ms_c, "rerank": float(np.mean(rerank_ms)),
                             "exact_scan_and_pack": float(np.mean(fb_ms)), "device_total": float(np.mean(dev_ms)),
                             "host_dedup": float(np.mean(host_ms))},
            "result": {"pairs_with_matches": int(n_match_pairs), "matches": int(n_matches),
                       "fallback_query_frac": fbq / max(q, 1), "early_rejected_query_frac": rejq / max(q, 1)},
        }
        if filt is not None:
            line["f_filter"] = filt
        if ba is not None:
            line["ba"] = ba
        if world == 1 and not args.no_cpu_baseline:
            from oracle import pyoracle as po
            nthreads = po.num_threads()
            n_sample = int(os.environ.get("R3D_REF_SAMPLE_PAIRS", "49"))
            sample = pairs[:n_sample]
            tc0 = time.perf_counter()
            o_ofs, o_m = po.match_pairs(sc["descs"], sc["xys"], sample, RATIO, n_threads=nthreads)
            tc = time.perf_counter() - tc0
            # the checker doubles as a parity probe on the sampled pairs
            got = m.to_dict()
            same = True
            for k, (I, J) in enumerate(sample):
                e = o_m[int(o_ofs[k]):int(o_ofs[k + 1])]
                g = got.get((int(I), int(J)))
                same &= (g is not None and len(g) == len(e) and set(zip(g["i"].tolist(), g["j"].tolist())) ==
                         set(zip(e["i"].tolist(), e["j"].tolist()))) or (g is None and len(e) == 0)
            line["cpu_baseline"] = {"value": len(sample) / tc, "unit": "pairs/s",
                                    "cores": min(nthreads, len(sample), effective_cpus()),
                                    "kind": "port",
                                    "sample": "%d pairs (I=0) of the same set, one pass, %d omp threads on %d usable CPUs "
                                              "(cgroup quota), %.1f s" % (len(sample), nthreads, effective_cpus(), tc),
                                    "parity_on_sample": bool(same)}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
