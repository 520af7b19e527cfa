"""Pair sharding across GPUs: image pairs are independent units, every GPU holds all
regions, the I-sorted pair list is cut into contiguous cost-balanced ranges (cost of a pair =
N_I * N_J), results are concatenated in pair order.  No data-path collective.  Same rule as the
in-process multi-device split of r3d_match_pairs (regard3d_b200/csrc/match_host.cu)."""
import numpy as np


def shard_bounds(pairs, counts, world):
    """Cut points c[0..world] into the pair list so that rank r owns pairs[c[r]:c[r+1]]."""
    pairs = np.asarray(pairs, np.int64).reshape(-1, 2)
    counts = np.asarray(counts, np.float64)
    cost = counts[pairs[:, 0]] * counts[pairs[:, 1]] + 1.0
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    cuts = [0]
    for k in range(1, world):
        target = cum[-1] * k / world
        cuts.append(int(min(np.searchsorted(cum, target, side="left"), len(pairs))))
    cuts.append(len(pairs))
    return np.maximum.accumulate(np.array(cuts, np.int64))


def my_shard(pairs, counts, rank, world):
    c = shard_bounds(pairs, counts, world)
    pairs = np.asarray(pairs).reshape(-1, 2)
    return pairs[c[rank]:c[rank + 1]], int(c[rank])


def partition_ba(prob, rank, world):
    """Point partition of a bundle-adjustment problem: rank r keeps a contiguous range
    of 3-D points cut so that every rank holds about the same number of observations, with ALL the
    observations of those points; cameras and intrinsics are replicated.  Returns (local problem dict
    for r3d_bundle_adjust, (p0, p1) = the rank's point range).  The ranges of all ranks tile
    [0, n_pts) and the local observations tile the observation set."""
    obs_pt = np.asarray(prob["obs_pt"], np.int64)
    n_pts = len(prob["points"])
    per_pt = np.bincount(obs_pt, minlength=n_pts).astype(np.float64)
    cum = np.concatenate([[0.0], np.cumsum(per_pt)])
    cuts = [0]
    for k in range(1, world):
        cuts.append(int(min(np.searchsorted(cum, cum[-1] * k / world, side="left"), n_pts)))
    cuts.append(n_pts)
    cuts = np.maximum.accumulate(np.array(cuts, np.int64))
    p0, p1 = int(cuts[rank]), int(cuts[rank + 1])
    sel = np.nonzero((obs_pt >= p0) & (obs_pt < p1))[0]
    local = {
        "poses": np.ascontiguousarray(prob["poses"], np.float64).copy(),
        "intrinsics": np.ascontiguousarray(prob["intrinsics"], np.float64).copy(),
        "points": np.ascontiguousarray(np.asarray(prob["points"], np.float64)[p0:p1]).copy(),
        "obs_cam": np.ascontiguousarray(np.asarray(prob["obs_cam"], np.uint32)[sel]),
        "obs_pt": np.ascontiguousarray((obs_pt[sel] - p0).astype(np.uint32)),
        "cam_intr": np.ascontiguousarray(prob["cam_intr"], np.uint32).copy(),
        "obs_xy": np.ascontiguousarray(np.asarray(prob["obs_xy"], np.float64)[sel]),
    }
    return local, (p0, p1)
