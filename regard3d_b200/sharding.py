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
