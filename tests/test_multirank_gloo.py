"""N > 1 host logic on CPU: two gloo ranks shard one pair list (cost-balanced contiguous ranges, no
data-path collective), each runs the matching stage on its shard (here: the CPU oracle stands in for
the device), rank 0 gathers and the union equals the single-process result."""
import os
import socket

import numpy as np
import pytest

from regard3d_b200 import sharding, synth


def test_shard_bounds_are_contiguous_and_balanced():
    counts = np.array([1000, 5000, 200, 3000, 3000, 800])
    pairs = synth.exhaustive_pairs(len(counts))
    for world in (1, 2, 3, 4, 8):
        c = sharding.shard_bounds(pairs, counts, world)
        assert c[0] == 0 and c[-1] == len(pairs) and (np.diff(c) >= 0).all()
        cost = counts[pairs[:, 0]].astype(float) * counts[pairs[:, 1]]
        per = [cost[c[r]:c[r + 1]].sum() for r in range(world)]
        assert max(per) <= cost.sum() / world + cost.max() + 1e-9      # off by at most one pair's cost


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as po
    sc = synth.make_scene(5, 400, 32, "msurf", seed=77)
    pairs = synth.exhaustive_pairs(5)
    counts = [len(d) for d in sc["descs"]]
    mine, start = sharding.my_shard(pairs, counts, rank, world)
    ofs, m = po.match_pairs(sc["descs"], sc["xys"], mine, 0.8, n_threads=2)
    payload = {"start": start, "pairs": mine, "ofs": ofs, "m": m}
    gathered = [None] * world
    dist.all_gather_object(gathered, payload)          # result gather on the host, not on the data path
    dist.barrier()
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), np.array(gathered, dtype=object), allow_pickle=True)
    dist.destroy_process_group()


def test_two_gloo_ranks_equal_single_process(oracle, tmp_path):
    mp = pytest.importorskip("torch.multiprocessing")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    gathered = np.load(tmp_path / "gathered.npy", allow_pickle=True)
    sc = synth.make_scene(5, 400, 32, "msurf", seed=77)
    pairs = synth.exhaustive_pairs(5)
    ofs, m = oracle.match_pairs(sc["descs"], sc["xys"], pairs, 0.8)
    got_pairs, got_m = [], []
    for g in sorted(gathered, key=lambda g: g["start"]):
        got_pairs.append(g["pairs"])
        got_m.append(g["m"])
    assert np.array_equal(np.concatenate(got_pairs), pairs)            # contiguous shards in pair order
    assert np.array_equal(np.concatenate(got_m), m)                    # concatenation == single-process result
