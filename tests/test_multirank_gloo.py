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


def _gather_worker(rank, world, port, out_dir, mode):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as po
    from regard3d_b200 import capi
    sc = synth.make_scene(6, 300, 32, "msurf", seed=78)
    pairs = synth.exhaustive_pairs(6)
    counts = [len(d) for d in sc["descs"]]
    mine, _ = sharding.my_shard(pairs, counts, rank, world)
    ofs, m = po.match_pairs(sc["descs"], sc["xys"], mine, 0.8, n_threads=2)
    handle = capi.Matches.from_csr(mine, ofs, m)        # the rank's PairWiseMatches (host container of the C ABI)
    g = sharding.Gather(rank, world, "cpu", mode=mode)
    for _ in range(2):                                  # twice: buffers are reused
        g(handle)
    assert g.ms.get("mode") == mode
    if rank == 0:
        parts = g.result()
        np.savez(os.path.join(out_dir, "g.npz"), pairs=np.concatenate([p[0] for p in parts]),
                 counts=np.concatenate([np.diff(p[1].astype(np.int64)) for p in parts]),
                 m=np.concatenate([p[2] for p in parts]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,mode", [(2, "shm"), (3, "shm"), (2, "p2p"), (3, "p2p")])
def test_gather_to_rank0_in_pair_order(r3dlib, oracle, tmp_path, world, mode):
    """The bench's strong-scaling gather (sharding.Gather: export CSR -> shared segment, or send/recv -> rank 0's buffer) returns the
    single-process PairWiseMatches: same pairs in std::map order, same match sequences."""
    mp = pytest.importorskip("torch.multiprocessing")
    mp.spawn(_gather_worker, args=(world, _free_port(), str(tmp_path), mode), nprocs=world, join=True)
    got = np.load(tmp_path / "g.npz")
    sc = synth.make_scene(6, 300, 32, "msurf", seed=78)
    pairs = synth.exhaustive_pairs(6)
    ofs, m = oracle.match_pairs(sc["descs"], sc["xys"], pairs, 0.8)
    cnt = np.diff(ofs.astype(np.int64))
    keep = cnt > 0                                       # empty pairs are absent from the map
    assert np.array_equal(got["pairs"], pairs[keep]) and np.array_equal(got["counts"], cnt[keep])
    assert np.array_equal(got["m"], m)


def test_partition_ba_tiles_points_and_observations():
    prob = synth.make_ba_problem(n_cams=6, n_pts=301, obs_per_pt=3, seed=5)
    for world in (1, 2, 3, 5):
        seen_pts, n_obs = [], 0
        for r in range(world):
            loc, (p0, p1) = sharding.partition_ba(prob, r, world)
            seen_pts.append((p0, p1))
            n_obs += len(loc["obs_xy"])
            assert loc["obs_pt"].max(initial=0) < max(p1 - p0, 1)
            assert np.array_equal(loc["poses"], prob["poses"]) and len(loc["cam_intr"]) == 6
        assert seen_pts[0][0] == 0 and seen_pts[-1][1] == 301
        assert all(seen_pts[k][1] == seen_pts[k + 1][0] for k in range(world - 1))
        assert n_obs == len(prob["obs_xy"])


def _ba_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as po
    prob = synth.make_ba_problem(n_cams=6, n_pts=301, obs_per_pt=3, seed=5)
    loc, _ = sharding.partition_ba(prob, rank, world)
    # the quantity the ranks exchange is a plain sum over observations: 0.5 * sum |r|^2 here
    res = po.ba_residuals(po.ba_prepare(**{k: loc[k] for k in ("poses", "intrinsics", "points", "obs_cam", "obs_pt", "cam_intr", "obs_xy")}))
    t = torch.tensor([0.5 * float((res ** 2).sum())], dtype=torch.float64)
    dist.all_reduce(t)
    if rank == 0:
        np.save(os.path.join(out_dir, "cost.npy"), t.numpy())
    dist.destroy_process_group()


def test_two_gloo_ranks_sum_partition_costs(oracle, tmp_path):
    mp = pytest.importorskip("torch.multiprocessing")
    mp.spawn(_ba_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    prob = synth.make_ba_problem(n_cams=6, n_pts=301, obs_per_pt=3, seed=5)
    full = oracle.ba_residuals(oracle.ba_prepare(**{k: prob[k] for k in ("poses", "intrinsics", "points", "obs_cam", "obs_pt", "cam_intr", "obs_xy")}))
    assert np.allclose(np.load(tmp_path / "cost.npy")[0], 0.5 * (full ** 2).sum(), rtol=1e-13)
