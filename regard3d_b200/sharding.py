"""Pair sharding across GPUs: image pairs are independent units, every GPU holds all
regions, the I-sorted pair list is cut into contiguous cost-balanced ranges (cost of a pair =
N_I * N_J), results are concatenated in pair order.  No data-path collective.  Same rule as the
in-process multi-device split of r3d_match_pairs (regard3d_b200/csrc/match_host.cu)."""
import os
import sys
import time

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
    # camera-side options are replicated like the cameras (the library counts prior blocks on rank 0 only)
    for k in ("intr_model", "intrinsics_ext", "prior_cam", "prior_center", "prior_weight"):
        if prob.get(k) is not None:
            local[k] = np.array(prob[k]).copy()
    return local, (p0, p1)


class Gather:
    """Per-rank PairWiseMatches -> rank 0, in pair order (shards are contiguous ranges of the I-sorted pair list, so
    rank order = pair order; the reference inserts into one std::map, src/R3DComputeMatches.cpp:483-486).
    Wire format per rank, 8-byte words: P pair ids (2 x u32) | P + 1 prefix offsets (u64) | T matches (2 x u32).
    The matches are HOST data on every rank (the coordinate de-duplication is a host step), so on one node the cheapest
    route is host to host: mode "shm" -- rank 0 owns a POSIX shared-memory segment (kept across calls, so its pages
    are faulted in once), every rank's r3d_matches_export_csr writes its slice of it directly (a multi-threaded
    memcpy), one barrier, done.  Mode "p2p" (R3D_GATHER=p2p, or when /dev/shm is unusable) stages through the GPUs:
    export into pinned memory -> H2D -> NCCL send/recv over NVLink -> D2H into rank 0's pinned buffer; with a CPU
    device (gloo) the same code runs without the staging copies.  The CPU tests exercise both modes."""

    def __init__(self, rank, world, device, mode=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.rank, self.world = torch, dist, rank, world
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.h_send = self.d_send = self.h_result = None
        self.d_recv = [None] * world
        self.sizes = None
        self.words = None
        self.h2d = self.d2h = 0
        self.mode = mode or os.environ.get("R3D_GATHER", "shm")
        self.shm = None          # np.memmap over the shared segment
        self.shm_cap = 0         # its capacity in words (identical on every rank)
        self.shm_gen = 0
        self.ms = {}

    def _grow(self, t, n, **kw):
        if t is None or t.numel() < n:
            return self.torch.empty(int(n * 1.25) + 1024, dtype=self.torch.int64, **kw)
        return t

    def _wait(self, ops):
        for req in (self.dist.batch_isend_irecv(ops) if ops else []):
            req.wait()

    def _shm_ensure(self, total):
        """(Re)create the shared segment when it is too small; every rank takes the same decision from the same sizes."""
        if self.shm is not None and self.shm_cap >= total:
            return True
        cap = int(total * 1.25) + 4096
        self.shm_gen += 1
        name = [None]
        if self.rank == 0:
            name[0] = "/dev/shm/r3d_gather_%d_%d_%d" % (os.getpid(), self.shm_gen, int(time.time() * 1e3) & 0xffffff)
        self.dist.broadcast_object_list(name, src=0)
        path = name[0]
        ok = self.torch.ones(1, dtype=self.torch.int64, device=self.device)
        self.shm = None
        if self.rank == 0:
            try:
                self.shm = np.memmap(path, dtype=np.int64, mode="w+", shape=(cap,))
            except (OSError, ValueError) as e:
                sys.stderr.write("[sharding.Gather] cannot create %s (%d words): %r\n" % (path, cap, e))
                ok[0] = 0
        self.dist.all_reduce(ok, op=self.dist.ReduceOp.MIN)  # also orders "created" before "opened"
        if int(ok.item()) == 1 and self.rank != 0:
            try:
                self.shm = np.memmap(path, dtype=np.int64, mode="r+", shape=(cap,))
            except (OSError, ValueError) as e:
                sys.stderr.write("[sharding.Gather] rank %d cannot map %s: %r\n" % (self.rank, path, e))
                ok[0] = 0
        self.dist.all_reduce(ok, op=self.dist.ReduceOp.MIN)
        if self.rank == 0:
            try:
                os.unlink(path)  # the mappings keep the segment alive; nothing is left behind when the job ends
            except OSError:
                pass
        if int(ok.item()) != 1:
            self.shm, self.shm_cap, self.mode = None, 0, "p2p"
            return False
        self.shm_cap = cap
        return True

    def __call__(self, m):
        torch, dist = self.torch, self.dist
        t0 = time.perf_counter()
        P, T = m.num_pairs, m.total
        mine = torch.tensor([P, T], dtype=torch.int64, device=self.device)
        lst = [torch.zeros(2, dtype=torch.int64, device=self.device) for _ in range(self.world)]
        dist.all_gather(lst, mine)
        self.sizes = np.stack([t.cpu().numpy() for t in lst])
        words = self.sizes[:, 0] * 2 + 1 + self.sizes[:, 1]
        self.words = words
        self.h2d = self.d2h = 0
        t1 = time.perf_counter()

        def export(buf):
            m.export_csr(buf[0:P].view(np.uint32), buf[P:2 * P + 1].view(np.uint64), buf[2 * P + 1:2 * P + 1 + T])
        if self.mode == "shm" and self._shm_ensure(int(words.sum())):
            off = int(words[:self.rank].sum())
            export(self.shm[off:off + int(words[self.rank])])
            t2 = time.perf_counter()
            dist.barrier()
            self.ms = {"mode": "shm", "sizes": 1e3 * (t1 - t0), "export": 1e3 * (t2 - t1), "barrier": 1e3 * (time.perf_counter() - t2)}
            return
        if self.rank == 0:
            self.h_result = self._grow(self.h_result, int(words.sum()), pin_memory=self.cuda)
            export(self.h_result.numpy())
            t2 = time.perf_counter()
            ops, off = [], int(words[0])
            for r in range(1, self.world):
                w = int(words[r])
                if self.cuda:
                    self.d_recv[r] = self._grow(self.d_recv[r], w, device=self.device)
                    ops.append(dist.P2POp(dist.irecv, self.d_recv[r][:w], r))
                else:
                    ops.append(dist.P2POp(dist.irecv, self.h_result[off:off + w], r))
                off += w
            self._wait(ops)
            t3 = time.perf_counter()
            if self.cuda:
                off = int(words[0])
                for r in range(1, self.world):
                    w = int(words[r])
                    self.h_result[off:off + w].copy_(self.d_recv[r][:w], non_blocking=True)
                    off += w
                    self.d2h += 8 * w
                torch.cuda.current_stream().synchronize()
            self.ms = {"mode": "p2p", "sizes": 1e3 * (t1 - t0), "export": 1e3 * (t2 - t1), "recv": 1e3 * (t3 - t2),
                       "d2h": 1e3 * (time.perf_counter() - t3)}
        else:
            w = int(words[self.rank])
            self.h_send = self._grow(self.h_send, w, pin_memory=self.cuda)
            export(self.h_send.numpy())
            src = self.h_send
            if self.cuda:
                self.d_send = self._grow(self.d_send, w, device=self.device)
                self.d_send[:w].copy_(self.h_send[:w], non_blocking=True)
                self.h2d += 8 * w
                src = self.d_send
            self._wait([dist.P2POp(dist.isend, src[:w], 0)])
            if self.cuda:
                torch.cuda.current_stream().synchronize()
            self.ms = {"mode": "p2p"}

    def result(self):
        """Rank 0: [(pairs[P,2] u32, ofs[P+1] u64, matches[T]) per rank], views into the gathered host buffer."""
        from regard3d_b200 import capi
        out, off = [], 0
        buf = self.shm if self.ms.get("mode") == "shm" else self.h_result.numpy()
        for r in range(self.world):
            w, P, T = int(self.words[r]), int(self.sizes[r, 0]), int(self.sizes[r, 1])
            b = buf[off:off + w]
            out.append((b[0:P].view(np.uint32).reshape(-1, 2), b[P:2 * P + 1].view(np.uint64),
                        b[2 * P + 1:2 * P + 1 + T].view(capi.indmatch_dtype)))
            off += w
        return out
