"""-m gpu, needs >= 2 GPUs in one process: r3d_create(device_ids = [0, 1]) -- the drop-in mode for the reference's
single-process GUI -- shards the pair list of r3d_match_pairs and the putative map of r3d_filter_pairs over the
context's devices (contiguous cost-balanced ranges, no collective) and returns what one device returns."""
import numpy as np
import pytest

from conftest import dict_sets, match_sets
from regard3d_b200 import synth

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs in one box")
def test_two_devices_in_one_context_equal_one_device_and_the_oracle(r3dlib, oracle):
    sc = synth.make_scene(7, 1500, 144, "liop", seed=31)
    pairs = synth.exhaustive_pairs(7)
    results = {}
    for devs in ((0,), (0, 1)):
        ctx = r3dlib.Context(devs)
        for v in range(7):
            ctx.upload_regions(v, sc["descs"][v], sc["xys"][v])
        m = ctx.match_pairs(pairs, 0.6)
        f = ctx.filter_pairs(m, sc["widths"], sc["heights"])
        results[devs] = (m.to_dict(), f.to_dict())
        ctx.close()
    one, two = results[(0,)], results[(0, 1)]
    for a, b in zip(one, two):
        assert sorted(a) == sorted(b)
        for k in a:
            assert np.array_equal(a[k], b[k]), k           # same sequences, pair by pair
    ofs, mm = oracle.match_pairs(sc["descs"], sc["xys"], pairs, 0.6)
    assert dict_sets(two[0]) == match_sets(ofs, mm, pairs)
    o_ofs, o_m = oracle.filter_pairs_F(sc["xys"], sc["widths"], sc["heights"], pairs, ofs, mm)
    assert dict_sets(two[1]) == match_sets(o_ofs, o_m, pairs)
