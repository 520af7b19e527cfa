"""-m gpu: R3D_MATCH_CASCADE_HASHING (cascade.cu) through the C ABI vs the CPU oracle -- identical hash tables and
identical (i, j) sequences (BASELINE config 4 names this matcher)."""
import numpy as np
import pytest

from cascade_ref import bucket_lists, hash_view, pack_code, zero_mean
from conftest import dict_sets, match_sets
from regard3d_b200 import synth

pytestmark = pytest.mark.gpu


def _upload(ctx, sc):
    ctx.clear_regions()
    for v, (d, x) in enumerate(zip(sc["descs"], sc["xys"])):
        ctx.upload_regions(v, d, x)


@pytest.mark.parametrize("kind,dim,as_u8", [("sift", 128, True), ("liop", 144, False), ("msurf", 64, False)])
def test_hash_tables_equal_the_numpy_restatement(gpu_ctx, oracle, kind, dim, as_u8):
    sc = synth.make_scene(3, 700, dim, kind, seed=31, as_u8=as_u8)
    _upload(gpu_ctx, sc)
    gpu_ctx.cascade_prepare([0, 1, 2])
    P = oracle.cascade_projections(dim)
    mean = zero_mean(sc["descs"], [0, 1, 2])
    for v in range(3):
        code, bucket = hash_view(sc["descs"][v], P, mean)
        g_code, g_bucket, g_ofs, g_ids = gpu_ctx.debug_cascade_view(v, 700, dim)
        assert np.array_equal(g_code, pack_code(code))
        assert np.array_equal(g_bucket.astype(np.int64), bucket)
        ofs, ids = bucket_lists(bucket, 700)
        assert np.array_equal(g_ofs, ofs) and np.array_equal(g_ids, ids)


@pytest.mark.parametrize("kind,dim,n,as_u8,ratio", [
    ("sift", 128, 3000, True, 0.6), ("sift", 128, 1500, False, 0.8), ("liop", 144, 2000, False, 0.8),
    ("msurf", 64, 2500, False, 0.7),
])
def test_cascade_match_pairs_equals_oracle(gpu_ctx, oracle, r3dlib, kind, dim, n, as_u8, ratio):
    sc = synth.make_scene(4, n, dim, kind, seed=32, as_u8=as_u8)
    pairs = synth.exhaustive_pairs(4)
    _upload(gpu_ctx, sc)
    ofs, m = oracle.cascade_match_pairs(sc["descs"], sc["xys"], pairs, ratio)
    exp = match_sets(ofs, m, pairs)
    res = gpu_ctx.match_pairs(pairs, ratio, r3dlib.MATCH_CASCADE_HASHING)
    assert dict_sets(res.to_dict()) == exp
    assert sum(len(s) for s in exp.values()) > 100
    # same SEQUENCE per pair (the reference's std::set order), not only the same set
    I, J, g = res.pair(0)
    k = [tuple(p) for p in pairs.tolist()].index((I, J))
    assert np.array_equal(g, m[int(ofs[k]):int(ofs[k + 1])])


def test_a_job_split_over_calls_needs_one_prepare(gpu_ctx, oracle, r3dlib):
    """The zero-mean descriptor belongs to the whole job: prepare(all views) + two half calls == one call == oracle;
    without the prepare each half is its own job (different hash tables, the reference's behaviour for that call)."""
    sc = synth.make_scene(5, 1200, 128, "sift", seed=33, as_u8=True)
    pairs = synth.exhaustive_pairs(5)
    _upload(gpu_ctx, sc)
    ofs, m = oracle.cascade_match_pairs(sc["descs"], sc["xys"], pairs, 0.7)
    exp = match_sets(ofs, m, pairs)
    gpu_ctx.cascade_prepare(list(range(5)))
    got = {}
    for part in (pairs[:4], pairs[4:]):
        got.update(dict_sets(gpu_ctx.match_pairs(part, 0.7, r3dlib.MATCH_CASCADE_HASHING).to_dict()))
    assert got == exp
    # a call on its own: the oracle of that sub-job
    _upload(gpu_ctx, sc)
    sub = pairs[7:]
    ofs2, m2 = oracle.cascade_match_pairs(sc["descs"], sc["xys"], sub, 0.7)
    assert dict_sets(gpu_ctx.match_pairs(sub, 0.7, r3dlib.MATCH_CASCADE_HASHING).to_dict()) == match_sets(ofs2, m2, sub)


def test_cascade_ragged_views(gpu_ctx, oracle, r3dlib):
    rng = np.random.default_rng(34)
    sizes = [0, 1, 2, 3, 40, 900]
    base = rng.integers(0, 255, (900, 128)).astype(np.uint8)
    descs = [np.clip(base[:n].astype(np.int32) + rng.integers(-6, 7, (n, 128)), 0, 255).astype(np.uint8) for n in sizes]
    xys = [rng.uniform(0, 500, (n, 2)).astype(np.float32) for n in sizes]
    pairs = synth.exhaustive_pairs(len(sizes))
    gpu_ctx.clear_regions()
    for v in range(len(sizes)):
        gpu_ctx.upload_regions(v, descs[v], xys[v])
    ofs, m = oracle.cascade_match_pairs(descs, xys, pairs, 0.9)
    assert dict_sets(gpu_ctx.match_pairs(pairs, 0.9, r3dlib.MATCH_CASCADE_HASHING).to_dict()) == match_sets(ofs, m, pairs)
