"""-m "not gpu": the cascade-hashing oracle (oracle/oracle_cascade.cpp, OpenMVG CASCADE_HASHING_L2 per SURVEY.md A.8)
against an independent numpy restatement, plus the properties the construction guarantees."""
import numpy as np
import pytest

from cascade_ref import hash_view, match_pair, zero_mean
from conftest import match_sets
from regard3d_b200 import synth


def test_projection_table_is_the_libstdcxx_normal_stream(oracle):
    P = oracle.cascade_projections(128)
    assert P.shape == (188, 128) and P.dtype == np.float32
    assert abs(float(P.mean())) < 0.02 and abs(float(P.std()) - 1.0) < 0.02
    # a fresh table of another dimension starts the same generator again: same leading values
    Q = oracle.cascade_projections(64)
    assert np.array_equal(Q[0, :64], P[0, :64])
    assert not np.array_equal(Q[1, :64], P[1, :64])  # row 1 starts 64 draws in for dim 64, 128 for dim 128


@pytest.mark.parametrize("kind,dim,as_u8", [("sift", 128, True), ("msurf", 64, False)])
def test_cascade_matches_equal_the_numpy_restatement(oracle, kind, dim, as_u8):
    sc = synth.make_scene(3, 400, dim, kind, seed=21, as_u8=as_u8)
    pairs = synth.exhaustive_pairs(3)
    ofs, m = oracle.cascade_match_pairs(sc["descs"], sc["xys"], pairs, 0.8)
    got = match_sets(ofs, m, pairs)
    P = oracle.cascade_projections(dim)
    mean = zero_mean(sc["descs"], [0, 1, 2])
    hv = [hash_view(d, P, mean) for d in sc["descs"]]
    n_total = 0
    for k, (I, J) in enumerate(pairs):
        raw = match_pair(sc["descs"][I], hv[I][0], hv[I][1], sc["descs"][J], hv[J][0], hv[J][1], 0.8)
        if not raw:
            assert (I, J) not in got
            continue
        arr = np.array(sorted(set(raw)), np.uint32)
        mm = np.zeros(len(arr), oracle.indmatch_dtype)
        mm["i"], mm["j"] = arr[:, 0], arr[:, 1]
        exp = oracle.coord_dedup(mm, sc["xys"][I], sc["xys"][J])  # the common tail (tested on its own elsewhere)
        assert got[(int(I), int(J))] == set(zip(exp["i"].tolist(), exp["j"].tolist()))
        n_total += len(exp)
    assert n_total > 50


def test_cascade_is_a_filter_on_correct_correspondences(oracle):
    """What cascade hashing returns are distinctive nearest neighbours among a candidate subset: on a synthetic scene
    almost all of them are true correspondences, and a good share of the exact matcher's matches is found."""
    sc = synth.make_scene(3, 3000, 128, "sift", seed=22, as_u8=True)
    pairs = synth.exhaustive_pairs(3)
    ofs, m = oracle.cascade_match_pairs(sc["descs"], sc["xys"], pairs, 0.6)
    ofs2, m2 = oracle.match_pairs(sc["descs"], sc["xys"], pairs, 0.6)
    a, b = match_sets(ofs, m, pairs), match_sets(ofs2, m2, pairs)
    truth = sc["truth"]
    good = tot = common = exact = 0
    for (I, J), s in a.items():
        good += sum(1 for (i, j) in s if truth[I][i] == truth[J][j] and truth[I][i] >= 0)
        tot += len(s)
        common += len(s & b.get((I, J), set()))
    for s in b.values():
        exact += len(s)
    assert tot > 200 and good >= 0.95 * tot
    assert common >= 0.2 * exact


def test_empty_and_tiny_views_give_nothing(oracle):
    rng = np.random.default_rng(5)
    sizes = [0, 1, 2, 3, 300]
    descs = [rng.integers(0, 255, (n, 128)).astype(np.uint8) for n in sizes]
    xys = [rng.uniform(0, 500, (n, 2)).astype(np.float32) for n in sizes]
    pairs = synth.exhaustive_pairs(len(sizes))
    ofs, m = oracle.cascade_match_pairs(descs, xys, pairs, 0.9)
    got = match_sets(ofs, m, pairs)
    for (I, J) in got:
        assert sizes[I] >= 3 and sizes[J] >= 1


def test_projection_table_equals_a_numpy_replay_of_libstdcxx(oracle):
    """CascadeHasher::Init draws std::normal_distribution<>(0, 1) from std::mt19937(default_seed = 5489).  libstdc++'s
    normal_distribution is Marsaglia's polar method on generate_canonical<double, 53> (two 32-bit draws per uniform, low
    word first) and hands out y * mult first, keeping x * mult for the next call.  numpy's legacy RandomState(5489) is
    the same init_genrand(5489) generator, so its raw 32-bit stream replays the table."""
    rs = np.random.RandomState(5489)
    raw = rs._bit_generator.random_raw(200000).astype(np.uint64)
    pos = [0]

    def canonical():
        lo, hi = raw[pos[0]], raw[pos[0] + 1]
        pos[0] += 2
        v = (float(lo) + float(hi) * 4294967296.0) / 18446744073709551616.0
        return v if v < 1.0 else np.nextafter(1.0, 0.0)

    out, saved = [], None
    need = 300
    while len(out) < need:
        if saved is not None:
            out.append(saved)
            saved = None
            continue
        while True:
            x = 2.0 * canonical() - 1.0
            y = 2.0 * canonical() - 1.0
            r2 = x * x + y * y
            if not (r2 > 1.0 or r2 == 0.0):
                break
        mult = np.sqrt(-2.0 * np.log(r2) / r2)
        saved = x * mult
        out.append(y * mult)
    ref = np.array(out[:need], np.float64).astype(np.float32)
    P = oracle.cascade_projections(128)
    got = P.reshape(-1)[:need]
    assert np.allclose(got, ref, rtol=2e-7, atol=0)          # libm log/sqrt may differ from numpy's in the last ulp
    assert np.array_equal(got, ref) or np.count_nonzero(got != ref) <= 2
