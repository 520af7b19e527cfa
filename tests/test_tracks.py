"""Track building: openMVG::tracks::TracksBuilder Build + Filter + ExportToSTL and GetTracksInImages -- what
Regard3D itself calls at src/threads/PreviewGeneratorThread.cpp:345-358 -- host code of the product against the oracle's
restatement (std::set / std::map / recursive union-find like upstream).  No GPU needed."""
import numpy as np

from regard3d_b200 import synth


def _random_matches(rng, n_views, n_feats, n_pairs_frac=0.7, per_pair=60, noise=0.15):
    """Feature f of every view belongs to 3-D point f (a ground-truth track); `noise` of the matches are wrong."""
    pairs, ofs, ms = [], [0], []
    for i in range(n_views):
        for j in range(i + 1, n_views):
            if rng.random() > n_pairs_frac:
                continue
            f = rng.choice(n_feats, per_pair, replace=False)
            g = f.copy()
            bad = rng.random(per_pair) < noise
            g[bad] = rng.integers(0, n_feats, bad.sum())
            m = np.unique(np.stack([f, g], 1), axis=0)
            pairs.append((i, j)); ms.append(m); ofs.append(ofs[-1] + len(m))
    return np.array(pairs, np.uint32), np.array(ofs, np.uint64), np.concatenate(ms).astype(np.uint32)


def test_tracks_equal_oracle(r3dlib, oracle):
    rng = np.random.default_rng(4)
    for trial in range(6):
        pairs, ofs, m2 = _random_matches(rng, n_views=5 + trial, n_feats=150, noise=0.1 * trial)
        m = np.zeros(len(m2), r3dlib.indmatch_dtype)
        m["i"], m["j"] = m2[:, 0], m2[:, 1]
        want = oracle.tracks_build(pairs, ofs, m, 2)
        got = r3dlib.Tracks.build(r3dlib.Matches.from_csr(pairs, ofs, m), 2).to_dict()
        assert got == want, trial                             # same track ids (union-find roots), same (view, feature) sets
        assert list(got) == sorted(got)                       # std::map order
        for t in got.values():
            assert len(t) >= 2
        want3 = oracle.tracks_build(pairs, ofs, m, 3)
        assert r3dlib.Tracks.build(r3dlib.Matches.from_csr(pairs, ofs, m), 3).to_dict() == want3


def test_tracks_drop_image_collisions_and_short_tracks(r3dlib):
    # (0,1): a-b ; (1,2): b-c ; (0,2): a2-c  -> features a and a2 of image 0 end up in one track: it must disappear
    pairs = np.array([[0, 1], [0, 2], [1, 2], [3, 4]], np.uint32)
    ofs = np.array([0, 2, 3, 4, 5], np.uint64)
    m = np.array([(10, 20), (11, 21), (12, 30), (20, 30), (5, 6)], r3dlib.indmatch_dtype)
    t = r3dlib.Tracks.build(r3dlib.Matches.from_csr(pairs, ofs, m), 2).to_dict()
    assert sorted(map(sorted, (d.items() for d in t.values()))) == [[(0, 11), (1, 21)], [(3, 5), (4, 6)]]
    assert r3dlib.Tracks.build(r3dlib.Matches.from_csr(pairs, ofs, m), 3).to_dict() == {}


def test_tracks_in_images(r3dlib, oracle):
    rng = np.random.default_rng(9)
    pairs, ofs, m2 = _random_matches(rng, n_views=6, n_feats=120, noise=0.0)
    m = np.zeros(len(m2), r3dlib.indmatch_dtype)
    m["i"], m["j"] = m2[:, 0], m2[:, 1]
    tr = r3dlib.Tracks.build(r3dlib.Matches.from_csr(pairs, ofs, m), 2)
    full = tr.to_dict()
    sub = tr.in_images([4, 1]).to_dict()                      # GetTracksInImages({1, 4})
    want = {k: {1: v[1], 4: v[4]} for k, v in full.items() if 1 in v and 4 in v}
    assert sub == want and len(sub) > 0
