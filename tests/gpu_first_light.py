"""Stand-alone GPU bring-up script (not a pytest file): prints diagnostics for the matching path.

    gpurun -- 'timeout 600 python tests/gpu_first_light.py > gpurun_out/first_light.log 2>&1'
"""
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np

from oracle import pyoracle as po
from regard3d_b200 import capi, synth


def as_set(m):
    return set(zip(m["i"].tolist(), m["j"].tolist()))


def compare(name, got, exp_ofs, exp_m, pairs):
    bad = 0
    tot = 0
    for k, (I, J) in enumerate(pairs):
        e = as_set(exp_m[int(exp_ofs[k]):int(exp_ofs[k + 1])])
        g = as_set(got.get((int(I), int(J)), np.zeros(0, capi.indmatch_dtype)))
        tot += len(e)
        if e != g:
            bad += 1
            if bad <= 3:
                print("   pair", (I, J), "expected", len(e), "got", len(g), "missing", len(e - g), "extra", len(g - e))
    print("[%s] pairs=%d mismatching=%d total expected matches=%d" % (name, len(pairs), bad, tot))
    return bad == 0


def stage(name, fn):
    print("=" * 20, name)
    sys.stdout.flush()
    t = time.time()
    try:
        ok = fn()
    except Exception:
        traceback.print_exc()
        ok = False
    print("-> %s: %s (%.2fs)" % (name, "OK" if ok else "FAIL", time.time() - t))
    sys.stdout.flush()
    return ok


def main():
    ctx = capi.Context((0,))
    sc = synth.make_scene(3, 1500, 64, "msurf", seed=3)
    pairs = synth.exhaustive_pairs(3)
    for v in range(3):
        ctx.upload_regions(v, sc["descs"][v], sc["xys"][v])
    exp_ofs, exp_m = po.match_pairs(sc["descs"], sc["xys"], pairs, 0.6)

    def s_exact():
        got = ctx.match_pairs(pairs, 0.6, capi.MATCH_EXACT_SCAN).to_dict()
        print("timing", ctx.match_timing())
        return compare("exact-scan", got, exp_ofs, exp_m, pairs)

    def s_keys():
        nq = sc["descs"][1].shape[0]
        keys, eps = ctx.debug_candidate_keys(0, 1, nq)
        A = sc["descs"][0].astype(np.float16).astype(np.float64)
        B = sc["descs"][1].astype(np.float16).astype(np.float64)
        n2a = (sc["descs"][0].astype(np.float64) ** 2).sum(1)
        n2b = (sc["descs"][1].astype(np.float64) ** 2).sum(1)
        D = n2b[:, None] + n2a[None, :] - 2 * B @ A.T
        nI = A.shape[0]
        npad = (nI + 255) // 256 * 256
        Dp = np.full((nq, npad), 1e30)
        Dp[:, :nI] = D
        cm = Dp.reshape(nq, npad // 16, 16).min(2)
        order = np.argsort(cm, 1)[:, :4]
        bits = int(np.ceil(np.log2(npad // 16)))
        kv = keys[:nq, :4].view(np.float32)
        kc = keys[:nq, :4] & ((1 << bits) - 1)
        print("eps_abs", eps)
        print("first rows keys(val,chunk):", [(float(kv[0, t]), int(kc[0, t])) for t in range(4)])
        print("numpy  chunk mins         :", [(float(cm[0, order[0, t]]), int(order[0, t])) for t in range(4)])
        agree = (kc[:, :3] == order[:, :3]).mean()
        exp_v = np.take_along_axis(cm, order, 1)
        relerr = np.abs(kv - exp_v) / np.maximum(np.abs(exp_v), 1e-6)
        print("chunk-id agreement (top3): %.4f   max rel err of key values (where ids agree): %.3e" %
              (agree, relerr[:, :3][kc[:, :3] == order[:, :3]].max() if agree > 0 else float("nan")))
        # approximate-vs-true distance error, to validate the eps model
        Dtrue = n2b[:, None] + n2a[None, :] - 2 * sc["descs"][1].astype(np.float64) @ sc["descs"][0].astype(np.float64).T
        print("max |fp16 distance - true distance| = %.3e (eps_abs %.3e)" % (np.abs(D - Dtrue).max(), eps))
        return agree > 0.99

    def s_tc():
        got = ctx.match_pairs(pairs, 0.6).to_dict()
        print("timing", ctx.match_timing())
        return compare("tensor-core", got, exp_ofs, exp_m, pairs)

    def s_nn():
        idx, dist = ctx.search_neighbours(0, 1, sc["descs"][1].shape[0])
        oi, od = po.search_neighbours(sc["descs"][0], sc["descs"][1])
        print("idx equal frac", (idx == oi).mean(), "dist bit-equal frac", (dist.view(np.uint32) == od.view(np.uint32)).mean())
        return bool((idx == oi).all() and (dist.view(np.uint32) == od.view(np.uint32)).all())

    ok = stage("exact scan vs oracle (3 x 1500, D=64)", s_exact)
    ok &= stage("candidate keys vs numpy", s_keys)
    ok &= stage("tensor-core path vs oracle", s_tc)
    ok &= stage("search_neighbours vs oracle", s_nn)

    def big(kind, dim, n_img, n_feat, as_u8=False):
        def run():
            ctx.clear_regions()
            s2 = synth.make_scene(n_img, n_feat, dim, kind, seed=11, as_u8=as_u8)
            prs = synth.exhaustive_pairs(n_img)
            for v in range(n_img):
                ctx.upload_regions(v, s2["descs"][v], s2["xys"][v])
            t0 = time.time()
            tc = ctx.match_pairs(prs, 0.6)
            t1 = time.time()
            tm = ctx.match_timing()
            print("TC wall %.1f ms; timing %s" % ((t1 - t0) * 1e3, tm))
            tc2 = ctx.match_pairs(prs, 0.6)
            tm2 = ctx.match_timing()
            print("TC 2nd  timing %s" % (tm2,))
            flops = 2.0 * n_feat * n_feat * dim * len(prs)
            print("candidate kernel: %.1f TFLOP/s algorithmic (2*N*N*D per pair)" % (flops / (tm2["ms_candidates"] * 1e-3) / 1e12))
            ex = ctx.match_pairs(prs, 0.6, capi.MATCH_EXACT_SCAN)
            print("EXACT timing %s" % (ctx.match_timing(),))
            d1, d2 = tc2.to_dict(), ex.to_dict()
            same = set(d1.keys()) == set(d2.keys()) and all(as_set(d1[k]) == as_set(d2[k]) for k in d1)
            print("TC == EXACT-scan:", same, "pairs", len(d1), "matches", tc2.total)
            o_ofs, o_m = po.match_pairs(s2["descs"][:2], s2["xys"][:2], prs[:1], 0.6)
            okp = compare("oracle pair (0,1)", d1, o_ofs, o_m, prs[:1])
            return same and okp
        return run

    ok &= stage("4 x 10000 msurf D=64", big("msurf", 64, 4, 10000))
    ok &= stage("4 x 10000 liop D=144", big("liop", 144, 4, 10000))
    ok &= stage("3 x 20000 sift u8 D=128", big("sift", 128, 3, 20000, True))
    ok &= stage("3 x 5000 sift f32 D=128 ragged", big("sift", 128, 3, 5003))
    print("ALL OK" if ok else "SOME FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
