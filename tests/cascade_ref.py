"""numpy restatement of the cascade-hashing tables and of one query, independent of the C++ oracle's control flow
(same algorithm: SURVEY.md A.8).  Float32 throughout; the projections accumulate over k in order, product and sum
rounded separately (elementwise numpy float32 ops round exactly like the scalar code)."""
import numpy as np


def zero_mean(descs, used):
    rows = []
    for v in used:
        d = descs[v].astype(np.float32)
        acc = np.zeros(d.shape[1] if d.ndim == 2 else 0, np.float32)
        for i in range(d.shape[0]):
            acc = acc + d[i]
        rows.append(acc / np.float32(d.shape[0]) if d.shape[0] else acc)
    acc = np.zeros_like(rows[0])
    for r in rows:
        acc = acc + r
    return acc / np.float32(len(rows))


def hash_view(desc, proj, mean):
    """-> code bits [n][dim] (bool), bucket ids [n][6]"""
    d = desc.astype(np.float32) - mean[None, :]
    n, dim = d.shape
    npj = proj.shape[0]
    acc = np.zeros((n, npj), np.float32)
    for k in range(dim):
        acc = acc + (proj[None, :, k] * d[:, None, k]).astype(np.float32)
    bits = acc > 0
    code = bits[:, :dim]
    bucket = np.zeros((n, 6), np.int64)
    for g in range(6):
        for k in range(10):
            bucket[:, g] = bucket[:, g] * 2 + bits[:, dim + 10 * g + k]
    return code, bucket


def pack_code(code):
    n, dim = code.shape
    words = (dim + 31) // 32
    out = np.zeros((n, words), np.uint32)
    for j in range(dim):
        out[:, j // 32] |= (code[:, j].astype(np.uint32) << np.uint32(j % 32))
    return out


def bucket_lists(bucket, n):
    ofs = np.zeros((6, 1025), np.uint32)
    ids = np.zeros((6, n), np.uint32)
    for g in range(6):
        order = np.argsort(bucket[:, g], kind="stable")
        ids[g] = order
        cnt = np.bincount(bucket[:, g], minlength=1024)
        ofs[g, 1:] = np.cumsum(cnt)
    return ofs, ids


def l2_upstream(a, b):
    """openMVG L2: 4-way unrolled float accumulate (dim % 4 == 0 here)."""
    a = a.astype(np.float32)
    b = b.astype(np.float32)
    acc = np.float32(0)
    for k in range(0, len(a), 4):
        d = a[k:k + 4] - b[k:k + 4]
        s = np.float32(np.float32(np.float32(d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]) + d[3] * d[3])
        acc = np.float32(acc + s)
    return acc


def match_pair(descI, codeI, bucketI, descJ, codeJ, bucketJ, ratio):
    """-> list of (i, j) before the two de-duplications, in query order"""
    nI = len(descI)
    ofs, ids = bucket_lists(bucketI, nI)
    fr = np.float32(ratio) * np.float32(ratio)
    out = []
    for q in range(len(descJ)):
        cand = []
        for g in range(6):
            b = int(bucketJ[q, g])
            cand.extend(ids[g, ofs[g, b]:ofs[g, b + 1]].tolist())
        if len(cand) <= 2:
            continue
        seen, by_h = set(), {}
        for c in cand:
            if c in seen:
                continue
            seen.add(c)
            h = int(np.count_nonzero(codeJ[q] != codeI[c]))
            by_h.setdefault(h, []).append(c)
        top = []
        for h in sorted(by_h):
            for c in by_h[h]:
                if len(top) < 10:
                    top.append(c)
        if len(top) < 2:
            continue
        e = sorted((float(l2_upstream(descI[c], descJ[q])), c) for c in top)
        if np.float32(e[0][0]) < np.float32(fr * np.float32(e[1][0])):
            out.append((e[0][1], q))
    return out
