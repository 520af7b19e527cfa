"""Synthetic scenes for tests and bench.

One scene model feeds every stage so that putative matches, F-inliers and bundle adjustment are
mutually consistent: 3-D points in a box, cameras on a ring looking at the centroid, pinhole
f = 1.1*max(w,h) (mirrors the reference's default focal, src/R3DProject.cpp:1152-1159),
pp = image centre, no distortion.  Everything is generated with numpy.random.Generator(PCG64(seed)).
"""
import numpy as np


def _rodrigues(aa):
    th = np.linalg.norm(aa)
    if th < 1e-12:
        return np.eye(3)
    k = aa / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def _log_so3(R):
    c = np.clip((np.trace(R) - 1) / 2, -1, 1)
    th = np.arccos(c)
    if th < 1e-12:
        return np.zeros(3)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / (2 * np.sin(th))
    return w * th


def look_at_pose(C, target):
    """Rotation (world->camera) and translation t = -R C for a camera at C looking at target."""
    z = target - C
    z = z / np.linalg.norm(z)
    up = np.array([0.0, 0.0, 1.0])
    x = np.cross(z, up)
    x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z])
    return R, -R @ C


def ring_cameras(n_cams, rng, radius=15.0, target=(5.0, 5.0, 2.0)):
    target = np.asarray(target, float)
    poses = np.zeros((n_cams, 6))
    Rs, ts = [], []
    for c in range(n_cams):
        ang = 2 * np.pi * c / n_cams
        C = target + np.array([radius * np.cos(ang), radius * np.sin(ang), rng.uniform(1.0, 4.0)])
        R, t = look_at_pose(C, target)
        poses[c, :3] = _log_so3(R)
        poses[c, 3:] = t
        Rs.append(R)
        ts.append(t)
    return poses, np.stack(Rs), np.stack(ts)


def project(R, t, X, f, w, h):
    p = X @ R.T + t
    z = p[:, 2]
    uv = np.stack([f * p[:, 0] / z + w / 2, f * p[:, 1] / z + h / 2], 1)
    vis = (z > 0.1) & (uv[:, 0] >= 0) & (uv[:, 0] < w) & (uv[:, 1] >= 0) & (uv[:, 1] < h)
    return uv, vis


def round_sig(x, sig=6):
    """Round to `sig` significant digits like the .feat text format (default ostream precision)."""
    x = np.asarray(x, np.float64)
    out = np.zeros_like(x)
    nz = x != 0
    mag = np.floor(np.log10(np.abs(x[nz])))
    scale = 10.0 ** (sig - 1 - mag)
    out[nz] = np.round(x[nz] * scale) / scale
    return out


def _descriptor_family(kind, dim, rng, n):
    g = rng.standard_normal((n, dim))
    if kind in ("liop", "sift"):
        g = np.abs(g)
    return g


def _finish_descriptor(kind, d):
    d = d / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-12)
    if kind == "sift":
        d = np.minimum(d, 0.2)
        d = d / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-12)
        d = np.clip(np.round(d * 512.0), 0, 255)
        return d.astype(np.float32)
    return d.astype(np.float32)


def make_scene(n_images, n_feats, dim, kind="msurf", seed=0, n_points=None, w=1920, h=1080,
               sigma=None, noise_px=0.5, as_u8=False):
    """Returns dict with descs[list of (n,dim)], xys[list of (n,2) f32], feats[list of (n,4)],
    truth[list of point-id per feature, -1 = distractor], widths, heights, poses, points, f.

    kind: 'msurf' (N(0,1)^D, unit norm), 'liop' (|N(0,1)|^D, unit norm), 'sift' (integer 0..255).
    sigma: per-view descriptor perturbation (relative to a unit-variance base); chosen so that
    roughly half of the true correspondences pass the 0.6 ratio test.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    if n_points is None:
        n_points = max(16, int(n_feats * 1.5))
    if sigma is None:
        sigma = {"msurf": 0.50, "liop": 0.35, "sift": 0.34}[kind]
    X = rng.uniform([0, 0, 0], [10, 10, 4], size=(n_points, 3))
    base = _descriptor_family(kind, dim, rng, n_points)
    poses, Rs, ts = ring_cameras(n_images, rng)
    f = 1.1 * max(w, h)
    descs, xys, feats, truth = [], [], [], []
    for c in range(n_images):
        uv, vis = project(Rs[c], ts[c], X, f, w, h)
        ids = np.nonzero(vis)[0]
        if len(ids) > n_feats:
            ids = rng.choice(ids, n_feats, replace=False)
        n_true = len(ids)
        n_dis = n_feats - n_true
        d_true = base[ids] + sigma * rng.standard_normal((n_true, dim))
        if kind in ("liop", "sift"):
            d_true = np.abs(d_true)
        d_dis = _descriptor_family(kind, dim, rng, n_dis)
        d = _finish_descriptor(kind, np.concatenate([d_true, d_dis], 0))
        xy_true = uv[ids] + noise_px * rng.standard_normal((n_true, 2))
        xy_dis = np.stack([rng.uniform(0, w, n_dis), rng.uniform(0, h, n_dis)], 1)
        xy = round_sig(np.concatenate([xy_true, xy_dis], 0)).astype(np.float32)
        tid = np.concatenate([ids, -np.ones(n_dis, int)])
        perm = rng.permutation(n_feats)
        d, xy, tid = d[perm], xy[perm], tid[perm]
        so = np.stack([rng.uniform(1.0, 8.0, n_feats), rng.uniform(0, 360, n_feats)], 1)
        descs.append(d.astype(np.uint8) if (as_u8 and kind == "sift") else d)
        xys.append(xy)
        feats.append(np.concatenate([xy, round_sig(so).astype(np.float32)], 1))
        truth.append(tid)
    return {
        "descs": descs, "xys": xys, "feats": feats, "truth": truth,
        "widths": np.full(n_images, w, np.uint32), "heights": np.full(n_images, h, np.uint32),
        "poses": poses, "points": X, "f": f, "w": w, "h": h, "kind": kind, "dim": dim,
    }


def exhaustive_pairs(n):
    """Pair_Builder exhaustivePairs(N): all (I,J), I<J, in std::set order (R3DComputeMatches.cpp:2042)."""
    i, j = np.triu_indices(n, 1)
    return np.stack([i, j], 1).astype(np.uint32)


def make_ba_problem(n_cams=200, n_pts=200000, obs_per_pt=5, seed=0, w=1920, h=1080,
                    noise_px=0.5, outlier_frac=0.02, perturb=True):
    """BA problem of SURVEY.md 8d (C5): each point observed by its `obs_per_pt` nearest-visible
    cameras; one shared radial-K3 intrinsic group; initial state = truth + noise."""
    rng = np.random.Generator(np.random.PCG64(seed))
    X = rng.uniform([0, 0, 0], [10, 10, 4], size=(n_pts, 3))
    poses, Rs, ts = ring_cameras(n_cams, rng)
    f = 1.1 * max(w, h)
    # camera centres
    Cs = np.stack([-Rs[c].T @ ts[c] for c in range(n_cams)])
    obs_cam = np.zeros((n_pts, obs_per_pt), np.uint32)
    obs_xy = np.zeros((n_pts, obs_per_pt, 2))
    # visibility: project every point in every camera in chunks of cameras
    vis_all = np.zeros((n_cams, n_pts), bool)
    uv_all = np.zeros((n_cams, n_pts, 2), np.float32)
    for c in range(n_cams):
        uv, vis = project(Rs[c], ts[c], X, f, w, h)
        vis_all[c] = vis
        uv_all[c] = uv
    d2 = ((X[None, :, :2] - Cs[:, None, :2]) ** 2).sum(-1)  # n_cams x n_pts
    d2 = np.where(vis_all, d2, np.inf)
    order = np.argsort(d2, axis=0)[:obs_per_pt]  # obs_per_pt x n_pts
    keep = np.isfinite(np.take_along_axis(d2, order, 0)).all(0)
    pts_idx = np.nonzero(keep)[0]
    X = X[pts_idx]
    order = order[:, pts_idx]
    n_pts = X.shape[0]
    obs_cam = order.T.astype(np.uint32).copy()
    obs_xy = np.stack([uv_all[obs_cam[:, k], pts_idx] for k in range(obs_per_pt)], 1).astype(np.float64)
    obs_xy += noise_px * rng.standard_normal(obs_xy.shape)
    n_obs = n_pts * obs_per_pt
    out = rng.random(n_obs) < outlier_frac
    obs_xy = obs_xy.reshape(n_obs, 2)
    obs_xy[out] += rng.choice([-50.0, 50.0], size=(out.sum(), 2))
    obs_pt = np.repeat(np.arange(n_pts, dtype=np.uint32), obs_per_pt)
    obs_cam = obs_cam.reshape(n_obs)
    intr = np.array([[f, w / 2, h / 2, 0.0, 0.0, 0.0]])
    poses0, X0, intr0 = poses.copy(), X.copy(), intr.copy()
    if perturb:
        poses0[:, :3] += 1e-2 * rng.standard_normal((n_cams, 3))
        poses0[:, 3:] += 1e-2 * rng.standard_normal((n_cams, 3))
        X0 += 5e-2 * rng.standard_normal(X0.shape)
        intr0[0, 0] *= 1 + 0.01 * rng.standard_normal()
    return {
        "poses": poses0, "intrinsics": intr0, "points": X0,
        "obs_cam": obs_cam, "obs_pt": obs_pt, "cam_intr": np.zeros(n_cams, np.uint32),
        "obs_xy": obs_xy, "truth": {"poses": poses, "points": X, "intrinsics": intr},
    }
