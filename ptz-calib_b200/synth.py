"""Synthetic PTZ scenes of the shapes BASELINE.json names.  The reference's datasets are not
available offline, so every test and benchmark input comes from here.

Model: single optical centre at the origin (t = 0).  View i has pan/tilt/roll -> R_i (camera <- local frame),
focal f_i, principal point at the image centre (ptz_incremental_optimizer.cc:326-327), optional radial k1.
A track is a direction on the unit sphere observed in >= 4 views (Filter(4), ptzray_optimizer.cc:541) with
N(0, sigma) pixel noise, rounded to float32 like cv::KeyPoint::pt (types.h:20).  Track weight = track length
(ptzray_optimizer.cc:805).  Initial guess = ground truth perturbed (rvec, focal), rays left to Pix2Ray.
"""
import numpy as np
from scipy.spatial import cKDTree

from . import abi
from .problem import BAProblem, RelocBatch

SEEDS = {1: 1001, 2: 1002, 3: 1003, 4: 1004, 5: 1005}


# ------------------------------------------------------------------------------------------ rotations
def rodrigues_np(r):
    """rotation vector(s) [...,3] -> matrices [...,3,3] (closed form; numpy, for data generation only)"""
    r = np.asarray(r, dtype=np.float64)
    th = np.linalg.norm(r, axis=-1, keepdims=True)
    small = th < 1e-12
    k = r / np.where(small, 1.0, th)
    K = np.zeros(r.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -k[..., 2], k[..., 1]
    K[..., 1, 0], K[..., 1, 2] = k[..., 2], -k[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -k[..., 1], k[..., 0]
    s, c = np.sin(th)[..., None], np.cos(th)[..., None]
    R = np.eye(3) + s * K + (1 - c) * (K @ K)
    return R


def log_so3(R):
    """rotation matrices [...,3,3] -> rotation vectors (theta < pi assumed)"""
    R = np.asarray(R, dtype=np.float64)
    c = np.clip((np.trace(R, axis1=-2, axis2=-1) - 1) * 0.5, -1, 1)
    th = np.arccos(c)
    v = np.stack([R[..., 2, 1] - R[..., 1, 2], R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] - R[..., 0, 1]], axis=-1)
    s = np.sin(th)
    fac = np.where(s > 1e-9, th / (2 * np.where(s > 1e-9, s, 1.0)), 0.5)
    return v * fac[..., None]


def ptz_rotation(pan, tilt, roll):
    """camera <- local rotation of a pan/tilt/roll head: R = Rz(roll) Rx(tilt) Ry(pan)"""
    pan, tilt, roll = np.broadcast_arrays(np.asarray(pan, float), np.asarray(tilt, float), np.asarray(roll, float))
    cp, sp, ct, st, cr, sr = np.cos(pan), np.sin(pan), np.cos(tilt), np.sin(tilt), np.cos(roll), np.sin(roll)
    z, o = np.zeros_like(pan), np.ones_like(pan)
    Ry = np.stack([np.stack([cp, z, -sp], -1), np.stack([z, o, z], -1), np.stack([sp, z, cp], -1)], -2)
    Rx = np.stack([np.stack([o, z, z], -1), np.stack([z, ct, -st], -1), np.stack([z, st, ct], -1)], -2)
    Rz = np.stack([np.stack([cr, -sr, z], -1), np.stack([sr, cr, z], -1), np.stack([z, z, o], -1)], -2)
    return Rz @ Rx @ Ry


def project(R, f, c, k1, n, dz=0.0):
    """n [..,3] local directions/points -> pixels with the Brown k1 term; returns (uv, z).  dz: displacement added to the depth
    (the d0 + d1 f + d2 f^2 of PTZRayDistDispFactor / Reproj2d3dDispFactor, ptzray_optimizer.cc:231-232,368-369)"""
    X = np.einsum("...ij,...j->...i", R, n)
    z = X[..., 2] + dz
    zz = np.where(np.abs(z) > 1e-12, z, 1e-12)
    x, y = X[..., 0] / zz, X[..., 1] / zz
    rad = 1.0 + k1 * (x * x + y * y)
    return np.stack([f * x * rad + c[..., 0], f * y * rad + c[..., 1]], -1), z


# ------------------------------------------------------------------------------------------ BA scenes
def _views(cfg, rng, V, width, height):
    if cfg == "ring":  # cfg 1: pan every 360/V deg at tilt -20 deg
        pan = np.deg2rad(np.arange(V) * (10.0 if V <= 36 else 360.0 / V))  # partial ring when scaled down
        tilt = np.full(V, np.deg2rad(-20.0))
        f = rng.uniform(1400, 2200, V)
        k1 = np.zeros(V)
    elif cfg == "broadcast":  # cfg 2: WorldCup14-like sweep, zoom-dependent radial distortion
        pan = np.deg2rad(rng.uniform(-35, 35, V))
        tilt = np.deg2rad(rng.uniform(-20, -5, V))
        f = np.exp(rng.uniform(np.log(1500), np.log(6000), V))
        k1 = -0.25 * (1500.0 / f)
    elif cfg == "band":  # cfg 4/5: Fibonacci-sphere directions over a band of tilt in [-35, 10] deg
        i = np.arange(V) + 0.5
        pan_range = min(2 * np.pi, V / 15.0)  # keep >= ~10 views over every direction when scaled down
        pan = ((i * (3 - np.sqrt(5)) * 0.5) % 1.0) * pan_range
        s_lo, s_hi = np.sin(np.deg2rad(-35.0)), np.sin(np.deg2rad(10.0))
        tilt = np.arcsin(s_lo + (s_hi - s_lo) * i / V)
        f = rng.uniform(1400, 2200, V)
        k1 = np.zeros(V)
    else:
        raise ValueError(cfg)
    roll = np.deg2rad(rng.normal(0, 0.3, V))
    R = ptz_rotation(pan, tilt, roll)
    c = np.tile(np.array([width * 0.5, height * 0.5]), (V, 1))
    return R, f, c, k1


def make_ba_scene(V, P, layout="ring", factor_type=abi.PTZ_BA_PTZRAY, seed=1001, width=1920, height=1080, sigma=0.5, mean_extra_len=1.0,
                  max_len=None, rot_noise_deg=0.5, focal_noise=0.02, num_pts3d=0, pts3d_views=3, track_seed=None, k1_init_ratio=0.8,
                  gt_init=False, neighbours=24, focal_range=None, disp_gt=None, k1_scale=-0.08):
    """One PTZ-BA problem: V views, ~P tracks (those with < 4 visible views are dropped, so the result has <= P).

    track_seed: tracks are drawn from an independent stream so that ranks of a multi-GPU run can each generate
    their own shard of tracks over the SAME views (views use `seed`).
    """
    rng = np.random.default_rng(seed)
    R, f, c, k1 = _views(layout, rng, V, width, height)
    if focal_range is not None:  # zoom sweep (drawn from its own stream: the other draws of the scene stay what they were)
        f = np.exp(np.random.default_rng(seed + 77).uniform(np.log(focal_range[0]), np.log(focal_range[1]), V))
    # true displacement of the projection centre along the optical axis, a polynomial of the focal length (DistDisp scenes)
    dz = np.zeros(V) if disp_gt is None else disp_gt[0] + disp_gt[1] * f + disp_gt[2] * f * f
    if factor_type == abi.PTZ_BA_PTZRAY:
        k1 = np.zeros(V)
    elif layout != "broadcast":
        k1 = k1_scale * (1500.0 / f)
    # initial guess (drawn before the tracks so it does not depend on track_seed)
    rvec_gt = log_so3(R)
    if gt_init:
        R0, f0, k10 = R, f.copy(), k1.copy()
    else:
        dR = rodrigues_np(np.deg2rad(rng.normal(0, rot_noise_deg, (V, 3))))
        R0 = dR @ R
        f0 = f * (1 + rng.normal(0, focal_noise, V))
        k10 = k1 * k1_init_ratio
    trng = np.random.default_rng(seed * 7919 + 13 if track_seed is None else track_seed)
    axes = R[:, 2, :]  # optical axis of view i in the local frame = third row of R
    tree = cKDTree(axes)
    kq = min(V, neighbours)
    obs_uv, obs_view, obs_track, weights, rays = [], [], [], [], []
    next_track = 0
    chunk = 200000
    for start in range(0, P, chunk):
        n = min(chunk, P - start)
        anchor = trng.integers(0, V, n)
        px = np.stack([trng.uniform(0, width, n), trng.uniform(0, height, n)], -1)
        d = np.stack([(px[:, 0] - c[anchor, 0]) / f[anchor], (px[:, 1] - c[anchor, 1]) / f[anchor], np.ones(n)], -1)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        ray = np.einsum("nji,nj->ni", R[anchor], d)  # R^T d
        _, nb = tree.query(ray, k=kq)
        nb = nb.reshape(n, kq)
        uv, z = project(R[nb], f[nb], c[nb], k1[nb], ray[:, None, :], dz[nb])
        vis = (z > 0.1) & (uv[..., 0] >= 0) & (uv[..., 0] < width) & (uv[..., 1] >= 0) & (uv[..., 1] < height)
        want = 4 + trng.poisson(mean_extra_len, n)
        if max_len is not None:
            want = np.minimum(want, max_len)
        # random subset of the visible neighbours, of size `want`
        score = np.where(vis, trng.random((n, kq)), 2.0)
        order = np.argsort(score, axis=1)
        rank = np.empty_like(order)
        np.put_along_axis(rank, order, np.broadcast_to(np.arange(kq), (n, kq)), axis=1)
        keep = vis & (rank < want[:, None])
        length = keep.sum(1)
        good = length >= 4
        tid = np.full(n, -1, np.int64)
        tid[good] = next_track + np.arange(good.sum())
        next_track += int(good.sum())
        sel = keep & good[:, None]
        ti, ki = np.nonzero(sel)
        noisy = uv[ti, ki] + trng.normal(0, sigma, (len(ti), 2))
        obs_uv.append(noisy.astype(np.float32))
        obs_view.append(nb[ti, ki].astype(np.int32))
        obs_track.append(tid[ti].astype(np.int32))
        weights.append(length[good].astype(np.float64))
        rays.append(ray[good])
    obs_uv = np.concatenate(obs_uv) if obs_uv else np.zeros((0, 2), np.float32)
    obs_view = np.concatenate(obs_view) if obs_view else np.zeros(0, np.int32)
    obs_track = np.concatenate(obs_track) if obs_track else np.zeros(0, np.int32)
    weights = np.concatenate(weights) if weights else np.zeros(0)
    rays = np.concatenate(rays) if rays else np.zeros((0, 3))

    intr = np.zeros((V, 9))
    intr[:, 0], intr[:, 1], intr[:, 2], intr[:, 3], intr[:, 4] = f0, f0, c[:, 0], c[:, 1], k10
    ext = np.zeros((V, 6))
    ext[:, :3] = log_so3(R0)
    gt = dict(R=R, f=f, c=c, k1=k1, rvec=rvec_gt, rays=rays, width=width, height=height, disp=None if disp_gt is None else np.asarray(disp_gt, float))

    pt_uv = pt_xyz = pt_view = tlw0 = None
    if num_pts3d > 0:
        # georeferencing terms (AddConstraints2d3d): world points X_w with X_l = R_lw X_w + t_lw seen from the origin
        rlw = rng.normal(0, 0.4, 3)
        Rlw = rodrigues_np(rlw)
        tlw = rng.normal(0, 1.0, 3) * np.array([20.0, 5.0, 20.0])
        views = rng.choice(V, size=min(pts3d_views, V), replace=False)
        per = int(np.ceil(num_pts3d / len(views)))
        pu, px3, pv = [], [], []
        for vi in views:
            pxs = np.stack([rng.uniform(0.1 * width, 0.9 * width, per), rng.uniform(0.1 * height, 0.9 * height, per)], -1)
            dd = np.stack([(pxs[:, 0] - c[vi, 0]) / f[vi], (pxs[:, 1] - c[vi, 1]) / f[vi], np.ones(per)], -1)
            Xl = (R[vi].T @ dd.T).T * rng.uniform(30, 90, (per, 1))
            Xw = (Rlw.T @ (Xl - tlw).T).T
            uvp, _ = project(R[vi], f[vi], c[vi], k1[vi], Xl, dz[vi])
            pu.append((uvp + rng.normal(0, sigma, uvp.shape)).astype(np.float32))
            px3.append(Xw)
            pv.append(np.full(per, vi, np.int32))
        pt_uv, pt_xyz, pt_view = np.concatenate(pu)[:num_pts3d], np.concatenate(px3)[:num_pts3d], np.concatenate(pv)[:num_pts3d]
        if gt_init:
            tlw0 = np.concatenate([rlw, tlw])
        else:
            # what EPnP on one annotated view would hand over: ground truth up to the view's own initial error
            tlw0 = np.concatenate([rlw + np.deg2rad(rng.normal(0, 0.3, 3)), tlw + rng.normal(0, 0.3, 3)])
        gt.update(tlw=np.concatenate([rlw, tlw]))
    return BAProblem(factor_type=factor_type, intr=intr, ext=ext, obs_uv=obs_uv, obs_view=obs_view, obs_track=obs_track, track_weight=weights,
                     pt_uv=pt_uv, pt_xyz=pt_xyz, pt_view=pt_view, tlw0=tlw0, gt=gt)


def make_config(cfg, scale=1.0, factor_type=None, **kw):
    """The five BASELINE.json configurations.  scale < 1 shrinks views/tracks for tests."""
    s = scale
    if cfg == 1:  # Synthetic-shaped: V=36 ring, P~4000, M~20000 (+ ~10 annotated points for the georef stage)
        return make_ba_scene(max(4, int(36 * s)), int(4000 * s), "ring", abi.PTZ_BA_PTZRAY if factor_type is None else factor_type,
                             seed=SEEDS[1], neighbours=12, **kw)
    if cfg == 2:  # WorldCup14-shaped: V=60, PTZRayDist, 1280x720, P~6000, M~40000
        return make_ba_scene(max(6, int(60 * s)), int(8000 * s), "broadcast", abi.PTZ_BA_PTZRAY_DIST if factor_type is None else factor_type,
                             seed=SEEDS[2], width=1280, height=720, mean_extra_len=2.5, **kw)
    if cfg == 4:  # scaled BA: V=1000, P=4e5, M=2e6
        return make_ba_scene(max(8, int(1000 * s)), int(400000 * s), "band", abi.PTZ_BA_PTZRAY if factor_type is None else factor_type,
                             seed=SEEDS[4], **kw)
    if cfg == 5:  # sharded BA: V=10000, P=1e7, M=5e7
        return make_ba_scene(max(8, int(10000 * s)), int(10000000 * s), "band", abi.PTZ_BA_PTZRAY if factor_type is None else factor_type,
                             seed=SEEDS[5], **kw)
    raise ValueError(cfg)


def make_distdisp_scene(num_pts3d=0, V=12, P=3000, seed=1001, **kw):
    """A scene on which PTZRayDistDispFactor's global disp[3] is OBSERVABLE: the data carry a true displacement polynomial
    d0 + d1 f + d2 f^2 of the projection centre, the views sweep the zoom range (f log-uniform in [500, 2600] px, so that 1, f, f^2
    separate and the wide end has enough perspective for the depth offset to differ from a focal change) and the distortion is
    mild (Pix2Ray, the reference's ray initialisation, ignores it: ptzray_optimizer.cc:768-797).  Converges in tens of iterations."""
    return make_ba_scene(V, P, "ring", abi.PTZ_BA_PTZRAY_DIST_DISP, seed=seed, neighbours=12, focal_range=(500.0, 2600.0),
                         disp_gt=(0.03, -2e-5, 6e-9), k1_scale=-0.002, num_pts3d=num_pts3d, **kw)


# ------------------------------------------------------------------------------------------ reloc batches
def krt21(f, fy, c, R, t, dist):
    out = np.zeros(R.shape[:-2] + (21,))
    out[..., 0], out[..., 1], out[..., 2], out[..., 3] = f, fy, c[..., 0], c[..., 1]
    out[..., 4:13] = R.reshape(R.shape[:-2] + (9,))
    out[..., 13:16] = t
    out[..., 16:21] = dist
    return out


def make_reloc_batch(B, factor_type=abi.PTZ_KRT_F, seed=1003, n_min=64, n_max=512, width=1920, height=1080, sigma=0.5, outlier_frac=0.05,
                     outlier_sigma=40.0, num_ref=36, dpan_deg=5.0, k1=-0.1, max_iter=200, max_reproj_error=100.0, query_seed=None, pts_per_query=0):
    """cfg 3: B independent queries against a calibrated reference ring; init exactly as run_ptz_reloc.cc:97-104."""
    rng = np.random.default_rng(seed)
    Rref, fref, cref, _ = _views("ring", rng, num_ref, width, height)
    dist_on = factor_type in (abi.PTZ_KRT_FDIST, abi.PTZ_KRT_FXFYDIST)
    k1v = k1 if dist_on else 0.0
    q = np.random.default_rng(seed * 104729 + 7 if query_seed is None else query_seed)
    ref_idx = q.integers(0, num_ref, B)
    dR = ptz_rotation(np.deg2rad(q.normal(0, dpan_deg, B)), np.deg2rad(q.normal(0, dpan_deg, B)), np.deg2rad(q.normal(0, 0.3, B)))
    Rq = dR @ Rref[ref_idx]
    fq = fref[ref_idx] * q.uniform(0.7, 1.4, B)
    cq = np.tile(np.array([width * 0.5, height * 0.5]), (B, 1))
    nm = q.integers(n_min, n_max + 1, B)
    off = np.zeros(B + 1, np.int64)
    off[1:] = np.cumsum(nm)
    N = int(off[-1])
    qi = np.repeat(np.arange(B), nm)
    uv_ref = np.zeros((N, 2), np.float32)
    uv_cur = np.zeros((N, 2), np.float32)
    todo = np.arange(N)
    # rejection-sample matches: a pixel of the query image whose direction also falls inside the reference image
    for _ in range(60):
        if len(todo) == 0:
            break
        b = qi[todo]
        px = np.stack([q.uniform(0, width, len(todo)), q.uniform(0, height, len(todo))], -1)
        d = np.stack([(px[:, 0] - cq[b, 0]) / fq[b], (px[:, 1] - cq[b, 1]) / fq[b], np.ones(len(todo))], -1)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        ray = np.einsum("nji,nj->ni", Rq[b], d)
        uvc, zc = project(Rq[b], fq[b], cq[b], k1v, ray)
        uvr, zr = project(Rref[ref_idx[b]], fref[ref_idx[b]], cref[ref_idx[b]], k1v, ray)
        ok = (zr > 0.1) & (uvr[:, 0] >= 2) & (uvr[:, 0] < width - 2) & (uvr[:, 1] >= 2) & (uvr[:, 1] < height - 2)
        ok &= (uvc[:, 0] >= 0) & (uvc[:, 0] < width) & (uvc[:, 1] >= 0) & (uvc[:, 1] < height)
        uv_ref[todo[ok]] = (uvr[ok] + q.normal(0, sigma, (int(ok.sum()), 2))).astype(np.float32)
        uv_cur[todo[ok]] = (uvc[ok] + q.normal(0, sigma, (int(ok.sum()), 2))).astype(np.float32)
        todo = todo[~ok]
    if len(todo):  # pathological leftovers: reuse the first match of the query (keeps the ragged shape)
        uv_ref[todo] = uv_ref[off[qi[todo]]]
        uv_cur[todo] = uv_cur[off[qi[todo]]]
    out = q.random(N) < outlier_frac
    # gross outliers are mismatches displaced by N(0, outlier_sigma) px: kept, since the reference has no robust loss
    # (krt_optimizer.cc:294); uniform-random ones would push most queries over the 100 px gate of CheckResults
    uv_cur[out] = (uv_cur[out] + q.normal(0, outlier_sigma, (int(out.sum()), 2))).astype(np.float32)
    dist = np.zeros((B, 5))
    dist[:, 0] = k1v
    t0 = np.zeros((B, 3))
    ref_cam = krt21(fref[ref_idx], fref[ref_idx], cref[ref_idx], Rref[ref_idx], t0, dist)
    # SetInitParams: K = [f_ref, centre of the test image], R, t, dist of the reference camera
    init_cam = krt21(fref[ref_idx], fref[ref_idx], cq, Rref[ref_idx], t0, dist)
    gt = dict(R=Rq, f=fq, ref_idx=ref_idx)
    pts = {}
    if pts_per_query > 0:
        # optional 2d-3d terms (Add2d3dConstraints): world points in front of the query camera, seen with pixel noise
        npq = np.full(B, pts_per_query, np.int64)
        poff = np.zeros(B + 1, np.int64)
        poff[1:] = np.cumsum(npq)
        pb = np.repeat(np.arange(B), npq)
        px = np.stack([q.uniform(0.1 * width, 0.9 * width, len(pb)), q.uniform(0.1 * height, 0.9 * height, len(pb))], -1)
        dcam = np.stack([(px[:, 0] - cq[pb, 0]) / fq[pb], (px[:, 1] - cq[pb, 1]) / fq[pb], np.ones(len(pb))], -1) * q.uniform(10, 80, (len(pb), 1))
        Xw = np.einsum("nji,nj->ni", Rq[pb], dcam)
        uvp, _ = project(Rq[pb], fq[pb], cq[pb], k1v, Xw)
        pts = dict(pt_offset=poff, pt_uv=(uvp + q.normal(0, sigma, uvp.shape)).astype(np.float32), pt_xyz=Xw)
    return RelocBatch(factor_type, off, uv_ref, uv_cur, ref_cam, init_cam, max_iter, max_reproj_error, gt, **pts)


# ------------------------------------------------------------------------------------------ pairwise matches of a scene
def make_matches_from_scene(prob, seed=11, edge_prob=1.0, n_collisions=0, n_short=0, extra_keypoints=0):
    """What feature matching would hand to TracksBuilder for a synthetic scene: keypoints per image (a view's observations,
    feature id = rank inside the view, plus `extra_keypoints` unmatched ones per image) and pairwise matches (every pair of a
    track's observations with probability `edge_prob`, consecutive ones always, so the track stays connected).

    n_collisions: extra matches that glue two tracks seen in a common image together -> that image is listed twice and
    Filter rejects the merged set.  n_short: extra 2- and 3-image sets -> Filter(4) rejects them.

    Returns (Matches, Views, expected) where expected = {frozenset((image, feature), ...)} of the tracks that must survive.
    """
    from .tracks import Matches, Views

    rng = np.random.default_rng(seed)
    V, M = prob.V, prob.M
    order = np.lexsort((prob.obs_track, prob.obs_view))  # keypoints of an image in (view, track) order
    view_s, track_s = prob.obs_view[order], prob.obs_track[order]
    cnt = np.bincount(view_s, minlength=V).astype(np.int64)
    nkp = cnt + extra_keypoints
    kp_offset = np.concatenate([[0], np.cumsum(nkp)]).astype(np.int64)
    first = np.concatenate([[0], np.cumsum(cnt)])[:-1]
    feat = (np.arange(M) - first[view_s]).astype(np.int32)
    kp_uv = np.zeros((int(kp_offset[-1]), 2), np.float32)
    kp_uv[kp_offset[view_s] + feat] = prob.obs_uv[order]
    if extra_keypoints:
        for v in range(V):
            kp_uv[kp_offset[v] + cnt[v]:kp_offset[v + 1]] = rng.uniform(0, 1000, (extra_keypoints, 2)).astype(np.float32)
    # observations grouped by track (inside a track: ascending view)
    g = np.lexsort((view_s, track_s))
    tv, tf, tt = view_s[g], feat[g], track_s[g]
    ea, eb = [], []
    d = 1
    while True:
        same = tt[:-d] == tt[d:] if d < len(tt) else np.zeros(0, bool)
        if not same.any():
            break
        idx = np.nonzero(same)[0]
        if d > 1 and edge_prob < 1.0:
            idx = idx[rng.random(len(idx)) < edge_prob]
        ea.append(idx)
        eb.append(idx + d)
        d += 1
    ea = np.concatenate(ea) if ea else np.zeros(0, np.int64)
    eb = np.concatenate(eb) if eb else np.zeros(0, np.int64)
    src, dst, q, t = tv[ea], tv[eb], tf[ea], tf[eb]  # ascending view inside a track: src < dst
    dead = set()
    if n_collisions:
        # a match between a node of track A and a node of track B, A and B both seen in image v (as neighbours in (view, track) order)
        start = np.concatenate([[0], np.nonzero(tt[1:] != tt[:-1])[0] + 1, [len(tt)]])
        done = 0
        for k in rng.permutation(M - 1):
            if done >= n_collisions:
                break
            if view_s[k] != view_s[k + 1] or track_s[k] == track_s[k + 1]:
                continue
            A, B = int(track_s[k]), int(track_s[k + 1])
            if A in dead or B in dead:
                continue
            ia = np.arange(start[A], start[A + 1])
            ib = np.arange(start[B], start[B + 1])
            ia, ib = ia[tv[ia] != view_s[k]], ib[tv[ib] != view_s[k]]
            pairs = [(x, y) for x in ia for y in ib if tv[x] != tv[y]]
            if not pairs:
                continue
            x, y = pairs[rng.integers(len(pairs))]
            if tv[x] > tv[y]:
                x, y = y, x
            src, dst, q, t = np.append(src, tv[x]), np.append(dst, tv[y]), np.append(q, tf[x]), np.append(t, tf[y])
            dead.update((A, B))
            done += 1
    if n_short:
        assert extra_keypoints >= 1 and V >= 3
        used = np.zeros(V, np.int64)
        for _ in range(n_short):
            L = int(rng.integers(2, 4))
            vs = np.sort(rng.choice(V, L, replace=False))
            if (used[vs] >= extra_keypoints).any():
                continue
            fs = cnt[vs] + used[vs]
            used[vs] += 1
            for a in range(L - 1):
                src, dst, q, t = np.append(src, vs[a]), np.append(dst, vs[a + 1]), np.append(q, fs[a]), np.append(t, fs[a + 1])
    # group by image pair, pairs in ascending (src, dst): the order of a vector<MatchesInfo> filled by two nested loops
    o = np.lexsort((np.arange(len(src)), dst, src))
    src, dst, q, t = src[o], dst[o], q[o], t[o]
    key = src.astype(np.int64) * V + dst
    pk, pstart = np.unique(key, return_index=True)
    match_offset = np.concatenate([pstart, [len(key)]]).astype(np.int64)
    matches = Matches((pk // V).astype(np.int32), (pk % V).astype(np.int32), match_offset, q.astype(np.int32), t.astype(np.int32))
    views = Views(np.ones(V, np.uint8), kp_offset, kp_uv)
    expected = set()
    bounds = np.concatenate([[0], np.nonzero(tt[1:] != tt[:-1])[0] + 1, [len(tt)]])
    for k in range(len(bounds) - 1):
        b, e = bounds[k], bounds[k + 1]
        if int(tt[b]) in dead:
            continue
        expected.add(frozenset(zip(tv[b:e].tolist(), tf[b:e].tolist())))
    return matches, views, expected
