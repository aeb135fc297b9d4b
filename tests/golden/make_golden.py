#!/usr/bin/env python
"""Generates the golden vectors under tests/golden/ that PIN the oracle (oracle/ptz_oracle.cpp).

The reference ships no tests or fixtures and cannot be built offline, so the anchors are the
third-party implementations the reference itself calls, available in this container as OpenCV-python 4.13
(cv2.Rodrigues, cv2.projectPoints, cv2.undistortPoints), 60-digit mpmath derivatives of an independent
restatement of every functor, and scipy.optimize.least_squares minima of gauge-fixed problems.

Nothing here imports the oracle or the product: the vectors are produced by independent code, then
tests/test_oracle_golden.py checks the oracle against them and tests/test_gpu_*.py check the CUDA path.

Run:  python tests/golden/make_golden.py      (needs cv2, mpmath, scipy; writes *.npz next to this file)
"""
import os
import sys

import cv2
import mpmath as mp
import numpy as np
from scipy.optimize import least_squares

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from ptz_calib_b200 import synth  # noqa: E402  (scene generator only; pure numpy)

mp.mp.dps = 60
BA_TYPES = ["PTZRay", "PTZRayDist", "PTZRayFxfyDist", "PTZRayDistDisp"]
KRT_TYPES = ["F", "FDist", "Fxfy", "FxfyDist"]


# ------------------------------------------------------------------------------------------------ cv2-based functors
def cv_project(pt3, rvec, tvec, fx, fy, cx, cy, dist_cv):
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    out, _ = cv2.projectPoints(np.asarray(pt3, np.float64).reshape(1, 1, 3), np.asarray(rvec, np.float64), np.asarray(tvec, np.float64), K,
                               np.asarray(dist_cv, np.float64))
    return out.reshape(2)


def hand_to_cv(d):  # hand-written factors read v[10..14] as (k1,k2,k3,p1,p2); OpenCV wants (k1,k2,p1,p2,k3)
    return np.array([d[0], d[1], d[3], d[4], d[2]])


def ba_ray_cv(t, intr, ext, ray, disp, uv):
    """ptzray_optimizer.cc:20-264 through cv2.projectPoints"""
    fx = intr[0]
    fy = intr[1] if t == 2 else intr[0]
    n = ray if t == 1 else ray / np.linalg.norm(ray)
    R = cv2.Rodrigues(np.asarray(ext[:3], np.float64))[0]
    if t == 1 and (R @ n)[2] < 0:
        return np.array([1e6, 1e6])
    dist = np.zeros(5) if t == 0 else hand_to_cv(intr[4:9])
    tz = disp[0] + disp[1] * fx + disp[2] * fx * fx if t == 3 else 0.0
    p = cv_project(n, ext[:3], [0, 0, tz], fx, fy, intr[2], intr[3], dist)
    return np.asarray(uv, np.float64) - p


def ba_pt_cv(t, intr, ext, tlw, disp, uv, xyz):
    """ptzray_optimizer.cc:268-401"""
    Rl = cv2.Rodrigues(np.asarray(tlw[:3], np.float64))[0]
    Xl = Rl @ xyz + tlw[3:]
    fx = intr[0]
    tz = disp[0] + disp[1] * fx + disp[2] * fx * fx if t == 3 else 0.0
    p = cv_project(Xl, ext[:3], [0, 0, tz], fx, intr[1], intr[2], intr[3], hand_to_cv(intr[4:9]))
    return np.asarray(uv, np.float64) - p


def krt_cv(t, cam, refK4, refd, uv1, uv2):
    """krt_optimizer.cc:22-197; t: 0 F, 1 FDist, 2 Fxfy, 3 FxfyDist"""
    K1 = np.array([[refK4[0], 0, refK4[2]], [0, refK4[1], refK4[3]], [0, 0, 1.0]])
    if t in (1, 3):
        und = cv2.undistortPoints(np.asarray(uv1, np.float32).reshape(1, 1, 2), K1, np.asarray(refd, np.float64), None, K1).reshape(2)
        assert und.dtype == np.float32
        if und[0] < 0 or und[0] >= refK4[2] * 2 or und[1] < 0 or und[1] >= refK4[3] * 2:
            return np.zeros(2), und
        pt = np.array([und[0], und[1], 1.0], np.float64)
    else:
        und = np.asarray(uv1, np.float32)
        pt = np.array([uv1[0], uv1[1], 1.0], np.float64)
    ray = np.linalg.inv(K1) @ pt
    if t != 2:
        ray = ray / np.linalg.norm(ray)
    fx = cam[0]
    fy = cam[1] if t in (2, 3) else cam[0]
    dist = hand_to_cv(cam[10:15]) if t in (1, 3) else np.zeros(5)
    p = cv_project(ray, cam[4:7], [0, 0, 0], fx, fy, cam[2], cam[3], dist)
    return np.asarray(uv2, np.float64) - p, und


def krt3d_cv(t, cam, uv, xyz):
    """krt_optimizer.cc:201-248: cv::projectPoints with v[10..14] in OpenCV order and the camera t"""
    fx = cam[0]
    fy = cam[1] if t in (2, 3) else cam[0]
    p = cv_project(xyz, cam[4:7], cam[7:10], fx, fy, cam[2], cam[3], cam[10:15])
    return np.asarray(uv, np.float64) - p


# ------------------------------------------------------------------------------------------------ mpmath functors
def mp_rod(r):
    t2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2]
    if t2 == 0:
        a, b = mp.mpf(1), mp.mpf(1) / 2
    else:
        th = mp.sqrt(t2)
        a, b = mp.sin(th) / th, (1 - mp.cos(th)) / t2
    x, y, z = r
    return [[1 - b * (y * y + z * z), b * x * y - a * z, b * x * z + a * y], [b * x * y + a * z, 1 - b * (x * x + z * z), b * y * z - a * x],
            [b * x * z - a * y, b * y * z + a * x, 1 - b * (x * x + y * y)]]


def mp_brown(x, y, k1, k2, k3, p1, p2):
    r2 = x * x + y * y
    rad = 1 + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2
    return x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x), y * rad + 2 * p2 * x * y + p1 * (r2 + 2 * y * y)


def mp_matvec(R, v):
    return [R[i][0] * v[0] + R[i][1] * v[1] + R[i][2] * v[2] for i in range(3)]


def mp_ba_ray(t, a, uv):
    """a = intr(9) [disp(3)] ext(6) ray(3), the functor argument order"""
    a = [mp.mpf(v) for v in a]
    intr = a[:9]
    o = 9
    disp = None
    if t == 3:
        disp = a[9:12]
        o = 12
    ext, ray = a[o : o + 6], a[o + 6 : o + 9]
    R = mp_rod(ext[:3])
    n = ray
    if t != 1:
        nn = mp.sqrt(ray[0] ** 2 + ray[1] ** 2 + ray[2] ** 2)
        n = [v / nn for v in ray]
    X = mp_matvec(R, n)
    fx = intr[0]
    fy = intr[1] if t == 2 else intr[0]
    Z = X[2] + (disp[0] + disp[1] * fx + disp[2] * fx * fx if t == 3 else 0)
    x, y = X[0] / Z, X[1] / Z
    if t == 0:
        xd, yd = x, y
    else:
        xd, yd = mp_brown(x, y, *intr[4:9])
    return [mp.mpf(float(uv[0])) - (fx * xd + intr[2]), mp.mpf(float(uv[1])) - (fy * yd + intr[3])]


def mp_ba_pt(t, a, uv, xyz):
    """a = intr(9) [disp(3)] ext(6) tlw(6)"""
    a = [mp.mpf(v) for v in a]
    intr = a[:9]
    o = 9
    disp = None
    if t == 3:
        disp = a[9:12]
        o = 12
    ext, tlw = a[o : o + 6], a[o + 6 : o + 12]
    Xw = [mp.mpf(float(v)) for v in xyz]
    Xl = [p + q for p, q in zip(mp_matvec(mp_rod(tlw[:3]), Xw), tlw[3:])]
    X = mp_matvec(mp_rod(ext[:3]), Xl)
    fx, fy = intr[0], intr[1]
    Z = X[2] + (disp[0] + disp[1] * fx + disp[2] * fx * fx if t == 3 else 0)
    xd, yd = mp_brown(X[0] / Z, X[1] / Z, *intr[4:9])
    return [mp.mpf(float(uv[0])) - (fx * xd + intr[2]), mp.mpf(float(uv[1])) - (fy * yd + intr[3])]


def mp_krt(t, cam, ray1, uv2):
    cam = [mp.mpf(v) for v in cam]
    n = [mp.mpf(float(v)) for v in ray1]
    X = mp_matvec(mp_rod(cam[4:7]), n)
    fx = cam[0]
    fy = cam[1] if t in (2, 3) else cam[0]
    x, y = X[0] / X[2], X[1] / X[2]
    if t in (1, 3):
        x, y = mp_brown(x, y, *cam[10:15])
    return [mp.mpf(float(uv2[0])) - (fx * x + cam[2]), mp.mpf(float(uv2[1])) - (fy * y + cam[3])]


def mp_jac(fun, a):
    """d fun / d a[j] for all j by 60-digit central differences (mp.diff)"""
    a = [mp.mpf(float(v)) for v in a]
    J = np.zeros((2, len(a)))
    for j in range(len(a)):
        for i in range(2):
            J[i, j] = float(mp.diff(lambda s: fun(a[:j] + [s] + a[j + 1 :])[i], a[j]))
    return J


# ------------------------------------------------------------------------------------------------ sampling
def sample_view(rng, near_zero=False, near_pi=False):
    if near_zero:
        rvec = rng.normal(0, 1e-9, 3) * rng.integers(0, 2)
    elif near_pi:
        ax = rng.normal(size=3)
        rvec = ax / np.linalg.norm(ax) * (np.pi - rng.uniform(1e-3, 0.2))
    else:
        rvec = rng.normal(0, 0.6, 3)
    return rvec


def make_functor_kat(n=1000, n_mp=12, seed=20250217):
    rng = np.random.default_rng(seed)
    out = {}
    for t, name in enumerate(BA_TYPES):
        intr, ext, ray, disp, uv, res = [], [], [], [], [], []
        for i in range(n):
            rvec = sample_view(rng, near_zero=(i % 10 == 0), near_pi=(i % 10 == 1))
            R = cv2.Rodrigues(rvec)[0]
            f = rng.uniform(800, 6000)
            c = np.array([rng.uniform(500, 1000), rng.uniform(300, 600)])
            d5 = np.array([rng.normal(0, 0.1), rng.normal(0, 0.02), rng.normal(0, 0.005), rng.normal(0, 1e-3), rng.normal(0, 1e-3)])
            px = np.array([rng.uniform(-200, 2 * c[0] + 200), rng.uniform(-200, 2 * c[1] + 200)])
            dirc = np.array([(px[0] - c[0]) / f, (px[1] - c[1]) / f, 1.0])
            if t == 1 and i % 10 == 2:
                dirc = -dirc  # behind the camera: constant penalty branch
            r = R.T @ dirc * rng.uniform(0.5, 2.0)
            it = np.concatenate([[f, f * rng.uniform(0.9, 1.1), c[0], c[1]], d5])
            ex = np.concatenate([rvec, rng.normal(0, 1, 3)])
            dp = np.array([rng.normal(0, 0.02), rng.normal(0, 1e-5), rng.normal(0, 1e-9)])
            u = (px + rng.normal(0, 2, 2)).astype(np.float32)
            intr.append(it); ext.append(ex); ray.append(r); disp.append(dp); uv.append(u)
            res.append(ba_ray_cv(t, it, ex, r, dp, u))
        out[f"ba{t}_intr"], out[f"ba{t}_ext"], out[f"ba{t}_ray"] = np.array(intr), np.array(ext), np.array(ray)
        out[f"ba{t}_disp"], out[f"ba{t}_uv"], out[f"ba{t}_res_cv"] = np.array(disp), np.array(uv), np.array(res)
        # exact Jacobians (all global coordinates, functor argument order) for the first n_mp generic + 2 edge samples
        idx = [k for k in range(n) if k % 10 not in (1, 2)][: n_mp]
        J, R60 = [], []
        for k in idx:
            a = np.concatenate([intr[k]] + ([disp[k]] if t == 3 else []) + [ext[k], ray[k]])
            J.append(mp_jac(lambda s: mp_ba_ray(t, s, uv[k]), a))
            R60.append([float(v) for v in mp_ba_ray(t, a, uv[k])])
        out[f"ba{t}_mp_idx"], out[f"ba{t}_jac_mp"], out[f"ba{t}_res_mp"] = np.array(idx), np.array(J), np.array(R60)
        print(name, "done", flush=True)
    # 2d-3d BA terms, types 0 (Reproj2d3dFactor) and 3 (Reproj2d3dDispFactor)
    for t in (0, 3):
        intr, ext, tlw, disp, uv, xyz, res = [], [], [], [], [], [], []
        for i in range(n // 4):
            rvec = sample_view(rng, near_zero=(i % 10 == 0))
            R = cv2.Rodrigues(rvec)[0]
            f = rng.uniform(800, 6000)
            c = np.array([rng.uniform(500, 1000), rng.uniform(300, 600)])
            d5 = np.array([rng.normal(0, 0.1), rng.normal(0, 0.02), rng.normal(0, 0.005), rng.normal(0, 1e-3), rng.normal(0, 1e-3)])
            tl = np.concatenate([rng.normal(0, 0.5, 3), rng.normal(0, 10, 3)])
            px = np.array([rng.uniform(0, 2 * c[0]), rng.uniform(0, 2 * c[1])])
            Xl = R.T @ np.array([(px[0] - c[0]) / f, (px[1] - c[1]) / f, 1.0]) * rng.uniform(20, 100)
            Xw = cv2.Rodrigues(tl[:3])[0].T @ (Xl - tl[3:])
            it = np.concatenate([[f, f * rng.uniform(0.95, 1.05), c[0], c[1]], d5])
            ex = np.concatenate([rvec, rng.normal(0, 1, 3)])
            dp = np.array([rng.normal(0, 0.02), rng.normal(0, 1e-5), rng.normal(0, 1e-9)])
            u = (px + rng.normal(0, 2, 2)).astype(np.float32)
            intr.append(it); ext.append(ex); tlw.append(tl); disp.append(dp); uv.append(u); xyz.append(Xw)
            res.append(ba_pt_cv(t, it, ex, tl, dp, u, Xw))
        k = f"pt{t}"
        out[k + "_intr"], out[k + "_ext"], out[k + "_tlw"], out[k + "_disp"] = np.array(intr), np.array(ext), np.array(tlw), np.array(disp)
        out[k + "_uv"], out[k + "_xyz"], out[k + "_res_cv"] = np.array(uv), np.array(xyz), np.array(res)
        J = []
        for q in range(6):
            a = np.concatenate([intr[q]] + ([disp[q]] if t == 3 else []) + [ext[q], tlw[q]])
            J.append(mp_jac(lambda s: mp_ba_pt(t, s, uv[q], xyz[q]), a))
        out[k + "_jac_mp"] = np.array(J)
        print("pt", t, "done", flush=True)
    # KRT 2d-2d
    for t, name in enumerate(KRT_TYPES):
        tt = {"F": 0, "FDist": 1, "Fxfy": 2, "FxfyDist": 3}[name]
        cam, refK, refd, uv1, uv2, res, und = [], [], [], [], [], [], []
        for i in range(n):
            rvec = sample_view(rng, near_zero=(i % 4 == 0)) * (0.15 if i % 4 else 1.0)
            f1 = rng.uniform(1000, 3000)
            c1 = np.array([960.0, 540.0]) if i % 2 else np.array([rng.uniform(500, 1000), rng.uniform(300, 600)])
            d1 = np.array([rng.normal(0, 0.15), rng.normal(0, 0.02), rng.normal(0, 1e-3), rng.normal(0, 1e-3), rng.normal(0, 0.005)])
            if tt in (0, 2):
                d1 = d1 * (i % 3 == 0)
            f2 = f1 * rng.uniform(0.7, 1.4)
            c2 = np.array([rng.uniform(500, 1000), rng.uniform(300, 600)])
            cm = np.concatenate([[f2, f2 * rng.uniform(0.95, 1.05), c2[0], c2[1]], rvec, rng.normal(0, 1, 3), d1 * rng.uniform(0.5, 1.5, 5)])
            p1 = np.array([rng.uniform(-30, 2 * c1[0] + 30), rng.uniform(-30, 2 * c1[1] + 30)]).astype(np.float32)
            p2 = np.array([rng.uniform(0, 2 * c2[0]), rng.uniform(0, 2 * c2[1])]).astype(np.float32)
            rk = np.array([f1, f1 * (rng.uniform(0.95, 1.05) if tt in (2, 3) else 1.0), c1[0], c1[1]])
            r, u = krt_cv(tt, cm, rk, d1, p1, p2)
            cam.append(cm); refK.append(rk); refd.append(d1); uv1.append(p1); uv2.append(p2); res.append(r); und.append(u)
        k = f"krt{tt}"
        out[k + "_cam"], out[k + "_refK"], out[k + "_refd"] = np.array(cam), np.array(refK), np.array(refd)
        out[k + "_uv1"], out[k + "_uv2"], out[k + "_res_cv"], out[k + "_und_cv"] = np.array(uv1), np.array(uv2), np.array(res), np.array(und)
        # exact Jacobians over the 15 coordinates; ray1 recomputed here from the cv2 undistorted pixel
        J, idx = [], []
        for q in range(n):
            if len(idx) >= n_mp:
                break
            if np.all(res[q] == 0):
                continue
            K1 = np.array([[refK[q][0], 0, refK[q][2]], [0, refK[q][1], refK[q][3]], [0, 0, 1.0]])
            ptv = np.array([und[q][0], und[q][1], 1.0], np.float64)
            ray1 = np.linalg.inv(K1) @ ptv
            if tt != 2:
                ray1 = ray1 / np.linalg.norm(ray1)
            J.append(mp_jac(lambda s: mp_krt(tt, s, ray1, uv2[q]), cam[q]))
            idx.append(q)
        out[k + "_mp_idx"], out[k + "_jac_mp"] = np.array(idx), np.array(J)
        print(name, "done", flush=True)
    # KRT 2d-3d through cv2.projectPoints (values and cv2's own analytic Jacobian)
    for tt in (0, 2):
        cam, uv, xyz, res, jac = [], [], [], [], []
        for i in range(n // 4):
            rvec = sample_view(rng)
            f = rng.uniform(1000, 3000)
            cm = np.concatenate([[f, f * rng.uniform(0.95, 1.05), 960, 540], rvec, rng.normal(0, 2, 3),
                                 [rng.normal(0, 0.1), rng.normal(0, 0.02), rng.normal(0, 1e-3), rng.normal(0, 1e-3), rng.normal(0, 0.005)]])
            R = cv2.Rodrigues(rvec)[0]
            px = np.array([rng.uniform(0, 1920), rng.uniform(0, 1080)])
            Xc = np.array([(px[0] - 960) / f, (px[1] - 540) / f, 1.0]) * rng.uniform(10, 80)
            X = R.T @ (Xc - cm[7:10])
            u = (px + rng.normal(0, 2, 2)).astype(np.float32)
            cam.append(cm); uv.append(u); xyz.append(X); res.append(krt3d_cv(tt, cm, u, X))
            fy = cm[1] if tt == 2 else cm[0]
            K = np.array([[cm[0], 0, cm[2]], [0, fy, cm[3]], [0, 0, 1.0]])
            _, jj = cv2.projectPoints(X.reshape(1, 1, 3), cm[4:7], cm[7:10], K, cm[10:15])
            jac.append(jj[:, :15])  # d(u,v)/d(rvec3, tvec3, fx, fy, cx, cy, k1,k2,p1,p2,k3)
        k = f"krt3d{tt}"
        out[k + "_cam"], out[k + "_uv"], out[k + "_xyz"], out[k + "_res_cv"], out[k + "_jac_cv"] = map(np.array, (cam, uv, xyz, res, jac))
    np.savez_compressed(os.path.join(HERE, "functor_kat.npz"), **out)


def make_opencv_kat(seed=7):
    rng = np.random.default_rng(seed)
    r = np.concatenate([rng.normal(0, 1, (200, 3)), rng.normal(0, 1e-6, (20, 3)), np.zeros((1, 3)), rng.normal(0, 1e-17, (5, 3))])
    R = np.array([cv2.Rodrigues(v)[0] for v in r])
    # matrix -> vector, incl. tiny rotations (s < 1e-5 -> 0) and near-pi
    ax = rng.normal(size=(30, 3))
    ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    r2 = np.concatenate([rng.normal(0, 1, (200, 3)), rng.normal(0, 3e-6, (20, 3)), ax * (np.pi - rng.uniform(1e-7, 1e-2, (30, 1)))])
    R2 = np.array([cv2.Rodrigues(v)[0] for v in r2])
    rinv = np.array([cv2.Rodrigues(m)[0].ravel() for m in R2])
    # undistortPoints
    K4 = np.array([[rng.uniform(1000, 3000), 0, rng.uniform(500, 1000), rng.uniform(300, 600)] for _ in range(300)])
    K4[:, 1] = K4[:, 0] * rng.uniform(0.95, 1.05, 300)
    d = np.stack([rng.normal(0, 0.15, 300), rng.normal(0, 0.03, 300), rng.normal(0, 2e-3, 300), rng.normal(0, 2e-3, 300), rng.normal(0, 0.01, 300)], 1)
    d[:20] = 0
    uv = np.stack([rng.uniform(-50, 2100, 300), rng.uniform(-50, 1200, 300)], 1).astype(np.float32)
    und = []
    for i in range(300):
        K = np.array([[K4[i, 0], 0, K4[i, 2]], [0, K4[i, 1], K4[i, 3]], [0, 0, 1.0]])
        und.append(cv2.undistortPoints(uv[i].reshape(1, 1, 2), K, d[i], None, K).reshape(2))
    np.savez_compressed(os.path.join(HERE, "opencv_kat.npz"), rod_r=r, rod_R=R, rodinv_R=R2, rodinv_r=rinv, und_K4=K4, und_d=d, und_uv=uv,
                        und_out=np.array(und, np.float32))


# ------------------------------------------------------------------------------------------------ minimiser KATs (scipy)
def np_rod(r):
    return cv2.Rodrigues(np.asarray(r, np.float64))[0]


def make_lm_kat(seed=99):
    """Gauge-fixed problems solved to the true minimum by scipy (independent minimiser, cv2 functors)."""
    out = {}
    # (a) reloc queries, F and FDist
    for name, tt in (("F", 0), ("FDist", 1)):
        b = synth.make_reloc_batch(6, factor_type=tt, seed=seed + tt, n_min=40, n_max=90, outlier_frac=0.0)
        sols, costs = [], []
        for q in range(b.B):
            o0, o1 = b.match_offset[q], b.match_offset[q + 1]
            ref, init = b.ref_cam[q], b.init_cam[q]
            # local frame: reference camera is the identity (krt_optimizer.cc:269-286); here R_init == R_ref so rvec0 = 0
            free = [0, 4, 5, 6] + ([10] if tt == 1 else [])
            x0 = np.zeros(15)
            x0[:4] = init[:4]
            x0[10:15] = init[16:21]

            def fun(z):
                cam = x0.copy()
                cam[free] = z
                r = [krt_cv(tt, cam, ref[:4], ref[16:21], b.uv_ref[i], b.uv_cur[i])[0] for i in range(o0, o1)]
                return np.concatenate(r)

            s = least_squares(fun, x0[free], method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15, x_scale=np.array([1e3, 1, 1, 1] + ([1] if tt == 1 else [])))
            full = x0.copy()
            full[free] = s.x
            sols.append(full)
            costs.append(s.cost)
        out[f"reloc{tt}_offset"], out[f"reloc{tt}_uv_ref"], out[f"reloc{tt}_uv_cur"] = b.match_offset, b.uv_ref, b.uv_cur
        out[f"reloc{tt}_ref"], out[f"reloc{tt}_init"], out[f"reloc{tt}_sol"], out[f"reloc{tt}_cost"] = b.ref_cam, b.init_cam, np.array(sols), np.array(costs)
        print("lm reloc", name, costs, flush=True)
    # (b) a tiny BA (free gauge): scipy from the same start; compare gauge-invariant quantities and the cost
    for t in (0, 1):
        p = synth.make_ba_scene(6, 60, "ring", factor_type=t, seed=seed + 10 + t, neighbours=6)
        V, P = p.V, p.P
        ray0 = np.zeros((P, 3))
        cnt = np.zeros(P)
        for k in range(p.M):  # Pix2Ray (ptzray_optimizer.cc:768-797)
            i = p.obs_view[k]
            K = np.array([[p.intr[i, 0], 0, p.intr[i, 2]], [0, p.intr[i, 1], p.intr[i, 3]], [0, 0, 1.0]])
            v = np.linalg.inv(np_rod(p.ext[i, :3])) @ np.linalg.inv(K) @ np.array([p.obs_uv[k, 0], p.obs_uv[k, 1], 1.0], np.float64)
            ray0[p.obs_track[k]] += v / np.linalg.norm(v)
            cnt[p.obs_track[k]] += 1
        ray0 /= cnt[:, None]
        ray0 /= np.linalg.norm(ray0, axis=1, keepdims=True)
        nci = 1 if t == 0 else 2  # live intrinsics: fx (+k1); fy is tied

        def unpack(z):
            intr = p.intr.copy()
            ext = p.ext.copy()
            intr[:, 0] = z[:V]
            o = V
            if t == 1:
                intr[:, 4] = z[o : o + V]
                o += V
            ext[:, :3] = z[o : o + 3 * V].reshape(V, 3)
            o += 3 * V
            return intr, ext, z[o:].reshape(P, 3)

        def fun(z):
            intr, ext, ray = unpack(z)
            r = [np.sqrt(p.track_weight[p.obs_track[k]]) * ba_ray_cv(t, intr[p.obs_view[k]], ext[p.obs_view[k]], ray[p.obs_track[k]], np.zeros(3), p.obs_uv[k])
                 for k in range(p.M)]
            return np.concatenate(r)

        z0 = np.concatenate([p.intr[:, 0]] + ([p.intr[:, 4]] if t == 1 else []) + [p.ext[:, :3].ravel(), ray0.ravel()])
        s = least_squares(fun, z0, method="trf", xtol=1e-15, ftol=1e-15, gtol=1e-13, x_scale="jac", max_nfev=400)
        intr, ext, ray = unpack(s.x)
        k = f"ba{t}"
        for nm in ("intr", "ext", "obs_uv", "obs_view", "obs_track", "track_weight"):
            out[f"{k}_{nm}"] = getattr(p, nm)
        out[f"{k}_ray0"], out[f"{k}_sol_intr"], out[f"{k}_sol_ext"], out[f"{k}_sol_ray"], out[f"{k}_cost"] = ray0, intr, ext, ray, s.cost
        print("lm ba", t, s.cost, s.status, s.nfev, flush=True)
    np.savez_compressed(os.path.join(HERE, "lm_kat.npz"), **out)


def make_epnp_kat(seed=31, cases=48):
    """cv2.solvePnP(SOLVEPNP_EPNP) on random non-planar point sets (double object points, float32 pixels as the reference's
    cv::Point2f): the first half exact projections (EPnP is then independent of the SVD's sign conventions), the second half with
    0.5 px noise.  Flat layout: n[c], K[c,9], dist[c,5], obj/pix concatenated with offsets, R[c,9], t[c,3], gt_R, gt_t."""
    rng = np.random.default_rng(seed)
    ns, Ks, ds, objs, pixs, Rs, ts, gR, gt, noise = [], [], [], [], [], [], [], [], [], []
    for c in range(cases):
        n = int(rng.integers(5, 40))
        rvec = rng.normal(0, 0.8, 3)
        R, _ = cv2.Rodrigues(rvec)
        t = np.array([rng.normal(0, 2), rng.normal(0, 2), rng.uniform(20, 60)])
        f = rng.uniform(1000, 3000)
        K = np.array([[f, 0, 960], [0, f * rng.uniform(0.95, 1.05), 540], [0, 0, 1.0]])
        dist = np.array([rng.normal(0, 0.05), rng.normal(0, 0.01), 0, 0, 0]) if c % 2 else np.zeros(5)
        obj = rng.uniform(-10, 10, (n, 3))
        sig = 0.0 if c < cases // 2 else 0.5
        pix, _ = cv2.projectPoints(obj, rvec, t, K, dist)
        pix = (pix.reshape(-1, 2) + rng.normal(0, sig, (n, 2))).astype(np.float32)
        ok, rv, tv = cv2.solvePnP(obj, pix, K, dist, flags=cv2.SOLVEPNP_EPNP)
        assert ok
        Rcv, _ = cv2.Rodrigues(rv)
        ns.append(n); Ks.append(K.ravel()); ds.append(dist); objs.append(obj); pixs.append(pix); Rs.append(Rcv.ravel()); ts.append(tv.ravel())
        gR.append(R.ravel()); gt.append(t); noise.append(sig)
    np.savez_compressed(os.path.join(HERE, "epnp_kat.npz"), n=np.array(ns, np.int32), K=np.array(Ks), dist=np.array(ds), obj=np.concatenate(objs),
                        pix=np.concatenate(pixs), R=np.array(Rs), t=np.array(ts), gt_R=np.array(gR), gt_t=np.array(gt), noise=np.array(noise),
                        cv2_version=np.array(cv2.__version__))


def make_shared_kat():
    """SetSharedIntrinsics (ptzray_optimizer.cc:497-505, :640-650, :821-848): the tiny BA scenes of lm_kat.npz again, now with (a) ALL views and
    (b) views {1, 3, 4} sharing ONE intrinsics block -- the block of the group's first view, its fixed cx, cy, dist included -- minimised
    by scipy with the cv2 functors.  Pins the oracle's restatement of shared blocks (cost, shared focal, relative rotations)."""
    k = np.load(os.path.join(HERE, "lm_kat.npz"))
    out = {}
    for t in (0, 1):
        intr0, ext0, uv, oview, otrack, w, ray0 = (k[f"ba{t}_{n}"] for n in ("intr", "ext", "obs_uv", "obs_view", "obs_track", "track_weight", "ray0"))
        V, P, M = len(intr0), len(ray0), len(uv)
        for name, ids in (("all", np.zeros(V, np.int32)), ("some", np.array([10, 7, 11, 7, 7, 12][:V], np.int32))):
            rep = np.array([int(np.nonzero(ids == ids[i])[0][0]) for i in range(V)])
            groups = sorted(set(rep.tolist()))           # one intrinsics block per group, at its first view
            gi = {g: n for n, g in enumerate(groups)}
            G = len(groups)

            def unpack(z):
                intr = intr0[rep].copy()                 # every view reads its group's block: fixed cx, cy, dist of the first view too
                intr[:, 0] = np.array([z[gi[rep[i]]] for i in range(V)])
                o = G
                if t == 1:
                    intr[:, 4] = np.array([z[o + gi[rep[i]]] for i in range(V)])
                    o += G
                ext = ext0.copy()
                ext[:, :3] = z[o : o + 3 * V].reshape(V, 3)
                return intr, ext, z[o + 3 * V :].reshape(P, 3)

            def fun(z):
                intr, ext, ray = unpack(z)
                return np.concatenate([np.sqrt(w[otrack[q]]) * ba_ray_cv(t, intr[oview[q]], ext[oview[q]], ray[otrack[q]], np.zeros(3), uv[q]) for q in range(M)])

            z0 = np.concatenate([intr0[groups, 0]] + ([intr0[groups, 4]] if t == 1 else []) + [ext0[:, :3].ravel(), ray0.ravel()])
            s = least_squares(fun, z0, method="trf", xtol=1e-15, ftol=1e-15, gtol=1e-13, x_scale="jac", max_nfev=600)
            intr, ext, ray = unpack(s.x)
            key = f"ba{t}_{name}"
            out[f"{key}_ids"], out[f"{key}_sol_intr"], out[f"{key}_sol_ext"], out[f"{key}_cost"] = ids, intr, ext, s.cost
            print("shared", key, s.cost, s.status, s.nfev, flush=True)
    np.savez_compressed(os.path.join(HERE, "shared_kat.npz"), **out)


def make_georef_kat(seed=123):
    """RunGeoreferencing (run_ptz_ba.cc:131-145): ray terms + annotated 2d-3d points with a free T_l_w, EVERY view annotated (two points
    each), minimised by scipy with the cv2 functors.  fy of an annotated view is driven by its 2d-3d terms only.  The gauge
    of the local frame is free, so the test compares the cost, T-invariant world-frame rotations R_i R_lw and the focal lengths."""
    out = {}
    for t in (0, 1):
        p = synth.make_ba_scene(6, 60, "ring", factor_type=t, seed=seed + t, neighbours=6, num_pts3d=12, pts3d_views=6)
        V, P, M, A = p.V, p.P, p.M, p.A
        ray0 = np.zeros((P, 3))
        cnt = np.zeros(P)
        for k in range(M):  # Pix2Ray
            i = p.obs_view[k]
            K = np.array([[p.intr[i, 0], 0, p.intr[i, 2]], [0, p.intr[i, 1], p.intr[i, 3]], [0, 0, 1.0]])
            v = np.linalg.inv(np_rod(p.ext[i, :3])) @ np.linalg.inv(K) @ np.array([p.obs_uv[k, 0], p.obs_uv[k, 1], 1.0], np.float64)
            ray0[p.obs_track[k]] += v / np.linalg.norm(v)
            cnt[p.obs_track[k]] += 1
        ray0 /= cnt[:, None]
        ray0 /= np.linalg.norm(ray0, axis=1, keepdims=True)

        def unpack(z):
            intr, ext = p.intr.copy(), p.ext.copy()
            intr[:, 0] = z[:V]
            intr[:, 1] = z[V : 2 * V]          # fy: only the 2d-3d terms read it (the ray factors tie fy := fx)
            o = 2 * V
            if t == 1:
                intr[:, 4] = z[o : o + V]
                o += V
            ext[:, :3] = z[o : o + 3 * V].reshape(V, 3)
            o += 3 * V
            return intr, ext, z[o : o + 3 * P].reshape(P, 3), z[o + 3 * P :]

        def fun(z):
            intr, ext, ray, tlw = unpack(z)
            r = [np.sqrt(p.track_weight[p.obs_track[k]]) * ba_ray_cv(t, intr[p.obs_view[k]], ext[p.obs_view[k]], ray[p.obs_track[k]], np.zeros(3), p.obs_uv[k])
                 for k in range(M)]
            r += [ba_pt_cv(t, intr[p.pt_view[a]], ext[p.pt_view[a]], tlw, np.zeros(3), p.pt_uv[a], p.pt_xyz[a]) for a in range(A)]
            return np.concatenate(r)

        z0 = np.concatenate([p.intr[:, 0], p.intr[:, 1]] + ([p.intr[:, 4]] if t == 1 else []) + [p.ext[:, :3].ravel(), ray0.ravel(), p.tlw0])
        s = least_squares(fun, z0, method="trf", xtol=1e-15, ftol=1e-15, gtol=1e-13, x_scale="jac", max_nfev=400)
        intr, ext, ray, tlw = unpack(s.x)
        key = f"geo{t}"
        for nm in ("intr", "ext", "obs_uv", "obs_view", "obs_track", "track_weight", "pt_uv", "pt_xyz", "pt_view", "tlw0"):
            out[f"{key}_{nm}"] = getattr(p, nm)
        out[f"{key}_sol_intr"], out[f"{key}_sol_ext"], out[f"{key}_sol_tlw"], out[f"{key}_cost"] = intr, ext, tlw, s.cost
        print("georef", key, s.cost, s.status, s.nfev, flush=True)
    np.savez_compressed(os.path.join(HERE, "georef_kat.npz"), **out)


def make_distdisp_kat(seed=77):
    """PTZRayDistDispFactor (ptzray_optimizer.cc:202-264): the global displacement polynomial disp[3] shared by every residual block, on a
    small scene where it is observable (synth.make_distdisp_scene: zoom sweep, true displacement in the data), minimised by scipy with the
    cv2 functors.  Free gauge: the test compares the cost, relative rotations and the displacement at a mid-range focal length."""
    t = 3
    p = synth.make_distdisp_scene(V=8, P=90, seed=seed)
    V, P, M = p.V, p.P, p.M
    ray0 = np.zeros((P, 3))
    cnt = np.zeros(P)
    for k in range(M):  # Pix2Ray
        i = p.obs_view[k]
        K = np.array([[p.intr[i, 0], 0, p.intr[i, 2]], [0, p.intr[i, 1], p.intr[i, 3]], [0, 0, 1.0]])
        v = np.linalg.inv(np_rod(p.ext[i, :3])) @ np.linalg.inv(K) @ np.array([p.obs_uv[k, 0], p.obs_uv[k, 1], 1.0], np.float64)
        ray0[p.obs_track[k]] += v / np.linalg.norm(v)
        cnt[p.obs_track[k]] += 1
    ray0 /= cnt[:, None]
    ray0 /= np.linalg.norm(ray0, axis=1, keepdims=True)

    def unpack(z):
        intr, ext = p.intr.copy(), p.ext.copy()
        intr[:, 0] = z[:V]
        intr[:, 4] = z[V : 2 * V]
        ext[:, :3] = z[2 * V : 5 * V].reshape(V, 3)
        return intr, ext, z[5 * V : 5 * V + 3 * P].reshape(P, 3), z[5 * V + 3 * P :]

    def fun(z):
        intr, ext, ray, disp = unpack(z)
        return np.concatenate([np.sqrt(p.track_weight[p.obs_track[k]]) * ba_ray_cv(t, intr[p.obs_view[k]], ext[p.obs_view[k]], ray[p.obs_track[k]], disp, p.obs_uv[k])
                               for k in range(M)])

    z0 = np.concatenate([p.intr[:, 0], p.intr[:, 4], p.ext[:, :3].ravel(), ray0.ravel(), np.zeros(3)])
    s = least_squares(fun, z0, method="trf", xtol=1e-15, ftol=1e-15, gtol=1e-13, x_scale="jac", max_nfev=600)
    intr, ext, ray, disp = unpack(s.x)
    out = {}
    for nm in ("intr", "ext", "obs_uv", "obs_view", "obs_track", "track_weight"):
        out[f"dd_{nm}"] = getattr(p, nm)
    out["dd_sol_intr"], out["dd_sol_ext"], out["dd_sol_disp"], out["dd_cost"], out["dd_status"], out["dd_nfev"] = intr, ext, disp, s.cost, s.status, s.nfev
    print("distdisp", s.cost, s.status, s.nfev, disp, flush=True)
    np.savez_compressed(os.path.join(HERE, "distdisp_kat.npz"), **out)


if __name__ == "__main__":
    if "--distdisp-only" in sys.argv:
        make_distdisp_kat()
        sys.exit(0)
    if "--georef-only" in sys.argv:
        make_georef_kat()
        sys.exit(0)
    if "--shared-only" in sys.argv:
        make_shared_kat()
        sys.exit(0)
    make_epnp_kat()
    make_opencv_kat()
    make_functor_kat()
    make_lm_kat()
    make_shared_kat()
    make_georef_kat()
    make_distdisp_kat()
    print("golden vectors written to", HERE)
