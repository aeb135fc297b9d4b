"""Host-side EPnP (include/ptzcalib_epnp.hpp, the cv::solvePnP(SOLVEPNP_EPNP) call of SetInitTransLocalToWorld,
ptzray_optimizer.cc:572) against golden vectors from cv2.solvePnP (tests/golden/epnp_kat.npz, made by make_golden.py)."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def epnp_results(tmp_path_factory):
    k = np.load(os.path.join(ROOT, "tests", "golden", "epnp_kat.npz"))
    d = tmp_path_factory.mktemp("epnp")
    exe, fin, fout = str(d / "epnp_check"), str(d / "in.bin"), str(d / "out.bin")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", os.path.join(ROOT, "tests", "cpp", "epnp_check.cpp"), "-o", exe], check=True)
    off = np.concatenate([[0], np.cumsum(k["n"])])
    with open(fin, "wb") as f:
        f.write(struct.pack("i", len(k["n"])))
        for c, n in enumerate(k["n"]):
            f.write(struct.pack("i", int(n)))
            f.write(k["K"][c].tobytes()); f.write(k["dist"][c].tobytes())
            f.write(np.ascontiguousarray(k["obj"][off[c]:off[c + 1]]).tobytes()); f.write(np.ascontiguousarray(k["pix"][off[c]:off[c + 1]]).tobytes())
    subprocess.run([exe, fin, fout], check=True)
    out = np.fromfile(fout, dtype=np.float64).reshape(len(k["n"]), 8, 13)
    return k, out


def pose_err(out, R, t):
    return max(np.abs(out[1:10] - R).max(), np.abs(out[10:13] - t).max() / np.abs(t).max())


def test_exact_projections_match_cv2(epnp_results):
    """On exact data all sign conventions of the control-point axes give the same pose: parity with cv2 to float32-pixel accuracy."""
    k, out = epnp_results
    ids = np.nonzero(k["noise"] == 0)[0]
    assert len(ids) >= 20
    for c in ids:
        assert out[c, 0, 0] == 1.0
        assert pose_err(out[c, 0], k["R"][c], k["t"][c]) < 2e-6, c
        assert pose_err(out[c, 0], k["gt_R"][c], k["gt_t"][c]) < 2e-6, c


def test_noisy_pixels_match_cv2_up_to_the_axis_signs(epnp_results):
    """With noise the estimate depends (at the 1e-4 level) on the sign an SVD happens to give the principal axes of the object
    points; cv2's pose must be reproduced by one of the 8 choices, and the default choice must be as close to the truth as cv2."""
    k, out = epnp_results
    ids = np.nonzero(k["noise"] > 0)[0]
    best = np.array([min(pose_err(out[c, s], k["R"][c], k["t"][c]) for s in range(8)) for c in ids])
    dflt = np.array([pose_err(out[c, 0], k["R"][c], k["t"][c]) for c in ids])
    assert np.median(best) < 1e-5 and best.max() < 2e-3
    assert dflt.max() < 1e-2
    mine = np.array([pose_err(out[c, 0], k["gt_R"][c], k["gt_t"][c]) for c in ids])
    cv = np.array([max(np.abs(k["R"][c] - k["gt_R"][c]).max(), np.abs(k["t"][c] - k["gt_t"][c]).max() / np.abs(k["gt_t"][c]).max()) for c in ids])
    assert np.median(mine) < 2.0 * np.median(cv) and mine.max() < 3.0 * cv.max()


def test_degenerate_inputs():
    """fewer than 4 points: refused; coplanar points: a finite answer through the pseudo-inverse (cvInvert CV_SVD)"""
    import tempfile

    d = tempfile.mkdtemp()
    exe, fin, fout = os.path.join(d, "e"), os.path.join(d, "i"), os.path.join(d, "o")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "cpp", "epnp_check.cpp"), "-o", exe], check=True)
    rng = np.random.default_rng(0)
    K = np.array([1500.0, 0, 960, 0, 1500, 540, 0, 0, 1])
    with open(fin, "wb") as f:
        f.write(struct.pack("i", 2))
        for n, planar in ((3, False), (12, True)):
            obj = rng.uniform(-5, 5, (n, 3))
            if planar:
                obj[:, 2] = 0
            X = obj + np.array([0, 0, 30.0])
            pix = (X[:, :2] / X[:, 2:] * 1500 + np.array([960, 540])).astype(np.float32)
            f.write(struct.pack("i", n)); f.write(K.tobytes()); f.write(np.zeros(5).tobytes()); f.write(obj.tobytes()); f.write(pix.tobytes())
    subprocess.run([exe, fin, fout], check=True)
    out = np.fromfile(fout, dtype=np.float64).reshape(2, 8, 13)
    assert out[0, 0, 0] == 0.0
    # (EPnP without its planar variant is unreliable on coplanar points — cv2 4.13 returns mirrored poses on such inputs too; the
    # caller's gates, ptzray_optimizer.cc:582-604, are what reject a bad initialisation.  Here: no NaN, no crash.)
    assert out[1, 0, 0] == 1.0 and np.isfinite(out[1, 0]).all()


def _rodrigues(r):
    th = np.linalg.norm(r)
    if th < 1e-12:
        return np.eye(3)
    k = r / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def test_init_trans_local_to_world_through_the_c_abi():
    """ptzgeo_init_tlw (PTZRayOptimizer::SetInitTransLocalToWorld, ptzray_optimizer.cc:562-633) is host code behind the C ABI: runs here
    without a device.  Exact projections of world points through T_i_l T_l_w must give T_l_w back; a view without points, one with fewer
    than 4 and one whose pixels no pose explains are passed over in order (:566-604)."""
    import ptz_calib_b200 as ptz

    rng = np.random.default_rng(5)
    Rlw, tlw = _rodrigues(np.array([0.3, -0.2, 0.5])), np.array([4.0, -3.0, 60.0])
    V = 5
    cams = np.zeros((V, 21))
    uv, xyz, view = [], [], []
    for i in range(V):
        Ri = _rodrigues(rng.normal(size=3) * 0.2)
        ti = np.zeros(3)                                          # a PTZ camera: no translation in the local frame
        fx = 1500.0 + 100 * i
        cams[i] = np.concatenate([[fx, fx * 1.01, 960, 540], Ri.ravel(), ti, [-0.05, 0.01, 0, 0, 0]])
        n = [0, 3, 8, 10, 12][i]
        # world points whose local-frame image falls inside the view: pick camera-frame points, pull them back to the world
        xc = np.stack([rng.uniform(-0.4, 0.4, n), rng.uniform(-0.25, 0.25, n), np.ones(n)], 1) * rng.uniform(40, 90, (n, 1))
        X = (Rlw.T @ ((Ri.T @ (xc - ti).T).T - tlw).T).T
        x, y = xc[:, 0] / xc[:, 2], xc[:, 1] / xc[:, 2]
        r2 = x * x + y * y
        d = 1 + cams[i, 16] * r2 + cams[i, 17] * r2 * r2
        uv.append(np.stack([fx * x * d + 960, fx * 1.01 * y * d + 540], 1))
        if i == 2:                                                # pixels unrelated to the points: no pose reprojects them within 300 px RMS
            uv[-1] = np.stack([rng.uniform(0, 1920, n), rng.uniform(0, 1080, n)], 1)
        xyz.append(X)
        view.append(np.full(n, i))
    uv, xyz, view = np.concatenate(uv), np.concatenate(xyz), np.concatenate(view)
    sh = rng.permutation(len(view))                               # any order of the rows: grouped by view inside
    got, used = ptz.init_trans_local_to_world(cams, uv[sh], xyz[sh], view[sh])
    assert used == 3
    assert np.abs(_rodrigues(got[:3]) - Rlw).max() < 1e-5
    assert np.abs(got[3:] - tlw).max() < 2e-3
    # no annotated view passes -> zeros, -1 (the reference leaves tlw_param_ at zero, :628-632)
    z, u0 = ptz.init_trans_local_to_world(cams[:2], uv[view < 2], xyz[view < 2], view[view < 2])
    assert u0 == -1 and not z.any()
    z, u0 = ptz.init_trans_local_to_world(cams, np.zeros((0, 2)), np.zeros((0, 3)), np.zeros(0, int))
    assert u0 == -1 and not z.any()
