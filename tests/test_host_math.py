"""The kernels' per-observation closed forms (csrc/ptz_math.cuh), compiled for the host, against the oracle's
dual-number Jacobians on the golden inputs.  CPU only; catches derivation mistakes before any GPU time is spent.
Tolerance: 1e-9 relative to the Jacobian scale (north-star gradient tolerance), values 1e-9 px."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from ptz_calib_b200 import abi, problem
from ptz_calib_b200.abi import as_ptr, f32, f64

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hm(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hm") / "libhost_math.so")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", os.path.join(HERE, "host_math_check.cpp"), "-o", so],
                   check=True)
    return C.CDLL(so)


def d(a):
    return as_ptr(f64(a), C.c_double)


@pytest.mark.parametrize("t", [0, 1, 2, 3])
def test_ba_obs_matches_oracle(hm, orc, functor_kat, t):
    k = functor_kat
    ncv = 5 if t == 0 else 6
    wo = ncv + 3 + (3 if t == 3 else 0)
    worst_r = worst_j = 0.0
    for i in range(0, len(k[f"ba{t}_uv"]), 3):
        intr, ext, ray, disp, uv = k[f"ba{t}_intr"][i], k[f"ba{t}_ext"][i], k[f"ba{t}_ray"][i], k[f"ba{t}_disp"][i], k[f"ba{t}_uv"][i]
        r, J = np.zeros(2), np.zeros((2, wo))
        hm.hm_ba_obs(C.c_int(t), d(intr), d(ext), d(ray), d(disp), as_ptr(f32(uv), C.c_float), d(r), as_ptr(J, C.c_double))
        p = problem.BAProblem(factor_type=t, intr=intr[None], ext=ext[None], obs_uv=uv[None], obs_view=[0], obs_track=[0], track_weight=[1.0], ray0=ray[None])
        e = orc.ba_eval(p, disp=disp if t == 3 else None)
        worst_r = max(worst_r, np.abs(r - e.residuals[0]).max() / max(1e3, np.abs(e.residuals[0]).max()))
        scale = max(np.abs(e.jac_obs[0]).max(), 1e-300)
        worst_j = max(worst_j, np.abs(J - e.jac_obs[0]).max() / scale)
    assert worst_r < 1e-12, worst_r
    assert worst_j < 1e-10, worst_j


@pytest.mark.parametrize("t", [0, 3])
def test_ba_pt_matches_oracle(hm, orc, functor_kat, t):
    k = functor_kat
    p_ = f"pt{t}"
    nci = 2 if t == 0 else 3
    for q in range(0, len(k[p_ + "_uv"]), 2):
        intr, ext, tlw, disp, uv, xyz = (k[p_ + s][q] for s in ("_intr", "_ext", "_tlw", "_disp", "_uv", "_xyz"))
        r, Jc, Jt, Jd = np.zeros(2), np.zeros((2, 6)), np.zeros((2, 6)), np.zeros((2, 3))
        hm.hm_ba_pt(C.c_int(1 if t == 3 else 0), d(intr), d(ext), d(tlw), d(disp), as_ptr(f32(uv), C.c_float), d(xyz), d(r), as_ptr(Jc, C.c_double),
                    as_ptr(Jt, C.c_double), as_ptr(Jd, C.c_double))
        prob = problem.BAProblem(factor_type=t, intr=intr[None], ext=ext[None], obs_uv=np.zeros((0, 2), np.float32), obs_view=[], obs_track=[], track_weight=[],
                                 pt_uv=uv[None], pt_xyz=xyz[None], pt_view=[0], tlw0=tlw)
        e = orc.ba_eval(prob, disp=disp if t == 3 else None)
        want = e.jac_pts[0]
        got = np.concatenate([Jc[:, :2], Jc[:, 2:3] if nci == 3 else np.zeros((2, 0)), Jc[:, 3:6], Jt] + ([Jd] if t == 3 else []), axis=1)
        assert np.abs(r - e.residuals[0]).max() < 1e-9
        assert np.abs(got - want).max() / np.abs(want).max() < 1e-10


@pytest.mark.parametrize("t", [0, 1, 2, 3])
def test_krt_obs_matches_oracle(hm, orc, functor_kat, t):
    k = functor_kat
    p = f"krt{t}"
    nf = len(abi.KRT_FREE[t])
    for i in range(0, len(k[p + "_uv1"]), 3):
        cam, refK, refd, uv1, uv2 = (k[p + s][i] for s in ("_cam", "_refK", "_refd", "_uv1", "_uv2"))
        r, J = np.zeros(2), np.zeros((2, nf))
        hm.hm_krt_obs(C.c_int(t), d(cam), d(refK), d(refd), as_ptr(f32(uv1), C.c_float), as_ptr(f32(uv2), C.c_float), d(r), as_ptr(J, C.c_double))
        ref21 = np.zeros(21)
        ref21[:4], ref21[4:13], ref21[16:21] = refK, np.eye(3).ravel(), refd
        res, jac, cost, g = orc.reloc_eval(t, uv1[None], uv2[None], ref21, cam)
        assert np.abs(r - res[0]).max() < 1e-9, (i, r, res[0])
        scale = max(np.abs(jac[0]).max(), 1e-300)
        assert np.abs(J - jac[0]).max() / scale < 1e-10, (i, J, jac[0])


def test_rodrigues_jac_small_angles(hm, orc):
    rng = np.random.default_rng(5)
    for mag in (0.0, 1e-12, 1e-6, 1e-3, 0.05, 0.0999, 0.1001, 0.5, 3.0):
        w = rng.normal(size=3)
        w = w / np.linalg.norm(w) * mag
        R, dR = np.zeros(9), np.zeros(27)
        hm.hm_rodrigues_jac(d(w), d(R), as_ptr(dR, C.c_double))
        assert np.abs(R.reshape(3, 3) - orc.rodrigues(w)).max() < 3e-16
        h = 1e-6
        for kk in range(3):
            e = np.zeros(3)
            e[kk] = h
            num = (orc.rodrigues(w + e) - orc.rodrigues(w - e)) / (2 * h)
            assert np.abs(dR[9 * kk : 9 * kk + 9].reshape(3, 3) - num).max() < 1e-9


@pytest.mark.parametrize("t", [0, 1, 2, 3])
def test_krt_obs3d_matches_oracle(hm, orc, functor_kat, t):
    """Factor2d3dDist / Factor2d3dFxfyDist (cv::projectPoints, OpenCV coefficient order, camera t): the oracle's values are
    pinned to cv2.projectPoints by the golden vectors; here the kernel's closed form is checked against the oracle."""
    k = functor_kat
    src = "krt3d0" if t in (0, 1) else "krt3d2"
    free = abi.KRT_FREE[t]
    for i in range(0, len(k[src + "_uv"]), 2):
        cam, uv, xyz = k[src + "_cam"][i], k[src + "_uv"][i], k[src + "_xyz"][i]
        r, J = np.zeros(2), np.zeros((2, len(free)))
        hm.hm_krt_obs3d(C.c_int(t), d(cam), as_ptr(f32(uv), C.c_float), d(xyz), d(r), as_ptr(J, C.c_double))
        want_r = orc.krt_2d3d_factor(t, cam, uv, xyz)
        want_J = orc.krt_2d3d_jac(t, cam, uv, xyz)[:, free]
        assert np.abs(r - want_r).max() < 1e-9
        assert np.abs(J - want_J).max() / np.abs(want_J).max() < 1e-10
