"""ctypes binding of the CPU oracle (oracle/ptz_oracle.cpp).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the product package."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from ptz_calib_b200 import abi, problem  # noqa: E402  (layout definitions only)
from ptz_calib_b200.abi import as_ptr, f32, f64  # noqa: E402

_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libptz_oracle.so")
    srcs = [os.path.join(_HERE, "ptz_oracle.cpp"), os.path.join(_HERE, "tracks_oracle.cpp"), os.path.join(_ROOT, "include", "ptzcalib_b200.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(so) < os.path.getmtime(p) for p in srcs if os.path.exists(p))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libptz_oracle.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_ptzba_solve.restype = C.c_int
        L.orc_ptzba_eval_mode.restype = C.c_int
        L.orc_ptzreloc_solve_batch.restype = C.c_int
        L.orc_ptzreloc_eval_mode.restype = C.c_int
        L.orc_ptzba_time_jacobian.restype = C.c_double
        L.orc_ptzba_time_jacobian.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_num_threads.restype = C.c_int
        L.orc_tracks_build.restype = C.c_int
        L.orc_tracks_flatten.restype = C.c_int
        _LIB = L
    return _LIB


def options(**kw):
    return problem.default_options(**kw)


def num_threads():
    return lib().orc_num_threads()


# ---------------------------------------------------------------------------------------- primitives
def rodrigues(r):
    R = np.zeros(9)
    lib().orc_rodrigues(as_ptr(f64(r), C.c_double), as_ptr(R, C.c_double))
    return R.reshape(3, 3)


def rodrigues_inv(R):
    r = np.zeros(3)
    lib().orc_rodrigues_inv(as_ptr(f64(R).reshape(9), C.c_double), as_ptr(r, C.c_double))
    return r


def undistort_point(uv, K4, dist5):
    out = np.zeros(2, np.float32)
    lib().orc_undistort_point(as_ptr(f32(uv), C.c_float), as_ptr(f64(K4), C.c_double), as_ptr(f64(dist5), C.c_double), as_ptr(out, C.c_float))
    return out


def ba_ray_factor(ftype, intr, ext, ray, uv, disp=None):
    res = np.zeros(2)
    lib().orc_ba_ray_factor(C.c_int(ftype), as_ptr(f64(intr), C.c_double), as_ptr(f64(ext), C.c_double), as_ptr(f64(ray), C.c_double),
                            as_ptr(None if disp is None else f64(disp), C.c_double), as_ptr(f32(uv), C.c_float), as_ptr(res, C.c_double))
    return res


def ba_pt_factor(ftype, intr, ext, tlw, uv, xyz, disp=None):
    res = np.zeros(2)
    lib().orc_ba_pt_factor(C.c_int(ftype), as_ptr(f64(intr), C.c_double), as_ptr(f64(ext), C.c_double), as_ptr(f64(tlw), C.c_double),
                           as_ptr(None if disp is None else f64(disp), C.c_double), as_ptr(f32(uv), C.c_float), as_ptr(f64(xyz), C.c_double),
                           as_ptr(res, C.c_double))
    return res


def krt_2d2d_factor(ftype, cam15, refK4, refd5, uv1, uv2):
    res = np.zeros(2)
    lib().orc_krt_2d2d_factor(C.c_int(ftype), as_ptr(f64(cam15), C.c_double), as_ptr(f64(refK4), C.c_double), as_ptr(f64(refd5), C.c_double),
                              as_ptr(f32(uv1), C.c_float), as_ptr(f32(uv2), C.c_float), as_ptr(res, C.c_double))
    return res


def krt_2d3d_factor(ftype, cam15, uv, xyz):
    res = np.zeros(2)
    lib().orc_krt_2d3d_factor(C.c_int(ftype), as_ptr(f64(cam15), C.c_double), as_ptr(f32(uv), C.c_float), as_ptr(f64(xyz), C.c_double), as_ptr(res, C.c_double))
    return res


def krt_2d3d_jac(ftype, cam15, uv, xyz):
    J = np.zeros((2, 15))
    lib().orc_krt_2d3d_jac(C.c_int(ftype), as_ptr(f64(cam15), C.c_double), as_ptr(f32(uv), C.c_float), as_ptr(f64(xyz), C.c_double), as_ptr(J, C.c_double))
    return J


def krt_to_local(ref21, init21):
    out = np.zeros(15)
    lib().orc_krt_to_local(as_ptr(f64(ref21), C.c_double), as_ptr(f64(init21), C.c_double), as_ptr(out, C.c_double))
    return out


def reloc_reproj_error(ftype, ref21, local15, uv_ref=None, uv_cur=None, pt_uv=None, pt_xyz=None):
    """(Cal2d2dReprojError, Cal2d3dReprojError) of KRTOptimizer at local15"""
    e22, e23 = C.c_double(0), C.c_double(0)
    n = 0 if uv_ref is None else len(uv_ref)
    m = 0 if pt_uv is None else len(pt_uv)
    a = [None if x is None else np.ascontiguousarray(x, dtype=dt) for x, dt in ((uv_ref, np.float32), (uv_cur, np.float32), (pt_uv, np.float32), (pt_xyz, np.float64))]
    lib().orc_ptzreloc_reproj_error(C.c_int(ftype), as_ptr(f64(ref21), C.c_double), as_ptr(f64(local15), C.c_double), C.c_int(n), as_ptr(a[0], C.c_float),
                                    as_ptr(a[1], C.c_float), C.c_int(m), as_ptr(a[2], C.c_float), as_ptr(a[3], C.c_double), C.byref(e22), C.byref(e23))
    return e22.value, e23.value


def krt_to_world(ftype, ref21, local15):
    out = np.zeros(21)
    lib().orc_krt_to_world(C.c_int(ftype), as_ptr(f64(ref21), C.c_double), as_ptr(f64(local15), C.c_double), as_ptr(out, C.c_double))
    return out


# ---------------------------------------------------------------------------------------- BA
def ba_eval(prob: problem.BAProblem, disp=None, jacobian_mode=0) -> problem.BAEval:
    c = prob.to_c()
    e, (res, jo, jp, g) = problem.alloc_ba_eval(prob)
    d = None if disp is None else f64(disp)
    rc = lib().orc_ptzba_eval_mode(C.byref(c), as_ptr(d, C.c_double), C.byref(e), C.c_int(jacobian_mode))
    if rc != 0:
        raise RuntimeError(f"orc_ptzba_eval rc={rc}")
    return problem.BAEval(e.cost, res, jo, jp, g)


def ba_init_rays(prob: problem.BAProblem):
    c = prob.to_c()
    ray = np.zeros((prob.P, 3))
    rc = lib().orc_ptzba_init_rays(C.byref(c), as_ptr(ray, C.c_double))
    if rc != 0:
        raise RuntimeError(f"orc_ptzba_init_rays rc={rc}")
    return ray


def ba_solve(prob: problem.BAProblem, opt=None, **kw):
    opt = opt or options(**kw)
    c = prob.to_c()
    r, arrs, log = problem.alloc_ba_result(prob)
    rc = lib().orc_ptzba_solve(C.byref(c), C.byref(opt), C.byref(r))
    if rc != 0:
        return rc, None
    return rc, problem.unpack_ba_result(prob, r, arrs, log)


def ba_time_jacobian(prob: problem.BAProblem, jacobian_mode, threads, repeats=1):
    c = prob.to_c()
    return lib().orc_ptzba_time_jacobian(C.addressof(c), jacobian_mode, threads, repeats)


# ---------------------------------------------------------------------------------------- reloc
def reloc_eval(ftype, uv_ref, uv_cur, ref21, local15, jacobian_mode=0):
    uv_ref, uv_cur = f32(uv_ref).reshape(-1, 2), f32(uv_cur).reshape(-1, 2)
    N = uv_ref.shape[0]
    nf = len(abi.KRT_FREE[ftype])
    res, jac, g, cost = np.zeros((N, 2)), np.zeros((N, 2, nf)), np.zeros(nf), C.c_double(0)
    rc = lib().orc_ptzreloc_eval_mode(C.c_int(ftype), C.c_int(N), as_ptr(uv_ref, C.c_float), as_ptr(uv_cur, C.c_float), as_ptr(f64(ref21), C.c_double),
                                      as_ptr(f64(local15), C.c_double), as_ptr(res, C.c_double), as_ptr(jac, C.c_double), C.byref(cost),
                                      as_ptr(g, C.c_double), C.c_int(jacobian_mode))
    if rc != 0:
        raise RuntimeError(f"orc_ptzreloc_eval rc={rc}")
    return res, jac, cost.value, g


def reloc_solve_batch(batch: problem.RelocBatch, opt=None, **kw) -> problem.RelocResult:
    opt = opt or options(**kw)
    c = batch.to_c()
    r, arrs = problem.alloc_reloc_result(batch.B)
    rc = lib().orc_ptzreloc_solve_batch(C.byref(c), C.byref(opt), C.byref(r))
    if rc != 0:
        raise RuntimeError(f"orc_ptzreloc_solve_batch rc={rc}")
    return problem.RelocResult(**arrs)


def reloc_solve_one(ftype, uv_ref, uv_cur, ref21, init21, opt=None, **kw):
    opt = opt or options(**kw)
    uv_ref, uv_cur = f32(uv_ref).reshape(-1, 2), f32(uv_cur).reshape(-1, 2)
    N = uv_ref.shape[0]
    out = np.zeros(15)
    cap = 512
    log = (abi.IterLog * cap)()
    n, term = C.c_int(0), C.c_int(0)
    lib().orc_ptzreloc_solve_one(C.c_int(ftype), C.c_int(N), as_ptr(uv_ref, C.c_float), as_ptr(uv_cur, C.c_float), as_ptr(f64(ref21), C.c_double),
                                 as_ptr(f64(init21), C.c_double), C.byref(opt), as_ptr(out, C.c_double), log, C.c_int(cap), C.byref(n), C.byref(term))
    rows = [dict(cost=l.cost, cost_change=l.cost_change, gradient_max_norm=l.gradient_max_norm, step_norm=l.step_norm, relative_decrease=l.relative_decrease,
                 trust_region_radius=l.trust_region_radius, step_is_successful=l.step_is_successful) for l in log[: n.value]]
    return out, rows, term.value


# ---------------------------------------------------------------------------------------- tracks (tracks_oracle.cpp)
def tracks_build(matches, min_track_length=4):
    """TracksBuilder::Build/Filter/ExportToSTL restated on the CPU; track ids are the reference's union-by-rank roots"""
    from ptz_calib_b200 import tracks as T

    rc, t = T.call_build(lib().orc_tracks_build, matches, min_track_length, "orc_tracks_build")
    assert rc == 0, rc
    return t


def tracks_flatten(tracks, views):
    from ptz_calib_b200 import tracks as T

    rc, o = T.call_flatten(lib().orc_tracks_flatten, tracks, views)
    assert rc == 0, rc
    return o
