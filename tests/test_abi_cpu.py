"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol include/ptzcalib_b200.h
declares, the ctypes mirror has the header's layout, and the product path fails loudly without a GPU (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from ptz_calib_b200 import abi, lib, synth
import ptz_calib_b200 as ptz

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "ptzcalib_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ptz\w*)\s*\(", src)) - {"ptzba_handle"})


def test_library_exports_every_declared_symbol():
    if not os.path.exists(lib.SO_PATH):
        import __graft_entry__ as g

        g.build()
    L = lib.load()
    names = header_functions()
    assert len(names) >= 16
    for n in names:
        assert hasattr(L, n), n
    assert sorted(lib.EXPORTS) == names


def test_struct_layout_matches_header_defaults():
    o = abi.SolverOptions()
    lib.load().ptz_solver_options_default(C.byref(o))
    d = ptz.default_options()
    for name, _ in abi.SolverOptions._fields_:
        if name == "linear_solver":
            continue
        assert getattr(o, name) == getattr(d, name), name
    assert o.function_tolerance == 1e-6 and o.gradient_tolerance == 1e-10 and o.parameter_tolerance == 1e-8
    assert o.initial_trust_region_radius == 1e4 and o.min_relative_decrease == 1e-3 and o.max_num_consecutive_invalid_steps == 5


def test_invalid_arguments_are_rejected_without_touching_the_gpu():
    L = lib.load()
    p = synth.make_config(1, scale=0.2)
    c = p.to_c()
    r, arrs, log = ptz.problem.alloc_ba_result(p)
    o = ptz.default_options(max_num_iterations=0)  # CheckValid: max_iter_ <= 0 (ptzray_optimizer.cc:521)
    assert L.ptzba_solve(C.byref(c), C.byref(o), C.byref(r)) == abi.PTZ_ERR_INVALID
    c.num_views = 0  # CheckValid: num_cams_ == 0
    o = ptz.default_options()
    assert L.ptzba_solve(C.byref(c), C.byref(o), C.byref(r)) == abi.PTZ_ERR_INVALID
    c = p.to_c()
    bad = p.obs_view.copy()
    bad[0] = p.V
    c.obs_view = bad.ctypes.data_as(C.POINTER(C.c_int32))
    assert L.ptzba_solve(C.byref(c), C.byref(o), C.byref(r)) == abi.PTZ_ERR_INVALID
    c = p.to_c()
    # SetSharedIntrinsics: a permutation of distinct ids shares nothing and is accepted (then fails for lack of a device here);
    # shared blocks together with 2d-3d points are the one combination that is not built
    ids = np.arange(p.V, dtype=np.int32)[::-1].copy()
    c.shared_ic_id = ids.ctypes.data_as(C.POINTER(C.c_int32))
    assert L.ptzba_solve(C.byref(c), C.byref(o), C.byref(r)) == abi.PTZ_ERR_NO_DEVICE
    q = synth.make_config(1, scale=0.2, num_pts3d=6)
    q.shared_ic_id = np.zeros(q.V, np.int32)
    cq = q.to_c()
    rq, _, _ = ptz.problem.alloc_ba_result(q)
    assert L.ptzba_solve(C.byref(cq), C.byref(o), C.byref(rq)) == abi.PTZ_ERR_UNSUPPORTED
    c.factor_type = 7
    assert L.ptzba_solve(C.byref(c), C.byref(o), C.byref(r)) == abi.PTZ_ERR_INVALID
    # ptzgeo_init_tlw: null output, negative count, decreasing offsets
    tlw, used, off = np.zeros(6), C.c_int32(7), np.array([0, 4, 2], np.int64)
    cams, uv, xyz = np.zeros((2, 21)), np.zeros((4, 2), np.float32), np.zeros((4, 3))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    assert L.ptzgeo_init_tlw(C.c_int32(2), vp(cams), vp(off), vp(uv), vp(xyz), None, C.byref(used)) == abi.PTZ_ERR_INVALID
    assert L.ptzgeo_init_tlw(C.c_int32(-1), vp(cams), vp(off), vp(uv), vp(xyz), vp(tlw), C.byref(used)) == abi.PTZ_ERR_INVALID
    assert L.ptzgeo_init_tlw(C.c_int32(2), vp(cams), vp(off), vp(uv), vp(xyz), vp(tlw), C.byref(used)) == abi.PTZ_ERR_INVALID
    assert L.ptzgeo_init_tlw(C.c_int32(0), None, None, None, None, vp(tlw), C.byref(used)) == 0 and used.value == -1


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point reports PTZ_ERR_NO_DEVICE instead of computing on the host."""
    if ptz.device_count() > 0:
        pytest.skip("a GPU is present")
    p = synth.make_config(1, scale=0.2)
    with pytest.raises(lib.PtzLibraryError, match="-5"):
        ptz.ba_solve(p)
    with pytest.raises(lib.PtzLibraryError, match="-5"):
        ptz.ba_eval(p)
    b = synth.make_reloc_batch(4)
    with pytest.raises(lib.PtzLibraryError, match="-5"):
        ptz.reloc_solve_batch(b)


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "ptz-calib_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("test-only oracle binding (oracle/oracle.py)", "").replace("oracle binding", "") or f in ("abi.py", "problem.py"), f
