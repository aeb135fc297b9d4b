"""The header-only C++ adaptor (include/ptzcalib_b200.hpp: PTZRayOptimizer / KRTOptimizer / TracksBuilder with the
reference's signatures) compiled with g++, linked to the CUDA library, run on the GPU and compared with the Python path."""
import os
import struct
import subprocess

import numpy as np
import pytest

import ptz_calib_b200 as ptz
from ptz_calib_b200 import abi, lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_exe(tmp_path):
    exe = str(tmp_path / "adaptor_check")
    so_dir = os.path.dirname(lib.SO_PATH)
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "cpp", "adaptor_check.cpp"), "-o", exe, "-L" + so_dir, "-lptzcalib_b200",
                    "-Wl,-rpath," + so_dir], check=True)
    return exe


def test_adaptor_compiles_without_gpu(tmp_path):
    """CPU: the adaptor header is valid C++17 against the C ABI and links to the library."""
    assert os.path.exists(build_exe(tmp_path))


@pytest.mark.gpu
@pytest.mark.parametrize("t", [abi.PTZ_BA_PTZRAY, abi.PTZ_BA_PTZRAY_DIST])
def test_adaptor_matches_python_path(tmp_path, orc, t):
    exe = build_exe(tmp_path)
    p = synth.make_config(1, scale=0.3, factor_type=t)
    b = synth.make_reloc_batch(1, n_min=120, n_max=120)
    cams = np.zeros((p.V, 21))
    for i in range(p.V):
        cams[i, :4] = p.intr[i, :4]
        cams[i, 4:13] = orc.rodrigues(p.ext[i, :3]).ravel()
        cams[i, 13:16] = p.ext[i, 3:]
        cams[i, 16:21] = p.intr[i, 4:9]
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("6i", p.V, p.P, p.M, t, 200, b.N))
        for a in (cams, p.obs_view, p.obs_track, p.obs_uv, b.ref_cam[0], b.init_cam[0], b.uv_ref, b.uv_cur):
            f.write(np.ascontiguousarray(a).tobytes())
    r = subprocess.run([exe, fin, fout], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = np.fromfile(fout, dtype=np.float64)
    head, cw, krt = out[:10], out[10 : 10 + 21 * p.V].reshape(p.V, 21), out[10 + 21 * p.V :]
    want = ptz.ba_solve(p, max_num_iterations=200)
    assert head[0] == 1.0 and want.converged
    assert int(head[1]) == want.num_iterations
    assert abs(head[2] - want.final_reproj_error_all) <= 1e-9 * want.final_reproj_error_all
    assert abs(head[3] - want.final_reproj_error_2d2d) <= 1e-9 * want.final_reproj_error_2d2d
    assert int(head[4]) == p.P and int(head[5]) == p.M  # TracksBuilder recovered every track; one Ray per observation
    assert np.abs(cw - want.cams_world).max() <= 1e-7
    assert head[6] == 0.0 and head[7] == 1.0  # capped solve: false, cameras untouched
    rr = ptz.reloc_solve_batch(b)
    assert head[8] == float(rr.success[0]) == 1.0 and int(head[9]) == int(rr.num_iter[0])
    assert np.abs(krt - rr.cam[0]).max() <= 1e-9
