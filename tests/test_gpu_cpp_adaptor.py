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


def build_exe(tmp_path, name="adaptor_check"):
    exe = str(tmp_path / name)
    so_dir = os.path.dirname(lib.SO_PATH)
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-o", exe, "-L" + so_dir, "-lptzcalib_b200",
                    "-ldl", "-Wl,-rpath," + so_dir], check=True)
    return exe


ORACLE_SO = os.path.join(ROOT, "oracle", "libptz_oracle.so")


def write_iba_input(path, p, max_iter, both_directions):
    gt, V = p.gt, p.V
    cams = np.zeros((V, 21))
    for i in range(V):
        cams[i, 0] = cams[i, 1] = gt["f"][i]
        cams[i, 2:4] = gt["c"][i]
        cams[i, 4:13] = gt["R"][i].ravel()
    with open(path, "wb") as f:
        f.write(struct.pack("4i", V, p.M, max_iter, both_directions))
        for a in (cams, p.obs_view, p.obs_track, p.obs_uv):
            f.write(np.ascontiguousarray(a).tobytes())


def parse_iba_output(path, V):
    """head(8) | V x (registered, krt21) | n, n x (kind, a, b) of the GPU driver | oracle: (rc, ok, registered, reproj, n), events, V x 22"""
    out = np.fromfile(path, dtype=np.float64)
    head, rows = out[:8], out[8 : 8 + 22 * V].reshape(V, 22)
    pos = 8 + 22 * V
    nt = int(out[pos])
    trace = out[pos + 1 : pos + 1 + 3 * nt].reshape(nt, 3).astype(np.int64)
    pos += 1 + 3 * nt
    if pos >= out.size:
        return head, rows, trace, None, None, None
    oh = out[pos : pos + 5]
    ne = int(oh[4])
    otrace = out[pos + 5 : pos + 5 + 3 * ne].reshape(ne, 3).astype(np.int64)
    orows = out[pos + 5 + 3 * ne :].reshape(V, 22)
    return head, rows, trace, oh, otrace, orows


def test_iba_oracle_driver_on_cpu(tmp_path):
    """CPU: the restated reference driver (oracle/iba_oracle.cpp) with the oracle's own BA / KRT / tracks registers a synthetic ring
    from unknown cameras and lands on the ground truth -- the checker is itself checked before the GPU driver is compared with it."""
    exe = build_exe(tmp_path, "iba_check")
    p = synth.make_config(1, scale=0.4)
    fin, fout = str(tmp_path / "iba_in.bin"), str(tmp_path / "iba_out.bin")
    write_iba_input(fin, p, 100, 1)
    r = subprocess.run([exe, fin, fout, ORACLE_SO, "oracle-only"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    _, _, trace, oh, otrace, orows = parse_iba_output(fout, p.V)
    assert trace.size == 0 and oh[0] == 0.0 and oh[1] == 1.0 and int(oh[2]) == p.V, r.stdout
    kinds = otrace[:, 0].tolist()
    assert kinds[0] == 1 and kinds[1] == 2 and otrace[1, 1] == 1 and kinds[2] == 3 and kinds[-1] == 3
    registered_in_order = [int(a) for k, a, b in otrace if k == 4 and b >= 0]
    assert len(set(registered_in_order)) == len(registered_in_order) == p.V - 2
    gt = p.gt
    assert np.abs(orows[:, 1] / gt["f"] - 1).max() < 5e-3
    R = orows[:, 5:14].reshape(p.V, 3, 3)
    for b in range(1, p.V):
        rel_est, rel_gt = R[b] @ R[0].T, gt["R"][b] @ gt["R"][0].T
        assert np.arccos(np.clip((np.trace(rel_est @ rel_gt.T) - 1) / 2, -1, 1)) < 2e-3


def test_adaptor_compiles_without_gpu(tmp_path):
    """CPU: the adaptor header is valid C++17 against the C ABI and links to the library."""
    assert os.path.exists(build_exe(tmp_path))
    assert os.path.exists(build_exe(tmp_path, "iba_check"))
    assert os.path.exists(build_exe(tmp_path, "georef_check"))


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
    assert np.abs(krt[:21] - rr.cam[0]).max() <= 1e-9
    # Cal2d2dReprojError before / after Solve, Cal2d3dReprojError without points (krt_optimizer.cc:406-500)
    ref_local = b.ref_cam[0].copy(); ref_local[4:13] = np.eye(3).ravel(); ref_local[13:16] = 0
    e0, _ = orc.reloc_reproj_error(abi.PTZ_KRT_F, ref_local, orc.krt_to_local(b.ref_cam[0], b.init_cam[0]), b.uv_ref, b.uv_cur)
    assert abs(krt[21] - e0) <= 1e-9 * e0
    assert abs(krt[22] - rr.final_rms[0]) <= 1e-9 * rr.final_rms[0] and krt[22] < krt[21]
    assert krt[23] == -1.0


@pytest.mark.gpu
@pytest.mark.parametrize("both_directions", [1, 0])
def test_incremental_driver_registers_the_whole_scene(tmp_path, orc, both_directions):
    """PtzIncrementalOptimizer (mirror of ptz_incremental_optimizer.cc:39-441) on a synthetic ring: cameras start unknown, the seed
    pair comes from the confidence ranking, every other image is registered through batched KRT solves from its homography to a
    registered neighbour, with a global BA each time the model grew by 10 %.  Checks: every image registered, relative rotations
    and focal lengths at the ground truth to the noise level (the gauge is free: the seed image carries R = I)."""
    exe = build_exe(tmp_path, "iba_check")
    p = synth.make_config(1, scale=0.5)
    gt = p.gt
    V = p.V
    fin, fout = str(tmp_path / "iba_in.bin"), str(tmp_path / "iba_out.bin")
    write_iba_input(fin, p, 100, both_directions)
    r = subprocess.run([exe, fin, fout, ORACLE_SO], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    head, rows, trace, orc_head, orc_trace, orc_rows = parse_iba_output(fout, V)
    assert head[0] == 1.0, r.stdout
    # decision-for-decision against the CPU restatement of the reference's SEQUENTIAL driver (oracle/iba_oracle.cpp) on the same
    # tables: seed pair, LM iterations of the two-view BA, every FindNextImages list (length, head), every registration attempt and
    # the registered neighbour it succeeded from, every global BA (outcome, model size), un-register events -- in the same order
    assert orc_head[0] == 0.0 and orc_head[1] == 1.0, r.stdout
    assert trace.shape == orc_trace.shape and (trace == orc_trace).all(), (trace.tolist(), orc_trace.tolist())
    assert (rows[:, 0] == orc_rows[:, 0]).all()
    both = rows[:, 0] == 1.0
    assert np.abs(rows[both, 1] - orc_rows[both, 1]).max() <= 1e-4                 # focal [px]
    assert np.abs(rows[both, 5:14] - orc_rows[both, 5:14]).max() <= 1e-6           # R entries
    assert abs(head[5] - orc_head[3]) <= 1e-6 * orc_head[3]                        # reprojection error of the last global BA
    reg = rows[:, 0] == 1.0
    if both_directions:
        assert reg.all(), r.stdout
    else:
        # the driver only registers image j from table entries (i registered -> j) (ptz_incremental_optimizer.cc:389): with one
        # direction per pair an image whose registered neighbours all have larger indices is never reached
        assert 2 <= reg.sum() < V, r.stdout
    assert head[2] >= 3 and head[3] >= reg.sum() - 2  # several global BAs; one reloc batch per registered image at least
    assert head[5] < 2.0  # final reprojection error of the last global BA [px] (noise 0.5 px, weighted)
    est_R = rows[:, 5:14].reshape(V, 3, 3)
    est_f = rows[:, 1]
    ids = np.nonzero(reg)[0]
    assert np.abs(est_f[ids] / gt["f"][ids] - 1).max() < 5e-3
    a = ids[0]
    for b in ids[1:]:
        rel_est = est_R[b] @ est_R[a].T
        rel_gt = gt["R"][b] @ gt["R"][a].T
        ang = np.arccos(np.clip((np.trace(rel_est @ rel_gt.T) - 1) / 2, -1, 1))
        assert ang < 2e-3, (a, b, ang)


@pytest.mark.gpu
def test_georeferencing_with_epnp_initialisation(tmp_path, orc):
    """run_ptz_ba.cc:142-145: BA with annotated 2d-3d points; T_l_w comes from EPnP on the first annotated view inside Solve
    (ptzray_optimizer.cc:562-633), the georeferencing BA refines it.  The EPnP start must sit at the ground truth up to the view's
    own initial error, and the solve must end where the Python path ends when it is handed the same start."""
    exe = build_exe(tmp_path, "georef_check")
    p = synth.make_config(1, scale=0.5, num_pts3d=18, pts3d_views=3)
    cams = np.zeros((p.V, 21))
    for i in range(p.V):
        cams[i, :4] = p.intr[i, :4]
        cams[i, 4:13] = orc.rodrigues(p.ext[i, :3]).ravel()
        cams[i, 13:16] = p.ext[i, 3:]
        cams[i, 16:21] = p.intr[i, 4:9]
    fin, fout = str(tmp_path / "g_in.bin"), str(tmp_path / "g_out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("5i", p.V, p.M, p.A, p.factor_type, 200))
        for a in (cams, p.obs_view, p.obs_track, p.obs_uv, p.pt_view, p.pt_uv, p.pt_xyz):
            f.write(np.ascontiguousarray(a).tobytes())
    r = subprocess.run([exe, fin, fout], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = np.fromfile(fout, dtype=np.float64)
    head, cw = out[:12], out[12:].reshape(p.V, 21)
    assert head[0] == 1.0, r.stdout
    tlw0, gt = head[6:12], p.gt["tlw"]
    # rotation within ~2 degrees (0.5 deg initial error per axis of the view + focal error), translation within a few % of the range
    assert np.abs(tlw0[:3] - gt[:3]).max() < 0.05, (tlw0, gt)
    assert np.abs(tlw0[3:] - gt[3:]).max() < 0.1 * np.linalg.norm(gt[3:]) + 3.0, (tlw0, gt)
    assert head[4] < 2.0  # final 2d-3d reprojection RMS [px]
    # the same initialisation through the C ABI (ptzgeo_init_tlw), as the Python mirror's from_matches uses it
    t_abi, used = ptz.init_trans_local_to_world(cams, p.pt_uv, p.pt_xyz, p.pt_view)
    assert used >= 0 and np.abs(t_abi - tlw0).max() <= 1e-9 * (1 + np.abs(tlw0).max())
    q = synth.make_config(1, scale=0.5, num_pts3d=18, pts3d_views=3)
    q.tlw0 = tlw0.copy()
    want = ptz.ba_solve(q, max_num_iterations=200)
    assert want.converged and int(head[1]) == want.num_iterations
    assert abs(head[2] - want.final_reproj_error_all) <= 1e-7 * want.final_reproj_error_all
    assert abs(head[4] - want.final_reproj_error_2d3d) <= 1e-6 * max(want.final_reproj_error_2d3d, 1e-3)
    assert np.abs(cw - want.cams_world).max() <= 1e-5
