"""PTZ-BA on the GPU (through the C ABI) against the CPU oracle.  Run on the B200 box: pytest -m gpu.

Tolerances are the north star's: residuals and gradients 1e-9 relative, final cost 1e-6 relative, refined
parameters 1e-6 rad / 1e-4 px focal (BASELINE.json).  The oracle's Jacobian here is the exact (dual-number) one;
the Ceres-CENTRAL emulation is compared separately with its documented noise floor."""
import os

import numpy as np
import pytest

import ptz_calib_b200 as ptz
from conftest import relerr
from ptz_calib_b200 import abi, synth

pytestmark = pytest.mark.gpu

TYPES = [abi.PTZ_BA_PTZRAY, abi.PTZ_BA_PTZRAY_DIST, abi.PTZ_BA_PTZRAY_FXFY_DIST, abi.PTZ_BA_PTZRAY_DIST_DISP]


def small_scene(t, **kw):
    if t == abi.PTZ_BA_PTZRAY_DIST:
        return synth.make_config(2, scale=0.25, **kw)
    return synth.make_config(1, scale=0.3, factor_type=t, **kw)


@pytest.mark.parametrize("t", TYPES)
def test_eval_matches_oracle(orc, t):
    p = small_scene(t)
    p = p.with_params(ray=orc.ba_init_rays(p))
    disp = np.array([0.01, -2e-6, 3e-10]) if t == abi.PTZ_BA_PTZRAY_DIST_DISP else None
    got, want = ptz.ba_eval(p, disp=disp), orc.ba_eval(p, disp=disp)
    assert np.abs(got.residuals - want.residuals).max() <= 1e-9 * max(1.0, np.abs(want.residuals).max())
    # column-wise scale: the disp columns (x f, x f^2) are orders of magnitude larger than the rest
    cs = np.abs(want.jac_obs).max(axis=(0, 1))
    assert (np.abs(got.jac_obs - want.jac_obs).max(axis=(0, 1)) <= 1e-9 * np.maximum(cs, 1e-300)).all()
    assert abs(got.cost - want.cost) <= 1e-12 * want.cost
    nd = p.V * p.ncv + 3 * p.P  # cameras + rays; then disp(3)
    assert np.abs(got.gradient[:nd] - want.gradient[:nd]).max() <= 1e-9 * np.abs(want.gradient[:nd]).max()
    assert (np.abs(got.gradient[nd:] - want.gradient[nd:]) <= 1e-9 * np.abs(want.gradient[nd:])).all()
    # against Ceres' own numeric Jacobian: at its noise floor (a few e-9 of the gradient scale); the disp columns are excluded
    # (Ceres' minimum step is a finite perturbation there, see tests/test_oracle_golden.py)
    num = orc.ba_eval(p, disp=disp, jacobian_mode=1)
    assert np.abs(got.gradient[:nd] - num.gradient[:nd]).max() <= 2e-8 * np.abs(num.gradient[:nd]).max()


@pytest.mark.parametrize("t", TYPES)
def test_eval_with_annotated_points(orc, t):
    p = small_scene(t, num_pts3d=12)
    p = p.with_params(ray=orc.ba_init_rays(p))
    disp = np.array([0.01, -2e-6, 3e-10]) if t == abi.PTZ_BA_PTZRAY_DIST_DISP else None
    got, want = ptz.ba_eval(p, disp=disp), orc.ba_eval(p, disp=disp)
    assert got.residuals.shape == (p.M + p.A, 2)
    assert np.abs(got.residuals - want.residuals).max() <= 1e-9 * max(1.0, np.abs(want.residuals).max())
    cs = np.abs(want.jac_pts).max(axis=(0, 1))
    assert (np.abs(got.jac_pts - want.jac_pts).max(axis=(0, 1)) <= 1e-9 * np.maximum(cs, 1e-300)).all()
    assert abs(got.cost - want.cost) <= 1e-12 * want.cost
    gs = np.maximum(np.abs(want.gradient), 1e-9 * np.abs(want.gradient[: p.V * p.ncv]).max())
    assert (np.abs(got.gradient - want.gradient) <= 1e-8 * gs).all()


def test_init_rays_match_pix2ray(orc):
    p = small_scene(abi.PTZ_BA_PTZRAY)
    r = ptz.ba_solve(p, max_num_iterations=1, function_tolerance=0.0)  # rays after one step are not the initial ones; use eval instead
    e = ptz.ba_eval(p)  # ray0 is None: the library computes Pix2Ray itself, the oracle too
    w = orc.ba_eval(p)
    assert np.abs(e.residuals - w.residuals).max() <= 1e-9 * np.abs(w.residuals).max()
    assert r.num_iterations >= 1


def rel_rot(ext, orc):
    R = np.array([orc.rodrigues(e[:3]) for e in ext])
    return R @ R[0].T


def check_solve(orc, p, got, want, label=""):
    assert got.termination == want.termination, (label, got.termination, want.termination)
    assert got.num_iterations == want.num_iterations, (label, got.num_iterations, want.num_iterations)
    assert got.num_successful_steps == want.num_successful_steps
    assert abs(got.initial_cost - want.initial_cost) <= 1e-11 * want.initial_cost
    assert abs(got.final_cost - want.final_cost) <= 1e-6 * want.final_cost, (label, got.final_cost, want.final_cost)
    assert got.num_residuals == want.num_residuals
    for a in ("init_reproj_error_all", "final_reproj_error_all", "final_reproj_error_2d2d"):
        assert abs(getattr(got, a) - getattr(want, a)) <= 1e-6 * getattr(want, a), a
    # per-iteration table (cost, radius, accept/reject decisions)
    assert len(got.log) == len(want.log)
    for lg, lw in zip(got.log, want.log):
        assert lg["step_is_successful"] == lw["step_is_successful"]
        assert abs(lg["cost"] - lw["cost"]) <= 1e-7 * lw["cost"]
        assert abs(lg["trust_region_radius"] - lw["trust_region_radius"]) <= 1e-4 * lw["trust_region_radius"]
    # refined parameters: 1e-4 px focal, 1e-6 rad on the gauge-invariant relative rotations and on the raw rvecs
    assert np.abs(got.intr[:, 0] - want.intr[:, 0]).max() <= 1e-4, np.abs(got.intr[:, 0] - want.intr[:, 0]).max()
    assert np.abs(got.intr[:, 4] - want.intr[:, 4]).max() <= 1e-7
    assert np.abs(rel_rot(got.ext, orc) - rel_rot(want.ext, orc)).max() <= 1e-6
    assert np.abs(got.ext - want.ext).max() <= 1e-6
    assert np.abs(got.ray - want.ray).max() <= 1e-6
    assert np.abs(got.cams_world - want.cams_world)[:, 4:13].max() <= 1e-6
    if p.factor_type == abi.PTZ_BA_PTZRAY_DIST_DISP:
        # disp = (d0, d1, d2) multiplies (1, f, f^2): compare the displacement it produces at a typical focal
        f = np.array([1.0, 1800.0, 1800.0 ** 2])
        assert abs((got.disp - want.disp) @ f) <= 1e-6 * max(1.0, abs(want.disp @ f))


@pytest.mark.parametrize("t", TYPES[:3])
def test_solve_matches_oracle(orc, t):
    p = small_scene(t)
    got = ptz.ba_solve(p, max_num_iterations=200)
    rc, want = orc.ba_solve(p, max_num_iterations=200)
    assert rc == 0
    check_solve(orc, p, got, want, f"type{t}")
    assert got.converged  # PTZRayOptimizer::Solve returns true


@pytest.mark.parametrize("vranks", [2, 5, 8])
def test_sharded_cg_protocol_with_virtual_ranks(orc, monkeypatch, vranks):
    """The multi-GPU linear solver (rows of the reduced system sharded across ranks, LL words pushed into the other ranks'
    inboxes, two-level reduction) run on ONE GPU with the grid split into virtual ranks: same answers as the oracle and as
    the single-rank kernel."""
    p = synth.make_config(4, scale=0.05)  # V = 50: every virtual rank owns a few rows
    monkeypatch.setenv("PTZ_CG_VRANKS", str(vranks))
    got = ptz.ba_solve(p, max_num_iterations=200)
    monkeypatch.delenv("PTZ_CG_VRANKS")
    rc, want = orc.ba_solve(p, max_num_iterations=200)
    assert rc == 0
    check_solve(orc, p, got, want, f"vranks{vranks}")
    for t in TYPES[:3]:  # the three live-column layouts (4, 5, 6 per view), a fixed number of LM steps
        q = synth.make_config(4, scale=0.05, factor_type=t)
        monkeypatch.setenv("PTZ_CG_VRANKS", str(vranks))
        a = ptz.ba_solve(q, max_num_iterations=8)
        monkeypatch.delenv("PTZ_CG_VRANKS")
        b = ptz.ba_solve(q, max_num_iterations=8)
        assert a.num_iterations == b.num_iterations and abs(a.final_cost - b.final_cost) <= 1e-9 * b.final_cost
        assert np.abs(a.ext - b.ext).max() <= 1e-8 and np.abs(a.intr - b.intr).max() <= 1e-6


@pytest.mark.parametrize("t", [abi.PTZ_BA_PTZRAY, abi.PTZ_BA_PTZRAY_DIST])
def test_georef_solve_matches_oracle(orc, t):
    """RunGeoreferencing (run_ptz_ba.cc:131-155): ray terms + annotated 2d-3d points, free T_l_w"""
    p = small_scene(t, num_pts3d=15)
    got = ptz.ba_solve(p, max_num_iterations=200)
    rc, want = orc.ba_solve(p, max_num_iterations=200)
    assert rc == 0
    check_solve(orc, p, got, want, f"georef{t}")
    assert abs(got.final_reproj_error_2d3d - want.final_reproj_error_2d3d) <= 1e-6 * want.final_reproj_error_2d3d
    assert np.abs(got.tlw - want.tlw).max() <= 1e-5
    assert np.abs(got.intr[:, 1] - want.intr[:, 1]).max() <= 1e-4  # fy of the annotated views is driven by the 2d-3d terms only
    assert np.abs(got.cams_world - want.cams_world).max() <= 1e-4


def test_iteration_cap_reports_no_convergence(orc):
    p = small_scene(abi.PTZ_BA_PTZRAY)
    got = ptz.ba_solve(p, max_num_iterations=1)
    rc, want = orc.ba_solve(p, max_num_iterations=1)
    assert got.termination == want.termination == abi.PTZ_NO_CONVERGENCE
    assert not got.converged
    assert abs(got.final_cost - want.final_cost) <= 1e-8 * want.final_cost
    ok, cams, rays = ptz.PTZRayOptimizer(p, 1).Solve()
    assert ok is False and cams is None  # outputs untouched unless CONVERGENCE (ptzray_optimizer.cc:482-487)


def test_edge_cases(orc):
    # a view without observations, a track list in arbitrary order, arbitrary observation order
    p = small_scene(abi.PTZ_BA_PTZRAY)
    rng = np.random.default_rng(3)
    keep = p.obs_view != 2
    sh = rng.permutation(int(keep.sum()))
    q = ptz.BAProblem(p.factor_type, p.intr, p.ext, p.obs_uv[keep][sh], p.obs_view[keep][sh], p.obs_track[keep][sh], p.track_weight)
    got = ptz.ba_solve(q, max_num_iterations=100)
    rc, want = orc.ba_solve(q, max_num_iterations=100)
    assert got.termination == want.termination and got.num_iterations == want.num_iterations
    assert abs(got.final_cost - want.final_cost) <= 1e-6 * want.final_cost
    assert np.array_equal(got.intr[2], p.intr[2]) and np.array_equal(got.ext[2], p.ext[2])  # untouched view
    # no observations at all: nothing to do, cost 0
    e = ptz.BAProblem(0, p.intr, p.ext, np.zeros((0, 2), np.float32), [], [], [])
    r = ptz.ba_solve(e)
    assert r.initial_cost == 0.0 and r.num_iterations == 0


def test_handle_reuse_and_stage_times():
    p = small_scene(abi.PTZ_BA_PTZRAY)
    h = ptz.BAHandle(p, max_num_iterations=200)
    a = h.run(200)
    h.reset()
    b = h.run(3)
    c = h.run(200)
    assert a.termination == c.termination and a.num_iterations == c.num_iterations
    assert a.final_cost == c.final_cost  # bit-reproducible: every reduction has a fixed order
    assert b.num_iterations == 3
    t = h.stage_times()
    assert t["launches_total"] > 0 and t["ms_run"] > 0 and t["ms_kernels_total"] <= t["ms_run"] and t["lm_iterations"] >= c.num_iterations
    assert {"resjac", "track_solve", "schur_diag", "schur_offdiag", "pcg", "track_backsub", "cost"} <= set(t["kernels"])
    h.close()


@pytest.mark.parametrize("t", [abi.PTZ_BA_PTZRAY, abi.PTZ_BA_PTZRAY_DIST])
def test_full_size_properties(t):
    """BASELINE cfg 4 at 1/4 size (V=250, M~5e5): no oracle at this size; size-independent properties instead."""
    p = synth.make_config(4, scale=0.25, factor_type=t)
    r = ptz.ba_solve(p, max_num_iterations=50)
    assert r.converged
    acc = [l["cost"] for l in r.log if l["step_is_successful"] == 1]
    assert all(b < a for a, b in zip(acc, acc[1:]))  # accepted steps strictly decrease the cost
    assert r.final_cost == min(acc)
    assert 0.5 < r.final_reproj_error_2d2d < 1.2  # sigma = 0.5 px per axis -> ~0.7 px RMS at the optimum
    assert np.abs(r.intr[:, 0] - p.gt["f"]).max() < 5.0
    # idempotence: restarting from the solution stops at once with the same cost
    q = p.with_params(intr=r.intr, ext=r.ext, ray=r.ray)
    r2 = ptz.ba_solve(q, max_num_iterations=50)
    assert r2.num_iterations <= 2 and abs(r2.final_cost - r.final_cost) <= 1e-6 * r.final_cost
    # gradient at the solution is small relative to the initial one
    g0 = ptz.ba_eval(p).gradient
    g1 = ptz.ba_eval(q).gradient
    assert np.abs(g1).max() < 1e-3 * np.abs(g0).max()


def test_device_structure_builder_equals_host_builder():
    """The orderings / pair lists / block CSR built on the GPU (ba_setup.cuh) reproduce the host builder (ba_structure.hpp,
    unit-tested on the CPU) exactly: with identical structure every fixed-order reduction gives bit-identical results."""
    rng = np.random.default_rng(11)
    for p in (small_scene(abi.PTZ_BA_PTZRAY), small_scene(abi.PTZ_BA_PTZRAY_DIST, num_pts3d=9), synth.make_config(4, scale=0.05)):
        sh = rng.permutation(p.M)
        keep = p.obs_view[sh] != 1  # one view without observations, arbitrary observation order
        q = ptz.BAProblem(p.factor_type, p.intr, p.ext, p.obs_uv[sh][keep], p.obs_view[sh][keep], p.obs_track[sh][keep], p.track_weight,
                          pt_uv=p.pt_uv, pt_xyz=p.pt_xyz, pt_view=p.pt_view, tlw0=p.tlw0)
        dev = ptz.ba_solve(q, max_num_iterations=30)
        host = ptz.ba_solve(q, max_num_iterations=30, verbose=4)
        assert dev.num_iterations == host.num_iterations and dev.termination == host.termination
        assert dev.final_cost == host.final_cost and dev.initial_cost == host.initial_cost
        assert np.array_equal(dev.intr, host.intr) and np.array_equal(dev.ext, host.ext) and np.array_equal(dev.ray, host.ray)
        e_dev, e_host = ptz.ba_eval(q), ptz.ba_eval(q)  # eval maps records back through the device-built permutation
        assert np.array_equal(e_dev.residuals, e_host.residuals)


@pytest.mark.parametrize("npts", [0, 15])
def test_distdisp_solve_matches_oracle(orc, npts):
    """PTZRayDistDisp / Reproj2d3dDispFactor (ptzray_optimizer.cc:202-264,335-401): the global disp[3] block in the dense border,
    coupled to every view and every ray.  On a scene where the displacement is observable (synth.make_distdisp_scene: zoom sweep,
    true displacement in the data) the solve converges and is compared like every other factor type: same termination, iteration
    count, accept/reject sequence, per-iteration cost 1e-7, final cost 1e-6, parameters 1e-6 rad / 1e-4 px."""
    p = synth.make_distdisp_scene(num_pts3d=npts)
    got = ptz.ba_solve(p, max_num_iterations=200)
    rc, want = orc.ba_solve(p, max_num_iterations=200)
    assert rc == 0 and want.termination == abi.PTZ_CONVERGENCE and want.num_iterations >= 8
    check_solve(orc, p, got, want, f"distdisp{npts}")
    assert got.converged
    # the estimate means something: the displacement polynomial at a mid-range focal is the one in the data
    f = np.array([1.0, 1500.0, 1500.0 ** 2])
    assert abs(got.disp @ f - p.gt["disp"] @ f) < 0.02
    assert 0.5 < got.final_reproj_error_2d2d < 0.9  # sigma = 0.5 px per axis
    if npts:
        assert abs(got.final_reproj_error_2d3d - want.final_reproj_error_2d3d) <= 1e-6 * want.final_reproj_error_2d3d
        assert np.abs(got.tlw - want.tlw).max() <= 1e-5
    # the ill-conditioned variant (pure rotation, no zoom sweep: disp is nearly a gauge) still follows the oracle over the first iterations
    q = synth.make_config(1, scale=0.3, factor_type=abi.PTZ_BA_PTZRAY_DIST_DISP, num_pts3d=npts)
    g8 = ptz.ba_solve(q, max_num_iterations=8)
    rc, w8 = orc.ba_solve(q, max_num_iterations=8)
    assert rc == 0 and g8.num_iterations == w8.num_iterations == 8
    for lg, lw in zip(g8.log, w8.log):
        assert lg["step_is_successful"] == lw["step_is_successful"]
        assert abs(lg["cost"] - lw["cost"]) <= 1e-6 * lw["cost"]


@pytest.mark.parametrize("cfg", [1, 2])
def test_full_cfg1_cfg2_solve_matches_oracle(orc, cfg):
    """BASELINE cfg 1 (Synthetic-shaped, V=36, PTZRay) and cfg 2 (WorldCup14-shaped, V=60, PTZRayDist) at their FULL size against
    the oracle's exact path (dense Schur + Cholesky, exact Jacobian)."""
    p = synth.make_config(cfg)
    got = ptz.ba_solve(p, max_num_iterations=200)
    rc, want = orc.ba_solve(p, max_num_iterations=200, num_threads=orc.num_threads())
    assert rc == 0
    check_solve(orc, p, got, want, f"cfg{cfg}")
    assert got.converged
    e, w = ptz.ba_eval(p), orc.ba_eval(p)
    assert np.abs(e.residuals - w.residuals).max() <= 1e-9 * np.abs(w.residuals).max()
    assert np.abs(e.gradient - w.gradient).max() <= 1e-9 * np.abs(w.gradient).max()


def test_full_cfg1_georef_matches_oracle(orc):
    """cfg 1 with its georeferencing stage (~10 annotated points in ~3 views, Reproj2d3dFactor, free T_l_w)"""
    p = synth.make_config(1, num_pts3d=10)
    got = ptz.ba_solve(p, max_num_iterations=200)
    rc, want = orc.ba_solve(p, max_num_iterations=200, num_threads=orc.num_threads())
    assert rc == 0
    check_solve(orc, p, got, want, "cfg1-georef")
    assert np.abs(got.tlw - want.tlw).max() <= 1e-5 and np.abs(got.cams_world - want.cams_world).max() <= 1e-4


@pytest.mark.parametrize("cfg,scale,t", [(1, 1.0, abi.PTZ_BA_PTZRAY), (4, 0.12, abi.PTZ_BA_PTZRAY), (2, 1.0, abi.PTZ_BA_PTZRAY_DIST)])
def test_georef_with_every_view_annotated(orc, cfg, scale, t):
    """RunGeoreferencing annotates whatever images the annotation file lists (run_ptz_ba.cc:131-145) -- possibly all of them.  The fy
    of an annotated view is driven by its 2d-3d terms alone; it is eliminated like a ray before the CG, so the number of
    annotated views is unbounded: 36, 60 and 120 annotated views here, four points each, against the oracle."""
    V = {1: 36, 2: 60, 4: 120}[cfg]
    p = synth.make_config(cfg, scale=scale, factor_type=t, num_pts3d=4 * V, pts3d_views=V)
    assert p.V == V and len(set(p.pt_view.tolist())) == V
    got = ptz.ba_solve(p, max_num_iterations=200)
    rc, want = orc.ba_solve(p, max_num_iterations=200, num_threads=orc.num_threads())
    assert rc == 0
    check_solve(orc, p, got, want, f"georef-all-{cfg}")
    assert abs(got.final_reproj_error_2d3d - want.final_reproj_error_2d3d) <= 1e-6 * want.final_reproj_error_2d3d
    assert np.abs(got.tlw - want.tlw).max() <= 1e-5
    assert np.abs(got.intr[:, 1] - want.intr[:, 1]).max() <= 1e-4
    assert np.abs(got.cams_world - want.cams_world).max() <= 1e-4
    e, w = ptz.ba_eval(p), orc.ba_eval(p)
    assert np.abs(e.gradient - w.gradient).max() <= 1e-9 * np.abs(w.gradient).max()


@pytest.mark.parametrize("t,groups", [(abi.PTZ_BA_PTZRAY, "all"), (abi.PTZ_BA_PTZRAY, "mixed"), (abi.PTZ_BA_PTZRAY_DIST, "all"), (abi.PTZ_BA_PTZRAY_DIST, "mixed"),
                                      (abi.PTZ_BA_PTZRAY_FXFY_DIST, "all")])
def test_shared_intrinsics_match_oracle(orc, t, groups):
    """PTZRayOptimizer::SetSharedIntrinsics (ptzray_optimizer.cc:497-505, wiring :640-650, :821-848): views with one id share ONE
    intrinsics block -- the one of the first view, fixed cx, cy, dist included.  A fixed-zoom scene (one true focal length); either
    every view in one group, or three groups of different sizes next to views that keep their own block.  Same iteration table,
    costs and parameters as the oracle, and the views of a group come out bit-identical."""
    def scene():
        if t == abi.PTZ_BA_PTZRAY:
            return synth.make_config(1, scale=0.7, factor_type=t, focal_range=(1800.0, 1800.0001))
        return synth.make_config(2, scale=0.4, factor_type=t, focal_range=(2500.0, 2500.0001))

    p = scene()
    V = p.V
    if groups == "all":
        ids = np.full(V, 7, np.int32)
    else:
        ids = np.arange(V, dtype=np.int32) + 100
        ids[[2, 9, 17]] = 5            # three views far apart
        ids[[3, 4]] = 3                # two neighbours
        ids[10:16] = 42                # six in a row
    p.shared_ic_id = ids
    got = ptz.ba_solve(p, max_num_iterations=200)
    rc, want = orc.ba_solve(p, max_num_iterations=200)
    assert rc == 0 and want.termination == abi.PTZ_CONVERGENCE
    check_solve(orc, p, got, want, f"shared-{t}-{groups}")
    for g in set(ids.tolist()):
        members = np.nonzero(ids == g)[0]
        assert (got.intr[members] == got.intr[members[0]]).all()
        assert np.array_equal(got.intr[members[0], 2:4], p.intr[members[0], 2:4])  # cx, cy of the group's first view
    # the shared model really is a different problem from the independent one
    free = ptz.ba_solve(scene(), max_num_iterations=200)
    assert free.final_cost < got.final_cost and np.ptp(free.intr[:, 0]) > 1e-3


@pytest.mark.parametrize("scale", [0.25, 1.0])
def test_cfg4_iterations_match_sparse_oracle(orc, scale):
    """BASELINE cfg 4 (V=1000, 2e6 observations) at a quarter and at FULL size: the first LM iterations against the oracle's
    block-sparse path (exact Jacobian, Schur + PCG at the same 1e-13): iteration table row by row, per-iteration cost 1e-7,
    and the refined parameters after those iterations.  The second and third linear solves on the GPU are deflated ones."""
    p = synth.make_config(4, scale=scale)
    n_it = 3
    got = ptz.ba_solve(p, max_num_iterations=n_it)
    rc, want = orc.ba_solve(p, max_num_iterations=n_it, linear_solver=1, num_threads=orc.num_threads())
    assert rc == 0 and got.num_iterations == want.num_iterations == n_it
    assert abs(got.initial_cost - want.initial_cost) <= 1e-11 * want.initial_cost
    for lg, lw in zip(got.log, want.log):
        assert lg["step_is_successful"] == lw["step_is_successful"]
        assert abs(lg["cost"] - lw["cost"]) <= 1e-7 * lw["cost"]
        assert abs(lg["gradient_max_norm"] - lw["gradient_max_norm"]) <= 1e-6 * lw["gradient_max_norm"]
        assert abs(lg["trust_region_radius"] - lw["trust_region_radius"]) <= 1e-4 * lw["trust_region_radius"]
    assert abs(got.final_cost - want.final_cost) <= 1e-7 * want.final_cost
    assert np.abs(got.intr[:, 0] - want.intr[:, 0]).max() <= 1e-4
    assert np.abs(got.ext - want.ext).max() <= 1e-6 and np.abs(got.ray - want.ray).max() <= 1e-6
    if scale < 1.0:  # (the evaluation hook returns every Jacobian block to the host: quarter size only)
        e, w = ptz.ba_eval(p), orc.ba_eval(p)
        assert np.abs(e.residuals - w.residuals).max() <= 1e-9 * np.abs(w.residuals).max()
        assert np.abs(e.gradient - w.gradient).max() <= 1e-9 * np.abs(w.gradient).max()


@pytest.mark.parametrize("cfg,kw", [(1, {}), (1, dict(num_pts3d=40, pts3d_views=10)), (0, dict(num_pts3d=12))])
def test_dense_direct_solve_equals_the_cg(monkeypatch, cfg, kw):
    """small reduced systems (n <= 160: cfg 1, the BAs of an IBA run) are factored by a dense Cholesky in one CTA (k_dense_chol) instead of the CG: same LM
    trajectory, cost and parameters as the CG path at its 1e-13 tolerance -- camera-only systems, with the tlw border, with disp"""
    def scene():
        return synth.make_distdisp_scene(**kw) if cfg == 0 else synth.make_config(cfg, **kw)

    dense = ptz.ba_solve(scene(), max_num_iterations=200)
    monkeypatch.setenv("PTZ_DENSE_MAX_N", "0")
    cg = ptz.ba_solve(scene(), max_num_iterations=200)
    monkeypatch.delenv("PTZ_DENSE_MAX_N")
    assert dense.converged and dense.linear_solver_iterations < cg.linear_solver_iterations  # (one 'iteration' per direct solve)
    assert dense.num_iterations == cg.num_iterations and dense.termination == cg.termination
    assert [l["step_is_successful"] for l in dense.log] == [l["step_is_successful"] for l in cg.log]
    loose = 1e3 if cfg == 0 else 1.0
    assert abs(dense.final_cost - cg.final_cost) <= 1e-9 * cg.final_cost
    assert np.abs(dense.ext - cg.ext).max() <= 1e-8 * loose and np.abs(dense.intr[:, 0] - cg.intr[:, 0]).max() <= 1e-6 * loose
    assert np.abs(dense.ray - cg.ray).max() <= 1e-8 * loose


def test_deflated_cg_same_trajectory(monkeypatch):
    """the deflated linear solver (Ritz vectors recycled from the first solve) and the plain one give the same LM trajectory, on
    one GPU and with the rows split over virtual ranks (the multi-GPU kernel variant)"""
    p = synth.make_config(4, scale=0.2)
    monkeypatch.setenv("PTZ_CG_DEFLATE", "0")
    plain = ptz.ba_solve(p, max_num_iterations=60)
    monkeypatch.setenv("PTZ_CG_DEFLATE", "1")
    defl = ptz.ba_solve(p, max_num_iterations=60)
    monkeypatch.setenv("PTZ_CG_VRANKS", "4")
    defl4 = ptz.ba_solve(p, max_num_iterations=60)
    monkeypatch.delenv("PTZ_CG_VRANKS")
    assert plain.converged and defl.linear_solver_iterations < 0.75 * plain.linear_solver_iterations
    for r in (defl, defl4):
        assert r.num_iterations == plain.num_iterations and r.termination == plain.termination
        assert abs(r.final_cost - plain.final_cost) <= 1e-10 * plain.final_cost
        assert np.abs(r.ext - plain.ext).max() <= 1e-9 and np.abs(r.intr[:, 0] - plain.intr[:, 0]).max() <= 1e-6
        assert [l["step_is_successful"] for l in r.log] == [l["step_is_successful"] for l in plain.log]


def test_multi_gpu_sharded_solve_matches_single_gpu():
    """Real ranks (one process per GPU, NCCL + the peer-memory CG): the sharded solve reproduces the single-GPU one.  Runs
    tests/scripts/multi_gpu_check.py under torchrun on 2 devices; skipped on a one-GPU box."""
    import subprocess
    import sys

    if ptz.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(root, "tests", "scripts", "multi_gpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "MULTI_GPU_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.parametrize("t", [0, 1])
def test_golden_scipy_minima_reproduced_on_the_gpu(orc, t):
    """the scipy / cv2 golden minima (tests/golden: georeferencing with every view annotated, shared intrinsics) straight against the
    CUDA path -- no oracle in between: cost to 1e-7, gauge-invariant rotations to 1e-5 / 1e-4"""
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    k = np.load(os.path.join(gdir, "georef_kat.npz"))
    g = f"geo{t}"
    p = ptz.BAProblem(factor_type=t, intr=k[f"{g}_intr"], ext=k[f"{g}_ext"], obs_uv=k[f"{g}_obs_uv"], obs_view=k[f"{g}_obs_view"], obs_track=k[f"{g}_obs_track"],
                          track_weight=k[f"{g}_track_weight"], pt_uv=k[f"{g}_pt_uv"], pt_xyz=k[f"{g}_pt_xyz"], pt_view=k[f"{g}_pt_view"], tlw0=k[f"{g}_tlw0"])
    r = ptz.ba_solve(p, function_tolerance=1e-16, parameter_tolerance=1e-15, gradient_tolerance=1e-12, max_num_iterations=500)
    want = float(k[f"{g}_cost"])
    assert abs(r.final_cost - want) / want < 1e-7
    Ra = np.array([orc.rodrigues(e[:3]) @ orc.rodrigues(r.tlw[:3]) for e in r.ext])
    Rb = np.array([orc.rodrigues(e[:3]) @ orc.rodrigues(k[f"{g}_sol_tlw"][:3]) for e in k[f"{g}_sol_ext"]])
    assert np.abs(Ra - Rb).max() < 1e-5
    lk, sk = np.load(os.path.join(gdir, "lm_kat.npz")), np.load(os.path.join(gdir, "shared_kat.npz"))
    for name in ("some", "all") if t == 1 else ("some",):
        key = f"ba{t}_{name}"
        q = ptz.BAProblem(factor_type=t, intr=lk[f"ba{t}_intr"], ext=lk[f"ba{t}_ext"], obs_uv=lk[f"ba{t}_obs_uv"], obs_view=lk[f"ba{t}_obs_view"],
                              obs_track=lk[f"ba{t}_obs_track"], track_weight=lk[f"ba{t}_track_weight"], shared_ic_id=sk[f"{key}_ids"])
        rs = ptz.ba_solve(q, function_tolerance=1e-16, parameter_tolerance=1e-15, gradient_tolerance=1e-12, max_num_iterations=500)
        ws = float(sk[f"{key}_cost"])
        assert abs(rs.final_cost - ws) / ws < 1e-7, (key, rs.final_cost, ws)
        Rc = np.array([orc.rodrigues(e[:3]) for e in rs.ext])
        Rd = np.array([orc.rodrigues(e[:3]) for e in sk[f"{key}_sol_ext"]])
        assert np.abs(Rc @ Rc[0].T - Rd @ Rd[0].T).max() < 1e-4
