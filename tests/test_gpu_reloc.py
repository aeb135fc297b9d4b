"""Batched PTZ relocalisation on the GPU (one query per CTA) against the CPU oracle and the scipy golden minima."""
import numpy as np
import pytest

import ptz_calib_b200 as ptz
from conftest import relerr
from ptz_calib_b200 import abi, problem, synth

pytestmark = pytest.mark.gpu

TYPES = [abi.PTZ_KRT_F, abi.PTZ_KRT_FDIST, abi.PTZ_KRT_FXFY, abi.PTZ_KRT_FXFYDIST]


@pytest.mark.parametrize("t", TYPES)
def test_eval_matches_oracle(orc, t):
    b = synth.make_reloc_batch(3, factor_type=t, n_min=50, n_max=80)
    for q in range(b.B):
        o0, o1 = b.match_offset[q], b.match_offset[q + 1]
        x = orc.krt_to_local(b.ref_cam[q], b.init_cam[q])
        x[4:7] += [0.01, -0.02, 0.005]
        rg, Jg, cg, gg = ptz.reloc_eval(t, b.uv_ref[o0:o1], b.uv_cur[o0:o1], b.ref_cam[q], x)
        ro, Jo, co, go = orc.reloc_eval(t, b.uv_ref[o0:o1], b.uv_cur[o0:o1], b.ref_cam[q], x)
        assert np.abs(rg - ro).max() <= 1e-9 * np.abs(ro).max()
        assert np.abs(Jg - Jo).max() <= 1e-9 * np.abs(Jo).max()
        assert abs(cg - co) <= 1e-12 * co
        assert np.abs(gg - go).max() <= 1e-9 * np.abs(go).max()


@pytest.mark.parametrize("t", TYPES)
def test_batch_matches_oracle(orc, t):
    b = synth.make_reloc_batch(300, factor_type=t, n_min=16, n_max=300)
    got = ptz.reloc_solve_batch(b)
    want = orc.reloc_solve_batch(b)
    assert np.array_equal(got.termination, want.termination)
    assert np.array_equal(got.success, want.success)
    assert np.array_equal(got.num_iter, want.num_iter)
    assert np.array_equal(got.iterations, want.iterations)
    assert relerr(got.initial_cost, want.initial_cost, floor=1e-30) <= 1e-11
    assert relerr(got.final_cost, want.final_cost, floor=1e-30) <= 1e-6
    assert relerr(got.final_rms, want.final_rms, floor=1e-30) <= 1e-6
    assert np.abs(got.local_cam15[:, 4:7] - want.local_cam15[:, 4:7]).max() <= 1e-6  # rad
    assert np.abs(got.local_cam15[:, :2] - want.local_cam15[:, :2]).max() <= 1e-4  # px
    assert np.abs(got.local_cam15[:, 10] - want.local_cam15[:, 10]).max() <= 1e-7
    assert np.abs(got.cam[:, 4:13] - want.cam[:, 4:13]).max() <= 1e-6  # world rotation
    assert np.abs(got.cam[:, :4] - want.cam[:, :4]).max() <= 1e-4
    assert got.success.mean() > 0.5


def test_golden_minima(lm_kat):
    """gauge-fixed minima from scipy.optimize.least_squares (tests/golden/lm_kat.npz), GPU run to tight tolerances"""
    k = lm_kat
    for t in (0, 1):
        b = problem.RelocBatch(t, k[f"reloc{t}_offset"], k[f"reloc{t}_uv_ref"], k[f"reloc{t}_uv_cur"], k[f"reloc{t}_ref"], k[f"reloc{t}_init"])
        r = ptz.reloc_solve_batch(b, function_tolerance=1e-15, parameter_tolerance=1e-14, gradient_tolerance=1e-12, max_num_iterations=200)
        sol = k[f"reloc{t}_sol"]
        assert np.all(r.termination == abi.PTZ_CONVERGENCE)
        assert relerr(r.final_cost, k[f"reloc{t}_cost"], floor=1e-30) < 1e-9
        assert np.abs(r.local_cam15[:, 4:7] - sol[:, 4:7]).max() < 1e-8
        assert np.abs(r.local_cam15[:, 0] - sol[:, 0]).max() < 1e-5


def test_ragged_and_edge_cases(orc):
    # very small, empty and very large (beyond the shared-memory staging cap) queries in one batch
    big = synth.make_reloc_batch(2, n_min=2500, n_max=2600, query_seed=5)
    small = synth.make_reloc_batch(6, n_min=1, n_max=5, query_seed=6)
    off = np.concatenate([[0], np.cumsum(np.concatenate([np.diff(big.match_offset), [0], np.diff(small.match_offset)]))])
    b = problem.RelocBatch(0, off, np.concatenate([big.uv_ref, small.uv_ref]), np.concatenate([big.uv_cur, small.uv_cur]),
                           np.concatenate([big.ref_cam, big.ref_cam[:1], small.ref_cam]), np.concatenate([big.init_cam, big.init_cam[:1], small.init_cam]))
    got = ptz.reloc_solve_batch(b)
    want = orc.reloc_solve_batch(b)
    assert np.array_equal(got.termination, want.termination)
    assert np.array_equal(got.success, want.success)
    assert relerr(got.final_cost[:2], want.final_cost[:2], floor=1e-30) <= 1e-6
    assert np.abs(got.local_cam15[:2] - want.local_cam15[:2]).max() <= 1e-4
    # KRTOptimizer class mirror on one query
    k = ptz.KRTOptimizer(200, 100.0, ptz.KRTOptimizer.F)
    c = big.init_cam[0]
    k.SetInitParams(np.array([[c[0], 0, c[2]], [0, c[1], c[3]], [0, 0, 1]]), c[4:13].reshape(3, 3), c[13:16], c[16:21])
    n = int(big.match_offset[1])
    k.Add2d2dConstraints(big.ref_cam[0], big.uv_ref[:n], big.uv_cur[:n], np.stack([np.arange(n), np.arange(n)], 1))
    ok, K, R, tt, dist = k.Solve()
    assert ok and abs(K[0, 0] - want.cam[0, 0]) < 1e-4 and k.num_iter_ == want.num_iter[0]


def test_full_size_properties():
    """BASELINE cfg 3 at 1/10 size: every query independent => any sub-batch reproduces its slice bit for bit"""
    b = synth.make_reloc_batch(10000)
    r = ptz.reloc_solve_batch(b)
    s = b.slice(1234, 1300)
    rs = ptz.reloc_solve_batch(s)
    assert np.array_equal(rs.cam, r.cam[1234:1300]) and np.array_equal(rs.success, r.success[1234:1300])
    ok = r.success == 1
    assert ok.mean() > 0.9
    f_err = np.abs(r.cam[ok, 0] - b.gt["f"][ok]) / b.gt["f"][ok]
    assert np.median(f_err) < 0.02  # 5 % gross outliers without a robust loss bias the focal, as in the reference
    # shards partition the batch
    parts = [b.shard(i, 4) for i in range(4)]
    assert sum(p.B for p in parts) == b.B and sum(p.N for p in parts) == b.N


@pytest.mark.parametrize("t", TYPES)
def test_batch_with_2d3d_constraints_matches_oracle(orc, t):
    """Add2d3dConstraints (krt_optimizer.cc:350-383): Factor2d3dDist / Factor2d3dFxfyDist next to the 2d-2d terms"""
    b = synth.make_reloc_batch(120, factor_type=t, n_min=16, n_max=120, pts_per_query=7)
    got = ptz.reloc_solve_batch(b)
    want = orc.reloc_solve_batch(b)
    assert np.array_equal(got.termination, want.termination) and np.array_equal(got.success, want.success)
    assert np.array_equal(got.iterations, want.iterations)
    assert relerr(got.initial_cost, want.initial_cost, floor=1e-30) <= 1e-11
    assert relerr(got.final_cost, want.final_cost, floor=1e-30) <= 1e-6
    assert relerr(got.final_rms, want.final_rms, floor=1e-30) <= 1e-6
    assert np.abs(got.local_cam15[:, 4:7] - want.local_cam15[:, 4:7]).max() <= 1e-6
    assert np.abs(got.local_cam15[:, :2] - want.local_cam15[:, :2]).max() <= 1e-4
    # the points do change the answer (they are not silently ignored)
    plain = ptz.reloc_solve_batch(synth.make_reloc_batch(120, factor_type=t, n_min=16, n_max=120))
    assert np.abs(plain.final_cost - got.final_cost).max() > 1e-3
    # the class mirror: SetInitParams / Add2d2dConstraints / Add2d3dConstraints / Solve on query 3
    q, k = 3, ptz.KRTOptimizer(200, 100.0, t)
    c = b.init_cam[q]
    k.SetInitParams(np.array([[c[0], 0, c[2]], [0, c[1], c[3]], [0, 0, 1]]), c[4:13].reshape(3, 3), c[13:16], c[16:21])
    lo, hi = int(b.match_offset[q]), int(b.match_offset[q + 1])
    k.Add2d2dConstraints(b.ref_cam[q], b.uv_ref[lo:hi], b.uv_cur[lo:hi], np.stack([np.arange(hi - lo)] * 2, 1))
    k.Add2d3dConstraints(b.pt_uv[b.pt_offset[q]:b.pt_offset[q + 1]], b.pt_xyz[b.pt_offset[q]:b.pt_offset[q + 1]])
    ok, K, R, tt, dist = k.Solve()
    assert ok == bool(got.success[q]) and k.num_iter_ == got.num_iter[q]
    if ok:
        assert K[0, 0] == got.cam[q, 0] and np.array_equal(R.reshape(9), got.cam[q, 4:13])


@pytest.mark.parametrize("t", [abi.PTZ_KRT_F, abi.PTZ_KRT_FDIST, abi.PTZ_KRT_FXFY, abi.PTZ_KRT_FXFYDIST])
def test_cal_reproj_errors_match_oracle(orc, t):
    """KRTOptimizer::Cal2d2dReprojError / Cal2d3dReprojError (krt_optimizer.cc:406-500) through the class mirror: at the initial local
    parameters (before Solve) and at the refined ones (after), against the oracle's restatement; and the identity
    Cal2d2dReprojError == sqrt(2)*sqrt(2*final_cost/num_residuals) of CheckResults when there are no 2d-3d terms."""
    b = synth.make_reloc_batch(6, factor_type=t, n_min=40, n_max=90, pts_per_query=5)
    for q in range(b.B):
        o0, o1 = int(b.match_offset[q]), int(b.match_offset[q + 1])
        p0, p1 = int(b.pt_offset[q]), int(b.pt_offset[q + 1])
        uv1, uv2 = b.uv_ref[o0:o1], b.uv_cur[o0:o1]
        ref, init = b.ref_cam[q], b.init_cam[q]
        k = ptz.KRTOptimizer(200, 100.0, t)
        c = init
        k.SetInitParams(np.array([[c[0], 0, c[2]], [0, c[1], c[3]], [0, 0, 1.0]]), c[4:13].reshape(3, 3), c[13:16], c[16:21])
        m = np.stack([np.arange(o1 - o0), np.arange(o1 - o0)], 1)
        k.Add2d2dConstraints(ref, uv1, uv2, m)
        local0 = orc.krt_to_local(ref, init)
        ref_local = ref.copy(); ref_local[4:13] = np.eye(3).reshape(9); ref_local[13:16] = 0
        want22, _ = orc.reloc_reproj_error(t, ref_local, local0, uv1, uv2)
        _, want23 = orc.reloc_reproj_error(t, ref, local0, None, None, b.pt_uv[p0:p1], b.pt_xyz[p0:p1])
        assert abs(k.Cal2d2dReprojError(ref, uv1, uv2, m) - want22) <= 1e-9 * want22
        assert abs(k.Cal2d3dReprojError(b.pt_uv[p0:p1], b.pt_xyz[p0:p1]) - want23) <= 1e-9 * want23
        assert k.Cal2d3dReprojError(np.zeros((0, 2)), np.zeros((0, 3))) == -1.0
        ok, K, R, tt, dist = k.Solve()
        single = orc.reloc_solve_batch(b.slice(q, q + 1).__class__(t, [0, o1 - o0], uv1, uv2, ref[None], init[None], b.max_iter, b.max_reproj_error))
        assert bool(ok) == bool(single.success[0])
        after22, _ = orc.reloc_reproj_error(t, ref_local, single.local_cam15[0], uv1, uv2)
        got22 = k.Cal2d2dReprojError(ref, uv1, uv2, m)
        assert abs(got22 - after22) <= 1e-6 * after22
        assert abs(got22 - single.final_rms[0]) <= 1e-6 * single.final_rms[0]
        assert got22 < want22  # the refinement reduced the error
