"""Host-side mirror of the reference's solver interface for the hot path, over the C ABI.

`PTZRayOptimizer` / `KRTOptimizer` keep the reference's names, argument meaning and error behaviour
(src/core/ptzray_optimizer.h:112-177, src/core/krt_optimizer.h:108-145) on flat numpy inputs; the functions
`ba_solve`, `ba_eval`, `reloc_solve_batch`, `reloc_eval` are the thin ctypes calls underneath.
Everything here runs on the GPU through csrc/libptzcalib_b200.so; there is no CPU path.
"""
import ctypes as C

import numpy as np

from . import abi, lib, problem, tracks
from .abi import as_ptr, f32, f64
from .problem import BAProblem, BAResult, RelocBatch, RelocResult, default_options
from .tracks import Matches, Tracks, Views, Observations, build_tracks, flatten_tracks

__all__ = ["abi", "problem", "BAProblem", "BAResult", "RelocBatch", "RelocResult", "default_options", "ba_solve", "ba_eval", "BAHandle",
           "reloc_solve_batch", "reloc_eval", "tracks", "Matches", "Tracks", "Views", "Observations", "build_tracks", "flatten_tracks", "PTZRayOptimizer", "KRTOptimizer", "find_best_match", "init_trans_local_to_world", "device_count", "nccl_init_from_torch", "nccl_finalize"]


def device_count() -> int:
    return lib.load().ptz_device_count()


# --------------------------------------------------------------------------------------------- BA
def ba_solve(prob: BAProblem, opt=None, alloc=None, **kw) -> BAResult:
    """ptzba_solve: one PTZRayOptimizer::Solve worth of work (ptzray_optimizer.cc:454-489) with host buffers.
    alloc(name, shape): optional provider of the output arrays (caller-owned, e.g. pinned and reused across calls)."""
    opt = opt or default_options(**kw)
    c = prob.to_c()
    r, arrs, log = problem.alloc_ba_result(prob, alloc=alloc)
    rc = lib.load().ptzba_solve(C.byref(c), C.byref(opt), C.byref(r))
    lib.check(rc, "ptzba_solve")
    return problem.unpack_ba_result(prob, r, arrs, log)


def ba_eval(prob: BAProblem, disp=None) -> problem.BAEval:
    """ptzba_eval: residuals, analytic Jacobian, cost, gradient at the problem's parameters."""
    c = prob.to_c()
    e, (res, jo, jp, g) = problem.alloc_ba_eval(prob)
    d = None if disp is None else f64(disp)
    rc = lib.load().ptzba_eval(C.byref(c), as_ptr(d, C.c_double), C.byref(e))
    lib.check(rc, "ptzba_eval")
    return problem.BAEval(e.cost, res, jo, jp, g)


class BAHandle:
    """Resident problem (ptzba_create/run/destroy): upload + structure set-up once, LM iterations on demand."""

    def __init__(self, prob: BAProblem, opt=None, **kw):
        self.prob = prob
        self.opt = opt or default_options(**kw)
        self._c = prob.to_c()
        self._h = C.c_void_p()
        rc = lib.load().ptzba_create(C.byref(self._c), C.byref(self.opt), C.byref(self._h))
        lib.check(rc, "ptzba_create")

    def reset(self):
        lib.check(lib.load().ptzba_reset(self._h), "ptzba_reset")

    def run(self, max_new_iterations: int, want_outputs=True) -> BAResult:
        r, arrs, log = problem.alloc_ba_result(self.prob)
        if not want_outputs:
            for k in arrs:
                setattr(r, k, as_ptr(None, C.c_double))
        rc = lib.load().ptzba_run(self._h, C.c_int(max_new_iterations), C.byref(r))
        lib.check(rc, "ptzba_run")
        return problem.unpack_ba_result(self.prob, r, arrs, log)

    def set_stage_timing(self, per_kernel: bool):
        """per-kernel CUDA events on/off (ptzba_set_stage_timing); the span of the runs is always timed"""
        lib.check(lib.load().ptzba_set_stage_timing(self._h, C.c_int(1 if per_kernel else 0)), "ptzba_set_stage_timing")

    def stage_times(self) -> dict:
        t = abi.StageTimesC()
        lib.check(lib.load().ptzba_get_stage_times(self._h, C.byref(t)), "ptzba_get_stage_times")
        out = dict(ms_run=t.ms_run, lm_iterations=t.lm_iterations, pcg_iterations=t.pcg_iterations, jacobian_evals=t.jacobian_evals,
                   cost_evals=t.cost_evals, nnz_blocks=t.nnz_blocks, num_pairs=int(t.num_pairs), deflated_solves=t.deflated_solves,
                   deflation_vectors=t.deflation_vectors, kernels={})
        for i, name in enumerate(abi.KERNEL_NAMES[:16]):
            if t.launches[i]:
                out["kernels"][name] = dict(ms=float(t.ms_kernel[i]), launches=int(t.launches[i]), stage=abi.KERNEL_STAGE[name])
        out["launches_total"] = sum(k["launches"] for k in out["kernels"].values())
        out["ms_kernels_total"] = sum(k["ms"] for k in out["kernels"].values())
        return out

    def close(self):
        if self._h:
            lib.load().ptzba_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# --------------------------------------------------------------------------------------------- reloc
def reloc_solve_batch(batch: RelocBatch, opt=None, **kw) -> RelocResult:
    """ptzreloc_solve_batch: B independent KRTOptimizer solves (krt_optimizer.cc:385-404), one CTA each."""
    opt = opt or default_options(**kw)
    c = batch.to_c()
    r, arrs = problem.alloc_reloc_result(batch.B)
    rc = lib.load().ptzreloc_solve_batch(C.byref(c), C.byref(opt), C.byref(r))
    lib.check(rc, "ptzreloc_solve_batch")
    return RelocResult(**arrs)


def reloc_eval(ftype, uv_ref, uv_cur, ref21, local15):
    uv_ref, uv_cur = f32(uv_ref).reshape(-1, 2), f32(uv_cur).reshape(-1, 2)
    N = uv_ref.shape[0]
    nf = len(abi.KRT_FREE[ftype])
    res, jac, g, cost = np.zeros((N, 2)), np.zeros((N, 2, nf)), np.zeros(nf), C.c_double(0)
    rc = lib.load().ptzreloc_eval(C.c_int(ftype), C.c_int(N), as_ptr(uv_ref, C.c_float), as_ptr(uv_cur, C.c_float), as_ptr(f64(ref21), C.c_double),
                                  as_ptr(f64(local15), C.c_double), as_ptr(res, C.c_double), as_ptr(jac, C.c_double), C.byref(cost), as_ptr(g, C.c_double))
    lib.check(rc, "ptzreloc_eval")
    return res, jac, cost.value, g


# --------------------------------------------------------------------------------------------- multi-GPU plumbing
def nccl_init_from_torch():
    """One process per GPU (torchrun).  Rank 0 makes the ncclUniqueId, torch.distributed broadcasts its 128 bytes,
    every rank joins the library's own communicator (used only to all-reduce camera blocks)."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    buf = (C.c_ubyte * 128)()
    if rank == 0:
        lib.check(lib.load().ptz_nccl_unique_id(buf), "ptz_nccl_unique_id")
    t = torch.tensor(list(buf), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    data = bytes(t.cpu().tolist())
    buf2 = (C.c_ubyte * 128).from_buffer_copy(data)
    lib.check(lib.load().ptz_nccl_init(buf2, C.c_int(rank), C.c_int(world)), "ptz_nccl_init")
    return rank, world


def nccl_finalize():
    lib.check(lib.load().ptz_nccl_finalize(), "ptz_nccl_finalize")


def init_trans_local_to_world(cams21, pt_uv, pt_xyz, pt_view):
    """PTZRayOptimizer::SetInitTransLocalToWorld (ptzray_optimizer.cc:562-633) through ptzgeo_init_tlw: cams21 [V,21] are the problem's
    views, pt_view the view index of every annotated point.  -> (tlw [6] = rvec | t, index of the view the estimate came from or -1).
    Host code (EPnP on a few tens of points): works without a device."""
    cams21 = f64(cams21).reshape(-1, 21)
    pt_view = np.asarray(pt_view, dtype=np.int64).reshape(-1)
    V = len(cams21)
    if len(pt_view) and (pt_view.min() < 0 or pt_view.max() >= V):
        raise ValueError("pt_view out of range")
    order = np.argsort(pt_view, kind="stable")  # per view, in the order given (pixels_[i][j], pts3d_[i][j])
    uv = np.ascontiguousarray(f32(pt_uv).reshape(-1, 2)[order])
    xyz = np.ascontiguousarray(f64(pt_xyz).reshape(-1, 3)[order])
    off = np.zeros(V + 1, np.int64)
    np.cumsum(np.bincount(pt_view, minlength=V), out=off[1:])
    tlw = np.zeros(6)
    used = C.c_int32(-1)
    L = lib.load()
    lib.check(L.ptzgeo_init_tlw(C.c_int32(V), cams21.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), uv.ctypes.data_as(C.c_void_p),
                                xyz.ctypes.data_as(C.c_void_p), tlw.ctypes.data_as(C.c_void_p), C.byref(used)), "ptzgeo_init_tlw")
    return tlw, int(used.value)


# --------------------------------------------------------------------------------------------- class mirrors
class PTZRayOptimizer:
    """Mirror of ptzcalib::PTZRayOptimizer (ptzray_optimizer.h:112-177) on flattened inputs.

    Construct with a BAProblem (views = candidate cameras, observations = tracks after FindTracks) and `max_iter`;
    `Solve()` returns True only on CONVERGENCE and only then exposes refined parameters (ptzray_optimizer.cc:482-487).
    """

    def __init__(self, prob: BAProblem, max_iter: int, factor_type=None):
        self.prob = prob if factor_type is None else BAProblem(**{**prob.__dict__, "factor_type": factor_type})
        self.max_iter = int(max_iter)
        self.result = None
        self._err = (0.0, 0.0, 0.0)

    @classmethod
    def from_matches(cls, matches: Matches, views: Views, cams21, max_iter: int, factor_type=abi.PTZ_BA_PTZRAY, min_track_length=4,
                     pt_uv=None, pt_xyz=None, pt_view=None, tlw0=None, shared_ic_ids=None, reference_track_ids=True):
        """The reference's constructor arguments (features = `views`, matches_info = `matches`, cameras = krt21 rows, cam_ids =
        views.is_candidate): FindTracks (ptzray_optimizer.cc:537-552) and the residual-block loop (:801-848) run on the GPU through
        ptztracks_build / ptztracks_flatten; the 2d-3d annotations are given per ORIGINAL image index (pt_view)."""
        cams21 = f64(cams21).reshape(-1, 21)
        # reference_track_ids: ids and order of the reference's sequential UnionFind (as the C++ adaptor), else the canonical device ids
        tr = build_tracks(matches, min_track_length, reference_track_ids=reference_track_ids)
        obs = flatten_tracks(tr, views)
        cand = np.nonzero(views.is_candidate)[0]
        dense = -np.ones(len(views.is_candidate), np.int64)
        dense[cand] = np.arange(len(cand))
        intr = np.concatenate([cams21[cand, 0:4], cams21[cand, 16:21]], axis=1)
        ext = np.zeros((len(cand), 6))
        from .synth import log_so3  # rvec of R (Camera::rvec, types.h:82-87)

        ext[:, :3] = log_so3(cams21[cand, 4:13].reshape(-1, 3, 3))
        ext[:, 3:] = cams21[cand, 13:16]
        kw = {}
        if pt_uv is not None and len(pt_uv):
            if len(pt_uv) != len(pt_xyz) or len(pt_uv) != len(pt_view):  # CheckValid, ptzray_optimizer.cc:524-532
                raise ValueError("pt_uv, pt_xyz and pt_view must have one row per annotated point")
            keep = dense[np.asarray(pt_view)] >= 0
            if tlw0 is None:
                # SetInitTransLocalToWorld (:562-633), as the C++ adaptor's Solve does when no T_l_w was given: EPnP on the first
                # annotated candidate view that passes the gates, zeros when none does
                tlw0, _ = init_trans_local_to_world(cams21[cand], f32(pt_uv)[keep], f64(pt_xyz)[keep], dense[np.asarray(pt_view)][keep])
            kw = dict(pt_uv=f32(pt_uv)[keep], pt_xyz=f64(pt_xyz)[keep], pt_view=dense[np.asarray(pt_view)][keep].astype(np.int32), tlw0=tlw0)
        if shared_ic_ids is not None:  # SetSharedIntrinsics (:497-505): one id per ORIGINAL image; a wrong length is ignored as there
            if len(shared_ic_ids) == len(views.is_candidate):
                kw["shared_ic_id"] = np.asarray(shared_ic_ids, dtype=np.int64)[cand].astype(np.int32)
        prob = BAProblem(factor_type=factor_type, intr=intr, ext=ext, obs_uv=obs.obs_uv, obs_view=obs.obs_view, obs_track=obs.obs_track,
                         track_weight=obs.track_weight, **kw)
        self = cls(prob, max_iter)
        self.tracks, self.observations, self.view_of = tr, obs, cand
        return self

    def SetSharedIntrinsics(self, shared_ic_ids):  # ptzray_optimizer.cc:497-505: one id per view of the problem; a wrong length is ignored
        if len(shared_ic_ids) == self.prob.V:
            self.prob.shared_ic_id = np.asarray(shared_ic_ids, dtype=np.int32)

    def CheckValid(self) -> bool:  # ptzray_optimizer.cc:515-535
        p = self.prob
        return p.V > 0 and self.max_iter > 0

    def Solve(self):
        """-> (ok, cams_world [V,21] or None, rays_world [P,3] or None)"""
        if not self.CheckValid():
            return False, None, None
        self.result = ba_solve(self.prob, max_num_iterations=self.max_iter)
        r = self.result
        self._err = (r.final_reproj_error_all, r.final_reproj_error_2d2d, r.final_reproj_error_2d3d)
        if r.converged:
            return True, r.cams_world, r.rays_world
        return False, None, None

    def final_reproj_error_all(self):
        return self._err[0]

    def final_reproj_error_2d2d(self):
        return self._err[1]

    def final_reproj_error_2d3d(self):
        return self._err[2]


def find_best_match(fname, img_pairs_name, pairs_matches):
    """FindBestMatch (run_ptz_reloc.cc:147-166): among the pairs whose SECOND image is `fname`, the one with the most matches (the
    first wins ties).  Returns (reference image name, matches) or ("", []) when there is none."""
    best = ("", [])
    if len(img_pairs_name) != len(pairs_matches):
        return best
    for (first, second), m in zip(img_pairs_name, pairs_matches):
        if second == fname and len(m) > len(best[1]):
            best = (first, m)
    return best


class KRTOptimizer:
    """Mirror of ptzcalib::KRTOptimizer (krt_optimizer.h:108-145) for one query; the batched call is reloc_solve_batch."""

    F, FDist, Fxfy, FxfyDist = abi.PTZ_KRT_F, abi.PTZ_KRT_FDIST, abi.PTZ_KRT_FXFY, abi.PTZ_KRT_FXFYDIST

    def __init__(self, max_iter: int, max_reproj_error: float, factor_type: int):
        self.max_iter, self.max_reproj_error, self.factor_type = int(max_iter), float(max_reproj_error), int(factor_type)
        self.num_iter_ = 0
        self._init = None
        self._ref = None
        self._uv = None
        self._pts = None

    def SetInitParams(self, K, R, t, dist):  # krt_optimizer.cc:257-263
        K, R = f64(K).reshape(3, 3), f64(R).reshape(3, 3)
        self._init = np.concatenate([[K[0, 0], K[1, 1], K[0, 2], K[1, 2]], R.reshape(9), f64(t).reshape(3), f64(dist).reshape(5)])

    def Add2d2dConstraints(self, cam_ref21, kpts_ref, kpts_curr, matches):  # krt_optimizer.cc:265-348
        """cam_ref21: krt21 of the reference camera; kpts_*: [n,2] pixels; matches: [m,2] (queryIdx, trainIdx)"""
        matches = np.asarray(matches, dtype=np.int64).reshape(-1, 2)
        self._ref = f64(cam_ref21).reshape(21)
        self._uv = (f32(kpts_ref).reshape(-1, 2)[matches[:, 0]], f32(kpts_curr).reshape(-1, 2)[matches[:, 1]])
        self._local15 = np.zeros(15)  # cam_curr_local_param_ (krt_optimizer.cc:269-286)
        lib.check(lib.load().ptzreloc_local_params(as_ptr(self._ref, C.c_double), as_ptr(self._init, C.c_double), as_ptr(self._local15, C.c_double)),
                  "ptzreloc_local_params")

    def _reproj(self, ref21, uv1, uv2, pts2d, pts3d):
        e22, e23 = C.c_double(0), C.c_double(0)
        n, m = (0 if uv1 is None else len(uv1)), (0 if pts2d is None else len(pts2d))
        rc = lib.load().ptzreloc_reproj_error(C.c_int(self.factor_type), as_ptr(f64(ref21), C.c_double), as_ptr(self._local15, C.c_double), C.c_int(n),
                                              as_ptr(uv1, C.c_float), as_ptr(uv2, C.c_float), C.c_int(m), as_ptr(pts2d, C.c_float),
                                              as_ptr(pts3d, C.c_double), C.byref(e22), C.byref(e23))
        lib.check(rc, "ptzreloc_reproj_error")
        return e22.value, e23.value

    def Cal2d2dReprojError(self, cam_ref21, kpts_ref, kpts_curr, matches):  # krt_optimizer.cc:406-455
        """RMS of the 2d-2d functor residuals at the current local parameters (initial before Solve, refined after)"""
        matches = np.asarray(matches, dtype=np.int64).reshape(-1, 2)
        ref = f64(cam_ref21).reshape(21).copy()
        ref[4:13] = np.eye(3).reshape(9)  # cam_ref_local: K, dist of the reference; R = I, t = 0 (:409-413)
        ref[13:16] = 0.0
        uv1 = np.ascontiguousarray(f32(kpts_ref).reshape(-1, 2)[matches[:, 0]])
        uv2 = np.ascontiguousarray(f32(kpts_curr).reshape(-1, 2)[matches[:, 1]])
        return self._reproj(ref, uv1, uv2, None, None)[0]

    def Cal2d3dReprojError(self, pts2d, pts3d):  # krt_optimizer.cc:457-500
        pts2d, pts3d = np.ascontiguousarray(f32(pts2d).reshape(-1, 2)), np.ascontiguousarray(f64(pts3d).reshape(-1, 3))
        if len(pts2d) != len(pts3d) or len(pts2d) == 0:
            return -1.0
        return self._reproj(self._ref, None, None, pts2d, pts3d)[1]

    def Add2d3dConstraints(self, pts2d, pts3d):  # krt_optimizer.cc:350-383
        """pts2d: [n,2] pixels of the current image; pts3d: [n,3] world points"""
        pts2d, pts3d = f32(pts2d).reshape(-1, 2), f64(pts3d).reshape(-1, 3)
        if len(pts2d) != len(pts3d) or len(pts2d) == 0:
            return
        self._pts = (pts2d, pts3d)

    def Solve(self):
        """-> (ok, K, R, t, dist); outputs are None unless ok (krt_optimizer.cc:398-403)"""
        uv1, uv2 = self._uv
        pts = {}
        if self._pts is not None:
            pts = dict(pt_offset=np.array([0, len(self._pts[0])], np.int64), pt_uv=self._pts[0], pt_xyz=self._pts[1])
        batch = RelocBatch(self.factor_type, np.array([0, len(uv1)], np.int64), uv1, uv2, self._ref[None], self._init[None], self.max_iter,
                           self.max_reproj_error, **pts)
        res = reloc_solve_batch(batch)
        self.num_iter_ = int(res.num_iter[0])
        self._local15 = np.ascontiguousarray(res.local_cam15[0])  # Ceres refines cam_curr_local_param_ in place, converged or not
        if not res.success[0]:
            return False, None, None, None, None
        c = res.cam[0]
        K = np.array([[c[0], 0, c[2]], [0, c[1], c[3]], [0, 0, 1.0]])
        return True, K, c[4:13].reshape(3, 3).copy(), c[13:16].copy(), c[16:21].copy()
