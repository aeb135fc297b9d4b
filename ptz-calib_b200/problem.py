"""Host-side containers for the two problem kinds of the hot path, as flat numpy arrays, and their packing into
the C-ABI structs of include/ptzcalib_b200.h.  Layout only; used by the product binding and by the oracle binding."""
import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import abi
from .abi import as_ptr, f32, f64, i32, i64


def default_options(lib_default_fn=None, **kw) -> abi.SolverOptions:
    """Ceres 1.14 defaults; keyword arguments override fields."""
    o = abi.SolverOptions()
    o.max_num_iterations = 50
    o.function_tolerance = 1e-6
    o.gradient_tolerance = 1e-10
    o.parameter_tolerance = 1e-8
    o.initial_trust_region_radius = 1e4
    o.max_trust_region_radius = 1e16
    o.min_trust_region_radius = 1e-32
    o.min_relative_decrease = 1e-3
    o.min_lm_diagonal = 1e-6
    o.max_lm_diagonal = 1e32
    o.max_num_consecutive_invalid_steps = 5
    o.jacobi_scaling = 1
    o.pcg_max_iterations = 2000
    o.pcg_rel_tolerance = 1e-13
    o.jacobian_mode = 0
    o.linear_solver = 0
    o.num_threads = 1
    o.verbose = 0
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


@dataclass
class BAProblem:
    """Input of PTZRayOptimizer::Solve after FindTracks (ptzray_optimizer.cc:454-467), flattened."""

    factor_type: int
    intr: np.ndarray  # [V,9]
    ext: np.ndarray  # [V,6]
    obs_uv: np.ndarray  # [M,2] float32
    obs_view: np.ndarray  # [M] int32
    obs_track: np.ndarray  # [M] int32
    track_weight: np.ndarray  # [P]
    ray0: Optional[np.ndarray] = None  # [P,3]
    pt_uv: Optional[np.ndarray] = None  # [A,2] float32
    pt_xyz: Optional[np.ndarray] = None  # [A,3]
    pt_view: Optional[np.ndarray] = None  # [A]
    tlw0: Optional[np.ndarray] = None  # [6]
    shared_ic_id: Optional[np.ndarray] = None  # [V] int32: SetSharedIntrinsics (views with equal ids share one intrinsics block)
    gt: dict = field(default_factory=dict)  # ground truth of a synthetic scene (not part of the ABI)

    def __post_init__(self):
        self.intr = f64(self.intr).reshape(-1, 9)
        self.ext = f64(self.ext).reshape(-1, 6)
        self.obs_uv = f32(self.obs_uv).reshape(-1, 2)
        self.obs_view = i32(self.obs_view).reshape(-1)
        self.obs_track = i32(self.obs_track).reshape(-1)
        self.track_weight = f64(self.track_weight).reshape(-1)
        if self.ray0 is not None:
            self.ray0 = f64(self.ray0).reshape(-1, 3)
        if self.pt_uv is not None and len(self.pt_uv):
            self.pt_uv = f32(self.pt_uv).reshape(-1, 2)
            self.pt_xyz = f64(self.pt_xyz).reshape(-1, 3)
            self.pt_view = i32(self.pt_view).reshape(-1)
        else:
            self.pt_uv = self.pt_xyz = self.pt_view = None
        if self.tlw0 is not None:
            self.tlw0 = f64(self.tlw0).reshape(6)

    @property
    def V(self):
        return self.intr.shape[0]

    @property
    def P(self):
        return self.track_weight.shape[0]

    @property
    def M(self):
        return self.obs_view.shape[0]

    @property
    def A(self):
        return 0 if self.pt_uv is None else self.pt_uv.shape[0]

    @property
    def ncv(self):
        return abi.ba_ncv(self.factor_type)

    @property
    def num_tangent(self):
        return self.V * self.ncv + 3 * self.P + (3 if self.factor_type == abi.PTZ_BA_PTZRAY_DIST_DISP else 0) + (6 if self.A > 0 else 0)

    def to_c(self) -> abi.BAProblemC:
        c = abi.BAProblemC()
        c.factor_type = int(self.factor_type)
        c.num_views, c.num_tracks, c.num_obs, c.num_pts3d = self.V, self.P, self.M, self.A
        c.intr = as_ptr(self.intr, C.c_double)
        c.ext = as_ptr(self.ext, C.c_double)
        c.obs_uv = as_ptr(self.obs_uv, C.c_float)
        c.obs_view = as_ptr(self.obs_view, C.c_int32)
        c.obs_track = as_ptr(self.obs_track, C.c_int32)
        c.track_weight = as_ptr(self.track_weight, C.c_double)
        c.ray0 = as_ptr(self.ray0, C.c_double)
        c.pt_uv = as_ptr(self.pt_uv, C.c_float)
        c.pt_xyz = as_ptr(self.pt_xyz, C.c_double)
        c.pt_view = as_ptr(self.pt_view, C.c_int32)
        c.tlw0 = as_ptr(self.tlw0, C.c_double)
        if self.shared_ic_id is not None:
            self.shared_ic_id = i32(self.shared_ic_id).reshape(-1)
        c.shared_ic_id = as_ptr(self.shared_ic_id, C.c_int32)
        return c

    def with_params(self, intr=None, ext=None, ray=None, tlw=None):
        """Same observations, different evaluation point."""
        import copy

        q = copy.copy(self)
        if intr is not None:
            q.intr = f64(intr).reshape(-1, 9)
        if ext is not None:
            q.ext = f64(ext).reshape(-1, 6)
        if ray is not None:
            q.ray0 = f64(ray).reshape(-1, 3)
        if tlw is not None:
            q.tlw0 = f64(tlw).reshape(6)
        return q

    def shard_tracks(self, rank: int, world: int) -> "BAProblem":
        """tracks (hence observations) are partitioned across ranks, balanced by observation
        count; every rank keeps all views.  Contiguous split of the track list by cumulative observations."""
        counts = np.bincount(self.obs_track, minlength=self.P)
        cum = np.cumsum(counts)
        total = cum[-1] if len(cum) else 0
        lo = np.searchsorted(cum, total * rank / world, side="left") if rank > 0 else 0
        hi = np.searchsorted(cum, total * (rank + 1) / world, side="left") if rank < world - 1 else self.P
        sel = (self.obs_track >= lo) & (self.obs_track < hi)
        return BAProblem(
            factor_type=self.factor_type, intr=self.intr, ext=self.ext, obs_uv=self.obs_uv[sel], obs_view=self.obs_view[sel],
            obs_track=self.obs_track[sel] - lo, track_weight=self.track_weight[lo:hi], ray0=None if self.ray0 is None else self.ray0[lo:hi],
            # the annotated 2d-3d points are NOT sharded: every rank passes all of them (the library evaluates them on rank 0 and the
            # all-reduce of the camera blocks carries them to the others)
            pt_uv=self.pt_uv, pt_xyz=self.pt_xyz, pt_view=self.pt_view, tlw0=self.tlw0, shared_ic_id=self.shared_ic_id, gt=self.gt)


@dataclass
class BAResult:
    termination: int
    num_iterations: int
    num_successful_steps: int
    num_unsuccessful_steps: int
    num_residuals: int
    linear_solver_iterations: int
    initial_cost: float
    final_cost: float
    init_reproj_error_all: float
    final_reproj_error_all: float
    final_reproj_error_2d2d: float
    final_reproj_error_2d3d: float
    intr: np.ndarray
    ext: np.ndarray
    ray: np.ndarray
    disp: np.ndarray
    tlw: np.ndarray
    cams_world: np.ndarray
    rays_world: np.ndarray
    log: list
    seconds_setup: float = 0.0
    seconds_solve: float = 0.0

    @property
    def converged(self):  # PTZRayOptimizer::Solve return value (ptzray_optimizer.cc:482)
        return self.termination == abi.PTZ_CONVERGENCE


def alloc_ba_result(prob: BAProblem, log_capacity=512, alloc=None):
    """alloc(name, shape) -> float64 array: lets a caller hand in its own (e.g. pinned, reused) output buffers"""
    V, P = prob.V, prob.P
    shapes = dict(intr=(V, 9), ext=(V, 6), ray=(max(P, 1), 3), disp=(3,), tlw=(6,), cams_world=(V, 21), rays_world=(max(P, 1), 3))
    arrs = {k: (np.zeros(sh) if alloc is None else alloc(k, sh)) for k, sh in shapes.items()}
    log = (abi.IterLog * log_capacity)()
    r = abi.BAResultC()
    for k, a in arrs.items():
        setattr(r, k, as_ptr(a, C.c_double))
    r.log = C.cast(log, C.POINTER(abi.IterLog))
    r.log_capacity = log_capacity
    r.log_count = 0
    return r, arrs, log


def unpack_ba_result(prob: BAProblem, r: abi.BAResultC, arrs, log) -> BAResult:
    rows = [dict(cost=l.cost, cost_change=l.cost_change, gradient_max_norm=l.gradient_max_norm, step_norm=l.step_norm,
                 relative_decrease=l.relative_decrease, trust_region_radius=l.trust_region_radius,
                 linear_solver_iterations=l.linear_solver_iterations, step_is_successful=l.step_is_successful)
            for l in log[: r.log_count]]
    return BAResult(r.termination, r.num_iterations, r.num_successful_steps, r.num_unsuccessful_steps, r.num_residuals,
                    r.linear_solver_iterations, r.initial_cost, r.final_cost, r.init_reproj_error_all, r.final_reproj_error_all,
                    r.final_reproj_error_2d2d, r.final_reproj_error_2d3d, arrs["intr"], arrs["ext"], arrs["ray"][: prob.P], arrs["disp"],
                    arrs["tlw"], arrs["cams_world"], arrs["rays_world"][: prob.P], rows, r.seconds_setup, r.seconds_solve)


@dataclass
class BAEval:
    cost: float
    residuals: np.ndarray  # [M+A, 2]
    jac_obs: np.ndarray  # [M, 2, ncv+3(+3)]
    jac_pts: np.ndarray  # [A, 2, ncv+6(+3)]
    gradient: np.ndarray  # [num_tangent]


def alloc_ba_eval(prob: BAProblem):
    dd = 3 if prob.factor_type == abi.PTZ_BA_PTZRAY_DIST_DISP else 0
    res = np.zeros((prob.M + prob.A, 2))
    jo = np.zeros((prob.M, 2, prob.ncv + 3 + dd))
    jp = np.zeros((prob.A, 2, prob.ncv + 6 + dd))
    g = np.zeros(prob.num_tangent)
    e = abi.BAEvalOutC()
    e.residuals = as_ptr(res, C.c_double)
    e.jac_obs = as_ptr(jo, C.c_double)
    e.jac_pts = as_ptr(jp, C.c_double)
    e.gradient = as_ptr(g, C.c_double)
    return e, (res, jo, jp, g)


@dataclass
class RelocBatch:
    """B independent KRTOptimizer problems (krt_optimizer.cc:257-404), ragged by match count."""

    factor_type: int
    match_offset: np.ndarray  # [B+1] int64
    uv_ref: np.ndarray  # [N,2] float32
    uv_cur: np.ndarray  # [N,2] float32
    ref_cam: np.ndarray  # [B,21]
    init_cam: np.ndarray  # [B,21]
    max_iter: int = 200  # run_ptz_reloc.cc:90
    max_reproj_error: float = 100.0  # run_ptz_reloc.cc:91
    gt: dict = field(default_factory=dict)
    pt_offset: Optional[np.ndarray] = None  # [B+1] int64: optional 2d-3d terms (Add2d3dConstraints)
    pt_uv: Optional[np.ndarray] = None  # [Np,2] float32
    pt_xyz: Optional[np.ndarray] = None  # [Np,3] world frame

    def __post_init__(self):
        if self.pt_offset is not None:
            self.pt_offset = i64(self.pt_offset).reshape(-1)
            self.pt_uv = f32(self.pt_uv).reshape(-1, 2)
            self.pt_xyz = f64(self.pt_xyz).reshape(-1, 3)
        self.match_offset = i64(self.match_offset).reshape(-1)
        self.uv_ref = f32(self.uv_ref).reshape(-1, 2)
        self.uv_cur = f32(self.uv_cur).reshape(-1, 2)
        self.ref_cam = f64(self.ref_cam).reshape(-1, 21)
        self.init_cam = f64(self.init_cam).reshape(-1, 21)

    @property
    def B(self):
        return self.ref_cam.shape[0]

    @property
    def N(self):
        return int(self.match_offset[-1]) if len(self.match_offset) else 0

    def to_c(self) -> abi.RelocBatchC:
        c = abi.RelocBatchC()
        c.factor_type = int(self.factor_type)
        c.num_queries = self.B
        c.match_offset = as_ptr(self.match_offset, C.c_int64)
        c.uv_ref = as_ptr(self.uv_ref, C.c_float)
        c.uv_cur = as_ptr(self.uv_cur, C.c_float)
        c.ref_cam = as_ptr(self.ref_cam, C.c_double)
        c.init_cam = as_ptr(self.init_cam, C.c_double)
        c.max_iter = int(self.max_iter)
        c.max_reproj_error = float(self.max_reproj_error)
        c.pt_offset = as_ptr(self.pt_offset, C.c_int64)
        c.pt_uv = as_ptr(self.pt_uv, C.c_float)
        c.pt_xyz = as_ptr(self.pt_xyz, C.c_double)
        return c

    def slice(self, lo: int, hi: int) -> "RelocBatch":
        """Queries [lo, hi): the unit of multi-GPU sharding (no collective)."""
        o0, o1 = int(self.match_offset[lo]), int(self.match_offset[hi])
        gt = {k: (v[lo:hi] if isinstance(v, np.ndarray) and len(v) == self.B else v) for k, v in self.gt.items()}
        pts = {}
        if self.pt_offset is not None:
            p0, p1 = int(self.pt_offset[lo]), int(self.pt_offset[hi])
            pts = dict(pt_offset=self.pt_offset[lo : hi + 1] - p0, pt_uv=self.pt_uv[p0:p1], pt_xyz=self.pt_xyz[p0:p1])
        return RelocBatch(self.factor_type, self.match_offset[lo : hi + 1] - o0, self.uv_ref[o0:o1], self.uv_cur[o0:o1], self.ref_cam[lo:hi],
                          self.init_cam[lo:hi], self.max_iter, self.max_reproj_error, gt, **pts)

    def shard(self, rank: int, world: int) -> "RelocBatch":
        """Contiguous split balanced by cumulative match count."""
        total = self.N
        lo = int(np.searchsorted(self.match_offset, total * rank / world, side="left")) if rank > 0 else 0
        hi = int(np.searchsorted(self.match_offset, total * (rank + 1) / world, side="left")) if rank < world - 1 else self.B
        lo, hi = min(lo, self.B), min(hi, self.B)
        return self.slice(lo, hi)


@dataclass
class RelocResult:
    cam: np.ndarray  # [B,21]
    success: np.ndarray
    termination: np.ndarray
    num_iter: np.ndarray
    iterations: np.ndarray
    initial_cost: np.ndarray
    final_cost: np.ndarray
    final_rms: np.ndarray
    local_cam15: np.ndarray


def alloc_reloc_result(B: int):
    arrs = dict(cam=np.zeros((B, 21)), success=np.zeros(B, np.int32), termination=np.zeros(B, np.int32), num_iter=np.zeros(B, np.int32),
                iterations=np.zeros(B, np.int32), initial_cost=np.zeros(B), final_cost=np.zeros(B), final_rms=np.zeros(B), local_cam15=np.zeros((B, 15)))
    r = abi.RelocResultC()
    for k, a in arrs.items():
        setattr(r, k, as_ptr(a, C.c_double if a.dtype == np.float64 else C.c_int32))
    return r, arrs
