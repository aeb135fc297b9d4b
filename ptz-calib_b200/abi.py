"""ctypes mirror of include/ptzcalib_b200.h (structs, enums).  Shared by the product binding (lib.py) and by
the test-only oracle binding (oracle/oracle.py): it is layout only, no code path."""
import ctypes as C

import numpy as np

PTZ_BA_PTZRAY, PTZ_BA_PTZRAY_DIST, PTZ_BA_PTZRAY_FXFY_DIST, PTZ_BA_PTZRAY_DIST_DISP = 0, 1, 2, 3
PTZ_KRT_F, PTZ_KRT_FDIST, PTZ_KRT_FXFY, PTZ_KRT_FXFYDIST = 0, 1, 2, 3
PTZ_CONVERGENCE, PTZ_NO_CONVERGENCE, PTZ_FAILURE = 0, 1, 2
PTZ_OK, PTZ_ERR_INVALID, PTZ_ERR_UNSUPPORTED, PTZ_ERR_CUDA, PTZ_ERR_NCCL, PTZ_ERR_NO_DEVICE = 0, -1, -2, -3, -4, -5

BA_TYPE_NAMES = {"PTZRay": 0, "PTZRayDist": 1, "PTZRayFxfyDist": 2, "PTZRayDistDisp": 3}
KRT_TYPE_NAMES = {"F": 0, "FDist": 1, "Fxfy": 2, "FxfyDist": 3}
KRT_FREE = {0: [0, 4, 5, 6], 1: [0, 4, 5, 6, 10], 2: [0, 1, 4, 5, 6], 3: [0, 1, 4, 5, 6, 10]}

dp = C.POINTER(C.c_double)
fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int32)
lp = C.POINTER(C.c_int64)


class SolverOptions(C.Structure):
    _fields_ = [
        ("max_num_iterations", C.c_int),
        ("function_tolerance", C.c_double),
        ("gradient_tolerance", C.c_double),
        ("parameter_tolerance", C.c_double),
        ("initial_trust_region_radius", C.c_double),
        ("max_trust_region_radius", C.c_double),
        ("min_trust_region_radius", C.c_double),
        ("min_relative_decrease", C.c_double),
        ("min_lm_diagonal", C.c_double),
        ("max_lm_diagonal", C.c_double),
        ("max_num_consecutive_invalid_steps", C.c_int),
        ("jacobi_scaling", C.c_int),
        ("pcg_max_iterations", C.c_int),
        ("pcg_rel_tolerance", C.c_double),
        ("jacobian_mode", C.c_int),
        ("linear_solver", C.c_int),
        ("num_threads", C.c_int),
        ("verbose", C.c_int),
    ]


class IterLog(C.Structure):
    _fields_ = [
        ("cost", C.c_double),
        ("cost_change", C.c_double),
        ("gradient_max_norm", C.c_double),
        ("step_norm", C.c_double),
        ("relative_decrease", C.c_double),
        ("trust_region_radius", C.c_double),
        ("linear_solver_iterations", C.c_int),
        ("step_is_successful", C.c_int),
    ]


class BAProblemC(C.Structure):
    _fields_ = [
        ("factor_type", C.c_int),
        ("num_views", C.c_int),
        ("num_tracks", C.c_int),
        ("num_obs", C.c_int),
        ("num_pts3d", C.c_int),
        ("intr", dp),
        ("ext", dp),
        ("obs_uv", fp),
        ("obs_view", ip),
        ("obs_track", ip),
        ("track_weight", dp),
        ("ray0", dp),
        ("pt_uv", fp),
        ("pt_xyz", dp),
        ("pt_view", ip),
        ("tlw0", dp),
        ("shared_ic_id", ip),
    ]


class BAResultC(C.Structure):
    _fields_ = [
        ("termination", C.c_int),
        ("num_iterations", C.c_int),
        ("num_successful_steps", C.c_int),
        ("num_unsuccessful_steps", C.c_int),
        ("num_residuals", C.c_int),
        ("linear_solver_iterations", C.c_int),
        ("initial_cost", C.c_double),
        ("final_cost", C.c_double),
        ("init_reproj_error_all", C.c_double),
        ("final_reproj_error_all", C.c_double),
        ("final_reproj_error_2d2d", C.c_double),
        ("final_reproj_error_2d3d", C.c_double),
        ("intr", dp),
        ("ext", dp),
        ("ray", dp),
        ("disp", dp),
        ("tlw", dp),
        ("cams_world", dp),
        ("rays_world", dp),
        ("log", C.POINTER(IterLog)),
        ("log_capacity", C.c_int),
        ("log_count", C.c_int),
        ("seconds_setup", C.c_double),
        ("seconds_solve", C.c_double),
    ]


class BAEvalOutC(C.Structure):
    _fields_ = [
        ("cost", C.c_double),
        ("residuals", dp),
        ("jac_obs", dp),
        ("jac_pts", dp),
        ("gradient", dp),
        ("num_tangent", C.c_int),
    ]


KERNEL_NAMES = ["view_prep", "resjac", "view_finalize", "track_accum", "pts", "track_solve", "schur_diag", "schur_offdiag", "precond", "pcg",
                "track_backsub", "cam_update", "cost", "scalars", "allreduce", "deflate"]
KERNEL_STAGE = {"view_prep": 1, "resjac": 1, "view_finalize": 1, "track_accum": 1, "pts": 1, "track_solve": 2, "schur_diag": 2, "schur_offdiag": 2,
                "precond": 2, "pcg": 3, "track_backsub": 4, "cam_update": 4, "cost": 4, "scalars": 4, "allreduce": 2, "deflate": 3}


class StageTimesC(C.Structure):
    _fields_ = [
        ("ms_kernel", C.c_float * 16),
        ("launches", C.c_int * 16),
        ("ms_run", C.c_float),
        ("lm_iterations", C.c_int),
        ("pcg_iterations", C.c_int),
        ("jacobian_evals", C.c_int),
        ("cost_evals", C.c_int),
        ("nnz_blocks", C.c_int),
        ("deflated_solves", C.c_int),
        ("deflation_vectors", C.c_int),
        ("reserved_", C.c_int),
        ("num_pairs", C.c_int64),
    ]


class RelocBatchC(C.Structure):
    _fields_ = [
        ("factor_type", C.c_int),
        ("num_queries", C.c_int),
        ("match_offset", lp),
        ("uv_ref", fp),
        ("uv_cur", fp),
        ("ref_cam", dp),
        ("init_cam", dp),
        ("max_iter", C.c_int),
        ("max_reproj_error", C.c_double),
        ("pt_offset", lp),
        ("pt_uv", fp),
        ("pt_xyz", dp),
    ]


class RelocResultC(C.Structure):
    _fields_ = [
        ("cam", dp),
        ("success", ip),
        ("termination", ip),
        ("num_iter", ip),
        ("iterations", ip),
        ("initial_cost", dp),
        ("final_cost", dp),
        ("final_rms", dp),
        ("local_cam15", dp),
    ]


def as_ptr(a, ctype):
    if a is None:
        return C.cast(None, C.POINTER(ctype))
    return a.ctypes.data_as(C.POINTER(ctype))


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def ba_ncv(factor_type):
    """tangent columns per view: [fx, fy, (k1), w1, w2, w3]"""
    return 5 if factor_type == PTZ_BA_PTZRAY else 6
