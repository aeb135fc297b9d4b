"""Loader of the CUDA shared library (csrc/libptzcalib_b200.so) behind the C ABI of include/ptzcalib_b200.h.
There is no CPU fallback: if the library is missing or no CUDA device is present, calls raise."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "csrc", "libptzcalib_b200.so")
_LIB = None

EXPORTS = [
    "ptz_solver_options_default", "ptz_last_error", "ptz_device_count",
    "ptzba_solve", "ptzba_eval", "ptzba_create", "ptzba_reset", "ptzba_run", "ptzba_get_stage_times", "ptzba_destroy",
    "ptz_nccl_unique_id", "ptz_nccl_init", "ptz_nccl_finalize",
    "ptzreloc_solve_batch", "ptzreloc_eval", "ptzreloc_solve_batch_dev",
    "ptztracks_build", "ptztracks_build_dev", "ptztracks_flatten",
]


class PtzLibraryError(RuntimeError):
    pass


def load():
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise PtzLibraryError(f"{SO_PATH} is not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
                                  "(there is no CPU fallback for the solver path)")
        L = C.CDLL(SO_PATH, mode=C.RTLD_GLOBAL)
        for name in EXPORTS:
            getattr(L, name).restype = C.c_int
        L.ptz_last_error.restype = C.c_char_p
        L.ptz_solver_options_default.restype = None
        _LIB = L
    return _LIB


def check(rc, what):
    if rc != 0:
        msg = load().ptz_last_error()
        raise PtzLibraryError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")
