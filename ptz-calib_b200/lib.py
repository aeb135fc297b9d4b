"""Loader of the CUDA shared library (csrc/libptzcalib_b200.so) behind the C ABI of include/ptzcalib_b200.h.
There is no CPU fallback: if the library is missing or no CUDA device is present, calls raise."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "csrc", "libptzcalib_b200.so")
_LIB = None

EXPORTS = [
    "ptz_solver_options_default", "ptz_last_error", "ptz_device_count", "ptz_measure_fp64_gflops",
    "ptzba_solve", "ptzba_eval", "ptzba_create", "ptzba_reset", "ptzba_run", "ptzba_get_stage_times", "ptzba_set_stage_timing", "ptzba_destroy",
    "ptz_nccl_unique_id", "ptz_nccl_init", "ptz_nccl_finalize",
    "ptzreloc_solve_batch", "ptzreloc_eval", "ptzreloc_solve_batch_dev", "ptzreloc_reproj_error", "ptzreloc_local_params",
    "ptztracks_build", "ptztracks_build_dev", "ptztracks_flatten", "ptztracks_reference_ids", "ptzgeo_init_tlw",
]


class PtzLibraryError(RuntimeError):
    pass


def _preload_nccl():
    """The library links libnccl.so.2.  PyTorch ships its own (newer) copy under site-packages/nvidia/nccl and needs ITS symbols: if
    ours pulled the system copy in first, a later `import torch` in the same process would fail to resolve them.  Loading torch's
    copy first (when there is one) makes both bind to the same library whatever the import order."""
    try:
        import importlib.util

        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
                return
    except Exception:
        pass


def load():
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise PtzLibraryError(f"{SO_PATH} is not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
                                  "(there is no CPU fallback for the solver path)")
        _preload_nccl()
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
