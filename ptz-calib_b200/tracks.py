"""Feature tracks from pairwise matches (TracksBuilder, src/core/tracks.cc:19-113) and their flattening into the observation
arrays of a BA problem (AddConstraints2d2d, ptzray_optimizer.cc:801-848), through ptztracks_build / ptztracks_flatten of the
C ABI.  `Matches`, `Tracks`, `Views`, `Observations` are flat numpy mirrors of vector<MatchesInfo>, Tracks, vector<ImageFeatures>
+ cam_ids and of the rows a ptzba_problem takes.  The same containers are used by the test-only oracle binding; the compute
entry points here always go to the CUDA library (no CPU path)."""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import lib
from .abi import as_ptr, ip, lp, fp, dp

u8p = C.POINTER(C.c_uint8)


class MatchesC(C.Structure):
    _fields_ = [("num_pairs", C.c_int32), ("pair_src", ip), ("pair_dst", ip), ("match_offset", lp), ("query_idx", ip), ("train_idx", ip),
                ("min_track_length", C.c_int32)]


class TracksC(C.Structure):
    _fields_ = [("num_nodes", C.c_int32), ("num_components", C.c_int32), ("num_tracks", C.c_int32), ("num_elems", C.c_int64),
                ("cap_tracks", C.c_int64), ("cap_elems", C.c_int64), ("track_id", ip), ("track_offset", lp), ("elem_img", ip), ("elem_feat", ip)]


class ViewsC(C.Structure):
    _fields_ = [("num_images", C.c_int32), ("is_candidate", u8p), ("kp_offset", lp), ("kp_uv", fp)]


class ObsC(C.Structure):
    _fields_ = [("num_rows", C.c_int32), ("num_obs", C.c_int32), ("cap_rows", C.c_int64), ("cap_obs", C.c_int64), ("row_track", ip),
                ("track_weight", dp), ("obs_uv", fp), ("obs_view", ip), ("obs_track", ip)]


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


@dataclass
class Matches:
    """vector<MatchesInfo> flattened: pair k = (pair_src[k], pair_dst[k]) owns rows match_offset[k]..match_offset[k+1]-1"""
    pair_src: np.ndarray
    pair_dst: np.ndarray
    match_offset: np.ndarray
    query_idx: np.ndarray
    train_idx: np.ndarray

    def __post_init__(self):
        self.pair_src, self.pair_dst = _i32(self.pair_src), _i32(self.pair_dst)
        self.match_offset = np.ascontiguousarray(self.match_offset, dtype=np.int64)
        self.query_idx, self.train_idx = _i32(self.query_idx), _i32(self.train_idx)

    @property
    def num_matches(self):
        return int(self.match_offset[-1]) if len(self.match_offset) else 0

    def to_c(self, min_track_length):
        c = MatchesC()
        c.num_pairs = len(self.pair_src)
        c.pair_src, c.pair_dst = as_ptr(self.pair_src, C.c_int32), as_ptr(self.pair_dst, C.c_int32)
        c.match_offset = as_ptr(self.match_offset, C.c_int64)
        c.query_idx, c.train_idx = as_ptr(self.query_idx, C.c_int32), as_ptr(self.train_idx, C.c_int32)
        c.min_track_length = int(min_track_length)
        return c


@dataclass
class Tracks:
    num_nodes: int
    num_components: int
    track_id: np.ndarray      # [T]
    track_offset: np.ndarray  # [T+1]
    elem_img: np.ndarray      # [E] per track ascending image id
    elem_feat: np.ndarray     # [E]

    @property
    def num_tracks(self):
        return len(self.track_id)

    def as_sets(self):
        """{frozenset of (image, feature)} — the id-free content of the tracks"""
        o = self.track_offset
        return {frozenset(zip(self.elem_img[o[t]:o[t + 1]].tolist(), self.elem_feat[o[t]:o[t + 1]].tolist())) for t in range(self.num_tracks)}

    def to_c(self):
        c = TracksC()
        c.num_nodes, c.num_components, c.num_tracks, c.num_elems = self.num_nodes, self.num_components, self.num_tracks, len(self.elem_img)
        c.cap_tracks, c.cap_elems = self.num_tracks, len(self.elem_img)
        c.track_id, c.track_offset = as_ptr(self.track_id, C.c_int32), as_ptr(self.track_offset, C.c_int64)
        c.elem_img, c.elem_feat = as_ptr(self.elem_img, C.c_int32), as_ptr(self.elem_feat, C.c_int32)
        return c


@dataclass
class Views:
    """cam_ids_ membership and the keypoints of every image (features_[i].keypoints[j].pt)"""
    is_candidate: np.ndarray  # [num_images] uint8
    kp_offset: np.ndarray     # [num_images+1]
    kp_uv: np.ndarray         # [sum, 2] float32

    def __post_init__(self):
        self.is_candidate = np.ascontiguousarray(self.is_candidate, dtype=np.uint8)
        self.kp_offset = np.ascontiguousarray(self.kp_offset, dtype=np.int64)
        self.kp_uv = np.ascontiguousarray(self.kp_uv, dtype=np.float32).reshape(-1, 2)

    def to_c(self):
        c = ViewsC()
        c.num_images = len(self.is_candidate)
        c.is_candidate = as_ptr(self.is_candidate, C.c_uint8)
        c.kp_offset, c.kp_uv = as_ptr(self.kp_offset, C.c_int64), as_ptr(self.kp_uv, C.c_float)
        return c


@dataclass
class Observations:
    row_track: np.ndarray     # [rows] index into the tracks
    track_weight: np.ndarray  # [rows]
    obs_uv: np.ndarray        # [M, 2] float32
    obs_view: np.ndarray      # [M] dense candidate-view index
    obs_track: np.ndarray     # [M] row


def call_build(fn, matches: Matches, min_track_length, what):
    """shared by the product (fn = ptztracks_build) and the oracle binding (fn = orc_tracks_build)"""
    N = matches.num_matches
    c = matches.to_c(min_track_length)
    r = TracksC()
    tid, toff = np.zeros(max(N, 1), np.int32), np.zeros(max(N, 1) + 1, np.int64)
    eimg, efeat = np.zeros(max(2 * N, 1), np.int32), np.zeros(max(2 * N, 1), np.int32)
    r.cap_tracks, r.cap_elems = N, 2 * N
    r.track_id, r.track_offset, r.elem_img, r.elem_feat = as_ptr(tid, C.c_int32), as_ptr(toff, C.c_int64), as_ptr(eimg, C.c_int32), as_ptr(efeat, C.c_int32)
    rc = fn(C.byref(c), C.byref(r))
    if rc != 0:
        return rc, None
    T, E = r.num_tracks, r.num_elems
    return 0, Tracks(r.num_nodes, r.num_components, tid[:T].copy(), toff[:T + 1].copy(), eimg[:E].copy(), efeat[:E].copy())


def call_flatten(fn, tracks: Tracks, views: Views):
    ct, cv = tracks.to_c(), views.to_c()
    T, E = tracks.num_tracks, len(tracks.elem_img)
    o = ObsC()
    rt, w = np.zeros(max(T, 1), np.int32), np.zeros(max(T, 1), np.float64)
    uv, ov, ot = np.zeros((max(E, 1), 2), np.float32), np.zeros(max(E, 1), np.int32), np.zeros(max(E, 1), np.int32)
    o.cap_rows, o.cap_obs = T, E
    o.row_track, o.track_weight, o.obs_uv, o.obs_view, o.obs_track = (as_ptr(rt, C.c_int32), as_ptr(w, C.c_double), as_ptr(uv, C.c_float),
                                                                     as_ptr(ov, C.c_int32), as_ptr(ot, C.c_int32))
    rc = fn(C.byref(ct), C.byref(cv), C.byref(o))
    if rc != 0:
        return rc, None
    R, M = o.num_rows, o.num_obs
    return 0, Observations(rt[:R].copy(), w[:R].copy(), uv[:M].copy(), ov[:M].copy(), ot[:M].copy())


def reference_ids(matches: Matches, tracks: Tracks, min_track_length=4) -> Tracks:
    """ptztracks_reference_ids: the tracks relabelled with, and sorted by, the reference's union-by-rank root ids (host pass)"""
    t = Tracks(tracks.num_nodes, tracks.num_components, tracks.track_id.copy(), tracks.track_offset.copy(), tracks.elem_img.copy(), tracks.elem_feat.copy())
    cm, ct = matches.to_c(min_track_length), t.to_c()
    lib.check(lib.load().ptztracks_reference_ids(C.byref(cm), C.byref(ct)), "ptztracks_reference_ids")
    return t


def build_tracks(matches: Matches, min_track_length=4, reference_track_ids=False) -> Tracks:
    """TracksBuilder::Build + Filter(min_track_length) + ExportToSTL on the GPU (FindTracks uses 4, ptzray_optimizer.cc:541).
    reference_track_ids: ids and order of the reference's sequential UnionFind instead of the canonical smallest-node ids."""
    L = lib.load()
    rc, t = call_build(L.ptztracks_build, matches, min_track_length, "ptztracks_build")
    lib.check(rc, "ptztracks_build")
    return reference_ids(matches, t, min_track_length) if reference_track_ids else t


def flatten_tracks(tracks: Tracks, views: Views) -> Observations:
    L = lib.load()
    rc, o = call_flatten(L.ptztracks_flatten, tracks, views)
    lib.check(rc, "ptztracks_flatten")
    return o
