"""GPU parity of track building (ptztracks_build / ptztracks_flatten, csrc/tracks.cu) against the CPU restatement of
TracksBuilder (oracle/tracks_oracle.cpp).  Integer work: the bar is exact equality — of the tracks as sets of (image,
feature), of every array once the reference's union-by-rank ids are re-labelled canonically."""
import ctypes as C

import numpy as np
import pytest

import ptz_calib_b200 as ptz
from ptz_calib_b200 import abi, lib, synth
from ptz_calib_b200.tracks import Matches, Tracks, Views

from test_tracks_oracle import flip_every_other_pair, random_match_graph

pytestmark = pytest.mark.gpu


def canonical(t: Tracks, m: Matches) -> Tracks:
    """re-label the oracle's tracks the way the CUDA path labels them: id = flat index of the smallest (image, feature) node,
    tracks in ascending id"""
    src = np.repeat(m.pair_src, np.diff(m.match_offset)).astype(np.int64)
    dst = np.repeat(m.pair_dst, np.diff(m.match_offset)).astype(np.int64)
    nodes = np.unique(np.concatenate([src << 32 | m.query_idx, dst << 32 | m.train_idx]))
    first = t.track_offset[:-1]
    ids = np.searchsorted(nodes, t.elem_img[first].astype(np.int64) << 32 | t.elem_feat[first]).astype(np.int32)
    order = np.argsort(ids, kind="stable")
    lens = np.diff(t.track_offset)
    off = np.concatenate([[0], np.cumsum(lens[order])]).astype(np.int64)
    gather = np.concatenate([np.arange(t.track_offset[k], t.track_offset[k + 1]) for k in order]) if len(order) else np.zeros(0, np.int64)
    return Tracks(t.num_nodes, t.num_components, ids[order], off, t.elem_img[gather], t.elem_feat[gather])


def assert_same(got: Tracks, want: Tracks):
    assert got.num_nodes == want.num_nodes and got.num_components == want.num_components
    for name in ("track_id", "track_offset", "elem_img", "elem_feat"):
        assert np.array_equal(getattr(got, name), getattr(want, name)), name


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("min_len", [2, 4])
def test_random_graphs_match_oracle(orc, seed, min_len):
    m = random_match_graph(100 + seed, num_images=20, feats=60, num_pairs=80, per_pair=10 + 10 * (seed % 3))
    got = ptz.build_tracks(m, min_len)
    assert_same(got, canonical(orc.tracks_build(m, min_len), m))


@pytest.mark.parametrize("seed", range(4))
def test_reference_track_ids_equal_the_oracle_exactly(orc, seed):
    """reference-id mode: ids, order and every array equal the restated TracksBuilder's WITHOUT any relabelling (the ids are the
    union-by-rank roots of the sequential forest, union_find.h:66-92)"""
    m = flip_every_other_pair(random_match_graph(seed, num_images=20, feats=80, num_pairs=90, per_pair=40))
    for min_len in (2, 4):
        assert_same(ptz.build_tracks(m, min_len, reference_track_ids=True), orc.tracks_build(m, min_len))
    p = synth.make_config(1, scale=0.5)
    m, v, _ = synth.make_matches_from_scene(p, n_collisions=7, n_short=9, extra_keypoints=3)
    assert_same(ptz.build_tracks(m, 4, reference_track_ids=True), orc.tracks_build(m, 4))


@pytest.mark.parametrize("cfg,scale", [(1, 1.0), (2, 0.5)])
def test_scene_matches_match_oracle_and_expected(orc, cfg, scale):
    p = synth.make_config(cfg, scale=scale)
    m, v, expected = synth.make_matches_from_scene(p, n_collisions=40, n_short=60, extra_keypoints=6, edge_prob=0.5)
    got = ptz.build_tracks(m, 4)
    want = canonical(orc.tracks_build(m, 4), m)
    assert_same(got, want)
    assert got.as_sets() == expected
    # flatten, all views candidates and a subset
    for cand in (np.ones(p.V, np.uint8), (np.arange(p.V) % 3 != 0).astype(np.uint8)):
        vv = Views(cand, v.kp_offset, v.kp_uv)
        og, oo = ptz.flatten_tracks(got, vv), orc.tracks_flatten(want, vv)
        for name in ("row_track", "track_weight", "obs_uv", "obs_view", "obs_track"):
            assert np.array_equal(getattr(og, name), getattr(oo, name)), name


def test_edge_cases(orc):
    t = ptz.build_tracks(Matches([], [], [0], [], []), 4)
    assert t.num_tracks == 0 and t.num_nodes == 0 and t.track_offset.tolist() == [0]
    t = ptz.build_tracks(Matches([0, 1], [1, 2], [0, 0, 0], [], []), 2)
    assert t.num_tracks == 0
    m = Matches([0, 0], [1, 1], [0, 2, 4], [3, 3, 3, 4], [7, 7, 7, 9])  # duplicate rows
    assert_same(ptz.build_tracks(m, 2), canonical(orc.tracks_build(m, 2), m))
    # everything rejected: two features of image 1 matched to the same feature of image 0
    m = Matches([0], [1], [0, 2], [5, 5], [1, 2])
    t = ptz.build_tracks(m, 2)
    assert t.num_tracks == 0 and t.num_components == 1 and t.num_nodes == 3
    # one long chain across many pairs (deep union-find paths), matches listed in an order that hooks roots repeatedly
    n = 4000
    m = Matches(np.arange(n - 1), np.arange(1, n), np.arange(n), np.zeros(n - 1), np.zeros(n - 1))
    t = ptz.build_tracks(m, 4)
    assert t.num_tracks == 1 and len(t.elem_img) == n and t.track_id.tolist() == [0]
    # negative index / capacity too small
    L = lib.load()
    from ptz_calib_b200 import tracks as T

    bad = Matches([0], [1], [0, 1], [-1], [2])
    rc, _ = T.call_build(L.ptztracks_build, bad, 2, "")
    assert rc == abi.PTZ_ERR_INVALID
    m = random_match_graph(3)
    c = m.to_c(2)
    r = T.TracksC()
    tid, toff, e1, e2 = np.zeros(1, np.int32), np.zeros(2, np.int64), np.zeros(1, np.int32), np.zeros(1, np.int32)
    r.cap_tracks, r.cap_elems = 1, 1
    r.track_id, r.track_offset = tid.ctypes.data_as(abi.ip), toff.ctypes.data_as(abi.lp)
    r.elem_img, r.elem_feat = e1.ctypes.data_as(abi.ip), e2.ctypes.data_as(abi.ip)
    assert L.ptztracks_build(C.byref(c), C.byref(r)) == abi.PTZ_ERR_INVALID
    assert r.num_tracks == orc.tracks_build(m, 2).num_tracks  # the counts needed come back


def test_full_size_tracks_then_bundle_adjustment():
    """cfg-4 size (2e6 observations, ~4e6 matches): no oracle at this size in seconds, so check against the scene's own tracks,
    then run the BA on the flattened rows and compare with the BA on the scene's arrays."""
    p = synth.make_config(4, scale=0.25)
    m, v, expected = synth.make_matches_from_scene(p)
    t = ptz.build_tracks(m, 4)
    assert t.num_tracks == p.P and int(t.track_offset[-1]) == p.M
    lens = np.diff(t.track_offset)
    assert lens.min() >= 4
    # every element is a real keypoint of its image, each (image, feature) appears once
    key = t.elem_img.astype(np.int64) << 32 | t.elem_feat
    assert len(np.unique(key)) == p.M
    o = ptz.flatten_tracks(t, v)
    a = np.lexsort((o.obs_uv[:, 1], o.obs_uv[:, 0], o.obs_view))
    b = np.lexsort((p.obs_uv[:, 1], p.obs_uv[:, 0], p.obs_view))
    assert np.array_equal(o.obs_view[a], p.obs_view[b]) and np.array_equal(o.obs_uv[a], p.obs_uv[b])
    q = ptz.BAProblem(factor_type=p.factor_type, intr=p.intr, ext=p.ext, obs_uv=o.obs_uv, obs_view=o.obs_view, obs_track=o.obs_track,
                      track_weight=o.track_weight)
    r1, r2 = ptz.ba_solve(p, max_num_iterations=30), ptz.ba_solve(q, max_num_iterations=30)
    assert r1.termination == r2.termination == abi.PTZ_CONVERGENCE
    assert abs(r1.final_cost - r2.final_cost) <= 1e-9 * r1.final_cost


def test_python_mirror_from_matches(orc):
    """PTZRayOptimizer.from_matches: matches + keypoints + cameras in, refined cameras out, everything between on the GPU; a subset of
    candidate views as in the incremental driver's global BAs."""
    p = synth.make_config(1, scale=0.5)
    m, v, _ = synth.make_matches_from_scene(p)
    cams = np.zeros((p.V, 21))
    for i in range(p.V):
        cams[i, :4] = p.intr[i, :4]
        cams[i, 4:13] = orc.rodrigues(p.ext[i, :3]).ravel()
        cams[i, 16:21] = p.intr[i, 4:9]
    ba = ptz.PTZRayOptimizer.from_matches(m, v, cams, max_iter=100)
    ok, cw, rays = ba.Solve()
    want = ptz.ba_solve(p, max_num_iterations=100)
    assert ok and want.converged and ba.prob.M == p.M and ba.prob.P == p.P
    assert abs(ba.final_reproj_error_all() - want.final_reproj_error_all) <= 1e-8 * want.final_reproj_error_all
    assert np.abs(cw - want.cams_world).max() <= 1e-6
    cand = (np.arange(p.V) < p.V // 2).astype(np.uint8)
    ba2 = ptz.PTZRayOptimizer.from_matches(m, Views(cand, v.kp_offset, v.kp_uv), cams, max_iter=100)
    ok2, cw2, _ = ba2.Solve()
    assert ok2 and ba2.prob.V == p.V // 2 and ba2.prob.M == int(cand[p.obs_view].sum())
    assert np.array_equal(ba2.view_of, np.arange(p.V // 2))
    lens = np.diff(ba2.tracks.track_offset)
    assert np.array_equal(ba2.prob.track_weight, lens[ba2.observations.row_track].astype(np.float64))  # weight counts every image of the track
