"""CPU tests of the track-building oracle (oracle/tracks_oracle.cpp, restating src/core/tracks.cc:19-118 and union_find.h) and of
the host-side containers.  The reference has no fixtures for this path, so the restatement is pinned against an independent
formulation: scipy connected components + a numpy Filter, and hand-worked cases."""
import ctypes as C

import numpy as np
import pytest
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components

import ptz_calib_b200 as ptz
from ptz_calib_b200 import abi, lib, synth
from ptz_calib_b200.tracks import Matches, Tracks, Views


def random_match_graph(seed, num_images=12, feats=40, num_pairs=30, per_pair=25):
    rng = np.random.default_rng(seed)
    pairs = set()
    while len(pairs) < num_pairs:
        a, b = rng.integers(0, num_images, 2)
        if a != b:
            pairs.add((int(min(a, b)), int(max(a, b))))
    pairs = sorted(pairs)
    off, q, t = [0], [], []
    for _ in pairs:
        n = int(rng.integers(0, per_pair))
        q += rng.integers(0, feats, n).tolist()
        t += rng.integers(0, feats, n).tolist()
        off.append(len(q))
    return Matches([p[0] for p in pairs], [p[1] for p in pairs], off, q, t)


def independent_tracks(m: Matches, min_len):
    """connected components of the match graph; keep those with >= min_len nodes, > 1 node and no image twice"""
    src = np.repeat(m.pair_src, np.diff(m.match_offset))
    dst = np.repeat(m.pair_dst, np.diff(m.match_offset))
    a = src.astype(np.int64) << 32 | m.query_idx
    b = dst.astype(np.int64) << 32 | m.train_idx
    nodes, inv = np.unique(np.concatenate([a, b]), return_inverse=True)
    n = len(nodes)
    if n == 0:
        return set(), 0, 0
    ia, ib = inv[:len(a)], inv[len(a):]
    ncomp, lab = connected_components(coo_matrix((np.ones(len(ia)), (ia, ib)), shape=(n, n)), directed=False)
    out = set()
    for c in range(ncomp):
        k = nodes[lab == c]
        imgs = k >> 32
        if len(k) >= min_len and len(k) > 1 and len(np.unique(imgs)) == len(imgs):
            out.add(frozenset(zip(imgs.tolist(), (k & 0xffffffff).tolist())))
    return out, n, ncomp


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("min_len", [2, 3, 4])
def test_oracle_matches_connected_components(orc, seed, min_len):
    m = random_match_graph(seed, per_pair=8 + 6 * (seed % 4))
    want, n, ncomp = independent_tracks(m, min_len)
    got = orc.tracks_build(m, min_len)
    assert got.num_nodes == n and got.num_components == ncomp
    assert got.as_sets() == want
    # std::map order: ascending track id, inside a track ascending image id
    assert np.all(np.diff(got.track_id) > 0)
    for k in range(got.num_tracks):
        assert np.all(np.diff(got.elem_img[got.track_offset[k]:got.track_offset[k + 1]]) > 0)


def test_oracle_hand_worked_cases(orc):
    # images 0..3; a 4-image chain (kept by Filter(4)), a 3-image chain (too short), and two chains glued through image 1
    m = Matches(pair_src=[0, 1, 2, 0, 1, 2, 0, 1], pair_dst=[1, 2, 3, 1, 2, 3, 2, 3], match_offset=[0, 1, 2, 3, 4, 5, 5, 6, 8],
                query_idx=[5, 6, 7, 10, 11, 20, 30, 31], train_idx=[6, 7, 8, 11, 12, 21, 32, 21])
    # chain A: (0,5)-(1,6)-(2,7)-(3,8); chain B: (0,10)-(1,11)-(2,12) (3 images); three 2-sets (0,20)-(2,21), (1,30)-(3,32), (1,31)-(3,21)
    t = orc.tracks_build(m, 4)
    assert t.as_sets() == {frozenset({(0, 5), (1, 6), (2, 7), (3, 8)})}
    t = orc.tracks_build(m, 3)
    assert t.as_sets() == {frozenset({(0, 5), (1, 6), (2, 7), (3, 8)}), frozenset({(0, 10), (1, 11), (2, 12)})}
    t = orc.tracks_build(m, 2)
    assert t.num_tracks == 5 and frozenset({(1, 31), (3, 21)}) in t.as_sets()
    # an image listed twice: (1,30) and (1,31) both match (3,32) -> {(0,20),(2,21),(3,32),(1,30),(1,31)} has image 1 twice
    m2 = Matches(pair_src=[0, 1, 1, 2], pair_dst=[2, 3, 3, 3], match_offset=[0, 1, 2, 3, 4], query_idx=[20, 30, 31, 21], train_idx=[21, 32, 32, 32])
    t = orc.tracks_build(m2, 2)
    assert t.num_tracks == 0 and t.num_components == 1 and t.num_nodes == 5
    # the reference's ids are union-by-rank roots: (0,5)-(1,6) first makes node 0 the root of chain A
    t = orc.tracks_build(m, 4)
    assert t.track_id.tolist() == [0]


def test_empty_and_duplicate_matches(orc):
    t = orc.tracks_build(Matches([], [], [0], [], []), 4)
    assert t.num_tracks == 0 and t.num_nodes == 0
    t = orc.tracks_build(Matches([0, 1], [1, 2], [0, 0, 0], [], []), 2)
    assert t.num_tracks == 0
    m = Matches([0, 0], [1, 1], [0, 2, 4], [3, 3, 3, 4], [7, 7, 7, 9])  # the same pair listed twice, duplicate rows
    t = orc.tracks_build(m, 2)
    assert t.as_sets() == {frozenset({(0, 3), (1, 7)}), frozenset({(0, 4), (1, 9)})}


def test_scene_matches_give_back_the_scene_tracks(orc):
    p = synth.make_config(1, scale=0.5)
    m, v, expected = synth.make_matches_from_scene(p, n_collisions=15, n_short=25, extra_keypoints=4, edge_prob=0.5)
    t = orc.tracks_build(m, 4)
    assert t.as_sets() == expected
    assert len(expected) < p.P  # the collisions removed some


def test_flatten_reproduces_the_scene_observations(orc):
    p = synth.make_config(1, scale=0.5)
    m, v, expected = synth.make_matches_from_scene(p)
    t = orc.tracks_build(m, 4)
    assert t.num_tracks == p.P
    o = orc.tracks_flatten(t, v)
    assert len(o.obs_view) == p.M and len(o.row_track) == p.P
    # same multiset of (view, u, v) rows, weights = track lengths, rows numbered in track order
    a = np.lexsort((o.obs_uv[:, 1], o.obs_uv[:, 0], o.obs_view))
    b = np.lexsort((p.obs_uv[:, 1], p.obs_uv[:, 0], p.obs_view))
    assert np.array_equal(o.obs_view[a], p.obs_view[b]) and np.array_equal(o.obs_uv[a], p.obs_uv[b])
    assert np.array_equal(np.bincount(o.obs_track), o.track_weight.astype(np.int64))
    assert np.all(np.diff(o.obs_track) >= 0) and np.array_equal(o.row_track, np.arange(p.P))
    # candidate subset: rows only for tracks that touch a candidate, weight still counts every image (ptzray_optimizer.cc:805)
    cand = np.zeros(p.V, np.uint8)
    cand[: p.V // 2] = 1
    v2 = Views(cand, v.kp_offset, v.kp_uv)
    o2 = orc.tracks_flatten(t, v2)
    assert len(o2.obs_view) == int(cand[p.obs_view].sum())
    assert o2.obs_view.max() == p.V // 2 - 1
    lens = np.diff(t.track_offset)
    assert np.array_equal(o2.track_weight, lens[o2.row_track].astype(np.float64))
    assert len(o2.row_track) < p.P


def test_tracks_entry_points_need_a_gpu_and_validate_arguments():
    L = lib.load()
    m = random_match_graph(0)
    from ptz_calib_b200 import tracks as T

    c = m.to_c(4)
    r = T.TracksC()
    assert L.ptztracks_build(C.byref(c), None) == abi.PTZ_ERR_INVALID
    off = np.zeros(1, np.int64)
    r.cap_tracks, r.cap_elems = 0, 0
    r.track_offset = off.ctypes.data_as(C.POINTER(C.c_int64))
    bad = m.match_offset.copy()
    bad[1], bad[2] = bad[2] + 1, bad[1]
    c.match_offset = bad.ctypes.data_as(C.POINTER(C.c_int64))
    assert L.ptztracks_build(C.byref(c), C.byref(r)) == abi.PTZ_ERR_INVALID
    if ptz.device_count() == 0:
        with pytest.raises(lib.PtzLibraryError, match="-5"):
            ptz.build_tracks(m, 4)
        p = synth.make_config(1, scale=0.2)
        mm, v, _ = synth.make_matches_from_scene(p)
        from oracle import oracle

        t = oracle.tracks_build(mm, 4)
        with pytest.raises(lib.PtzLibraryError, match="-5"):
            ptz.flatten_tracks(t, v)


def flip_every_other_pair(m: Matches) -> Matches:
    """every other pair stored the other way round (dst -> src), so that union-by-rank roots are not simply the smallest nodes"""
    flip = np.arange(len(m.pair_src)) % 2 == 1
    rows = np.repeat(flip, np.diff(m.match_offset))
    return Matches(np.where(flip, m.pair_dst, m.pair_src), np.where(flip, m.pair_src, m.pair_dst), m.match_offset,
                   np.where(rows, m.train_idx, m.query_idx), np.where(rows, m.query_idx, m.train_idx))


@pytest.mark.parametrize("seed", range(4))
def test_reference_id_pass_restores_the_oracle_ids(orc, seed):
    """CPU: ptztracks_reference_ids is host code.  The oracle's tracks, re-labelled canonically (smallest node) and re-sorted the way
    the device build emits them, come back with the reference's ids and order."""
    from ptz_calib_b200 import tracks as T

    m = flip_every_other_pair(random_match_graph(seed, num_images=16, feats=60, num_pairs=70, per_pair=30))
    want = orc.tracks_build(m, 2)
    src = np.repeat(m.pair_src, np.diff(m.match_offset)).astype(np.int64)
    dst = np.repeat(m.pair_dst, np.diff(m.match_offset)).astype(np.int64)
    nodes = np.unique(np.concatenate([src << 32 | m.query_idx, dst << 32 | m.train_idx]))
    first = want.track_offset[:-1]
    ids = np.searchsorted(nodes, want.elem_img[first].astype(np.int64) << 32 | want.elem_feat[first]).astype(np.int32)
    order = np.argsort(ids, kind="stable")
    lens = np.diff(want.track_offset)
    off = np.concatenate([[0], np.cumsum(lens[order])]).astype(np.int64)
    gather = np.concatenate([np.arange(want.track_offset[k], want.track_offset[k + 1]) for k in order])
    canon = Tracks(want.num_nodes, want.num_components, ids[order], off, want.elem_img[gather], want.elem_feat[gather])
    assert not np.array_equal(canon.track_id, want.track_id)  # (the two labellings do differ)
    got = T.reference_ids(m, canon, 2)
    for name in ("track_id", "track_offset", "elem_img", "elem_feat"):
        assert np.array_equal(getattr(got, name), getattr(want, name)), name
