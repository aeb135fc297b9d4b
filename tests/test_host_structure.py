"""ba_structure.hpp (orderings, chunks, pair lists, block-CSR pattern of the reduced camera system), on the CPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from ptz_calib_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hs(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hs") / "libhost_struct.so")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", os.path.join(HERE, "host_struct_check.cpp"), "-o", so], check=True)
    return C.CDLL(so)


def build(hs, p, chunk=128, extra=()):
    ex = np.asarray(extra, dtype=np.int64)
    hs.hs_build(p.V, p.P, p.M, p.obs_uv.ctypes.data_as(C.c_void_p), p.obs_view.ctypes.data_as(C.c_void_p), p.obs_track.ctypes.data_as(C.c_void_p), chunk,
                ex.ctypes.data_as(C.c_void_p), len(ex))
    nch, nub, nnzb, npairs = (hs.hs_size(i) for i in range(4))
    sizes = [p.M, p.M, p.M, p.V + 1, nch, nch, nch, p.V + 1, p.P + 1, p.M, nub, nub, npairs, npairs, p.V + 1, nnzb, p.V, nub, nub]
    names = ["perm", "o_view", "o_track", "view_off", "chunk_view", "chunk_begin", "chunk_cnt", "view_chunk_off", "t_off", "t_obs", "ub_row", "ub_col",
             "pair_a", "pair_b", "s_rowptr", "s_col", "diag_pos", "ub_pos", "ub_pos_t"]
    out = {}
    for i, (n, s) in enumerate(zip(names, sizes)):
        a = np.zeros(max(s, 1), np.int32)
        hs.hs_get(i, a.ctypes.data_as(C.c_void_p))
        out[n] = a[:s]
    po = np.zeros(nub + 1, np.int64)
    hs.hs_get_pair_off(po.ctypes.data_as(C.c_void_p))
    out["pair_off"] = po
    return out


@pytest.mark.parametrize("chunk", [128, 32])
def test_structure_invariants(hs, chunk):
    p = synth.make_config(1, scale=0.5)
    rng = np.random.default_rng(0)
    sh = rng.permutation(p.M)  # the caller's observation order is arbitrary
    p.obs_uv, p.obs_view, p.obs_track = p.obs_uv[sh].copy(), p.obs_view[sh].copy(), p.obs_track[sh].copy()
    s = build(hs, p, chunk)
    assert sorted(s["perm"]) == list(range(p.M))
    assert np.array_equal(s["o_view"], p.obs_view[s["perm"]]) and np.array_equal(s["o_track"], p.obs_track[s["perm"]])
    key = s["o_view"].astype(np.int64) * p.P + s["o_track"]
    assert np.all(np.diff(key) > 0)  # view-major, track ascending, no duplicates
    assert np.array_equal(s["view_off"], np.concatenate([[0], np.cumsum(np.bincount(p.obs_view, minlength=p.V))]))
    # chunks tile every view exactly and never straddle
    assert s["chunk_cnt"].sum() == p.M and s["chunk_cnt"].max() <= chunk and s["chunk_cnt"].min() >= 1
    for c in range(len(s["chunk_view"])):
        v = s["chunk_view"][c]
        assert s["view_off"][v] <= s["chunk_begin"][c] and s["chunk_begin"][c] + s["chunk_cnt"][c] <= s["view_off"][v + 1]
    # by-track lists
    for t in range(0, p.P, 7):
        idx = s["t_obs"][s["t_off"][t] : s["t_off"][t + 1]]
        assert np.all(s["o_track"][idx] == t) and np.all(np.diff(s["o_view"][idx]) > 0)
    # pairs: same track, views = (row, col) of their block, and every co-visible pair is present exactly once
    L = np.bincount(p.obs_track, minlength=p.P)
    assert len(s["pair_a"]) == int((L * (L - 1) // 2).sum())
    blk = np.repeat(np.arange(len(s["ub_row"])), np.diff(s["pair_off"]))
    assert np.array_equal(s["o_track"][s["pair_a"]], s["o_track"][s["pair_b"]])
    assert np.array_equal(s["o_view"][s["pair_a"]], s["ub_row"][blk]) and np.array_equal(s["o_view"][s["pair_b"]], s["ub_col"][blk])
    assert np.all(s["ub_row"] < s["ub_col"])
    k = s["ub_row"].astype(np.int64) * p.V + s["ub_col"]
    assert np.all(np.diff(k) > 0)
    # block CSR: symmetric pattern with ascending columns and the recorded slots
    for v in range(p.V):
        cols = s["s_col"][s["s_rowptr"][v] : s["s_rowptr"][v + 1]]
        assert np.all(np.diff(cols) > 0) and s["s_col"][s["diag_pos"][v]] == v
    assert np.array_equal(s["s_col"][s["ub_pos"]], s["ub_col"]) and np.array_equal(s["s_col"][s["ub_pos_t"]], s["ub_row"])
    rows_of = np.repeat(np.arange(p.V), np.diff(s["s_rowptr"]))
    assert np.array_equal(rows_of[s["ub_pos"]], s["ub_row"]) and np.array_equal(rows_of[s["ub_pos_t"]], s["ub_col"])
    assert len(s["s_col"]) == p.V + 2 * len(s["ub_row"])


def test_structure_union_pattern_and_empty(hs):
    """multi-GPU: extra (global) block keys are merged in with empty pair ranges; empty problems do not crash"""
    p = synth.make_config(1, scale=0.3)
    a, b = p.shard_tracks(0, 2), p.shard_tracks(1, 2)
    assert a.M + b.M == p.M and a.P + b.P == p.P
    sa, sb, sf = build(hs, a), build(hs, b), build(hs, p)
    ka = sa["ub_row"].astype(np.int64) << 32 | sa["ub_col"]
    kb = sb["ub_row"].astype(np.int64) << 32 | sb["ub_col"]
    union = np.union1d(ka, kb)
    assert np.array_equal(union, sf["ub_row"].astype(np.int64) << 32 | sf["ub_col"])
    sa2 = build(hs, a, extra=union)
    assert np.array_equal(sa2["ub_row"].astype(np.int64) << 32 | sa2["ub_col"], union)
    assert len(sa2["pair_a"]) == len(sa["pair_a"])
    cnt = dict(zip(ka.tolist(), np.diff(sa["pair_off"]).tolist()))
    for key, c in zip(union.tolist(), np.diff(sa2["pair_off"]).tolist()):
        assert c == cnt.get(key, 0)
    e = synth.make_ba_scene(4, 0)
    se = build(hs, e)
    assert len(se["pair_a"]) == 0 and len(se["s_col"]) == e.V
