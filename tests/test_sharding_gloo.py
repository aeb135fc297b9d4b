"""The N>1 path on the CPU: two processes over gloo.  Each rank takes its shard of the tracks (BAProblem.shard_tracks,
the partition bench.py and the library use), evaluates its part of the camera normal equations with the oracle and
all-reduces it — the same exchange the CUDA path performs with NCCL.  The sum must equal the
unsharded evaluation; reloc shards must tile the batch."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from ptz_calib_b200 import synth

    full = synth.make_config(1, scale=0.3)
    full = full.with_params(ray=orc.ba_init_rays(full))
    shard = full.shard_tracks(rank, world)
    e = orc.ba_eval(shard)
    ncv = full.ncv
    # camera part of the gradient and the cost are additive over shards; rays are owned by one rank
    cam = torch.from_numpy(np.concatenate([e.gradient[: full.V * ncv], [e.cost, float(shard.M), float(shard.P)]]))
    dist.all_reduce(cam)
    counts = np.bincount(full.obs_track, minlength=full.P)
    cum = np.cumsum(counts)
    lo = int(np.searchsorted(cum, cum[-1] * rank / world, side="left")) if rank > 0 else 0
    ref = orc.ba_eval(full)
    ok = np.abs(cam[: full.V * ncv].numpy() - ref.gradient[: full.V * ncv]).max() <= 1e-9 * np.abs(ref.gradient).max()
    ok = ok and abs(cam[-3].item() - ref.cost) <= 1e-12 * ref.cost and int(cam[-2].item()) == full.M and int(cam[-1].item()) == full.P
    ray_g = e.gradient[full.V * ncv :].reshape(-1, 3)
    ok = ok and np.abs(ray_g - ref.gradient[full.V * ncv :].reshape(-1, 3)[lo : lo + shard.P]).max() <= 1e-9 * np.abs(ref.gradient).max()
    # georeferencing on a sharded problem ("2d-3d terms computed on rank 0 and included in the all-reduce"): every rank holds
    # all annotated points (shard_tracks keeps them), rank 0 alone evaluates them; camera blocks, cost and the T_l_w gradient then sum
    # to the unsharded evaluation
    import dataclasses

    geo = synth.make_config(1, scale=0.3, num_pts3d=12, pts3d_views=3)
    geo = geo.with_params(ray=orc.ba_init_rays(geo))
    gs = geo.shard_tracks(rank, world)
    ok = ok and gs.A == geo.A and np.array_equal(gs.pt_view, geo.pt_view)
    mine = gs if rank == 0 else dataclasses.replace(gs, pt_uv=None, pt_xyz=None, pt_view=None, tlw0=None)
    eg = orc.ba_eval(mine)
    nv = geo.V * geo.ncv
    tl = eg.gradient[-6:] if rank == 0 else np.zeros(6)
    red = torch.from_numpy(np.concatenate([eg.gradient[:nv], tl, [eg.cost]]))
    dist.all_reduce(red)
    rg = orc.ba_eval(geo)
    gmax = np.abs(rg.gradient).max()
    ok = ok and np.abs(red[:nv].numpy() - rg.gradient[:nv]).max() <= 1e-9 * gmax
    ok = ok and np.abs(red[nv : nv + 6].numpy() - rg.gradient[-6:]).max() <= 1e-9 * gmax and np.abs(rg.gradient[-6:]).max() > 0
    ok = ok and abs(red[-1].item() - rg.cost) <= 1e-12 * rg.cost
    # reloc batch: contiguous shards balanced by matches, no exchange
    b = synth.make_reloc_batch(200, n_min=8, n_max=64)
    mine = b.shard(rank, world)
    sizes = torch.tensor([mine.B, mine.N], dtype=torch.int64)
    dist.all_reduce(sizes)
    ok = ok and sizes[0].item() == b.B and sizes[1].item() == b.N and abs(mine.N - b.N / world) <= 64
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(flag.item()))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_unsharded():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert q.get(timeout=10) == 1
