"""Run under torchrun on >= 2 GPUs of one box: the observation-sharded PTZ-BA (NCCL all-reduce of the camera blocks)
must reproduce the single-GPU solve of the same scene, and the sharded reloc batch its slices.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/scripts/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ptz_calib_b200 as ptz  # noqa: E402
from ptz_calib_b200 import abi, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    cases = [(abi.PTZ_BA_PTZRAY, 1, 0.5, {}), (abi.PTZ_BA_PTZRAY_DIST, 2, 0.3, {}), (abi.PTZ_BA_PTZRAY, 4, 0.1, {}),
             # georeferencing on a sharded problem: 2d-3d terms on rank 0, inside the all-reduce; every view annotated
             (abi.PTZ_BA_PTZRAY, 1, 0.5, dict(num_pts3d=54, pts3d_views=18)), (abi.PTZ_BA_PTZRAY_DIST, 4, 0.1, dict(num_pts3d=40, pts3d_views=8)),
             (abi.PTZ_BA_PTZRAY_DIST_DISP, 0, 1.0, dict(num_pts3d=12))]
    for t, cfg, scale, kw in cases:
        full = synth.make_distdisp_scene(**kw) if cfg == 0 else synth.make_config(cfg, scale=scale, factor_type=t, **kw)
        single = ptz.ba_solve(full, max_num_iterations=100)  # before the communicator exists: plain single-GPU solve
        ptz.nccl_init_from_torch()
        shard = full.shard_tracks(rank, world)
        got = ptz.ba_solve(shard, max_num_iterations=100)
        ptz.nccl_finalize()
        # (PTZRayDistDisp is ill-conditioned -- 1, f, f^2 columns of the global disp block: the ranks' sums come in another order and
        # the parameters follow at ~1e-6 relative, the cost at 1e-9)
        loose = 1e3 if t == abi.PTZ_BA_PTZRAY_DIST_DISP else 1.0
        diffs = dict(cost=abs(got.final_cost - single.final_cost) / single.final_cost / 1e-9, intr=np.abs(got.intr - single.intr).max() / (1e-6 * loose),
                     ext=np.abs(got.ext - single.ext).max() / (1e-8 * loose),
                     e2d2d=abs(got.final_reproj_error_2d2d - single.final_reproj_error_2d2d) / 1e-9)
        if full.A > 0:
            diffs.update(tlw=np.abs(got.tlw - single.tlw).max() / (1e-8 * loose), cams=np.abs(got.cams_world - single.cams_world).max() / (1e-7 * loose),
                         e2d3d=abs(got.final_reproj_error_2d3d - single.final_reproj_error_2d3d) / (1e-9 * loose))
        if t == abi.PTZ_BA_PTZRAY_DIST_DISP:
            diffs.update(disp=np.abs(got.disp - single.disp).max() / (1e-6 * max(1.0, np.abs(single.disp).max()) * loose))
        good = (got.termination == single.termination and got.num_iterations == single.num_iterations and got.num_residuals == single.num_residuals
                and all(v <= 1.0 for v in diffs.values()))
        if not good:
            print(f"[rank {rank}] differences in units of their tolerance: " + ", ".join(f"{k}={v:.3g}" for k, v in diffs.items()), flush=True)
        # rays stay on their owner rank: compare this rank's slice
        counts = np.bincount(full.obs_track, minlength=full.P)
        cum = np.cumsum(counts)
        lo = np.searchsorted(cum, cum[-1] * rank / world, side="left") if rank > 0 else 0
        good = good and np.abs(got.ray - single.ray[lo : lo + shard.P]).max() <= 1e-8 * loose
        print(f"[rank {rank}] cfg{cfg} type{t} A={full.A}: V={full.V} M={full.M} shard M={shard.M} iters {got.num_iterations}/{single.num_iterations} "
              f"cost {got.final_cost:.10e}/{single.final_cost:.10e} -> {'OK' if good else 'MISMATCH'}", flush=True)
        ok = ok and good
    # reloc: contiguous shards, no collective
    b = synth.make_reloc_batch(2000)
    whole = ptz.reloc_solve_batch(b)
    mine = b.shard(rank, world)
    part = ptz.reloc_solve_batch(mine)
    lo = int(np.searchsorted(b.match_offset, b.N * rank / world, side="left")) if rank > 0 else 0
    good = np.array_equal(part.cam, whole.cam[lo : lo + mine.B]) and np.array_equal(part.success, whole.success[lo : lo + mine.B])
    print(f"[rank {rank}] reloc shard {mine.B} queries -> {'OK' if good else 'MISMATCH'}", flush=True)
    ok = ok and good
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if int(flag.item()) == 1 else "FAIL", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
