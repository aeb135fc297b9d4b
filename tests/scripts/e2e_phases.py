"""Where does an end-to-end ptzba_solve spend its wall-clock?  Run plain or under torchrun (sharded); prints the library's
phase timers (PTZ_TIMING) next to the Python-level time of the call."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["PTZ_TIMING"] = "1"
import ptz_calib_b200 as ptz  # noqa: E402
from ptz_calib_b200 import synth  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
        ptz.nccl_init_from_torch()
    full = synth.make_ba_scene(1000 * world, 400000 * world, "band", 0, seed=1004)
    prob = full.shard_tracks(rank, world) if world > 1 else full
    for name in ("intr", "ext", "obs_uv", "obs_view", "obs_track", "track_weight"):
        setattr(prob, name, torch.from_numpy(np.ascontiguousarray(getattr(prob, name))).pin_memory().numpy())
    opt = ptz.default_options()
    for i in range(3):
        if world > 1:
            torch.distributed.barrier()
        t0 = time.perf_counter()
        r = ptz.ba_solve(prob, opt)
        dt = time.perf_counter() - t0
        print(f"[rank {rank}] solve {i}: python wall {dt * 1e3:.1f} ms, setup {r.seconds_setup * 1e3:.1f} ms, lm {r.seconds_solve * 1e3:.1f} ms, iters {r.num_iterations}", flush=True)
    if world > 1:
        ptz.nccl_finalize()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
