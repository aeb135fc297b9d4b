"""Per-LM-iteration CG counts of one sharded cfg-4-shaped solve (V = 1000 x ranks): python -m torch.distributed.run ... tests/scripts/defl_trace.py"""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ptz_calib_b200 as ptz
from ptz_calib_b200 import synth
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ptz.nccl_init_from_torch()
p = synth.make_ba_scene(1000 * world, 400000, "band", seed=synth.SEEDS[4], track_seed=900001 + rank)
r = ptz.ba_solve(p, max_num_iterations=200)
if rank == 0:
    print("iterations", r.num_iterations, "pcg total", r.linear_solver_iterations)
    print("pcg per LM iteration:", [l["linear_solver_iterations"] for l in r.log])
    print("radius:", ["%.2g" % l["trust_region_radius"] for l in r.log])
ptz.nccl_finalize()
dist.destroy_process_group()
