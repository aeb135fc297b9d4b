"""one complete cfg-4 solve (for ncu launch lists): python tests/scripts/one_solve.py [scale] [factor_type]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ptz_calib_b200 as ptz  # noqa: E402
from ptz_calib_b200 import synth  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
ft = int(sys.argv[2]) if len(sys.argv) > 2 else 0
p = synth.make_ba_scene(max(8, int(1000 * scale)), int(400000 * scale), "band", factor_type=ft, seed=synth.SEEDS[4], track_seed=900001)
h = ptz.BAHandle(p, max_num_iterations=200)
r = h.run(200)
print("iterations", r.num_iterations, "cost", r.final_cost, "pcg", r.linear_solver_iterations)
h.close()
