"""Per-LM-iteration CG counts of a single-GPU solve with many views (V = 4000 / 8000 over 2 M observations): the deflation behaviour of the
sharded runs reproduced on one GPU.  python tests/scripts/defl_trace1.py [V ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ptz_calib_b200 as ptz
from ptz_calib_b200 import synth
for V in [int(a) for a in sys.argv[1:]] or [4000]:
    p = synth.make_ba_scene(V, 400000, "band", seed=synth.SEEDS[4], track_seed=900001)
    t = time.time()
    r = ptz.ba_solve(p, max_num_iterations=200)
    print("V", V, "M", p.M, "iterations", r.num_iterations, "pcg total", r.linear_solver_iterations, "seconds_solve %.4f" % r.seconds_solve, "cost %.8e" % r.final_cost)
    print("  pcg per LM iteration:", [l["linear_solver_iterations"] for l in r.log])
