"""Workload of bench.py's hardware-counter side run (executed UNDER ncu by bench.py, never timed): a few LM iterations of the
bench scene and one reloc batch, so that ncu can report, per kernel launch, the DRAM bytes moved and the fp64 instructions
executed.  Usage: ncu ... python tests/scripts/ncu_counters.py <scale> <factor_type> <reloc_queries>"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ptz_calib_b200 as ptz  # noqa: E402
from ptz_calib_b200 import abi, synth  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
ft = int(sys.argv[2]) if len(sys.argv) > 2 else 0
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 0
p = synth.make_ba_scene(max(8, int(1000 * scale)), int(400000 * scale), "band", factor_type=ft, seed=synth.SEEDS[4], track_seed=900001)
h = ptz.BAHandle(p, max_num_iterations=3)
r = h.run(3)
h.close()
print("ba iterations", r.num_iterations, "pcg", r.linear_solver_iterations)
if nq > 0:
    b = synth.make_reloc_batch(nq, factor_type=abi.PTZ_KRT_F)
    rr = ptz.reloc_solve_batch(b)
    print("reloc", int(rr.success.sum()), "of", b.B, "matches", b.N, "lm iterations", int(rr.iterations.sum()))
