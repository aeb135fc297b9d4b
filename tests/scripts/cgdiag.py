"""CG kernel timing model: time of one k_cg launch = fixed + iterations * per_iteration (first LM iteration of cfg 4)."""
import sys
sys.path.insert(0, '/root/repo')
import ptz_calib_b200 as ptz
from ptz_calib_b200 import synth

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
p = synth.make_ba_scene(int(1000 * scale), int(400000 * scale), "band", seed=synth.SEEDS[4], track_seed=900001)
pts = []
for cap in (1, 10, 50, 100, 150):
    h = ptz.BAHandle(p, max_num_iterations=5, pcg_max_iterations=cap)
    h.run(1); h.reset(); a = h.stage_times()
    h.run(1); b = h.stage_times()
    ms = b["kernels"]["pcg"]["ms"] - a["kernels"]["pcg"]["ms"]
    it = b["pcg_iterations"] - a["pcg_iterations"]
    pts.append((it, ms * 1e3))
    h.close()
print("V", p.V, "(iterations, us):", [(i, round(t, 1)) for i, t in pts])
(i0, t0), (i1, t1) = pts[1], pts[-1]
per = (t1 - t0) / (i1 - i0)
print("per-iteration us: %.2f   fixed us: %.1f" % (per, t0 - per * i0))
