"""Per-kernel device time of the small configurations (cfg 1 / cfg 2): where a latency-bound LM iteration goes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ptz_calib_b200 as ptz
from ptz_calib_b200 import synth

for name, q in (("cfg1", synth.make_config(1, rot_noise_deg=2.0, focal_noise=0.08)), ("cfg2", synth.make_config(2, rot_noise_deg=2.0, focal_noise=0.08))):
    h = ptz.BAHandle(q, max_num_iterations=200)
    h.run(200); h.reset()
    a = h.stage_times()
    its = 0
    for _ in range(10):
        r = h.run(200); its += r.num_iterations; h.reset()
    b = h.stage_times()
    print(name, "V", q.V, "M", q.M, "LM its", its, "pcg its", b["pcg_iterations"] - a["pcg_iterations"], "ms_run/it", (b["ms_run"] - a["ms_run"]) / its)
    for k, v in b["kernels"].items():
        va = a["kernels"].get(k, dict(ms=0, launches=0))
        n = v["launches"] - va["launches"]
        if n:
            print("   %-16s launches/it %5.2f  avg us %8.2f  us/it %8.2f" % (k, n / its, 1e3 * (v["ms"] - va["ms"]) / n, 1e3 * (v["ms"] - va["ms"]) / its))
    h.close()
