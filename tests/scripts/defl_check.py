"""Deflated vs plain CG on the GPU: per-LM-iteration CG counts, final cost, stage-3 time.  Usage: defl_check.py [scale] [vranks]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402

import ptz_calib_b200 as ptz  # noqa: E402
from ptz_calib_b200 import synth  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
vr = sys.argv[2] if len(sys.argv) > 2 else None
ft = int(sys.argv[3]) if len(sys.argv) > 3 else 0
p = synth.make_ba_scene(max(8, int(1000 * scale)), int(400000 * scale), "band", factor_type=ft, seed=synth.SEEDS[4], track_seed=900001)
print(f"scene V={p.V} P={p.P} M={p.M} factor {ft} vranks {vr}", flush=True)
res = {}
for mode in ("0", "1"):
    os.environ["PTZ_CG_DEFLATE"] = mode
    if vr:
        os.environ["PTZ_CG_VRANKS"] = vr
    h = ptz.BAHandle(p, max_num_iterations=200)
    h.run(200)
    h.reset()
    a = h.stage_times()
    t0 = time.perf_counter()
    r = h.run(200)
    dt = time.perf_counter() - t0
    b = h.stage_times()
    pcg_ms = b["kernels"]["pcg"]["ms"] - a["kernels"]["pcg"]["ms"]
    dk = b["kernels"].get("deflate"); dk0 = a["kernels"].get("deflate", dict(ms=0.0, launches=0))
    if dk:
        print(f"   deflation set-up: {dk['ms'] - dk0['ms']:.3f} ms over {dk['launches'] - dk0['launches']} solves", flush=True)
    its = b["pcg_iterations"] - a["pcg_iterations"]
    lm = b["lm_iterations"] - a["lm_iterations"]
    print(f"deflate={mode}: term {r.termination} LM its {r.num_iterations} cost {r.final_cost:.12e} pcg its {its} ({its / max(lm, 1):.1f}/step) pcg ms {pcg_ms:.3f} "
          f"({1e3 * pcg_ms / max(its, 1):.2f} us/it) run {b['ms_run'] - a['ms_run']:.2f} ms wall {1e3 * dt:.1f} ms", flush=True)
    print("   per-step CG:", [l["linear_solver_iterations"] for l in r.log][1:], flush=True)
    res[mode] = r
    h.close()
a, b = res["0"], res["1"]
print("same iterations:", a.num_iterations == b.num_iterations, " cost rel diff: %.2e" % (abs(a.final_cost - b.final_cost) / a.final_cost),
      " max |d ext| %.2e  max |d f| %.2e" % (np.abs(a.ext - b.ext).max(), np.abs(a.intr[:, 0] - b.intr[:, 0]).max()))
