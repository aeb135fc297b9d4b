"""How exact must the linear solve be?  Ceres' SPARSE_SCHUR is a direct solve; the GPU path runs block-Jacobi CG to
`pcg_rel_tolerance`.  For a range of tolerances: complete cfg-4 solves, LM / CG iteration counts, time, and the distance of the final
parameters from the 1e-13 solve (the parity bars are 1e-6 relative on the cost, 1e-6 rad, 1e-4 px)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ptz_calib_b200 as ptz  # noqa: E402
from ptz_calib_b200 import synth  # noqa: E402


def main():
    cfgs = [("cfg4", synth.make_config(4)), ("cfg2", synth.make_config(2)), ("cfg1", synth.make_config(1))]
    for name, p in cfgs:
        ref = None
        for tol in (1e-13, 1e-11, 1e-10, 1e-9, 1e-8, 1e-7, 1e-6):
            opt = ptz.default_options(max_num_iterations=200, pcg_rel_tolerance=tol)
            ptz.ba_solve(p, opt)
            t0 = time.perf_counter()
            r = ptz.ba_solve(p, opt)
            dt = time.perf_counter() - t0
            if ref is None:
                ref = r
            d_rv = np.abs(r.ext[:, :3] - ref.ext[:, :3]).max()
            d_f = np.abs(r.intr[:, 0] - ref.intr[:, 0]).max()
            d_ray = np.abs(r.ray - ref.ray).max()
            accept = "".join("a" if e["step_is_successful"] else "r" for e in r.log[1:])
            print(f"{name} tol {tol:7.0e}: term {r.termination} LM it {r.num_iterations:3d} ok/rej {r.num_successful_steps}/{r.num_unsuccessful_steps} "
                  f"CG it {r.linear_solver_iterations:6d}  cost {r.final_cost:.12e} (rel diff {abs(r.final_cost - ref.final_cost) / ref.final_cost:.1e})  "
                  f"d rvec {d_rv:.1e} d f {d_f:.1e} d ray {d_ray:.1e}  {dt * 1e3:7.1f} ms {accept}", flush=True)


if __name__ == "__main__":
    main()
