"""Wall-clock of the whole incremental flow (PtzIncrementalOptimizer over the C ABI) on synthetic rings of growing size:
    python tests/scripts/iba_timing.py [scale ...]"""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ptz_calib_b200 import lib, synth  # noqa: E402


def main():
    d = tempfile.mkdtemp()
    exe = os.path.join(d, "iba_check")
    so_dir = os.path.dirname(lib.SO_PATH)
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "cpp", "iba_check.cpp"), "-o", exe, "-L" + so_dir, "-lptzcalib_b200",
                    "-Wl,-rpath," + so_dir], check=True)
    for scale in [float(a) for a in sys.argv[1:]] or [1.0, 2.0]:
        p = synth.make_config(1, scale=scale)
        cams = np.zeros((p.V, 21))
        cams[:, 0] = cams[:, 1] = p.gt["f"]
        cams[:, 2:4] = p.gt["c"]
        cams[:, 4:13] = p.gt["R"].reshape(p.V, 9)
        fin, fout = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        with open(fin, "wb") as f:
            f.write(struct.pack("4i", p.V, p.M, 100, 1))
            for a in (cams, p.obs_view, p.obs_track, p.obs_uv):
                f.write(np.ascontiguousarray(a).tobytes())
        for rep in range(2):  # the second run has the CUDA context / caches of nothing: separate processes, so both are cold starts
            r = subprocess.run([exe, fin, fout], capture_output=True, text=True, timeout=600)
            print(f"V={p.V} M={p.M}:", r.stdout.strip() or r.stderr.strip(), flush=True)


if __name__ == "__main__":
    main()
