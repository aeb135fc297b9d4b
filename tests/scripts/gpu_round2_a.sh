set -x
mkdir -p gpurun_out
nvidia-smi -L | head -2
(timeout 900 python tests/scripts/defl_check.py 0.25 > gpurun_out/defl_025.log 2>&1; echo rc=$? >> gpurun_out/defl_025.log)
(timeout 900 python tests/scripts/defl_check.py 1.0 > gpurun_out/defl_100.log 2>&1; echo rc=$? >> gpurun_out/defl_100.log)
(timeout 900 python tests/scripts/defl_check.py 0.25 4 > gpurun_out/defl_025_vr4.log 2>&1; echo rc=$? >> gpurun_out/defl_025_vr4.log)
(timeout 900 python tests/scripts/defl_check.py 0.5 1 1 > gpurun_out/defl_050_dist.log 2>&1; echo rc=$? >> gpurun_out/defl_050_dist.log)
tail -5 gpurun_out/defl_*.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
