set -x
mkdir -p gpurun_out
(timeout 900 python tests/scripts/defl_check.py 1.0 > gpurun_out/defl_100.log 2>&1; echo rc=$? >> gpurun_out/defl_100.log)
cat gpurun_out/defl_100.log
(timeout 900 python tests/scripts/defl_check.py 0.25 4 > gpurun_out/defl_025_vr4.log 2>&1; echo rc=$? >> gpurun_out/defl_025_vr4.log)
cat gpurun_out/defl_025_vr4.log
timeout 1200 python -m pytest tests/test_gpu_ba.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
