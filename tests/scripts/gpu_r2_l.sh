set -x
mkdir -p gpurun_out
python tests/scripts/small_breakdown.py 2>&1 | grep -E "cfg|pcg|track_solve"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log | cut -c1-500
