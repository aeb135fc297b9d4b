set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ba.py tests/test_gpu_tracks.py tests/test_gpu_cpp_adaptor.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log | cut -c1-600
