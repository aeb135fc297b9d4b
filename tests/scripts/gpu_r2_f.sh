set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_ba.py -m gpu -x -q -k "multi_gpu" > gpurun_out/pytest_multi.log 2>&1; echo rc=$? >> gpurun_out/pytest_multi.log
tail -25 gpurun_out/pytest_multi.log
for w in 16 8; do
PTZ_CG_WPB=$w timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 30 --warmup 3 --config cfg5 --no-reloc --no-e2e > gpurun_out/bench_cfg5_2gpu_wpb$w.json 2> gpurun_out/bench_cfg5_2gpu_wpb$w.err; echo rc=$?
tail -c 1500 gpurun_out/bench_cfg5_2gpu_wpb$w.err
tail -c 4000 gpurun_out/bench_cfg5_2gpu_wpb$w.json
done
