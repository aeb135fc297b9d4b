set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_ba.py -m gpu -x -q -k "multi_gpu" > gpurun_out/pytest_multi.log 2>&1; echo rc=$? >> gpurun_out/pytest_multi.log
tail -15 gpurun_out/pytest_multi.log
for d in 0 1; do
PTZ_CG_DEFLATE=$d timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 60 --warmup 3 --no-reloc --no-e2e > gpurun_out/bench2_defl$d.json 2> gpurun_out/bench2_defl$d.err; echo rc=$?
tail -c 3000 gpurun_out/bench2_defl$d.json
done
