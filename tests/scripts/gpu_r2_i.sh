set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ba.py -m gpu -x -q -k "multi_gpu" > gpurun_out/pytest_multi.log 2>&1; echo rc=$? >> gpurun_out/pytest_multi.log
grep -E "rank [01]\]|passed|failed|rc=" gpurun_out/pytest_multi.log | cut -c1-400 | tail -30
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 60 --warmup 3 --no-reloc > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo rc=$?
tail -c 600 gpurun_out/bench_2gpu.err
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_2gpu.json') if x.startswith('{')][-1]
j=json.loads(l)
for k in ('value','ms_per_step','ms_per_step_kernels_only','pcg_iterations_per_step','us_per_pcg_iteration'): print(k, j[k])
print({k:(v['avg_us'],v['launches']) for k,v in j['kernels'].items()})
print('e2e', j['e2e'])
PY
