set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 200 --warmup 3 --no-cpu --no-tracks --no-reloc --no-ncu > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; echo rc=$?
tail -c 1500 gpurun_out/bench_e.err
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_e.json') if x.startswith('{')][-1]
j=json.loads(l)
for k in ('value','ms_per_step','ms_per_step_with_per_kernel_events','ms_per_step_kernels_only','lm_iters_per_sec','pcg_iterations_per_step','us_per_pcg_iteration','gpu_launches'): print(k, j[k])
print('e2e', j['e2e'])
print('small', json.dumps(j['small_configs']))
PY
