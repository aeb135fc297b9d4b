set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
for m in 0 1; do
PTZ_SYNC_COPY=$m timeout 600 python bench.py --steps 200 --warmup 3 --no-cpu --no-tracks --no-reloc --no-ncu > gpurun_out/bench_sync$m.json 2> gpurun_out/bench_sync$m.err; echo rc=$?
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_sync$m.json') if x.startswith('{')][-1]
j=json.loads(l)
for k in ('value','ms_per_step','lm_iters_per_sec','pcg_iterations_per_step','us_per_pcg_iteration','gpu_launches'): print(k, j[k])
print('sumk', sum(v['avg_us']*v['launches'] for v in j['kernels'].values())/j['steps'])
print('e2e', j['e2e'])
print('small', json.dumps(j['small_configs']))
PY
done
