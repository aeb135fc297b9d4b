set -x
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 200 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo rc=$?
tail -c 600 gpurun_out/bench1.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench1.json') if x.startswith('{')][-1]
j=json.loads(l)
for k in ('value','ms_per_step','lm_iters_per_sec','pcg_iterations_per_step','us_per_pcg_iteration','gpu_launches'): print(k, j[k])
print('roofline', json.dumps(j['roofline'])[:900])
for r in j['roofline_kernels']: print('  ', r['kernel'], r['avg_us'], r['achieved'], r['frac'], r['traffic'], r['algorithmic_bytes_per_launch'], r['share_of_step'])
print('e2e', j['e2e'])
print('small', json.dumps(j['small_configs']))
print('reloc', json.dumps(j['reloc']))
print('cpu', json.dumps(j['cpu_baseline']))
print('clocks', j['clocks'])
PY
