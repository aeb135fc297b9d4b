set -x
mkdir -p gpurun_out
N=$1
nvidia-smi -L | wc -l
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 100 --warmup 3 > gpurun_out/bench_scale_$N.json 2> gpurun_out/bench_scale_$N.err; echo rc=$?
tail -c 500 gpurun_out/bench_scale_$N.err
python - <<PY
import json
j=json.loads([x for x in open('gpurun_out/bench_scale_$N.json') if x.startswith('{')][-1])
for k in ('value','ms_per_step','ms_per_step_kernels_only','pcg_iterations_per_step','us_per_pcg_iteration'): print(k, j[k])
print({k:(v['avg_us'],v['launches']) for k,v in j['kernels'].items()})
print('e2e', j['e2e']); print('reloc', j['reloc']['solves_per_sec'], j['reloc']['e2e_solves_per_sec'])
PY
