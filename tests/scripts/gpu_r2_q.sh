set -x
N=$1
PTZ_DEFL_DEBUG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 60 --warmup 3 --no-reloc --no-e2e 2> gpurun_out/q_$N.err | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('N=$N', d['value'], d['ms_per_step'], d['pcg_iterations_per_step'], d['us_per_pcg_iteration'], d['kernels'].get('deflate'), d['timed_and_instrumented_pass_identical']); print({k:(v['avg_us'],v['launches']) for k,v in d['kernels'].items()})
"
grep "rejected\|deflation basis" gpurun_out/q_$N.err | head -12
