set -x
for g in 16 8; do
PTZ_DEFL_DEBUG=1 PTZ_OD_GROUP=$g timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 60 --warmup 3 --no-reloc --no-e2e 2> gpurun_out/p_$g.err | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('group=$g', d['value'], d['pcg_iterations_per_step'], d['kernels'].get('deflate'), d['kernels']['schur_offdiag']['avg_us'], d['timed_and_instrumented_pass_identical'])
"
grep "rejected" gpurun_out/p_$g.err | head -5
done
