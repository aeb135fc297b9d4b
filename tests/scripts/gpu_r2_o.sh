set -x
mkdir -p gpurun_out
PTZ_OD_GROUP=8 timeout 900 python -m pytest tests/test_gpu_ba.py -m gpu -x -q -k "not multi_gpu" 2>&1 | tail -4
for h in 32 16 8; do
PTZ_OD_GROUP=$h timeout 300 python bench.py --steps 60 --warmup 3 --no-cpu --no-tracks --no-small --no-ncu --no-e2e --no-reloc 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('half=$h', d['value'], d['kernels']['schur_offdiag'])
"
done
