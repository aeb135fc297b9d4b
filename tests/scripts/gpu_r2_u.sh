set -x
timeout 900 python -m pytest tests/test_gpu_ba.py -m gpu -x -q -k "not multi_gpu" 2>&1 | tail -4
for w in 0 1; do
PTZ_BYTRACK_WARP=$w timeout 300 python bench.py --steps 120 --warmup 3 --no-cpu --no-tracks --no-small --no-ncu --no-e2e --no-reloc 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); k=d['kernels']; print('warp=$w value', d['value'], 'ms/step', d['ms_per_step'], ' '.join('%s=%.1f'%(n, k[n]['avg_us']) for n in k if k[n]['avg_us']>8))
"
done
