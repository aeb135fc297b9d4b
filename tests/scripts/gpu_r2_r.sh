set -x
python tests/scripts/small_breakdown.py 2>&1 | grep -E "cfg|pcg|deflate"
timeout 300 python bench.py --steps 120 --warmup 3 --no-cpu --no-tracks --no-small --no-ncu --no-e2e --no-reloc 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); k=d['kernels']; print('value', d['value'], 'ms/step', d['ms_per_step'], ' '.join('%s=%.1f'%(n, k[n]['avg_us']) for n in k if k[n]['avg_us']>20))
"
