set -x
PTZ_DEFL_DEBUG=1 python tests/scripts/defl_trace1.py 4000 8000 2>&1 | grep -E "^V|pcg per|rejected|polished" | cut -c1-260
timeout 600 python -m pytest tests/test_gpu_ba.py -m gpu -x -q -k "deflated or cfg4 or full_size or handle_reuse" 2>&1 | tail -3
timeout 300 python bench.py --steps 120 --warmup 3 --no-cpu --no-tracks --no-small --no-ncu --no-e2e --no-reloc 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); k=d['kernels']; print('value', d['value'], 'ms/step', d['ms_per_step'], d['pcg_iterations_per_step'], ' '.join('%s=%.1f'%(n, k[n]['avg_us']) for n in k if k[n]['avg_us']>60))
"
