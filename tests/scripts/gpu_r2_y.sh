set -x
timeout 900 python -m pytest tests/test_gpu_ba.py tests/test_gpu_cpp_adaptor.py -m gpu -x -q -k "not multi_gpu" 2>&1 | tail -3
timeout 300 python bench.py --steps 200 --warmup 3 --no-cpu --no-tracks --no-small --no-ncu --no-e2e --no-reloc 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); k=d['kernels']; print('value', d['value'], 'ms/step', d['ms_per_step'], d['pcg_iterations_per_step'], ' '.join('%s=%.1f/%d'%(n, k[n]['avg_us'], k[n]['launches']) for n in k if k[n]['avg_us']>50))
"
