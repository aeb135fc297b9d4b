#!/bin/bash
# A/B of library builds on the GPU box: for every ptz-calib_b200/csrc/ab/*.so run the kernel-only bench and print the per-kernel times
cd "$(dirname "$0")/../.."
for so in ptz-calib_b200/csrc/ab/*.so; do
  cp "$so" ptz-calib_b200/csrc/libptzcalib_b200.so
  echo "== $so"
  timeout 200 python bench.py --steps 120 --warmup 3 --no-e2e --no-cpu --no-tracks $AB_FLAGS 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    k=d['kernels']
    if d.get('reloc'): print('reloc solves/s %.3e  ms/batch %.3f  e2e %.3e' % (d['reloc']['solves_per_sec'], d['reloc']['ms_per_batch'], d['reloc']['e2e_solves_per_sec']))
    print('value', d['value'], 'ms/step', d['ms_per_step'], ' '.join('%s=%.1f'%(n, k[n]['avg_us']) for n in k if k[n]['avg_us']>20))
"
done
