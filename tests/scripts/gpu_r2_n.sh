set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_reloc.py tests/test_gpu_cpp_adaptor.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 40 --warmup 3 --no-cpu --no-tracks --no-small --no-ncu --no-e2e > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err; echo rc=$?
python - <<PY
import json
j=json.loads([x for x in open('gpurun_out/bench_n.json') if x.startswith('{')][-1])
print(j['value'], j['reloc'])
PY
