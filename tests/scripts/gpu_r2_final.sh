set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo rc=$? >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo rc=$?
tail -c 300 gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_reference.json 2>> gpurun_out/bench_final.err; echo rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_cfg4_final.csv python bench.py --steps 6 --warmup 3 --no-e2e --no-reloc --no-cpu --no-tracks --no-small --no-ncu > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_resjac|k_obs_what|k_schur_offdiag|k_track_accum|k_track_backsub|k_cost|k_cg' -s 40 -c 14 -o gpurun_out/r2_full_final python bench.py --steps 4 --warmup 3 --no-e2e --no-reloc --no-cpu --no-tracks --no-small --no-ncu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
