set -x
mkdir -p gpurun_out
python tests/scripts/small_breakdown.py 2>&1 | tail -45
timeout 900 python bench.py --steps 200 --warmup 3 > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; echo rc=$?
tail -c 400 gpurun_out/bench_k.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_cfg4.csv python bench.py --steps 6 --warmup 3 --no-e2e --no-reloc --no-cpu --no-tracks --no-small --no-ncu > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_resjac|k_obs_what|k_schur_offdiag|k_track_accum|k_track_backsub|k_cost|k_cg' -s 40 -c 14 -o gpurun_out/r2_full python bench.py --steps 4 --warmup 3 --no-e2e --no-reloc --no-cpu --no-tracks --no-small --no-ncu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
