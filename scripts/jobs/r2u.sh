mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2u_pytest.log | cut -c1-300
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"; cut -c1-2500 gpurun_out/r2u_bench.json; tail -3 gpurun_out/r2u_bench.err
for N in 400 800; do echo "== prof_one N=$N"; timeout 200 python scripts/prof_one.py $N 3 2>&1 | grep -E "^factor|resid" | tail -2 | cut -c1-250; done
echo "== factor timeline"; timeout 200 python scripts/factor_timeline.py 400 gpurun_out/r2u_factor_tl.txt 2>&1 | tail -60
echo "== solve timeline"; timeout 200 python scripts/solve_timeline.py 400 2>&1 | tail -40
B200_BENCH_SKIP_CPU=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2u_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2u_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
python scripts/agg_launches.py gpurun_out/r2u_launches_bench.csv
