mkdir -p gpurun_out
for N in 400 800; do timeout 200 python scripts/prof_one.py $N 3 2>&1 | grep -E "^factor|resid" | tail -2 | cut -c1-250; done
timeout 1000 python -m pytest tests -x -q -m gpu > gpurun_out/r2final4_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2final4_pytest.log | cut -c1-300
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2final4_bench.json 2> gpurun_out/r2final4_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2final4_bench.json
timeout 200 python scripts/factor_timeline.py 400 gpurun_out/r2final4_factor_tl.txt > gpurun_out/r2final4_factor_tl_summary.txt 2>&1; head -30 gpurun_out/r2final4_factor_tl_summary.txt | cut -c1-100
