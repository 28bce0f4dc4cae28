mkdir -p gpurun_out
timeout 300 python scripts/gpu_solve_check.py 40 100 400 800 > gpurun_out/r2f3_check.log 2>&1; echo "check rc=$?"; grep -E "mbndry|ALL|MISMATCH|rror" gpurun_out/r2f3_check.log | cut -c1-300
for N in 400 800; do timeout 200 python scripts/prof_one.py $N 3 2>&1 | grep -E "^factor" | tail -1 | cut -c1-250; done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not baseline_configs and not ip_loop_parity_full" 2>&1 | tail -3 | cut -c1-250
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2final3_bench.json 2> gpurun_out/r2final3_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2final3_bench.json
timeout 200 python scripts/factor_timeline.py 400 gpurun_out/r2final3_factor_tl.txt > gpurun_out/r2final3_factor_tl_summary.txt 2>&1; head -45 gpurun_out/r2final3_factor_tl_summary.txt | cut -c1-100
