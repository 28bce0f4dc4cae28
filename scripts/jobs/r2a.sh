mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
timeout 600 python scripts/gpu_solve_check.py > gpurun_out/r2a_solve_check.log 2>&1; echo "solve_check rc=$?"
tail -25 gpurun_out/r2a_solve_check.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2a_pytest.log
timeout 300 python scripts/solve_timeline.py 400 > gpurun_out/r2a_timeline.log 2>&1; echo "timeline rc=$?"; tail -3 gpurun_out/r2a_timeline.log
B200_BENCH_SKIP_CPU=1 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/r2a_bench.log | cut -c1-1500
