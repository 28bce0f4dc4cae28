mkdir -p gpurun_out
timeout 120 python scripts/gpu_solve_check.py 12 100 400 > gpurun_out/r2c_check.log 2>&1; echo "check rc=$?"; grep -E "mbndry|ALL|MISMATCH" gpurun_out/r2c_check.log | cut -c1-300
timeout 120 python scripts/solve_timeline.py 400 > gpurun_out/r2c_timeline.log 2>&1; echo "timeline rc=$?"; grep -E "solve plan|solve ms|solve:" gpurun_out/r2c_timeline.log
timeout 300 python -m pytest tests/test_vec_parity.py -x -q -m gpu > gpurun_out/r2c_vec.log 2>&1; echo "vec rc=$?"; tail -12 gpurun_out/r2c_vec.log | cut -c1-300
timeout 300 python -m pytest tests/test_schur_tc.py -x -q -m gpu > gpurun_out/r2c_tc.log 2>&1; echo "tc rc=$?"; tail -25 gpurun_out/r2c_tc.log | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "inertia_restoration or sharded" > gpurun_out/r2c_par.log 2>&1; echo "parity-subset rc=$?"; tail -8 gpurun_out/r2c_par.log | cut -c1-300
B200_BENCH_SKIP_CPU=1 timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r2c_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','kkt_factor_solve_ms_per_iter','roofline')})"
