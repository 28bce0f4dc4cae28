mkdir -p gpurun_out
timeout 150 python scripts/gpu_solve_check.py 12 40 100 400 > gpurun_out/r2f_check.log 2>&1; rc=$?; echo "check rc=$rc"; grep -E "mbndry|ALL|MISMATCH|rror" gpurun_out/r2f_check.log | cut -c1-300
if [ $rc -ne 0 ]; then tail -20 gpurun_out/r2f_check.log; fi
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not baseline_configs and not ip_loop_parity_full" > gpurun_out/r2f_par.log 2>&1; echo "parity-subset rc=$?"; tail -8 gpurun_out/r2f_par.log | cut -c1-300
timeout 300 python -m pytest tests/test_vec_parity.py tests/test_device_callers.py tests/test_schur_tc.py -x -q -m gpu > gpurun_out/r2f_vec_f1.log 2>&1; echo "vec+f1+tc rc=$?"; tail -12 gpurun_out/r2f_vec_f1.log | cut -c1-300
for N in 400 800; do echo "== prof_one N=$N"; timeout 200 python scripts/prof_one.py $N 3 2>&1 | grep -E "^factor|resid" | cut -c1-250; done
B200_BENCH_SKIP_CPU=1 timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r2f_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','kkt_factor_solve_ms_per_iter')}, d['roofline']['frac'], d['parity'])"
