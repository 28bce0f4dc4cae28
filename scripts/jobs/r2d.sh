mkdir -p gpurun_out
timeout 150 python scripts/gpu_solve_check.py 12 100 400 > gpurun_out/r2d_check.log 2>&1; rc=$?; echo "check rc=$rc"; grep -E "mbndry|ALL|MISMATCH|rror" gpurun_out/r2d_check.log | cut -c1-300
if [ $rc -ne 0 ]; then tail -20 gpurun_out/r2d_check.log; exit 1; fi
timeout 300 python -m pytest tests/test_schur_tc.py tests/test_vec_parity.py -x -q -m gpu > gpurun_out/r2d_tcvec.log 2>&1; echo "tc+vec rc=$?"; tail -12 gpurun_out/r2d_tcvec.log | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sharded or golden or synthetic or full_size_properties" > gpurun_out/r2d_par.log 2>&1; echo "parity-subset rc=$?"; tail -5 gpurun_out/r2d_par.log | cut -c1-300
for tc in 0 512 256; do for N in 400 800; do echo "== prof_one N=$N tc_min_r=$tc"; B200_TC_MIN_R=$tc timeout 200 python scripts/prof_one.py $N 3 2>&1 | grep -E "^factor|resid" | cut -c1-250; done; done
timeout 120 python scripts/solve_timeline.py 400 > gpurun_out/r2d_timeline.log 2>&1; echo "timeline rc=$?"; grep -E "solve ms" gpurun_out/r2d_timeline.log
B200_BENCH_SKIP_CPU=1 timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r2d_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','kkt_factor_solve_ms_per_iter')}, d['roofline']['frac'])"
