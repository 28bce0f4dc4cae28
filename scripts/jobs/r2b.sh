mkdir -p gpurun_out
timeout 150 python scripts/gpu_solve_check.py 6 12 40 100 > gpurun_out/r2b_check_small.log 2>&1; rc=$?; echo "check_small rc=$rc"; tail -12 gpurun_out/r2b_check_small.log | cut -c1-300
if [ $rc -ne 0 ]; then
  timeout 200 compute-sanitizer --tool memcheck --print-limit 20 python scripts/gpu_solve_check.py 12 > gpurun_out/r2b_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -v "^$" gpurun_out/r2b_memcheck.log | head -60 | cut -c1-250
  exit 1
fi
timeout 200 python scripts/gpu_solve_check.py 200 400 > gpurun_out/r2b_check_big.log 2>&1; echo "check_big rc=$?"; tail -12 gpurun_out/r2b_check_big.log | cut -c1-300
timeout 120 python scripts/solve_timeline.py 400 > gpurun_out/r2b_timeline.log 2>&1; echo "timeline rc=$?"; tail -3 gpurun_out/r2b_timeline.log
timeout 300 python -m pytest tests/test_vec_parity.py -x -q -m gpu > gpurun_out/r2b_vec.log 2>&1; echo "vec rc=$?"; tail -15 gpurun_out/r2b_vec.log | cut -c1-300
timeout 1100 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|error|final-iterate|Error|assert" gpurun_out/r2b_pytest.log | tail -20 | cut -c1-400
