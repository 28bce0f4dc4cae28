mkdir -p gpurun_out
timeout 150 python scripts/gpu_solve_check.py 40 100 400 > gpurun_out/r2i_check.log 2>&1; rc=$?; echo "check rc=$rc"; grep -E "mbndry|ALL|MISMATCH|rror" gpurun_out/r2i_check.log | cut -c1-300
timeout 200 python scripts/factor_timeline.py 400 2>&1 | tail -80
timeout 200 python -m pytest tests/test_device_callers.py -x -q -m gpu 2>&1 | tail -3 | cut -c1-250
