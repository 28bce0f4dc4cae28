mkdir -p gpurun_out
timeout 150 python scripts/gpu_solve_check.py 12 40 100 400 > gpurun_out/r2h_check.log 2>&1; rc=$?; echo "check rc=$rc"; grep -E "mbndry|ALL|MISMATCH|rror" gpurun_out/r2h_check.log | cut -c1-300
if [ $rc -ne 0 ]; then tail -20 gpurun_out/r2h_check.log; fi
for N in 400; do echo "== prof_one N=$N"; timeout 200 python scripts/prof_one.py $N 3 2>&1 | grep -E "^factor|resid" | cut -c1-250; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_big_panel -c 3 --launch-skip 150 -f -o gpurun_out/r2h_panel python scripts/prof_one.py 400 1 > gpurun_out/r2h_ncu.log 2>&1; echo "ncu rc=$?"
timeout 200 python -m pytest tests/test_device_callers.py -x -q -m gpu 2>&1 | tail -40 | cut -c1-250
