mkdir -p gpurun_out
timeout 250 python scripts/gpu_solve_check.py 100 400 800 > gpurun_out/r2t_check.log 2>&1; rc=$?; echo "check rc=$rc"; grep -E "mbndry|ALL|MISMATCH|rror" gpurun_out/r2t_check.log | cut -c1-300
echo "== default"; timeout 200 python scripts/factor_timeline.py 400 2>&1 | grep -E "^factor|chain role|^gap|diagonal LDL|linv" | head -60
for N in 400 800; do echo "== prof_one N=$N"; timeout 200 python scripts/prof_one.py $N 3 2>&1 | grep -E "^factor|resid" | tail -2 | cut -c1-250; done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not baseline_configs and not ip_loop_parity_full" 2>&1 | tail -3 | cut -c1-250
