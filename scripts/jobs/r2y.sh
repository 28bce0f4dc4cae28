mkdir -p gpurun_out
timeout 300 python scripts/gpu_solve_check.py 12 40 100 400 800 > gpurun_out/r2y_check.log 2>&1; rc=$?; echo "check rc=$rc"; grep -E "mbndry|lukvle|random|ALL|MISMATCH|rror" gpurun_out/r2y_check.log | cut -c1-330
for N in 400 800; do echo "== prof_one N=$N default"; timeout 200 python scripts/prof_one.py $N 3 2>&1 | grep -E "^factor|resid" | tail -3 | cut -c1-250; done
echo "== no warp2"; for N in 400 800; do B200_NO_WARP2=1 timeout 200 python scripts/prof_one.py $N 3 2>&1 | grep -E "^factor" | tail -1 | cut -c1-250; done
echo "== e2e breakdown"; timeout 200 python scripts/e2e_breakdown.py 400 2>&1 | tail -1
echo "== factor timeline"; timeout 200 python scripts/factor_timeline.py 400 gpurun_out/r2y_factor_tl.txt 2>&1 | grep -E "^factor| front_| chain  |chain role|^gap|front_smem phases" | head -40
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not baseline_configs and not ip_loop_parity_full" 2>&1 | tail -3 | cut -c1-250
