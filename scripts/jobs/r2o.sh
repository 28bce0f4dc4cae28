mkdir -p gpurun_out
timeout 150 python scripts/gpu_solve_check.py 40 100 400 > gpurun_out/r2o_check.log 2>&1; rc=$?; echo "check rc=$rc"; grep -E "mbndry|ALL|MISMATCH|rror" gpurun_out/r2o_check.log | cut -c1-300
echo "== default"; timeout 200 python scripts/factor_timeline.py 400 2>&1 | grep -E "^factor|chain role|^gap|^cb_update|^ +[0-9]+ (chain|schur|extend|cb_update|front_smem|linv)" | head -60
echo "== CB at end"; B200_CB_AT_END=1 timeout 200 python scripts/factor_timeline.py 400 gpurun_out/tl_cbend.txt 2>&1 | grep -E "^factor|chain role|^gap" | head -12
for N in 400 800; do echo "== prof_one N=$N"; timeout 200 python scripts/prof_one.py $N 3 2>&1 | grep -E "^factor|resid" | tail -2 | cut -c1-250; done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not baseline_configs and not ip_loop_parity_full" 2>&1 | tail -3 | cut -c1-250
