mkdir -p gpurun_out
timeout 150 python scripts/gpu_solve_check.py 100 400 > gpurun_out/r2l_check.log 2>&1; rc=$?; echo "check rc=$rc"; grep -E "mbndry|ALL|MISMATCH|rror" gpurun_out/r2l_check.log | cut -c1-300
echo "== CB per panel (default)"; timeout 200 python scripts/factor_timeline.py 400 2>&1 | grep -E "^factor|chain role|^gap|pivot paths|^cb_update|^ +1[0-4] (chain|schur|extend|cb_update)" | head -40
echo "== CB at end"; B200_CB_AT_END=1 timeout 200 python scripts/factor_timeline.py 400 gpurun_out/tl_cbend.txt 2>&1 | grep -E "^factor|chain role|^gap" | head -12
