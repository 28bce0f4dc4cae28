mkdir -p gpurun_out
timeout 250 python scripts/gpu_solve_check.py 100 400 800 > gpurun_out/r2q_check.log 2>&1; rc=$?; echo "check rc=$rc"; grep -E "mbndry|ALL|MISMATCH|rror" gpurun_out/r2q_check.log | cut -c1-300
echo "== default"; timeout 200 python scripts/factor_timeline.py 400 2>&1 | grep -E "^factor|chain role|^gap|^ +[0-9]+ (extend|chain)|^extend|^rows|^update|^cb_up" | head -60
for N in 400 800; do echo "== prof_one N=$N"; timeout 200 python scripts/prof_one.py $N 3 2>&1 | grep -E "^factor|resid" | tail -2 | cut -c1-250; done
echo "== CB at end"; for N in 400 800; do B200_CB_AT_END=1 timeout 200 python scripts/prof_one.py $N 3 2>&1 | grep -E "^factor" | tail -1 | cut -c1-250; done
echo "== buckets"; for B in 48 "48,96" "40,48,56,80,96,112"; do echo "B200_BUCKETS=$B"; B200_BUCKETS=$B timeout 200 python scripts/prof_one.py 400 3 2>&1 | grep -E "^factor" | tail -1 | cut -c1-250; done
