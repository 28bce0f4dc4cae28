mkdir -p gpurun_out
timeout 150 python scripts/gpu_solve_check.py 40 100 400 > gpurun_out/r2j_check.log 2>&1; rc=$?; echo "check rc=$rc"; grep -E "mbndry|ALL|MISMATCH|rror" gpurun_out/r2j_check.log | cut -c1-300
timeout 200 python scripts/factor_timeline.py 400 2>&1 | grep -v "^  *[0-9]* \(extend\|rows\|update\|schur\)" | tail -70
for N in 400 800; do echo "== prof_one N=$N"; timeout 200 python scripts/prof_one.py $N 3 2>&1 | grep -E "^factor|resid" | tail -2 | cut -c1-250; done
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "snapshot or small or kkt" 2>&1 | tail -3 | cut -c1-250
