mkdir -p gpurun_out
timeout 300 python scripts/gpu_solve_check.py 40 100 400 800 > gpurun_out/r2v_check.log 2>&1; rc=$?; echo "check rc=$rc"; grep -E "mbndry|lukvle|random|ALL|MISMATCH|rror" gpurun_out/r2v_check.log | cut -c1-330
echo "== check with B200_NO_WARP2=1"; B200_NO_WARP2=1 timeout 300 python scripts/gpu_solve_check.py 40 100 400 2>&1 | grep -E "mbndry|ALL|MISMATCH|rror" | cut -c1-330
echo "== no warp2"; B200_NO_WARP2=1 timeout 200 python scripts/prof_one.py 400 3 2>&1 | grep -E "^factor|resid" | tail -2 | cut -c1-250
if [ $rc -ne 0 ]; then export B200_NO_WARP2=1; echo "== retry without PDL"; B200_NO_PDL=1 timeout 300 python scripts/gpu_solve_check.py 100 400 2>&1 | grep -E "mbndry|ALL|MISMATCH|rror" | cut -c1-300; echo "== retry without pairing"; B200_SOLVE_NOPAIR=1 timeout 300 python scripts/gpu_solve_check.py 100 400 2>&1 | grep -E "mbndry|ALL|MISMATCH|rror" | cut -c1-300; fi
P='grep -E "^factor|resid|solve plan"'
for N in 400 800; do echo "== prof_one N=$N default"; timeout 200 python scripts/prof_one.py $N 3 2>&1 | grep -E "^factor|resid|solve plan" | tail -3 | cut -c1-250; done
echo "== no PDL"; B200_NO_PDL=1 timeout 200 python scripts/prof_one.py 400 3 2>&1 | grep -E "^factor" | tail -1 | cut -c1-250
echo "== no pairing"; B200_SOLVE_NOPAIR=1 timeout 200 python scripts/prof_one.py 400 3 2>&1 | grep -E "^factor" | tail -1 | cut -c1-250
for K in 8 16 24 48; do echo "== leaf_k=$K"; B200_LEAF_K=$K timeout 200 python scripts/prof_one.py 400 3 2>&1 | grep -E "^factor|analyse:" | tail -2 | cut -c1-330; done
for R in 0.05 0.3; do echo "== relax=$R"; B200_RELAX=$R timeout 200 python scripts/prof_one.py 400 3 2>&1 | grep -E "^factor|analyse:" | tail -2 | cut -c1-330; done
echo "== TC schur N=800 min_r=512"; B200_TC_MIN_R=512 timeout 200 python scripts/prof_one.py 800 3 2>&1 | grep -E "^factor|resid" | tail -2 | cut -c1-250
echo "== TC schur N=800 min_r=1024"; B200_TC_MIN_R=1024 timeout 200 python scripts/prof_one.py 800 3 2>&1 | grep -E "^factor|resid" | tail -2 | cut -c1-250
echo "== e2e breakdown"; timeout 200 python scripts/e2e_breakdown.py 400 2>&1 | tail -1
echo "== solve timeline"; timeout 200 python scripts/solve_timeline.py 400 2>&1 | tail -3
echo "== factor timeline"; timeout 200 python scripts/factor_timeline.py 400 gpurun_out/r2v_factor_tl.txt 2>&1 | grep -E "^factor|chain role|^gap|diagonal LDL" | head
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not baseline_configs and not ip_loop_parity_full" 2>&1 | tail -3 | cut -c1-250
