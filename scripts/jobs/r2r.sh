mkdir -p gpurun_out
echo "== default"; timeout 200 python scripts/factor_timeline.py 400 2>&1 | grep -E "^factor|chain role|^gap|diagonal LDL|^rows|^update|^cb_up" | head -60
echo "== one stream"; B200_ONE_STREAM=1 timeout 200 python scripts/factor_timeline.py 400 gpurun_out/tl_one.txt 2>&1 | grep -E "^factor|chain role|^gap|diagonal LDL|^rows|^update" | head -60
for N in 400; do echo "== prof_one N=$N"; timeout 200 python scripts/prof_one.py $N 3 2>&1 | grep -E "^factor|resid" | tail -2 | cut -c1-250; done
