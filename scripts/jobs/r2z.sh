mkdir -p gpurun_out
for M in 64 96; do echo "== smem_front_max=$M"; for N in 400 800; do B200_SMEM_MAX=$M timeout 200 python scripts/prof_one.py $N 3 2>&1 | grep -E "^factor|resid" | tail -2 | cut -c1-250; done; done
echo "== buckets"; for B in "80,96,112" "72,88,104,120"; do echo "B200_BUCKETS=$B"; B200_BUCKETS=$B timeout 200 python scripts/prof_one.py 400 3 2>&1 | grep -E "^factor" | tail -1 | cut -c1-250; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_front_smem" -c 6 --launch-skip 8 -f -o gpurun_out/r2z_smem python scripts/prof_one.py 400 1 > gpurun_out/r2z_ncu.log 2>&1; echo "ncu rc=$?"
