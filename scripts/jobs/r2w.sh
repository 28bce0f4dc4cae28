mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_front_warp|k_front_smem" -c 9 -f -o gpurun_out/r2w_fronts python scripts/prof_one.py 400 1 > gpurun_out/r2w_ncu_fronts.log 2>&1; echo "ncu fronts rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_solve -c 2 --launch-skip 2 -f -o gpurun_out/r2w_solve python scripts/prof_one.py 400 2 > gpurun_out/r2w_ncu_solve.log 2>&1; echo "ncu solve rc=$?"
for R in 0.0 0.02 0.05 0.08 0.1; do echo "== relax=$R"; B200_RELAX=$R timeout 200 python scripts/prof_one.py 400 3 2>&1 | grep -E "^factor|analyse:" | tail -2 | cut -c1-330; done
for R in 0.05 0.1; do echo "== N=800 relax=$R"; B200_RELAX=$R timeout 200 python scripts/prof_one.py 800 3 2>&1 | grep -E "^factor" | tail -1 | cut -c1-330; done
ls -la gpurun_out/*.ncu-rep
