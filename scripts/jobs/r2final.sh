mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 1000 python -m pytest tests -x -q -m gpu > gpurun_out/r2final_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2final_pytest.log | cut -c1-300
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2final_bench.json 2> gpurun_out/r2final_bench.err; echo "bench rc=$?"; cut -c1-1800 gpurun_out/r2final_bench.json; tail -3 gpurun_out/r2final_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2final_bench_reference.json 2>/dev/null; echo "ref rc=$?"; cut -c1-400 gpurun_out/r2final_bench_reference.json
for N in 400 800; do echo "== prof_one N=$N"; timeout 200 python scripts/prof_one.py $N 3 2>&1 | grep -E "^factor|resid|solve plan" | tail -4 | cut -c1-250; done
echo "== solve timeline"; timeout 200 python scripts/solve_timeline.py 400 2>&1 | tail -2; cp gpurun_out/solve_timeline.txt gpurun_out/r2final_solve_tl.txt
echo "== factor timeline"; timeout 200 python scripts/factor_timeline.py 400 gpurun_out/r2final_factor_tl.txt 2>&1 | grep -E "^factor|chain role|^gap|diagonal LDL" | head
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_solve|k_rhs_in|k_sol_out" --launch-skip 8 -c 8 -f -o gpurun_out/r2final_solve python scripts/prof_one.py 400 2 > gpurun_out/r2final_ncu_solve.log 2>&1; echo "ncu solve rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_big_update_cb|k_big_update|k_big_panel" --launch-skip 300 -c 12 -f -o gpurun_out/r2final_schur python scripts/prof_one.py 400 1 > gpurun_out/r2final_ncu_schur.log 2>&1; echo "ncu schur rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_front_warp|k_front_smem" -c 10 -f -o gpurun_out/r2final_fronts python scripts/prof_one.py 400 1 > gpurun_out/r2final_ncu_fronts.log 2>&1; echo "ncu fronts rc=$?"
B200_BENCH_SKIP_CPU=1 timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2final_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2final_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
python scripts/agg_launches.py gpurun_out/r2final_launches_bench.csv 2>&1 | tail -25
ls -la gpurun_out/*.ncu-rep
