mkdir -p gpurun_out
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2final_bench.json 2> gpurun_out/r2final_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2final_bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2final_bench_reference.json 2>/dev/null; echo "ref rc=$?"
timeout 200 python scripts/solve_timeline.py 400 > /dev/null 2>&1; cp gpurun_out/solve_timeline.txt gpurun_out/r2final_solve_tl.txt
timeout 200 python scripts/factor_timeline.py 400 gpurun_out/r2final_factor_tl.txt > gpurun_out/r2final_factor_tl_summary.txt 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"k_solve|k_rhs_in|k_sol_out" --launch-skip 8 -c 8 -f -o gpurun_out/r2final_solve python scripts/prof_one.py 400 2 > gpurun_out/r2final_ncu_solve.log 2>&1; echo "ncu solve rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:"k_big_update_cb|k_big_update|k_big_panel" --launch-skip 300 -c 9 -f -o gpurun_out/r2final_schur python scripts/prof_one.py 400 1 > gpurun_out/r2final_ncu_schur.log 2>&1; echo "ncu schur rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:"k_front_warp|k_front_smem" -c 8 -f -o gpurun_out/r2final_fronts python scripts/prof_one.py 400 1 > gpurun_out/r2final_ncu_fronts.log 2>&1; echo "ncu fronts rc=$?"
python scripts/ncu_extract.py r2 gpurun_out/r2final_solve.ncu-rep gpurun_out/r2final_schur.ncu-rep gpurun_out/r2final_fronts.ncu-rep > gpurun_out/r2final_extract.log 2>&1; echo "extract rc=$?"
cp profiles/r2_r2final_*_raw.csv profiles/r2_ncu_metrics.json gpurun_out/
for r in solve schur fronts; do ncu -i gpurun_out/r2final_$r.ncu-rep --page details --csv > gpurun_out/r2final_${r}_details.csv 2>/dev/null; done
rm -f gpurun_out/r2final_schur.ncu-rep gpurun_out/r2final_fronts.ncu-rep
B200_BENCH_SKIP_CPU=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2final_launches_bench.csv python bench.py --steps 1 --warmup 1 > gpurun_out/r2final_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
cuobjdump -sass ipopt_b200/lib/libb200ldlt.so | grep -oE "^\s+/\*[0-9a-f]+\*/\s+[A-Z0-9_.]+" | awk '{print $2}' | sed 's/\..*//' | sort | uniq -c | sort -rn > gpurun_out/r2final_sass_opcode_histogram.txt
du -sh gpurun_out; ls -la gpurun_out | head -40
