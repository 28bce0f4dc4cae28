mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_vec_parity.py tests/test_device_callers.py -x -q -m gpu > gpurun_out/r2e_vec_f1.log 2>&1; echo "vec+f1 rc=$?"; tail -15 gpurun_out/r2e_vec_f1.log | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ma97" > gpurun_out/r2e_shim.log 2>&1; echo "shim rc=$?"; tail -15 gpurun_out/r2e_shim.log | cut -c1-300
B200_BENCH_SKIP_CPU=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2e_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2e_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_solve -c 2 --launch-skip 2 -f -o gpurun_out/r2e_solve python scripts/prof_one.py 400 2 > gpurun_out/r2e_ncu_solve.log 2>&1; echo "ncu solve rc=$?"
B200_TC_MIN_R=512 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tc_ -c 6 --launch-skip 6 -f -o gpurun_out/r2e_tc python scripts/prof_one.py 800 1 > gpurun_out/r2e_ncu_tc.log 2>&1; echo "ncu tc rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_big_schur84|k_big_diag|k_big_trsm|k_big_update|k_front_smem" -c 12 --launch-skip 300 -f -o gpurun_out/r2e_factor python scripts/prof_one.py 400 1 > gpurun_out/r2e_ncu_factor.log 2>&1; echo "ncu factor rc=$?"
ls -la gpurun_out/*.ncu-rep
