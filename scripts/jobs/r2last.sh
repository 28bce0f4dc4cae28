mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not baseline_configs and not ip_loop_parity_full" 2>&1 | tail -2 | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2last_bench.json 2> gpurun_out/r2last_bench.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/r2last_bench.json
