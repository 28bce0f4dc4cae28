mkdir -p gpurun_out
timeout 300 python scripts/gpu_solve_check.py 12 40 100 400 800 > gpurun_out/r2x_check.log 2>&1; rc=$?; echo "check rc=$rc"; grep -E "mbndry|lukvle|random|ALL|MISMATCH|rror" gpurun_out/r2x_check.log | cut -c1-330
for D in 0 1 2 3 4; do echo "== SOLVE_DIRECT=$D"; B200_SOLVE_DIRECT=$D timeout 200 python scripts/gpu_solve_check.py 100 400 2>&1 | grep -E "mbndry|ALL|MISMATCH|rror" | cut -c1-330; B200_SOLVE_DIRECT=$D timeout 200 python scripts/prof_one.py 400 3 2>&1 | grep -E "^factor|solve plan" | tail -3 | cut -c1-250; done
for D in 1 2; do echo "== N=800 SOLVE_DIRECT=$D"; B200_SOLVE_DIRECT=$D timeout 200 python scripts/prof_one.py 800 3 2>&1 | grep -E "^factor|solve plan" | tail -3 | cut -c1-250; done
echo "== nopair D=1"; B200_SOLVE_NOPAIR=1 timeout 200 python scripts/prof_one.py 400 3 2>&1 | grep -E "^factor" | tail -1 | cut -c1-250
echo "== solve timeline (default)"; timeout 200 python scripts/solve_timeline.py 400 2>&1 | tail -2
cp gpurun_out/solve_timeline.txt gpurun_out/r2x_solve_tl_d1.txt
echo "== solve timeline D=2"; B200_SOLVE_DIRECT=2 timeout 200 python scripts/solve_timeline.py 400 2>&1 | tail -2
cp gpurun_out/solve_timeline.txt gpurun_out/r2x_solve_tl_d2.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not baseline_configs and not ip_loop_parity_full" 2>&1 | tail -3 | cut -c1-250
