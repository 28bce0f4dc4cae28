#!/bin/bash
# usage: scripts/gpu_driver_runs.sh "<problem> <N>" ...   (runs the b200 backend through the reference IP loop)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for p in "$@"; do set -- $p
  echo "=== $1 $2"
  ./tests/driver/ipopt_driver --backend b200 --problem $1 --N ${2:-0} --print-level ${PL:-5} --opt print_timing_statistics=yes ${EXTRA} \
     --json gpurun_out/b200_$1_${2:-0}.json --final gpurun_out/b200_final_$1_${2:-0}.bin > gpurun_out/b200_$1_${2:-0}.log 2>&1
  echo "rc=$?"; grep -E "^ +[0-9]+ |EXIT|DRIVER_JSON|LinearSystem" gpurun_out/b200_$1_${2:-0}.log | tail -${TAILN:-12} | cut -c1-400
done
