#!/bin/bash
cd "$(dirname "$0")/.."
EXTRA="--opt b200_verbose=2" PL=0 scripts/gpu_driver_runs.sh "MBndryCntrl1 400" > /dev/null 2>&1
grep -E "b200ldlt\] factor" gpurun_out/b200_MBndryCntrl1_400.log | awk '{print $3, $4, $7, $8, $9, $10, $11}' | tail -19
grep DRIVER_JSON gpurun_out/b200_MBndryCntrl1_400.log | cut -c1-700
