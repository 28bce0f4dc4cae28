#!/bin/bash
# A/B timing of environment switches in ONE gpurun call (each call is charged >= 1.5-2.5 min of box time):
#   gpurun --timeout 900 -- 'scripts/gpu_ab.sh 400 3 "" "B200_SCHUR_44=1" "B200_ONE_STREAM=1 B200_GATHER_GLOBAL=1"; python -m pytest tests -m gpu -x -q | tail -2'
# args: N nfac variant... (a variant is a space-separated list of VAR=value, "" = defaults); prints the last
# factor/solve line of scripts/prof_one.py for each variant.
cd "$(dirname "$0")/.."
N=$1; NF=$2; shift 2
for v in "$@"; do
  printf '%-48s ' "[${v:-defaults}]"
  env $v python scripts/prof_one.py "$N" "$NF" 2>&1 | grep -E "^factor" | tail -1
done
