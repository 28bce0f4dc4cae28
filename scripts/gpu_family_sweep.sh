#!/bin/bash
# GPU backend vs CPU oracle through the reference IP loop on a sweep of reference ScalableProblems.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export OMP_NUM_THREADS=16
for p in "LukVlE1 5000" "LukVlI1 5000" "LukVlE2 2000" "LukVlE5 2000" "MBndryCntrl1 100" "MBndryCntrl2 60" "MBndryCntrl3 60" "MBndryCntrl4 60" "MBndryCntrl5 60" "MDistCntrl1 60" "MDistCntrl2 60" "MDistCntrl3 60" "MDistCntrl3a 100" "MDistCntrl4 60" "MDistCntrl3a 600"; do set -- $p
  for be in b200 oracle; do
    timeout 900 ./tests/driver/ipopt_driver --backend $be --problem $1 --N $2 --print-level 0 --json gpurun_out/sweep_${be}_$1_$2.json > /dev/null 2>&1
  done
  python - "$1" "$2" <<'PY'
import json, sys
p, n = sys.argv[1], sys.argv[2]
try:
    a = json.load(open("gpurun_out/sweep_b200_%s_%s.json" % (p, n))); b = json.load(open("gpurun_out/sweep_oracle_%s_%s.json" % (p, n)))
    print("%-13s N=%-5s dim=%-8d | b200: st=%d it=%-3d nf=%-3d sing=%d wi=%d obj=%.12e fac=%.1fms sol=%.2fms | oracle: st=%d it=%-3d nf=%-3d sing=%d wi=%d obj=%.12e fac=%.0fms | rel.obj.diff=%.1e"
          % (p, n, a["kkt_dim"], a["status"], a["iterations"], a["n_factor"], a["n_singular"], a["n_wrong_inertia"], a["objective"], 1e3*a["t_factor_s"]/max(a["n_factor"]-1,1), 1e3*a["t_solve_s"]/max(a["n_solve"],1),
             b["status"], b["iterations"], b["n_factor"], b["n_singular"], b["n_wrong_inertia"], b["objective"], 1e3*b["t_factor_s"]/max(b["n_factor"]-1,1), abs(a["objective"]-b["objective"])/max(abs(b["objective"]),1e-300)))
except Exception as e:
    print(p, n, "FAILED", e)
PY
done
