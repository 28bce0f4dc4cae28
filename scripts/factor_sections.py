"""Per-level section times of one factorisation (plain launches, CUDA events between sections)."""
import os, sys
os.environ["B200_FACTOR_SECTIONS"] = "1"
import numpy as np
sys.path.insert(0, ".")
from ipopt_b200 import B200Ldlt
from ipopt_b200.kkt import mbndry_kkt
N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
dim, irn, jcn, val, nc = mbndry_kkt(N, sigma_spread=3.0, seed=1)
s = B200Ldlt(use_graph=0)
s.InitializeStructure(dim, len(irn), irn, jcn)
a = s.GetValuesArrayPtr()
a[:] = val
for it in range(3):
    if it == 2: print("---- last run ----", file=sys.stderr)
    print(s.factor(True, nc), s.info()["ms_factor_gpu"])
