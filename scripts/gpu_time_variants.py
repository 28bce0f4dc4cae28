import sys, os
import numpy as np
sys.path.insert(0, ".")
from ipopt_b200 import B200Ldlt
from ipopt_b200.kkt import mbndry_kkt
dim, irn, jcn, val, nc = mbndry_kkt(400, sigma_spread=3.0, seed=1)
for g in (1, 0):
    s = B200Ldlt(use_graph=g)
    s.InitializeStructure(dim, len(irn), irn, jcn)
    s.GetValuesArrayPtr()[:] = val
    for _ in range(4):
        st, neg = s.factor(True, nc)
    print("one_stream", os.environ.get("B200_ONE_STREAM"), "graph", g, "factor ms", s.info()["ms_factor_gpu"], st, neg)
    s.close()
