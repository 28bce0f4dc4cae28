import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from ipopt_b200 import B200Ldlt
from ipopt_b200.kkt import mbndry_kkt
for N in (40, 60, 100):
    dim, irn, jcn, val, nc = mbndry_kkt(N, sigma_spread=2.0, seed=2)
    for g in (0, 1):
        s = B200Ldlt(use_graph=g)
        s.InitializeStructure(dim, len(irn), irn, jcn)
        s.GetValuesArrayPtr()[:] = val
        st, neg = s.factor(True, nc)
        b = np.ones(dim); x = b.copy(); s.solve(x)
        r = s.residual(x, b)
        i = s.info()
        print("N", N, "graph", g, "st", st, neg, nc, "res %.2e" % (r[0] / (r[1] + r[2])), "maxfront", i["max_front"], "forced", i["num_forced"], "growth", i["num_growth"])
        s.close()
