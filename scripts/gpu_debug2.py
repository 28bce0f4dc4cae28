import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from ipopt_b200 import B200Ldlt
from ipopt_b200.kkt import mbndry_kkt, to_scipy
def sres(dim, irn, jcn, val, x, b):
    A = to_scipy(dim, irn, jcn, val); r = A @ x - b
    return np.abs(r).max() / (abs(A).max() * np.abs(x).max() + np.abs(b).max())
for mode in (0, 2):
    s = B200Ldlt(use_graph=mode)
    for N in (8, 15, 8):
        dim, irn, jcn, val, nc = mbndry_kkt(N, sigma_spread=1.0, seed=N)
        s.InitializeStructure(dim, len(irn), irn, jcn)
        for rep in range(3):
            v = val.copy(); v[:len(v)//3] *= (1.0 + 0.1*rep)
            s.GetValuesArrayPtr()[:] = v
            st, neg = s.factor(True, nc)
            b = np.arange(1.0, dim+1); x = b.copy(); s.solve(x)
            i = s.info()
            print("mode", mode, "N", N, "rep", rep, "st", st, neg, nc, "res %.2e" % sres(dim, irn, jcn, v, x, b), "forced", i["num_forced"], "tiny", i["num_tiny"], "2x2", i["num_2x2"])
    s.close()
