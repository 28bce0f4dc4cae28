import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from ipopt_b200 import B200Ldlt
from ipopt_b200.kkt import mbndry_kkt
N = int(sys.argv[1]) if len(sys.argv) > 1 else 15
dim, irn, jcn, val, nc = mbndry_kkt(N, sigma_spread=1.0, seed=N)
s = B200Ldlt(use_graph=int(sys.argv[2]) if len(sys.argv) > 2 else 0)
s.InitializeStructure(dim, len(irn), irn, jcn)
s.GetValuesArrayPtr()[:] = val
print(s.factor(True, nc))
b = np.arange(1.0, dim+1); x = b.copy(); print(s.solve(x)); print(s.residual(x, b))
print(s.factor(True, nc), s.last_error())
