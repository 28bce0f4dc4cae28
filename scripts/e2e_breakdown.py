"""Where the end-to-end (host-buffer) time of one factor + solve goes: wall clock around each C-ABI call vs its device time."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from ipopt_b200 import B200Ldlt
from ipopt_b200.kkt import mbndry_kkt
N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
dim, irn, jcn, val, nc = mbndry_kkt(N, sigma_spread=3.0, seed=1)
_, _, _, v0, _ = mbndry_kkt(N, w_zero=True)
s = B200Ldlt()
s.InitializeStructure(dim, len(irn), irn, jcn)
a = s.GetValuesArrayPtr()
a[:] = v0
s.factor(True, nc)
b = np.random.default_rng(0).standard_normal(dim)
acc = {"fill": 0.0, "factor_wall": 0.0, "factor_gpu": 0.0, "rhs_copy": 0.0, "solve_wall": 0.0, "solve_gpu": 0.0}
R = 10
for it in range(R + 2):
    t0 = time.perf_counter(); a[:] = val; t1 = time.perf_counter()
    s.factor(True, nc); t2 = time.perf_counter()
    fg = s.info()["ms_factor_gpu"]
    x = b.copy(); t3 = time.perf_counter()
    s.solve(x); t4 = time.perf_counter()
    sg = s.info()["ms_solve_gpu"]
    if it >= 2:
        acc["fill"] += (t1 - t0) * 1e3; acc["factor_wall"] += (t2 - t1) * 1e3; acc["factor_gpu"] += fg
        acc["rhs_copy"] += (t3 - t2) * 1e3; acc["solve_wall"] += (t4 - t3) * 1e3; acc["solve_gpu"] += sg
print("e2e breakdown N=%d (ms, mean of %d):" % (N, R), {k: round(v / R, 3) for k, v in acc.items()})
