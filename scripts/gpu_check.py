"""Quick GPU sanity run: factor+solve several synthetic KKTs, print inertia / residuals / timings."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from ipopt_b200 import B200Ldlt
from ipopt_b200.kkt import mbndry_kkt, lukvle1_kkt, random_kkt, to_scipy

def run(name, dim, irn, jcn, val, nc, first_val=None, **opts):
    s = B200Ldlt(verbose=2, **opts)
    assert s.InitializeStructure(dim, len(irn), irn, jcn) == 0
    a = s.GetValuesArrayPtr()
    a[:] = val if first_val is None else first_val
    st, neg = s.factor(True, nc)
    if first_val is not None:
        print(name, "first(W=0) status", st, "neg", neg, "expected", nc)
        a[:] = val
        st, neg = s.factor(True, nc)
    rng = np.random.default_rng(0)
    b = rng.standard_normal(dim)
    x = b.copy()
    st2 = s.solve(x)
    r, xi, bi = s.residual(x, b)
    info = s.info()
    print("%-22s dim=%d status=%d/%d neg=%d (exp %d) resid=%.2e |x|=%.2e 2x2=%d forced=%d tiny=%d growth=%d factor=%.3fms solve=%.3fms launches=%d/%d"
          % (name, dim, st, st2, neg, nc, r / (bi + 1e-300), xi, info["num_2x2"], info["num_forced"], info["num_tiny"], info["num_growth"],
             info["ms_factor_gpu"], info["ms_solve_gpu"], info["launches_factor"], info["launches_solve"]))
    for _ in range(3):
        st, neg = s.factor(True, nc)
        x = b.copy(); s.solve(x)
    info = s.info()
    print("    warm: factor=%.3fms solve=%.3fms" % (info["ms_factor_gpu"], info["ms_solve_gpu"]))
    s.close()
    return x

sizes = [int(a) for a in sys.argv[1:]] or [6, 30, 100]
dim, irn, jcn, val, nc = random_kkt(5, 2, density=0.5, seed=1)
run("random7", dim, irn, jcn, val, nc)
dim, irn, jcn, val, nc = random_kkt(300, 120, density=0.02, seed=1)
run("random420", dim, irn, jcn, val, nc)
for N in sizes:
    dim, irn, jcn, val, nc = mbndry_kkt(N, sigma_spread=3.0, seed=1)
    _, _, _, v0, _ = mbndry_kkt(N, w_zero=True)
    run("mbndry N=%d" % N, dim, irn, jcn, val, nc, first_val=v0)
dim, irn, jcn, val, nc = lukvle1_kkt(5000, sigma_spread=2.0, seed=1)
_, _, _, v0, _ = lukvle1_kkt(5000, w_zero=True)
run("lukvle1 N=5000", dim, irn, jcn, val, nc, first_val=v0)
