"""GPU check of the subtree / task-queue solve (default) against the level-per-launch solve kernels (use_graph=2)
on the same factors, plus timings.  python scripts/gpu_solve_check.py [N ...]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from ipopt_b200 import B200Ldlt
from ipopt_b200.kkt import mbndry_kkt, lukvle1_kkt, random_kkt

def one(name, dim, irn, jcn, v0, val, nc, reps=3):
    b = np.random.default_rng(0).standard_normal(dim)
    out = {}
    for mode in (1, 2):
        s = B200Ldlt(use_graph=mode)
        s.InitializeStructure(dim, len(irn), irn, jcn)
        a = s.GetValuesArrayPtr()
        a[:] = v0
        s.factor(False, 0)
        a[:] = val
        st, neg = s.factor(nc is not None, nc or 0)
        xs = []
        for it in range(reps):
            x = b.copy(); s.solve(x); xs.append(x)
        i = s.info()
        r, xi, bi = s.residual(xs[-1], b)
        out[mode] = (xs, i, st, neg, r / (bi + 1e-300))
        s.close()
    x1, x2 = out[1][0][-1], out[2][0][-1]
    det = all(np.array_equal(out[1][0][0], xx) for xx in out[1][0][1:])
    rel = np.linalg.norm(x1 - x2) / np.linalg.norm(x2)
    i1, i2 = out[1][1], out[2][1]
    print("%-22s dim %8d st %d/%d neg %d/%d  rel.diff new-vs-level %.2e  resid %.2e/%.2e  deterministic %s  factor %.3f ms  solve new %.3f ms (%d launches) level %.3f ms"
          % (name, dim, out[1][2], out[2][2], out[1][3], out[2][3], rel, out[1][4], out[2][4], det, i1["ms_factor_gpu"], i1["ms_solve_gpu"], i1["launches_solve"], i2["ms_solve_gpu"]), flush=True)
    return rel < 1e-9 and det

ok = True
sizes = [int(a) for a in sys.argv[1:]] or [6, 12, 40, 100, 200, 400]
for N in sizes:
    dim, irn, jcn, val, nc = mbndry_kkt(N, sigma_spread=3.0, seed=1)
    _, _, _, v0, _ = mbndry_kkt(N, w_zero=True)
    ok &= one("mbndry N=%d" % N, dim, irn, jcn, v0, val, nc)
for N in (50, 3000, 50000):
    dim, irn, jcn, val, nc = lukvle1_kkt(N, sigma_spread=2.0, seed=3)
    _, _, _, v0, _ = lukvle1_kkt(N, w_zero=True)
    ok &= one("lukvle1 N=%d" % N, dim, irn, jcn, v0, val, nc)
for seed in range(3):
    dim, irn, jcn, val, nc = random_kkt(400 + 300 * seed, 150 + 100 * seed, density=0.02, seed=seed)
    ok &= one("random seed %d" % seed, dim, irn, jcn, val, val, None)
print("ALL OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
