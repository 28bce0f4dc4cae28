import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from ipopt_b200 import B200Ldlt
from ipopt_b200.kkt import lukvle1_kkt, to_scipy
from oracle_api import OracleLdlt
for N, kw in [(3000, dict(sigma_spread=2.0, seed=3)), (500, dict(w_zero=True)), (300, dict(sigma_spread=2.0, seed=3))]:
    dim, irn, jcn, val, nc = lukvle1_kkt(N, **kw)
    o = OracleLdlt(); o.InitializeStructure(dim, len(irn), irn, jcn); o.GetValuesArrayPtr()[:] = val
    st, nego = o.factor(False, 0); print("oracle", N, st, nego, nc, o.stats())
    s = B200Ldlt(verbose=2); s.InitializeStructure(dim, len(irn), irn, jcn); s.GetValuesArrayPtr()[:] = val
    st, neg = s.factor(True, nego); print("gpu", N, st, neg, s.info())
    if dim < 3000:
        ev = np.linalg.eigvalsh(to_scipy(dim, irn, jcn, val).toarray()); print("eig neg", (ev < 0).sum(), "min|ev|", np.abs(ev).min())
