"""One analyse + NF factorisations + NS solves of a synthetic MBndryCntrl1-shaped KKT (profiling target)."""
import sys, time, os
import numpy as np
sys.path.insert(0, ".")
import ipopt_b200.capi as _capi
if os.environ.get("B200_LIBDIR"):      # experiments: a differently configured build of the library
    _capi.lib_path = lambda: os.path.join(os.environ["B200_LIBDIR"], "libb200ldlt.so")
from ipopt_b200 import B200Ldlt
from ipopt_b200.kkt import mbndry_kkt
N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
NF = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dim, irn, jcn, val, nc = mbndry_kkt(N, sigma_spread=3.0, seed=1)
_, _, _, v0, _ = mbndry_kkt(N, w_zero=True)
import os
kw = {}
if os.environ.get("B200_LEAF_K"): kw["leaf_k"] = int(os.environ["B200_LEAF_K"])
if os.environ.get("B200_RELAX"): kw["relax_frac"] = float(os.environ["B200_RELAX"])
if os.environ.get("B200_SMEM_MAX"): kw["smem_front_max"] = int(os.environ["B200_SMEM_MAX"])
s = B200Ldlt(verbose=1, tc_schur_min_r=int(os.environ.get("B200_TC_MIN_R", "0")), **kw)
s.InitializeStructure(dim, len(irn), irn, jcn)
a = s.GetValuesArrayPtr()
a[:] = v0
print("first", s.factor(True, nc))
a[:] = val
b = np.random.default_rng(0).standard_normal(dim)
for it in range(NF):
    st, neg = s.factor(True, nc)
    x = b.copy(); s.solve(x)
    i = s.info()
    print("factor %.3f ms (%d launches)  solve %.3f ms (%d launches) st=%d neg=%d" % (i["ms_factor_gpu"], i["launches_factor"], i["ms_solve_gpu"], i["launches_solve"], st, neg))
r, xi, bi = s.residual(x, b)
print("resid", r / bi, "info", {k: v for k, v in i.items() if k in ("nnz_L", "flops_panel", "flops_schur", "nsupernodes", "nlevels", "max_front")})
