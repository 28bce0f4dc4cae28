"""Host logic: the saddle-row pairing / ordering must leave NO structurally or numerically deficient pivot block
(no forced or noise pivots) when the GPU's pivot rule is emulated on the CPU (tests/mf_pivot_emulator.py)."""
import numpy as np
import pytest

from ipopt_b200 import SymbolicAnalysis
from ipopt_b200.kkt import lukvle1_kkt, mbndry_kkt, random_kkt
from mf_pivot_emulator import emulate_factor


@pytest.mark.parametrize("gen,arg,kw", [
    (lukvle1_kkt, 300, dict(sigma_spread=2.0, seed=3)), (lukvle1_kkt, 3000, dict(sigma_spread=2.0, seed=3)),
    (lukvle1_kkt, 500, dict(w_zero=True)), (lukvle1_kkt, 4000, dict(w_zero=True)),
    (mbndry_kkt, 20, dict(sigma_spread=6.0, seed=1)), (mbndry_kkt, 16, dict(w_zero=True)),
])
def test_no_forced_pivots(built_lib, gen, arg, kw):
    dim, irn, jcn, val, nc = gen(arg, **kw)
    S = SymbolicAnalysis(dim, irn, jcn, val)
    tot, bad = emulate_factor(S, dim, irn, jcn, val)
    assert tot["forced"] == 0 and tot["tiny"] == 0, (tot, len(bad))
    assert tot["n1"] + 2 * tot["n2"] == dim


def test_unpartnered_saddle_rows_come_after_their_neighbours(built_lib):
    dim, irn, jcn, val, nc = lukvle1_kkt(3000, sigma_spread=2.0, seed=3)
    S = SymbolicAnalysis(dim, irn, jcn, val)
    st = S.stats()
    assert st["n_saddle"] == nc and 0 < nc - st["n_pairs"] < 0.02 * nc    # a few rows stay unpartnered by design
    perm = S.get("perm")
    ip = np.empty(dim, dtype=np.int64)
    ip[perm] = np.arange(dim)
    nx = dim - nc
    for c in range(nc):
        pos = ip[nx + c]
        prev = perm[pos - 1]
        partnered = prev < nx and prev in (c, c + 1, c + 2)
        if not partnered:
            assert all(ip[x] < pos for x in (c, c + 1, c + 2))
