"""Pin the CPU oracle: solutions against numpy dense solves, inertia against dense eigenvalues,
status codes against the reference's protocol (IpSymLinearSolver.hpp:19-33)."""
import numpy as np
import pytest

from ipopt_b200.kkt import lukvle1_kkt, mbndry_kkt, random_kkt, to_scipy
from oracle_api import OracleLdlt


def _solve_check(dim, irn, jcn, val, nc, tol=1e-9, **kw):
    o = OracleLdlt(**kw)
    assert o.InitializeStructure(dim, len(irn), irn, jcn) == 0
    o.GetValuesArrayPtr()[:] = val
    A = to_scipy(dim, irn, jcn, val).toarray()
    ev = np.linalg.eigvalsh(A)
    st, neg = o.factor(True, int((ev < 0).sum()))
    assert neg == int((ev < 0).sum())
    assert st == 0
    b = np.random.default_rng(1).standard_normal((dim,))
    x = b.copy()
    assert o.solve(x) == 0
    xr = np.linalg.solve(A, b)
    assert np.linalg.norm(x - xr) <= tol * np.linalg.norm(xr) * max(1.0, np.linalg.cond(A) * 1e-7)
    return o


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_saddle_point(seed):
    dim, irn, jcn, val, nc = random_kkt(60, 25, density=0.08, seed=seed)
    _solve_check(dim, irn, jcn, val, nc)


def test_mbndry_zero_hessian_first_matrix():
    # the first matrix Ipopt factors has W == 0, D_x = I, zero (2,2) block
    dim, irn, jcn, val, nc = mbndry_kkt(8, w_zero=True)
    _solve_check(dim, irn, jcn, val, nc)


def test_mbndry_wide_sigma():
    dim, irn, jcn, val, nc = mbndry_kkt(10, sigma_spread=6.0, seed=3)
    _solve_check(dim, irn, jcn, val, nc, tol=1e-7)


def test_lukvle_upper_triangle_entries():
    dim, irn, jcn, val, nc = lukvle1_kkt(120, sigma_spread=1.0, seed=2)
    assert np.any(irn < jcn)  # W(i,i+1) entries are delivered in the upper triangle
    _solve_check(dim, irn, jcn, val, nc)


def test_delayed_pivots_no_pairing_possible():
    # zero diagonal everywhere in the leading block: pivots must be delayed / 2x2 found late
    rng = np.random.default_rng(5)
    n = 40
    import scipy.sparse as sp
    B = sp.random(n, n, density=0.15, random_state=3).toarray()
    A = np.triu(B, 1)
    A = A + A.T
    A[np.arange(n - 5, n), np.arange(n - 5, n)] = rng.uniform(1, 2, 5)
    i, j = np.nonzero(np.tril(A))
    dim, irn, jcn, val = n, (i + 1).astype(np.int32), (j + 1).astype(np.int32), A[i, j]
    ev = np.linalg.eigvalsh(A)
    if np.abs(ev).min() < 1e-8:
        pytest.skip("random matrix numerically singular")
    o = OracleLdlt()
    o.InitializeStructure(dim, len(irn), irn, jcn)
    o.GetValuesArrayPtr()[:] = val
    st, neg = o.factor(False, 0)
    assert st == 0 and neg == int((ev < 0).sum())
    b = np.ones(n)
    x = b.copy()
    o.solve(x)
    assert np.allclose(A @ x, b, atol=1e-8 * max(1.0, np.abs(x).max()))


def test_singular_and_wrong_inertia_status():
    # exactly singular: two identical constraint rows
    dim, irn, jcn, val, nc = random_kkt(10, 3, density=0.3, seed=4)
    A = to_scipy(dim, irn, jcn, val).toarray()
    A[12, :] = A[11, :]
    A[:, 12] = A[:, 11]
    A[12, 12] = A[11, 11] = 0.0
    A[12, 11] = A[11, 12] = 0.0
    i, j = np.nonzero(np.tril(A))
    o = OracleLdlt()
    o.InitializeStructure(dim, len(i), (i + 1).astype(np.int32), (j + 1).astype(np.int32))
    o.GetValuesArrayPtr()[:] = A[i, j]
    st, neg = o.factor(True, nc)
    assert st == 1  # SYMSOLVER_SINGULAR
    # wrong inertia is reported without solving, negevals still set
    dim, irn, jcn, val, nc = random_kkt(30, 10, density=0.1, seed=6)
    o = OracleLdlt()
    o.InitializeStructure(dim, len(irn), irn, jcn)
    o.GetValuesArrayPtr()[:] = val
    st, neg = o.factor(True, nc + 1)
    assert st == 2 and neg == nc and o.NumberOfNegEVals() == nc


def test_increase_quality_schedule():
    o = OracleLdlt(pivtol=1e-6, pivtolmax=0.1)
    steps = 0
    while o.IncreaseQuality():
        steps += 1
    assert steps == 3  # 1e-6 -> 1e-3 -> 3.2e-2 -> 0.1 (capped), cf. IpMumpsSolverInterface.cpp:592-610
