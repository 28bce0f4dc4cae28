"""Host logic: symbolic analysis (ordering, supernodes, assembly / extend-add maps) checked on the CPU."""
import numpy as np
import pytest

from ipopt_b200 import SymbolicAnalysis
from ipopt_b200.kkt import lukvle1_kkt, mbndry_kkt, random_kkt, to_scipy

from mf_emulator import emulate_solve


def _check_structure(S, dim):
    perm = S.get("perm")
    assert sorted(perm.tolist()) == list(range(dim))
    sn_start, parent = S.get("sn_start"), S.get("sn_parent")
    rows_ptr, rows, rel = S.get("rows_ptr"), S.get("rows"), S.get("rel")
    nsn = len(sn_start) - 1
    assert sn_start[0] == 0 and sn_start[-1] == dim and np.all(np.diff(sn_start) > 0)
    for s in range(nsn):
        rw = rows[rows_ptr[s]:rows_ptr[s + 1]]
        assert np.all(np.diff(rw) > 0) and (len(rw) == 0 or rw[0] >= sn_start[s + 1])
        p = parent[s]
        if p < 0:
            assert len(rw) == 0
            continue
        assert p > s
        kp = sn_start[p + 1] - sn_start[p]
        prow = np.concatenate([np.arange(sn_start[p], sn_start[p + 1]), rows[rows_ptr[p]:rows_ptr[p + 1]]])
        rl = rel[rows_ptr[s]:rows_ptr[s + 1]]
        assert np.array_equal(prow[rl], rw)  # relative indices address the same global rows in the parent
        assert np.all(np.diff(rl) > 0)
        assert rw[0] < sn_start[p + 1]      # first row of a child lies in the parent's pivot block
        del kp


@pytest.mark.parametrize("gen,arg", [(mbndry_kkt, 6), (mbndry_kkt, 17), (lukvle1_kkt, 50), (lukvle1_kkt, 300)])
def test_maps_reproduce_dense_solution(built_lib, gen, arg):
    dim, irn, jcn, val, nc = gen(arg, sigma_spread=1.0, delta_c=1e-2, seed=3)  # quasi-definite => no pivoting needed
    val0 = val.copy()
    val0[-nc:] = 0.0  # analysis sees the zero (2,2) block like Ipopt's first matrix -> pairing active
    S = SymbolicAnalysis(dim, irn, jcn, val0)
    st = S.stats()
    assert st["n_pairs"] > 0
    _check_structure(S, dim)
    A = to_scipy(dim, irn, jcn, val).toarray()
    b = np.random.default_rng(0).standard_normal(dim)
    x = emulate_solve(S, dim, irn, jcn, val, b)
    xr = np.linalg.solve(A, b)
    assert np.linalg.norm(x - xr) <= 1e-8 * np.linalg.norm(xr)


def test_duplicates_and_upper_triangle(built_lib):
    dim, irn, jcn, val, nc = lukvle1_kkt(40, delta_c=0.5, seed=1)   # has upper-triangle W entries
    # add explicit duplicates and flip some entries to the other triangle
    irn2 = np.concatenate([irn, irn[:25]])
    jcn2 = np.concatenate([jcn, jcn[:25]])
    val2 = np.concatenate([val * 1.0, val[:25] * 0.5])
    irn2[5:15], jcn2[5:15] = jcn2[5:15].copy(), irn2[5:15].copy()
    S = SymbolicAnalysis(dim, irn2, jcn2, val2, pair_saddle=0)
    _check_structure(S, dim)
    A = to_scipy(dim, irn2, jcn2, val2).toarray()
    b = np.arange(1.0, dim + 1)
    x = emulate_solve(S, dim, irn2, jcn2, val2, b)
    assert np.allclose(A @ x, b, rtol=0, atol=1e-8 * np.abs(b).max())


def test_natural_order_small_dense_front(built_lib):
    dim, irn, jcn, val, nc = random_kkt(12, 5, density=0.3, seed=2, delta_c=1.0)
    S = SymbolicAnalysis(dim, irn, jcn, val)
    st = S.stats()
    assert st["nsn"] == 1 and st["max_front"] == dim   # n <= dense_n: a single dense front
    A = to_scipy(dim, irn, jcn, val).toarray()
    b = np.ones(dim)
    x = emulate_solve(S, dim, irn, jcn, val, b)
    assert np.allclose(A @ x, b, atol=1e-9)


def test_pairs_are_adjacent_and_share_a_supernode(built_lib):
    dim, irn, jcn, val, nc = mbndry_kkt(12, w_zero=True)
    S = SymbolicAnalysis(dim, irn, jcn, val)
    st = S.stats()
    assert st["n_saddle"] == nc and st["n_pairs"] == nc
    perm, sn_start = S.get("perm"), S.get("sn_start")
    iperm = np.empty(dim, dtype=np.int64)
    iperm[perm] = np.arange(dim)
    sn_of = np.searchsorted(sn_start, np.arange(dim), side="right") - 1
    nx = dim - nc
    pos_c = iperm[nx:]
    # the predecessor of every saddle row in the elimination order is a primal variable in the same front
    prev = perm[pos_c - 1]
    assert np.all(prev < nx)
    assert np.all(sn_of[pos_c] == sn_of[pos_c - 1])


def test_stats_match_survey_sizes(built_lib):
    # MBndryCntrl1 N=30 -> dim 1920, 7440 triplets (also seen by the probe run)
    dim, irn, jcn, val, nc = mbndry_kkt(30)
    assert dim == 1920 and len(irn) == 7440
    dim, irn, jcn, val, nc = lukvle1_kkt(1000)
    assert dim == 1998 and len(irn) == 6991


@pytest.mark.parametrize("gen,arg,kw", [(mbndry_kkt, 40, dict(w_zero=True)), (lukvle1_kkt, 5000, dict(w_zero=True)),
                                        (random_kkt, (700, 250), dict(density=0.02, seed=4)), (mbndry_kkt, 9, dict(sigma_spread=2.0, seed=1))])
def test_column_counts_match_the_explicit_structures(built_lib, gen, arg, kw, monkeypatch):
    """The skeleton-graph column counts of the analysis (symbolic.cpp step 6) against the explicit column structures built
    one column at a time (B200_SYMBOLIC_CHECK makes the analysis fail on any difference)."""
    monkeypatch.setenv("B200_SYMBOLIC_CHECK", "1")
    dim, irn, jcn, val, nc = gen(*arg, **kw) if isinstance(arg, tuple) else gen(arg, **kw)
    S = SymbolicAnalysis(dim, irn, jcn, val)
    assert S.stats()["nnzL_true"] > 0
    _check_structure(S, dim)
