"""Host-logic check of the recursive-doubling inverse of a front's pivot block L11 (ipopt_b200/csrc/b200ldlt.cu
`build_linv_plan`, kernels `k_linv_diag` / `k_linv_gemm<1,2>` in solve_dataflow.cu): the same work lists and tile
ranges, executed with numpy on 64x64 tiles, must give L11^-1 (zero-padded to K64), and the block GEMVs the solve
tasks perform with it must equal the triangular solves."""
import numpy as np
import pytest

B = 64


def linv_items(nkb):
    """(level) -> list of (ib, jb, m0, m1) for pass 1 and pass 2, exactly as build_linv_plan enumerates them."""
    levels = []
    bt = 1
    while bt < nkb:
        g1, g2 = [], []
        a0 = 0
        while a0 + bt < nkb:
            a1, b1 = a0 + bt, min(a0 + 2 * bt, nkb)
            for ib in range(a1, b1):
                for jb in range(a0, a1):
                    g1.append((ib, jb, jb, a1))
                    g2.append((ib, jb, a1, ib + 1))
            a0 += 2 * bt
        levels.append((g1, g2))
        bt *= 2
    return levels


def emulate(L11):
    k = L11.shape[0]
    nkb = (k + B - 1) // B
    K64 = nkb * B
    Lp = np.zeros((K64, K64))
    Lp[:k, :k] = L11
    Li = np.zeros((K64, K64))
    T = np.zeros((K64, K64))

    def tile(A, i, j):
        return A[i * B:(i + 1) * B, j * B:(j + 1) * B]

    for b in range(nkb):                                   # k_linv_diag
        nd = min(B, k - b * B)
        D = np.eye(B)
        D[:nd, :nd] = np.tril(tile(Lp, b, b)[:nd, :nd], -1) + np.eye(nd)
        X = np.tril(np.linalg.inv(D))   # (forward substitution on e_j: exact zeros above the diagonal)
        X[nd:, :] = 0.0
        X[:, nd:] = 0.0
        tile(Li, b, b)[:] = X
    for g1, g2 in linv_items(nkb):
        for ib, jb, m0, m1 in g1:                          # k_linv_gemm<1>: T = L21 * Inv11
            acc = np.zeros((B, B))
            for m in range(m0, m1):
                acc += tile(Lp, ib, m) @ tile(Li, m, jb)
            tile(T, ib, jb)[:] = acc
        for ib, jb, m0, m1 in g2:                          # k_linv_gemm<2>: Linv21 = -Inv22 * T
            acc = np.zeros((B, B))
            for m in range(m0, m1):
                acc += tile(Li, ib, m) @ tile(T, m, jb)
            tile(Li, ib, jb)[:] = -acc
    return Li


@pytest.mark.parametrize("k", [64, 65, 130, 257, 448, 700])
def test_recursive_doubling_inverse(k):
    rng = np.random.default_rng(k)
    L = np.tril(rng.standard_normal((k, k)) * (1.0 / np.sqrt(k)), -1) + np.eye(k)   # (moderate condition number)
    Li = emulate(L)
    assert np.allclose(Li[:k, :k] @ L, np.eye(k), atol=1e-9)
    assert np.all(Li[k:, :] == 0) and np.all(Li[:, k:] == 0)
    assert np.all(np.triu(Li, 1) == 0)
    # every lower tile is written by exactly one level
    nkb = (k + B - 1) // B
    seen = set()
    for g1, g2 in linv_items(nkb):
        for it in g2:
            assert (it[0], it[1]) not in seen
            seen.add((it[0], it[1]))
    assert len(seen) == nkb * (nkb - 1) // 2


def test_block_gemv_solves_match_substitution():
    k, r = 300, 150
    rng = np.random.default_rng(1)
    L11 = np.tril(rng.standard_normal((k, k)) * (1.0 / np.sqrt(k)), -1) + np.eye(k)
    L21 = rng.standard_normal((r, k)) * 0.2
    Li = emulate(L11)
    nkb = (k + B - 1) // B
    w = rng.standard_normal(k + r)
    # forward: pivot blocks y_b = sum_{c<=b} Linv[b,c] w_c, then contribution rows
    wp = np.zeros(nkb * B); wp[:k] = w[:k]
    y = np.zeros(nkb * B)
    for b in range(nkb):
        for c in range(b + 1):
            y[b * B:(b + 1) * B] += Li[b * B:(b + 1) * B, c * B:(c + 1) * B] @ wp[c * B:(c + 1) * B]
    assert np.allclose(y[:k], np.linalg.solve(L11, w[:k]))
    cb = w[k:] - L21 @ y[:k]
    assert np.allclose(cb, w[k:] - L21 @ np.linalg.solve(L11, w[:k]))
    # backward: t = z - L21^T x_r, x_b = sum_{c>=b} Linv[c,b]^T t_c
    z, xr = rng.standard_normal(k), rng.standard_normal(r)
    t = np.zeros(nkb * B); t[:k] = z - L21.T @ xr
    x = np.zeros(nkb * B)
    for b in range(nkb):
        for c in range(b, nkb):
            x[b * B:(b + 1) * B] += Li[c * B:(c + 1) * B, b * B:(b + 1) * B].T @ t[c * B:(c + 1) * B]
    assert np.allclose(x[:k], np.linalg.solve(L11.T, z - L21.T @ xr))
