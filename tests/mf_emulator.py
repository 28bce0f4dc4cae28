"""numpy emulation of the multifrontal factorisation driven ONLY by the symbolic arrays the C++ analysis
exports (perm, supernodes, row structures, assembly map u_dst64/t2u, relative indices).  No pivoting: valid
for quasi-definite matrices.  It checks the maps the CUDA kernels rely on, on the CPU."""
import numpy as np


def emulate_solve(S, dim, irn, jcn, val, b):
    perm = S.get("perm")
    sn_start = S.get("sn_start")
    rows_ptr = S.get("rows_ptr")
    rows = S.get("rows")
    rel = S.get("rel")
    parent = S.get("sn_parent")
    uent_ptr = S.get("uent_ptr")
    u_dst64 = S.get("u_dst64")
    t2u = S.get("t2u")
    nsn = len(sn_start) - 1
    nu = len(u_dst64)
    uval = np.zeros(nu)
    np.add.at(uval, t2u, val)
    Ls, Ds, cbs = [None] * nsn, [None] * nsn, [None] * nsn
    children = [[] for _ in range(nsn)]
    for s in range(nsn):
        if parent[s] >= 0:
            children[parent[s]].append(s)
    for s in range(nsn):
        k = sn_start[s + 1] - sn_start[s]
        r = rows_ptr[s + 1] - rows_ptr[s]
        f = k + r
        F = np.zeros((f, f))
        P = np.zeros(f * k)
        P[u_dst64[uent_ptr[s]:uent_ptr[s + 1]]] = uval[uent_ptr[s]:uent_ptr[s + 1]]
        F[:, :k] = P.reshape((k, f)).T
        F = np.tril(F) + np.tril(F, -1).T
        for c in children[s]:
            rl = rel[rows_ptr[c]:rows_ptr[c + 1]]
            F[np.ix_(rl, rl)] += cbs[c]
            cbs[c] = None
        L = np.eye(f)[:, :k].copy()
        D = np.zeros(k)
        for j in range(k):
            D[j] = F[j, j]
            L[j + 1:, j] = F[j + 1:, j] / D[j]
            F[j + 1:, j + 1:] -= np.outer(L[j + 1:, j], F[j + 1:, j])
        Ls[s], Ds[s], cbs[s] = L, D, F[k:, k:].copy()
    x = b[perm].astype(float).copy()
    for s in range(nsn):
        a, e = sn_start[s], sn_start[s + 1]
        k = e - a
        rw = rows[rows_ptr[s]:rows_ptr[s + 1]]
        y = np.linalg.solve(Ls[s][:k, :k], x[a:e])
        x[a:e] = y
        x[rw] -= Ls[s][k:, :] @ y
    for s in range(nsn):
        a, e = sn_start[s], sn_start[s + 1]
        x[a:e] /= Ds[s]
    for s in range(nsn - 1, -1, -1):
        a, e = sn_start[s], sn_start[s + 1]
        k = e - a
        rw = rows[rows_ptr[s]:rows_ptr[s + 1]]
        v = x[a:e] - Ls[s][k:, :].T @ x[rw]
        x[a:e] = np.linalg.solve(Ls[s][:k, :k].T, v)
    out = np.zeros(dim)
    out[perm] = x
    return out
