"""Synthetic KKT (augmented-system) generators shaped like the reference's ScalableProblems.

They reproduce the *structure* Ipopt hands to the linear solver -- the triplet block order of
TripletHelper for the CompoundSymMatrix built in reference src/Algorithm/IpStdAugSystemSolver.cpp:309-468
((0,0): W entries then the Sigma_x+delta_x diagonal; (2,0): J_c; (2,2): the -delta_c diagonal delivered
as explicit entries) -- with synthetic Sigma values, so tests and bench run without the reference tree.
Sizes follow SURVEY.md section 8 (n_x, n_c, nnz formulas).  1-based triplets.
"""
import numpy as np


def mbndry_kkt(N, sigma_spread=0.0, delta_x=0.0, delta_c=0.0, seed=0, w_zero=False, alpha=0.01):
    """KKT pattern of MBndryCntrl1 (reference examples/ScalableProblems/MittelmannBndryCntrlDiri.cpp:60-82,
    395-433: (N+2)^2 grid values minus the 4 fixed corners, one 5-point-stencil PDE constraint per interior
    point {4,-1,-1,-1,-1}; Hessian diagonal h^2 on interior y and alpha*h on boundary controls)."""
    rng = np.random.default_rng(seed)
    M = N + 2
    h = 1.0 / (N + 1)
    idx = -np.ones((M, M), dtype=np.int64)
    k = 0
    for i in range(M):
        for j in range(M):
            if (i in (0, M - 1)) and (j in (0, M - 1)):
                continue
            idx[i, j] = k
            k += 1
    nx = k
    nc = N * N
    ii, jj = np.meshgrid(np.arange(1, N + 1), np.arange(1, N + 1), indexing="ij")
    ii, jj = ii.ravel(), jj.ravel()
    # W: interior diag then boundary diag (nnz_h_lag = N^2 + 4N)
    w_rows = [idx[ii, jj]]
    w_vals = [np.full(nc, h * h)]
    bd = np.concatenate([idx[1:N + 1, 0], idx[1:N + 1, M - 1], idx[0, 1:N + 1], idx[M - 1, 1:N + 1]])
    w_rows.append(bd)
    w_vals.append(np.full(4 * N, alpha * h))
    w_rows = np.concatenate(w_rows)
    w_vals = np.concatenate(w_vals)
    if w_zero:
        w_vals = np.zeros_like(w_vals)
    # Sigma_x + delta_x
    if sigma_spread > 0:
        sig = 10.0 ** rng.uniform(-sigma_spread, sigma_spread, nx)
    else:
        sig = np.ones(nx)
    dx_rows = np.arange(nx)
    dx_vals = sig + delta_x
    # J_c
    g = np.arange(nc)
    jr = np.repeat(g, 5)
    jc = np.stack([idx[ii, jj], idx[ii - 1, jj], idx[ii + 1, jj], idx[ii, jj - 1], idx[ii, jj + 1]], axis=1).ravel()
    jv = np.tile(np.array([4.0, -1.0, -1.0, -1.0, -1.0]), nc)
    irn = np.concatenate([w_rows, dx_rows, nx + jr, nx + g]) + 1
    jcn = np.concatenate([w_rows, dx_rows, jc, nx + g]) + 1
    val = np.concatenate([w_vals, dx_vals, jv, np.full(nc, -delta_c)])
    return nx + nc, irn.astype(np.int32), jcn.astype(np.int32), val.astype(np.float64), nc


def lukvle1_kkt(N, sigma_spread=0.0, delta_x=0.0, delta_c=0.0, seed=0, w_zero=False):
    """KKT pattern of LukVlE1 (reference examples/ScalableProblems/LuksanVlcek1.cpp:45-51,189-201,244-256:
    chained Rosenbrock, N-2 constraints each touching x_k, x_{k+1}, x_{k+2}; W has the diagonal and the
    (i,i+1) entries delivered in the UPPER triangle)."""
    rng = np.random.default_rng(seed)
    nx, nc = N, N - 2
    x = rng.uniform(-1.5, 1.5, N)
    wd_rows = np.arange(N)
    wd = 2.0 + 400.0 * rng.uniform(0.1, 2.0, N)
    wo_r = np.arange(N - 1)
    wo_c = wo_r + 1
    wo = -400.0 * x[:-1]
    if w_zero:
        wd = np.zeros_like(wd)
        wo = np.zeros_like(wo)
    sig = 10.0 ** rng.uniform(-sigma_spread, sigma_spread, nx) if sigma_spread > 0 else np.ones(nx)
    g = np.arange(nc)
    jr = np.repeat(g, 3)
    jc = np.stack([g, g + 1, g + 2], axis=1).ravel()
    jv = np.stack([3.0 * x[g + 1] ** 2 * np.ones(nc) + 0.1, 4.0 + 0.5 * np.cos(x[g]), -1.0 - 0.3 * np.sin(x[g + 2])], axis=1).ravel()
    irn = np.concatenate([wd_rows, wo_r, np.arange(nx), nx + jr, nx + g]) + 1
    jcn = np.concatenate([wd_rows, wo_c, np.arange(nx), jc, nx + g]) + 1
    val = np.concatenate([wd, wo, sig + delta_x, jv, np.full(nc, -delta_c)])
    return nx + nc, irn.astype(np.int32), jcn.astype(np.int32), val.astype(np.float64), nc


def random_kkt(nx, nc, density=0.02, seed=0, delta_c=0.0, spd_w=True):
    """Random sparse saddle-point matrix [[H, J^T],[J, -delta_c I]] with H SPD (expected inertia: nc negative)."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    B = sp.random(nx, nx, density=density, random_state=seed, format="coo")
    H = (B @ B.T).tocoo()
    H = sp.tril(H + sp.identity(nx) * (1.0 if spd_w else 0.0), format="coo")
    J = sp.random(nc, nx, density=max(density, 3.0 / nx), random_state=seed + 1, format="lil")
    for r in range(nc):  # guarantee full row rank with a shifted identity part
        J[r, (r * 7 + 3) % nx] += 2.0 + rng.uniform()
    J = J.tocoo()
    irn = np.concatenate([H.row, nx + J.row, nx + np.arange(nc)]) + 1
    jcn = np.concatenate([H.col, J.col, nx + np.arange(nc)]) + 1
    val = np.concatenate([H.data, J.data, np.full(nc, -delta_c)])
    return nx + nc, irn.astype(np.int32), jcn.astype(np.int32), val.astype(np.float64), nc


def to_scipy(dim, irn, jcn, val):
    """Full symmetric scipy CSC matrix from 1-based triplets of either triangle (duplicates summed)."""
    import scipy.sparse as sp
    i, j = irn.astype(np.int64) - 1, jcn.astype(np.int64) - 1
    off = i != j
    A = sp.coo_matrix((np.concatenate([val, val[off]]), (np.concatenate([i, j[off]]), np.concatenate([j, i[off]]))),
                      shape=(dim, dim))
    return A.tocsc()
