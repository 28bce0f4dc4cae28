"""the callers' host work on the device.
 * b200ldlt_assemble_augsys_device against the REFERENCE's own TripletHelper::FillValues on the CompoundSymMatrix that
   StdAugSystemSolver builds (tests/driver/libvecref.so -> unmodified libipopt.so): bit-identical value arrays;
 * b200ldlt_solve_refine_device against the host-buffer MultiSolve path + a numpy refinement loop."""
import ctypes as C

import numpy as np
import pytest

import vecref_api as R
from ipopt_b200 import B200Ldlt, SYMSOLVER_SUCCESS
from ipopt_b200.kkt import to_scipy

pytestmark = pytest.mark.gpu


def ref_fill(n_x, n_s, n_c, W, Wf, Jc, Jd, D, delta):
    L = R.lib()
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    L.vecref_augsys_fill.argtypes = ([C.c_int] * 3 + [C.c_int, ip, ip, dp, C.c_double] + [C.c_int, ip, ip, dp] * 2 +
                                     [dp, C.c_double] * 4 + [ip, ip, dp])
    cap = len(W[0]) + n_x + n_s + len(Jc[0]) + n_c + len(Jd[0]) + 2 * n_s
    irn, jcn, val = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap)

    def tri(T):
        i = np.ascontiguousarray(T[0] if len(T[0]) else [0], np.int32)
        j = np.ascontiguousarray(T[1] if len(T[1]) else [0], np.int32)
        v = np.ascontiguousarray(T[2] if len(T[2]) else [0.0], np.float64)
        return [len(T[0]), i.ctypes.data_as(ip), j.ctypes.data_as(ip), v.ctypes.data_as(dp)], (i, j, v)

    aw, kw = tri(W); ajc, kjc = tri(Jc); ajd, kjd = tri(Jd)
    dargs, keep = [], []
    for key in ("x", "s", "c", "d"):
        d = D[key]
        if d is None:
            dargs += [None, delta[key]]
        else:
            a = np.ascontiguousarray(d, np.float64); keep.append(a)
            dargs += [a.ctypes.data_as(dp), delta[key]]
    nnz = L.vecref_augsys_fill(n_x, n_s, n_c, *aw, Wf, *ajc, *ajd, *dargs, irn.ctypes.data_as(ip), jcn.ctypes.data_as(ip),
                               val.ctypes.data_as(dp))
    assert nnz == cap
    return irn, jcn, val


def random_blocks(rng, n_x, n_s, n_c):
    """A small NLP-like structure: W lower-triangular triplets (with the diagonal), J_c and J_d of full row rank."""
    wi, wj = [], []
    for i in range(n_x):
        wi.append(i + 1); wj.append(i + 1)
        for j in rng.choice(i, size=min(i, 2), replace=False) if i else []:
            wi.append(i + 1); wj.append(int(j) + 1)
    W = (np.array(wi), np.array(wj), rng.standard_normal(len(wi)))

    def jac(m):
        ii, jj = [], []
        for r in range(m):
            cols = rng.choice(n_x, size=min(n_x, 3), replace=False)
            ii += [r + 1] * len(cols); jj += [int(c) + 1 for c in cols]
        return (np.array(ii, int), np.array(jj, int), rng.standard_normal(len(ii)) + 0.5)
    return W, jac(n_c), jac(n_s)


@pytest.mark.parametrize("n_x,n_s,n_c,mode", [(40, 7, 11, 0), (300, 0, 120, 1), (500, 60, 200, 2), (1, 0, 0, 1)])
def test_assemble_matches_reference_fillvalues_and_refines(n_x, n_s, n_c, mode):
    import torch
    if not R.available():
        pytest.skip("tests/driver/libvecref.so not built")
    rng = np.random.default_rng(n_x)
    W, Jc, Jd = random_blocks(rng, n_x, n_s, n_c)
    D = {"x": np.exp(rng.uniform(-3, 6, n_x)), "s": np.exp(rng.uniform(-3, 6, n_s)),
         "c": None if mode == 0 else -np.exp(rng.uniform(-8, -2, n_c)), "d": None if mode != 2 else -np.exp(rng.uniform(-8, -2, n_s))}
    delta = {"x": [0.0, 1e-4, 0.0][mode], "s": [0.0, 1e-4, 0.0][mode], "c": [0.0, 1e-8, 0.0][mode], "d": [0.0, 1e-8, 1e-9][mode]}
    Wf = [1.0, 0.5, -1.0][mode]
    irn, jcn, val = ref_fill(n_x, n_s, n_c, W, Wf, Jc, Jd, D, delta)
    dim, nnz = n_x + 2 * n_s + n_c, len(val)
    s = B200Ldlt()
    assert s.InitializeStructure(dim, nnz, irn, jcn) == SYMSOLVER_SUCCESS
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float64)).cuda() if a is not None and len(a) else None
    t = {"W": dev(W[2]), "Jc": dev(Jc[2]), "Jd": dev(Jd[2]), **{k: dev(v) for k, v in D.items()}}
    torch.cuda.synchronize()
    p = lambda x: 0 if x is None else x.data_ptr()
    st = s.assemble_augsys_device(n_x, n_s, n_c, len(W[2]), len(Jc[2]), len(Jd[2]), W=p(t["W"]), W_factor=Wf, D_x=p(t["x"]),
                                  delta_x=delta["x"], D_s=p(t["s"]), delta_s=delta["s"], J_c=p(t["Jc"]), D_c=p(t["c"]),
                                  delta_c=delta["c"], J_d=p(t["Jd"]), D_d=p(t["d"]), delta_d=delta["d"])
    assert st == SYMSOLVER_SUCCESS, s.last_error()
    from ipopt_b200.sharded import _DevArr
    got = torch.as_tensor(_DevArr(s._L.b200ldlt_device_ptr(s._h, b"vals"), nnz, "<f8"), device="cuda").cpu().numpy()
    assert np.array_equal(got, val)            # bit-identical to the reference's FillValues
    # factor the assembled matrix (values never left the device) and solve with device-side refinement
    st, neg = s.refactor(False, 0)
    A = to_scipy(dim, irn, jcn, val)
    if st != SYMSOLVER_SUCCESS:
        pytest.skip("random block matrix numerically singular for the solver (status %d)" % st)
    b = rng.standard_normal(dim)
    db = torch.from_numpy(b.copy()).cuda(); torch.cuda.synchronize()
    st, steps, ratio = s.solve_refine_device(db.data_ptr(), min_steps=1, max_steps=10, tol=1e-12)
    assert st == SYMSOLVER_SUCCESS and 1 <= steps <= 10
    x = db.cpu().numpy()
    r = b - A @ x
    rr = np.abs(r).max() / (min(np.abs(x).max(), 1e6 * np.abs(b).max()) + np.abs(b).max())
    assert rr <= 1e-10 and abs(rr - ratio) <= 1e-3 * rr + 1e-15   # (at round-off level the two differ in the last digits)
    # host path on the same values: one solve + one numpy refinement step agrees
    xh = b.copy(); s.solve(xh)
    rh = b - A @ xh; s.solve(rh); xh += rh
    assert np.linalg.norm(x - xh) <= 1e-8 * np.linalg.norm(xh)
    s.close()
