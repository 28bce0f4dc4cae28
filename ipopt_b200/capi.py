"""ctypes mirror of the reference's solver plugin interface on top of the C ABI.

Method names, argument meaning and status codes follow
``Ipopt::SparseSymLinearSolverInterface`` (reference
src/Algorithm/LinearSolvers/IpSparseSymLinearSolverInterface.hpp:98-256) and
``ESymSolverStatus`` (IpSymLinearSolver.hpp:19-33) so the parity tests read like calls the
reference's ``TSymLinearSolver`` makes (IpTSymLinearSolver.cpp:159-312).
There is NO CPU fallback: constructing :class:`B200Ldlt` without a usable CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

SYMSOLVER_SUCCESS, SYMSOLVER_SINGULAR, SYMSOLVER_WRONG_INERTIA, SYMSOLVER_CALL_AGAIN, SYMSOLVER_FATAL_ERROR = range(5)

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    return os.path.join(_HERE, "lib", "libb200ldlt.so")


class Options(C.Structure):
    _fields_ = [("device", C.c_int), ("stream", C.c_void_p), ("ordering", C.c_int), ("pair_saddle", C.c_int),
                ("leaf_k", C.c_int), ("relax_frac", C.c_double), ("scaling", C.c_int), ("pivtol", C.c_double),
                ("pivtolmax", C.c_double), ("tiny", C.c_double), ("smem_front_max", C.c_int),
                ("use_graph", C.c_int), ("verbose", C.c_int), ("tc_schur_min_r", C.c_int)]


class AugSys(C.Structure):
    _fields_ = [("n_x", C.c_int), ("n_s", C.c_int), ("n_c", C.c_int), ("n_d", C.c_int), ("nnz_w", C.c_int),
                ("nnz_jc", C.c_int), ("nnz_jd", C.c_int), ("W", C.c_void_p), ("W_factor", C.c_double),
                ("D_x", C.c_void_p), ("delta_x", C.c_double), ("D_s", C.c_void_p), ("delta_s", C.c_double),
                ("J_c", C.c_void_p), ("D_c", C.c_void_p), ("delta_c", C.c_double), ("J_d", C.c_void_p),
                ("D_d", C.c_void_p), ("delta_d", C.c_double)]


class Info(C.Structure):
    _fields_ = [("n", C.c_int), ("nnz_in", C.c_int64), ("nnz_unique", C.c_int64), ("nsupernodes", C.c_int),
                ("nlevels", C.c_int), ("max_front", C.c_int), ("max_pivots", C.c_int), ("n_saddle", C.c_int),
                ("n_pairs", C.c_int), ("nnz_L", C.c_int64), ("nnz_L_true", C.c_int64), ("L_bytes", C.c_int64),
                ("cb_bytes", C.c_int64), ("flops_panel", C.c_double), ("flops_schur", C.c_double),
                ("t_order_s", C.c_double), ("t_symbolic_s", C.c_double), ("num_neg", C.c_int),
                ("num_forced", C.c_int), ("num_tiny", C.c_int), ("num_growth", C.c_int), ("num_2x2", C.c_int),
                ("ms_factor_gpu", C.c_float), ("ms_solve_gpu", C.c_float), ("launches_factor", C.c_int),
                ("launches_solve", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTED = ["b200ldlt_default_options", "b200ldlt_create", "b200ldlt_destroy", "b200ldlt_last_error",
            "b200ldlt_analyse", "b200ldlt_values_ptr", "b200ldlt_factor", "b200ldlt_factor_device",
            "b200ldlt_solve", "b200ldlt_solve_device", "b200ldlt_num_neg", "b200ldlt_increase_quality",
            "b200ldlt_refactor", "b200ldlt_get_info", "b200ldlt_symbolic_array", "b200ldlt_analyse_now",
            "b200ldlt_residual", "b200ldlt_set_pivtol",
            "b200ldlt_assemble_augsys_device", "b200ldlt_solve_refine_device"]

_lib = None


def load_library():
    """Load libb200ldlt.so (raises if it has not been built: run ``python __graft_entry__.py build``)."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise RuntimeError("libb200ldlt.so is not built (%s); run `make -C ipopt_b200/csrc`" % p)
    L = C.CDLL(p)
    vp, ip, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)
    L.b200ldlt_default_options.argtypes = [C.POINTER(Options)]
    L.b200ldlt_default_options.restype = None
    L.b200ldlt_create.argtypes = [C.POINTER(Options)]
    L.b200ldlt_create.restype = vp
    L.b200ldlt_destroy.argtypes = [vp]
    L.b200ldlt_destroy.restype = None
    L.b200ldlt_last_error.argtypes = [vp]
    L.b200ldlt_last_error.restype = C.c_char_p
    L.b200ldlt_analyse.argtypes = [vp, C.c_int, C.c_int, ip, ip]
    L.b200ldlt_values_ptr.argtypes = [vp]
    L.b200ldlt_values_ptr.restype = dp
    L.b200ldlt_factor.argtypes = [vp, C.c_int, C.c_int, ip]
    L.b200ldlt_factor_device.argtypes = [vp, vp, C.c_int, C.c_int, ip]
    L.b200ldlt_refactor.argtypes = [vp, C.c_int, C.c_int, ip]
    L.b200ldlt_assemble_augsys_device.argtypes = [vp, C.POINTER(AugSys)]
    L.b200ldlt_solve_refine_device.argtypes = [vp, vp, C.c_int, C.c_int, C.c_double, ip, dp]
    L.b200ldlt_device_ptr.argtypes = [vp, C.c_char_p]
    L.b200ldlt_device_ptr.restype = vp
    L.b200ldlt_solve.argtypes = [vp, C.c_int, dp]
    L.b200ldlt_solve_device.argtypes = [vp, C.c_int, vp]
    L.b200ldlt_num_neg.argtypes = [vp]
    L.b200ldlt_increase_quality.argtypes = [vp]
    L.b200ldlt_get_info.argtypes = [vp, C.POINTER(Info)]
    L.b200ldlt_symbolic_array.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int64), C.c_int64]
    L.b200ldlt_symbolic_array.restype = C.c_int64
    L.b200ldlt_analyse_now.argtypes = [vp, dp]
    L.b200ldlt_residual.argtypes = [vp, dp, dp, dp, dp, dp]
    L.b200ldlt_symbolic_create.argtypes = [C.c_int, C.c_int, ip, ip, dp, C.c_int, C.c_int, C.c_int, C.c_double]
    L.b200ldlt_symbolic_create.restype = vp
    L.b200ldlt_symbolic_free.argtypes = [vp]
    L.b200ldlt_symbolic_free.restype = None
    L.b200ldlt_symbolic_get.argtypes = [vp, C.c_char_p, C.POINTER(C.c_longlong), C.c_longlong]
    L.b200ldlt_symbolic_get.restype = C.c_longlong
    L.b200ldlt_symbolic_shard.argtypes = [vp, C.c_int, C.POINTER(C.c_longlong), C.c_longlong]
    L.b200ldlt_symbolic_shard.restype = C.c_longlong
    _lib = L
    return L


def _iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class SymbolicAnalysis:
    """Host-only symbolic analysis (ordering, elimination tree, supernodes, maps). No GPU needed."""
    STATS = ["n", "nsn", "nlevels", "max_front", "max_k", "n_saddle", "n_pairs", "nnzL", "nnzL_true", "cb_total",
             "flops_panel", "flops_schur", "ms_order", "ms_symbolic"]

    def __init__(self, dim, irn, jcn, vals=None, ordering=0, pair_saddle=1, leaf_k=0, relax_frac=-1.0):
        L = load_library()
        irn = np.ascontiguousarray(irn, dtype=np.int32)
        jcn = np.ascontiguousarray(jcn, dtype=np.int32)
        vp = None
        if vals is not None:
            vals = np.ascontiguousarray(vals, dtype=np.float64)
            vp = _dptr(vals)
        self._L = L
        self._h = L.b200ldlt_symbolic_create(int(dim), int(len(irn)), _iptr(irn), _iptr(jcn), vp, ordering,
                                             pair_saddle, leaf_k, relax_frac)
        if not self._h:
            raise RuntimeError("symbolic analysis failed")

    def get(self, name):
        n = self._L.b200ldlt_symbolic_get(self._h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.zeros(max(n, 1), dtype=np.int64)
        self._L.b200ldlt_symbolic_get(self._h, name.encode(), out.ctypes.data_as(C.POINTER(C.c_longlong)), n)
        return out[:n]

    def stats(self):
        return dict(zip(self.STATS, self.get("stats").tolist()))

    def shard(self, world):
        """(owner per supernode [-1 = top part, factorised by rank 0], number of subtrees below the cut)."""
        nsn = len(self.get("sn_start")) - 1
        out = np.zeros(max(nsn, 1), dtype=np.int64)
        nsub = self._L.b200ldlt_symbolic_shard(self._h, int(world), out.ctypes.data_as(C.POINTER(C.c_longlong)), nsn)
        return out[:nsn], int(nsub)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.b200ldlt_symbolic_free(self._h)
            self._h = None


class B200Ldlt:
    """Mirror of Ipopt::SparseSymLinearSolverInterface on top of the C ABI (GPU required)."""

    Triplet_Format = 0  # EMatrixFormat, IpSparseSymLinearSolverInterface.hpp:101-114

    def __init__(self, **opts):
        L = load_library()
        self._L = L
        o = Options()
        L.b200ldlt_default_options(C.byref(o))
        for k, v in opts.items():
            if not hasattr(o, k):
                raise TypeError("unknown option %s" % k)
            setattr(o, k, v)
        self.options = o
        self._h = L.b200ldlt_create(C.byref(o))
        if not self._h:
            raise RuntimeError("b200ldlt_create failed: no usable CUDA device (this backend has no CPU fallback)")
        self.dim = 0
        self.nonzeros = 0
        self._negevals = -1
        self._ia = self._ja = None

    # -- the 8 virtuals -------------------------------------------------------------------------
    def InitializeStructure(self, dim, nonzeros, ia, ja):
        self._ia = np.ascontiguousarray(ia, dtype=np.int32)
        self._ja = np.ascontiguousarray(ja, dtype=np.int32)
        assert len(self._ia) == nonzeros and len(self._ja) == nonzeros
        self.dim, self.nonzeros = int(dim), int(nonzeros)
        return self._L.b200ldlt_analyse(self._h, self.dim, self.nonzeros, _iptr(self._ia), _iptr(self._ja))

    def GetValuesArrayPtr(self):
        p = self._L.b200ldlt_values_ptr(self._h)
        return np.ctypeslib.as_array(p, shape=(max(self.nonzeros, 1),))[:self.nonzeros]

    def MultiSolve(self, new_matrix, ia, ja, nrhs, rhs_vals, check_NegEVals, numberOfNegEVals):
        """rhs_vals: float64 array of dim*nrhs (column-major), overwritten with the solution."""
        if new_matrix:
            neg = C.c_int(-1)
            st = self._L.b200ldlt_factor(self._h, int(bool(check_NegEVals)), int(numberOfNegEVals), C.byref(neg))
            self._negevals = neg.value
            if st != SYMSOLVER_SUCCESS:
                return st
        assert rhs_vals.dtype == np.float64 and rhs_vals.flags["C_CONTIGUOUS"] or rhs_vals.flags["F_CONTIGUOUS"]
        return self._L.b200ldlt_solve(self._h, int(nrhs), _dptr(rhs_vals))

    def NumberOfNegEVals(self):
        return self._L.b200ldlt_num_neg(self._h)

    def IncreaseQuality(self):
        return bool(self._L.b200ldlt_increase_quality(self._h))

    def ProvidesInertia(self):
        return True

    def MatrixFormat(self):
        return self.Triplet_Format

    # -- extras -----------------------------------------------------------------------------------
    def factor(self, check=False, expected=0):
        neg = C.c_int(-1)
        st = self._L.b200ldlt_factor(self._h, int(check), int(expected), C.byref(neg))
        return st, neg.value

    def factor_device(self, d_vals_ptr, check=False, expected=0):
        neg = C.c_int(-1)
        st = self._L.b200ldlt_factor_device(self._h, C.c_void_p(d_vals_ptr), int(check), int(expected), C.byref(neg))
        return st, neg.value

    def refactor(self, check=False, expected=0):
        neg = C.c_int(-1)
        st = self._L.b200ldlt_refactor(self._h, int(check), int(expected), C.byref(neg))
        return st, neg.value

    def solve(self, rhs, nrhs=1):
        return self._L.b200ldlt_solve(self._h, int(nrhs), _dptr(rhs))

    def solve_device(self, d_rhs_ptr, nrhs=1):
        return self._L.b200ldlt_solve_device(self._h, int(nrhs), C.c_void_p(d_rhs_ptr))

    def assemble_augsys_device(self, n_x, n_s, n_c, nnz_w, nnz_jc, nnz_jd, W=0, W_factor=1.0, D_x=0, delta_x=0.0, D_s=0,
                               delta_s=0.0, J_c=0, D_c=0, delta_c=0.0, J_d=0, D_d=0, delta_d=0.0):
        """Device pointers (ints, 0 = absent) of the blocks of the augmented system -> the handle's device value array
        (b200ldlt_assemble_augsys_device); follow with refactor()."""
        a = AugSys(n_x, n_s, n_c, n_s, nnz_w, nnz_jc, nnz_jd, W or None, W_factor, D_x or None, delta_x, D_s or None, delta_s,
                   J_c or None, D_c or None, delta_c, J_d or None, D_d or None, delta_d)
        return self._L.b200ldlt_assemble_augsys_device(self._h, C.byref(a))

    def solve_refine_device(self, d_rhs_ptr, min_steps=1, max_steps=10, tol=1e-10):
        steps, ratio = C.c_int(0), C.c_double(0.0)
        st = self._L.b200ldlt_solve_refine_device(self._h, C.c_void_p(d_rhs_ptr), int(min_steps), int(max_steps), float(tol),
                                                  C.byref(steps), C.byref(ratio))
        return st, steps.value, ratio.value

    def analyse_now(self, vals=None):
        return self._L.b200ldlt_analyse_now(self._h, None if vals is None else _dptr(np.ascontiguousarray(vals)))

    def residual(self, x, b):
        r, xi, bi = C.c_double(), C.c_double(), C.c_double()
        st = self._L.b200ldlt_residual(self._h, _dptr(x), _dptr(b), C.byref(r), C.byref(xi), C.byref(bi))
        if st != 0:
            raise RuntimeError(self.last_error())
        return r.value, xi.value, bi.value

    def info(self):
        i = Info()
        self._L.b200ldlt_get_info(self._h, C.byref(i))
        return i.as_dict()

    def symbolic(self, name):
        n = self._L.b200ldlt_symbolic_array(self._h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.zeros(max(n, 1), dtype=np.int64)
        self._L.b200ldlt_symbolic_array(self._h, name.encode(), out.ctypes.data_as(C.POINTER(C.c_int64)), n)
        return out[:n]

    def last_error(self):
        return self._L.b200ldlt_last_error(self._h).decode()

    def close(self):
        if getattr(self, "_h", None):
            self._L.b200ldlt_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()
