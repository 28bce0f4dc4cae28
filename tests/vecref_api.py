"""ctypes access to tests/driver/libvecref.so: the REFERENCE's own DenseVector / ExpansionMatrix / GenTMatrix /
SymTMatrix (unmodified libipopt.so of oracle/_ref) behind a C ABI.  TEST INFRASTRUCTURE ONLY (the checker of b200vec)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def available():
    return os.path.exists(os.path.join(ROOT, "tests", "driver", "libvecref.so"))


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(os.path.join(ROOT, "tests", "driver", "libvecref.so"))
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.vecref_op.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_double, C.c_double, dp, C.c_int, C.c_double, dp, C.c_int,
                                C.c_double, dp, ip, dp, dp]
        L.vecref_expansion.argtypes = [C.c_int, C.c_int, C.c_int, ip, C.c_double, C.c_double] + [dp, C.c_int, C.c_double] * 4 + [dp, ip, dp]
        L.vecref_tmat.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, dp, C.c_double, C.c_double, dp, C.c_int,
                                  C.c_double, dp, ip, dp]
        _lib = L
    return _lib


class HostVec:
    """(values, homogeneous, scalar) triple as the reference's DenseVector holds it."""

    def __init__(self, n, values=None, scalar=None):
        self.n = n
        self.d = np.zeros(max(n, 1))
        self.h = 1
        self.s = 0.0
        if values is not None:
            self.d[:n] = values
            self.h = 0
        elif scalar is not None:
            self.s = float(scalar)

    def args(self):
        return self.d.ctypes.data_as(C.POINTER(C.c_double)), self.h, self.s

    def expanded(self):
        return np.full(self.n, self.s) if self.h else self.d[: self.n].copy()


def _out_args(y):
    hy, sy = C.c_int(y.h), C.c_double(y.s)
    return hy, sy, (y.d.ctypes.data_as(C.POINTER(C.c_double)), C.byref(hy), C.byref(sy))


def op(name, y, x1=None, x2=None, a=0.0, b=0.0, c=0.0):
    n = y.n
    x1 = x1 or HostVec(n, scalar=0.0)
    x2 = x2 or HostVec(n, scalar=0.0)
    hy, sy, oa = _out_args(y)
    out = C.c_double(0.0)
    rc = lib().vecref_op(name.encode(), n, a, b, c, *x1.args(), *x2.args(), *oa, C.byref(out))
    assert rc == 0, name
    y.h, y.s = hy.value, sy.value
    return out.value


def expansion(which, nrows, ncols, pos, alpha, beta, y, x1=None, x2=None, x3=None, x4=None):
    z = HostVec(0, scalar=0.0)
    p = np.ascontiguousarray(pos if ncols else np.zeros(1), dtype=np.int32)
    hy, sy, oa = _out_args(y)
    vs = [v or z for v in (x1, x2, x3, x4)]
    rc = lib().vecref_expansion(which, nrows, ncols, p.ctypes.data_as(C.POINTER(C.c_int)), alpha, beta,
                                *vs[0].args(), *vs[1].args(), *vs[2].args(), *vs[3].args(), *oa)
    assert rc == 0
    y.h, y.s = hy.value, sy.value


def tmat(symmetric, trans, nrows, ncols, irow, jcol, values, alpha, beta, x, y):
    ir = np.ascontiguousarray(irow if len(irow) else np.zeros(1), dtype=np.int32)
    jc = np.ascontiguousarray(jcol if len(jcol) else np.zeros(1), dtype=np.int32)
    v = np.ascontiguousarray(values if len(values) else np.zeros(1), dtype=np.float64)
    hy, sy, oa = _out_args(y)
    ip = C.POINTER(C.c_int)
    rc = lib().vecref_tmat(int(symmetric), int(trans), nrows, ncols, len(irow), ir.ctypes.data_as(ip), jc.ctypes.data_as(ip),
                           v.ctypes.data_as(C.POINTER(C.c_double)), alpha, beta, *x.args(), *oa)
    assert rc == 0
    y.h, y.s = hy.value, sy.value
