"""ctypes access to the CPU oracle (oracle/lib/liboracle_ldlt.so). TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def oracle_lib():
    global _lib
    if _lib is None:
        p = os.path.join(ROOT, "oracle", "lib", "liboracle_ldlt.so")
        if not os.path.exists(p):
            import subprocess
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "lib/liboracle_ldlt.so"])
        L = C.CDLL(p)
        L.oracle_ldlt_create.restype = C.c_void_p
        L.oracle_ldlt_create.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int]
        L.oracle_ldlt_destroy.argtypes = [C.c_void_p]
        L.oracle_ldlt_analyse.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_ldlt_values_ptr.restype = C.POINTER(C.c_double)
        L.oracle_ldlt_values_ptr.argtypes = [C.c_void_p]
        L.oracle_ldlt_factor.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.oracle_ldlt_solve.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.oracle_ldlt_num_neg.argtypes = [C.c_void_p]
        L.oracle_ldlt_increase_quality.argtypes = [C.c_void_p]
        L.oracle_ldlt_stats.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.oracle_ldlt_set_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


class OracleLdlt:
    def __init__(self, pivtol=1e-6, pivtolmax=0.1, scaling=1, verbose=0):
        self.L = oracle_lib()
        self.h = self.L.oracle_ldlt_create(pivtol, pivtolmax, scaling, verbose)

    def InitializeStructure(self, dim, nonzeros, ia, ja):
        self.ia = np.ascontiguousarray(ia, dtype=np.int32)
        self.ja = np.ascontiguousarray(ja, dtype=np.int32)
        self.dim, self.nonzeros = dim, nonzeros
        return self.L.oracle_ldlt_analyse(self.h, dim, nonzeros, self.ia.ctypes.data_as(C.POINTER(C.c_int)),
                                          self.ja.ctypes.data_as(C.POINTER(C.c_int)))

    def GetValuesArrayPtr(self):
        return np.ctypeslib.as_array(self.L.oracle_ldlt_values_ptr(self.h), shape=(max(self.nonzeros, 1),))[:self.nonzeros]

    def factor(self, check=False, expected=0):
        neg = C.c_int(-1)
        st = self.L.oracle_ldlt_factor(self.h, int(check), int(expected), C.byref(neg))
        return st, neg.value

    def solve(self, rhs, nrhs=1):
        return self.L.oracle_ldlt_solve(self.h, nrhs, rhs.ctypes.data_as(C.POINTER(C.c_double)))

    def MultiSolve(self, new_matrix, ia, ja, nrhs, rhs_vals, check_NegEVals, numberOfNegEVals):
        if new_matrix:
            st, _ = self.factor(check_NegEVals, numberOfNegEVals)
            if st != 0:
                return st
        return self.solve(rhs_vals, nrhs)

    def NumberOfNegEVals(self):
        return self.L.oracle_ldlt_num_neg(self.h)

    def IncreaseQuality(self):
        return bool(self.L.oracle_ldlt_increase_quality(self.h))

    def stats(self):
        out = (C.c_double * 8)()
        self.L.oracle_ldlt_stats(self.h, out)
        return dict(zip(["t_analyse", "t_factor", "t_solve", "nnzL", "flops", "max_front", "num_delayed", "num_2x2"], list(out)))

    def close(self):
        if self.h:
            self.L.oracle_ldlt_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()
