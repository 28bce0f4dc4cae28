"""Host-side mirror of the reference's ``Ipopt::DenseVector`` / ``ExpansionMatrix`` / ``GenTMatrix`` / ``SymTMatrix``
interfaces on top of the b200vec C ABI (include/b200vec.h): same method names, argument meaning and "homogeneous"
state machine as reference src/LinAlg/IpDenseVector.hpp:17-360, IpExpansionMatrix.hpp, TMatrices/IpGenTMatrix.hpp,
TMatrices/IpSymTMatrix.hpp, so the parity tests read like calls the reference's algorithm code makes.

Values live in DEVICE memory (a torch CUDA tensor is used purely as the allocator); nothing here computes on the host.
There is NO CPU fallback: creating a :class:`VecContext` without a usable CUDA device raises.
"""
import ctypes as C

import numpy as np

from .capi import load_library


class _Vec(C.Structure):
    _fields_ = [("d", C.c_void_p), ("n", C.c_int), ("homogeneous", C.c_int), ("scalar", C.c_double)]


_sigs_done = False


def _lib():
    global _sigs_done
    L = load_library()
    if not _sigs_done:
        vp, pv, dp = C.c_void_p, C.POINTER(_Vec), C.POINTER(C.c_double)
        L.b200vec_create.argtypes = [C.c_int, vp]
        L.b200vec_create.restype = vp
        L.b200vec_destroy.argtypes = [vp]
        L.b200vec_destroy.restype = None
        L.b200vec_last_error.argtypes = [vp]
        L.b200vec_last_error.restype = C.c_char_p
        L.b200vec_sync.argtypes = [vp]
        L.b200vec_launches.argtypes = [vp]
        L.b200vec_launches.restype = C.c_int64
        for name in ("copy", "ew_divide", "ew_multiply", "ew_select", "ew_max", "ew_min"):
            getattr(L, "b200vec_" + name).argtypes = [vp, pv, pv]
        for name in ("scal", "set", "add_scalar"):
            getattr(L, "b200vec_" + name).argtypes = [vp, C.c_double, pv]
        L.b200vec_axpy.argtypes = [vp, C.c_double, pv, pv]
        L.b200vec_dot.argtypes = [vp, pv, pv, dp]
        for name in ("nrm2", "asum", "amax", "max", "min", "sum", "sumlogs"):
            getattr(L, "b200vec_" + name).argtypes = [vp, pv, dp]
        for name in ("ew_reciprocal", "ew_abs", "ew_sqrt", "ew_sgn"):
            getattr(L, "b200vec_" + name).argtypes = [vp, pv]
        L.b200vec_add_two_vectors.argtypes = [vp, C.c_double, pv, C.c_double, pv, C.c_double, pv]
        L.b200vec_frac_to_bound.argtypes = [vp, pv, pv, C.c_double, dp]
        L.b200vec_add_vector_quotient.argtypes = [vp, C.c_double, pv, pv, C.c_double, pv]
        L.b200vec_exp_mult.argtypes = [vp, C.c_int, C.c_int, vp, C.c_double, pv, C.c_double, pv]
        L.b200vec_exp_transmult.argtypes = [vp, C.c_int, C.c_int, vp, C.c_double, pv, C.c_double, pv]
        L.b200vec_exp_add_msinvz.argtypes = [vp, C.c_int, C.c_int, vp, C.c_double, pv, pv, pv]
        L.b200vec_exp_sinv_blrm_zmtdbr.argtypes = [vp, C.c_int, C.c_int, vp, C.c_double, pv, pv, pv, pv, pv]
        ip = C.POINTER(C.c_int)
        L.b200vec_tmat_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, ip, ip, C.c_int]
        L.b200vec_tmat_create.restype = vp
        L.b200vec_tmat_destroy.argtypes = [vp]
        L.b200vec_tmat_destroy.restype = None
        L.b200vec_tmat_mult.argtypes = [vp, vp, C.c_double, pv, C.c_double, pv]
        L.b200vec_tmat_transmult.argtypes = [vp, vp, C.c_double, pv, C.c_double, pv]
        _sigs_done = True
    return L


class VecContext:
    """One stream + reduction scratch (b200vec_create)."""

    def __init__(self, device=-1):
        import torch
        self._torch = torch
        self._L = _lib()
        self._h = self._L.b200vec_create(device, None)
        if not self._h:
            raise RuntimeError("b200vec_create failed: no usable CUDA device (there is no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device() if device < 0 else device)

    def check(self, rc):
        if rc != 0:
            raise RuntimeError("b200vec call failed: " + self._L.b200vec_last_error(self._h).decode())

    def launches(self):
        return self._L.b200vec_launches(self._h)

    def close(self):
        if self._h:
            self._L.b200vec_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


class DenseVector:
    """Mirror of Ipopt::DenseVector: a dense array OR a single scalar ("homogeneous")."""

    def __init__(self, ctx, n):
        self.ctx = ctx
        self._buf = ctx._torch.empty(max(n, 1), dtype=ctx._torch.float64, device=ctx.device)
        self._v = _Vec(self._buf.data_ptr(), n, 1, 0.0)

    # -- storage ------------------------------------------------------------------------------------
    def Dim(self):
        return self._v.n

    def IsHomogeneous(self):
        return bool(self._v.homogeneous)

    def Scalar(self):
        return self._v.scalar

    def SetValues(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert len(x) == self._v.n
        if self._v.n:
            self._buf[: self._v.n].copy_(self.ctx._torch.from_numpy(x))
            self.ctx._torch.cuda.current_stream().synchronize()   # the context's stream does not wait for torch's stream
        self._v.homogeneous = 0

    def ExpandedValues(self):
        if self._v.homogeneous:
            return np.full(self._v.n, self._v.scalar)
        self.ctx.check(self.ctx._L.b200vec_sync(self.ctx._h))
        return self._buf[: self._v.n].cpu().numpy()

    def _p(self):
        return C.byref(self._v)

    # -- the Vector interface -----------------------------------------------------------------------
    def Copy(self, x): self.ctx.check(self.ctx._L.b200vec_copy(self.ctx._h, x._p(), self._p()))
    def Scal(self, alpha): self.ctx.check(self.ctx._L.b200vec_scal(self.ctx._h, alpha, self._p()))
    def Set(self, alpha): self.ctx.check(self.ctx._L.b200vec_set(self.ctx._h, alpha, self._p()))
    def AddScalar(self, s): self.ctx.check(self.ctx._L.b200vec_add_scalar(self.ctx._h, s, self._p()))
    def Axpy(self, alpha, x): self.ctx.check(self.ctx._L.b200vec_axpy(self.ctx._h, alpha, x._p(), self._p()))

    def _red(self, name, *args):
        out = C.c_double(0.0)
        self.ctx.check(getattr(self.ctx._L, "b200vec_" + name)(self.ctx._h, *args, C.byref(out)))
        return out.value

    def Dot(self, x): return self._red("dot", x._p(), self._p())
    def Nrm2(self): return self._red("nrm2", self._p())
    def Asum(self): return self._red("asum", self._p())
    def Amax(self): return self._red("amax", self._p())
    def Max(self): return self._red("max", self._p())
    def Min(self): return self._red("min", self._p())
    def Sum(self): return self._red("sum", self._p())
    def SumLogs(self): return self._red("sumlogs", self._p())
    def ElementWiseDivide(self, x): self.ctx.check(self.ctx._L.b200vec_ew_divide(self.ctx._h, x._p(), self._p()))
    def ElementWiseMultiply(self, x): self.ctx.check(self.ctx._L.b200vec_ew_multiply(self.ctx._h, x._p(), self._p()))
    def ElementWiseSelect(self, x): self.ctx.check(self.ctx._L.b200vec_ew_select(self.ctx._h, x._p(), self._p()))
    def ElementWiseMax(self, x): self.ctx.check(self.ctx._L.b200vec_ew_max(self.ctx._h, x._p(), self._p()))
    def ElementWiseMin(self, x): self.ctx.check(self.ctx._L.b200vec_ew_min(self.ctx._h, x._p(), self._p()))
    def ElementWiseReciprocal(self): self.ctx.check(self.ctx._L.b200vec_ew_reciprocal(self.ctx._h, self._p()))
    def ElementWiseAbs(self): self.ctx.check(self.ctx._L.b200vec_ew_abs(self.ctx._h, self._p()))
    def ElementWiseSqrt(self): self.ctx.check(self.ctx._L.b200vec_ew_sqrt(self.ctx._h, self._p()))
    def ElementWiseSgn(self): self.ctx.check(self.ctx._L.b200vec_ew_sgn(self.ctx._h, self._p()))

    def AddTwoVectors(self, a, v1, b, v2, c):
        self.ctx.check(self.ctx._L.b200vec_add_two_vectors(self.ctx._h, a, v1._p(), b, v2._p(), c, self._p()))

    def FracToBound(self, delta, tau):
        return self._red("frac_to_bound", self._p(), delta._p(), C.c_double(tau))

    def AddVectorQuotient(self, a, z, s, c):
        self.ctx.check(self.ctx._L.b200vec_add_vector_quotient(self.ctx._h, a, z._p(), s._p(), c, self._p()))


class ExpansionMatrix:
    """Mirror of Ipopt::ExpansionMatrix (NRows x NCols, ExpandedPosIndices 0-based)."""

    def __init__(self, ctx, nrows, ncols, exp_pos):
        self.ctx, self.nrows, self.ncols = ctx, nrows, ncols
        p = np.ascontiguousarray(exp_pos, dtype=np.int32)
        assert len(p) == ncols
        self._pos = ctx._torch.from_numpy(p if ncols else np.zeros(1, np.int32)).to(ctx.device)
        ctx._torch.cuda.current_stream().synchronize()

    def _pp(self):
        return C.c_void_p(self._pos.data_ptr())

    def MultVector(self, alpha, x, beta, y):
        self.ctx.check(self.ctx._L.b200vec_exp_mult(self.ctx._h, self.nrows, self.ncols, self._pp(), alpha, x._p(), beta, y._p()))

    def TransMultVector(self, alpha, x, beta, y):
        self.ctx.check(self.ctx._L.b200vec_exp_transmult(self.ctx._h, self.nrows, self.ncols, self._pp(), alpha, x._p(), beta, y._p()))

    def AddMSinvZ(self, alpha, S, Z, X):
        self.ctx.check(self.ctx._L.b200vec_exp_add_msinvz(self.ctx._h, self.nrows, self.ncols, self._pp(), alpha, S._p(), Z._p(), X._p()))

    def SinvBlrmZMTdBr(self, alpha, S, R, Z, D, X):
        self.ctx.check(self.ctx._L.b200vec_exp_sinv_blrm_zmtdbr(self.ctx._h, self.nrows, self.ncols, self._pp(), alpha, S._p(),
                                                                 R._p(), Z._p(), D._p(), X._p()))


class TripletMatrix:
    """Mirror of Ipopt::GenTMatrix (symmetric=False) / SymTMatrix (symmetric=True): 1-based triplets, SetValues, MultVector."""

    def __init__(self, ctx, nrows, ncols, irow, jcol, symmetric=False):
        self.ctx = ctx
        ir = np.ascontiguousarray(irow, dtype=np.int32)
        jc = np.ascontiguousarray(jcol, dtype=np.int32)
        self.nnz = len(ir)
        ip = C.POINTER(C.c_int)
        self._h = ctx._L.b200vec_tmat_create(ctx._h, nrows, ncols, self.nnz, ir.ctypes.data_as(ip), jc.ctypes.data_as(ip),
                                              1 if symmetric else 0)
        if not self._h:
            raise RuntimeError("b200vec_tmat_create failed")
        self._vals = ctx._torch.zeros(max(self.nnz, 1), dtype=ctx._torch.float64, device=ctx.device)

    def SetValues(self, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        assert len(v) == self.nnz
        if self.nnz:
            self._vals[: self.nnz].copy_(self.ctx._torch.from_numpy(v))
            self.ctx._torch.cuda.current_stream().synchronize()

    def MultVector(self, alpha, x, beta, y):
        self.ctx.check(self.ctx._L.b200vec_tmat_mult(self._h, C.c_void_p(self._vals.data_ptr()), alpha, x._p(), beta, y._p()))

    def TransMultVector(self, alpha, x, beta, y):
        self.ctx.check(self.ctx._L.b200vec_tmat_transmult(self._h, C.c_void_p(self._vals.data_ptr()), alpha, x._p(), beta, y._p()))

    def close(self):
        if self._h:
            self.ctx._L.b200vec_tmat_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()
