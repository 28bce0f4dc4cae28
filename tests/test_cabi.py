"""C-ABI checks that need no GPU: the library loads, exports every symbol include/b200ldlt.h declares,
and refuses loudly (no CPU fallback) when no CUDA device is usable."""
import ctypes
import os
import re

import pytest

from ipopt_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header="b200ldlt.h", prefix="b200ldlt_"):
    hdr = open(os.path.join(ROOT, "include", header)).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(%s[a-z_0-9]+)\s*\(" % prefix, hdr)))


def test_header_symbols_are_exported(built_lib):
    lib = ctypes.CDLL(built_lib)
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    assert set(capi.EXPORTED) <= set(names)


def test_vector_header_symbols_are_exported(built_lib):
    """include/b200vec.h: every declared entry point is exported; without a GPU the context
    constructor refuses (no CPU path)."""
    lib = ctypes.CDLL(built_lib)
    names = _declared_symbols("b200vec.h", "b200vec_")
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
    import torch
    if not torch.cuda.is_available():
        from ipopt_b200.vec import VecContext
        with pytest.raises(RuntimeError, match="no usable CUDA device"):
            VecContext()


def test_status_codes_match_reference_enum():
    # ESymSolverStatus order, reference src/Algorithm/LinearSolvers/IpSymLinearSolver.hpp:19-33
    assert (capi.SYMSOLVER_SUCCESS, capi.SYMSOLVER_SINGULAR, capi.SYMSOLVER_WRONG_INERTIA,
            capi.SYMSOLVER_CALL_AGAIN, capi.SYMSOLVER_FATAL_ERROR) == (0, 1, 2, 3, 4)
    hdr = open(os.path.join(ROOT, "include", "b200ldlt.h")).read()
    for name, v in [("SUCCESS", 0), ("SINGULAR", 1), ("WRONG_INERTIA", 2), ("CALL_AGAIN", 3), ("FATAL_ERROR", 4)]:
        assert re.search(r"B200LDLT_%s\s*=\s*%d" % (name, v), hdr)


def test_no_cpu_fallback_without_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        capi.B200Ldlt()


def test_options_struct_layout_roundtrip(built_lib):
    L = capi.load_library()
    o = capi.Options()
    L.b200ldlt_default_options(ctypes.byref(o))
    assert o.device == -1 and o.pair_saddle == 1 and o.leaf_k == 32
    assert o.pivtol == 1e-8 and o.pivtolmax == 1e-4 and o.smem_front_max == 96 and o.scaling == 2 and o.use_graph == 1


def test_product_does_not_link_the_oracle(built_lib):
    import subprocess
    out = subprocess.run(["ldd", built_lib], capture_output=True, text=True).stdout
    assert "oracle" not in out
    for src in os.listdir(os.path.join(ROOT, "ipopt_b200", "csrc")) + os.listdir(os.path.join(ROOT, "ipopt_b200", "plugin")):
        p = os.path.join(ROOT, "ipopt_b200", "csrc", src)
        if not os.path.exists(p):
            p = os.path.join(ROOT, "ipopt_b200", "plugin", src)
        if os.path.isfile(p):
            assert "oracle" not in open(p).read().lower() or src == "B200LdltSolverInterface.hpp", src
