import os, sys, ctypes
os.environ["B200_SOLVE_TIMELINE"] = "1"
import numpy as np
sys.path.insert(0, ".")
from ipopt_b200 import B200Ldlt
from ipopt_b200.kkt import mbndry_kkt
N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
dim, irn, jcn, val, nc = mbndry_kkt(N, sigma_spread=3.0, seed=1)
s = B200Ldlt(verbose=1)
s.InitializeStructure(dim, len(irn), irn, jcn)
s.GetValuesArrayPtr()[:] = val
print(s.factor(True, nc))
b = np.random.default_rng(0).standard_normal(dim)
for _ in range(3):
    x = b.copy(); s.solve(x)
print("solve ms", s.info()["ms_solve_gpu"])
s._L.b200ldlt_dump_solve_timeline.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
print(s._L.b200ldlt_dump_solve_timeline(s._h, b"gpurun_out/solve_timeline.txt"))
