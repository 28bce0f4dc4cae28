"""Generate the committed golden fixtures by running the REFERENCE's own interior-point loop
(oracle/_ref/lib/libipopt.so, built from /root/reference) on the reference's own TNLPs with the CPU oracle
as linear solver.  Needs /root/reference (driver build); the outputs travel as small fixtures.

  python tests/golden/make_goldens.py

Writes, per case: <case>_final.npz (final x, z_L, z_U, lambda, objective, iterations, call counts) and
<case>_kkt.npz (KKT triplets + rhs + solution + inertia of selected factorisations, captured at the
SparseSymLinearSolverInterface boundary).
"""
import json
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DRIVER = os.path.join(ROOT, "tests", "driver", "ipopt_driver")

CASES = [
    ("hs071", 0, list(range(12))),
    ("LukVlE1", 1000, [0, 1, 3, 6]),
    ("MBndryCntrl1", 30, [0, 1, 7, 14]),
    ("MDistCntrl3a", 25, [0, 1, 9, 17]),
]


def read_final(path):
    with open(path, "rb") as f:
        n, m = struct.unpack("ii", f.read(8))
        obj, = struct.unpack("d", f.read(8))
        x = np.frombuffer(f.read(8 * n), dtype=np.float64)
        zl = np.frombuffer(f.read(8 * n), dtype=np.float64)
        zu = np.frombuffer(f.read(8 * n), dtype=np.float64)
        lam = np.frombuffer(f.read(8 * m), dtype=np.float64)
    return dict(obj=obj, x=x, z_L=zl, z_U=zu, lam=lam)


def read_dump(path):
    with open(path, "rb") as f:
        dim, nnz, nrhs, neg = struct.unpack("iiii", f.read(16))
        irn = np.frombuffer(f.read(4 * nnz), dtype=np.int32)
        jcn = np.frombuffer(f.read(4 * nnz), dtype=np.int32)
        val = np.frombuffer(f.read(8 * nnz), dtype=np.float64)
        rhs = np.frombuffer(f.read(8 * dim * nrhs), dtype=np.float64)
        sol = np.frombuffer(f.read(8 * dim * nrhs), dtype=np.float64)
    return dict(dim=dim, nrhs=nrhs, neg=neg, irn=irn, jcn=jcn, val=val, rhs=rhs, sol=sol)


def main():
    for name, N, which in CASES:
        with tempfile.TemporaryDirectory() as td:
            cmd = [DRIVER, "--backend", "oracle", "--problem", name, "--N", str(N), "--print-level", "0",
                   "--json", os.path.join(td, "r.json"), "--final", os.path.join(td, "final.bin"),
                   "--dump", os.path.join(td, "kkt"), "--dump-iters", ",".join(map(str, which))]
            subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
            summ = json.load(open(os.path.join(td, "r.json")))
            fin = read_final(os.path.join(td, "final.bin"))
            tag = "%s_%d" % (name, N)
            np.savez_compressed(os.path.join(HERE, tag + "_final.npz"), iterations=summ["iterations"],
                                n_factor=summ["n_factor"], n_solve=summ["n_solve"], status=summ["status"], **fin)
            snaps = {}
            for k in which:
                p = os.path.join(td, "kkt_%d.bin" % k)
                if not os.path.exists(p):
                    continue
                d = read_dump(p)
                if "irn" not in snaps:
                    snaps["irn"], snaps["jcn"], snaps["dim"] = d["irn"], d["jcn"], d["dim"]
                snaps["val_%d" % k] = d["val"]
                snaps["rhs_%d" % k] = d["rhs"]
                snaps["sol_%d" % k] = d["sol"]
                snaps["neg_%d" % k] = d["neg"]
            np.savez_compressed(os.path.join(HERE, tag + "_kkt.npz"), **snaps)
            print(tag, summ["iterations"], summ["objective"], sorted(k for k in snaps if k.startswith("val_")))


if __name__ == "__main__":
    sys.exit(main())
