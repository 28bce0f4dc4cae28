"""BASELINE.json config 5: MBndryCntrl1 N=800 KKT (dim 1 283 200, 5 126 400 triplets), elimination-tree subtrees sharded
across the GPUs of one node (torchrun), next to the single-GPU time of the same matrix measured on rank 0.
Synthetic Sigma values on the exact MBndryCntrl1 pattern (ipopt_b200/kkt.py); run:
  python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 scripts/bench_config5.py [N]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import torch.distributed as dist
from ipopt_b200 import B200Ldlt
from ipopt_b200.kkt import mbndry_kkt
from ipopt_b200.sharded import ShardedLdlt

N = int(sys.argv[1]) if len(sys.argv) > 1 else 800
K = 5
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
torch.cuda.set_device(lr)
dim, irn, jcn, val, nc = mbndry_kkt(N, sigma_spread=3.0, seed=1)
_, _, _, v0, _ = mbndry_kkt(N, w_zero=True)
b = np.random.default_rng(0).standard_normal(dim)
out = {"config": "MBndryCntrl1 N=%d KKT, dim %d, %d triplets" % (N, dim, len(irn)), "n_gpus": world}
if rank == 0:
    s = B200Ldlt(device=lr)
    s.InitializeStructure(dim, len(irn), irn, jcn)
    s.GetValuesArrayPtr()[:] = v0
    t0 = time.perf_counter(); st, neg = s.factor(True, nc); out["single_gpu_first_factor_s"] = time.perf_counter() - t0
    s.GetValuesArrayPtr()[:] = val
    fac, sol = [], []
    for _ in range(K):
        st, neg = s.factor(True, nc); assert st == 0 and neg == nc, (st, neg)
        x = b.copy(); s.solve(x)
        i = s.info(); fac.append(i["ms_factor_gpu"]); sol.append(i["ms_solve_gpu"])
    r, xi, bi = s.residual(x, b)
    i = s.info()
    out["single_gpu"] = {"factor_ms": float(np.median(fac)), "solve_ms": float(np.median(sol)), "step_ms": float(np.median(fac) + 2 * np.median(sol)),
                         "scaled_residual": r / (xi + bi), "nnz_L": i["nnz_L"], "flops": i["flops_panel"] + i["flops_schur"],
                         "max_front": i["max_front"], "levels": i["nlevels"], "supernodes": i["nsupernodes"],
                         "tri_solve_GBs": (16 * i["nnz_L"] + 32 * dim) / (np.median(sol) * 1e-3) / 1e9}
    s.close()
dist.barrier()
sh = ShardedLdlt(dim, irn, jcn, v0, device=lr)
d_val = torch.from_numpy(val).cuda(); d_b = torch.from_numpy(b).cuda()
for _ in range(2):
    st, neg = sh.factor_device(d_val, True, nc); sh.solve_device(d_b)
assert st == 0 and neg == nc, (st, neg)
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(sh.stream)
for _ in range(K):
    sh.factor_device(d_val, True, nc); sh.solve_device(d_b); sh.solve_device(d_b)
e1.record(sh.stream)
dist.barrier(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / K], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
x = sh.solve(b)
if rank == 0:
    r, xi, bi = sh.ranks[0].s.residual(x, b)
    own = sh.owner
    out["sharded"] = {"step_ms": float(t[0]), "subtrees": sh.n_subtrees, "top_fronts": int((own == -1).sum()), "scaled_residual": r / (xi + bi)}
    out["speedup_vs_single_gpu"] = out["single_gpu"]["step_ms"] / out["sharded"]["step_ms"]
    print(json.dumps(out))
sh.close()
dist.destroy_process_group()
