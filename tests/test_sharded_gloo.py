"""N>1 path on the CPU: two gloo processes run the elimination-tree sharding schedule (partition from the C++
symbolic phase, contribution blocks / update vectors / solution pieces exchanged with torch.distributed) with a
numpy multifrontal emulation standing in for the CUDA kernels.  Checks the partition and the exchange logic."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, N, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from ipopt_b200 import SymbolicAnalysis
    from ipopt_b200.kkt import mbndry_kkt, to_scipy
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dim, irn, jcn, val, nc = mbndry_kkt(N, sigma_spread=1.0, delta_c=1e-2, seed=5)   # quasi-definite: no pivoting needed
        v0 = val.copy(); v0[-nc:] = 0.0
        S = SymbolicAnalysis(dim, irn, jcn, v0)
        owner, nsub = S.shard(world)
        perm, sn_start, rows_ptr, rows, rel = S.get("perm"), S.get("sn_start"), S.get("rows_ptr"), S.get("rows"), S.get("rel")
        parent, uent_ptr, u_dst64, t2u = S.get("sn_parent"), S.get("uent_ptr"), S.get("u_dst64"), S.get("t2u")
        nsn = len(parent)
        assert nsub >= 2 * world and (owner == -1).any() and all((owner == g).any() for g in range(world))
        # a top front never lies below a subtree front
        for s in range(nsn):
            if owner[s] == -1 and parent[s] >= 0:
                assert owner[parent[s]] == -1
            if owner[s] >= 0 and parent[s] >= 0 and owner[parent[s]] >= 0:
                assert owner[parent[s]] == owner[s]
        uval = np.zeros(len(u_dst64)); np.add.at(uval, t2u, val)
        children = [[] for _ in range(nsn)]
        for s in range(nsn):
            if parent[s] >= 0:
                children[parent[s]].append(s)
        cut = [s for s in range(nsn) if owner[s] >= 0 and (parent[s] < 0 or owner[parent[s]] < 0)]
        Ls, Ds, cbs = {}, {}, {}

        def factor_front(s):
            k = sn_start[s + 1] - sn_start[s]; r = rows_ptr[s + 1] - rows_ptr[s]; f = k + r
            P = np.zeros(f * k); P[u_dst64[uent_ptr[s]:uent_ptr[s + 1]]] = uval[uent_ptr[s]:uent_ptr[s + 1]]
            F = np.zeros((f, f)); F[:, :k] = P.reshape((k, f)).T
            F = np.tril(F) + np.tril(F, -1).T
            for c in children[s]:
                rl = rel[rows_ptr[c]:rows_ptr[c + 1]]
                F[np.ix_(rl, rl)] += cbs[c]
            L = np.eye(f)[:, :k].copy(); D = np.zeros(k)
            for j in range(k):
                D[j] = F[j, j]; L[j + 1:, j] = F[j + 1:, j] / D[j]
                F[j + 1:, j + 1:] -= np.outer(L[j + 1:, j], F[j + 1:, j])
            Ls[s], Ds[s], cbs[s] = L, D, F[k:, k:].copy()

        mine = [s for s in range(nsn) if owner[s] == rank]
        top = [s for s in range(nsn) if owner[s] == -1]
        for s in mine:
            factor_front(s)
        # contribution blocks of the cut -> rank 0
        for s in cut:
            r = rows_ptr[s + 1] - rows_ptr[s]
            if owner[s] == 0 or r == 0:
                continue
            if rank == owner[s]:
                dist.send(torch.from_numpy(cbs[s].copy()), dst=0)
            elif rank == 0:
                buf = torch.empty((r, r), dtype=torch.float64); dist.recv(buf, src=int(owner[s])); cbs[s] = buf.numpy()
        if rank == 0:
            for s in top:
                factor_front(s)
        # ---- solve ----
        b = np.random.default_rng(0).standard_normal(dim)
        x = b[perm].astype(float).copy()
        cbv = {}

        def fwd(s):
            a, e = sn_start[s], sn_start[s + 1]; k = e - a
            w = np.concatenate([x[a:e], np.zeros(rows_ptr[s + 1] - rows_ptr[s])])
            for c in children[s]:
                w[rel[rows_ptr[c]:rows_ptr[c + 1]]] += cbv[c]
            y = np.linalg.solve(Ls[s][:k, :k], w[:k])
            cbv[s] = w[k:] - Ls[s][k:, :] @ y
            x[a:e] = y / Ds[s]

        def bwd(s):
            a, e = sn_start[s], sn_start[s + 1]; k = e - a
            rw = rows[rows_ptr[s]:rows_ptr[s + 1]]
            x[a:e] = np.linalg.solve(Ls[s][:k, :k].T, x[a:e] - Ls[s][k:, :].T @ x[rw])

        for s in mine:
            fwd(s)
        for s in cut:
            r = rows_ptr[s + 1] - rows_ptr[s]
            if owner[s] == 0 or r == 0:
                continue
            if rank == owner[s]:
                dist.send(torch.from_numpy(cbv[s].copy()), dst=0)
            elif rank == 0:
                buf = torch.empty(r, dtype=torch.float64); dist.recv(buf, src=int(owner[s])); cbv[s] = buf.numpy()
        idx_top = np.concatenate([np.arange(sn_start[s], sn_start[s + 1]) for s in top])
        if rank == 0:
            for s in top:
                fwd(s)
            for s in reversed(top):
                bwd(s)
        buf = torch.from_numpy(x[idx_top].copy())
        dist.broadcast(buf, 0)
        x[idx_top] = buf.numpy()
        for s in reversed(mine):
            bwd(s)
        first = sn_start[:-1].copy()
        for s in range(nsn):
            if parent[s] >= 0 and first[s] < first[parent[s]]:
                first[parent[s]] = first[s]
        for s in cut:
            if owner[s] == 0:
                continue
            a, e = first[s], sn_start[s + 1]
            if rank == owner[s]:
                dist.send(torch.from_numpy(x[a:e].copy()), dst=0)
            elif rank == 0:
                buf = torch.empty(e - a, dtype=torch.float64); dist.recv(buf, src=int(owner[s])); x[a:e] = buf.numpy()
        if rank == 0:
            sol = np.zeros(dim); sol[perm] = x
            A = to_scipy(dim, irn, jcn, val)
            res = np.abs(A @ sol - b).max() / np.abs(b).max()
            q.put(("ok", float(res), int(nsub), int((owner == -1).sum())))
        dist.barrier()
    except Exception as e:  # pragma: no cover
        import traceback
        q.put(("err", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_sharded_schedule_world2_gloo(built_lib, world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, 14, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
    assert out[0] == "ok", out
    assert out[1] < 1e-10 and out[2] >= 2 * world and out[3] >= 1


def test_shard_plan_balance(built_lib):
    from ipopt_b200 import SymbolicAnalysis
    from ipopt_b200.kkt import mbndry_kkt
    dim, irn, jcn, val, nc = mbndry_kkt(60, w_zero=True)
    S = SymbolicAnalysis(dim, irn, jcn, val)
    k = np.diff(S.get("sn_start")).astype(float); r = np.diff(S.get("rows_ptr")).astype(float)
    w = k ** 3 / 3 + k * k * r + k * r * (r + 1)
    for G in (2, 4, 8):
        owner, nsub = S.shard(G)
        assert nsub >= 4 * G            # north_star: shard only when the tree exposes enough independent subtrees
        loads = np.array([w[owner == g].sum() for g in range(G)])
        assert loads.min() > 0 and loads.max() <= 1.35 * loads.mean()
    owner1, n1 = S.shard(1)
    assert np.all(owner1 == 0) and n1 == 1
