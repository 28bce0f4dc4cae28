"""Multi-GPU KKT factor+solve by elimination-tree sharding.

One process per GPU (``torch.distributed``, backend NCCL); every rank holds a full ``b200ldlt`` handle and the same
matrix.  The supernodal elimination tree is cut (``b200ldlt_shard_setup``: deterministic, identical on all ranks)
into a top part and balanced disjoint subtrees:

  factor : every rank factorises its subtrees  ->  the contribution blocks of the cut are sent to rank 0
           (NCCL send/recv straight from / into the handles' device storage)  ->  rank 0 factorises the top part;
           the inertia counters are all-reduced.
  solve  : forward sweeps over the subtrees  ->  update vectors of the cut to rank 0  ->  rank 0 does the top
           forward+backward  ->  the top part of the solution is broadcast  ->  backward sweeps over the subtrees
           ->  the subtree pieces of the solution are gathered on rank 0.

PyTorch is plumbing only here (device-pointer views, NCCL point-to-point); all numerics run in libb200ldlt.so.
The same orchestration can run all ranks inside ONE process on one GPU (``local_world``), which is how the
single-GPU functional tests exercise the sharded code path.
"""
import ctypes as C

import numpy as np

from .capi import B200Ldlt, load_library


class _DevArr:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _bind(L):
    if getattr(L, "_shard_bound", False):
        return
    vp = C.c_void_p
    L.b200ldlt_shard_setup.argtypes = [vp, C.c_int, C.c_int]
    L.b200ldlt_shard_array.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int64), C.c_int64]
    L.b200ldlt_shard_array.restype = C.c_int64
    L.b200ldlt_device_ptr.argtypes = [vp, C.c_char_p]
    L.b200ldlt_device_ptr.restype = vp
    L.b200ldlt_shard_factor.argtypes = [vp, C.c_int, C.c_int]
    L.b200ldlt_shard_factor_finish.argtypes = [vp, C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.b200ldlt_shard_solve.argtypes = [vp, C.c_int, vp]
    L._shard_bound = True


class _Rank:
    """One rank's handle + torch views of its device arrays."""

    def __init__(self, rank, world, dim, irn, jcn, first_vals, device, stream, **opts):
        import torch
        self.torch = torch
        self.rank = rank
        # the handle enqueues on the SAME (non-default) stream torch / NCCL use, so phases and exchanges are ordered
        self.s = B200Ldlt(device=device, stream=stream.cuda_stream, **opts)
        L = self.s._L
        _bind(L)
        assert self.s.InitializeStructure(dim, len(irn), irn, jcn) == 0
        self.s.GetValuesArrayPtr()[:] = first_vals
        assert self.s.analyse_now(first_vals) == 0, self.s.last_error()
        assert L.b200ldlt_shard_setup(self.s._h, rank, world) == 0, self.s.last_error()
        self.h = self.s._h
        self.L = L
        self.dim = dim
        sym = self.s.symbolic
        self.sn_start, self.sn_parent = sym("sn_start"), sym("sn_parent")
        self.rows_ptr, self.cb_off = sym("rows_ptr"), sym("cb_off")
        self.owner = self._arr("owner")
        self.cut_roots = self._arr("cut_roots")
        self.top_fronts = self._arr("top_fronts")
        nsn = len(self.sn_parent)
        first = self.sn_start[:-1].copy()
        for s in range(nsn):                      # children precede parents (postorder)
            p = self.sn_parent[s]
            if p >= 0 and first[s] < first[p]:
                first[p] = first[s]
        self.subtree_first = first
        dev = "cuda:%d" % device
        self.CB = self._view("CB", int(self.cb_off[-1]), "<f8", dev)
        self.cbv = self._view("cbv", max(int(self.rows_ptr[-1]), 1), "<f8", dev)
        self.x = self._view("x", dim, "<f8", dev)
        self.counters = self._view("counters", 8, "<i4", dev)
        self.vals_dev = self._view("vals", len(irn), "<f8", dev)
        top_cols = [np.arange(self.sn_start[s], self.sn_start[s + 1]) for s in self.top_fronts]
        self.idx_top = torch.from_numpy(np.concatenate(top_cols) if top_cols else np.zeros(0, np.int64)).to(dev)
        self.rhs = torch.empty(dim, dtype=torch.float64, device=dev)

    def _arr(self, name):
        n = self.L.b200ldlt_shard_array(self.h, name.encode(), None, 0)
        out = np.zeros(max(n, 1), dtype=np.int64)
        self.L.b200ldlt_shard_array(self.h, name.encode(), out.ctypes.data_as(C.POINTER(C.c_int64)), n)
        return out[:n]

    def _view(self, name, n, typestr, dev):
        ptr = self.L.b200ldlt_device_ptr(self.h, name.encode())
        assert ptr, name
        return self.torch.as_tensor(_DevArr(ptr, n, typestr), device=dev)

    def factor_phase(self, phase, from_host):
        assert self.L.b200ldlt_shard_factor(self.h, phase, int(from_host)) == 0, self.s.last_error()

    def solve_phase(self, phase):
        assert self.L.b200ldlt_shard_solve(self.h, phase, C.c_void_p(self.rhs.data_ptr())) == 0, self.s.last_error()


class ShardedLdlt:
    """Sharded factor/solve.  ``local_world`` = number of ranks emulated inside this process (tests);
    otherwise one rank per process with torch.distributed already initialised (NCCL)."""

    def __init__(self, dim, irn, jcn, first_vals, local_world=None, device=0, **opts):
        import torch
        self.torch = torch
        if local_world:
            self.world, self.my = int(local_world), list(range(int(local_world)))
            self.dist = None
        else:
            import torch.distributed as dist
            self.dist = dist
            self.world, self.my = dist.get_world_size(), [dist.get_rank()]
        self.stream = torch.cuda.Stream(device)   # a real stream: handle 0 (legacy default) would mean "library-owned"
        with torch.cuda.stream(self.stream):
            self.ranks = {r: _Rank(r, self.world, dim, irn, jcn, first_vals, device, self.stream, **opts) for r in self.my}
        any_r = self.ranks[self.my[0]]
        self.dim = dim
        self.owner = any_r.owner
        par = any_r.sn_parent
        # roots of the subtrees below the cut (their parent is in the top part, or they are roots of the forest)
        self.cut_roots = [s for s in range(len(par)) if self.owner[s] >= 0 and (par[s] < 0 or self.owner[par[s]] < 0)]
        self.n_subtrees = len(self.cut_roots)
        self.top_fraction = None
        self._neg = -1

    # ---- data movement ------------------------------------------------------------------------------------
    def _move(self, name, off, length, src, dst, ops):
        """tensor `name`[off:off+length] of rank src -> same place on rank dst."""
        if src == dst or length == 0:
            return
        if src in self.ranks and dst in self.ranks:
            getattr(self.ranks[dst], name)[off:off + length].copy_(getattr(self.ranks[src], name)[off:off + length])
        elif src in self.ranks:
            ops.append(self.dist.P2POp(self.dist.isend, getattr(self.ranks[src], name)[off:off + length], dst))
        elif dst in self.ranks:
            ops.append(self.dist.P2POp(self.dist.irecv, getattr(self.ranks[dst], name)[off:off + length], src))

    def _flush(self, ops):
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()

    def _cut_to_root(self, name, offs):
        ops = []
        for s in self.cut_roots:
            o = int(self.owner[s])
            self._move(name, int(offs[s]), int(offs[s + 1] - offs[s]), o, 0, ops)
        self._flush(ops)

    # ---- numeric phases -----------------------------------------------------------------------------------
    def factor(self, vals, check=False, expected=0):
        with self.torch.cuda.stream(self.stream):
            return self._factor(vals, check, expected)

    def solve(self, rhs):
        """rhs: float64 numpy array (dim); returns the solution (valid on rank 0 / in local mode)."""
        with self.torch.cuda.stream(self.stream):
            return self._solve(rhs)

    def factor_device(self, d_vals, check=False, expected=0):
        """Same as factor() with the triplet values already on this rank's GPU (torch float64 tensor)."""
        with self.torch.cuda.stream(self.stream):
            return self._factor(d_vals, check, expected, device=True)

    def solve_device(self, d_rhs):
        """rhs/solution as a device tensor (dim); the solution is valid on rank 0."""
        with self.torch.cuda.stream(self.stream):
            return self._solve(d_rhs, device=True)

    def _factor(self, vals, check, expected, device=False):
        for R in self.ranks.values():
            if device:
                R.vals_dev.copy_(vals)
                R.factor_phase(0, False)
            else:
                R.s.GetValuesArrayPtr()[:] = vals
                R.factor_phase(0, True)
        any_r = self.ranks[self.my[0]]
        self._cut_to_root("CB", any_r.cb_off)
        if 0 in self.ranks:
            self.ranks[0].factor_phase(1, False)
        # inertia / status counters: sum over ranks
        if self.dist is not None:
            tot = self.ranks[self.my[0]].counters.clone()
            self.dist.all_reduce(tot)
        else:
            tot = sum(R.counters.clone() for R in self.ranks.values())
        tot_h = tot.cpu().numpy().astype(np.int32)
        st = 0
        for R in self.ranks.values():
            neg = C.c_int(-1)
            st = R.L.b200ldlt_shard_factor_finish(R.h, tot_h.ctypes.data_as(C.POINTER(C.c_int)), int(check), int(expected), C.byref(neg))
            self._neg = neg.value
        return st, self._neg

    def _solve(self, rhs, device=False):
        torch = self.torch
        for R in self.ranks.values():
            R.rhs.copy_(rhs if device else torch.from_numpy(np.ascontiguousarray(rhs)))
            R.solve_phase(0)
        any_r = self.ranks[self.my[0]]
        self._cut_to_root("cbv", any_r.rows_ptr)
        if 0 in self.ranks:
            self.ranks[0].solve_phase(1)
        # top part of the solution -> every rank
        if self.dist is not None:
            R = self.ranks[self.my[0]]
            buf = R.x[R.idx_top] if R.rank == 0 else torch.empty(len(R.idx_top), dtype=torch.float64, device=R.x.device)
            self.dist.broadcast(buf, 0)
            if R.rank != 0:
                R.x[R.idx_top] = buf
        else:
            buf = self.ranks[0].x[self.ranks[0].idx_top]
            for r, R in self.ranks.items():
                if r != 0:
                    R.x[R.idx_top] = buf
        for R in self.ranks.values():
            R.solve_phase(2)
        # subtree pieces of the (permuted) solution -> rank 0
        ops = []
        for s in self.cut_roots:
            o = int(self.owner[s])
            a, b = int(any_r.subtree_first[s]), int(any_r.sn_start[s + 1])
            self._move("x", a, b - a, o, 0, ops)
        self._flush(ops)
        out = None
        if 0 in self.ranks:
            self.ranks[0].solve_phase(3)
            if device:
                return self.ranks[0].rhs
            self.stream.synchronize()
            out = self.ranks[0].rhs.cpu().numpy()
        return out

    def info(self):
        return {r: R.s.info() for r, R in self.ranks.items()}

    def close(self):
        for R in self.ranks.values():
            R.s.close()
