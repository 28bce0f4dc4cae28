import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from ipopt_b200 import B200Ldlt
from ipopt_b200.kkt import mbndry_kkt
from ipopt_b200.sharded import ShardedLdlt
dim, irn, jcn, val, nc = mbndry_kkt(60, sigma_spread=3.0, seed=9)
_, _, _, v0, _ = mbndry_kkt(60, w_zero=True)
for world in (1, 2):
    sh = ShardedLdlt(dim, irn, jcn, v0, local_world=world)
    own = sh.owner
    print("world", world, "subtrees", sh.n_subtrees, "top", (own == -1).sum(), [int((own == g).sum()) for g in range(world)])
    st, neg = sh.factor(val, True, nc)
    print("  status", st, "neg", neg, "expected", nc, {r: R.counters.cpu().numpy().tolist() for r, R in sh.ranks.items()})
    b = np.ones(dim); x = sh.solve(b)
    R0 = sh.ranks[0]
    print("  resid", R0.s.residual(x, b))
    sh.close()
