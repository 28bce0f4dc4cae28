"""numpy emulation of the GPU numeric algorithm INCLUDING its pivot rule (Bunch-Kaufman restricted to the
supernode's pivot block, threshold test against the whole front column, retry passes, forced pivots).
Debugging / design tool: lets the pivoting logic be studied on the CPU."""
import numpy as np

ALPHA = 0.6403882032022076


def factor_front(F, k, u=1e-8, tiny=1e-15):
    """F: f x f symmetric (full). Returns dict(npiv stats, L, D blocks, perm) - only statistics are used."""
    f = F.shape[0]
    j, kend, progress, forced = 0, k, 0, False
    stats = dict(n1=0, n2=0, forced=0, tiny=0, neg=0, fails=0)
    order = list(range(f))
    while j < k:
        if j == kend:
            if progress > 0:
                kend, progress = k, 0
            else:
                forced, kend = True, k
        col = np.abs(F[j + 1:, j])
        cand = col[:kend - j - 1]
        lam = cand.max() if cand.size else 0.0
        r = j + 1 + int(cand.argmax()) if cand.size else -1
        gam = col[kend - j - 1:].max() if col[kend - j - 1:].size else 0.0
        ajj = abs(F[j, j])
        ok1 = ajj > tiny and ajj >= u * max(lam, gam)
        typ = 0
        if forced:
            typ = 1
        elif lam == 0.0:
            typ = 1 if ok1 else 0
        elif ok1 and ajj >= ALPHA * lam:
            typ = 1
        else:
            rowr = np.abs(F[r, j:]).copy()
            rowr[r - j] = 0.0
            sig = rowr[:kend - j].max()
            gamr = rowr[kend - j:].max() if rowr[kend - j:].size else 0.0
            oth = np.ones(f - j, bool); oth[0] = False; oth[r - j] = False
            cj = np.abs(F[j:, j])[oth].max() if oth.any() else 0.0
            cr = np.abs(F[j:, r])[oth].max() if oth.any() else 0.0
            arr = abs(F[r, r])
            if ok1 and ajj * sig >= ALPHA * lam * lam:
                typ = 1
            elif arr > tiny and arr >= ALPHA * sig and arr >= u * max(sig, gamr):
                typ = 2
            else:
                a, b, c = F[j, j], F[r, j], F[r, r]
                det = a * c - b * b
                if lam > tiny and abs(det) > 0 and (abs(c) * cj + abs(b) * cr) * u <= abs(det) and (abs(a) * cr + abs(b) * cj) * u <= abs(det):
                    typ = 3

        def swap(p, q):
            if p == q:
                return
            F[[p, q], :] = F[[q, p], :]
            F[:, [p, q]] = F[:, [q, p]]
            order[p], order[q] = order[q], order[p]
        if typ == 0:
            swap(j, kend - 1)
            kend -= 1
            stats["fails"] += 1
            continue
        if typ == 2:
            swap(j, r)
        if typ == 3:
            swap(j + 1, r)
        if typ != 3:
            d = F[j, j]
            if forced:
                cm = np.abs(F[j:, j]).max()
                if cm <= 1e-12 or not abs(d) > tiny:
                    d = -1.5e-8 if d < 0 else 1.5e-8
                    stats["tiny"] += 1
                else:
                    stats["forced"] += 1
            l = F[j + 1:, j] / d
            F[j + 1:, j + 1:] -= np.outer(l, F[j + 1:, j])
            F[j + 1:, j] = l
            stats["n1"] += 1
            stats["neg"] += d < 0
            j += 1
        else:
            a, b, c = F[j, j], F[j + 1, j], F[j + 1, j + 1]
            det = a * c - b * b
            C = F[j + 2:, j:j + 2].copy()
            Lm = np.stack([(c * C[:, 0] - b * C[:, 1]) / det, (a * C[:, 1] - b * C[:, 0]) / det], axis=1)
            F[j + 2:, j + 2:] -= Lm @ C.T
            F[j + 2:, j:j + 2] = Lm
            stats["n2"] += 1
            stats["neg"] += 1 if det < 0 else (2 if a < 0 else 0)
            j += 2
        progress += 1
    return stats, order


def emulate_factor(S, dim, irn, jcn, val, scaling=2, verbose=False):
    perm = S.get("perm"); sn_start = S.get("sn_start"); rows_ptr = S.get("rows_ptr"); rows = S.get("rows")
    rel = S.get("rel"); parent = S.get("sn_parent"); uent_ptr = S.get("uent_ptr"); u_dst64 = S.get("u_dst64"); t2u = S.get("t2u")
    nsn = len(sn_start) - 1
    uval = np.zeros(len(u_dst64)); np.add.at(uval, t2u, val)
    # scaling like the GPU (power-of-two inf-norm sweeps) on the original triplets
    i0, j0 = irn.astype(np.int64) - 1, jcn.astype(np.int64) - 1
    sc = np.ones(dim)
    import scipy.sparse as sp
    A = sp.coo_matrix((val, (i0, j0)), shape=(dim, dim)).tocsr()
    A = A + sp.tril(A, -1).T + sp.triu(A, 1).T
    for _ in range(scaling):
        B = abs(sp.diags(sc) @ A @ sp.diags(sc))
        m = np.asarray(B.max(axis=1).todense()).ravel()
        e = np.frexp(np.where(m > 0, m, 1.0))[1]
        sc = sc * np.ldexp(1.0, -(e // 2 + ((e % 2 != 0) & (e < 0))))
    # scaled unique values: need original (row, col) of each unique entry
    ur = np.zeros(len(u_dst64), np.int64); uc = np.zeros(len(u_dst64), np.int64)
    ur[t2u] = np.maximum(i0, j0); uc[t2u] = np.minimum(i0, j0)
    uval = uval * sc[ur] * sc[uc]
    cbs = [None] * nsn
    children = [[] for _ in range(nsn)]
    for s in range(nsn):
        if parent[s] >= 0:
            children[parent[s]].append(s)
    tot = dict(n1=0, n2=0, forced=0, tiny=0, neg=0, fails=0)
    badfronts = []
    for s in range(nsn):
        k = sn_start[s + 1] - sn_start[s]; r = rows_ptr[s + 1] - rows_ptr[s]; f = k + r
        P = np.zeros(f * k); P[u_dst64[uent_ptr[s]:uent_ptr[s + 1]]] = uval[uent_ptr[s]:uent_ptr[s + 1]]
        F = np.zeros((f, f)); F[:, :k] = P.reshape((k, f)).T
        F = np.tril(F) + np.tril(F, -1).T
        for c in children[s]:
            rl = rel[rows_ptr[c]:rows_ptr[c + 1]]
            F[np.ix_(rl, rl)] += cbs[c]; cbs[c] = None
        F0 = F.copy()
        st, order = factor_front(F, k)
        for key in tot:
            tot[key] += st[key]
        if st["forced"] or st["tiny"]:
            badfronts.append((s, k, r, st, F0, perm[sn_start[s]:sn_start[s + 1]]))
        cbs[s] = F[k:, k:].copy()
    return tot, badfronts
