// ORACLE -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// CPU restatement of the algorithm class the reference delegates to for this path: a sparse
// multifrontal LDL^T for symmetric indefinite matrices with threshold partial pivoting (1x1 / 2x2,
// Duff-Reid test) and DELAYED pivots, as done by the vendor solvers the reference binds:
//   * call protocol / status codes: reference src/Algorithm/LinearSolvers/IpSparseSymLinearSolverInterface.hpp:32-97,
//     IpSymLinearSolver.hpp:19-33;
//   * duplicate summing + triangle normalisation of the triplets: IpTripletToCSRConverter.cpp:154-197,352-359;
//   * inertia / singular / pivot-tolerance state machine: IpMumpsSolverInterface.cpp:247-306,448-541,592-610
//     (pivtol default 1e-6, pivtolmax 0.1, IncreaseQuality: pivtol <- min(pivtolmax, pivtol^0.5)).
// The arithmetic itself lives in MUMPS 5.x (coin-or-tools/ThirdParty-Mumps stable/3.0, reference
// .coin-or/Dependencies:3), which is NOT in /root/reference and not installed: PARITY UNPINNED against MUMPS.
// The oracle is pinned instead against (a) dense LAPACK-style eigenvalue inertia / numpy solves in
// tests/test_oracle.py, (b) the hs071 trace of reference doc/interfaces.dox:588-599 run through the
// reference's own IP loop (tests/golden), so its role as the checker of the CUDA path is sound.
//
// Deliberately independent of ipopt_b200/csrc: own ordering glue, own symbolic phase, different
// pivoting rule (threshold partial pivoting with delays instead of in-supernode Bunch-Kaufman).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

// BLAS-3 contribution-block update (ORACLE_BLAS3=1; bench.py's CPU-baseline legs): dgemm of the host BLAS the reference
// library itself is linked against (OpenBLAS, see oracle/refhost/Makefile), so the CPU column is not handicapped by
// rank-1 loops on the separator fronts.  Off by default: the parity fixtures were produced with the plain loops.
extern "C" void dgemm_(const char*, const char*, const int*, const int*, const int*, const double*, const double*, const int*,
                       const double*, const int*, const double*, double*, const int*);
extern "C" void openblas_set_num_threads(int);
static const bool g_blas3 = getenv("ORACLE_BLAS3") && atoi(getenv("ORACLE_BLAS3")) != 0;

extern "C" int METIS_NodeND(int64_t* nvtxs, int64_t* xadj, int64_t* adjncy, int64_t* vwgt, int64_t* options,
                            int64_t* perm, int64_t* iperm);

namespace {

double wall() {
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}

struct Front {          // factor data kept per supernode
  std::vector<int> ind; // global (permuted) ids: eliminated pivots first, then remaining rows
  int npiv = 0;
  std::vector<double> L;   // f x npiv column-major (unit diagonal implied; 2x2 sub-diagonal stored as 0)
  std::vector<double> d1;  // npiv: D diagonal
  std::vector<double> d2;  // npiv: off-diagonal of a 2x2 block at its first column, else 0
  std::vector<char> two;   // npiv: 1 at the first column of a 2x2 block
};

struct CB {             // contribution block handed to the parent
  std::vector<int> ind; // global ids, first ndelay are fully summed (delayed pivots)
  int ndelay = 0;
  std::vector<double> a; // m x m column-major, lower part valid
};

struct Oracle {
  int n = 0, nnz = 0;
  double pivtol = 1e-6, pivtolmax = 0.1;
  int scaling = 1, verbose = 0;
  std::vector<int> irn, jcn;
  std::vector<double> vals;
  bool analysed = false, factored = false;
  // symbolic
  std::vector<int> perm, iperm, parent_sn, sn_start;
  std::vector<int> cptr, cind, cmap;  // permuted lower CSC; cmap: triplet -> CSC slot
  std::vector<std::vector<int>> sn_rows, sn_children;
  // numeric
  std::vector<double> cval, scale;
  std::vector<Front> fronts;
  int num_neg = 0, num_delayed = 0, num_2x2 = 0, max_front = 0;
  int64_t nnzL = 0;
  double t_analyse = 0, t_factor = 0, t_solve = 0, flops = 0;
};

// ------------------------------------------------------------------------------------------------
void analyse(Oracle& O) {
  const int n = O.n;
  const int64_t nnz = O.nnz;
  double t0 = wall();
  // unique lower entries (original labels) with summed values
  std::vector<std::pair<int64_t, int>> key(nnz);
  for (int64_t e = 0; e < nnz; ++e) {
    int i = O.irn[e] - 1, j = O.jcn[e] - 1;
    key[e] = {(int64_t)std::min(i, j) * n + std::max(i, j), (int)e};
  }
  std::sort(key.begin(), key.end());
  std::vector<int> ur, uc;
  std::vector<double> uv;
  for (int64_t q = 0; q < nnz; ++q) {
    if (q == 0 || key[q].first != key[q - 1].first) {
      uc.push_back((int)(key[q].first / n)); ur.push_back((int)(key[q].first % n)); uv.push_back(0.0);
    }
    uv.back() += O.vals[key[q].second];
  }
  const int64_t nu = ur.size();
  std::vector<int64_t> xadj(n + 1, 0);
  for (int64_t u = 0; u < nu; ++u) if (ur[u] != uc[u]) { xadj[ur[u] + 1]++; xadj[uc[u] + 1]++; }
  for (int i = 0; i < n; ++i) xadj[i + 1] += xadj[i];
  std::vector<int> adj(xadj[n]);
  std::vector<double> adw(xadj[n]);
  {
    std::vector<int64_t> pos(xadj.begin(), xadj.end() - 1);
    for (int64_t u = 0; u < nu; ++u) if (ur[u] != uc[u]) {
      adw[pos[ur[u]]] = std::fabs(uv[u]); adj[pos[ur[u]]++] = uc[u];
      adw[pos[uc[u]]] = std::fabs(uv[u]); adj[pos[uc[u]]++] = ur[u];
    }
  }
  // pair every zero-diagonal row with its heaviest free neighbour (cf. MUMPS ICNTL(12) compressed
  // ordering for saddle-point matrices): simple one-pass greedy, heaviest edge per saddle row first.
  std::vector<double> diag(n, 0.0);
  for (int64_t u = 0; u < nu; ++u) if (ur[u] == uc[u]) diag[ur[u]] = uv[u];
  std::vector<int> mate(n, -1);
  {
    std::vector<int> sad;
    for (int i = 0; i < n; ++i) if (diag[i] == 0.0) sad.push_back(i);
    for (int pass = 0; pass < 2; ++pass)
      for (int s : sad) {
        if (mate[s] >= 0) continue;
        int best = -1; double bw = 0.0;
        for (int64_t p = xadj[s]; p < xadj[s + 1]; ++p) {
          int x = adj[p];
          if (diag[x] == 0.0 || mate[x] >= 0) continue;
          if (adw[p] > bw) { bw = adw[p]; best = x; }
        }
        if (best >= 0) { mate[s] = best; mate[best] = s; }
      }
  }
  // compressed graph + METIS
  std::vector<int> cn(n, -1), m1, m2;
  for (int i = 0; i < n; ++i) {
    if (cn[i] >= 0) continue;
    if (mate[i] < 0) { cn[i] = m1.size(); m1.push_back(i); m2.push_back(-1); }
    else {
      int x = diag[i] == 0.0 ? mate[i] : i, c = diag[i] == 0.0 ? i : mate[i];
      cn[x] = cn[c] = m1.size(); m1.push_back(x); m2.push_back(c);
    }
  }
  const int nc = m1.size();
  std::vector<int> corder(nc);
  std::iota(corder.begin(), corder.end(), 0);
  if (nc > 8) {
    std::vector<int64_t> cx(nc + 1, 0), ca, vw(nc), mp(nc), mip(nc);
    std::vector<int> mark(nc, -1);
    for (int v = 0; v < nc; ++v) {
      mark[v] = v;
      for (int mem : {m1[v], m2[v]}) {
        if (mem < 0) continue;
        for (int64_t p = xadj[mem]; p < xadj[mem + 1]; ++p) {
          int w = cn[adj[p]];
          if (mark[w] != v) { mark[w] = v; ca.push_back(w); }
        }
      }
      cx[v + 1] = ca.size();
      vw[v] = m2[v] >= 0 ? 2 : 1;
    }
    int64_t nv = nc;
    // ORACLE_METIS_SEED: another (equally valid) nested-dissection ordering, for solver-to-solver spread controls
    std::vector<int64_t> mopt(40, -1);   // METIS_NOPTIONS = 40, -1 = default; [8] = METIS_OPTION_SEED
    if (const char* e = getenv("ORACLE_METIS_SEED")) mopt[8] = atoll(e);
    if (!ca.empty() && METIS_NodeND(&nv, cx.data(), ca.data(), vw.data(), getenv("ORACLE_METIS_SEED") ? mopt.data() : nullptr, mp.data(), mip.data()) == 1)
      for (int v = 0; v < nc; ++v) corder[v] = (int)mp[v];
  }
  std::vector<int> perm(n), iperm(n);
  {
    int q = 0;
    for (int v = 0; v < nc; ++v) { perm[q++] = m1[corder[v]]; if (m2[corder[v]] >= 0) perm[q++] = m2[corder[v]]; }
    for (int k = 0; k < n; ++k) iperm[perm[k]] = k;
  }
  // etree (Liu) + postorder
  auto etree = [&](std::vector<int>& parent) {
    parent.assign(n, -1);
    std::vector<int> anc(n, -1);
    for (int k = 0; k < n; ++k)
      for (int64_t p = xadj[perm[k]]; p < xadj[perm[k] + 1]; ++p) {
        int i = iperm[adj[p]];
        while (i != -1 && i < k) { int nx = anc[i]; anc[i] = k; if (nx == -1) parent[i] = k; i = nx; }
      }
  };
  std::vector<int> parent;
  etree(parent);
  {
    std::vector<std::vector<int>> ch(n);
    std::vector<int> roots, post, stk, it(n, 0);
    for (int j = 0; j < n; ++j) (parent[j] < 0 ? roots : ch[parent[j]]).push_back(j);
    for (int r : roots) {
      stk.push_back(r);
      while (!stk.empty()) {
        int v = stk.back();
        if (it[v] < (int)ch[v].size()) stk.push_back(ch[v][it[v]++]);
        else { post.push_back(v); stk.pop_back(); }
      }
    }
    std::vector<int> p2(n);
    for (int t = 0; t < n; ++t) p2[t] = perm[post[t]];
    perm.swap(p2);
    for (int k = 0; k < n; ++k) iperm[perm[k]] = k;
    etree(parent);
  }
  // column counts by explicit symbolic elimination (child structures merged into the parent)
  std::vector<int> cc(n, 1);
  {
    std::vector<std::vector<int>> st(n), ch(n);
    for (int j = 0; j < n; ++j) if (parent[j] >= 0) ch[parent[j]].push_back(j);
    std::vector<int> mark(n, -1);
    for (int j = 0; j < n; ++j) {
      auto& out = st[j];
      for (int64_t p = xadj[perm[j]]; p < xadj[perm[j] + 1]; ++p) {
        int i = iperm[adj[p]];
        if (i > j && mark[i] != j) { mark[i] = j; out.push_back(i); }
      }
      for (int c : ch[j]) {
        for (int i : st[c]) if (i != j && mark[i] != j) { mark[i] = j; out.push_back(i); }
        std::vector<int>().swap(st[c]);
      }
      cc[j] = out.size() + 1;
    }
  }
  // supernodes: fundamental chains, saddle pairs kept together, tiny chains merged
  O.sn_start.clear();
  std::vector<int> sn_of(n);
  for (int j = 0; j < n; ++j) {
    bool join = false;
    if (j > 0 && parent[j - 1] == j) {
      int a = perm[j - 1], b = perm[j];
      if (cc[j - 1] == cc[j] + 1) join = true;
      if (mate[a] == b) join = true;
      if (j - O.sn_start.back() < 8 && cc[j - 1] <= cc[j] + 2) join = true;
    }
    if (!join) O.sn_start.push_back(j);
    sn_of[j] = O.sn_start.size() - 1;
  }
  const int nsn = O.sn_start.size();
  O.sn_start.push_back(n);
  O.parent_sn.assign(nsn, -1);
  O.sn_children.assign(nsn, {});
  for (int s = 0; s < nsn; ++s) {
    int last = O.sn_start[s + 1] - 1;
    if (parent[last] >= 0) { O.parent_sn[s] = sn_of[parent[last]]; O.sn_children[O.parent_sn[s]].push_back(s); }
  }
  // permuted lower CSC with triplet map
  {
    std::vector<std::pair<int64_t, int>> k2(nnz);
    for (int64_t e = 0; e < nnz; ++e) {
      int i = iperm[O.irn[e] - 1], j = iperm[O.jcn[e] - 1];
      k2[e] = {(int64_t)std::min(i, j) * n + std::max(i, j), (int)e};
    }
    std::sort(k2.begin(), k2.end());
    O.cptr.assign(n + 1, 0);
    O.cind.clear();
    O.cmap.assign(nnz, 0);
    for (int64_t q = 0; q < nnz; ++q) {
      if (q == 0 || k2[q].first != k2[q - 1].first) {
        O.cind.push_back((int)(k2[q].first % n));
        O.cptr[(int)(k2[q].first / n) + 1]++;
      }
      O.cmap[k2[q].second] = O.cind.size() - 1;
    }
    for (int j = 0; j < n; ++j) O.cptr[j + 1] += O.cptr[j];
  }
  // static row structures
  O.sn_rows.assign(nsn, {});
  {
    std::vector<int> mark(n, -1);
    for (int s = 0; s < nsn; ++s) {
      int a = O.sn_start[s], e = O.sn_start[s + 1];
      auto& rw = O.sn_rows[s];
      for (int j = a; j < e; ++j)
        for (int p = O.cptr[j]; p < O.cptr[j + 1]; ++p) {
          int i = O.cind[p];
          if (i >= e && mark[i] != s) { mark[i] = s; rw.push_back(i); }
        }
      for (int c : O.sn_children[s])
        for (int i : O.sn_rows[c]) if (i >= e && mark[i] != s) { mark[i] = s; rw.push_back(i); }
      std::sort(rw.begin(), rw.end());
    }
  }
  O.perm = perm; O.iperm = iperm;
  O.analysed = true;
  O.t_analyse = wall() - t0;
}

// ------------------------------------------------------------------------------------------------
// Dense partial factorisation of a symmetric front (lower part, column-major, ld = f): eliminate as many
// of the first p (fully-summed) columns as the threshold test allows.  On exit the first npiv positions
// hold L / D, order[] is the applied symmetric permutation (position -> original local index).
struct DenseResult { int npiv; };

inline double& AT(std::vector<double>& F, int f, int i, int j) { return F[(size_t)j * f + i]; }

void swap_sym_lower(std::vector<double>& F, int f, int a, int b, std::vector<int>& order) {
  if (a == b) return;
  if (a > b) std::swap(a, b);
  // rows/cols a<b of a lower-stored symmetric matrix
  for (int j = 0; j < a; ++j) std::swap(AT(F, f, a, j), AT(F, f, b, j));
  for (int i = a + 1; i < b; ++i) std::swap(AT(F, f, i, a), AT(F, f, b, i));
  for (int i = b + 1; i < f; ++i) std::swap(AT(F, f, i, a), AT(F, f, i, b));
  std::swap(AT(F, f, a, a), AT(F, f, b, b));
  std::swap(order[a], order[b]);
}

int factor_front(std::vector<double>& F, int f, int p, double u, std::vector<int>& order,
                 std::vector<double>& d1, std::vector<double>& d2, std::vector<char>& two, int& neg, int& n2,
                 double& flops, bool is_root, int& nforced) {
  const double tiny = 1e-15;  // relative to the equilibrated matrix (row maxima ~ 1): below this a pivot is rounding noise
  int npiv = 0, pend = p, progress = 0;
  std::vector<double> c1(f), c2(f);
  bool force = false;
  // The pivot search only ever reads the first p columns (all rows), so the rank-1/rank-2 updates are applied eagerly to
  // those columns only; the contribution block (rows/columns >= p) receives the same updates, in the same pivot order and
  // with the same operands (=> bit-identical result), in one cache-friendly pass at the end.  Wcb keeps the unscaled
  // pivot columns (= L*D) of the contribution rows for that pass.
  const int q = f - p;
  std::vector<double> Wcb((size_t)q * p);
  while (npiv < p) {
    if (npiv == pend) {
      if (progress > 0) { pend = p; progress = 0; }
      else if (is_root) { force = true; pend = p; }
      else break;  // the rest is delayed to the parent
    }
    const int j = npiv;
    double ajj = std::fabs(AT(F, f, j, j));
    double lam = 0.0, rest = 0.0;
    int r = -1;
    for (int i = j + 1; i < f; ++i) {
      double v = std::fabs(AT(F, f, i, j));
      if (i < pend) { if (v > lam) { lam = v; r = i; } }
      else rest = std::max(rest, v);
    }
    double cmax = std::max(lam, rest);
    int type = 0;
    if (force) type = 1;
    else if (ajj > tiny && ajj >= u * cmax) type = 1;
    else if (r >= 0 && lam > tiny) {
      // 2x2 candidate (j, r): Duff-Reid test  |D^-1| * [max col j ; max col r] <= 1/u
      double a = AT(F, f, j, j), b = AT(F, f, r, j), c = AT(F, f, r, r);
      double det = a * c - b * b;
      double mj = 0.0, mr = 0.0;
      for (int i = j + 1; i < f; ++i) {
        if (i == r) continue;
        mj = std::max(mj, std::fabs(AT(F, f, i, j)));
        mr = std::max(mr, std::fabs(i < r ? AT(F, f, r, i) : AT(F, f, i, r)));
      }
      double ad = std::fabs(det);
      if (ad > tiny && (std::fabs(c) * mj + std::fabs(b) * mr) * u <= ad && (std::fabs(a) * mr + std::fabs(b) * mj) * u <= ad)
        type = 2;
      else {
        // 1x1 on r?
        double arr = std::fabs(c), mx = std::max(mr, std::fabs(b));
        if (arr > tiny && arr >= u * mx) type = 3;
      }
    }
    if (type == 0) { swap_sym_lower(F, f, j, pend - 1, order); --pend; continue; }
    if (type == 3) { swap_sym_lower(F, f, j, r, order); type = 1; }
    if (type == 1) {
      double d = AT(F, f, j, j);
      // forced pivot at a root: the whole remaining column is rounding noise => numerically singular
      if (force && !(std::max(ajj, cmax) > 1e-12)) { d = d < 0 ? -1e-8 : 1e-8; ++nforced; }
      d1[j] = d; d2[j] = 0.0; two[j] = 0;
      if (d < 0) ++neg;
      for (int i = j + 1; i < f; ++i) { c1[i] = AT(F, f, i, j); AT(F, f, i, j) = c1[i] / d; }
      for (int i = p; i < f; ++i) Wcb[(size_t)j * q + (i - p)] = c1[i];
      const int m0 = j + 1;
#pragma omp parallel for schedule(static) if ((int64_t)(p - m0) * (f - m0) > 40000)
      for (int m = m0; m < p; ++m) {
        double cm = c1[m];
        if (cm == 0.0) continue;
        double* col = &F[(size_t)m * f];
        const double* lj = &F[(size_t)j * f];
        for (int i = m; i < f; ++i) col[i] -= lj[i] * cm;
      }
      flops += (double)(f - m0) * (f - m0);
      npiv += 1;
    } else {
      swap_sym_lower(F, f, j + 1, r, order);
      double a = AT(F, f, j, j), b = AT(F, f, j + 1, j), c = AT(F, f, j + 1, j + 1);
      double det = a * c - b * b;
      d1[j] = a; d1[j + 1] = c; d2[j] = b; d2[j + 1] = 0.0; two[j] = 1; two[j + 1] = 0;
      ++n2;
      if (det < 0) neg += 1; else if (a < 0) neg += 2;
      for (int i = j + 2; i < f; ++i) {
        c1[i] = AT(F, f, i, j); c2[i] = AT(F, f, i, j + 1);
        AT(F, f, i, j) = (c * c1[i] - b * c2[i]) / det;
        AT(F, f, i, j + 1) = (a * c2[i] - b * c1[i]) / det;
      }
      AT(F, f, j + 1, j) = 0.0;
      for (int i = p; i < f; ++i) { Wcb[(size_t)j * q + (i - p)] = c1[i]; Wcb[(size_t)(j + 1) * q + (i - p)] = c2[i]; }
      const int m0 = j + 2;
#pragma omp parallel for schedule(static) if ((int64_t)(p - m0) * (f - m0) > 40000)
      for (int m = m0; m < p; ++m) {
        double a1 = c1[m], a2 = c2[m];
        if (a1 == 0.0 && a2 == 0.0) continue;
        double* col = &F[(size_t)m * f];
        const double* l1 = &F[(size_t)j * f];
        const double* l2 = &F[(size_t)(j + 1) * f];
        for (int i = m; i < f; ++i) col[i] -= l1[i] * a1 + l2[i] * a2;
      }
      flops += 2.0 * (f - m0) * (f - m0);
      npiv += 2;
    }
    ++progress;
  }
  // deferred update of the contribution block: column m, pivots in elimination order
  if (g_blas3 && q > 0 && npiv > 0 && (int64_t)q * q * npiv > 200000) {
    // C(q x q, ld f) -= L21 (q x npiv, ld f) * Wcb^T  (Wcb is npiv rows of q: column-major q x npiv with ld q)
    const double one = 1.0, mone = -1.0;
    dgemm_("N", "T", &q, &q, &npiv, &mone, &F[p], &f, Wcb.data(), &q, &one, &F[(size_t)p * f + p], &f);
  } else if (q > 0 && npiv > 0) {
#pragma omp parallel for schedule(dynamic, 4) if ((int64_t)q * q * npiv > 2000000)
    for (int m = p; m < f; ++m) {
      double* col = &F[(size_t)m * f];
      for (int t = 0; t < npiv; ++t) {
        if (two[t]) {
          const double a1 = Wcb[(size_t)t * q + (m - p)], a2 = Wcb[(size_t)(t + 1) * q + (m - p)];
          const double* l1 = &F[(size_t)t * f];
          const double* l2 = &F[(size_t)(t + 1) * f];
          ++t;
          if (a1 == 0.0 && a2 == 0.0) continue;
          for (int i = m; i < f; ++i) col[i] -= l1[i] * a1 + l2[i] * a2;
        } else {
          const double cm = Wcb[(size_t)t * q + (m - p)];
          if (cm == 0.0) continue;
          const double* lj = &F[(size_t)t * f];
          for (int i = m; i < f; ++i) col[i] -= lj[i] * cm;
        }
      }
    }
  }
  return npiv;
}

int factor(Oracle& O) {
  const int n = O.n;
  double t0 = wall();
  if (!O.analysed) analyse(O);
  const int nsn = O.sn_start.size() - 1;
  // values into the permuted CSC (duplicates summed)
  O.cval.assign(O.cind.size(), 0.0);
  for (int e = 0; e < O.nnz; ++e) O.cval[O.cmap[e]] += O.vals[e];
  // symmetric inf-norm equilibration (MUMPS scales too: ICNTL(8)=77, IpMumpsSolverInterface.cpp:402-404)
  O.scale.assign(n, 1.0);
  for (int sw = 0; sw < O.scaling; ++sw) {
    std::vector<double> mx(n, 0.0);
    for (int j = 0; j < n; ++j)
      for (int p = O.cptr[j]; p < O.cptr[j + 1]; ++p) {
        int i = O.cind[p];
        double v = std::fabs(O.cval[p]) * O.scale[i] * O.scale[j];
        mx[i] = std::max(mx[i], v); mx[j] = std::max(mx[j], v);
      }
    for (int i = 0; i < n; ++i) if (mx[i] > 0) O.scale[i] /= std::sqrt(mx[i]);
  }
  O.fronts.assign(nsn, Front());
  std::vector<CB> cbs(nsn);
  O.num_neg = O.num_delayed = O.num_2x2 = 0; O.max_front = 0; O.nnzL = 0; O.flops = 0;
  int nforced = 0;
  // Fronts of one tree level (height above the leaves) are independent: levels with many fronts are processed by
  // concurrent threads (one front per thread, private work space), levels with few fronts one after the other with the
  // parallel loops inside factor_front.  The arithmetic of a front does not depend on the schedule.
  std::vector<int> level(nsn, 0);
  int nlev = 0;
  for (int s = 0; s < nsn; ++s) {   // children precede parents
    for (int c : O.sn_children[s]) level[s] = std::max(level[s], level[c] + 1);
    nlev = std::max(nlev, level[s] + 1);
  }
  std::vector<std::vector<int>> by_level(nlev);
  for (int s = 0; s < nsn; ++s) by_level[level[s]].push_back(s);
  struct Work { std::vector<double> F; std::vector<int> ind, order, pos; };
  struct Tally { int neg = 0, n2 = 0, delayed = 0, forced = 0, max_front = 0; int64_t nnzL = 0; };
  std::vector<double> front_flops(nsn, 0.0);
  auto process_front = [&](int s, Work& W, Tally& T) {
    std::vector<double>& F = W.F;
    std::vector<int>& ind = W.ind;
    std::vector<int>& order = W.order;
    std::vector<int>& pos = W.pos;
    if ((int)pos.size() != n) pos.assign(n, -1);
    const int a = O.sn_start[s], e = O.sn_start[s + 1];
    ind.clear();
    for (int j = a; j < e; ++j) ind.push_back(j);
    for (int c : O.sn_children[s])
      for (int t = 0; t < cbs[c].ndelay; ++t) ind.push_back(cbs[c].ind[t]);
    const int p = ind.size();
    for (int i : O.sn_rows[s]) ind.push_back(i);
    const int f = ind.size();
    T.max_front = std::max(T.max_front, f);
    for (int t = 0; t < f; ++t) pos[ind[t]] = t;
    F.assign((size_t)f * f, 0.0);
    for (int j = a; j < e; ++j)
      for (int q = O.cptr[j]; q < O.cptr[j + 1]; ++q) {
        int i = O.cind[q];
        int li = pos[i], lj = pos[j];
        AT(F, f, std::max(li, lj), std::min(li, lj)) += O.cval[q] * O.scale[i] * O.scale[j];
      }
    for (int c : O.sn_children[s]) {
      CB& cb = cbs[c];
      const int m = cb.ind.size();
      for (int jj = 0; jj < m; ++jj) {
        int lj = pos[cb.ind[jj]];
        for (int ii = jj; ii < m; ++ii) {
          int li = pos[cb.ind[ii]];
          AT(F, f, std::max(li, lj), std::min(li, lj)) += cb.a[(size_t)jj * m + ii];
        }
      }
      CB().a.swap(cb.a); cb.ind.clear(); cb.ind.shrink_to_fit();
    }
    order.resize(f);
    std::iota(order.begin(), order.end(), 0);
    Front& fr = O.fronts[s];
    fr.d1.assign(p, 0.0); fr.d2.assign(p, 0.0); fr.two.assign(p, 0);
    const bool is_root = O.parent_sn[s] < 0;
    int npiv = factor_front(F, f, p, O.pivtol, order, fr.d1, fr.d2, fr.two, T.neg, T.n2, front_flops[s], is_root, T.forced);
    fr.npiv = npiv;
    fr.d1.resize(npiv); fr.d2.resize(npiv); fr.two.resize(npiv);
    fr.ind.resize(f);
    for (int t = 0; t < f; ++t) fr.ind[t] = ind[order[t]];
    fr.L.assign((size_t)f * npiv, 0.0);
    for (int j = 0; j < npiv; ++j) {
      fr.L[(size_t)j * f + j] = 1.0;
      for (int i = j + 1; i < f; ++i) fr.L[(size_t)j * f + i] = AT(F, f, i, j);
    }
    T.nnzL += (int64_t)npiv * f - (int64_t)npiv * (npiv - 1) / 2;
    const int m = f - npiv;
    if (!is_root) {
      CB& cb = cbs[s];
      cb.ndelay = p - npiv;
      T.delayed += cb.ndelay;
      cb.ind.assign(fr.ind.begin() + npiv, fr.ind.end());
      cb.a.assign((size_t)m * m, 0.0);
      for (int jj = 0; jj < m; ++jj)
        for (int ii = jj; ii < m; ++ii) cb.a[(size_t)jj * m + ii] = AT(F, f, npiv + ii, npiv + jj);
    }
    for (int t = 0; t < f; ++t) pos[ind[t]] = -1;
  };
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = omp_get_max_threads();
#endif
  std::vector<Work> work(nthreads);
  std::vector<Tally> tally(nthreads);
  for (int l = 0; l < nlev; ++l) {
    const double tl0 = wall();
    const std::vector<int>& fl = by_level[l];
    if (g_blas3) openblas_set_num_threads((nthreads > 1 && (int)fl.size() >= std::max(2, nthreads / 4)) ? 1 : nthreads);
    if (nthreads > 1 && (int)fl.size() >= std::max(2, nthreads / 4)) {
#pragma omp parallel for schedule(dynamic, 1)
      for (int q = 0; q < (int)fl.size(); ++q) {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        process_front(fl[q], work[tid], tally[tid]);
      }
    } else {
      for (int s : fl) process_front(s, work[0], tally[0]);
    }
    if (getenv("ORACLE_PROFILE")) fprintf(stderr, "[oracle] level %d: %zu fronts %.1f ms\n", l, fl.size(), (wall() - tl0) * 1e3);
  }
  for (const Tally& T : tally) {
    O.num_neg += T.neg; O.num_2x2 += T.n2; O.num_delayed += T.delayed; nforced += T.forced;
    O.max_front = std::max(O.max_front, T.max_front); O.nnzL += T.nnzL;
  }
  for (int s = 0; s < nsn; ++s) O.flops += front_flops[s];
  O.factored = true;
  O.t_factor = wall() - t0;
  if (O.verbose)
    fprintf(stderr, "[oracle] factor %.3fs neg=%d 2x2=%d delayed=%d forced=%d maxfront=%d nnzL=%lld\n", O.t_factor,
            O.num_neg, O.num_2x2, O.num_delayed, nforced, O.max_front, (long long)O.nnzL);
  return nforced > 0 ? 1 : 0;
}

void solve(Oracle& O, double* b) {
  double t0 = wall();
  const int n = O.n;
  std::vector<double> x(n);
  for (int k = 0; k < n; ++k) x[k] = b[O.perm[k]] * O.scale[k];
  // NOTE: scale[] is indexed by permuted id (it was built on the permuted CSC)
  std::vector<double> y;
  for (const Front& fr : O.fronts) {
    const int f = fr.ind.size(), np = fr.npiv;
    if (!np) continue;
    y.resize(f);
    for (int t = 0; t < f; ++t) y[t] = x[fr.ind[t]];
    for (int j = 0; j < np; ++j) {
      double yj = y[j];
      if (yj == 0.0) continue;
      const double* col = &fr.L[(size_t)j * f];
      for (int i = j + 1; i < f; ++i) y[i] -= col[i] * yj;
    }
    for (int t = 0; t < f; ++t) x[fr.ind[t]] = y[t];
  }
  for (const Front& fr : O.fronts) {
    for (int j = 0; j < fr.npiv; ++j) {
      if (fr.two[j]) {
        double a = fr.d1[j], bb = fr.d2[j], c = fr.d1[j + 1], det = a * c - bb * bb;
        double u = x[fr.ind[j]], v = x[fr.ind[j + 1]];
        x[fr.ind[j]] = (c * u - bb * v) / det;
        x[fr.ind[j + 1]] = (a * v - bb * u) / det;
        ++j;
      } else x[fr.ind[j]] /= fr.d1[j];
    }
  }
  for (int s = (int)O.fronts.size() - 1; s >= 0; --s) {
    const Front& fr = O.fronts[s];
    const int f = fr.ind.size(), np = fr.npiv;
    if (!np) continue;
    y.resize(f);
    for (int t = 0; t < f; ++t) y[t] = x[fr.ind[t]];
    for (int j = np - 1; j >= 0; --j) {
      const double* col = &fr.L[(size_t)j * f];
      double acc = 0.0;
      for (int i = j + 1; i < f; ++i) acc += col[i] * y[i];
      y[j] -= acc;
    }
    for (int t = 0; t < np; ++t) x[fr.ind[t]] = y[t];
  }
  for (int k = 0; k < n; ++k) b[O.perm[k]] = x[k] * O.scale[k];
  O.t_solve = wall() - t0;
}

}  // namespace

extern "C" {

void* oracle_ldlt_create(double pivtol, double pivtolmax, int scaling, int verbose) {
  Oracle* O = new Oracle();
  if (pivtol > 0) O->pivtol = pivtol;
  if (pivtolmax > 0) O->pivtolmax = pivtolmax;
  O->scaling = scaling; O->verbose = verbose;
  return O;
}
void oracle_ldlt_destroy(void* h) { delete (Oracle*)h; }

int oracle_ldlt_analyse(void* h, int dim, int nonzeros, const int* irn, const int* jcn) {
  Oracle& O = *(Oracle*)h;
  if (dim <= 0 || nonzeros < 0) return 4;
  for (int e = 0; e < nonzeros; ++e)
    if (irn[e] < 1 || irn[e] > dim || jcn[e] < 1 || jcn[e] > dim) return 4;
  O.n = dim; O.nnz = nonzeros;
  O.irn.assign(irn, irn + nonzeros); O.jcn.assign(jcn, jcn + nonzeros);
  O.vals.assign(std::max(nonzeros, 1), 0.0);
  O.analysed = O.factored = false;
  return 0;
}
double* oracle_ldlt_values_ptr(void* h) { return ((Oracle*)h)->vals.data(); }

// status codes == Ipopt::ESymSolverStatus (IpSymLinearSolver.hpp:19-33)
int oracle_ldlt_factor(void* h, int check_inertia, int expected_neg, int* num_neg) {
  Oracle& O = *(Oracle*)h;
  if (O.n <= 0) return 4;
  int sing = factor(O);
  if (num_neg) *num_neg = O.num_neg;
  if (sing) return 1;
  if (check_inertia && O.num_neg != expected_neg) return 2;
  return 0;
}
int oracle_ldlt_solve(void* h, int nrhs, double* rhs) {
  Oracle& O = *(Oracle*)h;
  if (!O.factored) return 4;
  for (int c = 0; c < nrhs; ++c) solve(O, rhs + (size_t)c * O.n);
  return 0;
}
int oracle_ldlt_num_neg(void* h) { return ((Oracle*)h)->num_neg; }
int oracle_ldlt_increase_quality(void* h) {
  Oracle& O = *(Oracle*)h;
  if (O.pivtol >= O.pivtolmax) return 0;
  O.pivtol = std::min(O.pivtolmax, std::pow(O.pivtol, 0.5));  // IpMumpsSolverInterface.cpp:592-610
  return 1;
}
void oracle_ldlt_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
// out: [t_analyse, t_factor, t_solve, nnzL, flops, max_front, num_delayed, num_2x2]
void oracle_ldlt_stats(void* h, double* out) {
  Oracle& O = *(Oracle*)h;
  out[0] = O.t_analyse; out[1] = O.t_factor; out[2] = O.t_solve; out[3] = (double)O.nnzL; out[4] = O.flops;
  out[5] = O.max_front; out[6] = O.num_delayed; out[7] = O.num_2x2;
}

}  // extern "C"
