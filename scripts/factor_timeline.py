"""Kernel timeline of one factorisation (debug build switch B200_FACTOR_TIMELINE): %globaltimer records per kernel role."""
import os, sys, ctypes, collections
os.environ["B200_FACTOR_TIMELINE"] = "1"
import numpy as np
sys.path.insert(0, ".")
import ipopt_b200.capi as _capi
if os.environ.get("B200_LIBDIR"):      # experiments: a differently configured build of the library
    _capi.lib_path = lambda: os.path.join(os.environ["B200_LIBDIR"], "libb200ldlt.so")
from ipopt_b200 import B200Ldlt
from ipopt_b200.kkt import mbndry_kkt
N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/factor_timeline_N%d.txt" % N
PARSE_ONLY = os.environ.get("B200_TL_PARSE") is not None     # re-read an existing dump (no GPU)
if not PARSE_ONLY:
  dim, irn, jcn, val, nc = mbndry_kkt(N, sigma_spread=3.0, seed=1)
  s = B200Ldlt()
  s.InitializeStructure(dim, len(irn), irn, jcn)
  a = s.GetValuesArrayPtr()
  a[:] = val
  for it in range(3):
      st, neg = s.factor(True, nc)
  print("factor ms", s.info()["ms_factor_gpu"], "status", st, "2x2 pivots", s.info().get("num_2x2"))
  s._L.b200ldlt_dump_factor_timeline.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
  assert s._L.b200ldlt_dump_factor_timeline(s._h, out.encode()) == 0
recs = [list(map(int, l.split())) for l in open(out)]
pass
names = {1: "chain", 2: "rows", 3: "update", 4: "front_smem", 5: "front_warp", 6: "schur", 7: "extend", 8: "cb_update", 10: "linv_diag", 11: "linv_g1", 12: "linv_g2"}
# chain records carry the packed pivot-path statistics of the diagonal-block factorisation as a 7th word: not a time
for r in recs:
    if r[0] == 1 and len(r) > 13: r.append(("prof", r.pop(12), r.pop(12)))
def times(r): return [x for x in r[6:] if not isinstance(x, tuple)]
t0 = min(min(times(r)) for r in recs)
print("records", len(recs), " span %.1f us" % ((max(max(times(r)) for r in recs) - t0) / 1e3))
# per level and kind: first start, last end
lv = collections.defaultdict(lambda: [1 << 62, 0, 0])
for r in recs:
    key = (r[3], r[0]) if r[0] < 10 else (99, r[0])
    e = lv[key]
    e[0] = min(e[0], min(times(r))); e[1] = max(e[1], max(times(r))); e[2] += 1
print("level kind          start(us)   end(us)  dur(us)  records")
for key in sorted(lv, key=lambda k: lv[k][0]):
    e = lv[key]
    print("%5d %-12s %9.1f %9.1f %8.1f %6d" % (key[0], names.get(key[1], key[1]), (e[0] - t0) / 1e3, (e[1] - t0) / 1e3, (e[1] - e[0]) / 1e3, e[2]))
# chain statistics
ch = [r for r in recs if r[0] == 1]
if ch:
    ph = np.array([[r[7] - r[6], r[8] - r[7], r[9] - r[8], r[10] - r[9], r[11] - r[10]] for r in ch if r[2] >= 32], float)
    print("chain role (steps with both updates): n=%d  loads %.2f  B %.2f  C %.2f  LDLT %.2f  store %.2f  total %.2f us" % ((len(ph),) + tuple(ph.mean(0) / 1e3) + (ph.sum(1).mean() / 1e3,)))
    by = collections.defaultdict(list)
    for r in ch: by[r[1]].append(r)
    gaps = []; per = []
    for s_, L in by.items():
        L.sort(key=lambda r: r[2])
        for a_, b_ in zip(L[:-1], L[1:]):
            gaps.append(b_[6] - a_[11]); per.append(b_[6] - a_[6])
    print("gap end(chain p) -> start(chain p+1): mean %.2f us   period mean %.2f us  (n=%d)" % (np.mean(gaps) / 1e3, np.mean(per) / 1e3, len(gaps)))
    big = max(by, key=lambda q: len(by[q]))
    print("longest chain: front", big, "k", by[big][0][4], "f", by[big][0][5])
    rows = {(r[1], r[2]): r for r in recs if r[0] == 2}
    upd = collections.defaultdict(list)
    for r in recs:
        if r[0] == 3: upd[(r[1], r[2])].append(r)
    for r in by[big][:12]:
        rr = [x for x in recs if x[0] == 2 and x[1] == big and x[2] == r[2]]
        uu = upd.get((big, r[2]), [])
        print("  jb %4d chain %7.1f..%7.1f (ld %.1f B %.1f C %.1f D %.1f st %.1f)  rows %s  update %s" % (
            r[2], (r[6] - t0) / 1e3, (r[11] - t0) / 1e3, (r[7] - r[6]) / 1e3, (r[8] - r[7]) / 1e3, (r[9] - r[8]) / 1e3, (r[10] - r[9]) / 1e3, (r[11] - r[10]) / 1e3,
            " ".join("%.1f..%.1f" % ((x[6] - t0) / 1e3, (x[7] - t0) / 1e3) for x in rr),
            " ".join("%.1f..%.1f" % ((x[6] - t0) / 1e3, (x[7] - t0) / 1e3) for x in uu)))
if ch:
    nf = cf = ns = cs = 0
    for r in ch:
        if isinstance(r[-1], tuple):
            nf += r[-1][1] >> 40; cf += r[-1][1] & ((1 << 40) - 1); ns += r[-1][2] >> 40; cs += r[-1][2] & ((1 << 40) - 1)
    if nf + ns: print("diagonal LDL^T steps: fast 1x1 %d at %.0f cycles, other %d at %.0f cycles" % (nf, cf / max(nf, 1), ns, cs / max(ns, 1)))
for kind in (2, 3, 4, 5, 6, 7, 8):
    d = [times(r)[-1] - r[6] for r in recs if r[0] == kind]
    if d: print("%-12s n=%5d mean %.2f us max %.2f us" % (names[kind], len(d), np.mean(d) / 1e3, np.max(d) / 1e3))
fs = [r for r in recs if r[0] == 4]
if fs:
    ph = np.array([[r[7] - r[6], r[8] - r[7], r[9] - r[8]] for r in fs], float)
    print("front_smem phases (first/last CTA): assemble %.1f factor %.1f store %.1f us" % tuple(ph.mean(0) / 1e3))
