"""Development diagnostics: per-row timing trace of the chain kernel."""
import os, sys, numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rchol_b200 import problems, producer, capi
n, T, thr = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
which = capi.TRSV_FORWARD if len(sys.argv) < 5 or sys.argv[4] == "fwd" else capi.TRSV_BACKWARD
A = problems.laplace_3d(n); f = producer.factor(*A, threads=T)
G = (f.rowPtr, f.colIdx, f.val); b = problems.random_rhs(f.N)
s = capi.Solver(0, chain_threads=thr, chain_window=8192)
s.set_factor(*G, f.part if T > 0 else None)
y, tr = s.debug_trace(which, b)
y, tr = s.debug_trace(which, b)
print("stats", {k: v for k, v in s.stats().items() if k in ("chain_sm_mhz", "watchdog_row", "n_blocks")})
N = f.N
U = sp.csr_matrix((f.val, f.colIdx.astype(np.int64), f.rowPtr.astype(np.int64)), shape=(N, N))
if which == capi.TRSV_FORWARD:
    L = U.T.tocsr()
else:  # reversed index space
    perm = np.arange(N)[::-1]
    L = U[perm][:, perm].tocsr()
L.sort_indices()
part = f.part.astype(np.int64) if T > 0 else np.array([0, N])
if which == capi.TRSV_BACKWARD:
    part = (N - part)[::-1]
    lo, hi = part[-2], part[-1]      # last block in reversed space = first leaf
else:
    lo, hi = part[0], part[1]
fin = tr[:, 0].astype(np.int64); trips = tr[:, 1].astype(np.int64); st = tr[:, 2].astype(np.int64)
lp, lc = L.indptr, L.indices
t0 = st[lo:hi].min()
print("block rows", hi - lo, "total cycles", fin[lo:hi].max() - t0)
# levels
lev = np.zeros(hi - lo, np.int64)
hops = np.zeros(hi - lo, np.int64); nloc = np.zeros(hi - lo, np.int64)
for j in range(lo, hi):
    cs = lc[lp[j]:lp[j+1]-1]; cs = cs[cs >= lo]
    nloc[j-lo] = len(cs)
    if len(cs):
        lev[j-lo] = lev[cs-lo].max() + 1
        hops[j-lo] = fin[j] - fin[cs].max()
    else:
        hops[j-lo] = 0
nlev = lev.max() + 1
print("levels", nlev, "cycles/level %.0f" % ((fin[lo:hi].max() - t0) / nlev), "local deps/row %.1f" % nloc.mean())
h = hops[nloc > 0]
print("hop cycles: mean %.0f median %.0f p10 %.0f p90 %.0f p99 %.0f" % (h.mean(), np.median(h), np.percentile(h, 10), np.percentile(h, 90), np.percentile(h, 99)))
dur = fin[lo:hi] - st[lo:hi]
print("row (finish-start): mean %.0f median %.0f ; trips mean %.1f ; cycles/trip median %.0f" % (dur.mean(), np.median(dur), trips[lo:hi].mean(), np.median(dur / np.maximum(trips[lo:hi], 1))))
# critical path analysis: walk back from the last finished row along the latest-finishing dependency
j = lo + int(np.argmax(fin[lo:hi])); path = []
while True:
    cs = lc[lp[j]:lp[j+1]-1]; cs = cs[cs >= lo]
    if not len(cs): break
    c = cs[np.argmax(fin[cs])]
    path.append((fin[j] - fin[c], (j - lo) // 32 == (c - lo) // 32, j - c))
    j = c
path = np.array(path)
print("critical path hops", len(path), "mean hop %.0f; same-group hops %.2f (mean %.0f) cross-group (mean %.0f)" % (path[:, 0].mean(), path[:, 1].mean(), path[path[:, 1] == 1, 0].mean(), path[path[:, 1] == 0, 0].mean()))
big = path[path[:, 0] > 1000]
print("hops > 1000 cycles on the path: %d, total %.0f cycles of %.0f" % (len(big), big[:, 0].sum(), path[:, 0].sum()))
# group start lag: time between group start and its first row's finish
g0 = np.arange(lo, hi, 32)
print("group (first finish - start) median %.0f ; start-to-start of consecutive groups median %.0f" % (np.median(fin[g0] - st[g0]), np.median(np.diff(st[g0]))))
