"""Development diagnostics: group-level timing of the chain kernel on a large single block (no per-row python loop)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rchol_b200 import problems, producer, capi
n, T, thr, win = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
A = problems.laplace_3d(n); f = producer.factor(*A, threads=T)
G = (f.rowPtr, f.colIdx, f.val); b = problems.random_rhs(f.N)
s = capi.Solver(0, chain_threads=thr, chain_window=win)
s.set_factor(*G, f.part if T > 0 else None)
y, tr = s.debug_trace(capi.TRSV_FORWARD, b)
y, tr = s.debug_trace(capi.TRSV_FORWARD, b)
part = f.part.astype(np.int64) if T > 0 else np.array([0, f.N])
lo, hi = part[0], part[1]
fin = tr[lo:hi, 0].astype(np.int64); trips = tr[lo:hi, 1].astype(np.int64); st = tr[lo:hi, 2].astype(np.int64)
# unwrap 32-bit cycle counters along the row order (monotone-ish)
def unwrap(x):
    d = np.diff(x); wraps = np.cumsum(d < -(1 << 31)); return np.concatenate([[x[0]], x[1:] + (wraps << 32)])
fin = unwrap(fin); st = unwrap(st)
rows = hi - lo
g = np.arange(0, rows - 31, 32)
gstart = st[g]; gfirst = np.minimum.reduceat(fin, g); glast = np.maximum.reduceat(fin, g)
print("rows", rows, "total cycles %.3e" % (fin.max() - st.min()), "cycles/row %.1f" % ((fin.max() - st.min()) / rows))
print("per group: start->first finish median %.0f ; first->last finish median %.0f ; last(g)->last(g+1) median %.0f mean %.0f" % (
    np.median(gfirst - gstart), np.median(glast - gfirst), np.median(np.diff(glast)), np.mean(np.diff(glast))))
print("trips/row mean %.1f ; cycles per trip median %.0f" % (trips.mean(), np.median((fin - st) / np.maximum(trips, 1))))
# slack: time between a group's polling start and the completion of the previous group (negative = started late)
slack = glast[:-1] - gstart[1:]
print("lookahead (prev group last finish - this group poll start): median %.0f p10 %.0f p1 %.0f ; late starts %.3f" % (
    np.median(slack), np.percentile(slack, 10), np.percentile(slack, 1), np.mean(slack < 0)))
