"""Time marks (globaltimer, ns) of one hop of a dense-panel level (development tool).
Usage: python scripts/r02_dp_trace.py n T dir level"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import bench  # noqa: E402
from rchol_b200 import capi  # noqa: E402

n, T, direction, gi = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
d, _ = bench.build_problem(n, T)
with capi.Solver(0, dbg=2) as s:
    s.set_matrix(d["A_rp"], d["A_ci"], d["A_v"])
    s.set_factor(d["G_rp"], d["G_ci"], d["G_v"], d["part"] if T > 0 else None)
    s.set_rhs(d["b"])
    s.pcg_resident(1e-8, 3)
    print("groups", [(g["blocks"], g["max_rows"]) for g in s.groups(direction)])
    ms = s.time_group(direction, gi, 0, 2)
    tr = s.dp_trace().astype(np.int64)
print("level ms", ms)
def stats(name, v):
    v = np.asarray(v)
    if v.size == 0:
        print(f"{name:40s} (none)"); return
    print(f"{name:40s} n={v.size:5d} min {v.min():7d} p10 {int(np.percentile(v, 10)):7d} med {int(np.median(v)):7d} p90 {int(np.percentile(v, 90)):7d} max {v.max():7d}")
dn = tr[:, :, 4]
ctas = [c for c in range(160) if dn[c, 0] != 0]
T0 = min(dn[c, 0] for c in ctas)
print("CTAs with a dense task in the traced hop:", len(ctas), " all times in ns after the first of them started waiting for t1")
stats("dense(H)  start waiting (warp 0)", [tr[c, 0, 4] - T0 for c in ctas])
stats("dense(H)  t1 there (slowest warp of CTA)", [tr[c, :, 5].max() - T0 for c in ctas])
stats("dense(H)  synced (warp 0)", [tr[c, 0, 6] - T0 for c in ctas])
stats("dense(H)  x stored (warp 0)", [tr[c, 0, 7] - T0 for c in ctas])
order = np.argsort([tr[c, 0, 7] for c in ctas])
print("   x stored, by CTA (= task) index, every 16th:", [(ctas[k], int(tr[ctas[k], 0, 7] - T0)) for k in range(0, len(ctas), 16)])
nr = [(c, w) for c in range(160) for w in range(32) if tr[c, w, 3] != 0]
stats("near(H+1) indicator word there", [tr[c, w, 0] - T0 for c, w in nr])
stats("near(H+1) first gathers there", [tr[c, w, 1] - T0 for c, w in nr if tr[c, w, 1]])
stats("near(H+1) row tail done", [tr[c, w, 2] - T0 for c, w in nr])
stats("near(H+1) t1 stored", [tr[c, w, 3] - T0 for c, w in nr])
stats("near(H+1) tail time", [tr[c, w, 2] - max(tr[c, w, 1], tr[c, w, 0]) for c, w in nr])
c2 = [c for c in range(160) if tr[c, 0, 12] != 0]
stats("dense(H+1) start waiting (warp 0)", [tr[c, 0, 12] - T0 for c in c2])
stats("dense(H+1) t1 there (slowest warp)", [tr[c, :, 13].max() - T0 for c in c2])
stats("dense(H+1) x stored (warp 0)", [tr[c, 0, 15] - T0 for c in c2])
print("   x stored, by CTA (= task) index, every 16th:", [(c2[k], int(tr[c2[k], 0, 15] - T0)) for k in range(0, len(c2), 16)])
