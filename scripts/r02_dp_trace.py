"""Per-warp time marks of one hop of a dense-panel level (development tool).  Usage: python scripts/r02_dp_trace.py n T dir level"""
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
ctas = [c for c in range(160) if tr[c, 0, 0] != 0]
print("CTAs traced", len(ctas))
names = ["m0 hop start", "m1 rowptr issued", "m2 dense done", "m3 near-pre issued", "m4 prefetch done", "m5 barrier A passed", "m6 near rows done",
         "m7 dense-pre issued"]
for c in (ctas[0], ctas[len(ctas) // 2], ctas[-1]):
    base = tr[c, :, 0].min()
    print(f"CTA {c}: marks relative to the CTA's first warp entering the hop (min / median / max over warps)")
    for k in range(8):
        v = tr[c, :, k] - base
        print(f"   {names[k]:22s} {v.min():7d} {int(np.median(v)):7d} {v.max():7d}   slowest warp {int(v.argmax())}")
    b = tr[c, 0, 8:16] - base
    print("   barrier A (thread 0): cta-synced, released, polled, fenced:", b[:4].tolist(), " barrier B:", b[4:8].tolist())
# across CTAs: duration of the pieces for warp 0 / max over warps
def span(a, b):
    return np.array([(tr[c, :, b] - tr[c, :, a]).max() for c in ctas])
for a, b, nm in ((0, 2, "dense (m0->m2)"), (2, 4, "pre+prefetch (m2->m4)"), (5, 6, "near rows (m5->m6)"), (6, 7, "dot+dense-pre (m6->m7)")):
    v = span(a, b)
    print(f"{nm:26s} max over warps, per CTA: min {v.min()} median {int(np.median(v))} max {v.max()} (CTA {ctas[int(v.argmax())]})")
wa = np.array([tr[c, 0, 10] - tr[c, 0, 8] for c in ctas])
wb = np.array([tr[c, 0, 14] - tr[c, 0, 12] for c in ctas])
print("barrier A wait (cta-synced -> polled): min", wa.min(), "median", int(np.median(wa)), "max", wa.max())
print("barrier B wait (cta-synced -> polled): min", wb.min(), "median", int(np.median(wb)), "max", wb.max())
hop = np.array([tr[c, 0, 15] - tr[c, 0, 0] for c in ctas])
print("whole hop (m0 -> barrier B fenced), warp 0: min", hop.min(), "median", int(np.median(hop)), "max", hop.max())
