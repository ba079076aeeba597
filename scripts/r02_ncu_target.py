"""Target process for ncu: builds (or loads) a factor and applies the preconditioner a few times (development tool).
Usage: python scripts/r02_ncu_target.py n T reps key=value ...   (keys of capi.Solver)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rchol_b200 import capi  # noqa: E402

n, T, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
opts = {k: int(v) for k, v in (a.split("=") for a in sys.argv[4:])}
d, _ = bench.build_problem(n, T)
with capi.Solver(0, use_graph=False, **opts) as s:
    s.set_matrix(d["A_rp"], d["A_ci"], d["A_v"])
    s.set_factor(d["G_rp"], d["G_ci"], d["G_v"], d["part"] if T > 0 else None)
    for _ in range(reps):
        z = s.precond(d["b"])
    print("ok", float(abs(z).sum()))
