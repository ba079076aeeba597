"""Informational leg of bench.py: the headline workload with the leaf level on the cluster chain (chain_mode 4).  Runs in
its own process (bench.py starts it with a time-out) so that nothing it does can disturb the measured run; prints one JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rchol_b200 import capi  # noqa: E402

n, threads, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
d, _ = bench.build_problem(n, threads)
A = (d["A_rp"], d["A_ci"], d["A_v"]); G = (d["G_rp"], d["G_ci"], d["G_v"])
out = {}
with capi.Solver(0, chain_mode=4) as s:
    t0 = time.time()
    s.set_matrix(*A)
    s.set_factor(*G, d["part"] if threads > 0 else None)
    s.set_rhs(d["b"])
    out["setup_wall_s"] = time.time() - t0
    out["cluster_levels"] = [s.cluster_levels(capi.TRSV_FORWARD), s.cluster_levels(capi.TRSV_BACKWARD)]
    relres, itr = s.pcg_resident(bench.TOL, bench.MAXIT)
    ms, its = 0.0, 0
    for _ in range(steps):
        relres, itr = s.pcg_resident(bench.TOL, bench.MAXIT)
        ms += s.stats()["solve_ms"]; its += itr
    out.update(iterations=itr, relres=relres, ms_per_iter=ms / max(its, 1), device_bytes=s.stats()["device_bytes"])
    lv = {}
    for direction, dn in ((capi.TRSV_FORWARD, "fwd"), (capi.TRSV_BACKWARD, "bwd")):
        for gi, g in enumerate(s.groups(direction)):
            lv[f"{dn}:level{gi}:{g['blocks']}blocks"] = s.time_group(direction, gi, 0, 3)
    out["tree_levels_ms"] = lv
print(json.dumps(out), flush=True)
