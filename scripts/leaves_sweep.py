"""How the iteration time depends on the number of nested-dissection leaves (the reference's `threads` argument of
rchol(A, G, P, threads)): the same n^3 Laplacian factored with T = 8, 64, 256, 1024 leaves, solved to 1e-8 on one GPU.
With T = 8 (BASELINE.json configs[1]) only 8 SMs carry a chain; configs[2] (512^3, T = 2^k) leaves k free.
Usage: python scripts/leaves_sweep.py [n=256] [T ...]      - one JSON line per T (not a bench line; bench.py is)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rchol_b200 import capi, problems  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
leaves = [int(a) for a in sys.argv[2:]] or [8, 64, 256, 1024]
for T in leaves:
    d, info = bench.build_problem(n, T)
    N = d["A_rp"].shape[0] - 1
    B_iter = problems.algorithmic_bytes_per_iteration(N, int(d["A_rp"][-1]), int(d["G_rp"][-1]))
    out = dict(n=n, leaves=T, nnzG=int(d["G_rp"][-1]), bytes_per_iteration=B_iter)
    with capi.Solver(0) as s:
        t0 = time.time()
        s.set_matrix(d["A_rp"], d["A_ci"], d["A_v"])
        s.set_factor(d["G_rp"], d["G_ci"], d["G_v"], d["part"])
        s.set_rhs(d["b"])
        out["setup_wall_s"] = time.time() - t0
        s.pcg_resident(bench.TOL, bench.MAXIT)                       # warm-up
        relres, itr = s.pcg_resident(bench.TOL, bench.MAXIT)
        ms = s.stats()["solve_ms"]
        out.update(iterations=itr, relres=relres, ms_per_iter=ms / max(itr, 1), solve_ms=ms,
                   gbs_per_iter=B_iter * itr / ms / 1e6, frac_of_peak=B_iter * itr / ms / 1e6 / bench.measured_peak_gbs()[0])
        st = s.profile_iteration(2)
        out["split_ms"] = {k: st[k] for k in ("trsv_ms", "spmv_ms", "blas1_ms")}
    print(json.dumps(out), flush=True)
