"""Round-2 leaves sweep: the same n^3 Laplacian factored by the reference with T = 2^k nested-dissection leaves, solved to
1e-8 on one GPU; one JSON line per T with the per-level launch times of both triangular solves (k_bc_solve) and the host
cost of producing the factor.  Also prints what the GPU box offers (cores, RAM, /tmp) -- the 512^3 feasibility question.
Usage: python scripts/r02_sweep.py n T [T ...]   (not a bench line; bench.py is)."""
import json, os, resource, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rchol_b200 import capi, problems  # noqa: E402


def box():
    mem = {}
    for ln in open("/proc/meminfo"):
        k, v = ln.split(":")
        if k in ("MemTotal", "MemAvailable"):
            mem[k] = int(v.split()[0]) / 1e6
    st = os.statvfs("/tmp")
    return dict(cores=os.cpu_count(), affinity=len(os.sched_getaffinity(0)), mem_gb=mem,
                tmp_free_gb=st.f_bavail * st.f_frsize / 1e9)


n = int(sys.argv[1])
print(json.dumps(dict(box=box())), flush=True)
for T in [int(a) for a in sys.argv[2:]]:
    t0 = time.time()
    d, info = bench.build_problem(n, T)
    N = d["A_rp"].shape[0] - 1
    B_iter = problems.algorithmic_bytes_per_iteration(N, int(d["A_rp"][-1]), int(d["G_rp"][-1]))
    out = dict(n=n, leaves=T, nnzG=int(d["G_rp"][-1]), bytes_per_iteration=B_iter, produce=info, produce_wall_s=time.time() - t0,
               peak_rss_gb=resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6)
    with capi.Solver(0) as s:
        t0 = time.time()
        s.set_matrix(d["A_rp"], d["A_ci"], d["A_v"])
        s.set_factor(d["G_rp"], d["G_ci"], d["G_v"], d["part"])
        s.set_rhs(d["b"])
        out["setup_wall_s"] = time.time() - t0
        s.pcg_resident(bench.TOL, bench.MAXIT)                       # warm-up
        relres, itr = s.pcg_resident(bench.TOL, bench.MAXIT)
        st = s.stats()
        ms = st["solve_ms"]
        out.update(iterations=itr, relres=relres, ms_per_iter=ms / max(itr, 1), solve_ms=ms, device_gb=st["device_bytes"] / 1e9,
                   gbs_per_iter=B_iter * itr / ms / 1e6, frac_of_peak=B_iter * itr / ms / 1e6 / bench.measured_peak_gbs()[0])
        pr = s.profile_iteration(2)
        out["split_ms"] = {k: pr[k] for k in ("trsv_ms", "spmv_ms", "blas1_ms")}
        lv = {}
        for direction, dname in ((capi.TRSV_FORWARD, "fwd"), (capi.TRSV_BACKWARD, "bwd")):
            for gi, g in enumerate(s.groups(direction)):
                lv[f"{dname}{gi}"] = dict(blocks=g["blocks"], rows=g["rows"], max_rows=g.get("max_rows"),
                                          ms=s.time_group(direction, gi, 0, 2))
        out["levels"] = lv
    print(json.dumps(out), flush=True)
    del d
