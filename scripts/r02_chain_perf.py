"""Per-level timing of the triangular solves for several chain variants on one factor (development tool, not a bench line).
Usage: python scripts/r02_chain_perf.py n T variant [variant ...]
  variant = name:key=value,key=value   (keys of capi.Solver: chain_mode, recent, early, chain_window, ...; dbg=1 adds the
  cycle counters of the chain warp of CTA 0 for the LAST launch timed)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import bench  # noqa: E402
from rchol_b200 import capi, problems  # noqa: E402

n, T = int(sys.argv[1]), int(sys.argv[2])
d, info = bench.build_problem(n, T)
N = d["A_rp"].shape[0] - 1
B_iter = problems.algorithmic_bytes_per_iteration(N, int(d["A_rp"][-1]), int(d["G_rp"][-1]))
ref = None
for spec in sys.argv[3:]:
    name, _, kv = spec.partition(":")
    opts = {k: int(v) for k, v in (p.split("=") for p in kv.split(",") if p)}
    out = dict(n=n, leaves=T, variant=name, opts=opts)
    try:
        with capi.Solver(0, **opts) as s:
            t0 = time.time()
            s.set_matrix(d["A_rp"], d["A_ci"], d["A_v"])
            s.set_factor(d["G_rp"], d["G_ci"], d["G_v"], d["part"] if T > 0 else None)
            s.set_rhs(d["b"])
            st = s.stats()
            out.update(setup_wall_s=time.time() - t0, analysis_ms=st["analysis_ms"], device_gb=st["device_bytes"] / 1e9)
            for direction in (capi.TRSV_FORWARD, capi.TRSV_BACKWARD):
                lay = (capi.C.c_uint64 * 16)()
                s._check(s._L.rcg_debug_blocked_info(s._h, direction, lay))
                out[f"blob_gb_{direction}"] = dict(A=lay[4] / 1e9, B=lay[5] / 1e9, far_nnz=int(lay[6]), chunks=int(lay[1]))
            s.pcg_resident(bench.TOL, bench.MAXIT)
            relres, itr = s.pcg_resident(bench.TOL, bench.MAXIT)
            ms = s.stats()["solve_ms"]
            x = s.solution()
            if ref is None:
                ref = x
            out.update(iterations=itr, relres=relres, ms_per_iter=ms / max(itr, 1), gbs=B_iter * itr / ms / 1e6,
                       x_vs_first_variant=float(np.linalg.norm(x - ref) / np.linalg.norm(ref)))
            pr = s.profile_iteration(2)
            out["split_ms"] = {k: pr[k] for k in ("trsv_ms", "spmv_ms", "blas1_ms")}
            lv = {}
            for direction, dname in ((capi.TRSV_FORWARD, "fwd"), (capi.TRSV_BACKWARD, "bwd")):
                for gi, g in enumerate(s.groups(direction)):
                    ms_g = s.time_group(direction, gi, 0, 2)
                    e = dict(blocks=g["blocks"], rows=g["rows"], max_rows=g["max_rows"], ms=round(ms_g, 4))
                    if opts.get("dbg", 0) & 1:
                        c = s.counters()
                        mhz = s.stats()["chain_sm_mhz"] or 1965.0
                        e["cta0"] = dict(total_cyc=c[0], waitA=c[3], waitU=c[4], chunks=c[6], late_u=c[15],
                                         cyc_per_chunk=c[0] / max(c[6], 1), helper0=c[8:13])
                        # dense-panel levels (k_dp_solve, CTA 0): total, barrier waits, phase 0, phase 1, phase 2, hops, CTAs
                        e["dp"] = dict(total=c[3], sync=c[4], p0=c[5], p1=c[6], p2=c[7], hops=c[8], ctas=c[9])
                    lv[f"{dname}{gi}"] = e
            out["levels"] = lv
    except Exception as e:  # keep going: the other variants still tell something
        out["error"] = repr(e)[:400]
    print(json.dumps(out), flush=True)
