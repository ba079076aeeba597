"""Development probe (not a test): parity + timings of the blocked triangular solve for a few configurations.
usage: gpu_bc_probe.py n T [window,recent ...]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rchol_b200 import problems, producer, capi
from oracle import oracle


def relerr(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


n, T = int(sys.argv[1]), int(sys.argv[2])
cfgs = [tuple(int(v) for v in a.split(",")) for a in sys.argv[3:]] or [(0, 0)]
if os.environ.get("RCHOL_PROBE_CACHE"):      # share the problem (and its factorization) with bench.py's on-disk cache
    import bench
    d, _ = bench.build_problem(n, T)
    A = (d["A_rp"], d["A_ci"], d["A_v"]); G = (d["G_rp"], d["G_ci"], d["G_v"]); b = d["b"]
    part = d["part"] if T > 0 else None
    N = A[0].shape[0] - 1
    print(f"=== lap3d n={n} T={T} (bench cache) nnzG {int(G[0][-1])}", flush=True)
else:
    t = time.time(); A = problems.laplace_3d(n); f = producer.factor(*A, threads=T)
    print(f"=== lap3d n={n} T={T} factor {time.time()-t:.1f}s nnzG {f.nnz}", flush=True)
    G = (f.rowPtr, f.colIdx, f.val)
    b = problems.random_rhs(f.N)
    if T > 0:
        A = producer.ref_reorder(*A, f.P); b = problems.reorder_vector(b, f.P)
    part = f.part if T > 0 else None
    N = f.N
check = N <= 3_000_000
if check:
    t = time.time(); yo = oracle.trsv_forward(*G, b); zo = oracle.trsv_backward(*G, yo); print(f"oracle trsv {time.time()-t:.1f}s", flush=True)
for cfg in cfgs:
    win, rec = cfg[0], cfg[1] if len(cfg) > 1 else 0
    mode = cfg[2] if len(cfg) > 2 else 0
    dbg = cfg[3] if len(cfg) > 3 else 0
    sepw = cfg[4] if len(cfg) > 4 else 0
    early = cfg[5] if len(cfg) > 5 else 0
    capq = cfg[6] if len(cfg) > 6 else 0
    sa = cfg[7] if len(cfg) > 7 else 0
    fl2 = cfg[8] if len(cfg) > 8 else 0
    stile = cfg[9] if len(cfg) > 9 else 0
    esep = cfg[10] if len(cfg) > 10 else 0
    s = capi.Solver(0, chain_window=win, recent=rec, chain_mode=mode, dbg=dbg, sep_window=sepw, early=early, capb_quarters=capq, slots_a=sa, far_lanes2=fl2, sep_tile=stile, early_sep=esep)
    t = time.time(); s.set_matrix(*A); s.set_factor(*G, part); t_set = time.time() - t
    st = s.stats()
    print(f"--- window={win} recent={rec} mode={mode} sep_window={sepw} early={early} capq={capq} SA={sa} farlanes2={fl2} septile={stile} esep={esep} dbg={dbg}: set-up wall {t_set:.2f}s upload {st['upload_ms']:.0f} analysis {st['analysis_ms']:.0f} ms", flush=True)
    if check:
        y = s.trsv(capi.TRSV_FORWARD, b); z = s.trsv(capi.TRSV_BACKWARD, yo); zz = s.precond(b)
        print(f"    fwd relerr {relerr(y, yo):.2e} bwd {relerr(z, zo):.2e} precond {relerr(zz, zo):.2e}", flush=True)
    s.set_rhs(b)
    rr, it = s.pcg_resident(1e-8, int(os.environ.get('RCHOL_PROBE_MAXIT', 500)))
    st = s.stats()
    ph = [s.time_phase(p, 3) for p in range(4)]
    print(f"    pcg it={it} relres={rr:.3e} solve_ms={st['solve_ms']:.2f} ms/it={st['solve_ms']/max(it,1):.3f} | spmv {ph[0]:.3f} fwd {ph[1]:.3f} bwd {ph[2]:.3f} vec {ph[3]:.3f} ms | launches/it {st['launches_per_iteration']}", flush=True)
    for d, name in ((capi.TRSV_FORWARD, "fwd"), (capi.TRSV_BACKWARD, "bwd")):
        gs = s.groups(d)
        lay = s._L  # noqa
        for gi, g in enumerate(gs):
            ms = s.time_group(d, gi, 0, 3)
            print(f"      {name} level {gi}: blocks {g['blocks']} rows {g['rows']} chain-bytes {g['loc_nnz']} far-nnz {g['ext_nnz']} maxA {g['max_stage']}: {ms:.3f} ms"
                  f"  ({g['rows']/32/max(g['blocks'],1)/ms/1e3 if ms>0 else 0:.2f} chunks/us/block)", flush=True)
            if dbg:
                c = s.counters()
                nchk = max(c[6], 1)
                print(f"        CTA0 total {c[0]} cyc; critical per chunk {c[0]/nchk:.0f}: blockedA {c[3]/nchk:.0f} t' {c[4]/nchk:.0f} recent {c[5]/nchk:.0f} matvec {c[13]/nchk:.0f} preload+store {c[14]/nchk:.0f} late-t' {c[15]/nchk:.3f} barrier {c[7]/nchk:.0f} (chunks {c[6]});"
                      f" helper0 per own chunk: start {c[8]*9/nchk:.0f} blobB {c[9]*9/nchk:.0f} early {c[10]*9/nchk:.0f} waitprog {c[11]*9/nchk:.0f} late {c[12]*9/nchk:.0f}", flush=True)
    info = (capi.C.c_uint64 * 16)()
    s._L.rcg_debug_blocked_info(s._h, 0, info)
    lv = np.zeros(int(info[11]) * 10, np.uint64)
    s._L.rcg_debug_blocked_copy(s._h, 0, 9, lv.ctypes.data_as(capi.C.c_void_p), lv.nbytes)
    print("      fwd plan [depth first count capA capB SA SB groups helpers smem]:", lv.reshape(-1, 10).tolist(),
          "bytesA", int(info[4]), "bytesB", int(info[5]), "far", int(info[6]), flush=True)
    s.close()
