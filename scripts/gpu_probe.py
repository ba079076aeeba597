"""Development probe (not a test): parity + phase timings for a few kernel configurations."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rchol_b200 import problems, producer, capi
from oracle import oracle

def relerr(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))

def run(n, T, sweeps, do_pcg_oracle=True):
    print(f"=== lap3d n={n} T={T}", flush=True)
    t = time.time(); A = problems.laplace_3d(n); f = producer.factor(*A, threads=T); print("factor s", time.time() - t, "nnzG", f.nnz, flush=True)
    G = (f.rowPtr, f.colIdx, f.val)
    b = problems.random_rhs(f.N)
    if T > 0:
        A = producer.ref_reorder(*A, f.P); b = problems.reorder_vector(b, f.P)
    part = f.part if T > 0 else None
    t = time.time(); yo = oracle.trsv_forward(*G, b); zo = oracle.trsv_backward(*G, yo); qo = oracle.spmv(*A, b); print("oracle trsv+spmv s", time.time() - t, flush=True)
    first = True
    for (thr, win) in sweeps:
        s = capi.Solver(0, chain_threads=thr, chain_window=win)
        s.set_matrix(*A); s.set_factor(*G, part)
        if first:
            print("  spmv relerr", relerr(s.spmv(b), qo))
        y = s.trsv(capi.TRSV_FORWARD, b); z = s.trsv(capi.TRSV_BACKWARD, yo); zz = s.precond(b)
        print(f"  thr={thr} win={win}: fwd relerr {relerr(y, yo):.2e} bwd relerr {relerr(z, zo):.2e} precond relerr {relerr(zz, zo):.2e}", flush=True)
        s.set_rhs(b)
        rr, it = s.pcg_resident(1e-8, 500)
        st = s.stats()
        ph = [s.time_phase(p, 3) for p in range(4)]
        print(f"     pcg it={it} relres={rr:.3e} solve_ms={st['solve_ms']:.2f} ms/it={st['solve_ms']/max(it,1):.3f} | spmv {ph[0]:.3f} fwd {ph[1]:.3f} bwd {ph[2]:.3f} vec {ph[3]:.3f} ms | upload {st['upload_ms']:.0f} analysis {st['analysis_ms']:.0f} ms launches/it {st['launches_per_iteration']} sm_mhz {s.stats()['chain_sm_mhz']:.0f}", flush=True)
        if first and do_pcg_oracle:
            t = time.time(); o = oracle.pcg(A, b, 1e-8, 500, G); print(f"     oracle pcg it={o['itr']} relres={o['relres']:.3e} s={time.time()-t:.2f} timings={o['timings']}  x relerr {relerr(s.solution(), o['x']):.2e}", flush=True)
        first = False
        s.close()

if __name__ == "__main__":
    os.system("nproc; free -g | head -2; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv")
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    if which == "small":
        run(16, 0, [(128, 0)])
        run(32, 4, [(256, 0)])
        run(64, 8, [(0, 0), (256, 4096), (128, 1024)])
        run(64, 0, [(256, 8192)])
        run(128, 8, [(0, 0), (256, 2048), (512, 4096), (128, 1024)], do_pcg_oracle=False)
    elif which == "big":
        run(256, 8, [(256, 4096), (512, 4096), (512, 2048), (768, 4096)], do_pcg_oracle=False)
