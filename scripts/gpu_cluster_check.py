"""GPU check of the cluster chain (chain_mode 4) against the oracle and the 32-row chain: scripts/gpu_cluster_check.py [cases]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_problem, relerr
from rchol_b200 import capi
from oracle import oracle

cases = [("lap3d", 24, 4, {}), ("lap3d", 40, 8, {}), ("lap3d", 40, 0, dict(chain_window=1024)), ("aniso2d", 160, 4, {}),
         ("lap3d", 33, 2, dict(chain_window=2048)), ("lap3d", 48, 256, {})]
big = [a for a in sys.argv[1:] if a.isdigit()]
for kind, n, T, opts in cases:
    t0 = time.time()
    A, b, G, part, f = make_problem(kind, n, T)
    yo = oracle.trsv_forward(*G, b)
    zo = oracle.trsv_backward(*G, yo)
    try:
        with capi.Solver(0, chain_mode=4, **opts) as s:
            s.set_matrix(*A)
            s.set_factor(*G, part)
            for rep in range(2):
                e1 = relerr(s.trsv(capi.TRSV_FORWARD, b), yo)
                e2 = relerr(s.trsv(capi.TRSV_BACKWARD, yo), zo)
                e3 = relerr(s.precond(b), zo)
            x, relres, itr = s.pcg(b, 1e-8, 500)
            o = oracle.pcg(A, b, 1e-8, 500, G)
            print(f"{kind} {n} T={T} {opts}: fwd {e1:.2e} bwd {e2:.2e} precond {e3:.2e} | pcg itr {itr} (oracle {o['itr']}) relres {relres:.2e} | {time.time()-t0:.1f}s", flush=True)
    except Exception as e:
        print(f"{kind} {n} T={T} {opts}: FAILED {e}", flush=True)
for n in big:
    n = int(n)
    A, b, G, part, f = make_problem("lap3d", n, 8)
    for mode, kw in ((0, {}), (4, {}), (4, dict(chain_window=2048))):
        try:
            with capi.Solver(0, chain_mode=mode, **kw) as s:
                s.set_matrix(*A); s.set_factor(*G, part)
                z = s.precond(b)
                line = []
                for direction, dn in ((capi.TRSV_FORWARD, "fwd"), (capi.TRSV_BACKWARD, "bwd")):
                    for gi, g in enumerate(s.groups(direction)):
                        line.append(f"{dn}{gi}[{g['blocks']}b,{g['rows']}r] {s.time_group(direction, gi, 0, 3):.3f}")
                x, relres, itr = s.pcg(b, 1e-8, 500)
                st = s.stats()
                print(f"lap3d {n} T=8 mode {mode} {kw}: itr {itr} relres {relres:.2e} solve {st['solve_ms']:.1f} ms | " + " ".join(line), flush=True)
                if mode == 0: z0 = z
                else: print("   precond mode4 vs mode0:", relerr(z, z0))
            if mode == 4:
                for dbg in (1, 3):
                    with capi.Solver(0, chain_mode=4, dbg=dbg, **kw) as s:
                        s.set_matrix(*A); s.set_factor(*G, part)
                        for direction, dn in ((capi.TRSV_FORWARD, "fwd"), (capi.TRSV_BACKWARD, "bwd")):
                            gi = 0 if direction == capi.TRSV_FORWARD else len(s.groups(direction)) - 1
                            ms = s.time_group(direction, gi, 0, 1)
                            c = s.counters()
                            hops = max(c[10], 1)
                            names = ["stage wait", "loads", "told wait", "x wait", "recent", "matvec", "reduce+send"]
                            print(f"   rank {0 if dbg == 1 else 3} {dn} leaf level {ms:.3f} ms, {hops} hops, NS {c[2]}: " + ", ".join(f"{nm} {c[3 + i] / hops:.0f}" for i, nm in enumerate(names))
                                  + f" | helper: stage {c[11] / hops:.0f}, x {c[12] / hops:.0f}, gather {c[13] / hops:.0f} | producer: tiles {c[14] / hops:.0f}, slot {c[15] / hops:.0f}", flush=True)
        except Exception as e:
            print(f"lap3d {n} mode {mode}: FAILED {e}", flush=True)
