"""Small profiling target: a few PCG iterations on lap3d n^3 / T leaves (for ncu launch lists and full captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rchol_b200 import problems, producer, capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8
its = int(sys.argv[3]) if len(sys.argv) > 3 else 3
A = problems.laplace_3d(n); f = producer.factor(*A, threads=T)
b = problems.random_rhs(f.N)
if T > 0:
    A = producer.ref_reorder(*A, f.P); b = problems.reorder_vector(b, f.P)
s = capi.Solver(0, use_graph=False)
s.set_matrix(*A); s.set_factor(f.rowPtr, f.colIdx, f.val, f.part if T > 0 else None); s.set_rhs(b)
relres, itr = s.pcg_resident(1e-8, its)
print("iterations", itr, "relres", relres, "solve_ms", s.stats()["solve_ms"], "launches/it", s.stats()["launches_per_iteration"])
