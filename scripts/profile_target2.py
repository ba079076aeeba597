import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rchol_b200 import problems, producer, capi
A = problems.laplace_3d(64); f = producer.factor(*A, threads=8)
b = problems.random_rhs(f.N)
s = capi.Solver(0, use_graph=False, chain_threads=256)
s.set_factor(f.rowPtr, f.colIdx, f.val, f.part)
y = s.trsv(0, b)
