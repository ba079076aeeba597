"""Numerics of bigger dense-inverse chunks: blocked forward/backward solves with C x C explicit
inverses of the diagonal blocks (C = 32 ... 256) against the oracle's sequential substitution.  CPU only."""
import sys, os
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from conftest import make_problem, relerr
from blocked_reference import direction_matrix
from oracle import oracle


def blocked_solve(L, bounds, rhs, C):
    N = L.shape[0]
    x = np.zeros(N)
    Ld = L.tocsr()
    worst = 0.0
    for b in range(len(bounds) - 1):
        lo, hi = int(bounds[b]), int(bounds[b + 1])
        for s in range(lo, hi, C):
            e = min(hi, s + C)
            rows = Ld[s:e]
            D = rows[:, s:e].toarray()
            t = rhs[s:e] - rows[:, :s] @ x[:s]
            n = e - s
            W = np.zeros((n, n))
            for i in range(n):                      # same recurrence as the device set-up (row by row, all columns)
                ei = np.zeros(n); ei[i] = 1.0
                W[i] = (ei - D[i, :i] @ W[:i]) / D[i, i]
            worst = max(worst, np.abs(W).max())
            x[s:e] = W @ t
    return x, worst


if __name__ == "__main__":
    for kind, n, T in (("lap3d", 32, 8), ("lap3d", 40, 0), ("aniso2d", 160, 4)):
        A, b, G, part, f = make_problem(kind, n, T)
        yo = oracle.trsv_forward(*G, b)
        zo = oracle.precond(*G, b)
        for C in (32, 64, 128, 256):
            L, bounds, depth = direction_matrix(G, part, False)
            y, w1 = blocked_solve(L, bounds, b, C)
            L2, bounds2, depth2 = direction_matrix(G, part, True)
            z, w2 = blocked_solve(L2, bounds2, yo[::-1].copy(), C)
            print(f"{kind} n={n} T={T} C={C:3d}: fwd relerr {relerr(y, yo):.2e}  bwd relerr {relerr(z[::-1], zo):.2e}  max|Winv| {max(w1, w2):.2e}", flush=True)
