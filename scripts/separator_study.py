"""CPU study of the top separator blocks of the factor (round-2 design input): how dense is a separator's own triangular
block, and how accurate is a solve through explicitly inverted C x C diagonal blocks (C = 32 ... 1024) compared with
substitution?  Reads the problem cache bench.py writes (/tmp/rchol_b200_cache).  Usage: separator_study.py [n=256] [T=8]"""
import os, sys, time
import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import spsolve_triangular
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8
d, _ = bench.build_problem(n, T)
part = d["part"].astype(np.int64)
N = d["G_rp"].shape[0] - 1
U = sp.csr_matrix((d["G_v"], d["G_ci"].astype(np.int64), d["G_rp"].astype(np.int64)), shape=(N, N))
print("part", part.tolist())
rng = np.random.default_rng(1)
for name, bi in (("root separator", len(part) - 2), ("depth-1 separator", len(part) - 3)):
    lo, hi = int(part[bi]), int(part[bi + 1])
    m = hi - lo
    if m <= 0:
        continue
    B = U[lo:hi, lo:hi].tocsr()                  # the block's own upper-triangular part; L = B^T is what the forward solve uses
    L = B.T.tocsr()
    L.sort_indices()
    rl = np.diff(L.indptr)
    rows = np.repeat(np.arange(m), rl)
    dist = rows - L.indices
    print(f"{name}: block {bi}, rows {m}, entries {L.nnz}, per row mean {rl.mean():.1f} max {rl.max()}")
    for w in (32, 128, 256, 512, 1024, 4096, 16384):
        print(f"   entries within {w:6d} rows of the diagonal: {100.0 * (dist < w).mean():5.1f} %   "
              f"(dense band fill {100.0 * (dist < w).sum() / (m * w):5.2f} %)")
    b = rng.standard_normal(m)
    x_ref = spsolve_triangular(L, b, lower=True)
    for Cs in (32, 128, 256, 512, 1024):
        t0 = time.time()
        x = np.zeros(m)
        worst_cond = 0.0
        inblock = 0
        for s in range(0, m, Cs):
            e = min(s + Cs, m)
            D = L[s:e, s:e].toarray()
            inblock += np.count_nonzero(D)
            Dinv = np.linalg.inv(D)                                   # explicit inverse, as the GPU set-up would build it
            if s % (16 * Cs) == 0:
                worst_cond = max(worst_cond, np.linalg.cond(D))
            t = b[s:e] - L[s:e, :s] @ x[:s]
            x[s:e] = Dinv @ t
        err = np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref)
        print(f"   C = {Cs:5d}: hops {-(-m // Cs):6d}, entries inside diagonal blocks {100.0 * inblock / L.nnz:5.1f} %, "
              f"diag-block fill {100.0 * inblock / (m * Cs / 2):5.2f} %, rel. error vs substitution {err:.2e}, "
              f"cond(D) up to {worst_cond:.1e}   [{time.time() - t0:.0f}s]")
