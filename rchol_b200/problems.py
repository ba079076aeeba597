"""Synthetic problem generators for the PCG hot path (host side, numpy).

* ``laplace_3d(n)`` restates the reference generator /root/reference/c++/util/laplace_3d.hpp:8-65
  (7-point Dirichlet Laplacian, lexicographic ``idx = k + j*n + i*n*n``, diagonal 6, off-diagonals -1,
  columns ascending inside each row) as vectorised numpy so that 256^3 is generated in seconds.
* ``aniso_2d(n)`` is the 2D anisotropic random-weight SDDM of BASELINE.json configs[3]; it is *not* in
  the reference (SURVEY.md section 8d gives the spec).
* ``sdd_3d`` / ``sdd_to_sddm`` / ``sdd_rhs`` / ``sdd_recover`` restate the SDD front-end that only the reference's MATLAB
  binding has (matlab/rchol/sdd_to_sddm.m, sdd_3d.m, ex_sdd.m).
* ``reorder_matrix`` / ``reorder_vector`` restate /root/reference/c++/util/util.cpp:16-57 and
  util.hpp:147-155 (A(P,P) with re-sorted rows, b(P)).

All matrices are returned as ``(rowPtr, colIdx, val)`` with ``uint64`` indices and ``float64`` values,
i.e. the exact memory layout of the reference's ``SparseCSR`` (sparse.hpp:10-31).
"""
from __future__ import annotations

import numpy as np


def laplace_3d(n: int):
    n = int(n)
    N = n * n * n
    idx = np.arange(N, dtype=np.int64)
    k = idx % n
    j = (idx // n) % n
    i = idx // (n * n)
    n2 = n * n
    # slot order inside a row follows the reference push_back order (ascending column)
    offs = np.array([-n2, -n, -1, 0, 1, n, n2], dtype=np.int64)
    valid = np.stack([i > 0, j > 0, k > 0, np.ones(N, bool), k < n - 1, j < n - 1, i < n - 1], axis=1)
    cols = idx[:, None] + offs[None, :]
    vals = np.where(offs == 0, 6.0, -1.0)
    vals = np.broadcast_to(vals, (N, 7))
    rowlen = valid.sum(axis=1)
    rowPtr = np.zeros(N + 1, dtype=np.uint64)
    np.cumsum(rowlen, out=rowPtr[1:])
    colIdx = cols[valid].astype(np.uint64)
    val = np.ascontiguousarray(vals[valid], dtype=np.float64)
    return rowPtr, colIdx, val


def aniso_2d(n: int, seed: int = 12345, a_x: float = 1.0, a_y: float = 0.01):
    """5-point grid graph with edge weight ``a_dir * 10**U(-1,1)``; boundary (ghost) edges are drawn the
    same way and only added to the diagonal, so the matrix is a nonsingular SDDM."""
    n = int(n)
    rng = np.random.default_rng(seed)
    wx = a_x * 10.0 ** rng.uniform(-1.0, 1.0, size=(n, n + 1))   # wx[r, c] joins (r,c-1)-(r,c)
    wy = a_y * 10.0 ** rng.uniform(-1.0, 1.0, size=(n + 1, n))   # wy[r, c] joins (r-1,c)-(r,c)
    N = n * n
    r = np.repeat(np.arange(n, dtype=np.int64), n)
    c = np.tile(np.arange(n, dtype=np.int64), n)
    idx = r * n + c
    w_up = wy[r, c]
    w_left = wx[r, c]
    w_right = wx[r, c + 1]
    w_down = wy[r + 1, c]
    diag = w_up + w_left + w_right + w_down
    cols = np.stack([idx - n, idx - 1, idx, idx + 1, idx + n], axis=1)
    vals = np.stack([-w_up, -w_left, diag, -w_right, -w_down], axis=1)
    valid = np.stack([r > 0, c > 0, np.ones(N, bool), c < n - 1, r < n - 1], axis=1)
    rowlen = valid.sum(axis=1)
    rowPtr = np.zeros(N + 1, dtype=np.uint64)
    np.cumsum(rowlen, out=rowPtr[1:])
    colIdx = cols[valid].astype(np.uint64)
    val = np.ascontiguousarray(vals[valid], dtype=np.float64)
    return rowPtr, colIdx, val


def random_rhs(N: int, seed: int = 2024) -> np.ndarray:
    """b ~ U(0,1) (the reference's ``rand`` is unseeded, util.hpp:49-55; we fix the seed)."""
    return np.random.default_rng(seed).random(int(N))


def reorder_vector(x: np.ndarray, P: np.ndarray) -> np.ndarray:
    """xp[i] = x[P[i]]  (util.hpp:147-155)."""
    return np.ascontiguousarray(x[P.astype(np.int64)])


def unpermute_vector(xp: np.ndarray, P: np.ndarray) -> np.ndarray:
    """Inverse of ``reorder_vector``: y[P[i]] = xp[i] (python/ex_laplace_parallel.py:31-32)."""
    y = np.empty_like(xp)
    y[P.astype(np.int64)] = xp
    return y


def reorder_matrix(rowPtr, colIdx, val, P):
    """B = A(P,P) with each row re-sorted by new column index (util.cpp:16-49)."""
    P = P.astype(np.int64)
    N = P.shape[0]
    rp = rowPtr.astype(np.int64)
    inv = np.empty(N, dtype=np.int64)
    inv[P] = np.arange(N, dtype=np.int64)
    lens = (rp[1:] - rp[:-1])[P]
    newPtr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(lens, out=newPtr[1:])
    nnz = int(newPtr[-1])
    # source position of every new entry
    row_of = np.repeat(np.arange(N, dtype=np.int64), lens)
    within = np.arange(nnz, dtype=np.int64) - newPtr[row_of]
    src = rp[P][row_of] + within
    newCol = inv[colIdx.astype(np.int64)[src]]
    newVal = val[src]
    # sort inside rows: one global stable sort on (row, col)
    key = row_of * np.int64(N) + newCol if N < (1 << 31) else None
    if key is not None:
        order = np.argsort(key, kind="stable")
    else:  # pragma: no cover - very large N
        order = np.lexsort((newCol, row_of))
    return newPtr.astype(np.uint64), newCol[order].astype(np.uint64), np.ascontiguousarray(newVal[order])


# ------------------------------------------------------------------------------------------------------------
# SDD front-end: only the MATLAB binding of the reference has it
# (/root/reference/matlab/rchol/sdd_to_sddm.m:2-17, sdd_3d.m, ex_sdd.m:12-30)
# ------------------------------------------------------------------------------------------------------------
def sdd_3d(n: int):
    """3D SDD test matrix of matlab/rchol/sdd_3d.m: the 7-point Laplacian with the couplings of one direction flipped
    to +1 (kron(A1,-I) + kron(I,A1) + kron(I,A2) + 4 I), rows sorted.  Positive off-diagonals => not an M-matrix."""
    import scipy.sparse as sp
    e = np.ones(n)
    I = sp.identity(n, format="csr")
    D = sp.diags([-e[:-1], 2 * e, -e[:-1]], [-1, 0, 1], format="csr")
    A1 = sp.kron(D, I)
    A2 = sp.kron(I, D)
    A = (sp.kron(A1, -I) + sp.kron(I, A1) + sp.kron(I, A2) + 4 * sp.identity(n ** 3)).tocsr()
    A.eliminate_zeros()
    A.sort_indices()
    return A.indptr.astype(np.uint64), A.indices.astype(np.uint64), np.ascontiguousarray(A.data, dtype=np.float64)


def sdd_to_sddm(rowPtr, colIdx, val):
    """Ae = [D + Neg, -Pos; -Pos, D + Neg] (sdd_to_sddm.m:2-17): the 2N x 2N SDDM whose solution of Ae xe = [b; -b]
    carries the solution of the SDD system A x = b as x = (xe[:N] - xe[N:]) / 2.  Rows come out sorted by column."""
    rp = rowPtr.astype(np.int64)
    ci = colIdx.astype(np.int64)
    N = rp.shape[0] - 1
    row = np.repeat(np.arange(N, dtype=np.int64), np.diff(rp))
    offdiag_pos = (ci != row) & (val > 0)
    # upper half: row i keeps its diagonal and negative entries in columns [0, N), its positive ones go to column N + c
    # with the sign flipped; the lower half is the mirror image
    col_top = np.where(offdiag_pos, ci + N, ci)
    col_bot = np.where(offdiag_pos, ci, ci + N)
    v = np.where(offdiag_pos, -val, val)
    rows2 = np.concatenate([row, row + N])
    cols2 = np.concatenate([col_top, col_bot])
    vals2 = np.concatenate([v, v])
    order = np.lexsort((cols2, rows2))
    counts = np.bincount(rows2, minlength=2 * N)
    rp2 = np.zeros(2 * N + 1, np.uint64)
    np.cumsum(counts, out=rp2[1:])
    return rp2, cols2[order].astype(np.uint64), np.ascontiguousarray(vals2[order], dtype=np.float64)


def sdd_rhs(b: np.ndarray) -> np.ndarray:
    """be = [b; -b] (ex_sdd.m:13)."""
    return np.concatenate([b, -b])


def sdd_recover(xe: np.ndarray) -> np.ndarray:
    """x = (xe(1:N) - xe(N+1:end)) / 2 (ex_sdd.m:28)."""
    N = xe.shape[0] // 2
    return 0.5 * (xe[:N] - xe[N:])


# ------------------------------------------------------------------------------------------------------------
# on-disk container shared with the C++ front-end (rchol_b200/cxx/io.hpp)
# ------------------------------------------------------------------------------------------------------------
_MAGIC = b"RCHOLB2\x00"


def save_problem(path, A, G, P=None, part=None, b=None):
    """A, G: (rowPtr, colIdx, val) triples; P, part, b optional.  Layout: see rchol_b200/cxx/io.hpp."""
    u64 = lambda a: np.ascontiguousarray(a, dtype="<u8")
    f64 = lambda a: np.ascontiguousarray(a, dtype="<f8")
    N = A[0].shape[0] - 1
    P = np.zeros(0, np.uint64) if P is None else P
    part = np.zeros(0, np.uint64) if part is None else part
    b = np.zeros(0) if b is None else b
    if G[0].shape[0] - 1 != N or P.shape[0] not in (0, N) or b.shape[0] not in (0, N):
        raise ValueError("save_problem: inconsistent sizes")
    with open(path, "wb") as fh:
        fh.write(_MAGIC)
        fh.write(u64([1, N, int(A[0][-1]), int(G[0][-1]), P.shape[0], part.shape[0], b.shape[0]]).tobytes())
        for M in (A, G):
            fh.write(u64(M[0]).tobytes()); fh.write(u64(M[1]).tobytes()); fh.write(f64(M[2]).tobytes())
        fh.write(u64(P).tobytes()); fh.write(u64(part).tobytes()); fh.write(f64(b).tobytes())


def load_problem(path):
    """-> dict(A=(rowPtr, colIdx, val), G=(...), P, part, b) ; absent optional arrays come back as None."""
    with open(path, "rb") as fh:
        if fh.read(8) != _MAGIC:
            raise ValueError(f"{path}: not an rchol_b200 problem file")
        ver, N, nnzA, nnzG, nP, npart, nb = (int(v) for v in np.frombuffer(fh.read(56), dtype="<u8"))
        if ver != 1 or nP not in (0, N) or nb not in (0, N):
            raise ValueError(f"{path}: corrupt header")

        def arr(count, dtype):
            a = np.frombuffer(fh.read(8 * count), dtype=dtype)
            if a.shape[0] != count:
                raise ValueError(f"{path}: truncated file")
            return a.astype(np.uint64 if dtype == "<u8" else np.float64)
        A = (arr(N + 1, "<u8"), arr(nnzA, "<u8"), arr(nnzA, "<f8"))
        G = (arr(N + 1, "<u8"), arr(nnzG, "<u8"), arr(nnzG, "<f8"))
        P, part, b = arr(nP, "<u8"), arr(npart, "<u8"), arr(nb, "<f8")
    return dict(A=A, G=G, P=P if nP else None, part=part if npart else None, b=b if nb else None)


def algorithmic_bytes_per_iteration(N: int, nnzA: int, nnzG: int) -> int:
    """fp64 values, 4-byte column indices and row pointers, 17 vector passes."""
    return 12 * nnzA + 4 * (N + 1) + 2 * (12 * nnzG + 4 * (N + 1)) + 8 * N * 17
