"""Synthetic problem generators for the PCG hot path (host side, numpy).

* ``laplace_3d(n)`` restates the reference generator /root/reference/c++/util/laplace_3d.hpp:8-65
  (7-point Dirichlet Laplacian, lexicographic ``idx = k + j*n + i*n*n``, diagonal 6, off-diagonals -1,
  columns ascending inside each row) as vectorised numpy so that 256^3 is generated in seconds.
* ``aniso_2d(n)`` is the 2D anisotropic random-weight SDDM of BASELINE.json configs[3]; it is *not* in
  the reference (SURVEY.md section 8d gives the spec).
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


def algorithmic_bytes_per_iteration(N: int, nnzA: int, nnzG: int) -> int:
    """fp64 values, 4-byte column indices and row pointers, 17 vector passes."""
    return 12 * nnzA + 4 * (N + 1) + 2 * (12 * nnzG + 4 * (N + 1)) + 8 * N * 17
