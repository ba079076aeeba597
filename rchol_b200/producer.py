"""ctypes access to the UNMODIFIED reference factorization (``baseline/_ref/librchol_producer.so``).

BASELINE.json's north star fixes the factorization as an *input* of the hot path: "computed by the
reference rchol on the host with a fixed seed, so the GPU consumes the identical G and permutation".
This module is that input producer.  It is not on the solve path and the CUDA library never links it.

``factor(A, threads=0)``   -> ``rchol(A, G)``            /root/reference/c++/rchol/rchol.cpp:7-28
``factor(A, threads=2^k)`` -> ``rchol(A, G, P, threads)`` /root/reference/c++/rchol/rchol_parallel.cpp:37-92
plus ``part`` = the reference's local ``result_idx`` (rchol_parallel.cpp:64-70) with the ground vertex
dropped (last entry == N), which the C++ API computes but does not return (the Python/MATLAB bindings
do: python/rchol/rchol.py:45).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_LIB_PATH = os.path.join(_ROOT, "baseline", "_ref", "librchol_producer.so")
_lib = None

_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile the producer from /root/reference (only possible where the reference is mounted)."""
    if os.path.exists(_LIB_PATH) and not force:
        return _LIB_PATH
    if not os.path.isdir("/root/reference/c++"):
        raise RuntimeError("reference sources are not mounted and baseline/_ref/librchol_producer.so is absent")
    subprocess.check_call(["make", "-C", os.path.join(_ROOT, "baseline"), "-j8"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def available() -> bool:
    return os.path.exists(_LIB_PATH) or os.path.isdir("/root/reference/c++")


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.refprod_factor.restype = C.c_void_p
        L.refprod_factor.argtypes = [C.c_uint64, _u64p, _u64p, _f64p, C.c_int, C.c_uint, C.c_int]
        L.refprod_error.restype = C.c_char_p
        L.refprod_error.argtypes = [C.c_void_p]
        for name in ("refprod_G_n", "refprod_G_nnz", "refprod_P_len", "refprod_part_len"):
            getattr(L, name).restype = C.c_uint64
            getattr(L, name).argtypes = [C.c_void_p]
        for name in ("refprod_G_rowptr", "refprod_G_colidx", "refprod_P", "refprod_part"):
            getattr(L, name).restype = C.POINTER(C.c_uint64)
            getattr(L, name).argtypes = [C.c_void_p]
        L.refprod_G_val.restype = C.POINTER(C.c_double)
        L.refprod_G_val.argtypes = [C.c_void_p]
        L.refprod_free.argtypes = [C.c_void_p]
        L.refprod_laplace3d.argtypes = [C.c_int, _u64p, _u64p, _f64p]
        L.refprod_reorder.argtypes = [C.c_uint64, _u64p, _u64p, _f64p, _u64p, _u64p, _u64p, _f64p]
        _lib = L
    return _lib


@dataclass
class Factor:
    """G = CSR of the upper-triangular U (diag first, positive); P permutation; part block boundaries."""
    N: int
    rowPtr: np.ndarray
    colIdx: np.ndarray
    val: np.ndarray
    P: np.ndarray        # empty for the sequential API (natural order)
    part: np.ndarray     # len 2T, part[0]=0, part[-1]=N ; blocks in post-order [left.., right.., separator]

    @property
    def nnz(self) -> int:
        return int(self.rowPtr[-1])


def _copy(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).astype(dtype, copy=True)


def factor(rowPtr, colIdx, val, threads: int = 0, seed: int = 20240, quiet: bool = True) -> Factor:
    L = lib()
    N = rowPtr.shape[0] - 1
    h = L.refprod_factor(N, np.ascontiguousarray(rowPtr, np.uint64), np.ascontiguousarray(colIdx, np.uint64),
                         np.ascontiguousarray(val, np.float64), int(threads), int(seed), int(quiet))
    try:
        err = L.refprod_error(h)
        if err:
            raise ValueError(err.decode())
        n = L.refprod_G_n(h)
        nnz = L.refprod_G_nnz(h)
        f = Factor(
            N=int(n),
            rowPtr=_copy(L.refprod_G_rowptr(h), n + 1, np.uint64),
            colIdx=_copy(L.refprod_G_colidx(h), nnz, np.uint64),
            val=_copy(L.refprod_G_val(h), nnz, np.float64),
            P=_copy(L.refprod_P(h), L.refprod_P_len(h), np.uint64),
            part=_copy(L.refprod_part(h), L.refprod_part_len(h), np.uint64),
        )
    finally:
        L.refprod_free(h)
    return f


def ref_laplace_3d(n: int):
    """The reference's own generator, for cross-checking ``problems.laplace_3d``."""
    L = lib()
    N = n ** 3
    nnz = 7 * n ** 3 - 6 * n ** 2
    rp = np.empty(N + 1, np.uint64)
    ci = np.empty(nnz, np.uint64)
    v = np.empty(nnz, np.float64)
    L.refprod_laplace3d(n, rp, ci, v)
    return rp, ci, v


def ref_reorder(rowPtr, colIdx, val, P):
    """The reference's own ``reorder(A, P, B)`` (util.cpp:16-57), for cross-checking and for speed."""
    L = lib()
    N = rowPtr.shape[0] - 1
    rp = np.empty(N + 1, np.uint64)
    ci = np.empty(colIdx.shape[0], np.uint64)
    v = np.empty(colIdx.shape[0], np.float64)
    L.refprod_reorder(N, np.ascontiguousarray(rowPtr, np.uint64), np.ascontiguousarray(colIdx, np.uint64),
                      np.ascontiguousarray(val, np.float64), np.ascontiguousarray(P, np.uint64), rp, ci, v)
    return rp, ci, v
