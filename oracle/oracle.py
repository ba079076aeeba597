"""ORACLE -- test infrastructure, not product code.

ctypes access to ``oracle/_ref/liboracle.so`` (plain-C restatement of /root/reference/c++/util/pcg.cpp,
source: oracle/pcg_oracle.c) and, when it was built, to ``oracle/_ref/libpcg_ref.so`` (the UNMODIFIED
reference ``pcg`` compiled against genuine oneMKL kernels; see oracle/Makefile).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` legs may
import this module.  The CUDA product path never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "_ref", "liboracle.so")
_REFPCG_SO = os.path.join(_HERE, "_ref", "libpcg_ref.so")

_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")

_lib = None
_reflib = None


def build() -> None:
    """Compile the C restatement (always) and the reference-pcg oracle (only where /root/reference exists)."""
    need = not os.path.exists(_ORACLE_SO) or (os.path.isdir("/root/reference/c++") and not os.path.exists(_REFPCG_SO))
    if need:
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_ORACLE_SO)
        L.oracle_num_threads.restype = C.c_int
        L.oracle_spmv.argtypes = [C.c_uint64, _u64p, _u64p, _f64p, _f64p, _f64p]
        L.oracle_trsv_upper_transposed.argtypes = [C.c_uint64, _u64p, _u64p, _f64p, _f64p, _f64p]
        L.oracle_trsv_upper.argtypes = [C.c_uint64, _u64p, _u64p, _f64p, _f64p, _f64p]
        L.oracle_precond.argtypes = [C.c_uint64, _u64p, _u64p, _f64p, _f64p, _f64p, _f64p]
        L.oracle_pcg.restype = C.c_int
        L.oracle_pcg.argtypes = [C.c_uint64, _u64p, _u64p, _f64p, _f64p, C.c_double, C.c_int, _u64p, _u64p, _f64p,
                                 _f64p, C.POINTER(C.c_double), C.POINTER(C.c_int), _f64p, _f64p]
        _lib = L
    return _lib


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def spmv(rowPtr, colIdx, val, x):
    y = np.empty(rowPtr.shape[0] - 1, np.float64)
    lib().oracle_spmv(y.shape[0], _c(rowPtr, np.uint64), _c(colIdx, np.uint64), _c(val, np.float64), _c(x, np.float64), y)
    return y


def trsv_forward(rowPtr, colIdx, val, b):
    """y = U^{-T} b  (pcg.cpp:151)."""
    y = np.empty(rowPtr.shape[0] - 1, np.float64)
    lib().oracle_trsv_upper_transposed(y.shape[0], _c(rowPtr, np.uint64), _c(colIdx, np.uint64), _c(val, np.float64),
                                       _c(b, np.float64), y)
    return y


def trsv_backward(rowPtr, colIdx, val, y):
    """z = U^{-1} y  (pcg.cpp:155)."""
    z = np.empty(rowPtr.shape[0] - 1, np.float64)
    lib().oracle_trsv_upper(z.shape[0], _c(rowPtr, np.uint64), _c(colIdx, np.uint64), _c(val, np.float64),
                            _c(y, np.float64), z)
    return z


def precond(rowPtr, colIdx, val, r):
    """z = U^{-1} U^{-T} r  (pcg.cpp:141-159)."""
    n = rowPtr.shape[0] - 1
    z = np.empty(n, np.float64)
    scratch = np.empty(n, np.float64)
    lib().oracle_precond(n, _c(rowPtr, np.uint64), _c(colIdx, np.uint64), _c(val, np.float64), _c(r, np.float64), scratch, z)
    return z


def pcg(A, b, tol, maxit, G):
    """Restated reference loop.  ``A`` and ``G`` are (rowPtr, colIdx, val) triples.
    Returns dict(x, relres, itr, timings{trsv,spmv,blas1,loop}, hist)."""
    N = A[0].shape[0] - 1
    x = np.zeros(N, np.float64)
    relres = C.c_double(0)
    itr = C.c_int(0)
    timings = np.zeros(4, np.float64)
    hist = np.zeros(int(maxit) + 1, np.float64)
    rc = lib().oracle_pcg(N, _c(A[0], np.uint64), _c(A[1], np.uint64), _c(A[2], np.float64), _c(b, np.float64),
                          float(tol), int(maxit), _c(G[0], np.uint64), _c(G[1], np.uint64), _c(G[2], np.float64),
                          x, C.byref(relres), C.byref(itr), timings, hist)
    if rc != 0:
        raise MemoryError("oracle_pcg allocation failed")
    return dict(x=x, relres=relres.value, itr=itr.value,
                timings=dict(trsv=timings[0], spmv=timings[1], blas1=timings[2], loop=timings[3]),
                hist=hist[: itr.value + 1].copy())


def have_reference_pcg() -> bool:
    build()
    return os.path.exists(_REFPCG_SO)


def _load_ref():
    global _reflib
    if _reflib is None:
        if not have_reference_pcg():
            raise RuntimeError("oracle/_ref/libpcg_ref.so is not built (needs /root/reference at build time)")
        import torch  # noqa: F401  (makes sure libtorch_cpu.so and its OpenMP runtime are already resident)
        L = C.CDLL(_REFPCG_SO)
        L.refpcg_run.restype = C.c_int
        L.refpcg_run.argtypes = [C.c_uint64, _u64p, _u64p, _f64p, _f64p, C.c_double, C.c_int, _u64p, _u64p, _f64p,
                                 _f64p, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.refmkl_kernels.restype = C.c_int
        L.refmkl_kernels.argtypes = [C.c_uint64, _u64p, _u64p, _f64p, _u64p, _u64p, _f64p, _f64p, _f64p, _f64p, _f64p]
        _reflib = L
    return _reflib


def reference_mkl_kernels(A, G, r):
    """Real oneMKL SpMV and both triangular solves with the reference's descriptors: returns (A r, U^-T r, U^-1 U^-T r)."""
    L = _load_ref()
    N = A[0].shape[0] - 1
    Ar, y, z = (np.zeros(N, np.float64) for _ in range(3))
    L.refmkl_kernels(N, _c(A[0], np.uint64), _c(A[1], np.uint64), _c(A[2], np.float64), _c(G[0], np.uint64),
                     _c(G[1], np.uint64), _c(G[2], np.float64), _c(r, np.float64), Ar, y, z)
    return Ar, y, z


def reference_pcg(A, b, tol, maxit, G):
    """The UNMODIFIED reference ``pcg`` (real MKL SpMV/SpTRSV through libtorch_cpu.so; LP64 => nnz < 2^31)."""
    _load_ref()
    N = A[0].shape[0] - 1
    x = np.zeros(N, np.float64)
    relres = C.c_double(0)
    itr = C.c_int(0)
    _load_ref().refpcg_run(N, _c(A[0], np.uint64), _c(A[1], np.uint64), _c(A[2], np.float64), _c(b, np.float64),
                       float(tol), int(maxit), _c(G[0], np.uint64), _c(G[1], np.uint64), _c(G[2], np.float64),
                       x, C.byref(relres), C.byref(itr))
    L = _load_ref()
    L.rchol_b200_mkl_iteration_seconds.restype = C.c_double
    # seconds inside pcg::iteration alone (marks in mkl_adapter.cpp); the constructor's set-up copies are outside
    return dict(x=x, relres=relres.value, itr=itr.value, iteration_s=float(L.rchol_b200_mkl_iteration_seconds()))
