"""Generates the golden fixtures of tests/golden/ by running the REFERENCE here (needs /root/reference):
  * the factor G / permutation P / partition come from the unmodified reference factorization (baseline/),
  * x, relres, itr come from the unmodified reference `pcg` class (oracle/_ref/libpcg_ref.so, real oneMKL kernels),
  * A r, U^-T r and U^-1 U^-T r come from real oneMKL with the reference's descriptors (pcg.cpp:130-159).
Run:  python tests/golden/make_golden.py      (writes *.npz next to this file)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from rchol_b200 import problems, producer  # noqa: E402
from oracle import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, generator, threads (0 = sequential API), tol, maxit
    ("lap3d_8_seq", lambda: problems.laplace_3d(8), 0, 1e-8, 200),
    ("lap3d_12_t4", lambda: problems.laplace_3d(12), 4, 1e-8, 200),
    ("lap3d_10_t8_tol6", lambda: problems.laplace_3d(10), 8, 1e-6, 200),   # the examples' default tolerance
    ("aniso2d_24_t4", lambda: problems.aniso_2d(24), 4, 1e-8, 500),
]


def main():
    for name, gen, threads, tol, maxit in CASES:
        A = gen()
        f = producer.factor(*A, threads=threads, seed=20240)
        b = problems.random_rhs(f.N, seed=2024)
        if threads > 0:
            Ap = producer.ref_reorder(*A, f.P)
            bp = problems.reorder_vector(b, f.P)
        else:
            Ap, bp = A, b
        G = (f.rowPtr, f.colIdx, f.val)
        ref = oracle.reference_pcg(Ap, bp, tol, maxit, G)
        Ar, y, z = oracle.reference_mkl_kernels(Ap, G, bp)
        np.savez_compressed(os.path.join(HERE, name + ".npz"),
                            A_rowPtr=Ap[0], A_colIdx=Ap[1], A_val=Ap[2],
                            G_rowPtr=f.rowPtr, G_colIdx=f.colIdx, G_val=f.val, P=f.P, part=f.part, b=bp,
                            tol=tol, maxit=maxit, ref_x=ref["x"], ref_relres=ref["relres"], ref_itr=ref["itr"],
                            mkl_spmv=Ar, mkl_fwd=y, mkl_precond=z)
        print(name, "N", f.N, "nnzG", f.nnz, "itr", ref["itr"], "relres", ref["relres"])
    # the 3x3 known-answer test of SURVEY.md section 4 (checked against real MKL here as well)
    rp = np.array([0, 2, 4, 5], np.uint64); ci = np.array([0, 1, 1, 2, 2], np.uint64); v = np.array([2, 1, 3, 1, 4.0])
    b = np.array([1.0, 2.0, 3.0])
    Ar, y, z = oracle.reference_mkl_kernels((rp, ci, v), (rp, ci, v), b)
    np.savez_compressed(os.path.join(HERE, "kat3.npz"), rowPtr=rp, colIdx=ci, val=v, b=b, mkl_spmv=Ar, mkl_fwd=y, mkl_precond=z)
    print("kat3", Ar, y, z)


if __name__ == "__main__":
    main()
