"""CPU tests of the oracle (oracle/pcg_oracle.c): pinned against golden vectors that were produced by the
reference itself (tests/golden/make_golden.py: unmodified reference pcg + real oneMKL kernels)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden, relerr
from oracle import oracle


def test_kat_3x3():
    # U = [2 1 0; 0 3 1; 0 0 4], b = [1, 2, 3]
    k = load_golden("kat3")
    G = (k["rowPtr"], k["colIdx"], k["val"])
    np.testing.assert_allclose(oracle.trsv_forward(*G, k["b"]), [0.5, 0.5, 0.625], rtol=0, atol=1e-15)
    np.testing.assert_allclose(oracle.precond(*G, k["b"]), [0.19270833333333334, 0.11458333333333333, 0.15625], rtol=1e-15)
    np.testing.assert_allclose(oracle.spmv(*G, k["b"]), [4.0, 9.0, 12.0], rtol=0, atol=0)
    # and against what real MKL returned for the same calls
    np.testing.assert_allclose(oracle.trsv_forward(*G, k["b"]), k["mkl_fwd"], rtol=1e-15)
    np.testing.assert_allclose(oracle.precond(*G, k["b"]), k["mkl_precond"], rtol=1e-15)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_kernels_match_mkl_goldens(name):
    g = load_golden(name)
    assert relerr(oracle.spmv(*g["A"], g["b"]), g["mkl_spmv"]) < 1e-14
    y = oracle.trsv_forward(*g["G"], g["b"])
    assert relerr(y, g["mkl_fwd"]) < 1e-13
    assert relerr(oracle.trsv_backward(*g["G"], y), g["mkl_precond"]) < 1e-13
    assert relerr(oracle.precond(*g["G"], g["b"]), g["mkl_precond"]) < 1e-13


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_pcg_matches_reference_goldens(name):
    g = load_golden(name)
    o = oracle.pcg(g["A"], g["b"], float(g["tol"]), int(g["maxit"]), g["G"])
    assert o["itr"] == int(g["ref_itr"])
    assert abs(o["relres"] - float(g["ref_relres"])) <= 1e-3 * float(g["ref_relres"])   # summation order differs
    assert relerr(o["x"], g["ref_x"]) < 1e-12
    # semantics of pcg.cpp:82: the recurrence residual at exit is below tol, the one before is not
    assert o["hist"][-1] <= float(g["tol"]) and (len(o["hist"]) < 2 or o["hist"][-2] > float(g["tol"]))


def test_pcg_maxit_and_zero_iterations():
    g = load_golden("lap3d_8_seq")
    o = oracle.pcg(g["A"], g["b"], 1e-8, 3, g["G"])
    assert o["itr"] == 3 and o["relres"] > 1e-8
    o = oracle.pcg(g["A"], g["b"], 10.0, 50, g["G"])      # ||r0|| = ||b|| is not > 10 ||b||: no iteration at all
    assert o["itr"] == 0 and abs(o["relres"] - 1.0) < 1e-15 and not o["x"].any()


@pytest.mark.skipif(not os.path.isdir("/root/reference/c++"), reason="needs the reference sources to build libpcg_ref.so")
def test_oracle_vs_live_reference_pcg():
    """Where the reference is mounted: run the unmodified reference pcg (real MKL) live on a fresh seeded problem."""
    from conftest import make_problem
    A, b, G, part, f = make_problem("lap3d", 20, 4)
    ref = oracle.reference_pcg(A, b, 1e-8, 300, G)
    o = oracle.pcg(A, b, 1e-8, 300, G)
    assert o["itr"] == ref["itr"]
    assert relerr(o["x"], ref["x"]) < 1e-12
    Ar, y, z = oracle.reference_mkl_kernels(A, G, b)
    assert relerr(oracle.trsv_forward(*G, b), y) < 1e-13 and relerr(oracle.precond(*G, b), z) < 1e-13


def _random_upper(n, density, rng):
    """Random upper-triangular CSR with the layout the reference factor has: diagonal first (> 0), then sorted columns."""
    import scipy.sparse as sp
    M = sp.random(n, n, density=density, random_state=rng, format="csr", data_rvs=lambda k: -rng.uniform(0.05, 1.0, k))
    U = sp.triu(M, k=1).tocsr()
    U.sort_indices()
    rowsum = np.abs(U).sum(axis=1).A1 + np.abs(U).sum(axis=0).A1
    U = (U + sp.diags(rowsum + rng.uniform(0.5, 1.5, n))).tocsr()
    U.sort_indices()                                         # diagonal is the smallest column of an upper-triangular row
    return U


@pytest.mark.parametrize("n,density,seed", [(1, 1.0, 0), (2, 1.0, 1), (7, 0.0, 2), (50, 0.2, 3), (400, 0.02, 4), (400, 0.3, 5)])
def test_triangular_solves_and_spmv_against_scipy(n, density, seed):
    """Third statement of the three kernels (SciPy), independent of MKL and of the C restatement: random upper-triangular
    factors including the edge cases N = 1, a dense 2 x 2, and a diagonal-only factor (empty off-diagonal rows)."""
    from scipy.sparse.linalg import spsolve_triangular
    rng = np.random.default_rng(seed)
    U = _random_upper(n, density, rng)
    G = (U.indptr.astype(np.uint64), U.indices.astype(np.uint64), U.data.astype(np.float64))
    b = rng.standard_normal(n)
    y_ref = spsolve_triangular(U.T.tocsr(), b, lower=True)
    z_ref = spsolve_triangular(U, y_ref, lower=False)
    y = oracle.trsv_forward(*G, b)
    assert relerr(y, y_ref) < 1e-13
    assert relerr(oracle.trsv_backward(*G, y), z_ref) < 1e-13
    assert relerr(oracle.precond(*G, b), z_ref) < 1e-13
    assert relerr(oracle.spmv(*G, b), U @ b) < 1e-14
    # the preconditioner is symmetric positive definite: <u, M^-1 v> = <M^-1 u, v>
    u, v = rng.standard_normal(n), rng.standard_normal(n)
    assert abs(u @ oracle.precond(*G, v) - oracle.precond(*G, u) @ v) <= 1e-12 * (np.linalg.norm(u) * np.linalg.norm(v) + 1)


def test_pcg_with_exact_factor_converges_in_one_iteration():
    """With G = chol(A)^T the preconditioner is A^-1: one iteration, whatever the tolerance (pcg.cpp:82-112)."""
    import scipy.sparse as sp
    from rchol_b200 import problems
    rp, ci, v = problems.laplace_3d(4)
    N = rp.shape[0] - 1
    A = sp.csr_matrix((v, ci.astype(np.int64), rp.astype(np.int64)), shape=(N, N))
    U = sp.csr_matrix(np.triu(np.linalg.cholesky(A.toarray()).T))
    U.eliminate_zeros(); U.sort_indices()
    G = (U.indptr.astype(np.uint64), U.indices.astype(np.uint64), U.data)
    b = problems.random_rhs(N)
    o = oracle.pcg((rp, ci, v), b, 1e-10, 50, G)
    assert o["itr"] == 1 and o["relres"] < 1e-13
    assert relerr(o["x"], np.linalg.solve(A.toarray(), b)) < 1e-12
