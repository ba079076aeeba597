"""SDD front-end through the GPU path: the flow of the reference's matlab/ex_sdd.m:12-30 --
Ae = sdd_to_sddm(A), be = [b; -b], factor Ae, PCG on Ae, x = (xe[:N] - xe[N:]) / 2 -- with the factorization done by the
reference and the solve by the CUDA library, checked against the oracle and against the ORIGINAL SDD system."""
import numpy as np
import pytest

from conftest import needs_producer, relerr

pytestmark = pytest.mark.gpu


@needs_producer
@pytest.mark.parametrize("n,threads", [(12, 0), (16, 4), (20, 8)])
def test_sdd_system_solved_through_the_extended_sddm(n, threads):
    import scipy.sparse as sp
    from oracle import oracle
    from rchol_b200 import capi, problems, producer
    A = problems.sdd_3d(n)
    N = A[0].shape[0] - 1
    As = sp.csr_matrix((A[2], A[1].astype(np.int64), A[0].astype(np.int64)), shape=(N, N))
    b = problems.random_rhs(N)
    Ae, be = problems.sdd_to_sddm(*A), problems.sdd_rhs(b)
    f = producer.factor(*Ae, threads=threads, seed=5)
    G = (f.rowPtr, f.colIdx, f.val)
    tol = 1e-8
    with capi.Solver(0) as s:
        if threads > 0:
            s.set_matrix_permuted(*Ae, f.P)                       # reorder(Ae, P) on the device
            s.set_factor(*G, f.part)
            xe, relres, itr = s.pcg_original(be, tol, 500)        # be(P) and the un-permutation on the device
            Ap, bp = producer.ref_reorder(*Ae, f.P), problems.reorder_vector(be, f.P)
        else:
            s.set_matrix(*Ae)
            s.set_factor(*G, None)
            xe, relres, itr = s.pcg(be, tol, 500)
            Ap, bp = Ae, be
        zo = oracle.precond(*G, bp)
        assert relerr(s.precond(bp), zo) <= 1e-12
    o = oracle.pcg(Ap, bp, tol, 500, G)
    assert abs(itr - o["itr"]) <= 1 and relres <= 2 * tol
    x = problems.sdd_recover(xe)
    assert np.linalg.norm(b - As @ x) / np.linalg.norm(b) <= 4 * tol                 # ex_sdd.m:29 "Verify residual"
    assert np.linalg.norm(xe[:N] + xe[N:]) <= 1e-6 * np.linalg.norm(xe)              # xe = [x; -x] up to the tolerance
