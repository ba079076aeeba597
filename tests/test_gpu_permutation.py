"""GPU parity of the permutation steps either side of the hot path, through the C ABI:
reorder(A, P, B) of the reference (/root/reference/c++/util/util.cpp:16-57), reorder(x, P, xp) (util.hpp:147-155) and the
un-permutation of the solution (python/ex_laplace_parallel.py:31-32).  Integer/index work: the bar is bit-exact."""
import numpy as np
import pytest

from conftest import load_golden, make_problem, needs_producer

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from rchol_b200 import capi as m
    m.load()
    return m


def _same_csr(a, b):
    return all(np.array_equal(np.asarray(x), np.asarray(y)) for x, y in zip(a, b))


@pytest.mark.parametrize("name,n", [("lap3d_12_t4", 12), ("lap3d_10_t8_tol6", 10)])
def test_device_reorder_matches_the_reference_goldens(capi, name, n):
    """A(P,P) built on the device == the matrix the reference's reorder() produced when the golden was generated."""
    from rchol_b200 import problems
    g = load_golden(name)
    A0 = problems.laplace_3d(n)
    with capi.Solver(0) as s:
        s.set_matrix_permuted(*A0, g["P"])
        B = s.get_matrix()
    assert _same_csr(B, g["A"])


@needs_producer
@pytest.mark.parametrize("kind,n,threads", [("lap3d", 24, 8), ("aniso2d", 96, 4), ("lap3d", 5, 2)])
def test_device_reorder_vs_live_reference_and_host_restatement(capi, kind, n, threads):
    from rchol_b200 import problems, producer
    A0 = problems.laplace_3d(n) if kind == "lap3d" else problems.aniso_2d(n)
    f = producer.factor(*A0, threads=threads, seed=7)
    for P in (f.P, np.random.default_rng(3).permutation(f.N).astype(np.uint64), np.arange(f.N, dtype=np.uint64)):
        with capi.Solver(0) as s:
            s.set_matrix_permuted(*A0, P)
            B = s.get_matrix()
        assert _same_csr(B, producer.ref_reorder(*A0, P))          # the reference's own reorder, compiled unmodified
        assert _same_csr(B, problems.reorder_matrix(*A0, P))       # numpy restatement
        for i in range(min(f.N, 50)):                              # rows sorted by column
            r = B[1][int(B[0][i]):int(B[0][i + 1])]
            assert np.all(r[1:] > r[:-1])


def test_vector_permutations_round_trip(capi):
    rng = np.random.default_rng(11)
    for N in (1, 2, 33, 4097, 100003):
        P = rng.permutation(N).astype(np.uint64)
        x = rng.standard_normal(N)
        with capi.Solver(0) as s:
            s.set_permutation(P)
            xp = s.permute(x)
            assert np.array_equal(xp, x[P.astype(np.int64)])                 # xp[i] = x[P[i]]
            y = s.unpermute(xp)
            assert np.array_equal(y, x)                                       # y[P[i]] = xp[i]
            z = np.empty(N); z[P.astype(np.int64)] = x
            assert np.array_equal(s.unpermute(x), z)


@needs_producer
@pytest.mark.parametrize("kind,n,threads", [("lap3d", 20, 4), ("aniso2d", 64, 8)])
def test_solve_in_the_original_ordering(capi, kind, n, threads):
    """ex_laplace_parallel.py end to end: A, b in the original ordering in, y in the original ordering out."""
    import scipy.sparse as sp
    from rchol_b200 import problems
    Ap, bp, G, part, f = make_problem(kind, n, threads)
    A0 = problems.laplace_3d(n) if kind == "lap3d" else problems.aniso_2d(n)
    b0 = problems.random_rhs(f.N)
    with capi.Solver(0) as s:
        s.set_matrix_permuted(*A0, f.P)
        s.set_factor(*G, part)
        y, relres, itr = s.pcg_original(b0, 1e-8, 500)
        xp, relres_p, itr_p = s.pcg(bp, 1e-8, 500)                           # the permuted-ordering entry point
    assert itr == itr_p and relres == relres_p
    assert np.array_equal(y, problems.unpermute_vector(xp, f.P))               # bit-identical to permuting on the host
    A0s = sp.csr_matrix((A0[2], A0[1].astype(np.int64), A0[0].astype(np.int64)), shape=(f.N, f.N))
    true_rel = np.linalg.norm(A0s @ y - b0) / np.linalg.norm(b0)               # python/ex_laplace_parallel.py:33
    assert true_rel <= 2e-8 and abs(true_rel - relres) <= 1e-10


def test_permutation_errors(capi):
    from rchol_b200 import problems
    A0 = problems.laplace_3d(4)
    N = 64
    with capi.Solver(0) as s:
        dup = np.arange(N, dtype=np.uint64); dup[5] = 6
        with pytest.raises(capi.RcgError) as e:
            s.set_matrix_permuted(*A0, dup)
        assert e.value.code == capi.RCG_ERR_INVALID
        big = np.arange(N, dtype=np.uint64); big[0] = N
        with pytest.raises(capi.RcgError) as e:
            s.set_permutation(big)
        assert e.value.code == capi.RCG_ERR_INVALID
        s.set_matrix(*A0)
        with pytest.raises(capi.RcgError) as e:                                 # no permutation recorded
            s.permute(np.zeros(N))
        assert e.value.code == capi.RCG_ERR_STATE
        with pytest.raises(capi.RcgError):                                      # wrong length
            s.set_permutation(np.arange(N + 1, dtype=np.uint64))
