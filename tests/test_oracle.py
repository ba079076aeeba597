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
