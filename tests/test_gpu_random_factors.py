"""Out-of-distribution robustness of the GPU path: factors that do NOT come from rchol (random upper-triangular CSR with
long, dense-ish rows; a diagonal-only factor), no partition (generic schedule), checked against the oracle to the same
1e-12 bar (all five cases pass on the B200)."""
import numpy as np
import pytest

from conftest import relerr
from test_oracle import _random_upper

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,density,seed", [(2, 1.0, 1), (7, 0.0, 2), (400, 0.3, 5), (6000, 0.004, 6), (20000, 0.0008, 7)])
def test_random_upper_triangular_factor_without_partition(n, density, seed):
    from rchol_b200 import capi
    from oracle import oracle
    capi.load()
    rng = np.random.default_rng(seed)
    U = _random_upper(n, density, rng)
    G = (U.indptr.astype(np.uint64), U.indices.astype(np.uint64), U.data.astype(np.float64))
    b = rng.standard_normal(n)
    yo = oracle.trsv_forward(*G, b)
    zo = oracle.trsv_backward(*G, yo)
    with capi.Solver(0) as s:
        s.set_factor(*G, None)
        assert relerr(s.trsv(capi.TRSV_FORWARD, b), yo) <= 1e-12
        assert relerr(s.trsv(capi.TRSV_BACKWARD, yo), zo) <= 1e-12
        assert relerr(s.precond(b), zo) <= 1e-12
        assert s.stats()["watchdog_row"] == 0
