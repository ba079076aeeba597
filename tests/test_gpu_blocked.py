"""GPU tests of the blocked-inverse triangular solve (rchol_b200/csrc/rcg_blocked.cu), through the C ABI.

1. The layout the set-up kernels build on the device, replayed on the host (tests/blocked_emulator.py), solves the
   system to the 1e-12 gate: checks the set-up in isolation.
2. The solve kernel itself against the oracle for windows / recent-distances that force every entry class
   (recent, late, early, far-local, other blocks) and blocks that are not a multiple of 32 rows.
"""
import numpy as np
import pytest

from conftest import make_problem, needs_producer, relerr
from blocked_emulator import solve_from_layout
from blocked_reference import direction_matrix, build_layout, compare_layouts

pytestmark = pytest.mark.gpu
TRSV_TOL = 1e-12


@pytest.fixture(scope="module")
def capi():
    from rchol_b200 import capi as m
    m.load()
    return m


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle as o
    return o


@needs_producer
@pytest.mark.parametrize("kind,n,threads,opts", [("lap3d", 14, 4, dict()), ("lap3d", 20, 0, dict(chain_window=1024)),
                                                  ("aniso2d", 64, 8, dict(chain_window=1024, recent=1)),
                                                  ("lap3d", 33, 2, dict(chain_window=1024)),
                                                  ("lap3d", 40, 0, dict(chain_window=1024)),
                                                  ("lap3d", 33, 2, dict(chain_window=1024, early=6)),
                                                  ("lap3d", 40, 8, dict(sep_tile=4, early_sep=12)), ("aniso2d", 96, 4, dict(sep_tile=1))])
def test_device_layout_replayed_on_host(capi, oracle, kind, n, threads, opts):
    A, b, G, part, f = make_problem(kind, n, threads)
    yo = oracle.trsv_forward(*G, b)
    zo = oracle.trsv_backward(*G, yo)
    with capi.Solver(0, chain_mode=6, **opts) as s:   # the round-1 blocked chain (the default adds warp-per-block levels: test_gpu_fold.py)
        s.set_factor(*G, part)
        lay_f, lay_b = s.blocked_layout(capi.TRSV_FORWARD), s.blocked_layout(capi.TRSV_BACKWARD)
        assert lay_f["active"] and lay_b["active"] and lay_f["fold"] == 0
        kw = dict(Kr=lay_f["Kr"], E=lay_f["E"], Dfar=lay_f["Dfar"], Dfar_sep=lay_f["Dfar_sep"], tile_sep=lay_f["tile_sep"], E_sep=lay_f["E_sep"])
        L, bounds, depth = direction_matrix(G, part, False)
        compare_layouts(lay_f, build_layout(L, bounds, depth, False, **kw))
        L, bounds, depth = direction_matrix(G, part, True)
        compare_layouts(lay_b, build_layout(L, bounds, depth, True, reversed_=True, **kw))
        ye, st_f = solve_from_layout(lay_f, b, False)
        assert relerr(ye, yo) <= TRSV_TOL
        ze, st_b = solve_from_layout(lay_b, yo, True)
        assert relerr(ze, zo) <= TRSV_TOL
        assert relerr(s.trsv(capi.TRSV_FORWARD, b), yo) <= TRSV_TOL
        assert relerr(s.trsv(capi.TRSV_BACKWARD, yo), zo) <= TRSV_TOL


@needs_producer
@pytest.mark.parametrize("opts", [dict(), dict(recent=1), dict(chain_window=1024), dict(chain_window=8192), dict(plain_launch=True),
                                  dict(use_graph=False, chain_window=1024), dict(chain_mode=1), dict(chain_mode=3), dict(early=3, recent=2),
                                  dict(sep_tile=8, far_lanes2=32), dict(chain_mode=4), dict(chain_mode=4, use_graph=False, chain_window=2048)])
@pytest.mark.parametrize("kind,n,threads", [("lap3d", 40, 8), ("lap3d", 33, 2), ("aniso2d", 160, 4), ("lap3d", 40, 0)])
def test_blocked_solve_vs_oracle(capi, oracle, kind, n, threads, opts):
    A, b, G, part, f = make_problem(kind, n, threads)
    yo = oracle.trsv_forward(*G, b)
    zo = oracle.trsv_backward(*G, yo)
    with capi.Solver(0, **dict(dict(chain_mode=6), **opts)) as s:   # round-1 chains, kept selectable
        s.set_matrix(*A)
        s.set_factor(*G, part)
        for _ in range(2):   # twice: flags and progress counters are reset per solve
            assert relerr(s.trsv(capi.TRSV_FORWARD, b), yo) <= TRSV_TOL
            assert relerr(s.trsv(capi.TRSV_BACKWARD, yo), zo) <= TRSV_TOL
            assert relerr(s.precond(b), zo) <= TRSV_TOL
        x, relres, itr = s.pcg(b, 1e-8, 500)
        o = oracle.pcg(A, b, 1e-8, 500, G)
        assert abs(itr - o["itr"]) <= 1 and relres <= 2e-8


@needs_producer
def test_many_leaves_more_blocks_than_chain_ctas(capi, oracle):
    """T=256 leaves: 511 blocks, more than the 74 chain CTAs of a launch, so every chain CTA walks several blocks and
    the staging rings run across block boundaries."""
    A, b, G, part, f = make_problem("lap3d", 48, 256)
    yo = oracle.trsv_forward(*G, b)
    zo = oracle.trsv_backward(*G, yo)
    with capi.Solver(0, chain_mode=6) as s:
        s.set_factor(*G, part)
        assert relerr(s.trsv(capi.TRSV_FORWARD, b), yo) <= TRSV_TOL
        assert relerr(s.precond(b), zo) <= TRSV_TOL


@needs_producer
@pytest.mark.parametrize("kind,n,threads", [("lap3d", 48, 256), ("lap3d", 64, 8), ("lap3d", 5, 2)])
def test_cluster_chain_vs_oracle_and_32_row_chain(capi, oracle, kind, n, threads):
    """chain_mode 4 (rcg_cluster.cuh): leaf blocks on the cluster chain (128-row chunks, 4 CTAs per leaf, DSMEM exchange).
    Many small leaves (every cluster walks several blocks), leaves longer than the window (far tiles inside the own
    block, publisher back-pressure), and blocks shorter than one chunk; PCG with the fused r.z must converge like the
    oracle, and reruns must be bit-identical."""
    A, b, G, part, f = make_problem(kind, n, threads)
    yo = oracle.trsv_forward(*G, b)
    zo = oracle.trsv_backward(*G, yo)
    with capi.Solver(0, chain_mode=4) as s:
        s.set_matrix(*A)
        s.set_factor(*G, part)
        assert relerr(s.trsv(capi.TRSV_FORWARD, b), yo) <= TRSV_TOL
        assert relerr(s.trsv(capi.TRSV_BACKWARD, yo), zo) <= TRSV_TOL
        z1 = s.precond(b)
        z2 = s.precond(b)
        assert relerr(z1, zo) <= TRSV_TOL and np.array_equal(z1, z2)
        x, relres, itr = s.pcg(b, 1e-8, 500)
        x2, relres2, itr2 = s.pcg(b, 1e-8, 500)
        o = oracle.pcg(A, b, 1e-8, 500, G)
        assert abs(itr - o["itr"]) <= 1 and relres <= 2e-8
        assert itr2 == itr and np.array_equal(x, x2)
