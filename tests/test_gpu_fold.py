"""GPU tests of the folded chain and of the warp-per-block levels (rchol_b200/csrc/rcg_fold.cuh), through the C ABI.
(chain_mode 5 = folded chain + warp-per-block levels; the default, chain_mode 0, pairs the round-1 chain with the
warp-per-block levels: test_default_mode_* below.)

The recent entries of every 32-row chunk are folded with the inverse of the chunk's diagonal block into a dense panel at
set-up; the chain's hop is one panel apply by one warp, the mat-vec with the inverse moves to the near helpers.
1. The layout the set-up kernels build on the device equals the Python restatement (tests/blocked_reference.py) and,
   replayed on the host (tests/blocked_emulator.py), solves to the 1e-12 gate: the set-up in isolation.
2. The solve kernel against the oracle for fold depths / windows that force every entry class, blocks that are not a
   multiple of 32 rows, more blocks than chain CTAs, leaves longer than the window; reruns bit-identical; PCG iteration
   counts equal to the oracle's.
"""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden, make_problem, needs_producer, relerr
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
@pytest.mark.parametrize("kind,n,threads,opts", [("lap3d", 14, 4, dict()), ("lap3d", 20, 0, dict(chain_window=1024, recent=1)),
                                                  ("aniso2d", 64, 8, dict(chain_window=1024, recent=8)),
                                                  ("lap3d", 33, 2, dict(chain_window=1024, early=6)),
                                                  ("lap3d", 40, 8, dict(sep_tile=4, early_sep=12, recent=2)),
                                                  ("lap3d", 14, 4, dict(wb_min=2)), ("lap3d", 40, 8, dict(wb_min=4, recent=2)),
                                                  ("aniso2d", 64, 8, dict(wb_min=1, chain_window=1024)), ("lap3d", 14, 4, dict(wb_min=2, wb_ell=True)),
                                                  ("lap3d", 33, 2, dict(chain_window=1024, early=255)), ("lap3d", 40, 8, dict(sep_window=2048, wb_min=4))])
def test_device_layout_replayed_on_host(capi, oracle, kind, n, threads, opts):
    A, b, G, part, f = make_problem(kind, n, threads)
    yo = oracle.trsv_forward(*G, b)
    zo = oracle.trsv_backward(*G, yo)
    with capi.Solver(0, chain_mode=5, dp_min_rows=-1, **opts) as s:   # (separator levels on the chain: the dense-panel levels have no blobs)
        s.set_factor(*G, part)
        lay_f, lay_b = s.blocked_layout(capi.TRSV_FORWARD), s.blocked_layout(capi.TRSV_BACKWARD)
        assert lay_f["active"] and lay_b["active"] and lay_f["fold"] == 1 and lay_b["fold"] == 1
        kw = dict(Kr=lay_f["Kr"], E=lay_f["E"], Dfar=lay_f["Dfar"], Dfar_sep=lay_f["Dfar_sep"], tile_sep=lay_f["tile_sep"],
                  E_sep=lay_f["E_sep"], fold=True, wb_min=lay_f["wb_min"], Dfar_wb=lay_f["Dfar_wb"], wb_jagged=not opts.get("wb_ell", False))
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
@pytest.mark.parametrize("opts", [dict(), dict(recent=1), dict(recent=2), dict(recent=5), dict(recent=8, early=12), dict(chain_window=1024),
                                  dict(chain_window=2048, recent=1), dict(chain_window=8192), dict(plain_launch=True),
                                  dict(use_graph=False, chain_window=1024), dict(sep_window=4096), dict(chain_window=2048, sep_window=2048),
                                  dict(early=6), dict(early=4, recent=3), dict(capb_quarters=4), dict(slots_a=2), dict(slots_a=8, capb_quarters=5),
                                  dict(sep_tile=8, far_lanes2=32), dict(early_sep=16), dict(early_sep=4, early=5), dict(sep_tile=4, chain_window=1024),
                                  dict(wb_min=1), dict(wb_min=2), dict(wb_min=4, recent=2), dict(wb_min=-1), dict(wb_min=2, use_graph=False),
                                  dict(wb_min=2, wb_ell=True), dict(early=255), dict(early=255, early_sep=255, recent=2),
                                  dict(sep_window=2048), dict(slots_a=5)])
@pytest.mark.parametrize("kind,n,threads", [("lap3d", 40, 8), ("lap3d", 33, 2), ("aniso2d", 160, 4), ("lap3d", 40, 0)])
def test_folded_solve_vs_oracle(capi, oracle, kind, n, threads, opts):
    A, b, G, part, f = make_problem(kind, n, threads)
    yo = oracle.trsv_forward(*G, b)
    zo = oracle.trsv_backward(*G, yo)
    with capi.Solver(0, chain_mode=5, dp_min_rows=-1, **opts) as s:   # (with the dense-panel levels: tests/test_gpu_dense.py)
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
@pytest.mark.parametrize("kind,n,threads,opts", [("lap3d", 48, 256, dict()), ("lap3d", 64, 8, dict()), ("lap3d", 5, 2, dict()), ("aniso2d", 256, 64, dict()),
                                                 ("lap3d", 48, 256, dict(wb_min=-1)), ("lap3d", 64, 8, dict(wb_min=8)),
                                                 ("lap3d", 64, 0, dict(wb_min=1)), ("aniso2d", 256, 64, dict(wb_min=16))])
def test_folded_chain_many_blocks_long_leaves_and_repeatability(capi, oracle, kind, n, threads, opts):
    """Warp-per-block levels (T=256: the leaf level and the lower separator levels by default; forced on / off elsewhere;
    a single 262144-row block walked by one warp, far entries of the own block read back through L2).
    More blocks than chain CTAs (the rings run across block boundaries), leaves longer than the window (far tiles
    inside the own block, publisher back-pressure), blocks shorter than one chunk; bit-identical reruns; PCG converges
    like the oracle."""
    A, b, G, part, f = make_problem(kind, n, threads)
    yo = oracle.trsv_forward(*G, b)
    zo = oracle.trsv_backward(*G, yo)
    with capi.Solver(0, chain_mode=5, **opts) as s:
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


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_folded_chain_on_the_reference_goldens(capi, name):
    """Goldens generated by the unmodified reference pcg + real MKL (tests/golden/make_golden.py)."""
    g = load_golden(name)
    part = g["part"] if len(g["part"]) > 2 else None
    with capi.Solver(0, chain_mode=5) as s:
        s.set_matrix(*g["A"])
        s.set_factor(*g["G"], part)
        assert relerr(s.trsv(capi.TRSV_FORWARD, g["b"]), g["mkl_fwd"]) <= TRSV_TOL
        assert relerr(s.precond(g["b"]), g["mkl_precond"]) <= TRSV_TOL
        x, relres, itr = s.pcg(g["b"], float(g["tol"]), int(g["maxit"]))
        assert abs(itr - int(g["ref_itr"])) <= 1 and relres <= 2 * float(g["tol"])
        if itr == int(g["ref_itr"]):
            assert relerr(x, g["ref_x"]) <= 1e-9


@needs_producer
@pytest.mark.parametrize("kind,n,threads,opts", [("lap3d", 14, 4, dict(wb_min=2)), ("lap3d", 40, 8, dict(wb_min=4)),
                                                 ("aniso2d", 64, 8, dict(wb_min=1, chain_window=1024)), ("lap3d", 33, 2, dict(wb_min=2, wb_ell=True))])
def test_default_mode_layout_replayed_on_host(capi, oracle, kind, n, threads, opts):
    """chain_mode 0: round-1 chain layout for the levels with few blocks, warp-per-block layout for the others."""
    A, b, G, part, f = make_problem(kind, n, threads)
    yo = oracle.trsv_forward(*G, b)
    zo = oracle.trsv_backward(*G, yo)
    with capi.Solver(0, dp_min_rows=-1, **opts) as s:
        s.set_factor(*G, part)
        lay_f, lay_b = s.blocked_layout(capi.TRSV_FORWARD), s.blocked_layout(capi.TRSV_BACKWARD)
        assert lay_f["active"] and lay_f["fold"] == 0 and lay_f["wb_min"] == opts["wb_min"]
        kw = dict(Kr=lay_f["Kr"], E=lay_f["E"], Dfar=lay_f["Dfar"], Dfar_sep=lay_f["Dfar_sep"], tile_sep=lay_f["tile_sep"],
                  E_sep=lay_f["E_sep"], fold=False, wb_min=lay_f["wb_min"], Dfar_wb=lay_f["Dfar_wb"], wb_jagged=not opts.get("wb_ell", False))
        L, bounds, depth = direction_matrix(G, part, False)
        compare_layouts(lay_f, build_layout(L, bounds, depth, False, **kw))
        L, bounds, depth = direction_matrix(G, part, True)
        compare_layouts(lay_b, build_layout(L, bounds, depth, True, reversed_=True, **kw))
        ye, _ = solve_from_layout(lay_f, b, False)
        assert relerr(ye, yo) <= TRSV_TOL
        ze, _ = solve_from_layout(lay_b, yo, True)
        assert relerr(ze, zo) <= TRSV_TOL
        assert relerr(s.trsv(capi.TRSV_FORWARD, b), yo) <= TRSV_TOL
        assert relerr(s.trsv(capi.TRSV_BACKWARD, yo), zo) <= TRSV_TOL


@needs_producer
@pytest.mark.parametrize("kind,n,threads,opts", [("lap3d", 48, 256, dict()), ("lap3d", 64, 8, dict(wb_min=8)), ("lap3d", 40, 8, dict(wb_min=2)),
                                                 ("aniso2d", 256, 64, dict()), ("lap3d", 64, 0, dict(wb_min=1)), ("lap3d", 40, 8, dict(wb_min=1, use_graph=False))])
def test_default_mode_vs_oracle_and_repeatability(capi, oracle, kind, n, threads, opts):
    A, b, G, part, f = make_problem(kind, n, threads)
    yo = oracle.trsv_forward(*G, b)
    zo = oracle.trsv_backward(*G, yo)
    with capi.Solver(0, **opts) as s:
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
