"""CPU tests of the blocked triangular-solve ALGORITHM (chunks of 32 rows, dense inverse of the diagonal blocks,
recent / late / early / far entry classes, shared-memory window ring): the Python restatement of the device layout
(tests/blocked_reference.py) replayed by tests/blocked_emulator.py must reproduce the oracle's triangular solves to the
1e-12 gate of BASELINE.json.  The CUDA kernels are checked against the same restatement in tests/test_gpu_blocked.py."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden, make_problem, needs_producer, relerr
from blocked_reference import direction_matrix, build_layout, tree_depths
from blocked_emulator import solve_from_layout, w_pair_off, unpack_winv


def test_winv_indexing():
    """Full 32x32, [column pair][row] double2: every slot used exactly once, row-major after unpacking."""
    offs = sorted(w_pair_off(p, row) for p in range(16) for row in range(32))
    assert offs == list(range(0, 1024 * 8, 16))
    raw = np.zeros(1024)
    raw[w_pair_off(3, 5) // 8] = 1.0      # W[5][6]
    raw[w_pair_off(3, 5) // 8 + 1] = 2.0  # W[5][7]
    W = unpack_winv(raw)
    assert W[5, 6] == 1.0 and W[5, 7] == 2.0 and np.count_nonzero(W) == 2


def test_tree_depths_follow_the_reference_post_order():
    assert list(tree_depths(1)) == [0]
    assert list(tree_depths(3)) == [1, 1, 0]
    assert list(tree_depths(7)) == [2, 2, 1, 2, 2, 1, 0]


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("kw", [dict(), dict(Kr=1, Dfar=32), dict(fold=True, Kr=3), dict(fold=True, Kr=1, Dfar=32),
                                dict(fold=True, Kr=8, E=10), dict(fold=True, Kr=3, wb_min=2), dict(fold=True, Kr=2, wb_min=1, Dfar_wb=64), dict(wb_min=2), dict(Kr=1, Dfar=32, wb_min=1)])
def test_blocked_algorithm_on_goldens(name, kw):
    g = load_golden(name)
    part = g["part"] if len(g["part"]) > 2 else None
    L, bounds, depth = direction_matrix(g["G"], part, False)
    y, _ = solve_from_layout(build_layout(L, bounds, depth, False, **kw), g["b"], False)
    assert relerr(y, g["mkl_fwd"]) <= 1e-12
    L, bounds, depth = direction_matrix(g["G"], part, True)
    z, _ = solve_from_layout(build_layout(L, bounds, depth, True, reversed_=True, **kw), g["mkl_fwd"], True)
    assert relerr(z, g["mkl_precond"]) <= 1e-12


@needs_producer
def test_blocked_algorithm_with_every_entry_class():
    """One block of 64000 rows in natural order (plane distance 1600 rows) with a 1024-row window: recent, late, early
    and far-local entries all occur."""
    from oracle import oracle
    A, b, G, part, f = make_problem("lap3d", 40, 0)
    yo = oracle.trsv_forward(*G, b)
    L, bounds, depth = direction_matrix(G, part, False)
    lay = build_layout(L, bounds, depth, False, Dfar=32)
    y, st = solve_from_layout(lay, b, False)
    assert relerr(y, yo) <= 1e-12
    assert st["early_tot"] > 0 and st["late_slots"] > 0 and st["rec_slots"] > 0 and lay["tile_need"].max() > 0
    # folded layout: the recent entries become dense panel columns, the inverse moves to the helper's blob
    lay = build_layout(L, bounds, depth, False, Dfar=32, fold=True, Kr=3)
    y, st = solve_from_layout(lay, b, False)
    assert relerr(y, yo) <= 1e-12
    assert st["early_tot"] > 0 and st["late_slots"] > 0 and st["panel_cols"] > 0 and lay["tile_need"].max() > 0


def _solve_with_inverted_diagonal_blocks(L, bounds, b, C):
    """x = L^-1 b block by block (bounds = nested-dissection blocks), every block cut into chunks of C rows whose C x C
    diagonal block is inverted EXPLICITLY (what the GPU set-up stores) and everything left of it applied as a sparse panel."""
    x = np.zeros(L.shape[0])
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        for s in range(int(lo), int(hi), C):
            e = min(s + C, int(hi))
            Dinv = np.linalg.inv(L[s:e, s:e].toarray())
            x[s:e] = Dinv @ (b[s:e] - L[s:e, :s] @ x[:s])
    return x


@needs_producer
@pytest.mark.parametrize("C", [32, 128, 256, 1024])
def test_numerics_gate_of_large_inverted_chunks(C):
    """Gate for the kernels DESIGN.md (g) plans (128/256-row cluster chain, 1024-row dense-inverse separators): solving
    through explicitly inverted C x C diagonal blocks stays inside the 1e-12 bar of BASELINE.json in both directions - the
    factor is strongly diagonally dominant (profiles/r01_separator_study_256_T8.txt: 4e-16 on the real 256^3 factor)."""
    from oracle import oracle
    for kind, n, threads in (("lap3d", 28, 4), ("aniso2d", 160, 4)):
        A, b, G, part, f = make_problem(kind, n, threads)
        yo = oracle.trsv_forward(*G, b)
        zo = oracle.trsv_backward(*G, yo)
        L, bounds, depth = direction_matrix(G, part, False)
        assert relerr(_solve_with_inverted_diagonal_blocks(L, bounds, b, C), yo) <= 1e-13
        Lb, bb, db = direction_matrix(G, part, True)
        zb = _solve_with_inverted_diagonal_blocks(Lb, bb, yo[::-1].copy(), C)
        assert relerr(zb[::-1], zo) <= 1e-13
