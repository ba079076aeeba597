import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {k: d[k] for k in d.files}
    if "A_rowPtr" in out:
        out["A"] = (out["A_rowPtr"], out["A_colIdx"], out["A_val"])
        out["G"] = (out["G_rowPtr"], out["G_colIdx"], out["G_val"])
    return out


GOLDEN_CASES = ["lap3d_8_seq", "lap3d_12_t4", "lap3d_10_t8_tol6", "aniso2d_24_t4"]


def relerr(a, b):
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0))


_problem_cache = {}


def make_problem(kind, n, threads, seed=20240):
    """(A_perm, b_perm, G, part, factor) for a seeded synthetic problem; the factor comes from the reference
    factorization (baseline/_ref/librchol_producer.so)."""
    key = (kind, n, threads, seed)
    if key not in _problem_cache:
        from rchol_b200 import problems, producer
        A = problems.laplace_3d(n) if kind == "lap3d" else problems.aniso_2d(n)
        f = producer.factor(*A, threads=threads, seed=seed)
        b = problems.random_rhs(f.N)
        if threads > 0:
            A = producer.ref_reorder(*A, f.P)
            b = problems.reorder_vector(b, f.P)
        _problem_cache[key] = (A, b, (f.rowPtr, f.colIdx, f.val), f.part if threads > 0 else None, f)
    return _problem_cache[key]


def producer_available():
    from rchol_b200 import producer
    return producer.available()


needs_producer = pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "baseline", "_ref", "librchol_producer.so"))
                                    and not os.path.isdir("/root/reference/c++"),
                                    reason="reference producer library not built and /root/reference not mounted")
