"""CPU tests of the multi-GPU host logic (rchol_b200/multigpu.py): the sharding plan along the reference's
nested-dissection tree and the exchange pattern (two vector all-reduces over the top separators + scalar all-reduces),
run with world size 2, 4 and 8 over gloo and compared with the oracle's monolithic solve."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from conftest import ROOT, load_golden, relerr
from rchol_b200 import multigpu


def _csr(t, n):
    return sp.csr_matrix((t[2], t[1].astype(np.int64), t[0].astype(np.int64)), shape=(n, n))


@pytest.mark.parametrize("name,ranks", [("lap3d_12_t4", 1), ("lap3d_12_t4", 2), ("lap3d_12_t4", 4), ("lap3d_10_t8_tol6", 8),
                                        ("aniso2d_24_t4", 2)])
def test_plan_and_shard_invariants(name, ranks):
    g = load_golden(name)
    N = g["b"].shape[0]
    pl = multigpu.plan(g["part"], ranks)
    assert pl.top_depth == int(np.log2(ranks)) and len(pl.top_blocks) == ranks - 1
    # subtrees are contiguous, disjoint, in rank order; with the top separators they cover everything
    covered = np.zeros(N, int)
    for lo, hi in pl.sub_range:
        covered[lo:hi] += 1
    for bk, d in pl.top_blocks:
        assert d < pl.top_depth
        covered[int(g["part"][bk]):int(g["part"][bk + 1])] += 1
    assert np.all(covered == 1)
    A, U = _csr(g["A"], N), _csr(g["G"], N)
    p = np.random.default_rng(0).standard_normal(N)
    q = np.zeros(N)
    for r in range(ranks):
        loc = multigpu.shard(pl, r, g["A"], g["G"], g["b"])
        n, ns, idx = loc["index"].shape[0], loc["n_sub"], loc["index"]
        assert loc["bounds"][-1] == n and len(loc["depth"]) == len(loc["bounds"]) - 1
        assert np.array_equal(loc["b"], g["b"][idx])
        Al, Ul = _csr(loc["A"], n), _csr(loc["G"], n)
        # the factor's local rows are complete: U_loc equals the global U restricted to the local rows and columns
        assert abs(Ul - U[idx][:, idx]).max() == 0
        ql = Al @ p[idx]
        q[idx[:ns]] = ql[:ns]              # subtree rows are complete on their owner
        q[idx[ns:]] += ql[ns:]             # top rows: sum over ranks
        # block list: every local block is solved after the deeper ones it couples to
        blk = np.searchsorted(loc["bounds"].astype(np.int64), np.arange(n), side="right") - 1
        coo = Ul.tocoo()
        off = blk[coo.row] != blk[coo.col]
        assert np.all(loc["depth"][blk[coo.col[off]]] < loc["depth"][blk[coo.row[off]]])
    assert relerr(q, A @ p) < 1e-14


def _worker(rank, world, name, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = load_golden(name)
    pl = multigpu.plan(g["part"], world)
    loc = multigpu.shard(pl, rank, g["A"], g["G"], g["b"])
    n, ns = loc["index"].shape[0], loc["n_sub"]
    A, U = _csr(loc["A"], n), _csr(loc["G"], n)
    Uss, Ust, Utt = U[:ns, :ns].tocsr(), U[:ns, ns:].tocsr(), U[ns:, ns:].tocsr()

    def allreduce(x):
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).copy())
        dist.all_reduce(t)
        return t.numpy()

    def dot(a, b):
        lim = n if rank == 0 else ns        # the replicated top rows count once
        return float(allreduce(np.array([a[:lim] @ b[:lim]]))[0])

    def spmv(p):
        q = A @ p
        if n > ns:
            q[ns:] = allreduce(q[ns:])
        return q

    def precond(r):
        y = np.zeros(n)
        y[:ns] = spla.spsolve_triangular(Uss.T.tocsr(), r[:ns], lower=True)
        if n > ns:
            s = allreduce(Ust.T @ y[:ns])                                   # subtree -> top coupling, summed over ranks
            y[ns:] = spla.spsolve_triangular(Utt.T.tocsr(), r[ns:] - s, lower=True)
        z = np.zeros(n)
        if n > ns:
            z[ns:] = spla.spsolve_triangular(Utt, y[ns:], lower=False)
        z[:ns] = spla.spsolve_triangular(Uss, y[:ns] - (Ust @ z[ns:] if n > ns else 0.0), lower=False)
        return z

    # the loop of pcg.cpp:57-127 on the local vectors
    tol, maxit = float(g["tol"]), int(g["maxit"])
    b = loc["b"]
    x, r, p = np.zeros(n), b.copy(), np.zeros(n)
    bb = dot(b, b)
    rz_prev, it = 0.0, 0
    while np.sqrt(dot(r, r)) > np.sqrt(bb) * tol and it < maxit:
        z = precond(r)
        rz = dot(r, z)
        p = z.copy() if it == 0 else z + (rz / rz_prev) * p
        q = spmv(p)
        alpha = dot(p, r) / dot(p, q)
        x += alpha * p
        r -= alpha * q
        rz_prev = rz
        it += 1
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), x=x, index=loc["index"], n_sub=ns, it=it)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,world", [("lap3d_12_t4", 2), ("lap3d_12_t4", 4), ("aniso2d_24_t4", 2), ("lap3d_10_t8_tol6", 8)])
def test_distributed_pcg_over_gloo_matches_the_oracle(tmp_path, name, world):
    import torch.multiprocessing as mp
    from oracle import oracle
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, name, port, str(tmp_path)), nprocs=world, join=True)
    g = load_golden(name)
    pl = multigpu.plan(g["part"], world)
    pieces, its = [], set()
    for r in range(world):
        d = np.load(tmp_path / f"rank{r}.npz")
        pieces.append((d["x"], d["index"], int(d["n_sub"])))
        its.add(int(d["it"]))
    x = multigpu.assemble(pl, pieces)
    o = oracle.pcg(g["A"], g["b"], float(g["tol"]), int(g["maxit"]), g["G"])
    assert its == {o["itr"]} == {int(g["ref_itr"])}
    assert relerr(x, o["x"]) < 1e-9 and relerr(x, g["ref_x"]) < 1e-9
    # replicas of the top separators agree on every rank
    for r in range(1, world):
        assert np.array_equal(pieces[r][0][pieces[r][2]:], pieces[0][0][pieces[0][2]:])
