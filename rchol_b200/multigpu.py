"""Multi-GPU host logic: sharding of the PCG solve phase along the reference's nested-dissection tree.

Layout: with R = 2^g ranks, rank r owns the depth-g subtree r -- a contiguous range of the
permuted index space, because the reference lays the tree out in post-order [left subtree, right subtree, separator]
(/root/reference/c++/rchol/find_separator.cpp:138-150, rchol_lap.cpp:254-261) -- and a replica of the 2^g - 1
separators above depth g.  The local index space of a rank is [own subtree rows | top separator rows].

This module holds the pure index arithmetic (``plan``, ``shard``) -- tested on CPU with world-size-2 gloo runs against
the oracle (tests/test_multigpu_host.py) -- and the N > 1 leg of bench.py (``bench_main``), where the per-rank solves
run on the GPUs through the C ABI (rcg_dist_init / rcg_set_factor_blocks / rcg_pcg) with NCCL over NVLink.
"""
from __future__ import annotations

import json
import os
import time
from dataclasses import dataclass

import numpy as np


@dataclass
class Plan:
    nranks: int
    top_depth: int                 # g: blocks with depth < g are the replicated top separators
    sub_range: list                # per rank: (lo, hi) global row range of its subtree
    sub_blocks: list               # per rank: list of (global block index, depth) inside the subtree, in index order
    top_blocks: list               # list of (global block index, depth), in index order
    bounds: np.ndarray             # the global block boundaries (= part)


def _tree(nb: int):
    """depth and subtree extent (first block, number of blocks) of every block of a complete post-order tree."""
    depth = np.zeros(nb, np.int32)
    sub = [None] * nb

    def rec(start, total, d):
        if total == 1:
            depth[start] = d
            sub[start] = (start, 1)
        else:
            sep, half = start + total - 1, (total - 1) // 2
            depth[sep] = d
            sub[sep] = (start, total)
            rec(start, half, d + 1)
            rec(start + half, half, d + 1)
    rec(0, nb, 0)
    return depth, sub


def plan(part: np.ndarray, nranks: int) -> Plan:
    part = np.asarray(part, dtype=np.int64)
    nb = part.shape[0] - 1
    if nranks < 1 or (nranks & (nranks - 1)) != 0:
        raise ValueError("the number of ranks must be a power of two")
    T = (nb + 1) // 2
    if nranks > T:
        raise ValueError(f"{nranks} ranks need a partition with at least {nranks} leaves (have {T})")
    g = int(np.log2(nranks))
    depth, sub = _tree(nb)
    roots = [b for b in range(nb) if depth[b] == g]          # subtree roots, in index order = rank order
    assert len(roots) == nranks
    sub_range, sub_blocks = [], []
    for rb in roots:
        first, count = sub[rb]
        sub_range.append((int(part[first]), int(part[first + count])))
        sub_blocks.append([(b, int(depth[b])) for b in range(first, first + count)])
    top_blocks = [(b, int(depth[b])) for b in range(nb) if depth[b] < g]
    return Plan(nranks, g, sub_range, sub_blocks, top_blocks, part)


def _rows(rp, ci, v, ranges):
    """CSR rows of the concatenated index ranges."""
    rp = rp.astype(np.int64)
    lens, cols, vals = [], [], []
    for lo, hi in ranges:
        lens.append(rp[lo + 1:hi + 1] - rp[lo:hi])
        cols.append(ci[rp[lo]:rp[hi]])
        vals.append(v[rp[lo]:rp[hi]])
    lens = np.concatenate(lens) if lens else np.zeros(0, np.int64)
    out_rp = np.zeros(lens.shape[0] + 1, np.int64)
    np.cumsum(lens, out=out_rp[1:])
    return out_rp, (np.concatenate(cols) if cols else np.zeros(0, ci.dtype)), (np.concatenate(vals) if vals else np.zeros(0))


def _filter(rp, cols_local, vals, keep):
    """drop entries where keep is False"""
    row_of = np.repeat(np.arange(rp.shape[0] - 1, dtype=np.int64), np.diff(rp))
    counts = np.bincount(row_of[keep], minlength=rp.shape[0] - 1)
    new_rp = np.zeros(rp.shape[0], np.int64)
    np.cumsum(counts, out=new_rp[1:])
    return new_rp, cols_local[keep], vals[keep]


def shard(pl: Plan, rank: int, A, G, b=None):
    """Local problem of `rank`: dict with A, G (uint64/float64 CSR triples in the LOCAL index space), b, the local
    block list (bounds, depth), n_sub, and `index` = global row of every local row."""
    part = pl.bounds
    N = int(part[-1])
    lo, hi = pl.sub_range[rank]
    top_ranges = [(int(part[bk]), int(part[bk + 1])) for bk, _ in pl.top_blocks]
    ranges = [(lo, hi)] + top_ranges
    index = np.concatenate([np.arange(a, c, dtype=np.int64) for a, c in ranges]) if ranges else np.zeros(0, np.int64)
    n_sub = hi - lo
    g2l = np.full(N, -1, np.int64)
    g2l[index] = np.arange(index.shape[0], dtype=np.int64)

    # A: rows (subtree + top); columns restricted to (own subtree + top); the [top, top] block only on rank 0, so that the
    # sum over ranks of the local products is the global product for the top rows
    rp, ci, v = _rows(A[0], A[1], A[2], ranges)
    cl = g2l[ci.astype(np.int64)]
    keep = cl >= 0
    if rank != 0:
        row_of = np.repeat(np.arange(rp.shape[0] - 1, dtype=np.int64), np.diff(rp))
        keep &= ~((row_of >= n_sub) & (cl >= n_sub))
    a_rp, a_ci, a_v = _filter(rp, cl, v, keep)

    # G = CSR of U: subtree rows couple to the subtree and to the top separators, top rows to the top only
    rp, ci, v = _rows(G[0], G[1], G[2], ranges)
    cl = g2l[ci.astype(np.int64)]
    if (cl < 0).any():
        raise ValueError("the factor couples a subtree to a non-ancestor block: `part` does not describe it")

    bounds, depth = [0], []
    for bk, d in pl.sub_blocks[rank]:
        bounds.append(bounds[-1] + int(part[bk + 1] - part[bk]))
        depth.append(d)
    for bk, d in pl.top_blocks:
        bounds.append(bounds[-1] + int(part[bk + 1] - part[bk]))
        depth.append(d)
    out = dict(A=(a_rp.astype(np.uint64), a_ci.astype(np.uint64), np.ascontiguousarray(a_v)),
               G=(rp.astype(np.uint64), cl.astype(np.uint64), np.ascontiguousarray(v)),
               bounds=np.asarray(bounds, np.uint64), depth=np.asarray(depth, np.int32), n_sub=int(n_sub), index=index)
    if b is not None:
        out["b"] = np.ascontiguousarray(b[index])
    return out


def assemble(pl: Plan, pieces):
    """Global vector from the per-rank local vectors (subtree parts from their owners, top part from rank 0)."""
    N = int(pl.bounds[-1])
    x = np.zeros(N)
    for r, (loc, idx, n_sub) in enumerate(pieces):
        x[idx[:n_sub]] = loc[:n_sub]
        if r == 0:
            x[idx[n_sub:]] = loc[n_sub:]
    return x


# ---------------------------------------------------------------------------------------------------------------------
# N > 1 leg of bench.py
# ---------------------------------------------------------------------------------------------------------------------
def _roofline(value_gbs, world, B_iter, ms_per_iter):
    """Iteration-level roofline of the sharded solve: aggregate algorithmic GB/s against world x the measured HBM peak
    (the dominant kernels are the triangular-solve level launches, as on one GPU; their per-launch figures are in the
    N = 1 line)."""
    import bench as _bench
    peak, src = _bench.measured_peak_gbs()
    return dict(bound="hbm", kernel="PCG iteration, sharded (triangular-solve level launches k_wb_* / k_dp_* dominate: per-launch figures in the N=1 line; "
                       "the 2^g - 1 top separators are solved on every rank)", achieved=value_gbs,
                peak=peak * world, unit="GB/s", frac=value_gbs / (peak * world), traffic=None, peak_source=src + f" x {world} GPUs",
                bytes_per_iteration=B_iter, ms_per_iteration=ms_per_iter)


def bench_main(args, d, B_iter, rank, world, config):
    import torch
    import torch.distributed as dist
    from rchol_b200 import capi

    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    if not dist.is_initialized():
        dist.init_process_group("nccl", rank=rank, world_size=world)
    dev = torch.device("cuda", local_rank)

    A = (d["A_rp"], d["A_ci"], d["A_v"]); G = (d["G_rp"], d["G_ci"], d["G_v"])
    pl = plan(d["part"], world)
    t0 = time.time()
    loc = shard(pl, rank, A, G, d["b"])
    shard_s = time.time() - t0

    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid = torch.tensor(list(capi.nccl_unique_id()), dtype=torch.uint8, device=dev)
    dist.broadcast(uid, 0)
    uid_bytes = bytes(uid.cpu().tolist())

    def make_solver(uid_b):
        """Handle + NCCL communicator (a one-off per process, reported separately), matrices not yet set."""
        s = capi.Solver(local_rank)
        t0 = time.time()
        s.dist_init(world, rank, uid_b, loc["n_sub"], pl.top_depth)
        return s, 1e3 * (time.time() - t0)

    def set_matrices(s):
        s.set_matrix(*loc["A"])
        s.set_factor_blocks(*loc["G"], loc["bounds"], loc["depth"])

    def sync_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    s, comm_init_ms = make_solver(uid_bytes)
    set_matrices(s)
    s.set_rhs(loc["b"])
    for w in range(args.warmup):
        relres, itr = s.pcg_resident(1e-8, 500)
    dist.barrier(); torch.cuda.synchronize()
    launches0 = s.stats()["kernel_launches"]
    dev_ms, iters_total = 0.0, 0
    sampler = None
    if rank == 0:
        import bench as _bench                      # the clock sampler and the measured peak live with the bench contract
        sampler = _bench.ClockSampler(local_rank)
        sampler.start()
    for k in range(args.steps):
        dist.barrier(); torch.cuda.synchronize()
        relres, itr = s.pcg_resident(1e-8, 500)
        torch.cuda.synchronize()
        dev_ms += sync_max(s.stats()["solve_ms"])
        iters_total += itr
    launches = s.stats()["kernel_launches"] - launches0
    clocks = sampler.stop() if sampler else None
    x_loc = s.solution()
    st = s.stats()
    s.close()

    # end to end: per-rank upload of the local matrices + analysis + solve + download, max over ranks
    # (a fresh communicator per solver is not needed: the id is reused only once more)
    e2e_ms = None
    h2d = d2h = 0
    try:
        uid2 = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid2 = torch.tensor(list(capi.nccl_unique_id()), dtype=torch.uint8, device=dev)
        dist.broadcast(uid2, 0)
        uid_bytes = bytes(uid2.cpu().tolist())
        s2, comm_init_ms2 = make_solver(uid_bytes)     # communicator set-up: outside the per-solve region
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.time()
        set_matrices(s2)
        x2, relres2, itr2 = s2.pcg(loc["b"], 1e-8, 500)
        torch.cuda.synchronize()
        e2e_ms = sync_max(1e3 * (time.time() - t0))
        st2 = s2.stats()
        h2d, d2h = st2["h2d_bytes"] + 8 * x2.shape[0], st2["d2h_bytes"]
        s2.close()
    except Exception as e:  # pragma: no cover
        if rank == 0:
            print(f"[bench] e2e leg failed: {e}", flush=True)

    # gather the solution on rank 0 and check the true residual of the assembled vector there
    pieces = [None] * world
    dist.all_gather_object(pieces, (x_loc, loc["index"], loc["n_sub"]))
    if rank == 0:
        import scipy.sparse as sp
        N = int(d["A_rp"].shape[0] - 1)
        x = assemble(pl, pieces)
        Ag = sp.csr_matrix((d["A_v"], d["A_ci"].astype(np.int64), d["A_rp"].astype(np.int64)), shape=(N, N))
        true_rel = float(np.linalg.norm(Ag @ x - d["b"]) / np.linalg.norm(d["b"]))
        value = B_iter * iters_total / (dev_ms * 1e-3) / 1e9
        line = dict(metric="pcg_gbps_per_iter", value=value, unit="GB/s", n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=dev_ms / args.steps, higher_is_better=True, scaling="strong",
                    vs_baseline=None, dtype="f64", data="synthetic", config=config,
                    iterations=iters_total // args.steps, relres=relres, assembled_true_relres=true_rel,
                    ms_per_iter=dev_ms / max(iters_total, 1),
                    sharding=dict(ranks=world, top_depth=pl.top_depth, top_rows=int(loc["index"].shape[0] - loc["n_sub"]),
                                  subtree_rows=[hi - lo for lo, hi in pl.sub_range], shard_host_s=shard_s,
                                  collectives_per_iteration="2 vector all-reduces (top-separator rows) + 3 scalar all-reduces, NCCL"),
                    e2e=dict(value=(B_iter * iters_total / args.steps / (e2e_ms * 1e-3) / 1e9) if e2e_ms else None,
                             unit="GB/s", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h), ms_per_step=e2e_ms,
                             note="per rank: upload of the local matrices, analysis, solve, download; max over ranks; the NCCL "
                                  "communicator of the handle is created before the timed region (a one-off per process)",
                             nccl_comm_init_ms=comm_init_ms),
                    roofline=_roofline(value, world, B_iter, dev_ms / max(iters_total, 1)), clocks=clocks,
                    gpu_launches=int(launches), rank0_setup=dict(upload_ms=st["upload_ms"], analysis_ms=st["analysis_ms"]))
        print(json.dumps(line), flush=True)
    dist.barrier()
    return 0
