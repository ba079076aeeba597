"""torchrun target: sharded solve on N GPUs (C ABI + NCCL) checked against the oracle on rank 0."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
from rchol_b200 import capi, multigpu

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world)
dev = torch.device("cuda", local)

def relerr(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))

def problem(kind):
    if isinstance(kind, str):
        d = np.load(os.path.join(ROOT, "tests", "golden", kind))
        return (d["A_rowPtr"], d["A_colIdx"], d["A_val"]), (d["G_rowPtr"], d["G_colIdx"], d["G_val"]), d["b"], d["part"]
    from conftest import make_problem
    n, T = kind
    A, b, G, part, f = make_problem("lap3d", n, T)
    return A, G, b, part

cases = ["lap3d_12_t4.npz", "lap3d_10_t8_tol6.npz", (40, 16)] if world <= 4 else ["lap3d_10_t8_tol6.npz", (40, 16)]
ok = True
for case in cases:
    A, G, b, part = problem(case)
    if (len(part) // 2) < world:
        continue
    pl = multigpu.plan(part, world)
    loc = multigpu.shard(pl, rank, A, G, b)
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid = torch.tensor(list(capi.nccl_unique_id()), dtype=torch.uint8, device=dev)
    dist.broadcast(uid, 0)
    s = capi.Solver(local)
    s.dist_init(world, rank, bytes(uid.cpu().tolist()), loc["n_sub"], pl.top_depth)
    s.set_matrix(*loc["A"])
    s.set_factor_blocks(*loc["G"], loc["bounds"], loc["depth"])
    q = s.spmv(loc["b"])
    y = s.trsv(capi.TRSV_FORWARD, loc["b"])
    z = s.precond(loc["b"])
    x, relres, itr = s.pcg(loc["b"], 1e-8, 500)
    s.close()
    got = [None] * world
    dist.all_gather_object(got, dict(q=q, y=y, z=z, x=x, index=loc["index"], n_sub=loc["n_sub"], itr=itr, relres=relres))
    if rank == 0:
        from oracle import oracle
        asm = lambda k: multigpu.assemble(pl, [(g[k], g["index"], g["n_sub"]) for g in got])
        o = oracle.pcg(A, b, 1e-8, 500, G)
        e = dict(spmv=relerr(asm("q"), oracle.spmv(*A, b)), fwd=relerr(asm("y"), oracle.trsv_forward(*G, b)),
                 precond=relerr(asm("z"), oracle.precond(*G, b)), x=relerr(asm("x"), o["x"]))
        its = {g["itr"] for g in got}
        good = e["spmv"] < 1e-14 and e["fwd"] <= 1e-12 and e["precond"] <= 1e-12 and len(its) == 1 and abs(itr - o["itr"]) <= 1 and relres <= 2e-8
        if itr == o["itr"]:
            good = good and e["x"] < 1e-9
        print("case", case, "world", world, "errors", e, "iterations", its, "oracle", o["itr"], "relres", relres, "OK" if good else "FAIL", flush=True)
        ok = ok and good
    dist.barrier()
flag = torch.tensor([1 if ok else 0], device=dev)
dist.broadcast(flag, 0)
if rank == 0 and ok:
    print("DIST_CHECK_OK", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
