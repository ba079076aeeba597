#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json on B200 hardware.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload: 3D 7-point Laplacian 256^3, reference `rchol(A, G, P, threads=T)` factorization with T = 4096 nested-dissection
leaves (computed by the UNMODIFIED reference code on the host with a fixed seed -- an input, not timed), PCG to a relative
residual of 1e-8.  This is BASELINE.json configs[2] ("2^k ND partition") at the largest grid whose reference factorization
fits the GPU box: 512^3 needs ~280 GB of host memory (35 GB measured at 256^3, x8) and the box has 206 GB (DESIGN.md
"Workload"); T is the best of the measured leaves sweep (profiles/r02_leaves_sweep_256.jsonl).  BASELINE.json configs[1]
(the same grid with the reference example's 8-way partition) is measured beside it in the same line (`configs1`), and
configs[3] (the 2-D anisotropic random-weight SDDM, at 4096^2) in a child process with a time limit (`configs3`).
A "step" is one full PCG solve.  The metric is the algorithmic memory traffic of the PCG iterations
divided by the time they take: bytes per iteration (SURVEY.md 8d / BASELINE.md section 3)
    B_iter = 12 nnz(A) + 4 (N+1) + 2 [12 nnz(G) + 4 (N+1)] + 136 N
times the iterations done, over the device time of the solve -- "GB/s per iteration vs the HBM roofline".
`value` is measured with A, G and b resident in HBM (CUDA events around the solve); `e2e` is the same metric through the
drop-in entry point with HOST buffers (upload of A, G, b, device set-up/analysis, solve and download of x inside the timed
region).  The reference arm (--impl reference) times the reference's own CPU implementation of the same path on the host
cores on a bounded sample (a fixed number of iterations per step).
"""
from __future__ import annotations

import os
import sys

if "--impl" in sys.argv and "reference" in sys.argv:
    # The reference arm runs the reference's CPU code on ALL host cores.  torchrun exports OMP_NUM_THREADS=1 to every
    # rank; the OpenMP / MKL runtimes read the environment when they load, so it is overridden here, before numpy,
    # torch or the oracle libraries are imported.
    _n = str(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    for _k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[_k] = _n

import argparse
import json
import subprocess
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from rchol_b200 import problems, producer  # noqa: E402

TOL = 1e-8
MAXIT = 500
SEED = 20240


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------------------
# problem construction (inputs of the hot path; cached on local disk so that the two arms share the factor)
# ------------------------------------------------------------------------------------------------------------
def build_problem(n: int, threads: int, kind: str = "lap3d"):
    """A, G, part, b, P of the workload (`kind`: "lap3d" = 7-point Laplacian on n^3, "aniso2d" = the anisotropic
    random-weight 5-point SDDM on n^2).  Generated once per box (reference factorization on the host, fixed seed) and
    cached on local disk; under torchrun only local rank 0 generates, the other ranks wait for the cache and map it."""
    cache_root = os.environ.get("RCHOL_B200_CACHE", "/tmp/rchol_b200_cache")
    tag = os.path.join(cache_root, f"{kind}_{n}_T{threads}_s{SEED}")
    names = ["A_rp", "A_ci", "A_v", "G_rp", "G_ci", "G_v", "part", "b", "P"]
    ready = f"{tag}.ready"
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    multi = int(os.environ.get("WORLD_SIZE", 1)) > 1

    def load():
        t0 = time.time()
        d = {k: np.load(f"{tag}.{k}.npy", mmap_mode="r" if multi else None) for k in names}
        log(f"[bench] loaded cached problem {tag} in {time.time() - t0:.1f}s")
        return d, dict(cached=True, tag=tag)

    if os.path.exists(ready) and all(os.path.exists(f"{tag}.{k}.npy") for k in names):
        return load()
    if multi and local_rank != 0:
        t0 = time.time()
        while not os.path.exists(ready):
            if time.time() - t0 > 3600:
                raise RuntimeError("timed out waiting for local rank 0 to generate the problem")
            time.sleep(1.0)
        return load()
    t0 = time.time()
    A = problems.laplace_3d(n) if kind == "lap3d" else problems.aniso_2d(n)
    t1 = time.time()
    f = producer.factor(*A, threads=threads, seed=SEED)
    t2 = time.time()
    b = problems.random_rhs(f.N)
    if threads > 0:
        Ap = producer.ref_reorder(*A, f.P)
        bp = problems.reorder_vector(b, f.P)
    else:
        Ap, bp = A, b
    t3 = time.time()
    d = dict(A_rp=Ap[0], A_ci=Ap[1], A_v=Ap[2], G_rp=f.rowPtr, G_ci=f.colIdx, G_v=f.val, part=f.part, b=bp, P=f.P)
    log(f"[bench] generated {kind} {n}^{3 if kind == 'lap3d' else 2}: gen {t1 - t0:.1f}s, reference factorization (T={threads}) {t2 - t1:.1f}s, "
        f"reorder {t3 - t2:.1f}s, nnzG={f.nnz}")
    try:
        os.makedirs(cache_root, exist_ok=True)
        for k in names:                      # write-then-rename: a reader never sees a partial file
            tmp = f"{tag}.{k}.tmp.npy"
            np.save(tmp, d[k])
            os.replace(tmp, f"{tag}.{k}.npy")
        with open(ready, "w") as fh:
            fh.write("ok\n")
    except OSError as e:  # cache is an optimisation only (single process); the other ranks of a torchrun need it
        log(f"[bench] cache not written: {e}")
        if multi:
            raise
    return d, dict(cached=False, gen_s=t1 - t0, factor_s=t2 - t1, reorder_s=t3 - t2, tag=tag)


# ------------------------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md): sampled DURING the timed region
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.device)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()          # the exact PID we started
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1])); smax.append(float(p[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(smax) if smax else None,
                    reasons=sorted(reasons), samples=len(sm))


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the path, bounded sample
# ------------------------------------------------------------------------------------------------------------
_reference_child_failed = False


def _reference_sample_in_child(d, iters: int):
    """One timed call of the UNMODIFIED reference pcg in a forked child.  Why a child: the reference leaks its set-up
    copies by design (pcg.cpp:31-54, the `delete[]` is commented out: 16 B per entry of A and G, ~10 GB per call at
    256^3), so repeated calls in one process run the host out of memory.  The child shares the problem copy-on-write,
    warms MKL's thread pool on a tiny problem, runs `iters` iterations and reports the seconds spent inside
    pcg::iteration (marks in oracle/mkl_adapter.cpp: end of the last create_csr -> first destroy), i.e. without the
    set-up copies that a real solve spreads over ~37 iterations."""
    rfd, wfd = os.pipe()
    pid = os.fork()
    if pid == 0:                                             # child: nothing but the reference call
        status = 1
        try:
            os.close(rfd)
            from oracle import oracle
            w = problems.laplace_3d(6)
            oracle.reference_pcg(w, np.ones(w[0].shape[0] - 1), 1e-30, 2, w)          # warm-up (threads, MKL init)
            A = (d["A_rp"], d["A_ci"], d["A_v"]); G = (d["G_rp"], d["G_ci"], d["G_v"])
            t0 = time.time()
            r = oracle.reference_pcg(A, d["b"], 1e-30, iters, G)                       # tol unreachable: `iters` iterations
            import torch
            out = dict(iteration_s=r["iteration_s"], call_s=time.time() - t0, itr=int(r["itr"]),
                       omp_threads=int(oracle.num_threads()), mkl_threads=int(torch.get_num_threads()))
            os.write(wfd, json.dumps(out).encode())
            status = 0
        except BaseException as e:  # pragma: no cover
            try:
                os.write(wfd, json.dumps(dict(error=repr(e)[:300])).encode())
            except OSError:
                pass
        finally:
            os._exit(status)
    os.close(wfd)
    buf = b""
    while True:
        chunk = os.read(rfd, 65536)
        if not chunk:
            break
        buf += chunk
    os.close(rfd)
    _, st = os.waitpid(pid, 0)
    if st != 0 or not buf:
        raise RuntimeError(f"reference child ended with status {st}: {buf[-300:]!r}")
    out = json.loads(buf.decode())
    if "error" in out or out["iteration_s"] <= 0:
        raise RuntimeError(f"reference child: {out}")
    return out


def cpu_sample(d, iters: int, prefer_reference: bool):
    """Runs `iters` PCG iterations on the host cores; returns (seconds, iterations, kind, cores, detail)."""
    from oracle import oracle
    A = (d["A_rp"], d["A_ci"], d["A_v"]); G = (d["G_rp"], d["G_ci"], d["G_v"])
    global _reference_child_failed
    if (prefer_reference and not _reference_child_failed and oracle.have_reference_pcg()
            and int(d["G_rp"][-1]) < 2 ** 31 - 1):
        try:
            o = _reference_sample_in_child(d, iters)
            return o["iteration_s"], o["itr"], "reference", min(o["omp_threads"], o["mkl_threads"]), dict(
                what="unmodified reference pcg.cpp + oneMKL SpMV/SpTRSV (libtorch_cpu), OpenMP CBLAS-1 stand-ins; seconds "
                     "inside pcg::iteration (its six zero-filled work vectors and the final true-residual SpMV included, "
                     "its set-up copies not), one forked child per step (the reference leaks its set-up copies)",
                whole_call_s=o["call_s"], omp_threads=o["omp_threads"], mkl_threads=o["mkl_threads"],
                env_omp_num_threads=os.environ.get("OMP_NUM_THREADS"))
        except Exception as e:  # pragma: no cover
            # sticky: the port below starts an OpenMP pool in THIS process, after which forking again is not safe
            _reference_child_failed = True
            log(f"[bench] reference pcg unavailable ({e}); using the oracle port")
    t0 = time.time()
    o = oracle.pcg(A, d["b"], 1e-30, iters, G)
    dt = time.time() - t0
    return dt, o["itr"], "port", oracle.num_threads(), {k: float(v) for k, v in o["timings"].items()}


def run_reference_arm(args, d, B_iter, rank, world):
    if rank != 0:
        return
    iters = args.sample_iters
    times = []
    its_total = 0
    kind = cores = detail = None
    for step in range(args.warmup + args.steps):
        dt, it, kind, cores, detail = cpu_sample(d, iters, prefer_reference=True)
        log(f"[bench/reference] step {step}: {it} iterations in {dt:.2f}s ({kind})")
        if step >= args.warmup:
            times.append(dt)
            its_total += it
    total = sum(times)
    value = B_iter * its_total / total / 1e9
    N = d["A_rp"].shape[0] - 1
    line = dict(impl="reference", metric="pcg_gbps_per_iter", value=value, unit="GB/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * total / len(times), higher_is_better=True,
                scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                config=workload_config(args, d, N),
                cpu_baseline=dict(value=value, unit="GB/s", cores=cores, kind=kind,
                                  sample=f"{iters} PCG iterations per step on the full {args.n}^3 problem; seconds inside the "
                                         f"reference's pcg::iteration (its set-up copies excluded)"),
                e2e=dict(value=value, unit="GB/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0, detail=detail)
    print(json.dumps(line), flush=True)


def workload_config(args, d, N):
    return dict(workload=f"lap3d_{args.n}^3_rchol_T{args.threads}_pcg_tol1e-8", n=args.n, N=int(N),
                nnzA=int(d["A_rp"][-1]), nnzG=int(d["G_rp"][-1]), nd_leaves=args.threads, tol=TOL, maxit=MAXIT,
                factor_seed=SEED, l2_policy="inputs_exceed_l2 (A+G streamed per iteration >> 126 MB)",
                baseline_config="configs[2] (2^k ND partition) at 256^3: the reference factorization of 512^3 needs ~280 GB of host "
                                "memory, the GPU box has 206 GB; k = 12 from the measured leaves sweep; configs[1] (T=8) in `configs1`")


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def _relerr(a, b):
    nb = float(np.linalg.norm(b))
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0))


def measure_resident(d, threads, steps, warmup, peak, levels=True):
    """A, G, b resident in HBM: `warmup` untimed and `steps` timed PCG solves (CUDA events inside the library, on its own
    stream), then the per-launch split of the triangular solves.  Returns (result dict, solver) -- the solver is still
    open so that the caller can add parity checks on the same device copy."""
    from rchol_b200 import capi
    A = (d["A_rp"], d["A_ci"], d["A_v"]); G = (d["G_rp"], d["G_ci"], d["G_v"])
    part = d["part"] if threads > 0 else None
    N = A[0].shape[0] - 1
    nnzA, nnzG = int(A[0][-1]), int(G[0][-1])
    B_iter = problems.algorithmic_bytes_per_iteration(N, nnzA, nnzG)
    s = capi.Solver(0)
    t0 = time.time()
    s.set_matrix(*A)
    s.set_factor(*G, part)
    s.set_rhs(d["b"])
    setup_wall = time.time() - t0
    st0 = s.stats()
    log(f"[bench] T={threads}: upload {st0['upload_ms']:.0f} ms, analysis {st0['analysis_ms']:.0f} ms")
    for w in range(warmup):
        relres, itr = s.pcg_resident(TOL, MAXIT)
        log(f"[bench] T={threads} warm-up {w}: {itr} iterations, relres {relres:.3e}, {s.stats()['solve_ms']:.1f} ms")
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = s.stats()["kernel_launches"]
    dev_ms, iters_total = 0.0, 0
    wall0 = time.time()
    for k in range(steps):
        relres, itr = s.pcg_resident(TOL, MAXIT)
        dev_ms += s.stats()["solve_ms"]
        iters_total += itr
    wall = time.time() - wall0
    launches = s.stats()["kernel_launches"] - launches0
    clocks = sampler.stop()
    value = B_iter * iters_total / (dev_ms * 1e-3) / 1e9
    out = dict(B_iter=B_iter, N=N, nnzA=nnzA, nnzG=nnzG, value=value, dev_ms=dev_ms, iters_total=iters_total, relres=relres,
               wall=wall, launches=int(launches), clocks=clocks, setup=dict(upload_ms=st0["upload_ms"], analysis_ms=st0["analysis_ms"],
                                                                           wall_s=setup_wall))
    if levels:
        # The triangular solves run as ONE launch (warp-per-block levels: two) per tree level and direction: the dominant
        # kernels.  Every level is timed with CUDA events on the library's stream (rcg_time_group).  Algorithmic bytes of a
        # level: 12 B per factor entry of its rows, 4 B row pointer, right-hand side in and solution out (16 B).
        prof = s.profile_iteration(3)
        per_level, tot_ms, tot_bytes, nl, dom = {}, 0.0, 0, 0, None
        for direction, dname in ((capi.TRSV_FORWARD, "fwd"), (capi.TRSV_BACKWARD, "bwd")):
            for gi, g in enumerate(s.groups(direction)):
                ms = s.time_group(direction, gi, 0, 3)
                lbytes = 12 * (g["loc_nnz"] + g["ext_nnz"]) + 20 * g["rows"]
                tot_ms += ms; tot_bytes += lbytes; nl += 1
                per_level[f"{dname}:level{gi}:{g['blocks']}blocks"] = dict(
                    rows=g["rows"], max_rows=g["max_rows"], entries=g["loc_nnz"] + g["ext_nnz"], ms=ms, algorithmic_bytes=lbytes,
                    gbs=lbytes / ms / 1e6 if ms else 0.0)
                if dom is None or ms > dom[1]:
                    dom = (f"{dname} level {gi}: {g['blocks']} blocks, {g['rows']} rows", ms, lbytes)
        spmv_ms = prof["spmv_ms"]
        ach = tot_bytes / tot_ms / 1e6
        out["roofline"] = dict(
            bound="hbm", kernel="triangular-solve level launches (k_wb_pre + k_wb_solve / k_dp_pre + k_dp_solve / k_bc_solve): all launches of one forward + one backward solve",
            achieved=ach, peak=peak, unit="GB/s", frac=ach / peak, launches_per_solve_pair=nl, bytes_per_launch=tot_bytes / nl,
            ms_per_launch=tot_ms / nl, trsv_kernel_ms_per_iteration=tot_ms,
            largest_launch=dict(kernel=dom[0], ms=dom[1], gbs=dom[2] / dom[1] / 1e6),
            iteration=dict(bytes=B_iter, ms=dev_ms / max(iters_total, 1), gbs=value, frac=value / peak, trsv_ms=prof["trsv_ms"],
                           spmv_ms=spmv_ms, blas1_ms=prof["blas1_ms"],
                           spmv_gbs=(12 * nnzA + 4 * N + 24 * N) / spmv_ms / 1e6 if spmv_ms else 0.0),
            tree_levels=per_level)
    return out, s


def parity_at_bench_config(s, d, tag):
    """Parity AT THE BENCHMARKED SIZE, outside the timed region: both triangular solves and their composition against the
    oracle's sequential substitution (oracle/pcg_oracle.c restating pcg.cpp:141-159) on the bench right-hand side, and
    the iteration count of one converged CPU solve of the restated reference loop (cached beside the problem)."""
    from oracle import oracle
    from rchol_b200 import capi
    A = (d["A_rp"], d["A_ci"], d["A_v"]); G = (d["G_rp"], d["G_ci"], d["G_v"])
    b = np.asarray(d["b"])
    t0 = time.time()
    yo = oracle.trsv_forward(*G, b)
    zo = oracle.trsv_backward(*G, yo)
    out = dict(fwd=_relerr(s.trsv(capi.TRSV_FORWARD, b), yo), bwd=_relerr(s.trsv(capi.TRSV_BACKWARD, yo), zo),
               precond=_relerr(s.precond(b), zo), tolerance=1e-12, oracle_trsv_s=time.time() - t0)
    cache = f"{tag}.parity.json" if tag else None
    ref = None
    if cache and os.path.exists(cache):
        try:
            ref = json.load(open(cache))
        except Exception:
            ref = None
    if ref is None:
        t0 = time.time()
        o = oracle.pcg(A, b, TOL, MAXIT, G)
        ref = dict(itr_ref=int(o["itr"]), relres_ref=float(o["relres"]), cpu_solve_s=time.time() - t0, threads=oracle.num_threads())
        ref["x_norm"] = float(np.linalg.norm(o["x"]))
        np.save(f"{tag}.xref.npy", o["x"]) if tag else None
        if cache:
            try:
                json.dump(ref, open(cache, "w"))
            except OSError:
                pass
    out.update(ref)
    x = s.solution()
    try:
        out["x_relerr_vs_cpu"] = _relerr(x, np.load(f"{tag}.xref.npy"))
    except Exception:
        out["x_relerr_vs_cpu"] = None
    return out


# ------------------------------------------------------------------------------------------------------------
# secondary workloads measured beside the headline, each in its OWN process (a failure or time-out of a leg never
# costs the headline line): BASELINE.json configs[3] -- 2-D anisotropic random-weight SDDM (pure-chain separators)
# ------------------------------------------------------------------------------------------------------------
def run_leg(args):
    """Child process of the N = 1 arm: one secondary workload measured like the headline -- A, G, b resident, `steps`
    timed PCG solves after `warmup`, the per-level split of the triangular solves, parity against the oracle at this size
    (both triangular solves; iteration count of one converged CPU solve) -- printed as one JSON object."""
    d, info = build_problem(args.n, args.threads, args.leg)
    peak, _ = measured_peak_gbs()
    m, s = measure_resident(d, args.threads, args.steps, args.warmup, peak)
    parity = None
    if not args.no_parity:
        try:
            parity = parity_at_bench_config(s, d, info.get("tag"))
            parity["itr"] = m["iters_total"] // args.steps
        except Exception as e:  # pragma: no cover
            parity = dict(error=repr(e)[:300])
    s.close()
    r = m["roofline"]
    dims = 3 if args.leg == "lap3d" else 2
    out = dict(workload=f"{args.leg}_{args.n}^{dims}_rchol_T{args.threads}_pcg_tol1e-8", N=int(m["N"]), nnzA=m["nnzA"], nnzG=m["nnzG"],
               nd_leaves=args.threads, iterations=m["iters_total"] // args.steps, relres=m["relres"],
               ms_per_iter=m["dev_ms"] / max(m["iters_total"], 1), ms_per_solve=m["dev_ms"] / args.steps, value=m["value"], unit="GB/s",
               frac_of_peak=m["value"] / peak, bytes_per_iteration=m["B_iter"],
               trsv=dict(frac_of_peak=r["frac"], gbs=r["achieved"], ms_per_iteration=r["trsv_kernel_ms_per_iteration"],
                         launches_per_solve_pair=r["launches_per_solve_pair"], largest_launch=r["largest_launch"]),
               spmv_ms=r["iteration"]["spmv_ms"], spmv_gbs=r["iteration"]["spmv_gbs"], blas1_ms=r["iteration"]["blas1_ms"],
               tree_levels={k: dict(rows=v["rows"], max_rows=v["max_rows"], ms=v["ms"], gbs=v["gbs"]) for k, v in r["tree_levels"].items()},
               parity=parity, setup=m["setup"], gpu_launches=m["launches"], clocks=m["clocks"],
               factor_s=info.get("factor_s"), steps=args.steps, warmup=args.warmup)
    if parity and parity.get("cpu_solve_s") and parity.get("itr_ref"):
        out["cpu_port_ms_per_iter"] = 1e3 * parity["cpu_solve_s"] / parity["itr_ref"]
        out["cpu_port_threads"] = parity.get("threads")
    print(json.dumps(out), flush=True)


def leg_in_subprocess(kind: str, n: int, threads: int, steps: int, warmup: int, timeout_s: float):
    """Runs `bench.py --leg ...` as a child with a time limit; returns its JSON object, or {"error": ...}."""
    cmd = [sys.executable, os.path.abspath(__file__), "--leg", kind, "--n", str(n), "--threads", str(threads),
           "--steps", str(steps), "--warmup", str(warmup)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE")}
    t0 = time.time()
    try:
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=None, text=True, timeout=timeout_s, env=env, cwd=ROOT)
    except subprocess.TimeoutExpired:
        return dict(error=f"leg exceeded its {timeout_s:.0f} s limit (reference factorization on the host + solves)")
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    if p.returncode != 0 or not lines:
        return dict(error=f"leg ended with status {p.returncode}", stdout_tail=p.stdout[-300:])
    out = json.loads(lines[-1])
    out["leg_wall_s"] = time.time() - t0
    return out


def run_ours_single(args, d, B_iter, gen_info=None):
    gen_info = gen_info or {}
    from rchol_b200 import capi
    A = (d["A_rp"], d["A_ci"], d["A_v"]); G = (d["G_rp"], d["G_ci"], d["G_v"])
    part = d["part"] if args.threads > 0 else None
    N = A[0].shape[0] - 1
    peak, peak_src = measured_peak_gbs()

    m, s = measure_resident(d, args.threads, args.steps, args.warmup, peak)
    value, dev_ms, iters_total, relres = m["value"], m["dev_ms"], m["iters_total"], m["relres"]
    x_dev = s.solution()
    roofline = m["roofline"]
    roofline["peak_source"] = peak_src
    # dram__bytes_read+write of the level launches from an `ncu --set full` capture AT THE BENCH CONFIGURATION
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_trsv_traffic.json")))
        if tr.get("workload") == workload_config(args, d, N)["workload"]:
            roofline["traffic"] = tr["dram_bytes_per_solve_pair"] / roofline["launches_per_solve_pair"]
            roofline["traffic_over_algorithmic"] = tr["dram_bytes_per_solve_pair"] / (roofline["bytes_per_launch"] * roofline["launches_per_solve_pair"])
            roofline["traffic_source"] = tr.get("source")
        else:
            roofline["traffic"] = None
    except Exception:
        roofline["traffic"] = None

    parity = None
    if not args.no_parity:
        try:
            parity = parity_at_bench_config(s, d, gen_info.get("tag"))
            parity["itr"] = iters_total // args.steps
            log(f"[bench] parity at the bench configuration: {parity}")
        except Exception as e:  # pragma: no cover
            parity = dict(error=repr(e)[:300])
    # ---- a new matrix with the same sparsity pattern on the open handle (the reference's reuse flow) --
    refresh = None
    try:
        t0 = time.time(); s.set_matrix(*A); full_ms = 1e3 * (time.time() - t0)
        t0 = time.time(); s.update_matrix_values(A[2]); vals_ms = 1e3 * (time.time() - t0)
        relres_r, itr_r = s.pcg_resident(TOL, MAXIT)
        refresh = dict(set_matrix_ms=full_ms, update_matrix_values_ms=vals_ms, bytes_full=int(12 * m["nnzA"] + 8 * (N + 1)),
                       bytes_values_only=int(8 * m["nnzA"]), solve_after_refresh_identical=bool(np.array_equal(s.solution(), x_dev)),
                       set_factor_first_ms=m["setup"]["upload_ms"] + m["setup"]["analysis_ms"],
                       note="set_factor_first_ms: first set_matrix + set_factor of the process (cold allocator); the repeated "
                            "set-up is e2e.parts (a factorization samples a new sparsity pattern, so G has no value-only refresh)")
    except Exception as e:  # pragma: no cover - informational leg
        refresh = dict(error=repr(e)[:300])
    s.close()

    # ---- end to end through the drop-in entry point, host buffers -------------------------------------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e_ms, e2e_iters, h2d, d2h = 0.0, 0, 0, 0
    setup_ms = None
    for k in range(e2e_steps):
        t0 = time.time()
        x, relres_e, itr_e, st = capi.pcg(A, d["b"], TOL, MAXIT, G, part)
        e2e_ms += 1e3 * (time.time() - t0)
        e2e_iters += itr_e
        h2d, d2h = st["h2d_bytes"] + 8 * N, st["d2h_bytes"]
        setup_ms = dict(upload_ms=st["upload_ms"], analysis_ms=st["analysis_ms"], solve_ms=st["solve_ms"])
    e2e_value = B_iter * e2e_iters / (e2e_ms * 1e-3) / 1e9
    assert np.array_equal(x, x_dev), "one-shot and resident solves differ"

    # ---- reorder(A, P) on the device (upload of the ORIGINAL A + permutation + per-row sort) ------------
    reorder_info = None
    if args.threads > 0:
        try:
            A0 = problems.laplace_3d(args.n)
            with capi.Solver(0) as s3:
                t0 = time.time()
                s3.set_matrix_permuted(*A0, d["P"])
                t_dev = time.time() - t0
                same = all(np.array_equal(u, v) for u, v in zip(s3.get_matrix(), A))
            reorder_info = dict(upload_plus_device_reorder_ms=1e3 * t_dev, bit_identical_to_reference_reorder=bool(same),
                                reference_host_reorder_s=gen_info.get("reorder_s"))
            del A0
        except Exception as e:  # pragma: no cover - informational leg
            reorder_info = dict(error=str(e))

    # ---- CPU baseline beside it (bounded sample, rank 0) ------------------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        dt, it, kind, cores, detail = cpu_sample(d, args.sample_iters, prefer_reference=False)
        cpu = dict(value=B_iter * it / dt / 1e9, unit="GB/s", cores=cores, kind=kind,
                   sample=f"{it} PCG iterations (of {iters_total // args.steps} to convergence) on the full {args.n}^3 problem, "
                          f"{dt:.1f} s", ms_per_iter=1e3 * dt / it, detail=detail)

    # ---- BASELINE.json configs[1] beside the headline: the same grid with the reference's 8-way partition --------------
    configs1 = None
    if args.configs1 and args.threads != 8:
        try:
            d8, info8 = build_problem(args.n, 8)
            m8, s8 = measure_resident(d8, 8, max(1, min(args.steps, 3)), 2, peak)
            s8.close()
            r8 = m8["roofline"]
            configs1 = dict(workload=f"lap3d_{args.n}^3_rchol_T8_pcg_tol1e-8", nnzG=m8["nnzG"], iterations=m8["iters_total"] // max(1, min(args.steps, 3)),
                            relres=m8["relres"], ms_per_iter=m8["dev_ms"] / max(m8["iters_total"], 1), value=m8["value"], unit="GB/s",
                            frac_of_peak=m8["value"] / peak, trsv_frac_of_peak=r8["frac"], trsv_ms_per_iteration=r8["trsv_kernel_ms_per_iteration"],
                            largest_launch=r8["largest_launch"], factor_s=info8.get("factor_s"))
            del d8
        except Exception as e:  # pragma: no cover - informational leg
            configs1 = dict(error=repr(e)[:300])

    # ---- BASELINE.json configs[3] beside the headline: 2-D anisotropic random-weight SDDM, in its own process ----------
    configs3 = None
    if args.configs3:
        try:
            configs3 = leg_in_subprocess("aniso2d", args.aniso_n, args.aniso_threads, 2, 2, args.leg_timeout)
        except Exception as e:  # pragma: no cover - informational leg
            configs3 = dict(error=repr(e)[:300])
        configs3["baseline_config"] = (f"configs[3] (2-D anisotropic random-weight SDDM) at {args.aniso_n}^2 instead of 8192^2: the reference "
                                       "factorization of 8192^2 takes ~15 min of host time per run (measured: 57 s at 2048^2 on 8 cores, linear in N)")

    line = dict(metric="pcg_gbps_per_iter", value=value, unit="GB/s", n_gpus=1, steps=args.steps, warmup=args.warmup,
                ms_per_step=dev_ms / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None,
                dtype="f64", data="synthetic", config=workload_config(args, d, N),
                iterations=iters_total // args.steps, relres=relres, ms_per_iter=dev_ms / max(iters_total, 1),
                time_to_solution_ms=dict(resident=dev_ms / args.steps, end_to_end=e2e_ms / e2e_steps, parts=setup_ms),
                wall_ms_per_step=1e3 * m["wall"] / args.steps,
                roofline=roofline, cpu_baseline=cpu, parity=parity,
                e2e=dict(value=e2e_value, unit="GB/s", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                         steps=e2e_steps, ms_per_step=e2e_ms / e2e_steps, host_memory="pageable (caller's SparseCSR arrays)"),
                gpu_launches=m["launches"], clocks=m["clocks"], device_reorder=reorder_info, configs1=configs1, configs3=configs3,
                setup=m["setup"], refresh=refresh)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=int(os.environ.get("RCHOL_B200_BENCH_N", 256)))
    ap.add_argument("--threads", type=int, default=int(os.environ.get("RCHOL_B200_BENCH_T", 4096)),
                    help="leaves of the reference's nested-dissection partition (2^k)")
    ap.add_argument("--sample-iters", type=int, default=4, help="PCG iterations per CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison at the bench configuration")
    ap.add_argument("--no-configs1", dest="configs1", action="store_false",
                    help="skip the BASELINE.json configs[1] leg (same grid, the reference's 8-way partition)")
    ap.add_argument("--no-configs3", dest="configs3", action="store_false",
                    help="skip the BASELINE.json configs[3] leg (2-D anisotropic SDDM, run in a child process)")
    ap.add_argument("--aniso-n", type=int, default=int(os.environ.get("RCHOL_B200_BENCH_ANISO_N", 4096)))
    ap.add_argument("--aniso-threads", type=int, default=int(os.environ.get("RCHOL_B200_BENCH_ANISO_T", 4096)))
    ap.add_argument("--leg-timeout", type=float, default=900.0, help="time limit of a secondary-workload child, seconds")
    ap.add_argument("--leg", default=None, choices=["lap3d", "aniso2d"],
                    help="(internal) measure one secondary workload given by --n/--threads and print its JSON object")
    args = ap.parse_args()

    if args.leg:
        run_leg(args)
        return 0
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))

    if args.impl == "reference" and rank != 0:
        return 0
    d, info = build_problem(args.n, args.threads)
    N = d["A_rp"].shape[0] - 1
    B_iter = problems.algorithmic_bytes_per_iteration(N, int(d["A_rp"][-1]), int(d["G_rp"][-1]))
    if args.impl == "reference":
        run_reference_arm(args, d, B_iter, rank, world)
        return 0
    if world > 1 or args.gpus > 1:
        from rchol_b200 import multigpu
        return multigpu.bench_main(args, d, B_iter, rank, world, workload_config(args, d, N))
    run_ours_single(args, d, B_iter, info)
    return 0


if __name__ == "__main__":
    sys.exit(main())
