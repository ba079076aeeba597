"""ctypes binding of the C ABI in ``include/rchol_b200.h`` (``rchol_b200/lib/librchol_b200.so``).

This is the Python face of the drop-in boundary used by tests/ and bench.py; the C++ face is
``rchol_b200/cxx/pcg.hpp`` (same constructor signature as /root/reference/c++/util/pcg.hpp:13-16).
There is no fallback of any kind: if the CUDA library is missing or no B200 is visible, every call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "librchol_b200.so")

_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")

# every symbol include/rchol_b200.h declares (tests check that the library exports all of them)
EXPORTED = [
    "rcg_create", "rcg_create_with_options", "rcg_destroy", "rcg_last_error", "rcg_version",
    "rcg_set_matrix", "rcg_set_factor", "rcg_spmv", "rcg_trsv", "rcg_precond", "rcg_pcg",
    "rcg_set_rhs", "rcg_pcg_resident", "rcg_get_solution", "rcg_get_history", "rcg_pcg_oneshot",
    "rcg_get_stats", "rcg_profile_iteration", "rcg_time_phase", "rcg_debug_trace",
    "rcg_get_group_count", "rcg_get_group_info", "rcg_time_group",
    "rcg_set_factor_blocks", "rcg_nccl_unique_id", "rcg_dist_init", "rcg_dist_finalize",
    "rcg_debug_blocked_info", "rcg_debug_blocked_copy", "rcg_debug_counters", "rcg_debug_dp_trace",
    "rcg_set_matrix_permuted", "rcg_set_permutation", "rcg_permute_vector", "rcg_unpermute_vector",
    "rcg_pcg_original", "rcg_get_matrix", "rcg_detect_blocks", "rcg_update_matrix_values", "rcg_spmv_row_histogram",
]

TRSV_FORWARD, TRSV_BACKWARD = 0, 1
RCG_OK, RCG_ERR_CUDA, RCG_ERR_INVALID, RCG_ERR_STATE, RCG_ERR_STRUCTURE, RCG_ERR_NOMEM = 0, 1, 2, 3, 4, 5   # rchol_b200.h


class Options(C.Structure):
    _fields_ = [("chain_threads", C.c_int), ("chain_window", C.c_int), ("use_graph", C.c_int),
                ("spmv_lanes", C.c_int), ("chain_generic", C.c_int), ("chain_mode", C.c_int), ("reserved", C.c_int * 10)]


class Stats(C.Structure):
    _fields_ = [("N", C.c_uint64), ("nnzA", C.c_uint64), ("nnzG", C.c_uint64), ("n_blocks", C.c_uint64),
                ("tree_levels", C.c_uint64), ("upload_ms", C.c_double), ("analysis_ms", C.c_double),
                ("solve_ms", C.c_double), ("total_ms", C.c_double), ("trsv_ms", C.c_double),
                ("spmv_ms", C.c_double), ("blas1_ms", C.c_double), ("kernel_launches", C.c_uint64),
                ("launches_per_iteration", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("device_bytes", C.c_uint64), ("reserved", C.c_double * 8)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}
        d["chain_sm_mhz"] = self.reserved[0]   # SM clock seen by the last chain kernel (clock64 / globaltimer)
        d["watchdog_row"] = int(self.reserved[1])  # 1 + row whose dependency wait timed out (0 = none)
        d["dag_levels_fwd"] = int(self.reserved[2])  # total DAG levels (sum over blocks), forward / backward solve
        d["dag_levels_bwd"] = int(self.reserved[3])
        d["crit_cycles"] = dict(wait=self.reserved[4], prologue=self.reserved[5], batches=self.reserved[6], n_batch=self.reserved[7])
        return d


class RcgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rchol_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Loads the CUDA library; raises (loudly) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `make` (or __graft_entry__.build()); "
                           "rchol_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    H = C.c_void_p
    L.rcg_create.argtypes = [C.POINTER(H), C.c_int]
    L.rcg_create_with_options.argtypes = [C.POINTER(H), C.c_int, C.POINTER(Options)]
    L.rcg_destroy.argtypes = [H]
    L.rcg_last_error.restype = C.c_char_p
    L.rcg_last_error.argtypes = [H]
    L.rcg_version.restype = C.c_char_p
    L.rcg_set_matrix.argtypes = [H, C.c_uint64, _u64p, _u64p, _f64p]
    L.rcg_set_factor.argtypes = [H, C.c_uint64, _u64p, _u64p, _f64p, C.c_void_p, C.c_uint64]
    L.rcg_spmv.argtypes = [H, _f64p, _f64p]
    L.rcg_trsv.argtypes = [H, C.c_int, _f64p, _f64p]
    L.rcg_precond.argtypes = [H, _f64p, _f64p]
    L.rcg_pcg.argtypes = [H, _f64p, C.c_double, C.c_int, _f64p, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.rcg_set_rhs.argtypes = [H, _f64p]
    L.rcg_pcg_resident.argtypes = [H, C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.rcg_get_solution.argtypes = [H, _f64p]
    L.rcg_get_history.argtypes = [H, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    L.rcg_pcg_oneshot.argtypes = [C.c_int, C.c_uint64, _u64p, _u64p, _f64p, _f64p, C.c_double, C.c_int,
                                  _u64p, _u64p, _f64p, C.c_void_p, C.c_uint64, _f64p,
                                  C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(Stats)]
    L.rcg_get_stats.argtypes = [H, C.POINTER(Stats)]
    L.rcg_profile_iteration.argtypes = [H, C.c_int]
    L.rcg_time_phase.argtypes = [H, C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.rcg_get_group_count.argtypes = [H, C.c_int, C.POINTER(C.c_int)]
    L.rcg_get_group_info.argtypes = [H, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    L.rcg_time_group.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.rcg_set_factor_blocks.argtypes = [H, C.c_uint64, _u64p, _u64p, _f64p, _u64p, np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS"), C.c_uint64]
    L.rcg_nccl_unique_id.argtypes = [C.c_void_p]
    L.rcg_dist_init.argtypes = [H, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_int]
    L.rcg_dist_finalize.argtypes = [H]
    L.rcg_debug_trace.argtypes = [H, C.c_int, _f64p, _f64p, np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")]
    L.rcg_debug_blocked_info.argtypes = [H, C.c_int, C.POINTER(C.c_uint64)]
    L.rcg_debug_blocked_copy.argtypes = [H, C.c_int, C.c_int, C.c_void_p, C.c_uint64]
    L.rcg_debug_counters.argtypes = [H, C.POINTER(C.c_uint64)]
    L.rcg_debug_dp_trace.argtypes = [H, C.c_void_p, C.c_uint64]
    L.rcg_set_matrix_permuted.argtypes = [H, C.c_uint64, _u64p, _u64p, _f64p, _u64p]
    L.rcg_set_permutation.argtypes = [H, C.c_uint64, _u64p]
    L.rcg_permute_vector.argtypes = [H, _f64p, _f64p]
    L.rcg_unpermute_vector.argtypes = [H, _f64p, _f64p]
    L.rcg_pcg_original.argtypes = [H, _f64p, C.c_double, C.c_int, _f64p, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.rcg_get_matrix.argtypes = [H, _u64p, _u64p, _f64p]
    L.rcg_update_matrix_values.argtypes = [H, C.c_uint64, _f64p]
    L.rcg_spmv_row_histogram.argtypes = [C.c_uint64, _u64p, _u64p, C.POINTER(C.c_int)]
    L.rcg_detect_blocks.argtypes = [C.c_uint64, _u64p, _u64p, _u64p, np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS"),
                                    C.c_uint64, C.POINTER(C.c_uint64)]
    for name in EXPORTED:
        fn = getattr(L, name)
        if name not in ("rcg_last_error", "rcg_version"):
            fn.restype = C.c_int
    _lib = L
    return L


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Solver:
    """Handle-based interface: upload A and G once, then call the kernels or the PCG solve."""

    def __init__(self, device: int = 0, chain_threads: int = 0, chain_window: int = 0, use_graph: bool = True,
                 spmv_lanes: int = 0, chain_generic: bool = False, chain_mode: int = 0, backoff_ns: int = 0, dbg: int = 0, producers: int = 0,
                 recent: int = 0, plain_launch: bool = False, sep_window: int = 0, early: int = 0, capb_quarters: int = 0, slots_a: int = 0, far_lanes2: int = 0, sep_tile: int = 0, early_sep: int = 0,
                 wb_min: int = 0, wb_ell: bool = False, dp_min_rows: int = 0, dp_panel: int = 0, dp_leaf: bool = False, dp_max_blocks: int = 0):
        self._L = load()
        self._h = C.c_void_p()
        opt = Options()
        opt.chain_threads, opt.chain_window = int(chain_threads), int(chain_window)
        opt.use_graph, opt.spmv_lanes = int(bool(use_graph)), int(spmv_lanes)
        opt.chain_generic = int(bool(chain_generic))
        opt.chain_mode = int(chain_mode)
        opt.reserved[0] = int(backoff_ns)
        opt.reserved[1] = int(dbg)
        opt.reserved[2] = 1 if wb_ell else int(producers)
        opt.reserved[3] = int(recent)
        opt.reserved[4] = int(sep_window)
        opt.reserved[5] = int(early) | (int(early_sep) << 8)
        opt.reserved[7] = int(capb_quarters)
        opt.reserved[8] = int(slots_a)
        # wb_min: tree levels with at least this many blocks are solved warp-per-block (0 = default 64, -1 = never)
        opt.reserved[9] = int(far_lanes2) | (int(sep_tile) << 8) | ((0xFFFF if wb_min < 0 else int(wb_min)) << 16)
        # dense-panel levels: dp_min_rows = rows of a separator level's longest block from which the level is solved through
        # inverted panels (0 = default 128, -1 = never; multiples of 32); dp_panel = panel rows (0 = from the level's shape);
        # dp_leaf = the leaf level too
        # dp_max_blocks: levels with more blocks are not dense-panel levels (0 = default, -1 = no limit, else a power of two)
        mb = 0 if dp_max_blocks == 0 else 63 if dp_max_blocks < 0 else min(31, int(dp_max_blocks).bit_length())
        opt.reserved[6] = (int(bool(plain_launch)) | (2 if dp_leaf else 0) | (mb << 2) | ((255 if dp_min_rows < 0 else min(254, int(dp_min_rows) // 32)) << 8)
                           | (int(dp_panel) << 16))
        rc = self._L.rcg_create_with_options(C.byref(self._h), int(device), C.byref(opt))
        if rc != 0:
            raise RcgError(rc, (self._L.rcg_last_error(None) or b"").decode())
        self.N = 0

    def _check(self, rc):
        if rc != 0:
            raise RcgError(rc, (self._L.rcg_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.rcg_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_matrix(self, rowPtr, colIdx, val):
        rp = _u64(rowPtr)
        self.N = rp.shape[0] - 1
        self._check(self._L.rcg_set_matrix(self._h, self.N, rp, _u64(colIdx), _f64(val)))

    def update_matrix_values(self, val):
        """New values on the SAME sparsity pattern as the last set_matrix (the reference's reuse flow): 8 B per entry."""
        v = _f64(val)
        self._check(self._L.rcg_update_matrix_values(self._h, v.shape[0], v))

    # ---- permutation steps either side of the path, on the device (reference: util.cpp:16-57, util.hpp:147-155) ----
    def set_matrix_permuted(self, rowPtr, colIdx, val, P):
        """A in its ORIGINAL ordering + the permutation of rchol(A,G,P,threads): the handle holds A(P,P), rows re-sorted."""
        rp = _u64(rowPtr)
        self.N = rp.shape[0] - 1
        self._check(self._L.rcg_set_matrix_permuted(self._h, self.N, rp, _u64(colIdx), _f64(val), _u64(P)))

    def set_permutation(self, P):
        p = _u64(P)
        self.N = p.shape[0]
        self._check(self._L.rcg_set_permutation(self._h, self.N, p))

    def permute(self, x):
        out = np.empty(self.N, np.float64)
        self._check(self._L.rcg_permute_vector(self._h, _f64(x), out))
        return out

    def unpermute(self, xp):
        out = np.empty(self.N, np.float64)
        self._check(self._L.rcg_unpermute_vector(self._h, _f64(xp), out))
        return out

    def pcg_original(self, b, tol: float, maxit: int):
        """b and x in the ORIGINAL ordering (permuted / un-permuted on the device)."""
        x = np.empty(self.N, np.float64)
        relres, itr = C.c_double(0), C.c_int(0)
        self._check(self._L.rcg_pcg_original(self._h, _f64(b), float(tol), int(maxit), x, C.byref(relres), C.byref(itr)))
        return x, relres.value, itr.value

    def get_matrix(self):
        st = self.stats()
        rp = np.empty(self.N + 1, np.uint64); ci = np.empty(st["nnzA"], np.uint64); v = np.empty(st["nnzA"], np.float64)
        self._check(self._L.rcg_get_matrix(self._h, rp, ci, v))
        return rp, ci, v

    def set_factor(self, rowPtr, colIdx, val, part: Optional[np.ndarray] = None):
        rp = _u64(rowPtr)
        self.N = rp.shape[0] - 1
        if part is not None and len(part) >= 2:
            pa = _u64(part)
            pptr, plen = pa.ctypes.data_as(C.c_void_p), pa.shape[0]
        else:
            pa, pptr, plen = None, None, 0
        self._check(self._L.rcg_set_factor(self._h, self.N, rp, _u64(colIdx), _f64(val), pptr, plen))

    def set_factor_blocks(self, rowPtr, colIdx, val, bounds, depth):
        rp = _u64(rowPtr)
        self.N = rp.shape[0] - 1
        d = np.ascontiguousarray(depth, dtype=np.int32)
        self._check(self._L.rcg_set_factor_blocks(self._h, self.N, rp, _u64(colIdx), _f64(val), _u64(bounds), d, d.shape[0]))

    def dist_init(self, nranks: int, rank: int, unique_id: bytes, n_sub: int, top_depth: int):
        buf = C.create_string_buffer(unique_id, 128)
        self._check(self._L.rcg_dist_init(self._h, int(nranks), int(rank), buf, int(n_sub), int(top_depth)))

    def spmv(self, x):
        y = np.empty(self.N, np.float64)
        self._check(self._L.rcg_spmv(self._h, _f64(x), y))
        return y

    def trsv(self, which: int, rhs):
        out = np.empty(self.N, np.float64)
        self._check(self._L.rcg_trsv(self._h, int(which), _f64(rhs), out))
        return out

    def precond(self, r):
        z = np.empty(self.N, np.float64)
        self._check(self._L.rcg_precond(self._h, _f64(r), z))
        return z

    def pcg(self, b, tol: float, maxit: int):
        x = np.empty(self.N, np.float64)
        relres, itr = C.c_double(0), C.c_int(0)
        self._check(self._L.rcg_pcg(self._h, _f64(b), float(tol), int(maxit), x, C.byref(relres), C.byref(itr)))
        return x, relres.value, itr.value

    def set_rhs(self, b):
        self._check(self._L.rcg_set_rhs(self._h, _f64(b)))

    def pcg_resident(self, tol: float, maxit: int):
        relres, itr = C.c_double(0), C.c_int(0)
        self._check(self._L.rcg_pcg_resident(self._h, float(tol), int(maxit), C.byref(relres), C.byref(itr)))
        return relres.value, itr.value

    def solution(self):
        x = np.empty(self.N, np.float64)
        self._check(self._L.rcg_get_solution(self._h, x))
        return x

    def history(self):
        n = C.c_int(0)
        self._check(self._L.rcg_get_history(self._h, None, 0, C.byref(n)))
        h = np.zeros(max(n.value, 1), np.float64)
        self._check(self._L.rcg_get_history(self._h, h.ctypes.data_as(C.c_void_p), n.value, C.byref(n)))
        return h[: n.value]

    def stats(self) -> dict:
        s = Stats()
        self._check(self._L.rcg_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def profile_iteration(self, reps: int = 3) -> dict:
        self._check(self._L.rcg_profile_iteration(self._h, int(reps)))
        return self.stats()

    def groups(self, direction: int):
        """Dependency groups (tree levels) of one solve direction: list of dicts."""
        n = C.c_int(0)
        self._check(self._L.rcg_get_group_count(self._h, int(direction), C.byref(n)))
        out = []
        for g in range(n.value):
            info = (C.c_uint64 * 6)()
            self._check(self._L.rcg_get_group_info(self._h, int(direction), g, info))
            out.append(dict(blocks=info[0], rows=info[1], loc_nnz=info[2], ext_nnz=info[3], max_rows=info[4], max_stage=info[5]))
        return out

    def time_group(self, direction: int, group: int, kernel: int, reps: int = 3) -> float:
        ms = C.c_double(0)
        self._check(self._L.rcg_time_group(self._h, int(direction), int(group), int(kernel), int(reps), C.byref(ms)))
        return ms.value

    def debug_trace(self, which: int, rhs):
        out = np.empty(self.N, np.float64)
        tr = np.zeros((self.N, 4), np.uint32)
        self._check(self._L.rcg_debug_trace(self._h, int(which), _f64(rhs), out, tr))
        return out, tr

    def counters(self):
        out = (C.c_uint64 * 16)()
        self._check(self._L.rcg_debug_counters(self._h, out))
        return [int(v) for v in out]

    def dp_trace(self):
        """[CTA][warp][16] clock64 marks of one hop of the last traced dense-panel launch (dbg bit 1)."""
        a = np.zeros(160 * 32 * 16, dtype=np.uint64)
        self._check(self._L.rcg_debug_dp_trace(self._h, a.ctypes.data_as(C.c_void_p), a.size))
        return a.reshape(160, 32, 16)

    def cluster_levels(self, direction: int) -> int:
        """Tree levels of one direction that are launched on the cluster chain (chain_mode 4); 0 = 32-row chain everywhere."""
        info = (C.c_uint64 * 16)()
        self._check(self._L.rcg_debug_blocked_info(self._h, int(direction), info))
        return int(info[15]) & 0xFFFFFFFF

    def blocked_layout(self, direction: int) -> dict:
        """Raw copy of the blocked triangular-solve layout of one direction (tests/blocked_emulator.py interprets it)."""
        info = (C.c_uint64 * 16)()
        self._check(self._L.rcg_debug_blocked_info(self._h, int(direction), info))
        keys = ["active", "nchunks", "ntiles", "nblocks", "bytesA", "bytesB", "far_nnz", "Kr", "E", "Dfar", "N", "nlevels", "Dfar_sep", "tile_sep", "E_sep"]
        out = {k: int(info[i]) for i, k in enumerate(keys)}
        out["fold"] = int(info[15]) >> 32
        out["wb_min"] = (out["tile_sep"] >> 16) & 0xFFFFFF
        out["Dfar_wb"] = out["tile_sep"] >> 40
        out["tile_sep"] &= 0xFFFF
        if not out["active"]:
            return out

        def grab(what, dtype, count):
            a = np.zeros(max(int(count), 1), dtype=dtype)
            self._check(self._L.rcg_debug_blocked_copy(self._h, int(direction), what, a.ctypes.data_as(C.c_void_p), a.nbytes))
            return a[: int(count)]

        out["offA"] = grab(0, np.int64, out["nchunks"] + 1)
        out["offB"] = grab(1, np.int64, out["nchunks"] + 1)
        out["blobA"] = grab(2, np.uint8, out["bytesA"])
        out["blobB"] = grab(3, np.uint8, out["bytesB"])
        out["far_rp"] = grab(4, np.int64, out["N"] + 1)
        out["far_col"] = grab(5, np.uint32, out["far_nnz"])
        out["far_val"] = grab(6, np.float64, out["far_nnz"])
        out["tile_need"] = grab(7, np.uint32, out["ntiles"])
        out["blocks"] = grab(8, np.uint32, out["nblocks"] * 8).reshape(-1, 8)
        out["levels"] = grab(9, np.uint64, out["nlevels"] * 10).reshape(-1, 10)
        return out

    def time_phase(self, phase: int, reps: int = 3) -> float:
        ms = C.c_double(0)
        self._check(self._L.rcg_time_phase(self._h, int(phase), int(reps), C.byref(ms)))
        return ms.value


def spmv_row_histogram(rowPtr):
    """(entries by row-length bucket [<=2, 3..5, 6..12, 13..24, >24], lanes per row) -- the SpMV plan of set_matrix.
    Host-only (no GPU needed)."""
    L = load()
    rp = _u64(rowPtr)
    hist = np.zeros(5, np.uint64)
    lanes = C.c_int(0)
    rc = L.rcg_spmv_row_histogram(rp.shape[0] - 1, rp, hist, C.byref(lanes))
    if rc != 0:
        raise RcgError(rc, "rcg_spmv_row_histogram: bad arguments")
    return hist, lanes.value


def detect_blocks(rowPtr, colIdx, cap: int = 1 << 16):
    """Blocks (bounds, depth) that set_factor derives from G when no `part` is given; (None, None) = one block.
    Host-only (no GPU needed)."""
    L = load()
    rp = _u64(rowPtr)
    N = rp.shape[0] - 1
    bounds = np.zeros(cap + 1, np.uint64)
    depth = np.zeros(cap, np.int32)
    nb = C.c_uint64(0)
    rc = L.rcg_detect_blocks(N, rp, _u64(colIdx), bounds, depth, cap, C.byref(nb))
    if rc != 0:
        raise RcgError(rc, "rcg_detect_blocks: more blocks than the caller's capacity")
    if nb.value == 0:
        return None, None
    return bounds[: nb.value + 1].copy(), depth[: nb.value].copy()


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = load().rcg_nccl_unique_id(buf)
    if rc != 0:
        raise RcgError(rc, "ncclGetUniqueId failed")
    return buf.raw


def pcg(A, b, tol, maxit, G, part=None, device: int = 0):
    """One-shot mirror of the reference entry point ``pcg(A, b, tol, maxit, G, x, relres, itr)``
    (/root/reference/c++/util/pcg.hpp:13-16): returns ``(x, relres, itr, stats)``."""
    L = load()
    N = A[0].shape[0] - 1
    x = np.empty(N, np.float64)
    relres, itr = C.c_double(0), C.c_int(0)
    st = Stats()
    if part is not None and len(part) >= 2:
        pa = _u64(part)
        pptr, plen = pa.ctypes.data_as(C.c_void_p), pa.shape[0]
    else:
        pa, pptr, plen = None, None, 0
    rc = L.rcg_pcg_oneshot(int(device), N, _u64(A[0]), _u64(A[1]), _f64(A[2]), _f64(b), float(tol), int(maxit),
                           _u64(G[0]), _u64(G[1]), _f64(G[2]), pptr, plen, x, C.byref(relres), C.byref(itr),
                           C.byref(st))
    if rc != 0:
        raise RcgError(rc, (L.rcg_last_error(None) or b"").decode())
    return x, relres.value, itr.value, st.as_dict()
