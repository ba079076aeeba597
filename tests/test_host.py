"""CPU tests of the host side: generators, permutation helpers, factor structure, and the C-ABI library
(loads and exports every symbol the header declares -- no compute calls without a GPU)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, GOLDEN_CASES, load_golden, make_problem, needs_producer
from rchol_b200 import problems


def test_laplace_3d_closed_form_and_layout():
    for n in (1, 2, 3, 7):
        rp, ci, v = problems.laplace_3d(n)
        N = n ** 3
        assert rp.dtype == np.uint64 and ci.dtype == np.uint64 and v.dtype == np.float64
        assert rp.shape[0] == N + 1 and int(rp[-1]) == 7 * n ** 3 - 6 * n ** 2     # laplace_3d.hpp:60-64
        for i in range(N):
            cols = ci[int(rp[i]):int(rp[i + 1])]
            assert np.all(np.diff(cols.astype(np.int64)) > 0) and i in cols
        import scipy.sparse as sp
        A = sp.csr_matrix((v, ci.astype(np.int64), rp.astype(np.int64)), shape=(N, N))
        assert abs(A - A.T).max() == 0 and np.all(A.diagonal() == 6.0)
    assert int(problems.laplace_3d(3)[0][-1]) == 135


@needs_producer
def test_generators_match_reference():
    from rchol_b200 import producer
    for n in (3, 6, 11):
        a, b = problems.laplace_3d(n), producer.ref_laplace_3d(n)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    A = problems.laplace_3d(9)
    P = np.random.default_rng(1).permutation(9 ** 3).astype(np.uint64)
    a, b = problems.reorder_matrix(*A, P), producer.ref_reorder(*A, P)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    x = np.arange(9 ** 3, dtype=np.float64)
    assert np.array_equal(problems.unpermute_vector(problems.reorder_vector(x, P), P), x)


def test_aniso_2d_is_sddm():
    import scipy.sparse as sp
    n = 17
    rp, ci, v = problems.aniso_2d(n)
    assert int(rp[-1]) == 5 * n * n - 4 * n
    A = sp.csr_matrix((v, ci.astype(np.int64), rp.astype(np.int64)), shape=(n * n, n * n))
    assert abs(A - A.T).max() < 1e-15
    off = A - sp.diags(A.diagonal())
    assert off.max() <= 0 and np.all(A.diagonal() > 0)
    rowsum = np.asarray(A.sum(axis=1)).ravel()
    assert np.all(rowsum > -1e-12) and rowsum[0] > 0       # boundary rows strictly dominant (ghost edges)
    # same generator, same seed -> same matrix
    assert all(np.array_equal(a, b) for a, b in zip(problems.aniso_2d(n), (rp, ci, v)))


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_factor_structure_of_goldens(name):
    """Invariants the GPU solve relies on: G is CSR of an upper-triangular matrix, rows sorted,
    diagonal first and positive, and a row of block B only touches B and B's ancestor separators."""
    g = load_golden(name)
    rp, ci, v = (g["G_rowPtr"].astype(np.int64), g["G_colIdx"].astype(np.int64), g["G_val"])
    N = rp.shape[0] - 1
    part = g["part"].astype(np.int64)
    nb = part.shape[0] - 1
    assert part[0] == 0 and part[-1] == N and ((nb + 1) & nb) == 0
    anc = {}

    def rec(start, total, above):
        if total == 1:
            anc[start] = above
        else:
            sep, half = start + total - 1, (total - 1) // 2
            anc[sep] = above
            rec(start, half, above + [sep])
            rec(start + half, half, above + [sep])
    rec(0, nb, [])
    blk = np.searchsorted(part, np.arange(N), side="right") - 1
    for i in range(N):
        cols = ci[rp[i]:rp[i + 1]]
        assert cols[0] == i and v[rp[i]] > 0 and np.all(np.diff(cols) > 0)
        allowed = set(anc[blk[i]] + [blk[i]])
        assert set(blk[cols]).issubset(allowed)


@needs_producer
def test_producer_is_deterministic_and_validates_threads():
    from rchol_b200 import producer
    A = problems.laplace_3d(10)
    f1, f2 = producer.factor(*A, threads=4, seed=7), producer.factor(*A, threads=4, seed=7)
    assert np.array_equal(f1.colIdx, f2.colIdx) and np.array_equal(f1.val, f2.val) and np.array_equal(f1.P, f2.P)
    assert sorted(f1.P.tolist()) == list(range(1000)) and f1.part[-1] == 1000 and len(f1.part) == 8
    f3 = producer.factor(*A, threads=4, seed=8)
    assert not np.array_equal(f1.val, f3.val) or not np.array_equal(f1.colIdx, f3.colIdx)
    with pytest.raises(ValueError):
        producer.factor(*A, threads=3)       # rchol_parallel.cpp:42-43


def test_bytes_per_iteration_formula():
    # BASELINE.md section 3
    assert problems.algorithmic_bytes_per_iteration(10, 50, 30) == 12 * 50 + 4 * 11 + 2 * (12 * 30 + 4 * 11) + 136 * 10


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "rchol_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rcg_[a-z_0-9]+)\s*\(", text)))


def test_c_abi_library_exports_every_declared_symbol():
    from rchol_b200 import capi
    assert os.path.exists(capi.LIB_PATH), "run `make` / __graft_entry__.build() first"
    lib = ctypes.CDLL(capi.LIB_PATH)
    declared = _header_symbols()
    assert len(declared) >= 19
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/rchol_b200.h but not exported"
    assert sorted(capi.EXPORTED) == declared
    lib.rcg_version.restype = ctypes.c_char_p
    assert b"rchol_b200" in lib.rcg_version() and b"sm_100a" in lib.rcg_version()


def test_spmv_plan_from_the_row_length_histogram():
    """rcg_spmv_row_histogram (host only): entries by row-length bucket and the lanes per row chosen from them -- the
    smallest lane count whose bucket edge covers 90 % of the ENTRIES.  The bench matrices keep the plan they were measured
    with (7-point Laplacian: 8 lanes, 5-point SDDM: 4); a few very long rows among many short ones do not change the plan
    of the bulk, a matrix whose entries sit in long rows gets a warp per row; the plan is invariant under permutation."""
    from rchol_b200 import capi
    for n in (8, 20, 40):
        rp = problems.laplace_3d(n)[0]
        hist, lanes = capi.spmv_row_histogram(rp)
        lens = np.diff(rp.astype(np.int64))
        assert int(hist.sum()) == int(rp[-1]) and lanes == 8
        assert int(hist[1]) == int(lens[(lens > 2) & (lens <= 5)].sum()) and int(hist[2]) == int(lens[(lens > 5) & (lens <= 12)].sum())
    for n in (24, 96):
        hist, lanes = capi.spmv_row_histogram(problems.aniso_2d(n)[0])
        assert lanes == 4 and int(hist[2:].sum()) == 0
    rng = np.random.default_rng(3)
    lens = np.full(10000, 3, np.int64)
    lens[:5] = 2000                                            # 5 rows hold 25 % of the entries: the short rows cover only 75 %,
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    assert capi.spmv_row_histogram(rp)[1] == 32                # less than 90 %, so the plan serves the long rows (a warp per row)
    lens[:5] = 300                                             # 5 % of the entries in long rows: plan of the bulk
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    assert capi.spmv_row_histogram(rp)[1] == 4
    perm = rng.permutation(lens.shape[0])
    rp2 = np.concatenate([[0], np.cumsum(lens[perm])]).astype(np.uint64)
    h1, l1 = capi.spmv_row_histogram(rp); h2, l2 = capi.spmv_row_histogram(rp2)
    assert l1 == l2 and np.array_equal(h1, h2)
    assert capi.spmv_row_histogram(np.array([0, 40, 80, 120], np.uint64))[1] == 32
    assert capi.spmv_row_histogram(np.array([0, 1, 3, 4], np.uint64))[1] == 2


def test_no_cpu_fallback_without_a_gpu():
    """On a box without a B200 creating a solver must fail loudly (RCG_ERR_CUDA), never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    from rchol_b200 import capi
    with pytest.raises(capi.RcgError) as e:
        capi.Solver(0)
    assert e.value.code == 1 and "no CPU fallback" in str(e.value)
    g = load_golden("lap3d_8_seq")
    with pytest.raises(capi.RcgError):
        capi.pcg(g["A"], g["b"], 1e-8, 10, g["G"])


def test_product_sources_never_touch_the_oracle():
    """The CUDA library and the host mirror must not reference oracle/ (tier rule 3)."""
    for sub in ("rchol_b200/csrc", "rchol_b200/cxx"):
        for fn in os.listdir(os.path.join(ROOT, sub)):
            if fn.endswith((".cu", ".cuh", ".cpp", ".hpp", ".h")):
                assert "oracle" not in open(os.path.join(ROOT, sub, fn)).read().lower(), fn
    for fn in ("capi.py", "problems.py", "producer.py", "__init__.py"):
        src = open(os.path.join(ROOT, "rchol_b200", fn)).read()
        assert "import oracle" not in src and "from oracle" not in src, fn


def test_cxx_mirror_compiles_against_reference_style_driver(tmp_path):
    """The reference's ex_laplace driver body compiles unchanged against our sparse.hpp / pcg.hpp."""
    src = tmp_path / "drv.cpp"
    src.write_text('''
#include "sparse.hpp"
#include "util.hpp"
#include "pcg.hpp"
int main() {
  SparseCSR A; A = laplace_3d(3);
  int N = A.size(); std::vector<double> b(N); rand(b);
  SparseCSR G(A);
  double tol = 1e-6; int maxit = 200; double relres; int itr; std::vector<double> x;
  if (N < 0) pcg(A, b, tol, maxit, G, x, relres, itr);   // constructor-as-entry-point, like ex_laplace.cpp:42
  // the other helpers of util.hpp:34-46,147-164 with the reference's call shapes (ex_laplace_parallel.cpp:38-39,
  // rchol_parallel.cpp:71, find_separator.cpp:145)
  std::vector<size_t> P(N); for (int i = 0; i < N; i++) P[i] = (size_t)((i * 5 + 3) % N);   // 5 and 27 coprime: a permutation
  SparseCSR Ap; reorder(A, P, Ap);
  std::vector<size_t> rp, ci; std::vector<double> v; reorder(A, rp, ci, v, P);
  bool same = rp.size() == (size_t)N + 1 && ci.size() == Ap.nnz();
  for (size_t k = 0; same && k < ci.size(); k++) same = ci[k] == Ap.colIdx[k] && v[k] == Ap.val[k];
  std::vector<double> bp; reorder(b, P, bp);
  std::vector<double> bq = reorder(b, P);
  std::vector<double> back; unpermute(bp, P, back);
  same = same && bq == bp && back == b && bp[0] == b[P[0]];
  print(P, "P");
  return (same && A.nnz() == 135 && G.nnz() == 135 && G.ownMemory) ? 0 : 1;
}
''')
    exe = tmp_path / "drv"
    lib = os.path.join(ROOT, "rchol_b200", "lib")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-o", str(exe), str(src), "-I", os.path.join(ROOT, "rchol_b200", "cxx"),
                           "-L", lib, "-lrchol_b200_cxx", "-lrchol_b200", f"-Wl,-rpath,{lib}"])
    assert subprocess.call([str(exe)]) == 0


def test_sdd_front_end_matches_the_matlab_definition_and_the_cxx_mirror(tmp_path):
    """sdd_to_sddm (matlab/rchol/sdd_to_sddm.m:2-17): Ae = [D+Neg, -Pos; -Pos, D+Neg]; the numpy statement, a scipy
    restatement of the MATLAB lines and the C++ mirror (rchol_b200/cxx/util.cpp) agree bit for bit; Ae is an SDDM."""
    import scipy.sparse as sp
    from rchol_b200 import problems
    A = problems.sdd_3d(5)
    N = A[0].shape[0] - 1
    As = sp.csr_matrix((A[2], A[1].astype(np.int64), A[0].astype(np.int64)), shape=(N, N))
    di = sp.diags(As.diagonal())
    nod = As - di
    pos, neg = nod.maximum(0), nod.minimum(0)
    assert pos.nnz > 0                                                        # really SDD, not SDDM
    ref = sp.bmat([[di + neg, -pos], [-pos, di + neg]]).tocsr()
    ref.eliminate_zeros(); ref.sort_indices()
    Ae = problems.sdd_to_sddm(*A)
    assert np.array_equal(Ae[0], ref.indptr) and np.array_equal(Ae[1], ref.indices) and np.array_equal(Ae[2], ref.data)
    off = ref - sp.diags(ref.diagonal())
    assert off.nnz == 0 or off.data.max() <= 0                               # M-matrix pattern
    assert np.asarray(ref.sum(axis=1)).min() >= -1e-12                       # diagonally dominant
    b = problems.random_rhs(N)
    assert np.array_equal(problems.sdd_recover(problems.sdd_rhs(b)), b)
    # C++ mirror: dump what sdd_to_sddm builds for the same matrix
    np.asarray(A[0]).tofile(tmp_path / "rp.bin"); np.asarray(A[1]).tofile(tmp_path / "ci.bin"); np.asarray(A[2]).tofile(tmp_path / "v.bin")
    src = tmp_path / "sdd.cpp"
    src.write_text('''
#include <cstdio>
#include <vector>
#include "sparse.hpp"
#include "util.hpp"
template <class T> std::vector<T> rd(const char *p) { FILE *f = fopen(p, "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); rewind(f);
  std::vector<T> v(n / sizeof(T)); if (fread(v.data(), 1, n, f) != (size_t)n) return {}; fclose(f); return v; }
template <class T> void wr(const char *p, const T *d, size_t n) { FILE *f = fopen(p, "wb"); fwrite(d, sizeof(T), n, f); fclose(f); }
int main() {
  auto rp = rd<size_t>("rp.bin"), ci = rd<size_t>("ci.bin"); auto v = rd<double>("v.bin");
  SparseCSR A(rp, ci, v, true), Ae;
  sdd_to_sddm(A, Ae);
  wr("erp.bin", Ae.rowPtr, Ae.N + 1); wr("eci.bin", Ae.colIdx, Ae.nnz()); wr("ev.bin", Ae.val, Ae.nnz());
  std::vector<double> b = {1, 2, 3}, be, x; sdd_rhs(b, be); sdd_recover(be, x);
  return (be.size() == 6 && be[4] == -2 && x == b) ? 0 : 1;
}
''')
    exe = tmp_path / "sdd"
    lib = os.path.join(ROOT, "rchol_b200", "lib")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-o", str(exe), str(src), "-I", os.path.join(ROOT, "rchol_b200", "cxx"),
                           "-L", lib, "-lrchol_b200_cxx", "-lrchol_b200", f"-Wl,-rpath,{lib}"])
    assert subprocess.call([str(exe)], cwd=tmp_path) == 0
    assert np.array_equal(np.fromfile(tmp_path / "erp.bin", np.uint64), Ae[0])
    assert np.array_equal(np.fromfile(tmp_path / "eci.bin", np.uint64), Ae[1])
    assert np.array_equal(np.fromfile(tmp_path / "ev.bin", np.float64), Ae[2])


def test_problem_container_round_trips_between_python_and_cxx(tmp_path):
    """rchol_b200/cxx/io.hpp <-> rchol_b200/problems.py: a golden problem written by Python is loaded
    and re-written by C++ byte for byte, and read back identically."""
    from conftest import load_golden
    from rchol_b200 import problems
    g = load_golden("lap3d_12_t4")
    src_file, dst_file = tmp_path / "p.rchb", tmp_path / "q.rchb"
    problems.save_problem(src_file, g["A"], g["G"], g["P"], g["part"], g["b"])
    src = tmp_path / "io.cpp"
    src.write_text('''
#include "io.hpp"
int main(int argc, char **argv) {
  rchol_b200::Problem p;
  rchol_b200::load_problem(argv[1], p);
  if (p.A.size() != 1728 || p.G.size() != 1728 || p.P.size() != 1728 || p.b.size() != 1728 || p.part.size() != 8) return 2;
  rchol_b200::save_problem(argv[2], p.A, p.G, p.P, p.part, p.b);
  try { rchol_b200::load_problem(argv[0], p); return 3; } catch (const std::exception &) {}   // not a problem file
  return 0;
}
''')
    exe = tmp_path / "io"
    lib = os.path.join(ROOT, "rchol_b200", "lib")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-o", str(exe), str(src), "-I", os.path.join(ROOT, "rchol_b200", "cxx"),
                           "-L", lib, "-lrchol_b200_cxx", "-lrchol_b200", f"-Wl,-rpath,{lib}"])
    assert subprocess.call([str(exe), str(src_file), str(dst_file)]) == 0
    assert open(src_file, "rb").read() == open(dst_file, "rb").read()
    d = problems.load_problem(dst_file)
    for k in range(3):
        assert np.array_equal(d["A"][k], g["A"][k]) and np.array_equal(d["G"][k], g["G"][k])
    assert np.array_equal(d["P"], g["P"]) and np.array_equal(d["part"], g["part"]) and np.array_equal(d["b"], g["b"])
    problems.save_problem(src_file, g["A"], g["G"])                      # optional arrays absent
    d = problems.load_problem(src_file)
    assert d["P"] is None and d["part"] is None and d["b"] is None
    with pytest.raises(ValueError):
        problems.load_problem(src)                                       # not a problem file


@pytest.mark.skipif(not os.path.isdir("/root/reference/c++"), reason="reference sources not mounted")
def test_unmodified_reference_mains_build_on_our_pcg_and_fail_loudly_without_a_gpu():
    """Drop-in boundary: /root/reference/c++/ex_laplace.cpp and ex_laplace_parallel.cpp compile UNCHANGED
    against rchol_b200/cxx/{sparse,util,pcg}.hpp (`make refmains`) and link our pcg class + the reference factorization.
    Without a GPU the pcg call must end in an error that names the missing device: there is no CPU fallback."""
    subprocess.check_call(["make", "refmains"], cwd=ROOT, stdout=subprocess.DEVNULL)
    import torch
    for exe, args in (("ref_ex_laplace", ["-n", "6"]), ("ref_ex_laplace_parallel", ["-n", "8", "-t", "2"])):
        path = os.path.join(ROOT, "baseline", "_ref", exe)
        assert os.path.exists(path)
        out = subprocess.run([path] + args, capture_output=True, text=True, timeout=300)
        assert "Fill-in ratio" in out.stdout                     # the reference factorization ran
        if not torch.cuda.is_available():
            assert out.returncode != 0 and "no CPU fallback" in out.stderr, (out.returncode, out.stderr[-300:])
        else:
            assert out.returncode == 0 and "Relative residual" in out.stdout, out.stderr[-300:]


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libpcg_ref.so")) and not os.path.isdir("/root/reference/c++"),
                    reason="reference pcg not built")
@needs_producer
def test_bench_reference_arm_prints_the_contract_line(tmp_path):
    """`bench.py --impl reference` (the arm the driver runs beside ours): one JSON line on stdout with the contract's keys,
    kind "reference" (the unmodified pcg in a forked child), zero host<->device bytes, no GPU launches."""
    import json
    import sys
    env = dict(os.environ, RCHOL_B200_CACHE=str(tmp_path))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "20", "--threads", "4",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "pcg_gbps_per_iter" and j["unit"] == "GB/s" and j["higher_is_better"]
    assert j["value"] > 0 and j["steps"] == 2 and j["warmup"] == 1 and j["n_gpus"] == 1 and j["gpu_launches"] == 0
    assert j["cpu_baseline"]["kind"] == "reference" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"] == dict(value=j["value"], unit="GB/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert j["config"]["workload"] == "lap3d_20^3_rchol_T4_pcg_tol1e-8"


@needs_producer
def test_bench_secondary_workload_runs_in_a_child_and_never_raises(tmp_path, monkeypatch):
    """bench.py measures BASELINE.json configs[3] (2-D anisotropic SDDM) in a child process with a time limit, so that a
    failure of the leg never costs the headline line: without a GPU the child fails loudly (no CPU fallback) and the parent
    gets an error object; a leg that exceeds its limit is reported the same way.  The problem itself (generator + reference
    factorization, cached under its own tag) is built either way."""
    import bench
    monkeypatch.setenv("RCHOL_B200_CACHE", str(tmp_path))
    out = bench.leg_in_subprocess("aniso2d", 48, 4, 1, 1, 300)
    assert isinstance(out, dict)
    if "error" in out:                                   # (no GPU here; on a B200 the leg succeeds)
        assert "status" in out["error"]
    else:
        assert out["workload"] == "aniso2d_48^2_rchol_T4_pcg_tol1e-8" and out["parity"]["fwd"] <= 1e-12
    assert os.path.exists(os.path.join(str(tmp_path), "aniso2d_48_T4_s20240.ready"))
    d, info = bench.build_problem(48, 4, "aniso2d")
    assert info["cached"] and d["A_rp"].shape[0] == 48 * 48 + 1 and int(d["part"][-1]) == 48 * 48
    out = bench.leg_in_subprocess("aniso2d", 48, 4, 1, 1, 0.05)
    assert "exceeded" in out["error"]


# ---------------------------------------------------------------------------------------------------------------------
# stock signature pcg(A, b, tol, maxit, G, x, relres, itr) (/root/reference/c++/util/pcg.hpp:13-16): no `part`
# ---------------------------------------------------------------------------------------------------------------------
def _check_block_rule(G, bounds, depth):
    """An entry (i, c) of U with block(c) != block(i) needs depth(block(c)) < depth(block(i)) -- the rule of
    rcg_set_factor_blocks (k_validate_depths on the device)."""
    rp, ci = np.asarray(G[0], np.int64), np.asarray(G[1], np.int64)
    rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
    bi = np.searchsorted(bounds, rows, side="right") - 1
    bc = np.searchsorted(bounds, ci, side="right") - 1
    off = bi != bc
    assert np.all(depth[bc[off]] < depth[bi[off]])


@pytest.mark.parametrize("name", ["lap3d_12_t4", "lap3d_10_t8_tol6", "aniso2d_24_t4"])
def test_blocks_detected_from_the_factor_equal_the_reference_partition(name):
    """Without `part` the nested-dissection blocks are recovered from G itself (rcg_detect_blocks, host only): on the
    reference-generated goldens they equal result_idx (rchol_parallel.cpp:64-70) and the tree depths of
    rchol_lap.cpp:254-261."""
    from rchol_b200 import capi
    from blocked_reference import tree_depths
    g = load_golden(name)
    bounds, depth = capi.detect_blocks(g["G"][0], g["G"][1])
    assert bounds is not None
    assert np.array_equal(bounds, g["part"])
    assert np.array_equal(depth, tree_depths(len(depth)))
    _check_block_rule(g["G"], bounds.astype(np.int64), depth)


def test_a_factor_that_is_one_chain_stays_one_block():
    from rchol_b200 import capi
    g = load_golden("lap3d_8_seq")
    assert capi.detect_blocks(g["G"][0], g["G"][1]) == (None, None)


@needs_producer
@pytest.mark.parametrize("kind,n,threads", [("lap3d", 40, 8), ("lap3d", 48, 16), ("lap3d", 48, 256), ("aniso2d", 160, 4)])
def test_detected_blocks_obey_the_block_rule(kind, n, threads):
    from rchol_b200 import capi
    A, b, G, part, f = make_problem(kind, n, threads)
    bounds, depth = capi.detect_blocks(G[0], G[1])
    assert bounds is not None and bounds[0] == 0 and bounds[-1] == f.N and np.all(np.diff(bounds.astype(np.int64)) > 0)
    _check_block_rule(G, bounds.astype(np.int64), depth)
    if threads == 8 or kind == "aniso2d":
        assert np.array_equal(bounds, part)          # the reference's own partition, exactly
