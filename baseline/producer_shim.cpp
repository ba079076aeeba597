// Producer shim: a C ABI around the UNMODIFIED reference factorization, compiled from the sources
// where they lie under /root/reference/c++ (see baseline/Makefile).  It produces the fixed inputs of
// the hot path -- the factor G (CSR of the upper-triangular U), the nested-dissection permutation P
// and the block boundaries `part` (the reference's local `result_idx`,
// /root/reference/c++/rchol/rchol_parallel.cpp:64-70) -- with a fixed seed, so that the GPU path,
// the oracle and the CPU baseline all consume the identical G and P.
//
// Nothing here is product code and nothing here is on the solve path.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <exception>
#include <string>
#include <unistd.h>
#include <fcntl.h>

#include "sparse.hpp"
#include "rchol.hpp"
#include "rchol_parallel.hpp"
#include "find_separator.hpp"
#include "util.hpp"

extern "C" unsigned rchol_b200_fixed_seed = 20240u;

// ---- hook that rchol_parallel.cpp calls instead of find_separator (see hook.h) -------------------
static std::vector<size_t> g_last_part_sizes;
Separator_info rchol_b200_find_separator_hook(const SparseCSR &A, int depth, int target) {
  Separator_info s = find_separator(A, depth, target);
  g_last_part_sizes.assign(s.val->begin(), s.val->end());
  return s;
}

namespace {
struct Factor {
  SparseCSR G;                 // arrays new[]-ed by the reference, ownMemory=false
  std::vector<size_t> P;       // empty for the sequential API
  std::vector<size_t> part;    // block boundaries in permuted index space, last == N
  std::string err;
};

struct StdoutSilencer {        // the reference prints timings unconditionally
  int saved = -1;
  StdoutSilencer() {
    fflush(stdout);
    saved = dup(1);
    int nul = open("/dev/null", O_WRONLY);
    if (nul >= 0) { dup2(nul, 1); close(nul); }
  }
  ~StdoutSilencer() {
    fflush(stdout);
    if (saved >= 0) { dup2(saved, 1); close(saved); }
  }
};
}  // namespace

extern "C" {

// threads == 0  -> sequential API  rchol(A, G)            (/root/reference/c++/rchol/rchol.cpp:7)
// threads >= 1  -> parallel  API   rchol(A, G, P, threads) (/root/reference/c++/rchol/rchol_parallel.cpp:37)
// Runs on a fresh std::thread so that (a) the reference's thread_local generators are re-seeded on
// every call and (b) its sched_setaffinity(cpu 0) does not stick to the caller's thread.
void *refprod_factor(uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val,
                     int threads, unsigned seed, int quiet) {
  Factor *f = new Factor();
  rchol_b200_fixed_seed = seed;
  std::thread worker([&]() {
    try {
      SparseCSR A;
      A.N = N;
      A.rowPtr = const_cast<size_t *>(reinterpret_cast<const size_t *>(rowPtr));
      A.colIdx = const_cast<size_t *>(reinterpret_cast<const size_t *>(colIdx));
      A.val = const_cast<double *>(val);
      A.ownMemory = false;
      StdoutSilencer *sil = quiet ? new StdoutSilencer() : nullptr;
      try {
        if (threads <= 0) {
          rchol(A, f->G);
          f->part = {0, (size_t)N};
        } else {
          g_last_part_sizes.clear();
          rchol(A, f->G, f->P, threads);
          f->part.assign(1, 0);
          for (size_t s : g_last_part_sizes) f->part.push_back(f->part.back() + s);
        }
      } catch (...) { delete sil; throw; }
      delete sil;
    } catch (const std::exception &e) {
      f->err = e.what();
    }
  });
  worker.join();
  return f;
}

// Partition boundaries (part[0]=0 ... part[2T-1]=N) of the last parallel factorization, for C++ callers of the
// reference API `rchol(A, G, P, threads)` which does not return them.  Returns the number of boundaries.
uint64_t refprod_last_part(uint64_t *out, uint64_t capacity) {
  uint64_t n = g_last_part_sizes.size() + 1, acc = 0;
  if (out && capacity > 0) out[0] = 0;
  for (uint64_t i = 0; i + 1 < n; i++) {
    acc += g_last_part_sizes[i];
    if (out && i + 1 < capacity) out[i + 1] = acc;
  }
  return n;
}

const char *refprod_error(void *h) {
  Factor *f = static_cast<Factor *>(h);
  return f->err.empty() ? nullptr : f->err.c_str();
}
uint64_t refprod_G_n(void *h) { return static_cast<Factor *>(h)->G.N; }
uint64_t refprod_G_nnz(void *h) { return static_cast<Factor *>(h)->G.nnz(); }
uint64_t refprod_P_len(void *h) { return static_cast<Factor *>(h)->P.size(); }
uint64_t refprod_part_len(void *h) { return static_cast<Factor *>(h)->part.size(); }
const uint64_t *refprod_G_rowptr(void *h) { return (const uint64_t *)static_cast<Factor *>(h)->G.rowPtr; }
const uint64_t *refprod_G_colidx(void *h) { return (const uint64_t *)static_cast<Factor *>(h)->G.colIdx; }
const double *refprod_G_val(void *h) { return static_cast<Factor *>(h)->G.val; }
const uint64_t *refprod_P(void *h) { return (const uint64_t *)static_cast<Factor *>(h)->P.data(); }
const uint64_t *refprod_part(void *h) { return (const uint64_t *)static_cast<Factor *>(h)->part.data(); }

void refprod_free(void *h) {
  Factor *f = static_cast<Factor *>(h);
  if (!f) return;
  // G's arrays are leaked by design in the reference (rchol_lap.cpp:436-438 + sparse.cpp:10); free them here.
  delete[] f->G.rowPtr;
  delete[] f->G.colIdx;
  delete[] f->G.val;
  f->G.N = 0;
  delete f;
}

// The reference's own generators / permutation helpers, exposed for cross-checking our restatements.
// laplace_3d: /root/reference/c++/util/laplace_3d.hpp:8-65 ; reorder: /root/reference/c++/util/util.cpp:16-57
uint64_t refprod_laplace3d_nnz(int n) { uint64_t m = n; return 7 * m * m * m - 6 * m * m; }
void refprod_laplace3d(int n, uint64_t *rowPtr, uint64_t *colIdx, double *val) {
  SparseCSR A = laplace_3d(n);   // ownMemory=false: arrays leak unless we free them
  uint64_t N = A.N, nnz = A.nnz();
  memcpy(rowPtr, A.rowPtr, (N + 1) * sizeof(uint64_t));
  memcpy(colIdx, A.colIdx, nnz * sizeof(uint64_t));
  memcpy(val, A.val, nnz * sizeof(double));
  delete[] A.rowPtr; delete[] A.colIdx; delete[] A.val;
}
void refprod_reorder(uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val,
                     const uint64_t *P, uint64_t *outRowPtr, uint64_t *outColIdx, double *outVal) {
  SparseCSR A;
  A.N = N;
  A.rowPtr = const_cast<size_t *>(reinterpret_cast<const size_t *>(rowPtr));
  A.colIdx = const_cast<size_t *>(reinterpret_cast<const size_t *>(colIdx));
  A.val = const_cast<double *>(val);
  A.ownMemory = false;
  std::vector<size_t> perm(P, P + N), rp, ci;
  std::vector<double> v;
  reorder(A, rp, ci, v, perm);
  memcpy(outRowPtr, rp.data(), (N + 1) * sizeof(uint64_t));
  memcpy(outColIdx, ci.data(), ci.size() * sizeof(uint64_t));
  memcpy(outVal, v.data(), v.size() * sizeof(double));
}

}  // extern "C"
