// ORACLE BUILD ONLY.  C entry point around the UNMODIFIED reference `pcg` class
// (/root/reference/c++/util/pcg.hpp:13-16, constructor = entry point), so that tests can run the real
// reference loop (with MKL's own SpMV / SpTRSV) on the same inputs as the restatement and the GPU path.
#include <cstdint>
#include <cstring>
#include <vector>
#include "sparse.hpp"
#include "pcg.hpp"

static SparseCSR wrap(uint64_t N, const uint64_t *rp, const uint64_t *ci, const double *v) {
  SparseCSR A;
  A.N = N;
  A.rowPtr = const_cast<size_t *>(reinterpret_cast<const size_t *>(rp));
  A.colIdx = const_cast<size_t *>(reinterpret_cast<const size_t *>(ci));
  A.val = const_cast<double *>(v);
  A.ownMemory = false;
  return A;
}

extern "C" int refpcg_run(uint64_t N, const uint64_t *Arp, const uint64_t *Aci, const double *Av, const double *b,
                          double tol, int maxit, const uint64_t *Grp, const uint64_t *Gci, const double *Gv,
                          double *x_out, double *relres, int *itr) {
  SparseCSR A = wrap(N, Arp, Aci, Av), G = wrap(N, Grp, Gci, Gv);
  std::vector<double> bv(b, b + N), x;
  double rr = 0;
  int it = 0;
  pcg(A, bv, tol, maxit, G, x, rr, it);
  memcpy(x_out, x.data(), N * sizeof(double));
  *relres = rr;
  *itr = it;
  return 0;
}
