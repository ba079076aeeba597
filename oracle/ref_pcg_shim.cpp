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

// Genuine-MKL triangular solves and SpMV with exactly the descriptors the reference uses (pcg.cpp:130-159), exposed so
// that golden vectors for the individual kernels can be generated (tests/golden/make_golden.py).
#include "mkl_spblas.h"
extern "C" int refmkl_kernels(uint64_t N, const uint64_t *Arp, const uint64_t *Aci, const double *Av,
                              const uint64_t *Grp, const uint64_t *Gci, const double *Gv, const double *r,
                              double *Ar_out, double *y_out, double *z_out) {
  std::vector<size_t> pb(N + 1), pe(N + 1);
  sparse_matrix_t Am, Gm;
  for (uint64_t i = 0; i < N; i++) { pb[i] = Arp[i]; pe[i] = Arp[i + 1]; }
  mkl_sparse_d_create_csr(&Am, SPARSE_INDEX_BASE_ZERO, N, N, pb.data(), pe.data(), (size_t *)Aci, (double *)Av);
  std::vector<size_t> gb(N + 1), ge(N + 1);
  for (uint64_t i = 0; i < N; i++) { gb[i] = Grp[i]; ge[i] = Grp[i + 1]; }
  mkl_sparse_d_create_csr(&Gm, SPARSE_INDEX_BASE_ZERO, N, N, gb.data(), ge.data(), (size_t *)Gci, (double *)Gv);
  matrix_descr dg;
  dg.type = SPARSE_MATRIX_TYPE_GENERAL;
  mkl_sparse_d_mv(SPARSE_OPERATION_NON_TRANSPOSE, 1, Am, dg, r, 0, Ar_out);
  matrix_descr dt;
  dt.type = SPARSE_MATRIX_TYPE_TRIANGULAR;
  dt.mode = SPARSE_FILL_MODE_UPPER;
  dt.diag = SPARSE_DIAG_NON_UNIT;
  mkl_sparse_d_trsv(SPARSE_OPERATION_TRANSPOSE, 1, Gm, dt, r, y_out);
  mkl_sparse_d_trsv(SPARSE_OPERATION_NON_TRANSPOSE, 1, Gm, dt, y_out, z_out);
  mkl_sparse_destroy(Am);
  mkl_sparse_destroy(Gm);
  return 0;
}
