/* ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's solve-phase hot path, /root/reference/c++/util/pcg.cpp, with the
 * Intel-MKL calls (a third-party dependency whose source is not in the reference tree; version unpinned,
 * c++/Makefile:6 `-lmkl_intel_ilp64 -mkl`) replaced by plain C loops that restate the documented
 * semantics of those calls.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may call into this file; the CUDA product path never does.
 *
 * PARITY PINNING: the reference ships no tests, golden vectors or fixtures for this path (SURVEY.md
 * section 4), so the pins are (1) the 3x3 known-answer vectors in tests/golden (checked against real
 * oneMKL during the survey), and (2) oracle/_ref/libpcg_ref.so = the UNMODIFIED reference pcg.cpp compiled
 * here against the genuine oneMKL 2024.2 kernels that libtorch_cpu.so exports (oracle/Makefile,
 * oracle/mklshim/), against which tests/test_oracle.py checks this restatement (x, relres, iteration
 * count) on seeded problems.
 *
 * Index semantics follow pcg.cpp:31-54: zero-based CSR, row i occupies [rowPtr[i], rowPtr[i+1]),
 * 64-bit unsigned indices exactly as in SparseCSR (c++/sparse.hpp:10-31).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* q = A p : mkl_sparse_d_mv(NON_TRANSPOSE, 1, A, GENERAL, p, 0, q)   pcg.cpp:130-138 */
void oracle_spmv(uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val,
                 const double *p, double *q) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)N; i++) {
    double s = 0.0;
    for (uint64_t k = rowPtr[i]; k < rowPtr[i + 1]; k++) s += val[k] * p[colIdx[k]];
    q[i] = s;
  }
}

/* Forward solve U^T y = b : mkl_sparse_d_trsv(TRANSPOSE, 1, U, {TRIANGULAR, UPPER, NON_UNIT}, b, y)
 * pcg.cpp:146-151.  U is CSR; its transpose is traversed by columns, i.e. a right-looking scatter.
 * TRIANGULAR+UPPER means only entries with col >= row take part.  Sequential, like un-optimised MKL
 * (the reference never calls mkl_sparse_optimize). */
void oracle_trsv_upper_transposed(uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx,
                                  const double *val, const double *b, double *y) {
  memcpy(y, b, N * sizeof(double));
  for (uint64_t i = 0; i < N; i++) {
    double d = 0.0;
    for (uint64_t k = rowPtr[i]; k < rowPtr[i + 1]; k++)
      if (colIdx[k] == i) { d = val[k]; break; }
    double yi = y[i] / d;
    y[i] = yi;
    for (uint64_t k = rowPtr[i]; k < rowPtr[i + 1]; k++) {
      uint64_t c = colIdx[k];
      if (c > i) y[c] -= val[k] * yi;
    }
  }
}

/* Backward solve U z = y : mkl_sparse_d_trsv(NON_TRANSPOSE, ...)   pcg.cpp:154-155.  Row gather. */
void oracle_trsv_upper(uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val,
                       const double *y, double *z) {
  for (uint64_t ii = N; ii-- > 0;) {
    double s = y[ii], d = 0.0;
    for (uint64_t k = rowPtr[ii]; k < rowPtr[ii + 1]; k++) {
      uint64_t c = colIdx[k];
      if (c == ii) d = val[k];
      else if (c > ii) s -= val[k] * z[c];
    }
    z[ii] = s / d;
  }
}

/* ret = U^{-1} U^{-T} b   pcg.cpp:141-159 (the reference allocates and zeroes a scratch vector per call) */
void oracle_precond(uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val,
                    const double *b, double *scratch, double *ret) {
  oracle_trsv_upper_transposed(N, rowPtr, colIdx, val, b, scratch);
  oracle_trsv_upper(N, rowPtr, colIdx, val, scratch, ret);
}

/* --- CBLAS level-1 restatements (pcg.cpp:71,82,89,93-96,101-108,117-118) ------------------------ */
static double o_dot(uint64_t n, const double *a, const double *b) {
  double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
  for (int64_t i = 0; i < (int64_t)n; i++) s += a[i] * b[i];
  return s;
}
static double o_nrm2(uint64_t n, const double *a) { return sqrt(o_dot(n, a, a)); }
static void o_axpy(uint64_t n, double alpha, const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n; i++) y[i] += alpha * x[i];
}
static void o_scal(uint64_t n, double alpha, double *x) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n; i++) x[i] *= alpha;
}
static void o_copy(uint64_t n, const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n; i++) y[i] = x[i];
}

double oracle_dot(uint64_t n, const double *a, const double *b) { return o_dot(n, a, b); }
double oracle_nrm2(uint64_t n, const double *a) { return o_nrm2(n, a); }

/* pcg::iteration  pcg.cpp:57-127, statement for statement:
 *   x0 = 0 assumed (r = b, :70-71); loop while ||r|| > ||b|| tol && it < maxit (:82);
 *   z = M^{-1} r (:85); first: p = z (:89) else beta = (r.z)/(r_prev.z_prev) (:93-94), p = beta p + z (:95-96);
 *   q = A p (:100); alpha = (p.r)/(p.q) (:101-103); x += alpha p (:104); prev copies (:106-107);
 *   r -= alpha q (:108); finally relres = ||A x - b|| / ||b|| (TRUE residual, :116-118), itr (:119).
 * timings[0..3] (optional) accumulate seconds in: trsv pair, spmv, blas1, total loop.
 * hist (optional, length maxit+1) receives ||r_k|| / ||b|| of the recurrence residual at every loop test. */
int oracle_pcg(uint64_t N, const uint64_t *ArowPtr, const uint64_t *AcolIdx, const double *Aval,
               const double *b, double tol, int maxit, const uint64_t *GrowPtr,
               const uint64_t *GcolIdx, const double *Gval, double *x, double *relres, int *itr,
               double *timings, double *hist) {
  double *r = (double *)calloc(N, sizeof(double));
  double *prev_r = (double *)calloc(N, sizeof(double));
  double *prev_cond = (double *)calloc(N, sizeof(double));
  double *p = (double *)calloc(N, sizeof(double));
  double *temp = (double *)calloc(N, sizeof(double));
  double *q = (double *)calloc(N, sizeof(double));
  double *scratch = (double *)calloc(N, sizeof(double));
  if (!r || !prev_r || !prev_cond || !p || !temp || !q || !scratch) return 1;
  double t_trsv = 0, t_spmv = 0, t_blas = 0, t0, t_begin = now_s();
  memset(x, 0, N * sizeof(double));
  o_copy(N, b, r);
  int n_iters = 0;
  for (;;) {
    t0 = now_s();
    double nr = o_nrm2(N, r), nb = o_nrm2(N, b);
    t_blas += now_s() - t0;
    if (hist) hist[n_iters] = nr / nb;
    if (!(nr > nb * tol && n_iters < maxit)) break;

    t0 = now_s();
    oracle_precond(N, GrowPtr, GcolIdx, Gval, r, scratch, temp);
    t_trsv += now_s() - t0;

    t0 = now_s();
    if (n_iters == 0) {
      o_copy(N, temp, p);
    } else {
      double d1 = o_dot(N, r, temp);
      double d2 = o_dot(N, prev_r, prev_cond);
      o_scal(N, d1 / d2, p);
      o_axpy(N, 1.0, temp, p);
    }
    t_blas += now_s() - t0;

    t0 = now_s();
    oracle_spmv(N, ArowPtr, AcolIdx, Aval, p, q);
    t_spmv += now_s() - t0;

    t0 = now_s();
    double d1 = o_dot(N, p, r);
    double d2 = o_dot(N, p, q);
    double alpha = d1 / d2;
    o_axpy(N, alpha, p, x);
    o_copy(N, r, prev_r);
    o_copy(N, temp, prev_cond);
    o_axpy(N, -alpha, q, r);
    t_blas += now_s() - t0;
    n_iters++;
  }
  double t_loop = now_s() - t_begin;
  oracle_spmv(N, ArowPtr, AcolIdx, Aval, x, q);
  o_axpy(N, -1.0, b, q);
  *relres = o_nrm2(N, q) / o_nrm2(N, b);
  *itr = n_iters;
  if (timings) { timings[0] = t_trsv; timings[1] = t_spmv; timings[2] = t_blas; timings[3] = t_loop; }
  free(r); free(prev_r); free(prev_cond); free(p); free(temp); free(q); free(scratch);
  return 0;
}
