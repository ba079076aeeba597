// pcg -- drop-in for the reference's solve-phase entry point (/root/reference/c++/util/pcg.hpp:13-16).
// Same header name, class name and constructor signature, so the reference's example drivers
// (c++/ex_laplace.cpp:42, c++/ex_laplace_parallel.cpp:46) compile unchanged against it; the private MKL-typed
// members of the reference class are gone (they are not API).  The work is done on a B200 through the C ABI in
// include/rchol_b200.h; there is no CPU fallback -- a missing GPU surfaces as std::runtime_error.
//
// Semantics kept (pcg.cpp:57-127): zero initial guess, loop while ||r|| > tol*||b|| && it < maxit, `relres` is the
// TRUE residual ||Ax-b||/||b||, `itr` the completed iterations, x is resized to N and returned in the (permuted)
// ordering of A.
#ifndef pcg_hpp
#define pcg_hpp

#include <vector>

#include "sparse.hpp"

class pcg {
public:
  pcg(const SparseCSR &A, const std::vector<double> &b, double tol, int maxit, const SparseCSR &G,
      std::vector<double> &x, double &relres, int &itr);

  // Additive overload: `part` = the nested-dissection block boundaries (2T entries, part[0]=0, part.back()=N) that
  // the reference computes in rchol(A,G,P,threads) (rchol_parallel.cpp:64-70) but does not return from its C++ API
  // (its Python/MATLAB bindings do).  With it the block structure is taken as given; without it (the stock signature above) it is recovered from G.
  pcg(const SparseCSR &A, const std::vector<double> &b, double tol, int maxit, const SparseCSR &G,
      const std::vector<size_t> &part, std::vector<double> &x, double &relres, int &itr);

  // Additive overload for the whole ex_laplace_parallel flow: A and b in their ORIGINAL ordering plus the permutation P
  // that rchol(A,G,P,threads) returned.  reorder(A,P,Aperm) (util.cpp:16-57), reorder(b,P,bperm) (util.hpp:147-155) and
  // the un-permutation of the solution (python/ex_laplace_parallel.py:31-32) run on the device; x comes back in the
  // ORIGINAL ordering, so that ||A x - b|| / ||b|| can be checked against the caller's own A and b.
  pcg(const SparseCSR &A, const std::vector<double> &b, double tol, int maxit, const SparseCSR &G,
      const std::vector<size_t> &part, const std::vector<size_t> &P, std::vector<double> &x, double &relres, int &itr);

  // measurements of the last solve (milliseconds), for drivers that want to print them
  double upload_ms = 0, analysis_ms = 0, solve_ms = 0, total_ms = 0;

  static void set_device(int ordinal);   // CUDA device used by subsequent solves (default 0)

private:
  void run(const SparseCSR &A, const std::vector<double> &b, double tol, int maxit, const SparseCSR &G,
           const size_t *part, size_t npart, std::vector<double> &x, double &relres, int &itr);
};

#endif
