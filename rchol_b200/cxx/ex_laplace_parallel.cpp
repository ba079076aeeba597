// Example driver with the reference's command line (c++/ex_laplace_parallel.cpp:11-52): -n <grid> -t <leaves>.
// Factorization = the reference code (host); solve = this repository's GPU path behind the reference's pcg API.
// Factor once / solve many: -save <file> writes Aperm, G, P, part, bperm (io.hpp); -load <file> skips generation and
// factorization and solves the stored problem.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>

#include "io.hpp"
#include "pcg.hpp"
#include "rchol_ref.hpp"
#include "sparse.hpp"
#include "util.hpp"

int main(int argc, char *argv[]) {
  int n = 3, threads = 2;
  double tol = 1e-6;
  int maxit = 200;
  for (int i = 1; i + 1 < argc; i++) {
    if (!strcmp(argv[i], "-n")) n = atoi(argv[i + 1]);
    if (!strcmp(argv[i], "-t")) threads = atoi(argv[i + 1]);
    if (!strcmp(argv[i], "-tol")) tol = atof(argv[i + 1]);
    if (!strcmp(argv[i], "-maxit")) maxit = atoi(argv[i + 1]);
  }
  const char *save_path = nullptr, *load_path = nullptr;
  for (int i = 1; i + 1 < argc; i++) {
    if (!strcmp(argv[i], "-save")) save_path = argv[i + 1];
    if (!strcmp(argv[i], "-load")) load_path = argv[i + 1];
  }
  std::cout << std::setprecision(3);
  if (load_path) {
    rchol_b200::Problem pr;
    rchol_b200::load_problem(load_path, pr);
    double relres;
    int itr;
    std::vector<double> x;
    pcg solver(pr.A, pr.b, tol, maxit, pr.G, pr.part, x, relres, itr);
    std::cout << "Loaded " << load_path << ": N = " << pr.A.size() << ", blocks = " << (pr.part.empty() ? 1 : pr.part.size() - 1)
              << std::endl;
    std::cout << "# CG iterations: " << itr << std::endl;
    std::cout << "Relative residual: " << relres << std::endl;
    return 0;
  }
  SparseCSR A;
  A = laplace_3d(n);
  std::vector<double> b(A.size());
  rand(b);

  SparseCSR G;
  std::vector<size_t> P;
  std::vector<size_t> part;
  rchol(A, G, P, part, threads);   // additive overload of rchol_ref.hpp: also returns the block boundaries
  std::cout << "Fill-in ratio: " << 2. * G.nnz() / A.nnz() << std::endl;

  SparseCSR Aperm;
  reorder(A, P, Aperm);
  std::vector<double> bperm;
  reorder(b, P, bperm);

  if (save_path) rchol_b200::save_problem(save_path, Aperm, G, P, part, bperm);

  double relres;
  int itr;
  std::vector<double> x;
  pcg solver(Aperm, bperm, tol, maxit, G, part, x, relres, itr);
  std::cout << "# CG iterations: " << itr << std::endl;
  std::cout << "Relative residual: " << relres << std::endl;
  // the same solve with the permutation steps on the device: A, b in the original ordering in, y in the original
  // ordering out (what python/ex_laplace_parallel.py:31-33 verifies on the host)
  std::vector<double> y;
  double relres2;
  int itr2;
  pcg solver2(A, b, tol, maxit, G, part, P, y, relres2, itr2);
  std::vector<double> yref;
  unpermute(x, P, yref);
  double diff = 0;
  for (size_t i = 0; i < y.size(); i++) diff = std::max(diff, std::abs(y[i] - yref[i]));
  std::cout << "Device-permuted solve: " << itr2 << " iterations, max |y - unpermute(x)| = " << diff << std::endl;
  std::cout << "GPU ms: upload " << solver.upload_ms << " analysis " << solver.analysis_ms << " iterations "
            << solver.solve_ms << " total " << solver.total_ms << std::endl;
  return 0;
}
