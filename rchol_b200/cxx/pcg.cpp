// Host side of the drop-in `pcg` class: argument checks, then one call through the C ABI.
#include "pcg.hpp"

#include <cstdint>
#include <stdexcept>
#include <string>

#include "../../include/rchol_b200.h"

static_assert(sizeof(size_t) == sizeof(uint64_t), "SparseCSR indices must be 64-bit");

namespace {
int g_device = 0;
}

void pcg::set_device(int ordinal) { g_device = ordinal; }

pcg::pcg(const SparseCSR &A, const std::vector<double> &b, double tol, int maxit, const SparseCSR &G,
         std::vector<double> &x, double &relres, int &itr) {
  run(A, b, tol, maxit, G, nullptr, 0, x, relres, itr);
}

pcg::pcg(const SparseCSR &A, const std::vector<double> &b, double tol, int maxit, const SparseCSR &G,
         const std::vector<size_t> &part, std::vector<double> &x, double &relres, int &itr) {
  run(A, b, tol, maxit, G, part.data(), part.size(), x, relres, itr);
}

pcg::pcg(const SparseCSR &A, const std::vector<double> &b, double tol, int maxit, const SparseCSR &G,
         const std::vector<size_t> &part, const std::vector<size_t> &P, std::vector<double> &x, double &relres, int &itr) {
  if (A.N == 0 || A.N != G.N) throw std::invalid_argument("pcg: A and G must be non-empty and of equal size");
  if (b.size() != A.N || P.size() != A.N) throw std::invalid_argument("pcg: b and P must have one entry per row of A");
  x.resize(A.N);
  rcg_handle *h = nullptr;
  int rc = rcg_create(&h, g_device);
  if (rc != RCG_OK)
    throw std::runtime_error(std::string("rchol_b200 pcg failed (") + std::to_string(rc) + "): " + rcg_last_error(nullptr));
  auto u64 = [](const size_t *p) { return reinterpret_cast<const uint64_t *>(p); };
  rc = rcg_set_matrix_permuted(h, A.N, u64(A.rowPtr), u64(A.colIdx), A.val, u64(P.data()));
  if (rc == RCG_OK) rc = rcg_set_factor(h, G.N, u64(G.rowPtr), u64(G.colIdx), G.val, part.empty() ? nullptr : u64(part.data()), part.size());
  if (rc == RCG_OK) rc = rcg_pcg_original(h, b.data(), tol, maxit, x.data(), &relres, &itr);
  const std::string msg = rc == RCG_OK ? "" : rcg_last_error(h);
  rcg_stats st;
  rcg_get_stats(h, &st);
  rcg_destroy(h);
  if (rc != RCG_OK) throw std::runtime_error(std::string("rchol_b200 pcg failed (") + std::to_string(rc) + "): " + msg);
  upload_ms = st.upload_ms;
  analysis_ms = st.analysis_ms;
  solve_ms = st.solve_ms;
  total_ms = st.total_ms;
}

void pcg::run(const SparseCSR &A, const std::vector<double> &b, double tol, int maxit, const SparseCSR &G,
              const size_t *part, size_t npart, std::vector<double> &x, double &relres, int &itr) {
  if (A.N == 0 || A.N != G.N) throw std::invalid_argument("pcg: A and G must be non-empty and of equal size");
  if (b.size() != A.N) throw std::invalid_argument("pcg: right-hand side length differs from the matrix size");
  x.resize(A.N);   // pcg.cpp:67
  rcg_stats st;
  const int rc = rcg_pcg_oneshot(g_device, A.N, reinterpret_cast<const uint64_t *>(A.rowPtr),
                                 reinterpret_cast<const uint64_t *>(A.colIdx), A.val, b.data(), tol, maxit,
                                 reinterpret_cast<const uint64_t *>(G.rowPtr),
                                 reinterpret_cast<const uint64_t *>(G.colIdx), G.val,
                                 reinterpret_cast<const uint64_t *>(part), npart, x.data(), &relres, &itr, &st);
  if (rc != RCG_OK)
    throw std::runtime_error(std::string("rchol_b200 pcg failed (") + std::to_string(rc) + "): " + rcg_last_error(nullptr));
  upload_ms = st.upload_ms;
  analysis_ms = st.analysis_ms;
  solve_ms = st.solve_ms;
  total_ms = st.total_ms;
}
