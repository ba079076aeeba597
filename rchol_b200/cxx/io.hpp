// On-disk container for the inputs of the solve phase: A (permuted), G, the permutation P, the nested-dissection block
// boundaries `part` and a right-hand side -- factor once, solve many, and carry the same problem between the C++ and Python
// front-ends.  The reference has no such format (its MATLAB development main reads .mat files,
// /root/reference/matlab/rchol_lap/rchol_lap.cpp:173-176) and its C++ API does not even return `part`
// (rchol_parallel.cpp:62-70); the Python twin of this file is rchol_b200/problems.py (save_problem / load_problem).
//
// Layout (little endian, no padding): 8-byte magic "RCHOLB2\0", then u64 {version = 1, N, nnzA, nnzG, nP, npart, nb},
// then  A.rowPtr[N+1] A.colIdx[nnzA] (u64)  A.val[nnzA] (f64)  G.rowPtr[N+1] G.colIdx[nnzG] (u64)  G.val[nnzG] (f64)
//       P[nP] (u64, nP = 0 or N)  part[npart] (u64)  b[nb] (f64, nb = 0 or N).
#ifndef RCHOL_B200_IO_HPP
#define RCHOL_B200_IO_HPP

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "sparse.hpp"

namespace rchol_b200 {

struct Problem {
  SparseCSR A, G;
  std::vector<size_t> P, part;
  std::vector<double> b;
};

namespace detail {
inline void put(FILE *f, const void *p, size_t bytes, const std::string &path) {
  if (bytes && fwrite(p, 1, bytes, f) != bytes) { fclose(f); throw std::runtime_error("short write: " + path); }
}
inline void get(FILE *f, void *p, size_t bytes, const std::string &path) {
  if (bytes && fread(p, 1, bytes, f) != bytes) { fclose(f); throw std::runtime_error("truncated file: " + path); }
}
}  // namespace detail

inline void save_problem(const std::string &path, const SparseCSR &A, const SparseCSR &G, const std::vector<size_t> &P,
                         const std::vector<size_t> &part, const std::vector<double> &b) {
  static_assert(sizeof(size_t) == 8, "64-bit indices");
  if (A.N != G.N || (!P.empty() && P.size() != A.N) || (!b.empty() && b.size() != A.N))
    throw std::invalid_argument("save_problem: inconsistent sizes");
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) throw std::runtime_error("cannot open for writing: " + path);
  const char magic[8] = {'R', 'C', 'H', 'O', 'L', 'B', '2', 0};
  const uint64_t hdr[7] = {1, A.N, A.nnz(), G.nnz(), P.size(), part.size(), b.size()};
  detail::put(f, magic, 8, path);
  detail::put(f, hdr, sizeof(hdr), path);
  detail::put(f, A.rowPtr, 8 * (A.N + 1), path); detail::put(f, A.colIdx, 8 * A.nnz(), path); detail::put(f, A.val, 8 * A.nnz(), path);
  detail::put(f, G.rowPtr, 8 * (G.N + 1), path); detail::put(f, G.colIdx, 8 * G.nnz(), path); detail::put(f, G.val, 8 * G.nnz(), path);
  detail::put(f, P.data(), 8 * P.size(), path);
  detail::put(f, part.data(), 8 * part.size(), path);
  detail::put(f, b.data(), 8 * b.size(), path);
  if (fclose(f) != 0) throw std::runtime_error("close failed: " + path);
}

inline void load_problem(const std::string &path, Problem &out) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) throw std::runtime_error("cannot open: " + path);
  char magic[8];
  uint64_t hdr[7];
  detail::get(f, magic, 8, path);
  detail::get(f, hdr, sizeof(hdr), path);
  if (memcmp(magic, "RCHOLB2", 8) != 0 || hdr[0] != 1) { fclose(f); throw std::runtime_error("not an rchol_b200 problem file: " + path); }
  const uint64_t N = hdr[1], nnzA = hdr[2], nnzG = hdr[3], nP = hdr[4], npart = hdr[5], nb = hdr[6];
  if ((nP != 0 && nP != N) || (nb != 0 && nb != N)) { fclose(f); throw std::runtime_error("corrupt header: " + path); }
  auto read_csr = [&](SparseCSR &M, uint64_t nnz) {
    std::vector<size_t> rp(N + 1), ci(nnz);
    std::vector<double> v(nnz);
    detail::get(f, rp.data(), 8 * (N + 1), path); detail::get(f, ci.data(), 8 * nnz, path); detail::get(f, v.data(), 8 * nnz, path);
    if (rp[0] != 0 || rp[N] != nnz) { fclose(f); throw std::runtime_error("corrupt row pointers: " + path); }
    M.init(rp, ci, v);
  };
  read_csr(out.A, nnzA);
  read_csr(out.G, nnzG);
  out.P.resize(nP); out.part.resize(npart); out.b.resize(nb);
  detail::get(f, out.P.data(), 8 * nP, path);
  detail::get(f, out.part.data(), 8 * npart, path);
  detail::get(f, out.b.data(), 8 * nb, path);
  fclose(f);
}

}  // namespace rchol_b200

#endif
