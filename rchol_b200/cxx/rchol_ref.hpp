// Declarations of the reference's factorization API (the fixed-input producer, compiled UNMODIFIED from
// /root/reference/c++ into baseline/_ref/librchol_producer.so).  Signatures: c++/rchol/rchol.hpp:6 and
// c++/rchol/rchol_parallel.hpp:6.  `refprod_last_part` is the producer shim's accessor for the partition
// boundaries that the reference computes but does not return (baseline/producer_shim.cpp).
#ifndef RCHOL_REF_HPP
#define RCHOL_REF_HPP

#include <cstdint>
#include <vector>

#include "sparse.hpp"

void rchol(const SparseCSR &A, SparseCSR &G);
void rchol(const SparseCSR &A, SparseCSR &G, std::vector<size_t> &permutation, int threads);
extern "C" uint64_t refprod_last_part(uint64_t *out, uint64_t capacity);   // returns the number of boundaries

// Additive overload the reference lacks (its Python and MATLAB bindings return `part` as a third output,
// python/rchol/rchol.py:45, matlab/rchol/rchol.m:1,34-36): the block boundaries of rchol_parallel.cpp:64-70 with the
// ground vertex dropped, i.e. what pcg(..., part, ...) and rcg_set_factor take.
inline void rchol(const SparseCSR &A, SparseCSR &G, std::vector<size_t> &permutation, std::vector<size_t> &part, int threads) {
  rchol(A, G, permutation, threads);
  std::vector<uint64_t> buf(2 * (size_t)threads + 2);
  const uint64_t np = refprod_last_part(buf.data(), buf.size());
  part.assign(buf.begin(), buf.begin() + np);
}

#endif
