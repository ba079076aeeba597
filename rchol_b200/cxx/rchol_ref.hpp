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

#endif
