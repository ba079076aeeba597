// Host helpers either side of the hot path, with the reference's names and argument meaning
// (/root/reference/c++/util/util.hpp:32-55,147-164 and util.cpp:6-57, laplace_3d.hpp:8-65).
#ifndef UTIL_HPP
#define UTIL_HPP

#include <cstdint>
#include <iostream>
#include <string>
#include <vector>

#include "sparse.hpp"

SparseCSR laplace_3d(int n);                                                   // 7-point Dirichlet Laplacian, n^3 rows
void reorder(const SparseCSR &A, const std::vector<size_t> &P, SparseCSR &B);  // B = A(P,P), rows re-sorted
void rand(std::vector<double> &x, uint64_t seed = 2024);                       // U(0,1); the reference is unseeded
// same permutation, outputs as three arrays (util.hpp:36-37; what rchol_parallel.cpp:71 feeds the factorization)
void reorder(const SparseCSR &A, std::vector<size_t> &rowPtr, std::vector<size_t> &colIdx, std::vector<double> &val,
             const std::vector<size_t> &P);

// SDD front-end (only the reference's MATLAB binding has it: matlab/rchol/sdd_to_sddm.m:2-17, ex_sdd.m:12-30):
// Ae = [D + Neg, -Pos; -Pos, D + Neg] (2N x 2N SDDM, rows sorted), be = [b; -b], x = (xe[0:N] - xe[N:2N]) / 2
void sdd_to_sddm(const SparseCSR &A, SparseCSR &Ae);
void sdd_rhs(const std::vector<double> &b, std::vector<double> &be);
void sdd_recover(const std::vector<double> &xe, std::vector<double> &x);

template <typename T>
void reorder(std::vector<T> &x, std::vector<size_t> &p, std::vector<T> &xp) {   // xp[i] = x[p[i]]
  xp.clear();
  xp.reserve(p.size());
  for (size_t i = 0; i < p.size(); i++) xp.push_back(x[p[i]]);
}

template <typename T>
std::vector<T> reorder(std::vector<T> &x, std::vector<size_t> &p) {   // value-returning form (util.hpp:158-164)
  std::vector<T> xp;
  reorder(x, p, xp);
  return xp;
}

template <typename T>
void print(const T &x, std::string name) {   // "name:" then the elements on one line (util.hpp:40-46)
  std::cout << name << ":" << std::endl;
  for (size_t i = 0; i < x.size(); i++) std::cout << x[i] << " ";
  std::cout << std::endl;
}

template <typename T>
void unpermute(const std::vector<T> &xp, const std::vector<size_t> &p, std::vector<T> &x) {   // x[p[i]] = xp[i]
  x.resize(p.size());
  for (size_t i = 0; i < p.size(); i++) x[p[i]] = xp[i];
}

#endif
