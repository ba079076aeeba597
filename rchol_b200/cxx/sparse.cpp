// Implementation of the boundary type declared in sparse.hpp (behaviour of /root/reference/c++/sparse.cpp:5-67).
#include "sparse.hpp"

#include <algorithm>
#include <iostream>
#include <stdexcept>

SparseCSR::SparseCSR() : N(0), rowPtr(nullptr), colIdx(nullptr), val(nullptr), ownMemory(false) {}

SparseCSR::SparseCSR(const std::vector<size_t> &rp, const std::vector<size_t> &ci, const std::vector<double> &v,
                     bool mem)
    : N(0), rowPtr(nullptr), colIdx(nullptr), val(nullptr), ownMemory(false) {
  init(rp, ci, v, mem);
}

void SparseCSR::init(const std::vector<size_t> &rp, const std::vector<size_t> &ci, const std::vector<double> &v,
                     bool mem) {
  if (N != 0) throw std::logic_error("SparseCSR::init on a non-empty matrix");
  if (rp.empty()) throw std::invalid_argument("SparseCSR::init: empty row pointer array");
  N = rp.size() - 1;
  const size_t nz = rp[N];
  rowPtr = new size_t[N + 1];
  colIdx = new size_t[nz];
  val = new double[nz];
  std::copy(rp.begin(), rp.end(), rowPtr);
  std::copy(ci.begin(), ci.begin() + nz, colIdx);
  std::copy(v.begin(), v.begin() + nz, val);
  ownMemory = mem;
}

SparseCSR::SparseCSR(const SparseCSR &other) : N(other.N), ownMemory(true) {
  const size_t nz = other.nnz();
  rowPtr = new size_t[N + 1];
  colIdx = new size_t[nz];
  val = new double[nz];
  std::copy(other.rowPtr, other.rowPtr + N + 1, rowPtr);
  std::copy(other.colIdx, other.colIdx + nz, colIdx);
  std::copy(other.val, other.val + nz, val);
}

size_t SparseCSR::size() const { return N; }

size_t SparseCSR::nnz() const { return rowPtr[N]; }

SparseCSR::~SparseCSR() {
  if (N > 0 && ownMemory) {
    delete[] rowPtr;
    delete[] colIdx;
    delete[] val;
  }
  rowPtr = nullptr;
  colIdx = nullptr;
  val = nullptr;
  N = 0;
}

void print(const SparseCSR &A, std::string name) {
  const char *labels[3] = {" rowPtr:", " colIdx:", " value:"};
  std::cout << name << labels[0] << "\n";
  for (size_t i = 0; i <= A.N; i++) std::cout << A.rowPtr[i] << " ";
  std::cout << "\n" << name << labels[1] << "\n";
  for (size_t k = 0; k < A.nnz(); k++) std::cout << A.colIdx[k] << " ";
  std::cout << "\n" << name << labels[2] << "\n";
  for (size_t k = 0; k < A.nnz(); k++) std::cout << A.val[k] << " ";
  std::cout << std::endl;
}
