// SparseCSR -- the boundary type of the drop-in.  Same class name, members, layout and semantics as the
// reference's container (/root/reference/c++/sparse.hpp:10-31, sparse.cpp:5-67): zero-based CSR with 64-bit
// unsigned indices, public raw arrays, and an `ownMemory` flag that decides whether the destructor frees them.
// Written from the interface, not copied: code that includes this header links against either implementation.
#ifndef sparse_hpp
#define sparse_hpp

#include <cstddef>
#include <string>
#include <vector>

class SparseCSR {
public:
  SparseCSR();
  SparseCSR(const std::vector<size_t> &rowPtr, const std::vector<size_t> &colIdx, const std::vector<double> &val,
            bool mem = true);
  SparseCSR(const SparseCSR &);   // deep copy, owns its arrays

  // fills an empty matrix from three vectors (copies them)
  void init(const std::vector<size_t> &rowPtr, const std::vector<size_t> &colIdx, const std::vector<double> &val,
            bool mem = true);

  size_t size() const;   // number of rows
  size_t nnz() const;    // rowPtr[N]

  ~SparseCSR();

public:
  size_t N = 0;
  size_t *rowPtr;
  size_t *colIdx;
  double *val;
  bool ownMemory;
};

void print(const SparseCSR &, std::string name);

#endif
