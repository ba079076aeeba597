#include "util.hpp"

#include <algorithm>
#include <random>
#include <utility>

SparseCSR laplace_3d(int n) {
  const size_t m = (size_t)n, m2 = m * m, N = m * m * m;
  std::vector<size_t> rp(N + 1, 0), ci;
  std::vector<double> v;
  ci.reserve(7 * N);
  v.reserve(7 * N);
  for (size_t idx = 0; idx < N; idx++) {
    const size_t k = idx % m, j = (idx / m) % m, i = idx / m2;
    auto put = [&](size_t c, double w) { ci.push_back(c); v.push_back(w); };
    if (i > 0) put(idx - m2, -1.0);
    if (j > 0) put(idx - m, -1.0);
    if (k > 0) put(idx - 1, -1.0);
    put(idx, 6.0);
    if (k + 1 < m) put(idx + 1, -1.0);
    if (j + 1 < m) put(idx + m, -1.0);
    if (i + 1 < m) put(idx + m2, -1.0);
    rp[idx + 1] = ci.size();
  }
  return SparseCSR(rp, ci, v, false);   // like the reference: returned by value, arrays not owned
}

void reorder(const SparseCSR &A, const std::vector<size_t> &P, SparseCSR &B) {
  const size_t N = P.size();
  std::vector<size_t> inv(N);
  for (size_t i = 0; i < N; i++) inv[P[i]] = i;
  std::vector<size_t> rp(N + 1, 0), ci(A.nnz());
  std::vector<double> v(A.nnz());
  for (size_t i = 0; i < N; i++) rp[i + 1] = rp[i] + (A.rowPtr[P[i] + 1] - A.rowPtr[P[i]]);
#pragma omp parallel for schedule(static)
  for (long long ii = 0; ii < (long long)N; ii++) {
    const size_t i = (size_t)ii, src = A.rowPtr[P[i]], len = rp[i + 1] - rp[i];
    std::vector<std::pair<size_t, double>> row(len);
    for (size_t t = 0; t < len; t++) row[t] = {inv[A.colIdx[src + t]], A.val[src + t]};
    std::sort(row.begin(), row.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
    for (size_t t = 0; t < len; t++) { ci[rp[i] + t] = row[t].first; v[rp[i] + t] = row[t].second; }
  }
  B.init(rp, ci, v);
}

void reorder(const SparseCSR &A, std::vector<size_t> &rowPtr, std::vector<size_t> &colIdx, std::vector<double> &val,
             const std::vector<size_t> &P) {
  SparseCSR B;
  reorder(A, P, B);
  rowPtr.assign(B.rowPtr, B.rowPtr + B.N + 1);
  colIdx.assign(B.colIdx, B.colIdx + B.nnz());
  val.assign(B.val, B.val + B.nnz());
}

void rand(std::vector<double> &x, uint64_t seed) {
  std::mt19937_64 gen(seed);
  std::uniform_real_distribution<double> dist(0.0, 1.0);
  for (double &e : x) e = dist(gen);
}

// Row i of the upper half keeps the diagonal and the negative entries in columns [0, N) and sends every positive
// off-diagonal entry w at column c to column N + c with value -w; the lower half is the mirror image.
void sdd_to_sddm(const SparseCSR &A, SparseCSR &Ae) {
  const size_t N = A.N;
  std::vector<size_t> rp(2 * N + 1, 0), ci(2 * A.nnz());
  std::vector<double> v(2 * A.nnz());
  for (size_t i = 0; i < N; i++) rp[i + 1] = rp[i + 1 + N] = A.rowPtr[i + 1] - A.rowPtr[i];
  for (size_t i = 0; i < 2 * N; i++) rp[i + 1] += rp[i];
#pragma omp parallel for schedule(static)
  for (long long ii = 0; ii < (long long)(2 * N); ii++) {
    const size_t r = (size_t)ii, i = r % N, shift_same = r < N ? 0 : N, shift_other = r < N ? N : 0;
    const size_t first = A.rowPtr[i], len = A.rowPtr[i + 1] - first;
    std::vector<std::pair<size_t, double>> row(len);
    for (size_t t = 0; t < len; t++) {
      const size_t c = A.colIdx[first + t];
      const double w = A.val[first + t];
      const bool pos = c != i && w > 0;
      row[t] = {c + (pos ? shift_other : shift_same), pos ? -w : w};
    }
    std::sort(row.begin(), row.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
    for (size_t t = 0; t < len; t++) { ci[rp[r] + t] = row[t].first; v[rp[r] + t] = row[t].second; }
  }
  Ae.init(rp, ci, v);
}

void sdd_rhs(const std::vector<double> &b, std::vector<double> &be) {
  be.resize(2 * b.size());
  for (size_t i = 0; i < b.size(); i++) { be[i] = b[i]; be[i + b.size()] = -b[i]; }
}

void sdd_recover(const std::vector<double> &xe, std::vector<double> &x) {
  const size_t N = xe.size() / 2;
  x.resize(N);
  for (size_t i = 0; i < N; i++) x[i] = 0.5 * (xe[i] - xe[i + N]);
}
