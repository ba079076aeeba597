// rchol_b200 -- one-off device set-up ("analysis") of the solve phase.  Replaces pcg::create_sparse
// (/root/reference/c++/util/pcg.cpp:31-54, which deep-copies A and G into MKL handles) with:
//   * upload of the caller's SparseCSR arrays and narrowing of the 64-bit column indices to 32 bits,
//   * validation of G (upper triangular, diagonal first and positive: rchol_lap.cpp:358-387 `coalesce`),
//   * L = U^T by rows (count / scan / scatter / per-row sort) for the forward solve,
//   * J U J (index reversal) for the backward solve, so that ONE lower-triangular kernel serves both,
//   * the nested-dissection block schedule from `part` (rchol_parallel.cpp:64-70, rchol_lap.cpp:254-261).
#include <omp.h>

#include <algorithm>
#include <cstdlib>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "rcg_common.cuh"

namespace {

double wall_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------------------------------------------------
// exclusive scan of int64 (three small kernels; set-up only)
// ---------------------------------------------------------------------------------------------------------
constexpr int SCAN_T = 1024;
constexpr int SCAN_I = 4;
constexpr int SCAN_TILE = SCAN_T * SCAN_I;

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t *total, int64_t *smem /*>=33*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int64_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int64_t w = lane < nw ? smem[lane] : 0;
    int64_t winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int64_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    smem[lane] = winc - w;          // exclusive warp offsets
    if (lane == 31) smem[32] = winc;  // block total
  }
  __syncthreads();
  int64_t res = smem[warp] + inc - v;
  *total = smem[32];
  __syncthreads();
  return res;
}

__global__ void k_scan_tile_sums(const int64_t *__restrict__ in, int64_t n, int64_t *__restrict__ tile_sums) {
  __shared__ int64_t sm[34];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_I;
  int64_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_I; i++)
    if (base + i < n) s += in[base + i];
  int64_t total;
  block_exclusive_scan(s, &total, sm);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void k_scan_tile_offsets(int64_t *tile_sums, int64_t ntiles) {
  __shared__ int64_t sm[34];
  int64_t carry = 0;
  for (int64_t base = 0; base < ntiles; base += blockDim.x) {
    int64_t i = base + threadIdx.x;
    int64_t v = i < ntiles ? tile_sums[i] : 0;
    int64_t total;
    int64_t ex = block_exclusive_scan(v, &total, sm);
    if (i < ntiles) tile_sums[i] = carry + ex;
    carry += total;
  }
}

__global__ void k_scan_apply(int64_t *data, int64_t n, const int64_t *__restrict__ tile_offsets) {
  __shared__ int64_t sm[34];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_I;
  int64_t v[SCAN_I];
  int64_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_I; i++) {
    v[i] = base + i < n ? data[base + i] : 0;
    s += v[i];
  }
  int64_t total;
  int64_t ex = block_exclusive_scan(s, &total, sm) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_I; i++) {
    if (base + i < n) data[base + i] = ex;
    ex += v[i];
  }
}

int exclusive_scan_inplace(rcg_handle *h, int64_t *data, int64_t n) {
  int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  int64_t *tile_sums = nullptr;
  RCG_CUDA(h, cudaMalloc(&tile_sums, sizeof(int64_t) * (size_t)ntiles));
  k_scan_tile_sums<<<(unsigned)ntiles, SCAN_T, 0, h->stream>>>(data, n, tile_sums);
  k_scan_tile_offsets<<<1, SCAN_T, 0, h->stream>>>(tile_sums, ntiles);
  k_scan_apply<<<(unsigned)ntiles, SCAN_T, 0, h->stream>>>(data, n, tile_sums);
  h->stats.kernel_launches += 3;
  RCG_CUDA(h, cudaGetLastError());
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  RCG_CUDA(h, cudaFree(tile_sums));
  return RCG_OK;
}

// ---------------------------------------------------------------------------------------------------------
// upload helpers
// ---------------------------------------------------------------------------------------------------------
// Host -> device copies of the caller's (pageable) SparseCSR arrays.  A plain cudaMemcpy of pageable memory is staged
// by the driver on ONE thread (measured 11 GB/s for the 10 GB of 256^3 / T=8).  Here the host cores fill two pinned
// staging buffers (narrowing the 64-bit column indices to 32 bits and range-checking them on the way, so only 12 B
// per entry cross PCIe instead of 16) while the copy engine drains the other buffer.
constexpr size_t STAGE_BYTES = (size_t)64 << 20;
}  // namespace

// Pinning 2 x 64 MiB costs tens of milliseconds per handle, and the reference's one-shot `pcg(...)` creates a handle per
// solve: released staging buffers are kept (at most four per process) and handed to the next handle.
static std::mutex g_stage_mu;
static std::vector<void *> g_stage_free;

void rcg_release_stage_buffer(void *p) {   // rcg_destroy
  if (!p) return;
  std::lock_guard<std::mutex> lock(g_stage_mu);
  if (g_stage_free.size() < 4) g_stage_free.push_back(p);
  else cudaFreeHost(p);
}

namespace {

int stage_init(rcg_handle *h) {
  if (h->stage_buf[0]) return RCG_OK;
  for (int i = 0; i < 2; i++) {
    {
      std::lock_guard<std::mutex> lock(g_stage_mu);
      if (!g_stage_free.empty()) { h->stage_buf[i] = g_stage_free.back(); g_stage_free.pop_back(); }
    }
    if (!h->stage_buf[i]) RCG_CUDA(h, cudaHostAlloc(&h->stage_buf[i], STAGE_BYTES, cudaHostAllocPortable));
    RCG_CUDA(h, cudaEventCreateWithFlags(&h->stage_ev[i], cudaEventDisableTiming));
  }
  return RCG_OK;
}

// Host threads that fill the pinned staging buffers: all cores (at most 16), divided by the ranks that share the node
// (torchrun exports LOCAL_WORLD_SIZE; measured at N = 8 on a 16-core box: 8 x 16 staging threads made the 1.2 GB upload of
// a rank slower than the 7.7 GB upload of a single rank).  RCHOL_B200_HOST_THREADS overrides.
int host_threads() {
  int t = omp_get_num_procs();
  if (const char *e = getenv("RCHOL_B200_HOST_THREADS")) {
    const int v = atoi(e);
    if (v > 0) return v > 64 ? 64 : v;
  }
  if (const char *e = getenv("LOCAL_WORLD_SIZE")) {
    const int lw = atoi(e);
    if (lw > 1) t = t / lw;
  }
  return t < 1 ? 1 : (t > 16 ? 16 : t);
}

// Lanes per row of the CSR SpMV from the row-length histogram (north star: "warp-per-row or merge-based, chosen per
// row-length histogram").  The histogram counts ENTRIES by the length of the row they sit in, with the bucket edges at the
// lengths up to which 2 / 4 / 8 / 16 lanes per row keep their lanes busy (2, 5, 12, 24); the choice is the smallest lane
// count whose bucket edge covers at least 90 % of the entries -- a matrix with a few very long rows among many short ones is
// served by its bulk, a matrix whose entries sit mostly in long rows gets a warp per row.  Integer sums: the result does not
// depend on the thread count or on the order of the rows (a permuted matrix gets the same plan).
int spmv_lanes_from_rows(uint64_t N, const uint64_t *rowPtr, uint64_t entries[5]) {
  unsigned long long e0 = 0, e1 = 0, e2 = 0, e3 = 0, e4 = 0;
  const int T = host_threads();
#pragma omp parallel for num_threads(T) schedule(static) reduction(+ : e0, e1, e2, e3, e4)
  for (long long i = 0; i < (long long)N; i++) {
    const uint64_t len = rowPtr[i + 1] - rowPtr[i];
    if (len <= 2) e0 += len;
    else if (len <= 5) e1 += len;
    else if (len <= 12) e2 += len;
    else if (len <= 24) e3 += len;
    else e4 += len;
  }
  const unsigned long long e[5] = {e0, e1, e2, e3, e4};
  unsigned long long total = 0, cum = 0;
  for (int k = 0; k < 5; k++) { total += e[k]; if (entries) entries[k] = e[k]; }
  static const int lanes[5] = {2, 4, 8, 16, 32};
  for (int k = 0; k < 5; k++) {
    cum += e[k];
    if ((long double)cum * 10.0L >= (long double)total * 9.0L) return lanes[k];
  }
  return 32;
}

// fill(dst, first, count) writes elements [first, first + count) of the device array into the pinned buffer `dst`
template <class Fill>
int staged_h2d(rcg_handle *h, void *dst, size_t n, size_t elem_bytes, Fill fill) {
  RCG_TRY(stage_init(h));
  const size_t per = STAGE_BYTES / elem_bytes;
  int slot = 0;
  for (size_t first = 0; first < n; first += per, slot ^= 1) {
    const size_t cnt = std::min(per, n - first);
    RCG_CUDA(h, cudaEventSynchronize(h->stage_ev[slot]));   // the copy that last read this buffer has finished
    fill(h->stage_buf[slot], first, cnt);
    RCG_CUDA(h, cudaMemcpyAsync(static_cast<char *>(dst) + first * elem_bytes, h->stage_buf[slot], cnt * elem_bytes,
                                cudaMemcpyHostToDevice, h->stream));
    RCG_CUDA(h, cudaEventRecord(h->stage_ev[slot], h->stream));
  }
  return RCG_OK;
}

int staged_copy(rcg_handle *h, void *dst, const void *src, size_t bytes) {
  const int T = host_threads();
  return staged_h2d(h, dst, bytes, 1, [&](void *stage, size_t first, size_t cnt) {
    const size_t blk = (size_t)1 << 20;
    const long nblk = (long)((cnt + blk - 1) / blk);
#pragma omp parallel for num_threads(T) schedule(static)
    for (long b = 0; b < nblk; b++) {
      const size_t o = (size_t)b * blk;
      memcpy(static_cast<char *>(stage) + o, static_cast<const char *>(src) + first + o, std::min(blk, cnt - o));
    }
  });
}

// 64-bit -> 32-bit column indices; *bad is set when an index is >= N
int staged_narrow(rcg_handle *h, uint32_t *dst, const uint64_t *src, size_t n, uint64_t N, bool *bad) {
  const int T = host_threads();
  int any_bad = 0;
  int rc = staged_h2d(h, dst, n, sizeof(uint32_t), [&](void *stage, size_t first, size_t cnt) {
    uint32_t *out = static_cast<uint32_t *>(stage);
    const uint64_t *in = src + first;
    const long blk = 1 << 18;
    const long nblk = ((long)cnt + blk - 1) / blk;
    int b_any = 0;
#pragma omp parallel for num_threads(T) schedule(static) reduction(| : b_any)
    for (long b = 0; b < nblk; b++) {
      const long lo = b * blk, hi = std::min<long>((long)cnt, lo + blk);
      uint64_t mx = 0;
      for (long i = lo; i < hi; i++) {
        const uint64_t c = in[i];
        mx = c > mx ? c : mx;
        out[i] = (uint32_t)c;
      }
      b_any |= (mx >= N) ? 1 : 0;
    }
    any_bad |= b_any;
  });
  *bad = any_bad != 0;
  return rc;
}

// rowPtr must start at 0, be non-decreasing and end at nnz
__global__ void k_check_rowptr(const int64_t *__restrict__ rp, int64_t N, int64_t nnz, int *err) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (; i < N; i += stride) bad |= (rp[i + 1] < rp[i]);
  if (blockIdx.x == 0 && threadIdx.x == 0) bad |= (rp[0] != 0) || (rp[N] != nnz);
  if (bad) atomicExch(err, 1);
}

int grid_for(const rcg_handle *h, int64_t work, int threads, int per_sm = 16) {
  int64_t g = (work + threads - 1) / threads;
  int64_t cap = (int64_t)h->sm_count * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// Uploads a SparseCSR (64-bit indices) and produces {int64 rowptr, u32 col, fp64 val} on the device.
int upload_csr(rcg_handle *h, uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val,
               CsrDev &out) {
  if (!rowPtr || (N && (!colIdx || !val))) {
    h->err = "null matrix pointer";
    return RCG_ERR_INVALID;
  }
  if (N == 0 || N >= 0xFFFFFFFFull) {
    h->err = "matrix dimension must be in [1, 2^32-2]";
    return RCG_ERR_INVALID;
  }
  const int64_t nnz = (int64_t)rowPtr[N];
  if (nnz <= 0) {
    h->err = "empty matrix";
    return RCG_ERR_INVALID;
  }
  int *derr = nullptr;
  // the copies and the row-pointer check; on any failure the caller's `out` is left empty (nothing stays allocated)
  auto body = [&]() -> int {
    double t0 = wall_ms();
    out.nnz = nnz;
    RCG_CUDA(h, cudaMalloc(&out.rowptr, sizeof(int64_t) * (N + 1)));
    RCG_CUDA(h, cudaMalloc(&out.col, sizeof(uint32_t) * (size_t)nnz));
    RCG_CUDA(h, cudaMalloc(&out.val, sizeof(double) * (size_t)nnz));
    RCG_CUDA(h, cudaMalloc(&derr, sizeof(int)));
    RCG_CUDA(h, cudaMemsetAsync(derr, 0, sizeof(int), h->stream));
    bool bad_col = false;
    RCG_TRY(staged_copy(h, out.rowptr, rowPtr, sizeof(int64_t) * (N + 1)));
    RCG_TRY(staged_narrow(h, out.col, colIdx, (size_t)nnz, N, &bad_col));
    RCG_TRY(staged_copy(h, out.val, val, sizeof(double) * (size_t)nnz));
    RCG_CUDA(h, cudaStreamSynchronize(h->stream));
    h->stats.upload_ms += wall_ms() - t0;
    h->stats.h2d_bytes += sizeof(int64_t) * (N + 1) + (size_t)nnz * 12;

    double t1 = wall_ms();
    k_check_rowptr<<<grid_for(h, (int64_t)N, 256), 256, 0, h->stream>>>(out.rowptr, (int64_t)N, nnz, derr);
    h->stats.kernel_launches += 1;
    int herr = 0;
    RCG_CUDA(h, cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    RCG_CUDA(h, cudaStreamSynchronize(h->stream));
    h->stats.analysis_ms += wall_ms() - t1;
    if (bad_col) herr = 1;
    if (herr) {
      h->err = "malformed CSR: column index out of range or row pointers not monotone";
      return RCG_ERR_INVALID;
    }
    return RCG_OK;
  };
  const int rc = body();
  cudaFree(derr);
  if (rc != RCG_OK) rcg_free_csr(out);
  return rc;
}

// ---------------------------------------------------------------------------------------------------------
// validation of G = CSR(U)
// ---------------------------------------------------------------------------------------------------------
__global__ void k_validate_upper(const int64_t *__restrict__ rp, const uint32_t *__restrict__ col,
                                 const double *__restrict__ val, uint32_t N, int *err) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t stride = gridDim.x * blockDim.x;
  bool bad = false;
  for (; i < N; i += stride) {
    int64_t s = rp[i], e = rp[i + 1];
    if (e <= s) { bad = true; continue; }
    bad |= (col[s] != i) || !(val[s] > 0.0);
    uint32_t prev = i;
    for (int64_t k = s + 1; k < e; k++) {
      uint32_t c = col[k];
      bad |= (c <= prev);
      prev = c;
    }
  }
  if (bad) atomicExch(err, 1);
}

// every entry (i, c) of U must have block(c) == block(i) or an ancestor of block(i):
// ancestors of a block a are exactly the blocks whose subtree range [sub_lo[a], hi[a]) contains the row.
__global__ void k_validate_blocks(const int64_t *__restrict__ rp, const uint32_t *__restrict__ col, uint32_t N,
                                  const uint32_t *__restrict__ part, const uint32_t *__restrict__ sub_lo, int nb,
                                  int *err) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t stride = gridDim.x * blockDim.x;
  bool bad = false;
  for (; i < N; i += stride) {
    for (int64_t k = rp[i] + 1; k < rp[i + 1]; k++) {
      uint32_t c = col[k];
      int lo = 0, hi = nb;   // find a with part[a] <= c < part[a+1]
      while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (part[mid] <= c) lo = mid; else hi = mid;
      }
      bad |= (sub_lo[lo] > i);
    }
  }
  if (bad) atomicExch(err, 1);
}

// general block lists: an entry (i, c) of U with block(c) != block(i) needs depth(block(c)) < depth(block(i))
__global__ void k_validate_depths(const int64_t *__restrict__ rp, const uint32_t *__restrict__ col, uint32_t N,
                                  const uint32_t *__restrict__ part, const uint32_t *__restrict__ depth, int nb, int *err) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t stride = gridDim.x * blockDim.x;
  bool bad = false;
  for (; i < N; i += stride) {
    int lo = 0, hi = nb;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (part[mid] <= i) lo = mid; else hi = mid; }
    const int bi = lo;
    for (int64_t k = rp[i] + 1; k < rp[i + 1]; k++) {
      const uint32_t c = col[k];
      lo = 0; hi = nb;
      while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (part[mid] <= c) lo = mid; else hi = mid; }
      bad |= (lo != bi) && !(depth[lo] < depth[bi]);
    }
  }
  if (bad) atomicExch(err, 1);
}

// ---------------------------------------------------------------------------------------------------------
// transpose  U (CSR) -> L = U^T (CSR, sorted rows)
// ---------------------------------------------------------------------------------------------------------
__global__ void k_count_cols(const uint32_t *__restrict__ col, int64_t nnz, int64_t *__restrict__ cnt) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) atomicAdd((unsigned long long *)&cnt[col[i]], 1ull);
}

// 8 lanes per U row; scatters (row, val) to the L row of every column.  Order inside an L row is arbitrary here.
__global__ void k_scatter_transpose(const int64_t *__restrict__ rp, const uint32_t *__restrict__ col,
                                    const double *__restrict__ val, uint32_t N, int64_t *__restrict__ cursor,
                                    uint32_t *__restrict__ lcol, double *__restrict__ lval) {
  const int sub = threadIdx.x & 7;
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 3;
  for (; row < N; row += stride) {
    const int64_t s = rp[row], e = rp[row + 1];
    for (int64_t k = s + sub; k < e; k += 8) {
      uint32_t c = col[k];
      int64_t pos = (int64_t)atomicAdd((unsigned long long *)&cursor[c], 1ull);
      lcol[pos] = (uint32_t)row;
      lval[pos] = val[k];
    }
  }
}

constexpr int SORT_SMALL = 64;

// thread-per-row insertion sort (rows arrive nearly sorted because U rows are scattered in ascending order);
// longer rows are appended to `long_rows` for the block-wide bitonic sort; n_long[0] = their number, n_long[1] = the
// length of the longest of them (capped at 2^32 - 1), so that the host needs 8 bytes back and no row pointers.
__global__ void k_sort_rows_small(const int64_t *__restrict__ rp, uint32_t *__restrict__ col, double *val,
                                  uint32_t N, uint32_t *__restrict__ long_rows, unsigned int *n_long) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t stride = gridDim.x * blockDim.x;
  for (; j < N; j += stride) {
    const int64_t s = rp[j];
    const int64_t len64 = rp[j + 1] - s;
    const int len = (int)len64;
    if (len64 > SORT_SMALL) {
      unsigned int slot = atomicAdd(n_long, 1u);
      long_rows[slot] = j;
      const unsigned int l32 = (unsigned int)(len64 > 0xFFFFFFFFll ? 0xFFFFFFFFll : len64);
      if (l32 > *(volatile unsigned int *)(n_long + 1)) atomicMax(n_long + 1, l32);   // (the maximum only grows: most rows skip the atomic)
      continue;
    }
    for (int a = 1; a < len; a++) {
      uint32_t kc = col[s + a];
      double kv = val ? val[s + a] : 0.0;
      int b = a - 1;
      while (b >= 0 && col[s + b] > kc) {
        col[s + b + 1] = col[s + b];
        if (val) val[s + b + 1] = val[s + b];
        b--;
      }
      col[s + b + 1] = kc;
      if (val) val[s + b + 1] = kv;
    }
  }
}

// CTA-per-row bitonic sort of (col, val) pairs.  Rows up to `smem_cap` entries are sorted in shared memory,
// longer ones in place in global memory through a padded scratch area.
__global__ void k_sort_rows_bitonic(const int64_t *__restrict__ rp, uint32_t *__restrict__ col, double *val,
                                    const uint32_t *__restrict__ long_rows, int smem_cap, uint32_t *__restrict__ gkeys,
                                    double *__restrict__ gvals, int64_t gstride) {
  extern __shared__ unsigned char sm_raw[];
  const uint32_t j = long_rows[blockIdx.x];
  const int64_t s = rp[j];
  const int len = (int)(rp[j + 1] - s);
  int n2 = 1;
  while (n2 < len) n2 <<= 1;
  uint32_t *keys;
  double *vals;
  if (n2 <= smem_cap) {
    vals = reinterpret_cast<double *>(sm_raw);
    keys = reinterpret_cast<uint32_t *>(sm_raw + sizeof(double) * (size_t)smem_cap);
  } else {
    keys = gkeys + (int64_t)blockIdx.x * gstride;
    vals = gvals + (int64_t)blockIdx.x * gstride;
  }
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    keys[i] = i < len ? col[s + i] : 0xFFFFFFFFu;
    vals[i] = (i < len && val) ? val[s + i] : 0.0;
  }
  __syncthreads();
  for (int k = 2; k <= n2; k <<= 1) {
    for (int jj = k >> 1; jj > 0; jj >>= 1) {
      for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        int ixj = i ^ jj;
        if (ixj > i) {
          bool up = ((i & k) == 0);
          uint32_t a = keys[i], b = keys[ixj];
          if ((a > b) == up) {
            keys[i] = b; keys[ixj] = a;
            double t = vals[i]; vals[i] = vals[ixj]; vals[ixj] = t;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    col[s + i] = keys[i];
    if (val) val[s + i] = vals[i];
  }
}

// Sorts every segment [rp[i], rp[i+1]) of (key[, val]) ascending by key: thread per short segment, CTA-wide bitonic
// sort for long ones.  Used for the rows of L = U^T, for the level segments and for the level-space rows.
int sort_segments(rcg_handle *h, const int64_t *rp, uint32_t *key, double *val, uint32_t nseg) {
  uint32_t *long_rows = nullptr;
  unsigned int *n_long = nullptr;   // {number of long segments, length of the longest}
  RCG_CUDA(h, cudaMalloc(&long_rows, sizeof(uint32_t) * std::max<uint32_t>(nseg, 1)));
  RCG_CUDA(h, cudaMalloc(&n_long, 2 * sizeof(unsigned int)));
  auto body = [&]() -> int {
    RCG_CUDA(h, cudaMemsetAsync(n_long, 0, 2 * sizeof(unsigned int), h->stream));
    k_sort_rows_small<<<grid_for(h, nseg, 128), 128, 0, h->stream>>>(rp, key, val, nseg, long_rows, n_long);
    h->stats.kernel_launches += 1;
    unsigned int hl[2] = {0, 0};
    RCG_CUDA(h, cudaMemcpyAsync(hl, n_long, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
    RCG_CUDA(h, cudaStreamSynchronize(h->stream));
    const unsigned int hn_long = hl[0];
    const int64_t maxlen = (int64_t)hl[1];
    if (hn_long == 0) return RCG_OK;
    // (The order of `long_rows` varies from run to run; every segment is sorted on its own, so the result does not.)
    if (maxlen >= ((int64_t)1 << 30)) { h->err = "a row with 2^30 or more entries cannot be sorted on the device"; return RCG_ERR_INVALID; }
    const int smem_cap = 8192;                      // entries: 8192 * 12 B = 96 KB of shared memory
    const size_t smem_bytes = (size_t)smem_cap * 12;
    RCG_CUDA(h, cudaFuncSetAttribute(k_sort_rows_bitonic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    int64_t gstride = 0;
    uint32_t *gkeys = nullptr;
    double *gvals = nullptr;
    unsigned int batch = hn_long;
    if (maxlen > smem_cap) {
      gstride = 1;
      while (gstride < maxlen) gstride <<= 1;
      // the long segments are processed in batches so that the padded scratch area (gstride entries per CTA of a
      // launch, 12 B each) stays at or below 1 GiB
      const int64_t fit = ((int64_t)1 << 30) / (12 * gstride);
      batch = (unsigned int)std::max<int64_t>(1, std::min<int64_t>(fit, (int64_t)hn_long));
      cudaError_t ae = cudaMalloc(&gkeys, sizeof(uint32_t) * (size_t)gstride * batch);
      if (ae == cudaSuccess) ae = cudaMalloc(&gvals, sizeof(double) * (size_t)gstride * batch);
      if (ae != cudaSuccess) { cudaFree(gkeys); cudaFree(gvals); }
      RCG_CUDA(h, ae);
    }
    for (unsigned int first = 0; first < hn_long; first += batch) {
      const unsigned int cnt = std::min(batch, hn_long - first);
      k_sort_rows_bitonic<<<cnt, 512, smem_bytes, h->stream>>>(rp, key, val, long_rows + first, smem_cap, gkeys, gvals, gstride);
      h->stats.kernel_launches += 1;
    }
    cudaError_t se = cudaGetLastError();
    if (se == cudaSuccess) se = cudaStreamSynchronize(h->stream);
    cudaFree(gkeys); cudaFree(gvals);
    RCG_CUDA(h, se);
    return RCG_OK;
  };
  const int rc = body();
  cudaFree(long_rows);
  cudaFree(n_long);
  return rc;
}

// ---------------------------------------------------------------------------------------------------------
// J U J : reversed copy of U (row N-1-i, column N-1-c), which is lower triangular with the diagonal last
// ---------------------------------------------------------------------------------------------------------
__global__ void k_reverse_entries(const uint32_t *__restrict__ col, const double *__restrict__ val, int64_t nnz,
                                  uint32_t N, uint32_t *__restrict__ rcol, double *__restrict__ rval) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < nnz; i += stride) {
    rcol[nnz - 1 - i] = N - 1 - col[i];
    rval[nnz - 1 - i] = val[i];
  }
}
__global__ void k_reverse_rowptr(const int64_t *__restrict__ rp, int64_t nnz, uint32_t N, int64_t *__restrict__ rrp) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i <= (int64_t)N; i += stride) rrp[i] = nnz - rp[N - i];
}

// ---------------------------------------------------------------------------------------------------------
// per-row finishing: number of external entries, reciprocal of the diagonal
// `bounds` = ascending block boundaries in the direction's solve index space (nb+1 entries)
// ---------------------------------------------------------------------------------------------------------
__global__ void k_count_external(const int64_t *__restrict__ rp, const uint32_t *__restrict__ col, double *__restrict__ val,
                                 uint32_t N, const uint32_t *__restrict__ bounds, int nb, int64_t *__restrict__ ext_cnt,
                                 int *err) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t stride = gridDim.x * blockDim.x;
  bool bad = false;
  for (; j < N; j += stride) {
    int lo = 0, hi = nb;
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (bounds[mid] <= j) lo = mid; else hi = mid;
    }
    const uint32_t blo = bounds[lo];
    const int64_t s = rp[j], e = rp[j + 1];
    int64_t k = s;
    while (k < e - 1 && col[k] < blo) k++;
    ext_cnt[j] = k - s;
    bad |= (col[e - 1] != j);
    val[e - 1] = 1.0 / val[e - 1];
  }
  if (bad) atomicExch(err, 1);
}

// loc_rp = rp - ext_rp (element-wise), in place over rp
__global__ void k_sub_rowptr(int64_t *__restrict__ rp, const int64_t *__restrict__ ext_rp, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) rp[i] -= ext_rp[i];
}

// 8 lanes per row: copy the external prefix and the local suffix of every row into the two split matrices.
// `rp` has already been turned into the local row pointers; the combined start of row j is rp[j] + ext_rp[j].
__global__ void k_split_rows(const int64_t *__restrict__ loc_rp, const int64_t *__restrict__ ext_rp,
                             const uint32_t *__restrict__ col, const double *__restrict__ val, uint32_t N,
                             const uint32_t *__restrict__ bounds, int nb,
                             uint32_t *__restrict__ lcol, double *__restrict__ lval, uint32_t *__restrict__ ecol,
                             double *__restrict__ eval) {
  const int sub = threadIdx.x & 7;
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 3;
  for (; row < N; row += stride) {
    const int64_t ls = loc_rp[row], le = loc_rp[row + 1], es = ext_rp[row], ee = ext_rp[row + 1];
    const int64_t src = ls + es;
    const int64_t ne = ee - es, nl = le - ls;
    for (int64_t i = sub; i < ne; i += 8) { ecol[es + i] = col[src + i]; eval[es + i] = val[src + i]; }
    // local part: off-diagonal values are stored as -v/diag (the diagonal slot, last, already holds 1/diag), so
    // x_j = init_j/diag + sum v'_jc x_c and the chain kernel's last dependent operation per row is a single FMA
    // Local columns are stored relative to the first row of the block (the chain kernel indexes its solution
    // window with them directly).
    const double dinv = val[src + ne + nl - 1];
    int blo_i = 0, bhi_i = nb;
    while (bhi_i - blo_i > 1) {
      const int mid = (blo_i + bhi_i) >> 1;
      if (bounds[mid] <= (uint32_t)row) blo_i = mid; else bhi_i = mid;
    }
    const uint32_t blo = bounds[blo_i];
    for (int64_t i = sub; i < nl; i += 8) {
      double v = val[src + ne + i];
      if (v != v) v = __longlong_as_double(0x7FF8000000000000ll);   // canonical NaN: never the solve's sentinel
      lcol[ls + i] = col[src + ne + i] - blo;
      lval[ls + i] = (i == nl - 1) ? v : -(v * dinv);   // off-diagonals: negated and scaled by 1/diag
    }
  }
}

// per block: external / local entry counts and the largest staging group (aligned runs of 32 rows from the block start)
__global__ void k_block_stats(const int64_t *__restrict__ loc_rp, const int64_t *__restrict__ ext_rp,
                              const uint32_t *__restrict__ bounds, int nb, unsigned long long *__restrict__ ext_nnz,
                              unsigned long long *__restrict__ loc_nnz, unsigned int *__restrict__ max_stage) {
  const int b = blockIdx.x;
  if (b >= nb) return;
  const uint32_t lo = bounds[b], hi = bounds[b + 1];
  unsigned int m = 0;
  for (uint32_t j0 = lo + 32u * threadIdx.x; j0 < hi; j0 += 32u * blockDim.x) {
    const uint32_t j1 = min(hi, j0 + 32u);
    const int64_t e0 = loc_rp[j0] & ~3ll;
    m = max(m, (unsigned int)(loc_rp[j1] - e0));
  }
  atomicMax(&max_stage[b], m);
  if (threadIdx.x == 0) {
    ext_nnz[b] = (unsigned long long)(ext_rp[hi] - ext_rp[lo]);
    loc_nnz[b] = (unsigned long long)(loc_rp[hi] - loc_rp[lo]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// level space: inside every block, rows are re-stored sorted by (DAG level, row index)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int block_of(const uint32_t *__restrict__ bounds, int nb, uint32_t j) {
  int lo = 0, hi = nb;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (bounds[mid] <= j) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void k_levels_to_int(const double *__restrict__ lev, uint32_t N, const uint32_t *__restrict__ bounds, int nb,
                                uint32_t *__restrict__ ilev, unsigned int *__restrict__ blkmax) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (; j < N; j += stride) {
    const double l = lev[j];
    const uint32_t il = (l >= 0.0 && l < 4.0e9) ? (uint32_t)l : 0u;   // NaN (watchdog) -> level 0; caught by validation
    ilev[j] = il;
    atomicMax(&blkmax[block_of(bounds, nb, j)], il);
  }
}

// histogram of (block, level) and, second use, stable-by-sort placement of the rows into their level segment
__global__ void k_level_count(const uint32_t *__restrict__ ilev, uint32_t N, const uint32_t *__restrict__ bounds, int nb,
                              const int64_t *__restrict__ lvl_off, int64_t *__restrict__ cnt) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (; j < N; j += stride)
    atomicAdd((unsigned long long *)&cnt[lvl_off[block_of(bounds, nb, j)] + ilev[j]], 1ull);
}
__global__ void k_level_place(const uint32_t *__restrict__ ilev, uint32_t N, const uint32_t *__restrict__ bounds, int nb,
                              const int64_t *__restrict__ lvl_off, int64_t *__restrict__ cursor, uint32_t *__restrict__ perm) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (; j < N; j += stride) {
    const int64_t pos = (int64_t)atomicAdd((unsigned long long *)&cursor[lvl_off[block_of(bounds, nb, j)] + ilev[j]], 1ull);
    perm[pos] = j;
  }
}
__global__ void k_invert_perm(const uint32_t *__restrict__ perm, uint32_t N, uint32_t *__restrict__ inv,
                              uint32_t *__restrict__ vecidx, int reversed) {
  uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (; v < N; v += stride) {
    const uint32_t j = perm[v];
    inv[j] = v;
    vecidx[v] = reversed ? N - 1 - j : j;
  }
}
// Row lengths of the level-space matrices.  Every ND block is cut into segments of `seg` level-space rows (the
// chain kernel's solution window covers one segment); a local entry whose column falls into an EARLIER segment of
// the block becomes an external entry (its row is solved by an earlier launch), so that the chain kernel never has to
// read a column older than its window.
__global__ void k_perm_count(const int64_t *__restrict__ lrp, const uint32_t *__restrict__ lcol, const int64_t *__restrict__ erp,
                             const uint32_t *__restrict__ perm, const uint32_t *__restrict__ inv, uint32_t N,
                             const uint32_t *__restrict__ bounds, int nb, uint32_t seg, int64_t *__restrict__ nl_out,
                             int64_t *__restrict__ ne_out) {
  uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (; v < N; v += stride) {
    const uint32_t j = perm[v];
    const uint32_t blo = bounds[block_of(bounds, nb, j)];
    const uint32_t segv = (v - blo) / seg;
    const int64_t ls = lrp[j], le = lrp[j + 1];
    int64_t same = 0;
    for (int64_t k = ls; k < le; k++) same += ((inv[blo + lcol[k]] - blo) / seg == segv);
    nl_out[v] = same;
    ne_out[v] = (erp[j + 1] - erp[j]) + (le - ls - same);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { nl_out[N] = 0; ne_out[N] = 0; }
}
// Copy row perm[v] of the solve-space matrices to row v of the level-space ones.
// loc columns: block-relative solve-space -> segment-relative level-space; ext columns: -> vector-space.
__global__ void k_permute_rows(const int64_t *__restrict__ lrp, const uint32_t *__restrict__ lcol, const double *__restrict__ lval,
                               const int64_t *__restrict__ erp, const uint32_t *__restrict__ ecol, const double *__restrict__ eval,
                               const uint32_t *__restrict__ perm, const uint32_t *__restrict__ inv, uint32_t N,
                               const uint32_t *__restrict__ bounds, int nb, uint32_t seg, int reversed,
                               const int64_t *__restrict__ nlrp, uint32_t *__restrict__ nlcol, double *__restrict__ nlval,
                               const int64_t *__restrict__ nerp, uint32_t *__restrict__ necol, double *__restrict__ neval) {
  uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (; v < N; v += stride) {
    const uint32_t j = perm[v];
    const uint32_t blo = bounds[block_of(bounds, nb, j)];
    const uint32_t segv = (v - blo) / seg;
    const uint32_t seglo = blo + segv * seg;
    int64_t ld = nlrp[v], ed = nerp[v];
    for (int64_t k = erp[j]; k < erp[j + 1]; k++) {
      const uint32_t c = ecol[k];
      necol[ed] = reversed ? N - 1 - c : c;
      neval[ed++] = eval[k];
    }
    const int64_t ls = lrp[j], le = lrp[j + 1];
    const double dinv = lval[le - 1];   // diagonal slot = 1/diag; off-diagonals are stored as -v/diag
    for (int64_t k = ls; k < le; k++) {
      const uint32_t c = blo + lcol[k];   // solve-space column
      const uint32_t vc = inv[c];
      if ((vc - blo) / seg == segv) {
        nlcol[ld] = vc - seglo;
        nlval[ld++] = lval[k];
      } else {
        // earlier segment: un-scale (the pre kernel computes rhs - sum v x, the chain kernel applies 1/diag)
        necol[ed] = reversed ? N - 1 - c : c;
        neval[ed++] = -lval[k] / dinv;
      }
    }
  }
}
__global__ void k_gather_u32(const uint32_t *__restrict__ src, const uint32_t *__restrict__ idx, uint32_t n,
                             uint32_t *__restrict__ dst) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = src[idx[i]];
}
// one thread per 32-row group: bit l of the mask is set when row l of the group starts a new DAG level
// (bit 0 always: a batch of the critical warp never crosses a group boundary)
__global__ void k_group_masks(const uint32_t *__restrict__ lvl, const uint32_t *__restrict__ rbounds, int nrb,
                              const uint32_t *__restrict__ grp0, uint32_t ngrp, uint32_t *__restrict__ mask) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (; g < ngrp; g += stride) {
    int lo = 0, hi = nrb;                      // refined block of the group: largest rb with grp0[rb] <= g
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (grp0[mid] <= g) lo = mid; else hi = mid;
    }
    const uint32_t j0 = rbounds[lo] + 32u * (g - grp0[lo]);
    const uint32_t nr = min(32u, rbounds[lo + 1] - j0);
    uint32_t m = 1u;
    for (uint32_t l = 1; l < nr; l++)
      if (lvl[j0 + l] != lvl[j0 + l - 1]) m |= (1u << l);
    mask[g] = m;
  }
}
// Number of trailing off-diagonal entries of every level-space row whose column lies within the last `K` DAG
// levels (at most RCG_NEAR_MAX): these "near" entries are applied by the critical warp of k_tri_chain_lv, everything
// older by the helper warps.  The count is packed into bits 28..31 of the diagonal slot's column entry.
__global__ void k_pack_near(const int64_t *__restrict__ rp, uint32_t *__restrict__ col, const uint32_t *__restrict__ lvl,
                            uint32_t N, const uint32_t *__restrict__ rbounds, int nrb, uint32_t K, uint32_t nmax) {
  uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (; v < N; v += stride) {
    const uint32_t lo = rbounds[block_of(rbounds, nrb, v)];
    const uint32_t lv = lvl[v];
    const int64_t s = rp[v], d = rp[v + 1] - 1;
    uint32_t nn = 0;
    for (int64_t k = d - 1; k >= s && nn < nmax; k--) {
      if (lvl[lo + col[k]] + K < lv) break;
      nn++;
    }
    col[d] = (col[d] & 0x0FFFFFFFu) | (nn << 28);
  }
}
// after sorting the level-space rows by column the diagonal (largest column) must be last
__global__ void k_check_diag_last(const int64_t *__restrict__ rp, const uint32_t *__restrict__ col, uint32_t N,
                                  const uint32_t *__restrict__ rbounds, int nrb, int *err) {
  uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t stride = gridDim.x * blockDim.x;
  bool bad = false;
  for (; v < N; v += stride) bad |= (col[rp[v + 1] - 1] != v - rbounds[block_of(rbounds, nrb, v)]);
  if (bad) atomicExch(err, 1);
}

// ---------------------------------------------------------------------------------------------------------
// nested-dissection tree from `part` (post-order: [left subtree, right subtree, separator])
// ---------------------------------------------------------------------------------------------------------
struct TreeInfo {
  std::vector<int> depth;        // per block
  std::vector<uint32_t> sub_lo;  // first row of the subtree rooted at the block
  int max_depth = 0;
};

void build_tree(const std::vector<uint32_t> &bounds, int start, int total, int depth, TreeInfo &t) {
  // same arithmetic as /root/reference/c++/rchol_lap/rchol_lap.cpp:254-261
  if (total == 1) {
    t.depth[start] = depth;
    t.sub_lo[start] = bounds[start];
  } else {
    int sep = start + total - 1, half = (total - 1) / 2;
    t.depth[sep] = depth;
    t.sub_lo[sep] = bounds[start];
    build_tree(bounds, start, half, depth + 1, t);
    build_tree(bounds, start + half, half, depth + 1, t);
  }
  t.max_depth = std::max(t.max_depth, depth);
}

int alloc_csr(rcg_handle *h, CsrDev &m, uint64_t N, int64_t nnz, bool with_rowptr) {
  m.nnz = nnz;
  if (with_rowptr) RCG_CUDA(h, cudaMalloc(&m.rowptr, sizeof(int64_t) * (N + 1)));
  RCG_CUDA(h, cudaMalloc(&m.col, sizeof(uint32_t) * (size_t)(nnz + 8)));
  RCG_CUDA(h, cudaMalloc(&m.val, sizeof(double) * (size_t)(nnz + 8)));
  RCG_CUDA(h, cudaMemsetAsync(m.col + nnz, 0, sizeof(uint32_t) * 8, h->stream));
  RCG_CUDA(h, cudaMemsetAsync(m.val + nnz, 0, sizeof(double) * 8, h->stream));
  return RCG_OK;
}

// `comb` = the direction's lower-triangular matrix with sorted rows (diagonal last).  Splits it into d.M.loc /
// d.M.ext, frees `comb`, and builds the dependency groups.
int finish_direction(rcg_handle *h, DirectionDev &d, CsrDev &comb, const std::vector<uint32_t> &bounds_solve,
                     const std::vector<int> &depth_solve, int max_depth, bool root_first) {
  const uint32_t N = (uint32_t)h->N;
  const int nb = (int)bounds_solve.size() - 1;
  if (rcg_use_blocked(h)) return rcg_build_blocked(h, d, comb, bounds_solve, depth_solve, max_depth, root_first);
  uint32_t *dbounds = nullptr;
  int *derr = nullptr;
  RCG_CUDA(h, cudaMalloc(&dbounds, sizeof(uint32_t) * (nb + 1)));
  RCG_CUDA(h, cudaMalloc(&derr, sizeof(int)));
  RCG_CUDA(h, cudaMemcpyAsync(dbounds, bounds_solve.data(), sizeof(uint32_t) * (nb + 1), cudaMemcpyHostToDevice, h->stream));
  RCG_CUDA(h, cudaMemsetAsync(derr, 0, sizeof(int), h->stream));

  CsrDev &loc = d.M.loc, &ext = d.M.ext;
  RCG_CUDA(h, cudaMalloc(&ext.rowptr, sizeof(int64_t) * ((size_t)N + 1)));
  RCG_CUDA(h, cudaMemsetAsync(ext.rowptr, 0, sizeof(int64_t) * ((size_t)N + 1), h->stream));
  k_count_external<<<grid_for(h, N, 256), 256, 0, h->stream>>>(comb.rowptr, comb.col, comb.val, N, dbounds, nb,
                                                              ext.rowptr, derr);
  h->stats.kernel_launches += 1;
  int herr = 0;
  RCG_CUDA(h, cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  RCG_CUDA(h, cudaFree(derr));
  if (herr) {
    h->err = "factor row without a trailing diagonal after transposition (G is not triangular)";
    return RCG_ERR_STRUCTURE;
  }
  RCG_TRY(exclusive_scan_inplace(h, ext.rowptr, (int64_t)N + 1));
  int64_t ext_total = 0;
  RCG_CUDA(h, cudaMemcpy(&ext_total, ext.rowptr + N, sizeof(int64_t), cudaMemcpyDeviceToHost));
  // comb.rowptr becomes the local row pointer array
  k_sub_rowptr<<<grid_for(h, (int64_t)N + 1, 256), 256, 0, h->stream>>>(comb.rowptr, ext.rowptr, (int64_t)N + 1);
  loc.rowptr = comb.rowptr;
  comb.rowptr = nullptr;
  RCG_TRY(alloc_csr(h, loc, N, comb.nnz - ext_total, false));
  RCG_TRY(alloc_csr(h, ext, N, ext_total, false));
  k_split_rows<<<grid_for(h, (int64_t)N * 8, 256), 256, 0, h->stream>>>(loc.rowptr, ext.rowptr, comb.col, comb.val, N,
                                                                       dbounds, nb, loc.col, loc.val, ext.col, ext.val);
  h->stats.kernel_launches += 2;
  RCG_CUDA(h, cudaGetLastError());
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  rcg_free_csr(comb);

  // per-block statistics (external / local entry counts, largest 32-row staging group)
  std::vector<unsigned long long> hext(nb), hloc(nb);
  std::vector<unsigned int> hstage(nb);
  auto block_stats = [&]() -> int {
    unsigned long long *dext = nullptr, *dloc = nullptr;
    unsigned int *dstage = nullptr;
    RCG_CUDA(h, cudaMalloc(&dext, sizeof(unsigned long long) * nb));
    RCG_CUDA(h, cudaMalloc(&dloc, sizeof(unsigned long long) * nb));
    RCG_CUDA(h, cudaMalloc(&dstage, sizeof(unsigned int) * nb));
    RCG_CUDA(h, cudaMemsetAsync(dstage, 0, sizeof(unsigned int) * nb, h->stream));
    k_block_stats<<<nb, 256, 0, h->stream>>>(loc.rowptr, ext.rowptr, dbounds, nb, dext, dloc, dstage);
    h->stats.kernel_launches += 1;
    RCG_CUDA(h, cudaMemcpyAsync(hext.data(), dext, sizeof(unsigned long long) * nb, cudaMemcpyDeviceToHost, h->stream));
    RCG_CUDA(h, cudaMemcpyAsync(hloc.data(), dloc, sizeof(unsigned long long) * nb, cudaMemcpyDeviceToHost, h->stream));
    RCG_CUDA(h, cudaMemcpyAsync(hstage.data(), dstage, sizeof(unsigned int) * nb, cudaMemcpyDeviceToHost, h->stream));
    RCG_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(dext); cudaFree(dloc); cudaFree(dstage);
    return RCG_OK;
  };
  RCG_TRY(block_stats());

  // groups: forward = deepest level first, backward (reversed space) = root first
  d.groups.clear();
  d.blocks_host.clear();
  std::vector<int> bidx;   // bounds-order index of every entry of blocks_host
  for (int g = 0; g <= max_depth; g++) {
    const int want = root_first ? g : max_depth - g;
    GroupHost G;
    G.first = (int)d.blocks_host.size();
    for (int b = 0; b < nb; b++) {
      if (depth_solve[b] != want) continue;
      BlockDesc bd{bounds_solve[b], bounds_solve[b + 1]};
      if (bd.hi == bd.lo) continue;   // empty separator
      d.blocks_host.push_back(bd);
      bidx.push_back(b);
      G.count++;
      G.max_rows = std::max(G.max_rows, bd.hi - bd.lo);
      G.rows += bd.hi - bd.lo;
      G.ext_nnz += (int64_t)hext[b];
      G.loc_nnz += (int64_t)hloc[b];
      G.max_stage = std::max(G.max_stage, hstage[b]);
    }
    if (G.count > 0) d.groups.push_back(G);
  }
  RCG_CUDA(h, cudaMalloc(&d.blocks, sizeof(BlockDesc) * std::max<size_t>(1, d.blocks_host.size())));
  RCG_CUDA(h, cudaMemcpyAsync(d.blocks, d.blocks_host.data(), sizeof(BlockDesc) * d.blocks_host.size(),
                              cudaMemcpyHostToDevice, h->stream));
  RCG_CUDA(h, cudaMalloc(&d.w, sizeof(double) * ((size_t)N + 4)));
  RCG_CUDA(h, cudaMemsetAsync(d.w, 0, sizeof(double) * ((size_t)N + 4), h->stream));
  RCG_CUDA(h, cudaMalloc(&d.vecidx, sizeof(uint32_t) * (size_t)N));

  // ---- level space ------------------------------------------------------------------------------------------
  // 1. DAG level of every row inside its block: the chain kernel itself in MODE 1 (level = 1 + max over the row's
  //    local columns), on the solve-space matrices.
  RCG_TRY(rcg_compute_levels(h, loc, d.blocks, d.groups, d.w));
  uint32_t *ilev = nullptr, *perm = nullptr, *inv = nullptr;
  unsigned int *dblkmax = nullptr;
  RCG_CUDA(h, cudaMalloc(&ilev, sizeof(uint32_t) * (size_t)N));
  RCG_CUDA(h, cudaMalloc(&perm, sizeof(uint32_t) * (size_t)N));
  RCG_CUDA(h, cudaMalloc(&inv, sizeof(uint32_t) * (size_t)N));
  RCG_CUDA(h, cudaMalloc(&dblkmax, sizeof(unsigned int) * nb));
  RCG_CUDA(h, cudaMemsetAsync(dblkmax, 0, sizeof(unsigned int) * nb, h->stream));
  k_levels_to_int<<<grid_for(h, N, 256), 256, 0, h->stream>>>(d.w, N, dbounds, nb, ilev, dblkmax);
  h->stats.kernel_launches += 1;
  std::vector<unsigned int> blkmax(nb);
  RCG_CUDA(h, cudaMemcpyAsync(blkmax.data(), dblkmax, sizeof(unsigned int) * nb, cudaMemcpyDeviceToHost, h->stream));
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  // 2. counting sort by (block, level); rows of one level are then sorted by index, so the order is deterministic
  std::vector<int64_t> lvl_off(nb + 1, 0);
  for (int b = 0; b < nb; b++)
    lvl_off[b + 1] = lvl_off[b] + (bounds_solve[b + 1] > bounds_solve[b] ? (int64_t)blkmax[b] + 1 : 0);
  const int64_t nlev = lvl_off[nb];
  int64_t *dlvl_off = nullptr, *lptr = nullptr, *cursor = nullptr;
  RCG_CUDA(h, cudaMalloc(&dlvl_off, sizeof(int64_t) * (nb + 1)));
  RCG_CUDA(h, cudaMalloc(&lptr, sizeof(int64_t) * ((size_t)nlev + 1)));
  RCG_CUDA(h, cudaMalloc(&cursor, sizeof(int64_t) * ((size_t)nlev + 1)));
  RCG_CUDA(h, cudaMemcpyAsync(dlvl_off, lvl_off.data(), sizeof(int64_t) * (nb + 1), cudaMemcpyHostToDevice, h->stream));
  RCG_CUDA(h, cudaMemsetAsync(lptr, 0, sizeof(int64_t) * ((size_t)nlev + 1), h->stream));
  k_level_count<<<grid_for(h, N, 256), 256, 0, h->stream>>>(ilev, N, dbounds, nb, dlvl_off, lptr);
  h->stats.kernel_launches += 1;
  RCG_TRY(exclusive_scan_inplace(h, lptr, nlev + 1));
  RCG_CUDA(h, cudaMemcpyAsync(cursor, lptr, sizeof(int64_t) * ((size_t)nlev + 1), cudaMemcpyDeviceToDevice, h->stream));
  k_level_place<<<grid_for(h, N, 256), 256, 0, h->stream>>>(ilev, N, dbounds, nb, dlvl_off, cursor, perm);
  h->stats.kernel_launches += 1;
  RCG_TRY(sort_segments(h, lptr, perm, nullptr, (uint32_t)nlev));
  k_invert_perm<<<grid_for(h, N, 256), 256, 0, h->stream>>>(perm, N, inv, d.vecidx, d.reversed ? 1 : 0);
  h->stats.kernel_launches += 1;
  h->stats.reserved[root_first ? 3 : 2] = (double)nlev;   // total DAG levels of the direction (sum over blocks)
  cudaFree(dlvl_off); cudaFree(lptr); cudaFree(cursor); cudaFree(dblkmax);
  // 3. level-space matrices, with every block cut into window-sized segments
  uint32_t seg = 1;
  {
    const uint32_t cw = h->opt.chain_window > 0 ? (uint32_t)h->opt.chain_window : 2048u;
    while ((seg << 1) <= cw && (seg << 1) != 0) seg <<= 1;
    seg = std::max(64u, seg * 2u);   // one segment = the chain kernel's window (current + previous chunk)
  }
  CsrDev nloc, next;
  RCG_CUDA(h, cudaMalloc(&nloc.rowptr, sizeof(int64_t) * ((size_t)N + 4)));
  RCG_CUDA(h, cudaMalloc(&next.rowptr, sizeof(int64_t) * ((size_t)N + 4)));
  RCG_CUDA(h, cudaMemsetAsync(nloc.rowptr, 0, sizeof(int64_t) * ((size_t)N + 4), h->stream));
  RCG_CUDA(h, cudaMemsetAsync(next.rowptr, 0, sizeof(int64_t) * ((size_t)N + 4), h->stream));
  k_perm_count<<<grid_for(h, N, 256), 256, 0, h->stream>>>(loc.rowptr, loc.col, ext.rowptr, perm, inv, N, dbounds, nb, seg,
                                                          nloc.rowptr, next.rowptr);
  h->stats.kernel_launches += 1;
  RCG_TRY(exclusive_scan_inplace(h, nloc.rowptr, (int64_t)N + 1));
  RCG_TRY(exclusive_scan_inplace(h, next.rowptr, (int64_t)N + 1));
  int64_t nl_total = 0, ne_total = 0;
  RCG_CUDA(h, cudaMemcpy(&nl_total, nloc.rowptr + N, sizeof(int64_t), cudaMemcpyDeviceToHost));
  RCG_CUDA(h, cudaMemcpy(&ne_total, next.rowptr + N, sizeof(int64_t), cudaMemcpyDeviceToHost));
  RCG_TRY(alloc_csr(h, nloc, N, nl_total, false));
  RCG_TRY(alloc_csr(h, next, N, ne_total, false));
  k_permute_rows<<<grid_for(h, N, 256), 256, 0, h->stream>>>(loc.rowptr, loc.col, loc.val, ext.rowptr, ext.col, ext.val, perm,
                                                            inv, N, dbounds, nb, seg, d.reversed ? 1 : 0, nloc.rowptr,
                                                            nloc.col, nloc.val, next.rowptr, next.col, next.val);
  h->stats.kernel_launches += 1;
  RCG_CUDA(h, cudaGetLastError());
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  // level of every level-space row (for the group masks below)
  uint32_t *lvl_ls = nullptr;
  RCG_CUDA(h, cudaMalloc(&lvl_ls, sizeof(uint32_t) * (size_t)N));
  k_gather_u32<<<grid_for(h, N, 256), 256, 0, h->stream>>>(ilev, perm, N, lvl_ls);
  h->stats.kernel_launches += 1;
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  rcg_free_csr(loc);
  rcg_free_csr(ext);
  cudaFree(perm); cudaFree(inv); cudaFree(ilev);
  loc = nloc;
  ext = next;
  RCG_TRY(sort_segments(h, loc.rowptr, loc.col, loc.val, N));

  // 4. refined blocks (one per segment) and dependency groups: tree level by tree level, segment by segment
  std::vector<uint32_t> rbounds;      // ascending boundaries of the refined blocks (level space)
  std::vector<int> r_parent, r_segidx;
  for (int b = 0; b < nb; b++) {
    const uint32_t lo = bounds_solve[b], hi = bounds_solve[b + 1];
    int k = 0;
    for (uint32_t s0 = lo; s0 < hi; s0 += seg, k++) {
      rbounds.push_back(s0);
      r_parent.push_back(b);
      r_segidx.push_back(k);
    }
  }
  rbounds.push_back(N);
  const int nrb = (int)rbounds.size() - 1;
  cudaFree(dbounds);
  RCG_CUDA(h, cudaMalloc(&dbounds, sizeof(uint32_t) * (nrb + 1)));
  RCG_CUDA(h, cudaMemcpyAsync(dbounds, rbounds.data(), sizeof(uint32_t) * (nrb + 1), cudaMemcpyHostToDevice, h->stream));
  {
    int *derr2 = nullptr, herr2 = 0;
    RCG_CUDA(h, cudaMalloc(&derr2, sizeof(int)));
    RCG_CUDA(h, cudaMemsetAsync(derr2, 0, sizeof(int), h->stream));
    k_check_diag_last<<<grid_for(h, N, 256), 256, 0, h->stream>>>(loc.rowptr, loc.col, N, dbounds, nrb, derr2);
    h->stats.kernel_launches += 1;
    RCG_CUDA(h, cudaMemcpyAsync(&herr2, derr2, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    RCG_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(derr2);
    if (herr2) {
      h->err = "internal error: level ordering is not a topological order of the factor's dependency graph";
      return RCG_ERR_STRUCTURE;
    }
  }
  std::vector<unsigned long long> rext(nrb), rloc(nrb);
  std::vector<unsigned int> rstage(nrb);
  {
    unsigned long long *dext = nullptr, *dloc = nullptr;
    unsigned int *dstage = nullptr;
    RCG_CUDA(h, cudaMalloc(&dext, sizeof(unsigned long long) * nrb));
    RCG_CUDA(h, cudaMalloc(&dloc, sizeof(unsigned long long) * nrb));
    RCG_CUDA(h, cudaMalloc(&dstage, sizeof(unsigned int) * nrb));
    RCG_CUDA(h, cudaMemsetAsync(dstage, 0, sizeof(unsigned int) * nrb, h->stream));
    k_block_stats<<<nrb, 256, 0, h->stream>>>(loc.rowptr, ext.rowptr, dbounds, nrb, dext, dloc, dstage);
    h->stats.kernel_launches += 1;
    RCG_CUDA(h, cudaMemcpyAsync(rext.data(), dext, sizeof(unsigned long long) * nrb, cudaMemcpyDeviceToHost, h->stream));
    RCG_CUDA(h, cudaMemcpyAsync(rloc.data(), dloc, sizeof(unsigned long long) * nrb, cudaMemcpyDeviceToHost, h->stream));
    RCG_CUDA(h, cudaMemcpyAsync(rstage.data(), dstage, sizeof(unsigned int) * nrb, cudaMemcpyDeviceToHost, h->stream));
    RCG_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(dext); cudaFree(dloc); cudaFree(dstage);
  }
  // 32-row groups of the refined blocks and their level-start masks (critical-warp batching in k_tri_chain_lv)
  std::vector<uint32_t> r_grp0(nrb);
  uint32_t ngrp_total = 0;
  for (int rb = 0; rb < nrb; rb++) { r_grp0[rb] = ngrp_total; ngrp_total += (rbounds[rb + 1] - rbounds[rb] + 31) / 32; }
  {
    uint32_t *dgrp0 = nullptr;
    RCG_CUDA(h, cudaMalloc(&dgrp0, sizeof(uint32_t) * std::max(1, nrb)));
    RCG_CUDA(h, cudaMemcpyAsync(dgrp0, r_grp0.data(), sizeof(uint32_t) * nrb, cudaMemcpyHostToDevice, h->stream));
    RCG_CUDA(h, cudaMalloc(&d.grp_mask, sizeof(uint32_t) * std::max<uint32_t>(1, ngrp_total)));
    k_group_masks<<<grid_for(h, ngrp_total, 256), 256, 0, h->stream>>>(lvl_ls, dbounds, nrb, dgrp0, ngrp_total, d.grp_mask);
    k_pack_near<<<grid_for(h, N, 256), 256, 0, h->stream>>>(loc.rowptr, loc.col, lvl_ls, N, dbounds, nrb, 12u, 6u);
    h->stats.kernel_launches += 2;
    RCG_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(dgrp0);
    cudaFree(lvl_ls);
  }
  std::vector<GroupHost> tree_groups;
  tree_groups.swap(d.groups);
  d.blocks_host.clear();
  for (int g = 0; g <= max_depth; g++) {
    const int want = root_first ? g : max_depth - g;
    for (int k = 0;; k++) {        // segment index
      GroupHost G;
      G.depth = want;
      G.first = (int)d.blocks_host.size();
      for (int rb = 0; rb < nrb; rb++) {
        if (r_segidx[rb] != k || depth_solve[r_parent[rb]] != want) continue;
        BlockDesc bd{rbounds[rb], rbounds[rb + 1], r_grp0[rb], 0u};
        d.blocks_host.push_back(bd);
        G.count++;
        G.max_rows = std::max(G.max_rows, bd.hi - bd.lo);
        G.rows += bd.hi - bd.lo;
        G.ext_nnz += (int64_t)rext[rb];
        G.loc_nnz += (int64_t)rloc[rb];
        G.max_stage = std::max(G.max_stage, rstage[rb]);
      }
      if (G.count == 0) break;
      d.groups.push_back(G);
    }
  }
  cudaFree(d.blocks);
  d.blocks = nullptr;
  RCG_CUDA(h, cudaMalloc(&d.blocks, sizeof(BlockDesc) * std::max<size_t>(1, d.blocks_host.size())));
  RCG_CUDA(h, cudaMemcpyAsync(d.blocks, d.blocks_host.data(), sizeof(BlockDesc) * d.blocks_host.size(),
                              cudaMemcpyHostToDevice, h->stream));
  cudaFree(dbounds);
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  return RCG_OK;
}

}  // namespace

int rcg_exclusive_scan(rcg_handle *h, int64_t *data, int64_t n) { return exclusive_scan_inplace(h, data, n); }

void rcg_free_csr(CsrDev &a) {
  cudaFree(a.rowptr); cudaFree(a.col); cudaFree(a.val);
  a = CsrDev();
}
void rcg_free_direction(DirectionDev &d) {
  rcg_free_csr(d.M.loc);
  rcg_free_csr(d.M.ext);
  cudaFree(d.blocks);
  cudaFree(d.vecidx);
  cudaFree(d.w);
  cudaFree(d.grp_mask);
  rcg_free_blocked(d.bc);
  d = DirectionDev();
}

int rcg_setup_matrix(rcg_handle *h, uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val) {
  if (h->haveG && N != h->N) {   // (checked first: a rejected call leaves the handle's matrix in place)
    h->err = "matrix dimension differs from the factor's";
    return RCG_ERR_INVALID;
  }
  if (h->haveA) { rcg_free_csr(h->A); h->haveA = false; }
  RCG_TRY(upload_csr(h, N, rowPtr, colIdx, val, h->A));
  h->N = N;
  h->haveA = true;
  h->a_resorted = false;
  h->stats.N = N;
  h->stats.nnzA = (uint64_t)h->A.nnz;
  // lanes per row of the SpMV from the row-length histogram (spmv_lanes_from_rows above)
  h->spmv_lanes = h->opt.spmv_lanes > 0 ? h->opt.spmv_lanes : spmv_lanes_from_rows(N, rowPtr, nullptr);
  return RCG_OK;
}

// ---------------------------------------------------------------------------------------------------------
// permutation steps either side of the path: reorder(A, P, B) of the reference
// (/root/reference/c++/util/util.cpp:16-57) and the vector permutations (util.hpp:147-155,
// python/ex_laplace_parallel.py:31-32) on the device
// ---------------------------------------------------------------------------------------------------------
namespace {

// inv[P[i]] = i; err is set when P is not a permutation of 0..N-1 (out of range or a value taken twice)
__global__ void k_perm_invert_checked(const uint32_t *__restrict__ P, uint32_t N, uint32_t *__restrict__ inv, int *err) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (; i < N; i += stride) {
    const uint32_t v = P[i];
    if (v >= N) { atomicExch(err, 1); continue; }
    if (atomicExch(inv + v, i) != 0xFFFFFFFFu) atomicExch(err, 1);
  }
}

// row lengths of B = A(P,P): len_B[i] = len_A[P[i]]
__global__ void k_perm_row_lengths(const int64_t *__restrict__ rp, const uint32_t *__restrict__ P, uint32_t N,
                                   int64_t *__restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (; i < N; i += stride) {
    const uint32_t s = P[i];
    out[i] = rp[s + 1] - rp[s];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[N] = 0;
}

// B row i <- A row P[i] with the columns renamed by the inverse permutation; 8 lanes per row (rows of the path's
// matrices hold 3-7 entries), unsorted: sort_segments() re-sorts every row afterwards like the reference does
__global__ void k_perm_fill_rows(const int64_t *__restrict__ rp, const uint32_t *__restrict__ col,
                                 const double *__restrict__ val, const uint32_t *__restrict__ P,
                                 const uint32_t *__restrict__ inv, const int64_t *__restrict__ nrp, uint32_t N,
                                 uint32_t *__restrict__ ncol, double *__restrict__ nval) {
  const uint32_t sub = threadIdx.x & 7u;
  uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const uint32_t stride = (gridDim.x * blockDim.x) >> 3;
  for (; i < N; i += stride) {
    const uint32_t s = P[i];
    const int64_t a0 = rp[s], a1 = rp[s + 1], b0 = nrp[i];
    for (int64_t e = a0 + sub; e < a1; e += 8) {
      ncol[b0 + (e - a0)] = inv[col[e]];
      nval[b0 + (e - a0)] = val[e];
    }
  }
}

__global__ void k_vec_gather(const double *__restrict__ src, const uint32_t *__restrict__ P, uint32_t N,
                             double *__restrict__ dst) {   // dst[i] = src[P[i]]
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (; i < N; i += stride) dst[i] = src[P[i]];
}

__global__ void k_vec_scatter(const double *__restrict__ src, const uint32_t *__restrict__ P, uint32_t N,
                              double *__restrict__ dst) {  // dst[P[i]] = src[i]
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (; i < N; i += stride) dst[P[i]] = src[i];
}

__global__ void k_widen_u32(const uint32_t *__restrict__ in, int64_t n, uint64_t *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = in[i];
}

// uploads P (narrowed to 32 bits) and builds the checked inverse; on success *dP (and *dinv if wanted) are device arrays
int upload_permutation(rcg_handle *h, uint64_t N, const uint64_t *P, uint32_t **dP, uint32_t **dinv) {
  if (!P) { h->err = "null permutation"; return RCG_ERR_INVALID; }
  if (N == 0 || N >= 0xFFFFFFFFull) { h->err = "matrix dimension must be in [1, 2^32-2]"; return RCG_ERR_INVALID; }
  const uint32_t n32 = (uint32_t)N;
  uint32_t *p32 = nullptr, *inv = nullptr;
  int *derr = nullptr;
  RCG_CUDA(h, cudaMalloc(&p32, sizeof(uint32_t) * N));
  RCG_CUDA(h, cudaMalloc(&inv, sizeof(uint32_t) * N));
  RCG_CUDA(h, cudaMalloc(&derr, sizeof(int)));
  RCG_CUDA(h, cudaMemsetAsync(derr, 0, sizeof(int), h->stream));
  RCG_CUDA(h, cudaMemsetAsync(inv, 0xFF, sizeof(uint32_t) * N, h->stream));
  bool bad = false;
  double t0 = wall_ms();
  RCG_TRY(staged_narrow(h, p32, P, (size_t)N, N, &bad));
  h->stats.upload_ms += wall_ms() - t0;
  h->stats.h2d_bytes += sizeof(uint32_t) * N;
  int herr = 0;
  if (!bad) {
    k_perm_invert_checked<<<grid_for(h, n32, 256), 256, 0, h->stream>>>(p32, n32, inv, derr);
    h->stats.kernel_launches += 1;
    RCG_CUDA(h, cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  }
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(derr);
  if (bad || herr) {
    cudaFree(p32); cudaFree(inv);
    h->err = "P is not a permutation of 0..N-1";
    return RCG_ERR_INVALID;
  }
  *dP = p32;
  if (dinv) *dinv = inv; else cudaFree(inv);
  return RCG_OK;
}

}  // namespace

int rcg_setup_permutation(rcg_handle *h, uint64_t N, const uint64_t *P) {
  if ((h->haveA || h->haveG) && N != h->N) { h->err = "permutation length differs from the matrix dimension"; return RCG_ERR_INVALID; }
  uint32_t *dP = nullptr;
  RCG_TRY(upload_permutation(h, N, P, &dP, nullptr));
  cudaFree(h->perm);
  h->perm = dP;
  h->N = N;
  return RCG_OK;
}

int rcg_setup_matrix_permuted(rcg_handle *h, uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val,
                              const uint64_t *P) {
  if (h->haveG && N != h->N) { h->err = "matrix dimension differs from the factor's"; return RCG_ERR_INVALID; }
  uint32_t *dP = nullptr, *dinv = nullptr;
  RCG_TRY(upload_permutation(h, N, P, &dP, &dinv));
  CsrDev A0;
  int rc = upload_csr(h, N, rowPtr, colIdx, val, A0);
  if (rc != RCG_OK) { cudaFree(dP); cudaFree(dinv); return rc; }
  double t0 = wall_ms();
  const uint32_t n32 = (uint32_t)N;
  CsrDev B;
  B.nnz = A0.nnz;
  RCG_CUDA(h, cudaMalloc(&B.rowptr, sizeof(int64_t) * (N + 1)));
  RCG_CUDA(h, cudaMalloc(&B.col, sizeof(uint32_t) * (size_t)A0.nnz));
  RCG_CUDA(h, cudaMalloc(&B.val, sizeof(double) * (size_t)A0.nnz));
  k_perm_row_lengths<<<grid_for(h, n32, 256), 256, 0, h->stream>>>(A0.rowptr, dP, n32, B.rowptr);
  h->stats.kernel_launches += 1;
  RCG_TRY(exclusive_scan_inplace(h, B.rowptr, (int64_t)N + 1));
  k_perm_fill_rows<<<grid_for(h, (int64_t)n32 * 8, 256), 256, 0, h->stream>>>(A0.rowptr, A0.col, A0.val, dP, dinv, B.rowptr,
                                                                              n32, B.col, B.val);
  h->stats.kernel_launches += 1;
  RCG_CUDA(h, cudaGetLastError());
  RCG_TRY(sort_segments(h, B.rowptr, B.col, B.val, n32));   // "sort elements" of util.cpp:40
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  rcg_free_csr(A0);
  cudaFree(dinv);
  if (h->haveA) { rcg_free_csr(h->A); h->haveA = false; }
  cudaFree(h->perm);
  h->perm = dP;
  h->A = B;
  h->N = N;
  h->haveA = true;
  h->a_resorted = true;
  h->stats.N = N;
  h->stats.nnzA = (uint64_t)B.nnz;
  // (the histogram of row lengths is invariant under the symmetric permutation: the same plan as for the permuted matrix)
  h->spmv_lanes = h->opt.spmv_lanes > 0 ? h->opt.spmv_lanes : spmv_lanes_from_rows(N, rowPtr, nullptr);
  h->stats.analysis_ms += wall_ms() - t0;
  return RCG_OK;
}

// device vectors: inverse = false: dst[i] = src[P[i]] ; inverse = true: dst[P[i]] = src[i]
int rcg_apply_permutation(rcg_handle *h, const double *src, double *dst, bool inverse) {
  if (!h->perm) { h->err = "no permutation set (rcg_set_permutation / rcg_set_matrix_permuted)"; return RCG_ERR_STATE; }
  const uint32_t n32 = (uint32_t)h->N;
  if (inverse) k_vec_scatter<<<grid_for(h, n32, 256), 256, 0, h->stream>>>(src, h->perm, n32, dst);
  else k_vec_gather<<<grid_for(h, n32, 256), 256, 0, h->stream>>>(src, h->perm, n32, dst);
  h->stats.kernel_launches += 1;
  RCG_CUDA(h, cudaGetLastError());
  return RCG_OK;
}

// SURVEY.md 8f row 1 (the reference's reuse flow, python/rchol/rchol.py:25-40, python/ex_reuse_partition.py: a new matrix
// with the SAME sparsity pattern keeps perm / part): only the values of A cross PCIe (8 B per entry instead of 12 B plus the
// row pointers), the structure on the device, the SpMV's lane choice and the captured iteration graph stay as they are.
int rcg_refresh_matrix_values(rcg_handle *h, uint64_t nnz, const double *val) {
  if (!h->haveA) { h->err = "rcg_update_matrix_values: no matrix set"; return RCG_ERR_STATE; }
  if (h->a_resorted) {
    h->err = "rcg_update_matrix_values: the matrix was permuted and re-sorted on the device (rcg_set_matrix_permuted); its "
             "entries are not in the caller's order -- call rcg_set_matrix_permuted again";
    return RCG_ERR_STATE;
  }
  if (!val || nnz != (uint64_t)h->A.nnz) { h->err = "rcg_update_matrix_values: nnz differs from the matrix on the device"; return RCG_ERR_INVALID; }
  const double t0 = wall_ms();
  RCG_TRY(staged_copy(h, h->A.val, val, sizeof(double) * (size_t)nnz));
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  h->stats.upload_ms += wall_ms() - t0;
  h->stats.h2d_bytes += sizeof(double) * (size_t)nnz;
  return RCG_OK;
}

int rcg_download_matrix(rcg_handle *h, uint64_t *rowPtr, uint64_t *colIdx, double *val) {
  if (!h->haveA) { h->err = "rcg_set_matrix has not been called"; return RCG_ERR_STATE; }
  if (!rowPtr || !colIdx || !val) { h->err = "null output pointer"; return RCG_ERR_INVALID; }
  uint64_t *c64 = nullptr;
  RCG_CUDA(h, cudaMalloc(&c64, sizeof(uint64_t) * (size_t)h->A.nnz));
  k_widen_u32<<<grid_for(h, h->A.nnz, 256), 256, 0, h->stream>>>(h->A.col, h->A.nnz, c64);
  h->stats.kernel_launches += 1;
  RCG_CUDA(h, cudaMemcpyAsync(rowPtr, h->A.rowptr, sizeof(int64_t) * (h->N + 1), cudaMemcpyDeviceToHost, h->stream));
  RCG_CUDA(h, cudaMemcpyAsync(colIdx, c64, sizeof(uint64_t) * (size_t)h->A.nnz, cudaMemcpyDeviceToHost, h->stream));
  RCG_CUDA(h, cudaMemcpyAsync(val, h->A.val, sizeof(double) * (size_t)h->A.nnz, cudaMemcpyDeviceToHost, h->stream));
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(c64);
  return RCG_OK;
}

static int setup_factor_impl(rcg_handle *h, uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val,
                             const std::vector<uint32_t> &bounds, const TreeInfo &tree, bool strict_tree);

// ---------------------------------------------------------------------------------------------------------
// Blocks of a factor that arrives WITHOUT `part` -- the reference's stock signature pcg(A, b, tol, maxit, G, x, relres, itr)
// (/root/reference/c++/util/pcg.hpp:13-16, called like that in ex_laplace_parallel.cpp:46) carries no partition.
// The nested-dissection structure is recovered from G itself.  With jmin(i) = smallest row j < i that has an entry in
// column i of U (the oldest unknown row i of L = U^T depends on), one sweep over i with a stack of open index ranges:
//   no jmin            -> i opens a new range (a new independent chain starts);
//   jmin in the top range -> the top range grows by i;
//   jmin further down  -> every range from the one that holds jmin to the top closes and becomes a child of a new range
//                         whose own rows start at i (a separator: it depends on all of them).
// Ranges of one tree depth never reference each other (a row only reaches into its own range or, through its jmin
// chain, into descendants), which is the block rule of rcg_setup_factor_blocks -- checked on the device afterwards
// (k_validate_depths), so a structure the sweep gets wrong is an error, never a wrong result.  Small ranges (artifacts:
// a row whose lower neighbours all lie outside its leaf) are absorbed by their parent.
// ---------------------------------------------------------------------------------------------------------
namespace {
struct DetNode { uint32_t lo, own_lo, hi; int kid0, nkids; };

bool detect_blocks(uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, std::vector<uint32_t> &bounds, std::vector<int> &depth) {
  const uint32_t NONE = 0xFFFFFFFFu;
  std::vector<uint32_t> jmin(N, NONE);
#pragma omp parallel for schedule(dynamic, 4096)
  for (int64_t j = 0; j < (int64_t)N; j++)
    for (uint64_t p = rowPtr[j]; p < rowPtr[j + 1]; p++) {
      const uint64_t c = colIdx[p];
      if (c <= (uint64_t)j || c >= N) continue;
      uint32_t old = __atomic_load_n(&jmin[c], __ATOMIC_RELAXED);
      while ((uint32_t)j < old && !__atomic_compare_exchange_n(&jmin[c], &old, (uint32_t)j, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    }
  std::vector<DetNode> nodes;
  std::vector<int> kids, stack, tmp;
  for (uint32_t i = 0; i < (uint32_t)N; i++) {
    const uint32_t jm = jmin[i];
    if (jm == NONE || stack.empty()) {
      nodes.push_back({i, i, i + 1u, 0, 0});
      stack.push_back((int)nodes.size() - 1);
      continue;
    }
    DetNode &t = nodes[stack.back()];
    if (jm >= t.lo) { t.hi = i + 1u; continue; }
    tmp.clear();
    uint32_t lo = i;
    while (!stack.empty()) {
      const int c = stack.back();
      stack.pop_back();
      tmp.push_back(c);
      lo = nodes[c].lo;
      if (nodes[c].lo <= jm) break;
    }
    DetNode s{lo, i, i + 1u, (int)kids.size(), (int)tmp.size()};
    for (int q = (int)tmp.size() - 1; q >= 0; q--) kids.push_back(tmp[q]);   // ascending index order
    nodes.push_back(s);
    stack.push_back((int)nodes.size() - 1);
    if (nodes.size() > (size_t)1 << 22) return false;   // not a nested-dissection factor: too fragmented
  }
  // bottom-up (children precede their parents in `nodes`): absorb small trailing children, single-child chains and small
  // subtrees
  const uint32_t thr = (uint32_t)std::min<uint64_t>(2048, std::max<uint64_t>(8, N / 256));
  std::vector<char> gone(nodes.size(), 0);
  for (size_t n = 0; n < nodes.size(); n++) {
    DetNode &d = nodes[n];
    auto absorb_all = [&](auto &&self, int c) -> void {
      DetNode &k = nodes[c];
      for (int q = 0; q < k.nkids; q++) self(self, kids[k.kid0 + q]);
      gone[c] = 1;
    };
    if (d.hi - d.lo < thr) {   // small subtree: one block
      for (int q = 0; q < d.nkids; q++) absorb_all(absorb_all, kids[d.kid0 + q]);
      d.nkids = 0;
      d.own_lo = d.lo;
      continue;
    }
    {   // a small leaf between two siblings joins the last block of the sibling before it (nothing there references it)
      int kept = 0;
      for (int q = 0; q < d.nkids; q++) {
        const int c = kids[d.kid0 + q];
        const DetNode &k = nodes[c];
        if (kept > 0 && k.nkids == 0 && k.hi - k.lo < thr) {
          nodes[kids[d.kid0 + kept - 1]].hi = k.hi;
          gone[c] = 1;
        } else {
          kids[d.kid0 + kept++] = c;
        }
      }
      d.nkids = kept;
    }
    while (d.nkids > 0) {
      const int c = kids[d.kid0 + d.nkids - 1];
      const DetNode &k = nodes[c];
      if (k.nkids == 0 && k.hi - k.lo < thr) { d.own_lo = k.lo; gone[c] = 1; d.nkids--; }
      else break;
    }
    if (d.nkids == 1) {   // chain: merge with the only child (adopt its children)
      const int c = kids[d.kid0];
      const DetNode k = nodes[c];
      d.own_lo = k.own_lo;
      d.kid0 = k.kid0;
      d.nkids = k.nkids;
      gone[c] = 1;
    }
  }
  // depths, top-down (parents follow their children in `nodes`)
  std::vector<int> dep(nodes.size(), 0);
  int max_depth = 0;
  for (size_t n = nodes.size(); n-- > 0;) {
    if (gone[n]) continue;
    const DetNode &d = nodes[n];
    for (int q = 0; q < d.nkids; q++) {
      dep[kids[d.kid0 + q]] = dep[n] + 1;
      max_depth = std::max(max_depth, dep[n] + 1);
    }
  }
  if (max_depth > 60) return false;
  std::vector<std::pair<uint32_t, size_t>> order;
  for (size_t n = 0; n < nodes.size(); n++)
    if (!gone[n]) order.push_back({nodes[n].own_lo, n});
  std::sort(order.begin(), order.end());
  bounds.clear();
  depth.clear();
  uint32_t expect = 0;
  for (const auto &o : order) {
    const DetNode &d = nodes[o.second];
    if (d.own_lo != expect) return false;   // (cannot happen: the blocks tile [0, N))
    bounds.push_back(d.own_lo);
    depth.push_back(dep[o.second]);
    expect = d.hi;
  }
  if (expect != (uint32_t)N) return false;
  bounds.push_back((uint32_t)N);
  return true;
}
}  // namespace

static int setup_factor_impl(rcg_handle *h, uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val,
                             const std::vector<uint32_t> &bounds, const TreeInfo &tree, bool strict_tree);

// Host only (no GPU needed): the row-length histogram of a CSR matrix as the SpMV plan sees it -- entries5[k] = entries in rows
// of length <= 2, 3..5, 6..12, 13..24, > 24 -- and the lanes per row chosen from it.
extern "C" int rcg_spmv_row_histogram(uint64_t N, const uint64_t *rowPtr, uint64_t *entries5, int *lanes) {
  if (!rowPtr || !lanes || N == 0) return RCG_ERR_INVALID;
  *lanes = spmv_lanes_from_rows(N, rowPtr, entries5);
  return RCG_OK;
}

extern "C" int rcg_detect_blocks(uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, uint64_t *bounds_out,
                                 int32_t *depth_out, uint64_t cap, uint64_t *nblocks) {
  if (!rowPtr || !colIdx || !nblocks) return RCG_ERR_INVALID;
  *nblocks = 0;
  std::vector<uint32_t> db;
  std::vector<int> dd;
  if (N < 64 || N >= 0xFFFFFFFFull || !detect_blocks(N, rowPtr, colIdx, db, dd) || db.size() <= 2) return RCG_OK;
  if (dd.size() > cap || !bounds_out || !depth_out) return RCG_ERR_INVALID;
  for (size_t i = 0; i < db.size(); i++) bounds_out[i] = db[i];
  for (size_t i = 0; i < dd.size(); i++) depth_out[i] = dd[i];
  *nblocks = dd.size();
  return RCG_OK;
}

int rcg_setup_factor(rcg_handle *h, uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val,
                     const uint64_t *part, uint64_t npart) {
  if (h->haveA && N != h->N) {
    h->err = "factor dimension differs from the matrix's";
    return RCG_ERR_INVALID;
  }
  if (!(part && npart >= 2) && N >= 64 && N < 0xFFFFFFFFull && rowPtr && colIdx && !h->opt.chain_generic) {
    // stock signature: recover the blocks from G (a factor that is one chain comes back as one block)
    std::vector<uint32_t> db;
    std::vector<int> dd;
    const double t0 = wall_ms();
    if (detect_blocks(N, rowPtr, colIdx, db, dd) && db.size() > 2) {
      TreeInfo tree;
      tree.depth = dd;
      tree.sub_lo.assign(dd.size(), 0);
      for (int v : dd) tree.max_depth = std::max(tree.max_depth, v);
      h->stats.reserved[7] = wall_ms() - t0;   // host time of the block detection (ms)
      const int rc = setup_factor_impl(h, N, rowPtr, colIdx, val, db, tree, false);
      if (rc != RCG_ERR_STRUCTURE) return rc;
      h->err.clear();   // the detected structure did not validate: solve as one block
    }
  }
  // ---- partition -> block boundaries ---------------------------------------------------------------
  std::vector<uint32_t> bounds;
  if (part && npart >= 2) {
    uint64_t nb = npart - 1;
    if (((nb + 1) & nb) != 0 || part[0] != 0 || part[npart - 1] != N) {
      h->err = "part must hold 2T boundaries (T a power of two) from 0 to N";
      return RCG_ERR_INVALID;
    }
    for (uint64_t i = 0; i < npart; i++) {
      if (i && part[i] < part[i - 1]) { h->err = "part is not monotone"; return RCG_ERR_INVALID; }
      bounds.push_back((uint32_t)part[i]);
    }
  } else {
    bounds = {0u, (uint32_t)N};
  }
  const int nb = (int)bounds.size() - 1;
  TreeInfo tree;
  tree.depth.assign(nb, 0);
  tree.sub_lo.assign(nb, 0);
  build_tree(bounds, 0, nb, 0, tree);
  return setup_factor_impl(h, N, rowPtr, colIdx, val, bounds, tree, true);
}

// General form: `nblocks` consecutive blocks with boundaries bounds[0..nblocks] and a depth per block; a row of a block
// may only couple to its own block and to blocks of SMALLER depth (solved later in the forward solve, earlier in the
// backward solve).  Used by the multi-GPU layout, whose local index space is [own subtree | replicated top separators].
int rcg_setup_factor_blocks(rcg_handle *h, uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val,
                            const uint64_t *bounds64, const int32_t *depth, uint64_t nblocks) {
  if (h->haveA && N != h->N) {
    h->err = "factor dimension differs from the matrix's";
    return RCG_ERR_INVALID;
  }
  if (!bounds64 || !depth || nblocks == 0 || bounds64[0] != 0 || bounds64[nblocks] != N) {
    h->err = "blocks must cover [0, N)";
    return RCG_ERR_INVALID;
  }
  std::vector<uint32_t> bounds;
  TreeInfo tree;
  for (uint64_t i = 0; i <= nblocks; i++) {
    if (i && bounds64[i] < bounds64[i - 1]) { h->err = "block boundaries are not monotone"; return RCG_ERR_INVALID; }
    bounds.push_back((uint32_t)bounds64[i]);
  }
  for (uint64_t i = 0; i < nblocks; i++) {
    if (depth[i] < 0 || depth[i] > 62) { h->err = "block depth out of range"; return RCG_ERR_INVALID; }
    tree.depth.push_back(depth[i]);
    tree.sub_lo.push_back(0);
    tree.max_depth = std::max(tree.max_depth, (int)depth[i]);
  }
  return setup_factor_impl(h, N, rowPtr, colIdx, val, bounds, tree, false);
}

static int setup_factor_impl(rcg_handle *h, uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val,
                             const std::vector<uint32_t> &bounds, const TreeInfo &tree, bool strict_tree) {
  if (h->haveG) { rcg_free_direction(h->fwd); rcg_free_direction(h->bwd); h->haveG = false; }
  const int nb = (int)bounds.size() - 1;
  CsrDev U, L, R;
  // whatever of U, L, R is still allocated when the function returns (an error path) is released; the success path has
  // handed L and R to the layouts and freed U (rcg_free_csr leaves null pointers behind, so this is a no-op then)
  struct Scope { CsrDev *m[3]; ~Scope() { for (CsrDev *c : m) rcg_free_csr(*c); } } scope{{&U, &L, &R}};
  RcgPhases ph;
  RCG_TRY(upload_csr(h, N, rowPtr, colIdx, val, U));
  ph.mark(h->stream, "set_factor: upload of G");
  h->N = N;
  const int64_t nnz = U.nnz;
  const uint32_t n32 = (uint32_t)N;
  double t0 = wall_ms();

  int *derr = nullptr;
  RCG_CUDA(h, cudaMalloc(&derr, sizeof(int)));
  RCG_CUDA(h, cudaMemsetAsync(derr, 0, sizeof(int), h->stream));
  k_validate_upper<<<grid_for(h, n32, 256), 256, 0, h->stream>>>(U.rowptr, U.col, U.val, n32, derr);
  h->stats.kernel_launches += 1;
  int herr = 0;
  RCG_CUDA(h, cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  if (herr) {
    rcg_free_csr(U); cudaFree(derr);
    h->err = "G must be CSR of an upper-triangular matrix with sorted rows and a positive leading diagonal";
    return RCG_ERR_STRUCTURE;
  }
  if (nb > 1) {
    uint32_t *dpart = nullptr, *dsub = nullptr;
    RCG_CUDA(h, cudaMalloc(&dpart, sizeof(uint32_t) * (nb + 1)));
    RCG_CUDA(h, cudaMalloc(&dsub, sizeof(uint32_t) * nb));
    RCG_CUDA(h, cudaMemcpyAsync(dpart, bounds.data(), sizeof(uint32_t) * (nb + 1), cudaMemcpyHostToDevice, h->stream));
    RCG_CUDA(h, cudaMemcpyAsync(dsub, tree.sub_lo.data(), sizeof(uint32_t) * nb, cudaMemcpyHostToDevice, h->stream));
    if (!strict_tree) {   // depth-based rule: reuse the sub_lo slot for the depths
      std::vector<uint32_t> dep(tree.depth.begin(), tree.depth.end());
      RCG_CUDA(h, cudaMemcpyAsync(dsub, dep.data(), sizeof(uint32_t) * nb, cudaMemcpyHostToDevice, h->stream));
      k_validate_depths<<<grid_for(h, n32, 256), 256, 0, h->stream>>>(U.rowptr, U.col, n32, dpart, dsub, nb, derr);
    } else
    k_validate_blocks<<<grid_for(h, n32, 256), 256, 0, h->stream>>>(U.rowptr, U.col, n32, dpart, dsub, nb, derr);
    h->stats.kernel_launches += 1;
    RCG_CUDA(h, cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    RCG_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(dpart); cudaFree(dsub);
    if (herr) {
      rcg_free_csr(U); cudaFree(derr);
      h->err = "G couples a block to a non-ancestor block: `part` does not describe this factor";
      return RCG_ERR_STRUCTURE;
    }
  }
  RCG_CUDA(h, cudaFree(derr));
  ph.mark(h->stream, "set_factor: validation");

  // ---- forward direction: L = U^T ---------------------------------------------------------------------
  L.nnz = nnz;
  RCG_CUDA(h, cudaMalloc(&L.rowptr, sizeof(int64_t) * (N + 4)));   // padded: staged in 16-byte aligned slices
  RCG_CUDA(h, cudaMalloc(&L.col, sizeof(uint32_t) * (size_t)nnz));
  RCG_CUDA(h, cudaMalloc(&L.val, sizeof(double) * (size_t)nnz));
  RCG_CUDA(h, cudaMemsetAsync(L.rowptr, 0, sizeof(int64_t) * (N + 4), h->stream));
  k_count_cols<<<grid_for(h, nnz, 256), 256, 0, h->stream>>>(U.col, nnz, L.rowptr);
  h->stats.kernel_launches += 1;
  RCG_TRY(exclusive_scan_inplace(h, L.rowptr, (int64_t)N + 1));
  {
    int64_t *cursor = nullptr;
    RCG_CUDA(h, cudaMalloc(&cursor, sizeof(int64_t) * N));
    RCG_CUDA(h, cudaMemcpyAsync(cursor, L.rowptr, sizeof(int64_t) * N, cudaMemcpyDeviceToDevice, h->stream));
    k_scatter_transpose<<<grid_for(h, (int64_t)N * 8, 256), 256, 0, h->stream>>>(U.rowptr, U.col, U.val, n32, cursor,
                                                                               L.col, L.val);
    h->stats.kernel_launches += 1;
    RCG_TRY(sort_segments(h, L.rowptr, L.col, L.val, n32));
    RCG_CUDA(h, cudaFree(cursor));
  }

  ph.mark(h->stream, "set_factor: transpose + row sort (L)");
  // ---- backward direction: reversed U ------------------------------------------------------------------
  R.nnz = nnz;
  RCG_CUDA(h, cudaMalloc(&R.rowptr, sizeof(int64_t) * (N + 4)));
  RCG_CUDA(h, cudaMalloc(&R.col, sizeof(uint32_t) * (size_t)nnz));
  RCG_CUDA(h, cudaMalloc(&R.val, sizeof(double) * (size_t)nnz));
  k_reverse_entries<<<grid_for(h, nnz, 256), 256, 0, h->stream>>>(U.col, U.val, nnz, n32, R.col, R.val);
  k_reverse_rowptr<<<grid_for(h, (int64_t)N + 1, 256), 256, 0, h->stream>>>(U.rowptr, nnz, n32, R.rowptr);
  h->stats.kernel_launches += 2;
  RCG_CUDA(h, cudaGetLastError());
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  rcg_free_csr(U);
  h->bwd.reversed = true;
  h->fwd.reversed = false;

  ph.mark(h->stream, "set_factor: reversed copy (R), free U");
  // ---- schedules ---------------------------------------------------------------------------------------
  RCG_TRY(finish_direction(h, h->fwd, L, bounds, tree.depth, tree.max_depth, /*root_first=*/false));
  ph.mark(h->stream, "set_factor: forward layout");
  std::vector<uint32_t> rbounds(nb + 1);
  std::vector<int> rdepth(nb);
  for (int i = 0; i <= nb; i++) rbounds[i] = n32 - bounds[nb - i];
  for (int b = 0; b < nb; b++) rdepth[b] = tree.depth[nb - 1 - b];
  RCG_TRY(finish_direction(h, h->bwd, R, rbounds, rdepth, tree.max_depth, /*root_first=*/true));
  ph.mark(h->stream, "set_factor: backward layout");

  h->stats.analysis_ms += wall_ms() - t0;
  h->haveG = true;
  h->nnzG = (uint64_t)nnz;
  h->n_blocks = nb;
  h->tree_levels = tree.max_depth + 1;
  h->stats.nnzG = (uint64_t)nnz;
  h->stats.n_blocks = (uint64_t)nb;
  h->stats.tree_levels = (uint64_t)h->tree_levels;
  return RCG_OK;
}
