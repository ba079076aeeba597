// rchol_b200 -- device memory through a caching allocator (see rcg_common.cuh "device memory").
#define RCG_POOL_IMPL
#include <cstdlib>
#include <map>
#include <mutex>
#include <unordered_map>

#include "rcg_common.cuh"

namespace {
// Freed blocks are kept per device, keyed by size, and handed out again to a request of (almost) the same size -- the
// reference's one-shot `pcg(...)` constructor repeats exactly the same allocations for every solve of a time-stepping or
// many-right-hand-side loop.  Measured at 256^3 / T = 4096 (gpurun_out/c41_bench.err): the second one-shot solve spent
// 1400 ms in cudaMalloc / cudaMallocAsync calls (the driver re-mapping physical memory) that cost 35 ms in the first.
struct Cache {
  std::multimap<size_t, void *> free_blocks;          // size -> block
  std::unordered_map<void *, size_t> live;            // blocks handed out
  size_t cached_bytes = 0;
};
Cache g_cache[64];
std::mutex g_mu;
int g_enabled = -1;
constexpr size_t CACHE_CAP = (size_t)96 << 30;        // more cached than this: everything goes back to the driver

bool enabled() {
  if (g_enabled < 0) {
    const char *e = getenv("RCG_POOL");
    g_enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  return g_enabled == 1;
}

void flush(Cache &C) {
  for (auto &kv : C.free_blocks) cudaFree(kv.second);
  C.free_blocks.clear();
  C.cached_bytes = 0;
}
}  // namespace

// Same contract as cudaMalloc.
cudaError_t rcg_pool_malloc(void **p, size_t bytes) {
  int dev = 0;
  if (!enabled() || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return cudaMalloc(p, bytes);
  if (bytes == 0) bytes = 1;
  const size_t want = (bytes + 511) & ~(size_t)511;
  std::lock_guard<std::mutex> lock(g_mu);
  Cache &C = g_cache[dev];
  auto it = C.free_blocks.lower_bound(want);
  if (it != C.free_blocks.end() && it->first <= want + want / 8 + 4096) {   // (at most 12.5 % larger than asked for)
    *p = it->second;
    C.live[*p] = it->first;
    C.cached_bytes -= it->first;
    C.free_blocks.erase(it);
    return cudaSuccess;
  }
  cudaError_t e = cudaMalloc(p, want);
  if (e != cudaSuccess && !C.free_blocks.empty()) {   // out of memory with blocks cached: give them back and retry
    cudaGetLastError();
    flush(C);
    e = cudaMalloc(p, want);
  }
  if (e == cudaSuccess) C.live[*p] = want;
  return e;
}

// Same contract as cudaFree: all work of the device that may still use the memory is complete before it is reused.
cudaError_t rcg_pool_free(void *p) {
  if (!p) return cudaSuccess;
  int dev = 0;
  if (!enabled() || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return cudaFree(p);
  cudaError_t e = cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lock(g_mu);
  Cache &C = g_cache[dev];
  auto it = C.live.find(p);
  if (it == C.live.end()) return cudaFree(p);   // (not ours: allocated on another device's cache or before the cache)
  const size_t sz = it->second;
  C.live.erase(it);
  if (C.cached_bytes + sz > CACHE_CAP) flush(C);
  C.free_blocks.emplace(sz, p);
  C.cached_bytes += sz;
  return e;
}
