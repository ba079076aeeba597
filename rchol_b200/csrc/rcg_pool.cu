// rchol_b200 -- device memory through the stream-ordered pool (see rcg_common.cuh "device memory").
#define RCG_POOL_IMPL
#include <cstdlib>
#include <mutex>

#include "rcg_common.cuh"

namespace {
struct PoolDev {
  bool init = false, on = false;
  cudaStream_t stream = nullptr;
};
PoolDev g_pool[64];
std::mutex g_mu;

PoolDev *pool_of_current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(g_mu);
  PoolDev &P = g_pool[dev];
  if (!P.init) {
    P.init = true;
    const char *e = getenv("RCG_POOL");
    int supported = 0;
    cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, dev);
    if (supported && !(e && atoi(e) == 0)) {
      cudaMemPool_t mp = nullptr;
      unsigned long long keep = ~0ull;
      if (cudaDeviceGetDefaultMemPool(&mp, dev) == cudaSuccess &&
          cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess &&
          cudaStreamCreateWithFlags(&P.stream, cudaStreamNonBlocking) == cudaSuccess)
        P.on = true;
    }
    cudaGetLastError();
  }
  return &P;
}
}  // namespace

// Same contract as cudaMalloc: the memory is usable on any stream when the call returns (the pool's own stream is idle
// apart from these operations, so the synchronisation is free).
cudaError_t rcg_pool_malloc(void **p, size_t bytes) {
  PoolDev *P = pool_of_current_device();
  if (!P || !P->on) return cudaMalloc(p, bytes);
  cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 1, P->stream);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(P->stream);
}

// Same contract as cudaFree: all work of the device that may still use the memory is complete before it is reused.
cudaError_t rcg_pool_free(void *p) {
  if (!p) return cudaSuccess;
  PoolDev *P = pool_of_current_device();
  if (!P || !P->on) return cudaFree(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return e;
  return cudaFreeAsync(p, P->stream);
}
