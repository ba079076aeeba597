// Microbenchmark: one warp advancing a dependency chain level by level through shared memory (no polling):
// every step each lane loads two values written in the previous step by other lanes, does two DFMAs and stores.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int VAR>
__global__ void k_crit(unsigned long long *out, int steps, const uint32_t *lvl) {
  __shared__ double win[2048];
  __shared__ double part[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) { win[i] = 1.0 + i * 1e-9; part[i] = 0.5; }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  volatile double *vw = win;
  double sink = 0;
  long long t0 = clock64();
  uint32_t pos = 32;
  for (int s = 0; s < steps; s++) {
    // batch = up to `nb` rows starting at pos (VAR 0: fixed 4 rows; VAR 1: batch found with a ballot on staged levels)
    uint32_t nb = 4;
    if (VAR == 1) {
      const uint32_t l0 = lvl[(pos) & 2047], mine = lvl[(pos + lane) & 2047];
      const unsigned m = __ballot_sync(0xffffffffu, mine == l0);
      nb = __ffs(~m) - 1;   // length of the leading run
      if (nb == 0 || nb > 32) nb = 32;
    }
    const bool act = lane < nb;
    const uint32_t r = (pos + lane) & 2047;
    // two "near" dependencies: rows produced in the previous step(s)
    const double x0 = vw[(pos - 1 - lane) & 2047];
    const double x1 = vw[(pos - 3 - lane) & 2047];
    double acc = part[r];
    acc = fma(-0.25, x0, acc);
    acc = fma(-0.125, x1, acc);
    if (act) vw[r] = acc;
    __syncwarp();
    sink += acc;
    pos += nb;
  }
  long long t1 = clock64();
  if (lane == 0) { out[0] = t1 - t0; out[1] = (unsigned long long)sink; }
}
int main() {
  unsigned long long *d, h[2];
  cudaMalloc(&d, 16);
  uint32_t hl[4096];
  for (int i = 0; i < 4096; i++) hl[i] = i / 4;
  uint32_t *dl; cudaMalloc(&dl, sizeof(hl)); cudaMemcpy(dl, hl, sizeof(hl), cudaMemcpyHostToDevice);
  const int T = 20000;
  k_crit<0><<<1, 32>>>(d, T, dl); cudaDeviceSynchronize(); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("critical warp, fixed batches   : %.1f cycles/level\n", (double)h[0] / T);
  k_crit<1><<<1, 32>>>(d, T, dl); cudaDeviceSynchronize(); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("critical warp, ballot batches  : %.1f cycles/level\n", (double)h[0] / T);
  k_crit<1><<<1, 288>>>(d, T, dl); cudaDeviceSynchronize(); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("critical warp (+8 idle warps)  : %.1f cycles/level\n", (double)h[0] / T);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
