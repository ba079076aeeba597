// Microbenchmarks for the sync-free polling loop (development tool, not product code).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define SENT 0xFFFFFFFFFFFFFFFFull
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long ldv(uint32_t a) { unsigned long long v; asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void stv(uint32_t a, unsigned long long v) { asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }

// 1. LDS latency (dependent pointer chase through shared memory)
__global__ void k_lds_chase(unsigned long long *out, int trips) {
  __shared__ uint32_t buf[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = (i * 17 + 1) & 1023;
  __syncthreads();
  uint32_t p = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < trips; i++) p = buf[p];
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = p; }
}
// 2. dependent DFMA chain
__global__ void k_dfma_chain(unsigned long long *out, int trips, double a, double b) {
  double x = a;
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < trips; i++) x = fma(x, b, a);
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = (unsigned long long)x; }
}
// 3. pure spin trips: poll a slot that never becomes ready, several loop shapes
template <int SHAPE>
__global__ void k_spin(unsigned long long *out, int trips) {
  __shared__ unsigned long long win[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) win[i] = SENT;
  __syncthreads();
  uint32_t a = smem_u32(win) + 8 * (threadIdx.x & 255);
  asm volatile("" : "+r"(a));
  volatile unsigned long long *vw = win;
  unsigned long long acc = 0;
  int n = 0;
  long long t0 = clock64();
  if (SHAPE == 0) {            // branch-free
    for (int i = 0; i < trips; i++) acc += (ldv(a) != SENT);
  } else if (SHAPE == 1) {     // branch on the polled value (never taken)
    for (int i = 0; i < trips; i++) { unsigned long long b = ldv(a); if (b != SENT) { acc += b; n++; } }
  } else if (SHAPE == 2) {     // the product loop's shape: while(pend) { if (k<kd) {poll; if ready {...}} if (k>=kd) {store; pend=false} }
    bool pend = true; int k = 0, kd = 1; int t = 0;
    while (pend) {
      if (++t > trips) { k = kd; }
      if (k < kd) { unsigned long long b = ldv(a); if (b != SENT) { acc += b; k++; } }
      if (k >= kd) { stv(a, 1); pend = false; }
    }
  } else if (SHAPE == 3) {     // generic volatile pointer
    for (int i = 0; i < trips; i++) { unsigned long long b = vw[threadIdx.x & 255]; if (b != SENT) { acc += b; n++; } }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = acc + n; }
}
// 4. cross-warp hop: warp 0 lane 0 and warp 1 lane 0 ping-pong through shared memory
template <int SHAPE>
__global__ void k_pingpong(unsigned long long *out, int rounds) {
  __shared__ unsigned long long slot[64];
  if (threadIdx.x < 64) slot[threadIdx.x] = SENT;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t base = smem_u32(slot);
  asm volatile("" : "+r"(base));
  long long t0 = clock64();
  if (lane == 0 && warp < 2) {
    for (int r = 0; r < rounds; r++) {
      // warp 0 writes slot[2r%32 ...]: use alternating slots and values r
      uint32_t mine = base + 8 * (warp), theirs = base + 8 * (1 - warp);
      if (warp == 0) {
        stv(mine, (unsigned long long)r);
        while (ldv(theirs) != (unsigned long long)r) {}
      } else {
        while (ldv(theirs) != (unsigned long long)r) {}
        stv(mine, (unsigned long long)r);
      }
    }
  } else if (SHAPE == 1 && warp >= 2) {
    // background spinners: other warps polling never-ready slots, like the product kernel's waiting warps
    uint32_t a = base + 8 * (8 + (threadIdx.x & 31));
    while (ldv(base + 8 * 1) != (unsigned long long)(rounds - 1)) { if (ldv(a) != SENT) break; }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; }
}
// 5. intra-warp chain: lane l waits for lane l-1 (32-deep chain) using the product loop shape
__global__ void k_intrawarp_chain(unsigned long long *out, int reps) {
  __shared__ unsigned long long win[64];
  uint32_t base = smem_u32(win);
  asm volatile("" : "+r"(base));
  const int lane = threadIdx.x & 31;
  long long total = 0;
  for (int rep = 0; rep < reps; rep++) {
    if (threadIdx.x < 64) win[threadIdx.x] = SENT;
    __syncwarp();
    long long t0 = clock64();
    bool pend = true; int k = (lane == 0) ? 1 : 0; const int kd = 1;
    double acc = 1.0;
    uint32_t a0 = base + 8 * (lane ? lane - 1 : 0), mine = base + 8 * lane;
    while (pend) {
      if (k < kd) { unsigned long long b = ldv(a0); if (b != SENT) { acc = fma(-0.5, __longlong_as_double((long long)b), acc); k++; } }
      if (k >= kd) { double res = acc * 0.75; stv(mine, (unsigned long long)__double_as_longlong(res)); pend = false; }
    }
    __syncwarp();
    total += clock64() - t0;
  }
  if (threadIdx.x == 0) out[0] = total;
}

// 6. branch-free trip (predicated), uniform exit through a vote: intra-warp 32-lane chain
__global__ void k_chain_uniform(unsigned long long *out, int reps) {
  __shared__ unsigned long long win[64];
  uint32_t base = smem_u32(win);
  asm volatile("" : "+r"(base));
  const int lane = threadIdx.x & 31;
  long long total = 0;
  for (int rep = 0; rep < reps; rep++) {
    if (threadIdx.x < 64) win[threadIdx.x] = SENT;
    __syncwarp();
    long long t0 = clock64();
    bool pend = true; int k = (lane == 0) ? 1 : 0; const int kd = 1;
    double acc = 1.0;
    uint32_t a0 = base + 8 * (lane ? lane - 1 : 0), mine = base + 8 * lane;
    do {
      const unsigned long long b = ldv(a0);
      const bool ready = (b != SENT) & (k < kd);
      const double x = __longlong_as_double((long long)b);
      const double nacc = fma(-0.5, x, acc);
      acc = ready ? nacc : acc;
      k += ready;
      const bool fin = pend & (k >= kd);
      if (fin) { stv(mine, (unsigned long long)__double_as_longlong(acc * 0.75)); }
      pend = pend & !fin;
    } while (__any_sync(0xffffffffu, pend));
    total += clock64() - t0;
  }
  if (threadIdx.x == 0) out[0] = total;
}
// 7. cross-warp chain with all lanes uniform: warp w waits for warp w-1 (NW-deep chain per round), lane l row = l
//    every lane of warp w depends on the same lane of warp w-1; measures the cross-warp hop of the branch-free trip
__global__ void k_crosswarp_uniform(unsigned long long *out, int reps) {
  __shared__ unsigned long long win[32 * 33];
  uint32_t base = smem_u32(win);
  asm volatile("" : "+r"(base));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NW = blockDim.x >> 5;
  long long total = 0;
  for (int rep = 0; rep < reps; rep++) {
    for (int i = threadIdx.x; i < 32 * 33; i += blockDim.x) win[i] = SENT;
    __syncthreads();
    long long t0 = clock64();
    bool pend = true; int k = (warp == 0) ? 1 : 0; const int kd = 1;
    double acc = 1.0;
    uint32_t a0 = base + 8 * ((warp ? warp - 1 : 0) * 32 + lane), mine = base + 8 * (warp * 32 + lane);
    do {
      const unsigned long long b = ldv(a0);
      const bool ready = (b != SENT) & (k < kd);
      const double x = __longlong_as_double((long long)b);
      const double nacc = fma(-0.5, x, acc);
      acc = ready ? nacc : acc;
      k += ready;
      const bool fin = pend & (k >= kd);
      if (fin) { stv(mine, (unsigned long long)__double_as_longlong(acc * 0.75)); }
      pend = pend & !fin;
    } while (__any_sync(0xffffffffu, pend));
    __syncthreads();
    total += clock64() - t0;
  }
  if (threadIdx.x == 0) out[0] = total;
}
int main() {
  unsigned long long *d, h[2];
  cudaMalloc(&d, 16);
  auto rep = [&](const char *name, double per) { cudaDeviceSynchronize(); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("%-40s %8.1f cycles\n", name, (double)h[0] / per); };
  const int T = 20000;
  k_lds_chase<<<1, 32>>>(d, T); rep("LDS dependent chase / load", T);
  k_dfma_chain<<<1, 32>>>(d, T, 1.0, 0.999); rep("DFMA dependent chain / op", T);
  k_spin<0><<<1, 32>>>(d, T); rep("spin branch-free / trip (1 warp)", T);
  k_spin<1><<<1, 32>>>(d, T); rep("spin if(ready) / trip (1 warp)", T);
  k_spin<2><<<1, 32>>>(d, T); rep("spin product-shape / trip (1 warp)", T);
  k_spin<3><<<1, 32>>>(d, T); rep("spin generic volatile / trip (1 warp)", T);
  k_spin<2><<<1, 256>>>(d, T); rep("spin product-shape / trip (8 warps)", T);
  k_spin<2><<<1, 1024>>>(d, T); rep("spin product-shape / trip (32 warps)", T);
  k_pingpong<0><<<1, 64>>>(d, T); rep("ping-pong hop (2 warps) / hop", 2.0 * T);
  k_pingpong<1><<<1, 256>>>(d, T); rep("ping-pong hop (+6 spinning warps) / hop", 2.0 * T);
  k_pingpong<1><<<1, 1024>>>(d, T); rep("ping-pong hop (+30 spinning warps) / hop", 2.0 * T);
  k_intrawarp_chain<<<1, 32>>>(d, 1000); rep("intra-warp 32-lane chain / lane hop", 1000.0 * 31);
  k_chain_uniform<<<1, 32>>>(d, 1000); rep("uniform intra-warp chain / lane hop", 1000.0 * 31);
  k_crosswarp_uniform<<<1, 256>>>(d, 1000); rep("uniform cross-warp chain (8w) / hop", 1000.0 * 7);
  k_crosswarp_uniform<<<1, 1024>>>(d, 1000); rep("uniform cross-warp chain (32w) / hop", 1000.0 * 31);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
