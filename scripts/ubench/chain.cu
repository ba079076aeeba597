// Microbenchmarks: cost of one intra-warp dependent hop through shared memory (development tool).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define SENT 0xFFFFFFFFFFFFFFFFull
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long ldv(uint32_t a) { unsigned long long v; asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void stv(uint32_t a, unsigned long long v) { asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ void stv_if(uint32_t a, unsigned long long v, bool p) {
  asm volatile("{.reg .pred q; setp.ne.u32 q, %2, 0; @q st.volatile.shared.u64 [%0], %1;}" ::"r"(a), "l"(v), "r"((uint32_t)p) : "memory");
}
// VAR: 0 vote-exit + dfma ; 1 fixed trips + dfma ; 2 fixed trips, no dfma ; 3 fixed trips, hi-word test, predicated store
//      4 shuffle-based (no shared memory) ; 5 like 3 but two lanes per level (16 levels)
template <int VAR>
__global__ void k_chain(unsigned long long *out, int reps) {
  __shared__ unsigned long long win[64];
  uint32_t base = smem_u32(win);
  asm volatile("" : "+r"(base));
  const int lane = threadIdx.x & 31;
  long long total = 0;
  double sink = 0;
  for (int rep = 0; rep < reps; rep++) {
    if (threadIdx.x < 64) win[threadIdx.x] = SENT;
    __syncwarp();
    long long t0 = clock64();
    bool pend = true; int k = (lane == 0) ? 1 : 0; const int kd = 1;
    double acc = 1.0 + lane;
    uint32_t a0 = base + 8 * (lane ? lane - 1 : 0), mine = base + 8 * lane;
    if (VAR == 0) {
      do {
        const unsigned long long b = ldv(a0);
        const bool ready = (b != SENT) & (k < kd);
        const double nacc = fma(-0.5, __longlong_as_double((long long)b), acc);
        acc = ready ? nacc : acc; k += ready;
        const bool fin = pend & (k >= kd);
        stv_if(mine, (unsigned long long)__double_as_longlong(acc), fin);
        pend = pend & !fin;
      } while (__any_sync(0xffffffffu, pend));
    } else if (VAR == 1 || VAR == 2 || VAR == 3) {
      for (int t = 0; t < 34; t++) {
        const unsigned long long b = ldv(a0);
        const bool ready = (VAR == 3 ? ((uint32_t)(b >> 32) != 0xFFFFFFFFu) : (b != SENT)) & (k < kd);
        const double x = __longlong_as_double((long long)b);
        const double nacc = (VAR == 2) ? x : fma(-0.5, x, acc);
        acc = ready ? nacc : acc; k += ready;
        const bool fin = pend & (k >= kd);
        stv_if(mine, (unsigned long long)__double_as_longlong(acc), fin);
        pend = pend & !fin;
      }
    } else if (VAR == 4) {
      // lane l needs lane l-1's value: shuffle every trip; ready flags travel with the value
      unsigned long long mybits = (lane == 0) ? (unsigned long long)__double_as_longlong(acc) : SENT;
      for (int t = 0; t < 34; t++) {
        const unsigned long long b = __shfl_up_sync(0xffffffffu, mybits, 1);
        const bool ready = (b != SENT) & (k < kd) & (lane > 0);
        const double nacc = fma(-0.5, __longlong_as_double((long long)b), acc);
        acc = ready ? nacc : acc; k += ready;
        const bool fin = pend & (k >= kd);
        mybits = fin ? (unsigned long long)__double_as_longlong(acc) : mybits;
        pend = pend & !fin;
      }
    }
    __syncwarp();
    total += clock64() - t0;
    sink += acc;
  }
  if (threadIdx.x == 0) { out[0] = total; out[1] = (unsigned long long)sink; }
}
// throughput of pure polling trips (no dependency): N warps polling ready slots, measures cycles per trip per warp
__global__ void k_poll_only(unsigned long long *out, int trips) {
  __shared__ unsigned long long win[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) win[i] = 12345;
  __syncthreads();
  uint32_t a = smem_u32(win) + 8 * (threadIdx.x & 1023);
  asm volatile("" : "+r"(a));
  unsigned long long acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < trips; i++) { unsigned long long b = ldv(a); a = a ^ (uint32_t)((b & 1) << 3); acc += b; }   // dependent address
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = acc; }
}
int main() {
  unsigned long long *d, h[2];
  cudaMalloc(&d, 16);
  auto rep = [&](const char *name, double per) { cudaDeviceSynchronize(); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("%-52s %8.1f cycles\n", name, (double)h[0] / per); };
  const int R = 2000;
  k_chain<0><<<1, 32>>>(d, R); rep("chain: vote exit + dfma / hop", R * 31.0);
  k_chain<1><<<1, 32>>>(d, R); rep("chain: fixed trips + dfma / trip (34 trips)", R * 34.0);
  k_chain<2><<<1, 32>>>(d, R); rep("chain: fixed trips, no dfma / trip", R * 34.0);
  k_chain<3><<<1, 32>>>(d, R); rep("chain: fixed trips, hi-word test / trip", R * 34.0);
  k_chain<4><<<1, 32>>>(d, R); rep("chain: shuffle based / trip", R * 34.0);
  k_poll_only<<<1, 32>>>(d, 20000); rep("dependent-address poll / trip (1 warp)", 20000.0);
  k_poll_only<<<1, 256>>>(d, 20000); rep("dependent-address poll / trip (8 warps)", 20000.0);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
