// Microbenchmark for the blocked chain's critical warp (rcg_blocked.cu): what does ONE warp pay per chunk for
//   [recent: 6 LDS.128 + 8 gathers + 8 DFMA] -> STS t, syncwarp -> [mat-vec: 16 LDS.128 W + 16 LDS.128 bcast + 32 DFMA]
//   -> STS x, syncwarp -> release (MEMBAR.CTA + STS | mbarrier arrive | plain STS)
// with the other warps idle, sleeping-polling, or gathering from shared memory all the time?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mv mv.cu ; run: ./mv
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lds128d(double &a, double &b, uint32_t addr) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr) : "memory");
}
__device__ __forceinline__ void lds128u(uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d, uint32_t addr) {
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
}
__device__ __forceinline__ double lds64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }

// PHASES bit 0: recent, bit 1: mat-vec, bit 2: store+release ; REL 0: st.release.cta, 1: mbarrier.arrive, 2: plain volatile store
// NOISE 0: other warps exit, 1: poll a flag with nanosleep(300), 2: poll without sleeping, 3: gather loop
template <int PHASES, int REL, int NOISE, int WPRE>
__global__ void __launch_bounds__(512, 1) k_mv(unsigned long long *out, int iters) {
  extern __shared__ __align__(128) unsigned char sm[];
  double *win = reinterpret_cast<double *>(sm);                 // 4096 doubles
  double *scratch = win + 4096 + 16;                            // 32
  unsigned char *ring = reinterpret_cast<unsigned char *>(scratch + 32);   // 4 slots x 12288 B
  uint64_t *bar = reinterpret_cast<uint64_t *>(ring + 4 * 12288);
  uint32_t *flag = reinterpret_cast<uint32_t *>(bar + 8);
  for (int i = threadIdx.x; i < 4096 + 16; i += blockDim.x) win[i] = 1.0 + 1e-9 * i;
  for (int i = threadIdx.x; i < 4 * 12288 / 8; i += blockDim.x) reinterpret_cast<double *>(ring)[i] = 1e-3;
  for (int s = 0; s < 4; s++)   // window byte offsets of the "recent" batch
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
      reinterpret_cast<uint32_t *>(ring + s * 12288 + 8192 + 2048)[i] = 8u * ((i * 37u + s * 11u) & 4095u);
  if (threadIdx.x == 0) {
    *flag = 0;
    for (int i = 0; i < 8; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + i)));
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t win_s = smem_u32(win), sc_s = smem_u32(scratch), flag_s = smem_u32(flag);
  if (warp != 0) {
    if (NOISE == 0 || (warp & 3) == 0) return;
    double acc = 0;
    uint32_t o = lane * 8;
    for (;;) {
      uint32_t f;
      asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(f) : "r"(flag_s) : "memory");
      if (f == 0xFFFFFFFFu) break;
      if (NOISE == 1) __nanosleep(300);
      if (NOISE == 3) {
#pragma unroll
        for (int u = 0; u < 8; u++) { acc += lds64(win_s + o); o = (o * 5u + 8u * u + 8u) & 32760u; }
      }
    }
    if (acc == 123.456) out[1] = 1;
    return;
  }
  double x = 1.0;
  const long long t0 = clock64();
  for (int k = 0; k < iters; k++) {
    const uint32_t a_s = smem_u32(ring + (k & 3) * 12288);
    const uint32_t w_s = a_s + 16u * lane, r_s = w_s + 8192u;
    double w[2 * (WPRE > 0 ? WPRE : 1)];
#pragma unroll
    for (int p = 0; p < WPRE; p++) lds128d(w[2 * p], w[2 * p + 1], w_s + 512u * p);
    double t = x;
    if (PHASES & 1) {
      uint32_t o0, o1, o2, o3, o4, o5, o6, o7;
      double v0, v1, v2, v3, v4, v5, v6, v7;
      lds128u(o0, o1, o2, o3, r_s + 2048u);
      lds128u(o4, o5, o6, o7, r_s + 2560u);
      lds128d(v0, v1, r_s);
      lds128d(v2, v3, r_s + 512u);
      lds128d(v4, v5, r_s + 1024u);
      lds128d(v6, v7, r_s + 1536u);
      const double x0 = lds64(win_s + o0), x1 = lds64(win_s + o1), x2 = lds64(win_s + o2), x3 = lds64(win_s + o3);
      const double x4 = lds64(win_s + o4), x5 = lds64(win_s + o5), x6 = lds64(win_s + o6), x7 = lds64(win_s + o7);
      double t1, t2, t3;
      t = fma(-v0, x0, t); t1 = -v1 * x1; t2 = -v2 * x2; t3 = -v3 * x3;
      t = fma(-v4, x4, t); t1 = fma(-v5, x5, t1); t2 = fma(-v6, x6, t2); t3 = fma(-v7, x7, t3);
      t = (t + t1) + (t2 + t3);
    }
    if (PHASES & 2) {
      sts64(sc_s + 8u * lane, t);
      __syncwarp();
      double a0 = 0, a1 = 0, a2 = 0, a3 = 0, x0, x1, y0, y1;
#pragma unroll
      for (int p = 0; p < WPRE; p += 2) {
        lds128d(x0, x1, sc_s + 16u * p);
        a0 = fma(w[2 * p], x0, a0); a1 = fma(w[2 * p + 1], x1, a1);
        lds128d(x0, x1, sc_s + 16u * p + 16u);
        a2 = fma(w[2 * p + 2], x0, a2); a3 = fma(w[2 * p + 3], x1, a3);
      }
#pragma unroll
      for (int p = WPRE; p < 16; p += 2) {
        lds128d(y0, y1, w_s + 512u * p);
        lds128d(x0, x1, sc_s + 16u * p);
        a0 = fma(y0, x0, a0); a1 = fma(y1, x1, a1);
        lds128d(y0, y1, w_s + 512u * p + 512u);
        lds128d(x0, x1, sc_s + 16u * p + 16u);
        a2 = fma(y0, x0, a2); a3 = fma(y1, x1, a3);
      }
      t = (a0 + a1) + (a2 + a3);
    }
    x = t * 1e-3 + 1.0;
    if (PHASES & 4) {
      sts64(win_s + 8u * ((32u * k + lane) & 4095u), x);
      __syncwarp();
      if (lane == 0) {
        if (REL == 0) asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(flag_s), "r"(k + 1) : "memory");
        if (REL == 1) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar + (k & 7))) : "memory");
        if (REL == 2) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(flag_s), "r"(k + 1) : "memory");
      }
    }
  }
  const long long t1 = clock64();
  if (lane == 0) {
    out[0] = (unsigned long long)(t1 - t0);
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(flag_s), "r"(0xFFFFFFFFu) : "memory");
  }
  if (x == 123.456) out[1] = 2;
}

template <int PHASES, int REL, int NOISE, int WPRE>
void run(const char *name) {
  unsigned long long *d, h[2];
  cudaMalloc(&d, 16);
  const int iters = 20000, smem = (4096 + 16 + 32) * 8 + 4 * 12288 + 128;
  cudaFuncSetAttribute(k_mv<PHASES, REL, NOISE, WPRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_mv<PHASES, REL, NOISE, WPRE><<<1, 512, smem>>>(d, iters);
  k_mv<PHASES, REL, NOISE, WPRE><<<1, 512, smem>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-58s %7.1f cycles/chunk  %s\n", name, (double)h[0] / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<2, 0, 0, 8>("matvec only, W half preloaded, idle");
  run<2, 0, 0, 0>("matvec only, W in loop, idle");
  run<2, 0, 0, 16>("matvec only, W fully preloaded, idle");
  run<1, 0, 0, 8>("recent only, idle");
  run<3, 0, 0, 8>("recent + matvec, idle");
  run<7, 0, 0, 8>("recent + matvec + st.release, idle");
  run<7, 1, 0, 8>("recent + matvec + mbarrier.arrive, idle");
  run<7, 2, 0, 8>("recent + matvec + plain store, idle");
  run<7, 0, 1, 8>("all + st.release, 12 warps poll with nanosleep(300)");
  run<7, 0, 2, 8>("all + st.release, 12 warps poll without sleeping");
  run<7, 0, 3, 8>("all + st.release, 12 warps gather from smem all the time");
  run<7, 1, 3, 8>("all + mbarrier.arrive, 12 warps gather all the time");
  return 0;
}
