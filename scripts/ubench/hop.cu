// Microbenchmark for DESIGN.md (g) row 1: what does ONE hop of the blocked-inverse chain cost when the dense block of a
// chunk is C x C (C = 32 ... 256 rows) and is split row-wise over a thread-block cluster of S CTAs?
//
// Per hop every CTA of the cluster holds the full input vector t (C doubles, = the x of the previous hop) in its own
// shared memory, multiplies its R x C slab (R = C / S rows) of the dense matrix with it -- NT threads, thread = (row,
// column group), partial sums reduced through shared memory -- and pushes its R results into the shared memory of
// every CTA of the cluster with st.async (data + mbarrier complete_tx in one instruction, DSMEM).  The next hop starts
// when a CTA's mbarrier has seen all C values.  The matrix slabs are either resident in shared memory (latency floor)
// or streamed from HBM by a TMA producer warp through a ring of stages (cp.async.bulk + mbarrier), distinct bytes for
// every hop, which is what the real kernel would do.
//
// The matrix is a cyclic shift (x_h[i] = x_{h-1}[(i+1) % C]), so the result after H hops is known exactly and every
// value must have crossed the cluster correctly.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o hop hop.cu && ./hop [hops] [clusters]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);           \
      exit(1);                                                                                  \
    }                                                                                           \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_cluster(uint32_t bar, uint32_t parity) {   // acquire at cluster scope (remote st.async)
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0u;
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0u;
}
// 8 bytes into the shared memory of a CTA of the cluster + complete_tx(8) on an mbarrier of the SAME remote CTA
__device__ __forceinline__ void st_async_f64(uint32_t remote_addr, double v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr),
               "l"(__double_as_longlong(v)), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void tma_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

// bounded waits: a protocol mistake must not hang the box
struct Guard {
  unsigned int *err;
  long long t0;
  uint32_t n;
};
#define WAIT(cond, code)                                                        \
  do {                                                                          \
    G.n = 0;                                                                    \
    while (!(cond)) {                                                           \
      if ((++G.n & 255u) == 0u) {                                               \
        if (*(volatile unsigned int *)G.err != 0u) break;                       \
        const long long now_ = clock64();                                       \
        if (G.n == 256u) G.t0 = now_;                                           \
        else if (now_ - G.t0 > 1500000000ll) { atomicCAS(G.err, 0u, (code)); break; } \
      }                                                                         \
    }                                                                           \
  } while (0)

template <int C, int S, int NT>
struct Cfg {
  static constexpr int R = C / S;                 // rows per CTA
  static constexpr int G = NT / R;                // column groups
  static constexpr int CG = C / G;                // columns per thread
  static constexpr int STAGE = R * C * 8;         // bytes of one slab
  static constexpr int NS_ = (200 * 1024 - 2 * C * 8 - NT * 8 - 256) / STAGE;
  static constexpr int NS = NS_ > 4 ? 4 : (NS_ < 1 ? 1 : NS_);
  static constexpr int SMEM = 256 + 2 * C * 8 + NT * 8 + NS * STAGE;
  static_assert(C % S == 0 && NT % R == 0 && C % G == 0 && G >= 1 && CG >= 1, "shape");
};

// Wg: [cluster][hop % Hw][rank][col][R] doubles.  xout: [cluster][C].  clk: [cluster] cycles of the hop loop (rank 0)
template <int C, int S, int NT, bool STREAM>
__global__ void __launch_bounds__(NT + 32, 1) k_hop(const double *__restrict__ Wg, double *__restrict__ xout,
                                                    unsigned long long *__restrict__ clk, int hops, int Hw, unsigned int *err) {
  using K = Cfg<C, S, NT>;
  constexpr int R = K::R, G_ = K::G, CG = K::CG, NS = K::NS;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem);               // [0,2) xfull, [2,2+NS) wfull, [2+NS,2+2NS) wempty
  double *xbuf = reinterpret_cast<double *>(smem + 256);             // [2][C]
  double *part = xbuf + 2 * C;                                       // [G][R]
  double *Wst = part + NT;                                           // [NS][C][R]
  const uint32_t tid = threadIdx.x, lane = tid & 31u;
  const uint32_t rank = cluster_ctarank();
  const uint32_t cluster = blockIdx.x / S;
  Guard G{err, 0, 0};

  const uint32_t xfull0 = smem_u32(&bars[0]);
  if (tid == 0) {
    mbar_init(xfull0, 1);
    mbar_init(xfull0 + 8, 1);
    for (int s = 0; s < NS; s++) {
      mbar_init(smem_u32(&bars[2 + s]), 1);
      mbar_init(smem_u32(&bars[2 + NS + s]), NT / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(xfull0, C * 8);        // hop 0 reads buffer 0
    mbar_expect_tx(xfull0 + 8, C * 8);    // hop 1 reads buffer 1
  }
  const double *Wc = Wg + ((size_t)cluster * Hw * S + rank) * (size_t)(R * C);
  if (!STREAM) {   // resident slab (stage 0)
    for (int i = tid; i < R * C; i += NT + 32) Wst[i] = Wc[i];
  }
  __syncthreads();
  cluster_sync_all();

  if (tid >= NT) {
    // ------------------------------------------------------------------------------------------ TMA producer warp
    if (STREAM && lane == 0) {
      for (int h = 0; h < hops; h++) {
        const int s = h % NS;
        const uint32_t full = smem_u32(&bars[2 + s]), empty = smem_u32(&bars[2 + NS + s]);
        WAIT(mbar_try(empty, (((uint32_t)(h / NS)) & 1u) ^ 1u), 1u);
        mbar_expect_tx(full, K::STAGE);
        const unsigned char *src = reinterpret_cast<const unsigned char *>(Wc + (size_t)(h % Hw) * S * (size_t)(R * C));
        const uint32_t dst = smem_u32(Wst + (size_t)s * R * C);
        for (int o = 0; o < K::STAGE; o += 16384) tma_load(dst + o, src + o, (uint32_t)(K::STAGE - o < 16384 ? K::STAGE - o : 16384), full);
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------ compute threads
    const uint32_t r = tid % R, g = tid / R;
    // x of "hop -1": x0[i] = 1 + i
    if (tid < (uint32_t)R) {
      const double v = 1.0 + (double)(rank * R + r);
      for (uint32_t p = 0; p < (uint32_t)S; p++)
        st_async_f64(mapa(smem_u32(&xbuf[rank * R + r]), p), v, mapa(xfull0, p));
    }
    long long t0 = 0;
    double wreg[CG];
    for (int h = 0; h < hops; h++) {
      if (h == 8 && tid == 0) t0 = clock64();
      const int s = STREAM ? h % NS : 0;
      const double *Ws = Wst + (size_t)s * R * C;
      if (STREAM) WAIT(mbar_try(smem_u32(&bars[2 + s]), ((uint32_t)(h / NS)) & 1u), 2u);
#pragma unroll
      for (int i = 0; i < CG; i++) wreg[i] = Ws[(g * CG + i) * R + r];           // before x arrives: off the critical path
      const uint32_t b = (uint32_t)h & 1u;
      WAIT(mbar_try_cluster(xfull0 + 8u * b, ((uint32_t)h >> 1) & 1u), 3u);
      if (tid == NT - 1) mbar_expect_tx(xfull0 + 8u * b, C * 8);                  // phase of hop h + 2
      const double *t = xbuf + b * C + g * CG;
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int i = 0; i < CG; i += 2) {
        a0 = fma(wreg[i], t[i], a0);
        if (i + 1 < CG) a1 = fma(wreg[i + 1], t[i + 1], a1);
      }
      part[g * R + r] = a0 + a1;
      if (STREAM) {
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[2 + NS + s]));
      }
      asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
      if (tid < (uint32_t)R) {
        double x = 0.0;
#pragma unroll
        for (int q = 0; q < G_; q++) x += part[q * R + r];
        const uint32_t nb = b ^ 1u;
        const uint32_t dst = smem_u32(&xbuf[nb * C + rank * R + r]);
#pragma unroll
        for (uint32_t p = 0; p < (uint32_t)S; p++) st_async_f64(mapa(dst, p), x, mapa(xfull0 + 8u * nb, p));
      }
    }
    // result of the last hop
    {
      const uint32_t b = (uint32_t)hops & 1u;
      WAIT(mbar_try_cluster(xfull0 + 8u * b, ((uint32_t)hops >> 1) & 1u), 4u);
      if (tid == 0) clk[cluster] = (unsigned long long)(clock64() - t0);
      if (rank == 0 && tid < (uint32_t)C) xout[(size_t)cluster * C + tid] = xbuf[b * C + tid];
      if (rank == 0 && C > NT)
        for (uint32_t i = tid + NT; i < (uint32_t)C; i += NT) xout[(size_t)cluster * C + i] = xbuf[b * C + i];
    }
  }
  __syncthreads();
  cluster_sync_all();   // nobody leaves while a peer may still write into its shared memory
}

__global__ void k_fill(double *W, size_t n, int C, int R, int S) {   // [..][rank][col][R]: cyclic shift matrix
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t in_slab = i % ((size_t)R * C);
    const int rank = (int)((i / ((size_t)R * C)) % S);
    const int col = (int)(in_slab / R), row = rank * R + (int)(in_slab % R);
    W[i] = col == (row + 1) % C ? 1.0 : 0.0;
  }
}

static double *g_W = nullptr;
static size_t g_Wbytes = 0;
static double *g_x = nullptr;
static unsigned long long *g_clk = nullptr;
static unsigned int *g_err = nullptr;

template <int C, int S, int NT, bool STREAM>
void run(int hops, int clusters) {
  using K = Cfg<C, S, NT>;
  size_t slab_all = (size_t)C * C * 8;                                   // bytes per hop per cluster
  int Hw = STREAM ? (int)(g_Wbytes / (slab_all * clusters)) : 1;
  if (Hw > hops) Hw = hops;
  if (Hw < 1) Hw = 1;
  const size_t n = (size_t)clusters * Hw * C * C;
  k_fill<<<1184, 256>>>(g_W, n, C, K::R, S);
  CK(cudaMemset(g_err, 0, 4));
  CK(cudaMemset(g_x, 0, (size_t)clusters * C * 8));
  auto kern = k_hop<C, S, NT, STREAM>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * S);
  cfg.blockDim = dim3(NT + 32);
  cfg.dynamicSmemBytes = K::SMEM;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = S; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms = 0;
  for (int rep = 0; rep < 2; rep++) {   // second run is the measured one
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, kern, (const double *)g_W, g_x, g_clk, hops, Hw, g_err));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaEventElapsedTime(&ms, e0, e1));
  }
  static double hx[8 * 256];
  static unsigned long long hclk[64];
  unsigned int herr = 0;
  CK(cudaMemcpy(hx, g_x, (size_t)clusters * C * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hclk, g_clk, clusters * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&herr, g_err, 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int c = 0; c < clusters; c++)
    for (int i = 0; i < C; i++) bad += hx[c * C + i] != 1.0 + (double)((i + hops) % C);
  double cyc = 0;
  for (int c = 0; c < clusters; c++) cyc = cyc > (double)hclk[c] ? cyc : (double)hclk[c];
  cyc /= (hops - 8);
  printf("C=%3d S=%d R=%3d NT=%3d G=%2d CG=%2d %s NS=%d clusters=%d | %7.1f cycles/hop  %6.2f cycles/row | kernel %7.3f ms  %7.1f GB/s | %s%s\n",
         C, S, K::R, NT, K::G, K::CG, STREAM ? "stream  " : "resident", STREAM ? K::NS : 0, clusters, cyc, cyc / C, ms,
         STREAM ? (double)slab_all * hops * clusters / ms / 1e6 : 0.0, bad ? "WRONG RESULT " : "ok", herr ? " TIMEOUT" : "");
  fflush(stdout);
  CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
}

int main(int argc, char **argv) {
  const int hops = argc > 1 ? atoi(argv[1]) : 4000;
  const int clusters = argc > 2 ? atoi(argv[2]) : 8;
  g_Wbytes = (size_t)2 << 30;
  CK(cudaMalloc(&g_W, g_Wbytes));
  CK(cudaMalloc(&g_x, 64 * 256 * 8));
  CK(cudaMalloc(&g_clk, 64 * 8));
  CK(cudaMalloc(&g_err, 4));
  cudaDeviceProp pr;
  CK(cudaGetDeviceProperties(&pr, 0));
  printf("%s, %d SMs, hops %d, clusters %d (one cluster = one nested-dissection leaf)\n", pr.name, pr.multiProcessorCount, hops, clusters);
  // one CTA per leaf (today's situation, chunk of 32), then bigger chunks, then clusters
  run<32, 1, 128, false>(hops, clusters);
  run<32, 1, 512, false>(hops, clusters);
  run<32, 1, 128, true>(hops, clusters);
  run<64, 1, 256, false>(hops, clusters);
  run<64, 1, 512, false>(hops, clusters);
  run<64, 1, 512, true>(hops, clusters);
  run<128, 1, 512, false>(hops, clusters);
  run<128, 1, 512, true>(hops, clusters);
  run<64, 2, 256, false>(hops, clusters);
  run<64, 2, 256, true>(hops, clusters);
  run<128, 2, 512, false>(hops, clusters);
  run<128, 2, 512, true>(hops, clusters);
  run<128, 4, 128, false>(hops, clusters);
  run<128, 4, 256, false>(hops, clusters);
  run<128, 4, 512, false>(hops, clusters);
  run<128, 4, 256, true>(hops, clusters);
  run<128, 4, 512, true>(hops, clusters);
  run<256, 4, 512, false>(hops, clusters);
  run<256, 4, 512, true>(hops, clusters);
  run<256, 8, 256, false>(hops, clusters);
  run<256, 8, 512, false>(hops, clusters);
  run<256, 8, 256, true>(hops, clusters);
  run<256, 8, 512, true>(hops, clusters);
  run<256, 8, 512, true>(hops, 1);
  run<128, 4, 512, true>(hops, 1);
  run<256, 8, 512, true>(hops, 16);
  return 0;
}
