// rchol_b200 -- device-side helpers shared by the kernel translation units (sm_100a).
#pragma once
#include "rcg_common.cuh"

namespace {

constexpr unsigned long long SENTINEL = 0xFFFFFFFFFFFFFFFFull;   // "not solved yet" marker in the window
constexpr unsigned long long CANON_NAN = 0x7FF8000000000000ull;
constexpr uint32_t WATCHDOG_TRIPS = 1u << 21;   // polling trips before a row gives up (~0.1 s)

// ---------------------------------------------------------------------------------------------------------
// reductions: every CTA publishes a partial; the last CTA to finish sums them in index order (deterministic)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// returns the CTA total in every thread of warp 0 (valid in thread 0)
__device__ __forceinline__ double block_sum(double v, double *sm /* >= 32 doubles */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = lane < nw ? sm[lane] : 0.0;
    t = warp_sum(t);
  }
  return t;
}

// Sum of `n` partials in a fixed order, identical in every CTA that calls it.
__device__ __forceinline__ double sum_partials(const double *__restrict__ part, int n, double *sm) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
  double t = block_sum(s, sm);
  __shared__ double bcast;
  if (threadIdx.x == 0) bcast = t;
  __syncthreads();
  return bcast;
}

// publish this CTA's partials; the last CTA reduces all of them into out[0..K)
template <int K>
__device__ __forceinline__ void publish_and_finalize(const double (&v)[K], double *partials /*K x stride*/, int stride,
                                                     unsigned int *counter, double *const (&out)[K], double *sm) {
  __shared__ bool is_last;
  double t[K];
#pragma unroll
  for (int k = 0; k < K; k++) t[k] = block_sum(v[k], sm);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; k++) partials[k * stride + blockIdx.x] = t[k];
    __threadfence();
    unsigned int done = atomicAdd(counter, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
#pragma unroll
    for (int k = 0; k < K; k++) {
      double s = 0.0;
      for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += __ldcg(&partials[k * stride + i]);
      double tot = block_sum(s, sm);
      if (threadIdx.x == 0) *out[k] = tot;
    }
    if (threadIdx.x == 0) *counter = 0u;
  }
}

// ---------------------------------------------------------------------------------------------------------
// SpMV: LPR lanes per row, consecutive rows in consecutive lane groups so that the column/value streams of a

// ---------------------------------------------------------------------------------------------------------
// mbarrier / bulk-copy (TMA) primitives -- sm_90+ PTX, SASS: SYNCS.* / UBLKCP
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
// shared-memory accesses by 32-bit shared-space address (keeps address arithmetic out of the polling loop: with
// generic pointers the compiler rebuilds the cluster-window address from SR_CgaCtaId on every trip)
__device__ __forceinline__ unsigned long long lds_volatile_u64(uint32_t addr) {
  unsigned long long v;
  asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_volatile_u64(uint32_t addr, unsigned long long v) {
  asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}

// predicated shared loads: lanes whose predicate is false keep the old register and generate no bank traffic
__device__ __forceinline__ void lds_u32_if(uint32_t &v, uint32_t addr, bool p) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.shared.u32 %0, [%1];\n\t}" : "+r"(v) : "r"(addr), "r"((uint32_t)p) : "memory");
}
__device__ __forceinline__ void lds_f64_if(double &v, uint32_t addr, bool p) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q ld.shared.f64 %0, [%1];\n\t}" : "+d"(v) : "r"(addr), "r"((uint32_t)p) : "memory");
}
// predicated stores (no branch, so the polling loop stays convergent: measured on B200, a divergent trip with
// BSSY/BSYNC/YIELD costs ~160-410 cycles against ~45 for a convergent one -- scripts/ubench/spin.cu)
__device__ __forceinline__ void sts_volatile_u64_if(uint32_t addr, unsigned long long v, bool p) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.volatile.shared.u64 [%0], %1;\n\t}" ::"r"(addr), "l"(v),
               "r"((uint32_t)p)
               : "memory");
}
__device__ __forceinline__ void stg_f64_if(double *ptr, double v, bool p) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.f64 [%0], %1;\n\t}" ::"l"(ptr), "d"(v),
               "r"((uint32_t)p)
               : "memory");
}

// global -> shared bulk copy, completion counted in bytes on `bar`; src/dst 16-byte aligned, bytes % 16 == 0
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}


__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

}  // namespace
