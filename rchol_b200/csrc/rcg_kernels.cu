// rchol_b200 -- the kernels of the PCG iteration (sm_100a).  fp64 throughout; nothing here is a dense
// contraction, so no tensor-core path exists (BASELINE.json north_star); the bound is HBM bandwidth for SpMV and
// the vector updates, and the dependency chain of the factor for the triangular solves.
//
//   k_spmv<LPR>        q = A p (+ fused p.q and p.r)            replaces mkl_sparse_d_mv      pcg.cpp:130-138
//   k_tri_external     x_B = rhs_B - M[B, solved blocks] x      (throughput part of a solve)
//   k_tri_chain        block-local sync-free triangular solve   replaces mkl_sparse_d_trsv    pcg.cpp:151,155
//   k_p_update         p = z + (r.z / r_prev.z_prev) p          replaces ddot x2, dscal, daxpy pcg.cpp:89-96
//   k_xr_update        x += a p ; r -= a q ; r.r                replaces ddot x2, daxpy x2, dcopy x2, dnrm2
//                                                                                            pcg.cpp:101-108,82
#include <cstdio>

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
// warp are one contiguous, coalesced segment of the CSR arrays.
// ---------------------------------------------------------------------------------------------------------
template <int LPR, bool DOTS>
__global__ void __launch_bounds__(256) k_spmv(const int64_t *__restrict__ rowptr, const uint32_t *__restrict__ col,
                                              const double *__restrict__ val, const double *__restrict__ x,
                                              double *__restrict__ y, const double *__restrict__ r, uint32_t N,
                                              double *partials, int pstride, unsigned int *counter, PcgScalars *scal) {
  __shared__ double sm[32];
  const int sub = threadIdx.x % LPR;
  const uint32_t rows_per_cta = blockDim.x / LPR;
  double pq = 0.0, pr = 0.0;
  for (uint32_t base = blockIdx.x * rows_per_cta; base < N; base += gridDim.x * rows_per_cta) {
    const uint32_t row = base + threadIdx.x / LPR;
    double s = 0.0;
    if (row < N) {
      const int64_t e = rowptr[row + 1];
      for (int64_t k = rowptr[row] + sub; k < e; k += LPR) s = fma(val[k], __ldg(&x[col[k]]), s);
    }
#pragma unroll
    for (int o = LPR >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (row < N && sub == 0) {
      y[row] = s;
      if (DOTS) {
        const double xr = x[row];
        pq = fma(xr, s, pq);
        pr = fma(xr, r[row], pr);
      }
    }
  }
  if (DOTS) {
    double v[2] = {pq, pr};
    double *const out[2] = {&scal->pq, &scal->pr};
    publish_and_finalize<2>(v, partials, pstride, counter, out, sm);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Triangular solve, throughput part: for the rows of the group's blocks,
//   out[j] = rhs[j] - sum over external entries (columns in blocks solved by earlier groups) of M[j,c] out[c].
// LPR lanes per row, grid.y = block of the group.
// ---------------------------------------------------------------------------------------------------------
template <int LPR>
__global__ void __launch_bounds__(256) k_tri_external(const int64_t *__restrict__ rowptr, const uint32_t *__restrict__ col,
                                                      const double *__restrict__ val, const BlockDesc *__restrict__ blocks,
                                                      const double *__restrict__ rhs, double *out, uint32_t N,
                                                      int reversed) {
  const BlockDesc b = blocks[blockIdx.y];
  const int sub = threadIdx.x % LPR;
  const uint32_t rows_per_cta = blockDim.x / LPR;
  for (uint32_t base = b.lo + blockIdx.x * rows_per_cta; base < b.hi; base += gridDim.x * rows_per_cta) {
    const uint32_t j = base + threadIdx.x / LPR;
    double acc = 0.0;
    if (j < b.hi) {
      const int64_t e = rowptr[j + 1];
      for (int64_t k = rowptr[j] + sub; k < e; k += LPR) {
        const uint32_t c = col[k];
        acc = fma(val[k], out[reversed ? N - 1 - c : c], acc);
      }
    }
#pragma unroll
    for (int o = LPR >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (j < b.hi && sub == 0) {
      const uint32_t vj = reversed ? N - 1 - j : j;
      out[vj] = rhs[vj] - acc;
    }
  }
}

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

// ---------------------------------------------------------------------------------------------------------
// Triangular solve, dependency-chain part (the latency-bound heart of the path).
//
// One CTA per nested-dissection block, one thread per row, rows taken in index order; warp w owns the aligned
// 32-row staging groups w, w+NW, ... of a chunk.  The block is swept in chunks of C rows (C a power of two):
//   * solution window: 2C slots of shared memory indexed by (row - block.lo) mod 2C, i.e. the current chunk and
//     the previous one.  A slot holds the solved value; SENTINEL marks "not solved yet" (the value itself is the
//     ready flag -- sync-free).  Entries whose column is older than the previous chunk are final in HBM -> plain
//     loads, taken first because rows are sorted.
//   * matrix entries: every warp streams the contiguous CSR segment of its next 32-row group into its private
//     staging buffer with two bulk copies (TMA, cp.async.bulk + mbarrier) issued one group ahead, so the polling
//     loop touches shared memory only.  Entries beyond the staging capacity are read from HBM (rare).
//   * polling loop: divergent per-lane loop; a lane stores its result inside the loop before leaving it, so a
//     lane that waits at the reconvergence point never owes a value to a lane that is still spinning.
// init[] = right-hand side, or the output of k_tri_external when the group has external entries.
// ---------------------------------------------------------------------------------------------------------
struct ChainArgs {
  const int64_t *rowptr;
  const uint32_t *col;
  const double *val;
  const BlockDesc *blocks;
  const double *init;
  double *out;
  const double *dotvec;
  double *dot_partials;
  uint32_t N;
  int reversed;
  uint32_t C;          // chunk rows (power of two)
  uint32_t win_slots;  // 2C, or C when every block of the launch fits one chunk
  uint32_t cap;        // staging capacity per warp / slot in entries (multiple of 4)
  uint32_t slots;      // staging slots of the pipelined kernel
  unsigned long long *clk;  // nullable: {SM cycles, nanoseconds} of CTA 0 (clock-rate probe)
  uint32_t *trace;          // nullable diagnostics: per row {finish cycle, polling-loop trips, start cycle, first-ready cycle}
};

__global__ void __launch_bounds__(1024) k_tri_chain(const ChainArgs P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double red[32];
  const uint32_t NW = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long *win = reinterpret_cast<unsigned long long *>(smem_raw);
  volatile unsigned long long *vwin = win;
  double *sval_all = reinterpret_cast<double *>(smem_raw + (size_t)P.win_slots * 8);
  uint32_t *scol_all = reinterpret_cast<uint32_t *>(sval_all + (size_t)NW * P.cap);
  uint64_t *bars = reinterpret_cast<uint64_t *>(scol_all + (size_t)NW * P.cap);
  const double *sval = sval_all + (size_t)warp * P.cap;
  const uint32_t *scol = scol_all + (size_t)warp * P.cap;
  uint64_t *bar = bars + warp;

  const BlockDesc b = P.blocks[blockIdx.x];
  const uint32_t rows = b.hi - b.lo;
  const uint32_t C = P.C, mask = P.win_slots - 1;
  const uint32_t nchunks = (rows + C - 1) / C;
  const uint32_t N = P.N;
  const bool rev = P.reversed != 0;

  if (lane == 0) mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  // ---- registers prefetched one group ahead -------------------------------------------------------------
  int64_t n_rs = 0, n_re = 0;     // row start / end (entry indices) of this lane's row in the NEXT group
  double n_init = 0.0, n_dv = 0.0;
  uint32_t n_j0 = 0, n_nr = 0;    // next group: first row, number of rows (0 = none)
  uint32_t parity = 0;

  // prefetch the lane registers of group (chunk ci, group g) and issue its bulk copies
  auto prefetch_group = [&](uint32_t ci, uint32_t g) {
    const uint32_t clo = b.lo + ci * C;
    const uint32_t cn = min(C, b.hi - clo);
    n_j0 = clo + 32u * g;
    n_nr = min(32u, clo + cn - n_j0);
    if (lane < n_nr) {
      const uint32_t j = n_j0 + lane;
      const uint32_t vj = rev ? N - 1 - j : j;
      n_rs = P.rowptr[j];
      n_re = P.rowptr[j + 1];
      n_init = P.init[vj];
      n_dv = P.dotvec ? P.dotvec[vj] : 0.0;
    }
    const int64_t e0 = __shfl_sync(0xffffffffu, n_rs, 0);
    const int64_t e1 = __shfl_sync(0xffffffffu, n_re, (int)n_nr - 1);
    if (lane == 0) {
      const int64_t e0s = e0 & ~3ll;
      uint32_t n = (uint32_t)min((long long)P.cap, (long long)((e1 - e0s + 3) & ~3ll));
      mbar_expect_tx(bar, n * 12u);
      bulk_g2s(const_cast<double *>(sval), P.val + e0s, n * 8u, bar);
      bulk_g2s(const_cast<uint32_t *>(scol), P.col + e0s, n * 4u, bar);
    }
  };
  // first group of this warp: chunk 0, group `warp` (or the first chunk that has such a group)
  auto first_group_from = [&](uint32_t ci, uint32_t g) -> bool {
    // advance (ci, g) to the next existing group of this warp; g already points NW past the previous one
    while (ci < nchunks) {
      const uint32_t cn = min(C, rows - ci * C);
      const uint32_t ng = (cn + 31) >> 5;
      if (g < ng) { prefetch_group(ci, g); return true; }
      ci++;
      g = warp;
    }
    n_nr = 0;
    return false;
  };
  first_group_from(0, warp);

  double dot = 0.0;
  for (uint32_t ci = 0; ci < nchunks; ci++) {
    const uint32_t clo = b.lo + ci * C;
    const uint32_t cn = min(C, b.hi - clo);
    const uint32_t wbase = (P.win_slots == C) ? 0u : (ci & 1u) * C;
    for (uint32_t i = threadIdx.x; i < cn; i += blockDim.x) win[wbase + i] = SENTINEL;
    __syncthreads();
    // columns >= smem_lo are in the window (current or previous chunk); older ones are final in HBM
    const uint32_t smem_lo = (ci > 0 && P.win_slots != C) ? clo - C : clo;
    const uint32_t ng = (cn + 31) >> 5;
    for (uint32_t g = warp; g < ng; g += NW) {
      // ---- take over the prefetched group ----------------------------------------------------------------
      const uint32_t j0 = n_j0, nr = n_nr;
      const int64_t rs = n_rs, re = n_re;
      double acc = n_init;
      const double dv = n_dv;
      const int64_t e0s = __shfl_sync(0xffffffffu, rs, 0) & ~3ll;
      const int64_t e1 = __shfl_sync(0xffffffffu, re, (int)nr - 1);
      const uint32_t nst = (uint32_t)min((long long)P.cap, (long long)((e1 - e0s + 3) & ~3ll));
      mbar_wait(bar, parity);
      parity ^= 1u;

      bool pend = lane < nr;
      const uint32_t j = j0 + lane;
      const uint32_t vj = rev ? N - 1 - j : j;
      uint32_t kr = (uint32_t)(rs - e0s);            // relative entry index of the lane's current entry
      const uint32_t kd = (uint32_t)(re - 1 - e0s);  // relative index of the diagonal slot
      double dinv = 0.0;
      uint32_t c = 0;
      double v = 0.0;
      auto load_entry = [&](uint32_t k, uint32_t &cc, double &vv) {   // cc = absolute column (stored block-relative)
        if (k < nst) { cc = scol[k] + b.lo; vv = sval[k]; }
        else { cc = P.col[e0s + k] + b.lo; vv = P.val[e0s + k]; }
      };
      if (pend) {
        uint32_t dc;
        load_entry(kd, dc, dinv);
        acc *= dinv;   // local off-diagonal values are pre-scaled by 1/diag at set-up
        // entries older than the window: final values in HBM
        while (kr < kd) {
          load_entry(kr, c, v);
          if (c >= smem_lo) break;
          acc = fma(v, P.out[rev ? N - 1 - c : c], acc);
          ++kr;
        }
      }
      // ---- sync-free polling on the window ---------------------------------------------------------------
      while (pend) {
        if (kr < kd) {
          const unsigned long long bits = vwin[(c - b.lo) & mask];
          if (bits != SENTINEL) {
            acc = fma(v, __longlong_as_double((long long)bits), acc);
            ++kr;
            if (kr < kd) load_entry(kr, c, v);
          }
        }
        if (kr >= kd) {
          double res = acc;
          unsigned long long rb = (unsigned long long)__double_as_longlong(res);
          if (res != res) { rb = CANON_NAN; res = __longlong_as_double((long long)CANON_NAN); }
          vwin[(j - b.lo) & mask] = rb;
          P.out[vj] = res;
          dot = fma(res, dv, dot);
          pend = false;
        }
      }
      __syncwarp();
      // ---- stage this warp's next group (same chunk, or its first group of a later chunk) -------------
      first_group_from(ci, g + NW);
    }
    __syncthreads();
  }
  if (P.dot_partials) {
    double t = block_sum(dot, red);
    if (threadIdx.x == 0) P.dot_partials[blockIdx.x] = t;
  }
}

// Fast variant: every staging group of the launch fits a staging slot (cap >= max_stage), so the polling loop
// reads shared memory only.  Warp-specialised TMA pipeline:
//   * warp NW (producer) runs ahead of the solve and streams, for every 32-row group, the group's CSR segment
//     (values, columns), its 33 row pointers and its slice of the start vector (and of the dot vector) into a ring
//     of S staging slots with bulk copies (cp.async.bulk, completion on the slot's `full` mbarrier); a slot is
//     reused once the consumer of its previous group has arrived on the slot's `empty` mbarrier.  Measured on
//     B200: issuing the copies only one group ahead exposes ~8000 cycles per group and warp.
//   * warps 0..NW-1 (consumers) take the groups of a chunk round-robin and run the sync-free polling loop:
//     convergent, branch-free trips (a divergent trip costs ~160-410 cycles against ~45 for a convergent one --
//     scripts/ubench/spin.cu); two entries are kept in registers and the second one is only turned into an
//     address one trip after its loads were issued, so the only latency a trip waits for is the poll itself.
// Off-diagonal values are stored as v' = -v/diag at set-up: x_j = init_j/diag + sum v'_jc x_c, so the last
// dependent operation of a row is a single DFMA.
constexpr uint32_t SLOT_RP = 40, SLOT_VEC = 36;   // doubles reserved for the row-pointer / vector slices of a slot

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(1024) k_tri_chain_fast(const ChainArgs P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double red[32];
  const uint32_t NW = (blockDim.x >> 5) - 1, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t S = P.slots;
  const uint32_t slot_bytes = P.cap * 12u + (SLOT_RP + 2u * SLOT_VEC) * 8u;
  unsigned long long *win = reinterpret_cast<unsigned long long *>(smem_raw);
  unsigned char *slots = smem_raw + (size_t)P.win_slots * 8;
  uint64_t *full = reinterpret_cast<uint64_t *>(slots + (size_t)S * slot_bytes);
  uint64_t *empty = full + S;

  const BlockDesc b = P.blocks[blockIdx.x];
  const uint32_t rows = b.hi - b.lo;
  const uint32_t C = P.C, mask = P.win_slots - 1;
  const uint32_t nchunks = (rows + C - 1) / C;
  const uint32_t G = (rows + 31) >> 5;          // 32-row groups of the block (C is a multiple of 32)
  const uint32_t gpc = C >> 5;                  // groups per chunk
  const uint32_t N = P.N;
  const bool rev = P.reversed != 0;
  long long clk0 = 0;
  unsigned long long ns0 = 0;
  if (P.clk && blockIdx.x == 0 && threadIdx.x == 0) {
    clk0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
  }
  if (threadIdx.x < S) { mbar_init(full + threadIdx.x, 1); mbar_init(empty + threadIdx.x, 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  double dot = 0.0;
  if (warp == NW) {
    // ================================ producer warp ==========================================================
    for (uint32_t gb = 0; gb < G; gb += 32) {
      // row-pointer values at the group boundaries of the next 32 groups (lane l: start of group gb+l)
      const uint32_t jb = min(b.lo + 32u * (gb + lane), b.hi);
      const int64_t E = P.rowptr[jb];
      const int64_t Elast = P.rowptr[min(b.lo + 32u * (gb + 32u), b.hi)];
      const uint32_t ge = min(32u, G - gb);
      for (uint32_t l = 0; l < ge; l++) {
        const uint32_t g = gb + l;
        const uint32_t slot = g % S, use = g / S;
        const int64_t e0 = __shfl_sync(0xffffffffu, E, (int)l);
        int64_t e1 = __shfl_sync(0xffffffffu, E, (int)((l + 1) & 31));
        if (l == 31) e1 = Elast;
        if (use > 0) mbar_wait(empty + slot, (use - 1) & 1u);
        if (lane == 0) {
          const uint32_t j0 = b.lo + 32u * g;
          const uint32_t nr = min(32u, b.hi - j0);
          unsigned char *sl = slots + (size_t)slot * slot_bytes;
          double *d_val = reinterpret_cast<double *>(sl);
          uint32_t *d_col = reinterpret_cast<uint32_t *>(sl + (size_t)P.cap * 8);
          double *d_rp = reinterpret_cast<double *>(sl + (size_t)P.cap * 12);
          double *d_init = d_rp + SLOT_RP;
          double *d_dot = d_init + SLOT_VEC;
          const int64_t e0s = e0 & ~3ll;
          const uint32_t n = (uint32_t)((e1 - e0s + 3) & ~3ll);            // <= cap by construction
          const uint32_t rp0 = j0 & ~1u;                                   // 16-byte aligned slice of rowptr
          const uint32_t nrp = (j0 - rp0 + nr + 1u + 1u) & ~1u;
          const uint32_t v_lo = rev ? N - j0 - nr : j0;                    // first vector index of the group
          const uint32_t v0 = v_lo & ~1u;
          const uint32_t nv = (v_lo - v0 + nr + 1u) & ~1u;
          const uint32_t bytes = n * 12u + nrp * 8u + nv * 8u * (P.dotvec ? 2u : 1u);
          mbar_expect_tx(full + slot, bytes);
          bulk_g2s(d_val, P.val + e0s, n * 8u, full + slot);
          bulk_g2s(d_col, P.col + e0s, n * 4u, full + slot);
          bulk_g2s(d_rp, P.rowptr + rp0, nrp * 8u, full + slot);
          bulk_g2s(d_init, P.init + v0, nv * 8u, full + slot);
          if (P.dotvec) bulk_g2s(d_dot, P.dotvec + v0, nv * 8u, full + slot);
        }
        __syncwarp();
      }
    }
  } else {
    // ================================ consumer warps =========================================================
    uint32_t win_s = smem_u32(win);
    asm volatile("" : "+r"(win_s));   // keep the shared-space address in a register (no S2UR in the loop)
    const uint32_t nthr_c = NW * 32u;
    for (uint32_t ci = 0; ci < nchunks; ci++) {
      const uint32_t clo = b.lo + ci * C;
      const uint32_t cn = min(C, b.hi - clo);
      const uint32_t wbase = (P.win_slots == C) ? 0u : (ci & 1u) * C;
      for (uint32_t i = threadIdx.x; i < cn; i += nthr_c) win[wbase + i] = SENTINEL;
      asm volatile("bar.sync 1, %0;" ::"r"(nthr_c) : "memory");
      // columns >= smem_lo are in the window (current or previous chunk); older ones are final in HBM
      const uint32_t smem_lo = (ci > 0 && P.win_slots != C) ? clo - C : clo;
      const bool has_old = smem_lo > b.lo;
      const uint32_t ng = (cn + 31) >> 5;
      for (uint32_t gl = warp; gl < ng; gl += NW) {
        const uint32_t g = ci * gpc + gl;
        const uint32_t slot = g % S, use = g / S;
        const uint32_t j0 = clo + 32u * gl;
        const uint32_t nr = min(32u, clo + cn - j0);
        unsigned char *sl = slots + (size_t)slot * slot_bytes;
        uint32_t sval_s = smem_u32(sl), scol_s = sval_s + P.cap * 8u;
        asm volatile("" : "+r"(sval_s), "+r"(scol_s));
        const int64_t *s_rp = reinterpret_cast<const int64_t *>(sl + (size_t)P.cap * 12) + (j0 & 1u);
        const uint32_t v_lo = rev ? N - j0 - nr : j0;
        const double *s_init = reinterpret_cast<const double *>(sl + (size_t)P.cap * 12) + SLOT_RP + (v_lo & 1u);
        const double *s_dot = s_init + SLOT_VEC;
        mbar_wait(full + slot, use & 1u);

        bool pend = lane < nr;
        const uint32_t j = j0 + lane;
        const uint32_t vj = rev ? N - 1 - j : j;
        const uint32_t vi = rev ? nr - 1 - lane : lane;          // position of the row in the staged vector slice
        const int64_t e0s = s_rp[0] & ~3ll;
        const int64_t rs = pend ? s_rp[lane] : e0s, re = pend ? s_rp[lane + 1] : e0s + 1;
        double acc = pend ? s_init[vi] : 0.0;
        const double dv = (pend && P.dotvec) ? s_dot[vi] : 0.0;
        const uint32_t my_a = win_s + 8u * ((j - b.lo) & mask);
        uint32_t kr = (uint32_t)(rs - e0s);           // relative index of the lane's current entry
        const uint32_t kd = (uint32_t)(re - 1 - e0s);  // relative index of the diagonal slot (holds 1/diag)
        acc *= lds_f64(sval_s + 8u * kd);
        // NaN payloads propagate through DMUL/DFMA: canonicalise once here so that no value in the window can ever
        // equal SENTINEL (matrix values are canonicalised at set-up)
        if (acc != acc) acc = __longlong_as_double((long long)CANON_NAN);
        if (has_old && pend) {
          const double *sval = reinterpret_cast<const double *>(sl);
          const uint32_t *scol = reinterpret_cast<const uint32_t *>(sl + (size_t)P.cap * 8);
          uint32_t ko = kr;
          while (ko < kd && scol[ko] + b.lo < smem_lo) ++ko;
#pragma unroll 4
          for (uint32_t k = kr; k < ko; ++k) {
            const uint32_t c = scol[k] + b.lo;
            acc = fma(sval[k], P.out[rev ? N - 1 - c : c], acc);
          }
          kr = ko;
        }
        // Entry kr (window byte address a0, value v0) is polled; entry kr+1 is kept raw (relative column c1, value
        // v1).  Indices are clamped to the diagonal slot, whose column is the row itself.
        uint32_t kn = min(kr, kd);
        uint32_t a0 = win_s + 8u * (lds_u32(scol_s + 4u * kn) & mask);
        double v0 = lds_f64(sval_s + 8u * kn);
        kn = min(kr + 1u, kd);
        uint32_t c1 = lds_u32(scol_s + 4u * kn);
        double v1 = lds_f64(sval_s + 8u * kn);
        uint32_t trips = 0, t_start = 0, pend_u = pend ? 1u : 0u;
        if (P.trace) t_start = (uint32_t)clock64();
        double *const outp = P.out + vj;
        // ---- sync-free polling on the window: convergent, branch-free trips -----------------------------------
        // One trip = one PTX block so that the predicates stay in predicate registers (through C++ every predicated
        // helper costs a P2R/ISETP pair): poll a0; if solved and an entry is pending: acc -= v0*x, advance to entry
        // kr+1 and refill the look-ahead entry; if the row is complete: publish it (window, HBM, dot).
        // A solved value never equals SENTINEL (NaNs are canonicalised on entry), so testing the high word is enough.
        for (;;) {
          asm volatile(
              "{\n\t"
              ".reg .pred pr, pf;\n\t"
              ".reg .b64 bits;\n\t"
              ".reg .f64 x;\n\t"
              ".reg .u32 lo, hi, t, kn, ad;\n\t"
              "and.b32 t, %4, %9;\n\t"                       // a1 = win_s + 8*(c1 & mask)
              "shl.b32 t, t, 3;\n\t"
              "add.u32 t, t, %10;\n\t"
              "ld.volatile.shared.b64 bits, [%3];\n\t"       // poll
              "mov.b64 {lo, hi}, bits;\n\t"
              "mov.b64 x, bits;\n\t"
              "setp.ne.u32 pr, hi, 0xFFFFFFFF;\n\t"
              "setp.lt.and.u32 pr, %5, %11, pr;\n\t"         // ready = solved && kr < kd
              "@pr fma.rn.f64 %0, %1, x, %0;\n\t"            // acc += (-v0) * x   (values are stored negated)
              "@pr add.u32 %5, %5, 1;\n\t"
              "@pr mov.u32 %3, t;\n\t"
              "@pr mov.f64 %1, %2;\n\t"
              "add.u32 kn, %5, 1;\n\t"
              "min.u32 kn, kn, %11;\n\t"
              "mad.lo.u32 ad, kn, 4, %12;\n\t"
              "@pr ld.shared.u32 %4, [ad];\n\t"
              "mad.lo.u32 ad, kn, 8, %13;\n\t"
              "@pr ld.shared.f64 %2, [ad];\n\t"
              "setp.ge.u32 pf, %5, %11;\n\t"
              "setp.ne.and.u32 pf, %6, 0, pf;\n\t"           // fin = pend && kr >= kd
              "@pf st.volatile.shared.f64 [%14], %0;\n\t"
              "@pf st.global.f64 [%15], %0;\n\t"
              "@pf fma.rn.f64 %7, %0, %8, %7;\n\t"
              "@pf mov.u32 %6, 0;\n\t"
              "}"
              : "+d"(acc), "+d"(v0), "+d"(v1), "+r"(a0), "+r"(c1), "+r"(kr), "+r"(pend_u), "+d"(dot)
              : "d"(dv), "r"(mask), "r"(win_s), "r"(kd), "r"(scol_s), "r"(sval_s), "r"(my_a), "l"(outp)
              : "memory");
          ++trips;
          if (P.trace && pend && !pend_u) {
            P.trace[4 * (size_t)j + 0] = (uint32_t)clock64();
            P.trace[4 * (size_t)j + 1] = trips;
            P.trace[4 * (size_t)j + 2] = t_start;
            P.trace[4 * (size_t)j + 3] = blockIdx.x * 1024u + threadIdx.x;
          }
          pend = pend_u != 0u;
          if (!__any_sync(0xffffffffu, pend)) break;
          if (trips > WATCHDOG_TRIPS) {   // a dependency never arrived (corrupt input or a bug): never hang the GPU
            if (pend) {
              if (P.clk) atomicExch(P.clk + 2, 1ull + j);
              sts_volatile_u64_if(my_a, CANON_NAN, true);
              P.out[vj] = __longlong_as_double((long long)CANON_NAN);
            }
            break;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + slot);   // the slot may be refilled
      }
      asm volatile("bar.sync 1, %0;" ::"r"(nthr_c) : "memory");
    }
  }
  if (P.dot_partials) {
    double t = block_sum(dot, red);
    if (threadIdx.x == 0) P.dot_partials[blockIdx.x] = t;
  }
  if (P.clk && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long ns1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
    P.clk[0] = (unsigned long long)(clock64() - clk0);
    P.clk[1] = ns1 - ns0;
  }
}

// ---------------------------------------------------------------------------------------------------------
// fused vector kernels
// ---------------------------------------------------------------------------------------------------------
// r = b, x = 0, p = 0 ; bb = rr = b.b ; it = 0        (pcg.cpp:67-81; x0 = 0 is assumed by the reference)
__global__ void __launch_bounds__(256) k_init_solve(const double *__restrict__ b, double *__restrict__ x,
                                                    double *__restrict__ r, double *__restrict__ p, uint32_t N,
                                                    double *partials, int pstride, unsigned int *counter,
                                                    PcgScalars *scal) {
  __shared__ double sm[32];
  double s = 0.0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const double bi = b[i];
    r[i] = bi;
    x[i] = 0.0;
    p[i] = 0.0;
    s = fma(bi, bi, s);
  }
  double v[1] = {s};
  double *const out[1] = {&scal->bb};
  publish_and_finalize<1>(v, partials, pstride, counter, out, sm);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    scal->it = 0;
    scal->rz = 0.0;
    scal->rz_prev = 0.0;
    scal->pq = 0.0;
    scal->pr = 0.0;
  }
}

// rz = sum of the backward solve's per-block partials; beta = rz / rz_prev (0 in the first iteration, where the
// reference copies z into p, pcg.cpp:87-90); p = z + beta p.
__global__ void __launch_bounds__(256) k_p_update(const double *__restrict__ z, double *__restrict__ p, uint32_t N,
                                                  const double *__restrict__ rz_partials, int n_partials,
                                                  PcgScalars *scal) {
  __shared__ double sm[32];
  const double rz = sum_partials(rz_partials, n_partials, sm);
  const int it = scal->it;
  const double beta = it == 0 ? 0.0 : rz / scal->rz_prev;
  if (it == 0) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) p[i] = z[i];
  } else {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
      p[i] = fma(beta, p[i], z[i]);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) scal->rz = rz;
}

// alpha = (p.r)/(p.q) ; x += alpha p ; r -= alpha q ; rr = r.r ; it++ ; rz_prev = rz      (pcg.cpp:101-110)
__global__ void __launch_bounds__(256) k_xr_update(const double *__restrict__ p, const double *__restrict__ q,
                                                   double *__restrict__ x, double *__restrict__ r, uint32_t N,
                                                   double *partials, int pstride, unsigned int *counter,
                                                   PcgScalars *scal) {
  __shared__ double sm[32];
  const double alpha = scal->pr / scal->pq;
  double s = 0.0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const double pi = p[i];
    x[i] = fma(alpha, pi, x[i]);
    const double ri = fma(-alpha, q[i], r[i]);
    r[i] = ri;
    s = fma(ri, ri, s);
  }
  double v[1] = {s};
  double *const out[1] = {&scal->rr};
  publish_and_finalize<1>(v, partials, pstride, counter, out, sm);
  // bookkeeping by the CTA that finalised would race with readers of scal->it in this same kernel only if they
  // read after the write; alpha and it are read at kernel entry by every CTA, so defer the update to the last CTA:
  // publish_and_finalize resets the counter last, so the "last CTA" is the one that sees counter == 0 afterwards.
}

// it++ and rz_prev = rz, in a 1-thread kernel after k_xr_update (keeps every reader of `it` race free)
__global__ void k_advance(PcgScalars *scal) {
  scal->rz_prev = scal->rz;
  scal->it += 1;
}

// q = A x was computed; s = sum (q - b)^2     (true residual, pcg.cpp:116-118)
__global__ void __launch_bounds__(256) k_residual_norm(const double *__restrict__ q, const double *__restrict__ b,
                                                       uint32_t N, double *partials, int pstride,
                                                       unsigned int *counter, double *out_norm2) {
  __shared__ double sm[32];
  double s = 0.0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const double d = q[i] - b[i];
    s = fma(d, d, s);
  }
  double v[1] = {s};
  double *const out[1] = {out_norm2};
  publish_and_finalize<1>(v, partials, pstride, counter, out, sm);
}

template <int LPR>
void launch_spmv_t(rcg_handle *h, const double *x, double *y, const double *r, bool dots, int grid) {
  if (dots)
    k_spmv<LPR, true><<<grid, 256, 0, h->stream>>>(h->A.rowptr, h->A.col, h->A.val, x, y, r, (uint32_t)h->N, h->partials,
                                                   h->partial_cap, h->counters + 0, h->scal);
  else
    k_spmv<LPR, false><<<grid, 256, 0, h->stream>>>(h->A.rowptr, h->A.col, h->A.val, x, y, r, (uint32_t)h->N,
                                                    h->partials, h->partial_cap, h->counters + 0, h->scal);
}

}  // namespace

int rcg_launch_spmv(rcg_handle *h, const double *x, double *y, const double *dot_r, bool with_dots) {
  const int grid = h->reduce_grid;
  switch (h->spmv_lanes) {
    case 2: launch_spmv_t<2>(h, x, y, dot_r, with_dots, grid); break;
    case 4: launch_spmv_t<4>(h, x, y, dot_r, with_dots, grid); break;
    case 8: launch_spmv_t<8>(h, x, y, dot_r, with_dots, grid); break;
    case 16: launch_spmv_t<16>(h, x, y, dot_r, with_dots, grid); break;
    default: launch_spmv_t<32>(h, x, y, dot_r, with_dots, grid); break;
  }
  h->stats.kernel_launches += 1;
  RCG_CUDA(h, cudaGetLastError());
  return RCG_OK;
}

// One triangular solve = for every dependency group: [external part] + chain kernel.
// `dotvec` (nullable): accumulate sum_j out[j]*dotvec[j] into per-block partials at h->partials + 2*partial_cap.
static uint32_t floor_pow2(uint32_t v) {
  uint32_t p = 1;
  while ((p << 1) <= v && (p << 1) != 0) p <<= 1;
  return p;
}

int rcg_launch_trisolve(rcg_handle *h, DirectionDev &d, const double *rhs, double *out, const double *dotvec) {
  const uint32_t N = (uint32_t)h->N;
  static bool attr_set = false;
  const int SMEM_MAX = 232448 - 1024;   // 227 KB per CTA minus the static part and some slack
  if (!attr_set) {
    RCG_CUDA(h, cudaFuncSetAttribute(k_tri_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
    RCG_CUDA(h, cudaFuncSetAttribute(k_tri_chain_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
    attr_set = true;
  }
  double *rz_part = h->partials + 2 * (size_t)h->partial_cap;
  for (const GroupHost &g : d.groups) {
    // ---- external part ---------------------------------------------------------------------------------
    if (g.ext_nnz > 0) {
      const double mean = (double)g.ext_nnz / (double)g.rows;
      const int lpr = mean <= 6.0 ? 4 : mean <= 24.0 ? 8 : 32;
      const uint32_t rows_per_cta = 256 / lpr;
      uint32_t gx = (g.max_rows + rows_per_cta - 1) / rows_per_cta;
      uint32_t cap = (uint32_t)std::max(1, h->sm_count * 8 / g.count);
      if (gx > cap) gx = cap;
      dim3 grid(gx, (unsigned)g.count);
      const CsrDev &E = d.M.ext;
      if (lpr == 4)
        k_tri_external<4><<<grid, 256, 0, h->stream>>>(E.rowptr, E.col, E.val, d.blocks + g.first, rhs, out, N, d.reversed);
      else if (lpr == 8)
        k_tri_external<8><<<grid, 256, 0, h->stream>>>(E.rowptr, E.col, E.val, d.blocks + g.first, rhs, out, N, d.reversed);
      else
        k_tri_external<32><<<grid, 256, 0, h->stream>>>(E.rowptr, E.col, E.val, d.blocks + g.first, rhs, out, N, d.reversed);
      h->stats.kernel_launches += 1;
    }
    // ---- chain part: pick warps / chunk / staging capacity for this group ----------------------------------
    int threads = h->opt.chain_threads > 0 ? h->opt.chain_threads : 256;
    threads = std::min(992, std::max(32, (threads + 31) / 32 * 32));   // + 1 producer warp
    // never more warps than 32-row groups in the largest block
    const int max_groups = (int)((g.max_rows + 31) / 32);
    if (threads > max_groups * 32) threads = std::max(32, max_groups * 32);
    const uint32_t NW = (uint32_t)threads / 32;
    uint32_t C = floor_pow2(h->opt.chain_window > 0 ? (uint32_t)h->opt.chain_window : 8192u);
    if (C < 32) C = 32;
    uint32_t need = 32;
    while (need < g.max_rows) need <<= 1;          // smallest power of two covering the largest block
    uint32_t win_slots;
    if (need <= 2 * C) { C = need; win_slots = need; }   // single chunk: one buffer is enough
    else win_slots = 2 * C;
    const uint32_t cap_need = std::max(64u, (g.max_stage + 3u) & ~3u);
    ChainArgs a;
    a.rowptr = d.M.loc.rowptr; a.col = d.M.loc.col; a.val = d.M.loc.val;
    a.blocks = d.blocks + g.first;
    a.init = g.ext_nnz > 0 ? out : rhs;
    a.out = out;
    a.dotvec = dotvec;
    a.dot_partials = dotvec ? rz_part + g.first : nullptr;
    a.N = N; a.reversed = d.reversed ? 1 : 0;
    // ---- pipelined kernel: NW consumer warps + 1 producer warp, S staging slots ---------------------------
    {
      const size_t slot_bytes = (size_t)cap_need * 12 + (40 + 2 * 36) * 8;
      uint32_t ws = win_slots, Cp = C;
      auto slots_fit = [&](uint32_t w) -> int64_t {
        return ((int64_t)SMEM_MAX - (int64_t)w * 8 - 512) / (int64_t)(slot_bytes + 16);
      };
      const int64_t want = std::max<int64_t>(2 * (int64_t)NW, 12);
      while (slots_fit(ws) < want && ws > 2048) { ws >>= 1; Cp = ws / 2; }
      int64_t S = std::min<int64_t>(slots_fit(ws), 32);
      if (!h->opt.chain_generic && (S >= (int64_t)NW + 1 || (S >= 2 && max_groups <= (int)S))) {
        a.C = Cp; a.win_slots = ws; a.cap = cap_need; a.slots = (uint32_t)S;
        a.clk = h->clk_probe;
        a.trace = h->trace;
        const size_t smem = (size_t)ws * 8 + (size_t)S * slot_bytes + (size_t)S * 16;
        k_tri_chain_fast<<<g.count, threads + 32, smem, h->stream>>>(a);
        h->stats.kernel_launches += 1;
        continue;
      }
    }
    // ---- fallback: staging groups too large for shared memory -------------------------------------------
    auto smem_bytes = [&](uint32_t ws, uint32_t cp) { return (size_t)ws * 8 + (size_t)NW * cp * 12 + (size_t)NW * 8; };
    auto cap_fit = [&](uint32_t ws) -> uint32_t {
      const int64_t room = (int64_t)SMEM_MAX - (int64_t)ws * 8 - (int64_t)NW * 8;
      return room <= 0 ? 0u : (uint32_t)(room / ((int64_t)NW * 12)) & ~3u;
    };
    while (cap_fit(win_slots) < std::min(cap_need, 1024u) && win_slots > 64) { win_slots >>= 1; C = win_slots / 2; }
    const uint32_t cap = std::max(4u, std::min(cap_need, cap_fit(win_slots)));
    a.C = C; a.win_slots = win_slots; a.cap = cap; a.slots = 0;
    a.clk = h->clk_probe;
    a.trace = h->trace;
    k_tri_chain<<<g.count, threads, smem_bytes(win_slots, cap), h->stream>>>(a);
    h->stats.kernel_launches += 1;
  }
  RCG_CUDA(h, cudaGetLastError());
  return RCG_OK;
}

int rcg_launch_init_solve(rcg_handle *h) {
  k_init_solve<<<h->reduce_grid, 256, 0, h->stream>>>(h->b, h->x, h->r, h->p, (uint32_t)h->N, h->partials,
                                                     h->partial_cap, h->counters + 1, h->scal);
  h->stats.kernel_launches += 1;
  RCG_CUDA(h, cudaGetLastError());
  return RCG_OK;
}

int rcg_launch_p_update(rcg_handle *h) {
  const double *rz_part = h->partials + 2 * (size_t)h->partial_cap;
  k_p_update<<<h->reduce_grid, 256, 0, h->stream>>>(h->z, h->p, (uint32_t)h->N, rz_part, (int)h->bwd.blocks_host.size(),
                                                   h->scal);
  h->stats.kernel_launches += 1;
  RCG_CUDA(h, cudaGetLastError());
  return RCG_OK;
}

int rcg_launch_xr_update(rcg_handle *h) {
  k_xr_update<<<h->reduce_grid, 256, 0, h->stream>>>(h->p, h->q, h->x, h->r, (uint32_t)h->N, h->partials, h->partial_cap,
                                                    h->counters + 2, h->scal);
  k_advance<<<1, 1, 0, h->stream>>>(h->scal);
  h->stats.kernel_launches += 2;
  RCG_CUDA(h, cudaGetLastError());
  return RCG_OK;
}

int rcg_launch_residual_norm(rcg_handle *h, double *out_host_norm2) {
  // q = A x ; ||q - b||^2 -> scal->pq is reused as the output slot
  RCG_TRY(rcg_launch_spmv(h, h->x, h->q, nullptr, false));
  k_residual_norm<<<h->reduce_grid, 256, 0, h->stream>>>(h->q, h->b, (uint32_t)h->N, h->partials, h->partial_cap,
                                                        h->counters + 3, &h->scal->pq);
  h->stats.kernel_launches += 1;
  RCG_CUDA(h, cudaGetLastError());
  RCG_CUDA(h, cudaMemcpyAsync(out_host_norm2, &h->scal->pq, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  return RCG_OK;
}
