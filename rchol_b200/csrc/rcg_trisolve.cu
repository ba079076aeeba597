// rchol_b200 -- sparse triangular solves with the rchol factor (sm_100a): replaces the two mkl_sparse_d_trsv calls
// of /root/reference/c++/util/pcg.cpp:141-159.
//
// A solve runs tree level by tree level (dependency groups).  Per group, three kernels:
//   k_tri_pre    w[v]  = rhs[vec(v)] - sum over external entries M[v,c] out[c]      throughput (row-gather SpMV)
//   k_tri_chain* w     = (block-local lower-triangular solve of w, in place)        latency-bound dependency chain
//   k_tri_post   out[vec(v)] = w[v]  (+ fused dot with another vector)               throughput
// `v` is the LEVEL-SPACE index: inside every nested-dissection block the rows are stored sorted by their level in the
// block's dependency DAG (set-up, rcg_setup.cu), so that the 32 rows a warp works on are mutually independent and
// the sync-free chain kernel advances one DAG level per polling trip instead of one row.  `vec(v)` maps back to the
// index space of the caller's vectors (and folds the index reversal of the backward solve).
#include <algorithm>
#include <cstdio>

#include "rcg_device.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------------
// pre: start vector of a group's rows in level space.  LPR lanes per row over the external entries (columns are
// vector-space indices of rows solved by earlier groups); LPR == 1 is the pure gather for groups without them.
// ---------------------------------------------------------------------------------------------------------
template <int LPR>
__global__ void __launch_bounds__(256) k_tri_pre(const int64_t *__restrict__ rowptr, const uint32_t *__restrict__ col,
                                                 const double *__restrict__ val, const BlockDesc *__restrict__ blocks,
                                                 const uint32_t *__restrict__ vecidx, const double *__restrict__ rhs,
                                                 const double *out, double *__restrict__ w, uint32_t col_min,
                                                 const double *__restrict__ corr) {
  // col_min / corr (multi-GPU, top separators only): external entries with column < col_min (the rank's own subtree)
  // are left out here -- their sum over ALL ranks arrives in corr[], indexed like the vectors minus col_min.
  const BlockDesc b = blocks[blockIdx.y];
  const int sub = threadIdx.x % LPR;
  const uint32_t rows_per_cta = blockDim.x / LPR;
  for (uint32_t base = b.lo + blockIdx.x * rows_per_cta; base < b.hi; base += gridDim.x * rows_per_cta) {
    const uint32_t v = base + threadIdx.x / LPR;
    double acc = 0.0;
    if (LPR > 1 && v < b.hi) {
      const int64_t e = rowptr[v + 1];
      for (int64_t k = rowptr[v] + sub; k < e; k += LPR) {
        const uint32_t c = col[k];
        if (c >= col_min) acc = fma(val[k], out[c], acc);
      }
    }
#pragma unroll
    for (int o = LPR >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (v < b.hi && sub == 0) {
      const uint32_t i = vecidx[v];
      w[v] = (corr ? rhs[i] - corr[i - col_min] : rhs[i]) - acc;
    }
  }
}

// multi-GPU forward solve: contribution of the rank's own subtree to the right-hand side of the top separators,
// sbuf[i - n_sub] = sum over external entries with column < n_sub of M[v,c] out[c]   (then summed over ranks by NCCL)
__global__ void __launch_bounds__(256) k_tri_couple(const int64_t *__restrict__ rowptr, const uint32_t *__restrict__ col,
                                                    const double *__restrict__ val, const BlockDesc *__restrict__ blocks,
                                                    const uint32_t *__restrict__ vecidx, const double *__restrict__ out,
                                                    double *__restrict__ sbuf, uint32_t n_sub) {
  const BlockDesc b = blocks[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const uint32_t wpc = blockDim.x >> 5;
  for (uint32_t v = b.lo + blockIdx.x * wpc + (threadIdx.x >> 5); v < b.hi; v += gridDim.x * wpc) {
    double acc = 0.0;
    const int64_t e = rowptr[v + 1];
    for (int64_t k = rowptr[v] + lane; k < e; k += 32) {
      const uint32_t c = col[k];
      if (c < n_sub) acc = fma(val[k], out[c], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) sbuf[vecidx[v] - n_sub] = acc;
  }
}

// post: scatter the solved block back to vector space; optional fused dot product (r.z of the PCG recurrence)
__global__ void __launch_bounds__(256) k_tri_post(const BlockDesc *__restrict__ blocks, const uint32_t *__restrict__ vecidx,
                                                  const double *__restrict__ w, double *__restrict__ out,
                                                  const double *__restrict__ dotvec, double *dot_partials,
                                                  uint32_t dot_limit) {
  __shared__ double red[32];
  const BlockDesc b = blocks[blockIdx.y];
  double dot = 0.0;
  for (uint32_t v = b.lo + blockIdx.x * blockDim.x + threadIdx.x; v < b.hi; v += gridDim.x * blockDim.x) {
    const uint32_t i = vecidx[v];
    const double x = w[v];
    out[i] = x;
    if (dotvec && i < dot_limit) dot = fma(x, dotvec[i], dot);
  }
  if (dot_partials) {
    const double t = block_sum(dot, red);
    if (threadIdx.x == 0) dot_partials[blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
}

constexpr uint32_t SLOT_RP = 40, SLOT_VEC = 36;   // doubles reserved for the row-pointer / start-vector slices of a slot

// ---------------------------------------------------------------------------------------------------------
// Dependency-chain kernels.  MODE 0: x_j = w_j/diag + sum v'_jc x_c  (values pre-scaled by -1/diag at set-up).
//                           MODE 1: level_j = 1 + max_c level_c       (set-up: DAG levels of the block, as doubles)
// ---------------------------------------------------------------------------------------------------------
struct ChainArgs {
  const int64_t *rowptr;
  const uint32_t *col;      // block-relative columns
  const double *val;
  const BlockDesc *blocks;
  double *w;                // in: start vector, out: solution (same index space as the rows)
  const uint32_t *grp_mask; // per 32-row group: level-start bits (k_tri_chain_lv)
  uint32_t C;               // chunk rows (power of two)
  uint32_t win_slots;       // 2C, or C when every block of the launch fits one chunk
  uint32_t cap;             // staging capacity per warp / slot in entries (multiple of 4)
  uint32_t slots;           // staging slots of the pipelined kernel
  uint32_t producers;       // TMA producer warps
  unsigned long long *clk;  // nullable: {SM cycles, nanoseconds, watchdog row} of CTA 0
  uint32_t *trace;          // nullable diagnostics: per row {finish cycle, polling trips, start cycle, cta*1024+thread}
  uint32_t dbg;             // timing experiments only (results become wrong): 1 = critical warp skips the level sweep,
                            // 2 = helpers skip their entries
};

template <int MODE>
__device__ __forceinline__ double combine(double acc, double v, double x) {
  return MODE == 0 ? fma(v, x, acc) : fmax(acc, x + 1.0);
}

// Fallback: one staging buffer per warp, divergent polling loop, entries beyond the staging capacity read from HBM.
// Used only when a 32-row staging group does not fit the pipelined kernel's slots (very long rows).
template <int MODE>
__global__ void __launch_bounds__(1024) k_tri_chain(const ChainArgs P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
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

  if (lane == 0) mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  int64_t n_rs = 0, n_re = 0;
  double n_init = 0.0;
  uint32_t n_j0 = 0, n_nr = 0;
  uint32_t parity = 0;

  auto prefetch_group = [&](uint32_t ci, uint32_t g) {
    const uint32_t clo = b.lo + ci * C;
    const uint32_t cn = min(C, b.hi - clo);
    n_j0 = clo + 32u * g;
    n_nr = min(32u, clo + cn - n_j0);
    if (lane < n_nr) {
      const uint32_t j = n_j0 + lane;
      n_rs = P.rowptr[j];
      n_re = P.rowptr[j + 1];
      n_init = P.w[j];
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
  auto next_group_from = [&](uint32_t ci, uint32_t g) {
    while (ci < nchunks) {
      const uint32_t cn = min(C, rows - ci * C);
      const uint32_t ng = (cn + 31) >> 5;
      if (g < ng) { prefetch_group(ci, g); return; }
      ci++;
      g = warp;
    }
    n_nr = 0;
  };
  next_group_from(0, warp);

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
      const uint32_t j0 = n_j0, nr = n_nr;
      const int64_t rs = n_rs, re = n_re;
      double acc = n_init;
      const int64_t e0s = __shfl_sync(0xffffffffu, rs, 0) & ~3ll;
      const int64_t e1 = __shfl_sync(0xffffffffu, re, (int)nr - 1);
      const uint32_t nst = (uint32_t)min((long long)P.cap, (long long)((e1 - e0s + 3) & ~3ll));
      mbar_wait(bar, parity);
      parity ^= 1u;

      bool pend = lane < nr;
      const uint32_t j = j0 + lane;
      uint32_t kr = (uint32_t)(rs - e0s);
      const uint32_t kd = (uint32_t)(re - 1 - e0s);
      uint32_t c = 0;
      double v = 0.0;
      auto load_entry = [&](uint32_t k, uint32_t &cc, double &vv) {   // cc = absolute column (stored block-relative)
        if (k < nst) { cc = scol[k] + b.lo; vv = sval[k]; }
        else { cc = P.col[e0s + k] + b.lo; vv = P.val[e0s + k]; }
      };
      if (pend) {
        uint32_t dc;
        double dinv;
        load_entry(kd, dc, dinv);
        if (MODE == 0) {
          acc *= dinv;
          if (acc != acc) acc = __longlong_as_double((long long)CANON_NAN);
        } else {
          acc = 0.0;
        }
        while (kr < kd) {
          load_entry(kr, c, v);
          if (c >= smem_lo) break;
          acc = combine<MODE>(acc, v, P.w[c]);
          ++kr;
        }
      }
      uint32_t trips = 0;
      while (pend) {
        if (kr < kd) {
          const unsigned long long bits = vwin[(c - b.lo) & mask];
          if (bits != SENTINEL) {
            acc = combine<MODE>(acc, v, __longlong_as_double((long long)bits));
            ++kr;
            if (kr < kd) load_entry(kr, c, v);
          }
        }
        if (++trips > WATCHDOG_TRIPS) {
          if (P.clk) atomicExch(P.clk + 2, 1ull + j);
          acc = __longlong_as_double((long long)CANON_NAN);
          kr = kd;
        }
        if (kr >= kd) {
          vwin[(j - b.lo) & mask] = (unsigned long long)__double_as_longlong(acc);
          P.w[j] = acc;
          pend = false;
        }
      }
      __syncwarp();
      next_group_from(ci, g + NW);
    }
    __syncthreads();
  }
}

// Pipelined variant: every staging group of the launch fits a staging slot (cap >= max_stage), so the polling loop
// reads shared memory only.  Warp-specialised TMA pipeline:
//   * warp NW (producer) runs ahead of the solve and streams, for every 32-row group, the group's CSR segment
//     (values, columns), its 33 row pointers and its slice of the start vector into a ring of S staging slots with
//     bulk copies (cp.async.bulk, completion on the slot's `full` mbarrier); a slot is reused once the consumer of its
//     previous group has arrived on the slot's `empty` mbarrier.  Measured on B200: issuing the copies only one group
//     ahead exposes ~8000 cycles per group and warp.
//   * warps 0..NW-1 (consumers) take the groups of a chunk round-robin.  A lane first consumes, without waiting, the
//     leading entries of its row whose columns are already solved (in level order that is most of the row), then
//     runs the sync-free polling loop on the rest: convergent, branch-free trips (a divergent trip costs ~160-410
//     cycles against ~75-100 for a convergent one -- scripts/ubench/); two entries are kept in registers and the
//     second one is only turned into an address one trip after its loads were issued.


// TMA producer: streams the 32-row groups g = p, p + np, ... of block `b` into the staging ring.  Several producer
// warps share the work (measured: one warp issuing four bulk copies per group caps the pipeline at ~700 cycles per
// group, more than the consumers need).
__device__ __forceinline__ void produce_groups(const ChainArgs &P, const BlockDesc &b, uint32_t G, uint32_t S,
                                               uint32_t slot_bytes, unsigned char *slots, uint64_t *full, uint64_t *empty,
                                               uint32_t p, uint32_t np, uint32_t lane) {
  for (uint32_t base = p; base < G; base += 32u * np) {
    // lane l of the warp prefetches the row-pointer bounds of group base + l*np
    const uint32_t gl = base + lane * np;
    const uint32_t jl = min(b.lo + 32u * gl, b.hi);
    const int64_t E0 = P.rowptr[jl];
    const int64_t E1 = P.rowptr[min(jl + 32u, b.hi)];
    for (uint32_t l = 0; l < 32u; l++) {
      const uint32_t g = base + l * np;
      if (g >= G) break;
      const uint32_t slot = g % S, use = g / S;
      const int64_t e0 = __shfl_sync(0xffffffffu, E0, (int)l);
      const int64_t e1 = __shfl_sync(0xffffffffu, E1, (int)l);
      if (use > 0) mbar_wait(empty + slot, (use - 1) & 1u);
      if (lane == 0) {
        const uint32_t j0 = b.lo + 32u * g;
        const uint32_t nr = min(32u, b.hi - j0);
        unsigned char *sl = slots + (size_t)slot * slot_bytes;
        double *d_val = reinterpret_cast<double *>(sl);
        uint32_t *d_col = reinterpret_cast<uint32_t *>(sl + (size_t)P.cap * 8);
        double *d_rp = reinterpret_cast<double *>(sl + (size_t)P.cap * 12);
        double *d_init = d_rp + SLOT_RP;
        const int64_t e0s = e0 & ~3ll;
        const uint32_t n = (uint32_t)((e1 - e0s + 3) & ~3ll);            // <= cap by construction
        const uint32_t rp0 = j0 & ~1u;                                   // 16-byte aligned slices
        const uint32_t nrp = (j0 - rp0 + nr + 1u + 1u) & ~1u;
        const uint32_t nv = (j0 - rp0 + nr + 1u) & ~1u;
        mbar_expect_tx(full + slot, n * 12u + nrp * 8u + nv * 8u);
        bulk_g2s(d_val, P.val + e0s, n * 8u, full + slot);
        bulk_g2s(d_col, P.col + e0s, n * 4u, full + slot);
        bulk_g2s(d_rp, P.rowptr + rp0, nrp * 8u, full + slot);
        bulk_g2s(d_init, P.w + rp0, nv * 8u, full + slot);
      }
      __syncwarp();
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(1024) k_tri_chain_fast(const ChainArgs P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t NP = P.producers, NW = (blockDim.x >> 5) - NP, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t S = P.slots;
  const uint32_t slot_bytes = P.cap * 12u + (SLOT_RP + SLOT_VEC) * 8u;
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
  long long clk0 = 0;
  unsigned long long ns0 = 0;
  if (P.clk && blockIdx.x == 0 && threadIdx.x == 0) {
    clk0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
  }
  if (threadIdx.x < S) { mbar_init(full + threadIdx.x, 1); mbar_init(empty + threadIdx.x, 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  if (warp >= NW) {
    // ================================ producer warps ===========================================================
    produce_groups(P, b, G, S, slot_bytes, slots, full, empty, warp - NW, NP, lane);
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
      // relative columns >= rel_lo are in the window (current or previous chunk); older ones are final in HBM
      const uint32_t rel_lo = (ci > 0 && P.win_slots != C) ? (ci - 1) * C : ci * C;
      const bool has_old = rel_lo > 0;
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
        const double *s_init = reinterpret_cast<const double *>(sl + (size_t)P.cap * 12) + SLOT_RP + (j0 & 1u);
        mbar_wait(full + slot, use & 1u);

        bool pend = lane < nr;
        const uint32_t j = j0 + lane;
        const int64_t e0s = s_rp[0] & ~3ll;
        // idle lanes point at the group's first real entry (entries before it belong to the previous row, whose diagonal
        // slot carries packed bits in its column)
        const int64_t rs = pend ? s_rp[lane] : s_rp[0], re = pend ? s_rp[lane + 1] : s_rp[0] + 1;
        double acc = pend ? s_init[lane] : 0.0;
        const uint32_t my_a = win_s + 8u * ((j - b.lo) & mask);
        uint32_t kr = (uint32_t)(rs - e0s);           // relative index of the lane's current entry
        const uint32_t kd = (uint32_t)(re - 1 - e0s);  // relative index of the diagonal slot (holds 1/diag)
        if (MODE == 0) {
          acc *= lds_f64(sval_s + 8u * kd);
          // NaN payloads propagate through DMUL/DFMA: canonicalise once here so that no value in the window can ever
          // equal SENTINEL (matrix values are canonicalised at set-up)
          if (acc != acc) acc = __longlong_as_double((long long)CANON_NAN);
        } else {
          acc = 0.0;
        }
        if (P.dbg & 2u) kr = kd;   // timing experiment only
        if (pend) {
          const double *sval = reinterpret_cast<const double *>(sl);
          const uint32_t *scol = reinterpret_cast<const uint32_t *>(sl + (size_t)P.cap * 8);
          if (has_old) {   // entries older than the window: final values in HBM, independent loads
            uint32_t ko = kr;
            while (ko < kd && scol[ko] < rel_lo) ++ko;
#pragma unroll 4
            for (uint32_t k = kr; k < ko; ++k) acc = combine<MODE>(acc, sval[k], P.w[b.lo + scol[k]]);
            kr = ko;
          }
          // entries whose column is already solved: consume without waiting (in level order: most of the row)
          while (kr < kd) {
            const unsigned long long bits = lds_volatile_u64(win_s + 8u * (scol[kr] & mask));
            if ((uint32_t)(bits >> 32) == 0xFFFFFFFFu) break;
            acc = combine<MODE>(acc, sval[kr], __longlong_as_double((long long)bits));
            ++kr;
          }
        }
        // Entry kr (window byte address a0, value v0) is polled; entry kr+1 is kept raw (relative column c1, value
        // v1).  Indices are clamped to the diagonal slot, whose column is the row itself.
        uint32_t kn = min(kr, kd);
        uint32_t a0 = win_s + 8u * (lds_u32(scol_s + 4u * kn) & mask);
        double v0 = lds_f64(sval_s + 8u * kn);
        kn = min(kr + 1u, kd);
        uint32_t c1 = lds_u32(scol_s + 4u * kn);
        double v1 = lds_f64(sval_s + 8u * kn);
        uint32_t trips = 0, pend_u = pend ? 1u : 0u;
        double *const outp = P.w + j;
        // ---- sync-free polling on the window: convergent, branch-free trips -----------------------------------
        // The whole loop is ONE PTX block so that the predicates (pending, ready, finished) live in predicate
        // registers across trips and the loop-back test is a single vote: poll a0; if solved and an entry is
        // pending: combine, advance to entry kr+1 and refill the look-ahead entry; if the row is complete: publish
        // it (window + HBM).  A solved value never equals SENTINEL (NaNs are canonicalised on entry), so testing the
        // high word is enough.  The loop also ends after WATCHDOG_TRIPS trips (corrupt input must not hang the GPU).
#define RCG_CHAIN_LOOP(COMBINE)                                                                                  \
        asm volatile(                                                                                            \
            "{\n\t"                                                                                              \
            ".reg .pred pr, pf, pp, pany, pw;\n\t"                                                               \
            ".reg .b64 bits;\n\t"                                                                                \
            ".reg .f64 x;\n\t"                                                                                   \
            ".reg .u32 lo, hi, t, kn, ad;\n\t"                                                                   \
            "setp.ne.u32 pp, %6, 0;\n\t"                                                                         \
            "RCG_LOOP:\n\t"                                                                                      \
            "and.b32 t, %4, %8;\n\t"                       /* a1 = win_s + 8*(c1 & mask) */                      \
            "shl.b32 t, t, 3;\n\t"                                                                               \
            "add.u32 t, t, %9;\n\t"                                                                              \
            "ld.volatile.shared.b64 bits, [%3];\n\t"       /* poll */                                            \
            "mov.b64 {lo, hi}, bits;\n\t"                                                                        \
            "mov.b64 x, bits;\n\t"                                                                               \
            "setp.ne.u32 pr, hi, 0xFFFFFFFF;\n\t"                                                                \
            "setp.lt.and.u32 pr, %5, %10, pr;\n\t"         /* ready = solved && kr < kd */                       \
            COMBINE                                                                                              \
            "@pr add.u32 %5, %5, 1;\n\t"                                                                         \
            "@pr mov.u32 %3, t;\n\t"                                                                             \
            "@pr mov.f64 %1, %2;\n\t"                                                                            \
            "add.u32 kn, %5, 1;\n\t"                                                                             \
            "min.u32 kn, kn, %10;\n\t"                                                                           \
            "mad.lo.u32 ad, kn, 4, %11;\n\t"                                                                     \
            "@pr ld.shared.u32 %4, [ad];\n\t"                                                                    \
            "mad.lo.u32 ad, kn, 8, %12;\n\t"                                                                     \
            "@pr ld.shared.f64 %2, [ad];\n\t"                                                                    \
            "setp.ge.and.u32 pf, %5, %10, pp;\n\t"         /* fin = pend && kr >= kd */                          \
            "@pf st.volatile.shared.f64 [%13], %0;\n\t"                                                          \
            "@pf st.global.f64 [%14], %0;\n\t"                                                                   \
            "and.pred pp, pp, !pf;\n\t"                                                                          \
            "add.u32 %7, %7, 1;\n\t"                                                                             \
            "setp.le.u32 pw, %7, %15;\n\t"                                                                       \
            "vote.sync.any.pred pany, pp, 0xffffffff;\n\t"                                                       \
            "and.pred pany, pany, pw;\n\t"                                                                       \
            "@pany bra RCG_LOOP;\n\t"                                                                            \
            "selp.u32 %6, 1, 0, pp;\n\t"                                                                         \
            "}"                                                                                                  \
            : "+d"(acc), "+d"(v0), "+d"(v1), "+r"(a0), "+r"(c1), "+r"(kr), "+r"(pend_u), "+r"(trips)             \
            : "r"(mask), "r"(win_s), "r"(kd), "r"(scol_s), "r"(sval_s), "r"(my_a), "l"(outp), "r"(WATCHDOG_TRIPS) \
            : "memory")
        if (MODE == 0) {
          RCG_CHAIN_LOOP("@pr fma.rn.f64 %0, %1, x, %0;\n\t");   // acc += v0 * x (values stored negated and scaled)
        } else {
          RCG_CHAIN_LOOP("add.f64 x, x, 0d3FF0000000000000;\n\t@pr max.f64 %0, %0, x;\n\t");   // level = max(level, level_c + 1)
        }
#undef RCG_CHAIN_LOOP
        if (pend_u) {   // watchdog: a dependency never arrived
          if (P.clk) atomicExch(P.clk + 2, 1ull + j);
          sts_volatile_u64_if(my_a, CANON_NAN, true);
          P.w[j] = __longlong_as_double((long long)CANON_NAN);
        }
        if (P.trace && pend) {
          P.trace[4 * (size_t)j + 0] = (uint32_t)clock64();
          P.trace[4 * (size_t)j + 1] = trips;
          P.trace[4 * (size_t)j + 2] = 0;
          P.trace[4 * (size_t)j + 3] = blockIdx.x * 1024u + threadIdx.x;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + slot);   // the slot may be refilled
      }
      asm volatile("bar.sync 1, %0;" ::"r"(nthr_c) : "memory");
    }
  }
  if (P.clk && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long ns1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
    P.clk[0] = (unsigned long long)(clock64() - clk0);
    P.clk[1] = ns1 - ns0;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Role-specialised variant (the default for MODE 0).  Measured on B200 (scripts/ubench/critical.cu): ONE warp that
// sweeps the DAG levels of a block through shared memory -- load the values the previous level just stored, two
// DFMAs, store, __syncwarp -- advances a level in ~58 cycles, against ~230 for the polling kernel above, where every
// hop is a poll/vote trip.  So the row is split:
//   * helper warps (1..H) take the 32-row groups round-robin and pre-reduce, for every row, all entries EXCEPT the
//     trailing ones whose columns lie within the last 12 DAG levels (at most 6; in level space the columns of a row
//     are sorted by level, so the rest is old and nearly always solved long ago) into a partial-sum window `part`
//     (sync-free polling as above, rarely waits);
//   * the critical warp (warp 0) walks the groups in order and, inside a group, the DAG levels in order (batches given
//     by the set-up's level-start bit mask): partial + the row's trailing "near" entries, whose columns belong to
//     earlier levels of the SAME segment and were therefore stored by this very warp -- no polling, no flags;
//   * warp H+1 is the TMA producer (same staging ring; a slot is released by the helper AND the critical warp).
// Requires the block (a window-sized segment) to fit the window: no column is ever older than the window.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_tri_chain_lv(const ChainArgs P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // Warp roles.  Warp w issues on scheduler w % 4: the critical warp (warp 0) gets scheduler 0 for itself -- the other
  // warps of that scheduler (4, 8, ...) exit at once -- because spinning helpers on the same scheduler would take
  // most of its issue slots (measured: 68 -> ~300 cycles per DAG level).  Of the remaining warps the last one is the
  // TMA producer, the others are helpers.
  const uint32_t nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t nwork = nwarps - (nwarps + 3) / 4;            // warps with w % 4 != 0
  const uint32_t NP = P.producers;
  const uint32_t H = nwork - NP;                               // helpers
  const uint32_t widx = warp - (warp + 3) / 4;                 // index of this warp among the w % 4 != 0 warps
  const bool is_critical = warp == 0, is_idle = warp != 0 && (warp & 3u) == 0;
  const bool is_producer = !is_critical && !is_idle && widx >= H;
  const uint32_t S = P.slots;
  const uint32_t slot_bytes = P.cap * 12u + (SLOT_RP + SLOT_VEC) * 8u;
  const uint32_t W = P.win_slots;               // window slots (>= rows of the block); slot W holds 0.0
  unsigned long long *win = reinterpret_cast<unsigned long long *>(smem_raw);
  unsigned long long *part = win + (W + 2);
  unsigned char *slots = reinterpret_cast<unsigned char *>(part + W);
  uint64_t *full = reinterpret_cast<uint64_t *>(slots + (size_t)S * slot_bytes);
  uint64_t *empty = full + S;

  const BlockDesc b = P.blocks[blockIdx.x];
  const uint32_t rows = b.hi - b.lo;
  const uint32_t G = (rows + 31) >> 5;
  long long clk0 = 0;
  unsigned long long ns0 = 0;
  if (P.clk && blockIdx.x == 0 && threadIdx.x == 0) {
    clk0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
  }
  if (threadIdx.x < S) { mbar_init(full + threadIdx.x, 1); mbar_init(empty + threadIdx.x, 2); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  for (uint32_t i = threadIdx.x; i < rows; i += blockDim.x) { win[i] = SENTINEL; part[i] = SENTINEL; }
  if (threadIdx.x == 0) { win[W] = 0ull; win[W + 1] = 0ull; }   // +0.0: padding column of rows with < 2 off-diagonals
  __syncthreads();

  uint32_t win_s = smem_u32(win), part_s = smem_u32(part);
  asm volatile("" : "+r"(win_s), "+r"(part_s));

  if (is_idle) return;
  if (is_producer) {
    // ================================ producer warps ===========================================================
    produce_groups(P, b, G, S, slot_bytes, slots, full, empty, widx - H, NP, lane);
  } else if (!is_critical) {
    // ================================ helper warps: partial sums of the old entries ===============================
    const uint32_t mask = 0xFFFFFFFFu;   // columns are < rows <= W: no wrap
    for (uint32_t g = widx; g < G; g += H) {
      const uint32_t slot = g % S, use = g / S;
      const uint32_t j0 = b.lo + 32u * g;
      const uint32_t nr = min(32u, b.hi - j0);
      unsigned char *sl = slots + (size_t)slot * slot_bytes;
      uint32_t sval_s = smem_u32(sl), scol_s = sval_s + P.cap * 8u;
      asm volatile("" : "+r"(sval_s), "+r"(scol_s));
      const int64_t *s_rp = reinterpret_cast<const int64_t *>(sl + (size_t)P.cap * 12) + (j0 & 1u);
      const double *s_init = reinterpret_cast<const double *>(sl + (size_t)P.cap * 12) + SLOT_RP + (j0 & 1u);
      mbar_wait(full + slot, use & 1u);

      const bool pend = lane < nr;
      const uint32_t j = j0 + lane;
      const int64_t e0s = s_rp[0] & ~3ll;
      // idle lanes point at the group's first real entry (entries before it belong to the previous row, whose diagonal
        // slot carries packed bits in its column)
        const int64_t rs = pend ? s_rp[lane] : s_rp[0], re = pend ? s_rp[lane + 1] : s_rp[0] + 1;
      double acc = pend ? s_init[lane] : 0.0;
      uint32_t kr = (uint32_t)(rs - e0s);
      const uint32_t kdiag = (uint32_t)(re - 1 - e0s);
      // the row's trailing `nn` off-diagonals (columns within the last few DAG levels; count packed into the diagonal's
      // column entry at set-up) belong to the critical warp
      const uint32_t kd = kdiag - (pend ? (lds_u32(scol_s + 4u * kdiag) >> 28) : 0u);
      acc *= lds_f64(sval_s + 8u * kdiag);
      if (acc != acc) acc = __longlong_as_double((long long)CANON_NAN);   // no value may ever equal SENTINEL
      const uint32_t my_a = part_s + 8u * (j - b.lo);
      if (P.dbg & 2u) kr = kd;
      if (pend) {   // entries whose column is already solved: consume without waiting (nearly all of them)
        const double *sval = reinterpret_cast<const double *>(sl);
        const uint32_t *scol = reinterpret_cast<const uint32_t *>(sl + (size_t)P.cap * 8);
        while (kr < kd) {
          const unsigned long long bits = lds_volatile_u64(win_s + 8u * scol[kr]);
          if ((uint32_t)(bits >> 32) == 0xFFFFFFFFu) break;
          acc = fma(sval[kr], __longlong_as_double((long long)bits), acc);
          ++kr;
        }
      }
      uint32_t kn = min(kr, kd);
      uint32_t a0 = win_s + 8u * lds_u32(scol_s + 4u * kn);
      double v0 = lds_f64(sval_s + 8u * kn);
      kn = min(kr + 1u, kd);
      uint32_t c1 = lds_u32(scol_s + 4u * kn);
      double v1 = lds_f64(sval_s + 8u * kn);
      uint32_t trips = 0, pend_u = pend ? 1u : 0u;
      // same convergent polling loop as k_tri_chain_fast; the finished partial sum goes to the `part` window only
      asm volatile(
          "{\n\t"
          ".reg .pred pr, pf, pp, pany, pw, pz;\n\t"
          ".reg .b64 bits;\n\t"
          ".reg .f64 x;\n\t"
          ".reg .u32 lo, hi, t, kn, ad;\n\t"
          "setp.ne.u32 pp, %6, 0;\n\t"
          "RCG_HLOOP:\n\t"
          "and.b32 t, %4, %8;\n\t"
          "shl.b32 t, t, 3;\n\t"
          "add.u32 t, t, %9;\n\t"
          "ld.volatile.shared.b64 bits, [%3];\n\t"
          "mov.b64 {lo, hi}, bits;\n\t"
          "mov.b64 x, bits;\n\t"
          "setp.ne.u32 pr, hi, 0xFFFFFFFF;\n\t"
          "setp.lt.and.u32 pr, %5, %10, pr;\n\t"
          "@pr fma.rn.f64 %0, %1, x, %0;\n\t"
          "@pr add.u32 %5, %5, 1;\n\t"
          "@pr mov.u32 %3, t;\n\t"
          "@pr mov.f64 %1, %2;\n\t"
          "add.u32 kn, %5, 1;\n\t"
          "min.u32 kn, kn, %10;\n\t"
          "mad.lo.u32 ad, kn, 4, %11;\n\t"
          "@pr ld.shared.u32 %4, [ad];\n\t"
          "mad.lo.u32 ad, kn, 8, %12;\n\t"
          "@pr ld.shared.f64 %2, [ad];\n\t"
          "setp.ge.and.u32 pf, %5, %10, pp;\n\t"
          "@pf st.volatile.shared.f64 [%13], %0;\n\t"
          "and.pred pp, pp, !pf;\n\t"
          "add.u32 %7, %7, 1;\n\t"
          "setp.le.u32 pw, %7, %14;\n\t"
          "vote.sync.any.pred pz, pr, 0xffffffff;\n\t"   /* nobody progressed: stay off the shared-memory pipe */
          "@!pz nanosleep.u32 %15;\n\t"
          "vote.sync.any.pred pany, pp, 0xffffffff;\n\t"
          "and.pred pany, pany, pw;\n\t"
          "@pany bra RCG_HLOOP;\n\t"
          "selp.u32 %6, 1, 0, pp;\n\t"
          "}"
          : "+d"(acc), "+d"(v0), "+d"(v1), "+r"(a0), "+r"(c1), "+r"(kr), "+r"(pend_u), "+r"(trips)
          : "r"(mask), "r"(win_s), "r"(kd), "r"(scol_s), "r"(sval_s), "r"(my_a), "r"(WATCHDOG_TRIPS), "r"(P.C)
          : "memory");
      if (pend_u) {   // watchdog: publish a NaN so that nobody waits for this row
        if (P.clk) atomicExch(P.clk + 2, 1ull + j);
        sts_volatile_u64_if(my_a, CANON_NAN, true);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + slot);
    }
  } else {
    // ================================ critical warp: the DAG levels, in order ====================================
    uint32_t mask_next = G ? P.grp_mask[b.grp0] : 0u;
    for (uint32_t g = 0; g < G; g++) {
      const uint32_t slot = g % S;
      const uint32_t g0row = 32u * g;               // first row of the group inside the block
      const uint32_t j0 = b.lo + g0row;
      const uint32_t nr = min(32u, b.hi - j0);
      const uint32_t lmask = mask_next;
      if (g + 1 < G) mask_next = P.grp_mask[b.grp0 + g + 1];   // prefetch
      unsigned char *sl = slots + (size_t)slot * slot_bytes;
      uint32_t sval_s = smem_u32(sl), scol_s = sval_s + P.cap * 8u;
      asm volatile("" : "+r"(sval_s), "+r"(scol_s));
      const bool mine = lane < nr;
      const uint32_t p_a = part_s + 8u * (g0row + lane), w_a = win_s + 8u * (g0row + lane);
      // Partial sums come from the helper warps, normally long before they are needed.  `ready` = lanes whose partial
      // has been seen; a batch only waits (re-polls) when one of ITS lanes is missing -- a row of a later level of this
      // group may legitimately still wait for a row of an earlier level of the same group.  The helper starts a group
      // after the slot's `full` barrier completed, so once the first batch's partial sums are there the staged CSR
      // segment is too: no mbarrier wait on this warp.
      unsigned long long pbits = mine ? lds_volatile_u64(p_a) : 0ull;
      uint32_t ready = __ballot_sync(0xffffffffu, (uint32_t)(pbits >> 32) != 0xFFFFFFFFu);
      uint32_t trips = 0;
      auto wait_for = [&](uint32_t bm) {
        while ((bm & ~ready) != 0u) {
          if ((uint32_t)(pbits >> 32) == 0xFFFFFFFFu) pbits = lds_volatile_u64(p_a);
          if (++trips > WATCHDOG_TRIPS) {
            if (P.clk && (uint32_t)(pbits >> 32) == 0xFFFFFFFFu) atomicExch(P.clk + 2, 1ull + j0 + lane);
            pbits = CANON_NAN;
          }
          ready = __ballot_sync(0xffffffffu, (uint32_t)(pbits >> 32) != 0xFFFFFFFFu);
        }
      };
      const uint32_t rest0 = lmask >> 1;
      uint32_t e0 = rest0 ? (uint32_t)__ffs((int)rest0) : 32u;     // end of the first batch
      wait_for((e0 >= 32u ? 0xFFFFFFFFu : ((1u << e0) - 1u)) & (nr >= 32u ? 0xFFFFFFFFu : ((1u << nr) - 1u)));

      const int64_t *s_rp = reinterpret_cast<const int64_t *>(sl + (size_t)P.cap * 12) + (j0 & 1u);
      const int64_t e0s = s_rp[0] & ~3ll;
      const uint32_t kdiag = mine ? (uint32_t)(s_rp[lane + 1] - 1 - e0s) : (uint32_t)(s_rp[0] - e0s);
      const uint32_t nn = mine ? (lds_u32(scol_s + 4u * kdiag) >> 28) : 0u;
      // The row's trailing `nn` (<= 6) off-diagonal entries, NEWEST first (k = 0 is the last entry of the row).
      // Entries whose column lies before this group are settled: they are summed once, here (`pre`).  The others are
      // live -- their columns are rows of this very group, solved by earlier batches -- and are applied per batch.
      uint32_t la[6];
      double lv[6], pre = 0.0;
      uint32_t nlive = 0;
#pragma unroll
      for (int k = 5; k >= 0; k--) {           // oldest first, so that `pre` is summed in column order
        const bool have = (uint32_t)k < nn;
        const uint32_t c = have ? lds_u32(scol_s + 4u * (kdiag - 1u - (uint32_t)k)) : 0u;
        const double v = have ? lds_f64(sval_s + 8u * (kdiag - 1u - (uint32_t)k)) : 0.0;
        const bool live = have && c >= g0row;
        if (have && !live) pre = fma(v, __longlong_as_double((long long)lds_volatile_u64(win_s + 8u * c)), pre);
        la[k] = win_s + 8u * (live ? c : W);   // column W holds 0.0
        lv[k] = live ? v : 0.0;
        nlive += live ? 1u : 0u;
      }
      const uint32_t lmax = __reduce_max_sync(0xffffffffu, nlive);   // live entries to apply per batch (usually 1-2)
      double res = 0.0;
      const uint32_t mine_mask = nr >= 32u ? 0xFFFFFFFFu : ((1u << nr) - 1u);
      if ((ready & mine_mask) == mine_mask) {
        // ---- fast path (the rule): every partial sum of the group is already there.  A lone warp on its scheduler
        // pays ~5 cycles per dependent instruction, so the batch loop is kept to the bare chain:
        // load the live columns, FMA, predicated store, __syncwarp.
        const uint32_t my_batch = (uint32_t)__popc(lmask & (0xFFFFFFFFu >> (31u - lane))) - 1u;   // batch of my row
        const uint32_t nbatch = (uint32_t)__popc(lmask);
        const double base = __longlong_as_double((long long)pbits) + pre;
        if (lmax <= 2u) {
          for (uint32_t bt = 0; bt < nbatch; bt++) {
            const double x1 = __longlong_as_double((long long)lds_volatile_u64(la[1]));
            const double x0 = __longlong_as_double((long long)lds_volatile_u64(la[0]));
            const double acc = fma(lv[0], x0, fma(lv[1], x1, base));
            const bool act = mine && my_batch == bt;
            sts_volatile_u64_if(w_a, (unsigned long long)__double_as_longlong(acc), act);
            res = act ? acc : res;
            __syncwarp();
          }
        } else {
          for (uint32_t bt = 0; bt < nbatch; bt++) {
            double acc = base;
#pragma unroll
            for (int k = 5; k >= 0; k--)
              if ((uint32_t)k < lmax) acc = fma(lv[k], __longlong_as_double((long long)lds_volatile_u64(la[k])), acc);
            const bool act = mine && my_batch == bt;
            sts_volatile_u64_if(w_a, (unsigned long long)__double_as_longlong(acc), act);
            res = act ? acc : res;
            __syncwarp();
          }
        }
      } else {
      uint32_t s0 = 0;
      for (;;) {
        // batch = rows [s0, e0) of the group = one DAG level (or the part of it inside this group)
        const bool act = mine && lane >= s0 && lane < e0;
        double acc = __longlong_as_double((long long)pbits) + pre;
#pragma unroll
        for (int k = 5; k >= 0; k--)
          if ((uint32_t)k < lmax) acc = fma(lv[k], __longlong_as_double((long long)lds_volatile_u64(la[k])), acc);
        sts_volatile_u64_if(w_a, (unsigned long long)__double_as_longlong(acc), act);
        res = act ? acc : res;
        __syncwarp();   // (no global store in here: __syncwarp orders memory and would wait for it)
        s0 = e0;
        if (s0 >= nr) break;
        const uint32_t rest = lmask >> s0 >> 1;
        e0 = rest ? s0 + (uint32_t)__ffs((int)rest) : 32u;
        const uint32_t hi_m = e0 >= 32u ? 0xFFFFFFFFu : ((1u << e0) - 1u);
        const uint32_t bm = hi_m & ~((1u << s0) - 1u) & mine_mask;
        if ((bm & ~ready) != 0u) wait_for(bm);
      }
      }
      if (mine) P.w[j0 + lane] = res;   // one coalesced store per group
      if (lane == 0) mbar_arrive(empty + slot);
    }
  }
  if (P.clk && blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long ns1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
    P.clk[0] = (unsigned long long)(clock64() - clk0);
    P.clk[1] = ns1 - ns0;
  }
}

uint32_t floor_pow2(uint32_t v) {
  uint32_t p = 1;
  while ((p << 1) <= v && (p << 1) != 0) p <<= 1;
  return p;
}

constexpr int SMEM_MAX = 232448 - 1024;   // 227 KB per CTA minus some slack

// Picks warps / chunk / staging for one dependency group and launches the chain kernel on its blocks.
template <int MODE>
int launch_chain(rcg_handle *h, const CsrDev &loc, const BlockDesc *blocks_dev, const GroupHost &g, double *w,
                 const uint32_t *grp_mask) {
  if (!(h->smem_optin_mask & (1u << MODE))) {   // per handle = per device: the attribute belongs to the device's instance of the kernel
    RCG_CUDA(h, cudaFuncSetAttribute(k_tri_chain<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
    RCG_CUDA(h, cudaFuncSetAttribute(k_tri_chain_fast<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
    h->smem_optin_mask |= 1u << MODE;
  }
  int threads = h->opt.chain_threads > 0 ? h->opt.chain_threads : 512;
  threads = std::min(992, std::max(32, (threads + 31) / 32 * 32));   // + 1 producer warp
  const int max_groups = (int)((g.max_rows + 31) / 32);
  if (threads > max_groups * 32) threads = std::max(32, max_groups * 32);   // never more warps than 32-row groups
  uint32_t NW = (uint32_t)threads / 32;
  uint32_t C = floor_pow2(h->opt.chain_window > 0 ? (uint32_t)h->opt.chain_window : 2048u);
  if (C < 32) C = 32;
  uint32_t need = 32;
  while (need < g.max_rows) need <<= 1;          // smallest power of two covering the largest block
  uint32_t win_slots;
  if (need <= 2 * C) { C = need; win_slots = need; }   // single chunk: one buffer is enough
  else win_slots = 2 * C;
  const uint32_t cap_need = std::max(64u, (g.max_stage + 3u) & ~3u);
  ChainArgs a;
  a.rowptr = loc.rowptr; a.col = loc.col; a.val = loc.val;
  a.blocks = blocks_dev;
  a.w = w;
  a.clk = h->clk_probe;
  a.trace = h->trace;
  a.grp_mask = grp_mask;
  a.dbg = (uint32_t)h->opt.reserved[1];
  // ---- role-specialised kernel: critical warp + H helper warps + producer warp ------------------------------
  // (opt-in: chain_mode == 2.  Measured on B200 it does not yet beat the polling kernel -- its per-group prologue on a
  //  lone warp costs more than the polling trips it saves)
  if (MODE == 0 && !h->opt.chain_generic && grp_mask && h->opt.chain_mode == 2 && need <= 16384) {
    if (!(h->smem_optin_mask & (1u << 8))) {
      RCG_CUDA(h, cudaFuncSetAttribute(k_tri_chain_lv, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
      h->smem_optin_mask |= 1u << 8;
    }
    const size_t slot_bytes = (size_t)cap_need * 12 + (SLOT_RP + SLOT_VEC) * 8;
    const size_t win_bytes = ((size_t)need + 2) * 8 + (size_t)need * 8;
    const int64_t S = std::min<int64_t>(((int64_t)SMEM_MAX - (int64_t)win_bytes - 512) / (int64_t)(slot_bytes + 16), 32);
    if (S >= 3) {
      // warps: 1 critical + idle partners on its scheduler + helpers + 1 producer; helpers need a slot each
      uint32_t nwarps = std::max(8u, std::min<uint32_t>(NW, 32u));
      const uint32_t NPr = h->opt.reserved[2] > 0 ? (uint32_t)h->opt.reserved[2] : 2u;
      a.producers = NPr;
      auto helpers_of = [&](uint32_t nw) { return nw - (nw + 3) / 4 - NPr; };
      while (nwarps > 8 && (int64_t)helpers_of(nwarps) + 1 > S) nwarps--;
      a.C = h->opt.reserved[0] > 0 ? (uint32_t)h->opt.reserved[0] : 100u;   // helper back-off (ns); lv kernel has no chunks
      a.win_slots = need; a.cap = cap_need; a.slots = (uint32_t)S;
      const size_t smem = win_bytes + (size_t)S * slot_bytes + (size_t)S * 16;
      k_tri_chain_lv<<<g.count, nwarps * 32, smem, h->stream>>>(a);
      h->stats.kernel_launches += 1;
      return RCG_OK;
    }
  }
  // ---- pipelined kernel: NW consumer warps + 1 producer warp, S staging slots ---------------------------
  if (!h->opt.chain_generic) {
    const size_t slot_bytes = (size_t)cap_need * 12 + (SLOT_RP + SLOT_VEC) * 8;
    uint32_t ws = win_slots, Cp = C;
    auto slots_fit = [&](uint32_t wsl) -> int64_t {
      return ((int64_t)SMEM_MAX - (int64_t)wsl * 8 - 512) / (int64_t)(slot_bytes + 16);
    };
    // Keep the window (a segment of the block fits it, so no column is ever older than the window); give the rest
    // of shared memory to staging slots and use as many consumer warps as the slots can feed.  Only when fewer than
    // 3 slots fit (very long rows) is the window shrunk (columns older than it are then read from HBM).
    while (slots_fit(ws) < 3 && ws > 2048) { ws >>= 1; Cp = ws / 2; }
    int64_t S = std::min<int64_t>(slots_fit(ws), 32);
    uint32_t nw = NW;
    if (S >= 3 && S < (int64_t)nw + 2) nw = (uint32_t)std::max<int64_t>(1, S - 2);
    if (S >= (int64_t)nw + 1 || (S >= 2 && max_groups <= (int)S)) {
      a.C = Cp; a.win_slots = ws; a.cap = cap_need; a.slots = (uint32_t)S;
      a.producers = h->opt.reserved[2] > 0 ? (uint32_t)h->opt.reserved[2] : 2u;
      if (nw + a.producers > 32u) nw = 32u - a.producers;
      const size_t smem = (size_t)ws * 8 + (size_t)S * slot_bytes + (size_t)S * 16;
      k_tri_chain_fast<MODE><<<g.count, (nw + a.producers) * 32, smem, h->stream>>>(a);
      h->stats.kernel_launches += 1;
      return RCG_OK;
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
  a.C = C; a.win_slots = win_slots; a.cap = cap; a.slots = 0; a.producers = 0;
  k_tri_chain<MODE><<<g.count, threads, smem_bytes(win_slots, cap), h->stream>>>(a);
  h->stats.kernel_launches += 1;
  return RCG_OK;
}

uint32_t aux_grid_x(const rcg_handle *h, const GroupHost &g, uint32_t rows_per_cta) {
  uint32_t gx = (g.max_rows + rows_per_cta - 1) / rows_per_cta;
  const uint32_t cap = (uint32_t)std::max(1, h->sm_count * 8 / std::max(1, g.count));
  return std::max(1u, std::min(gx, cap));
}

}  // namespace

// number of dot-partial slots a direction's post kernels write (one per CTA), and each group's first slot
int rcg_post_slots(const rcg_handle *h, const DirectionDev &d, std::vector<int> *first_slot) {
  // blocked solve: one partial per block, written by the block's publisher warp; dense-panel levels: one per panel
  if (d.bc.on) return (int)(d.bc.nblocks + d.bc.dp.npanels);
  int total = 0;
  if (first_slot) first_slot->clear();
  for (const GroupHost &g : d.groups) {
    if (first_slot) first_slot->push_back(total);
    total += (int)aux_grid_x(h, g, 256) * g.count;
  }
  return total;
}

// One triangular solve = for every dependency group: pre (start vector, external part) + chain + post (scatter, dot).
// `dotvec` (nullable): sum_j out[j]*dotvec[j] is accumulated into per-CTA partials at h->partials + 2*partial_cap.
// only_group / only_kernel (measurement): run a single group and a single kernel of it (0 chain, 1 pre, 2 post).
int rcg_launch_trisolve(rcg_handle *h, DirectionDev &d, const double *rhs, double *out, const double *dotvec,
                        int only_group, int only_kernel) {
  if (d.bc.on) return rcg_launch_blocked(h, d, rhs, out, dotvec, only_group, only_kernel);
  double *rz_part = h->partials + 2 * (size_t)h->partial_cap;
  std::vector<int> first_slot;
  rcg_post_slots(h, d, &first_slot);
  const bool dist_fwd = h->dist.on && !d.reversed && h->N > h->dist.n_sub;
  const uint32_t dot_limit = h->dist.on ? h->dist.dot_limit : 0xFFFFFFFFu;
  bool coupled = false;
  int gi = -1;
  for (const GroupHost &g : d.groups) {
    ++gi;
    if (only_group >= 0 && gi != only_group) continue;
    const BlockDesc *blocks = d.blocks + g.first;
    const bool top = dist_fwd && g.depth < h->dist.top_depth;
    if (top && !coupled) {
      // Multi-GPU forward solve, before the first top separator: every rank adds up what its own subtree contributes
      // to the right-hand sides of ALL top-separator rows, and NCCL sums that over the ranks (SURVEY 8e "reduce").
      for (const GroupHost &t : d.groups) {
        if (t.depth >= h->dist.top_depth) continue;
        dim3 grid(aux_grid_x(h, t, 8), (unsigned)t.count);
        k_tri_couple<<<grid, 256, 0, h->stream>>>(d.M.ext.rowptr, d.M.ext.col, d.M.ext.val, d.blocks + t.first, d.vecidx, out,
                                                 h->dist.sbuf, h->dist.n_sub);
        h->stats.kernel_launches += 1;
      }
      RCG_CUDA(h, cudaGetLastError());
      RCG_TRY(rcg_allreduce_sum(h, h->dist.sbuf, h->N - h->dist.n_sub));
      coupled = true;
    }
    // ---- pre --------------------------------------------------------------------------------------------
    if (only_kernel < 0 || only_kernel == 1) {
      const CsrDev &E = d.M.ext;
      const double mean = g.rows ? (double)g.ext_nnz / (double)g.rows : 0.0;
      const int lpr = g.ext_nnz == 0 ? 1 : mean <= 6.0 ? 4 : mean <= 24.0 ? 8 : 32;
      dim3 grid(aux_grid_x(h, g, 256 / lpr), (unsigned)g.count);
      const uint32_t cmin = top ? h->dist.n_sub : 0u;
      const double *corr = top ? h->dist.sbuf : nullptr;
      if (lpr == 1) k_tri_pre<1><<<grid, 256, 0, h->stream>>>(E.rowptr, E.col, E.val, blocks, d.vecidx, rhs, out, d.w, cmin, corr);
      else if (lpr == 4) k_tri_pre<4><<<grid, 256, 0, h->stream>>>(E.rowptr, E.col, E.val, blocks, d.vecidx, rhs, out, d.w, cmin, corr);
      else if (lpr == 8) k_tri_pre<8><<<grid, 256, 0, h->stream>>>(E.rowptr, E.col, E.val, blocks, d.vecidx, rhs, out, d.w, cmin, corr);
      else k_tri_pre<32><<<grid, 256, 0, h->stream>>>(E.rowptr, E.col, E.val, blocks, d.vecidx, rhs, out, d.w, cmin, corr);
      h->stats.kernel_launches += 1;
    }
    // ---- chain ------------------------------------------------------------------------------------------
    if (only_kernel < 0 || only_kernel == 0) RCG_TRY(launch_chain<0>(h, d.M.loc, blocks, g, d.w, d.grp_mask));
    // ---- post -------------------------------------------------------------------------------------------
    if (only_kernel < 0 || only_kernel == 2) {
      dim3 grid(aux_grid_x(h, g, 256), (unsigned)g.count);
      k_tri_post<<<grid, 256, 0, h->stream>>>(blocks, d.vecidx, d.w, out, dotvec, dotvec ? rz_part + first_slot[gi] : nullptr,
                                             dot_limit);
      h->stats.kernel_launches += 1;
    }
  }
  RCG_CUDA(h, cudaGetLastError());
  return RCG_OK;
}

// Set-up: DAG level of every row inside its block (longest path over local entries), as doubles in `w`.
int rcg_compute_levels(rcg_handle *h, const CsrDev &loc, const BlockDesc *blocks_dev, const std::vector<GroupHost> &groups,
                       double *w) {
  for (const GroupHost &g : groups) RCG_TRY(launch_chain<1>(h, loc, blocks_dev + g.first, g, w, nullptr));
  RCG_CUDA(h, cudaGetLastError());
  return RCG_OK;
}
