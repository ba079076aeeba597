// rchol_b200 -- dense-panel levels of the triangular solves (included by rcg_blocked.cu, inside its anonymous namespace).
//
// Replaces, for the separator levels of the nested-dissection tree, the per-block chain of 32-row hops of
// mkl_sparse_d_trsv (reference: c++/util/pcg.cpp:151,155).  A separator block is short and nearly dense next to its
// diagonal; its chain of rows/32 hops on ONE CTA leaves the rest of the GPU idle (measured at 256^3 / T = 4096: the seven
// top separator levels took 8.6 of 15.3 ms per PCG iteration).  Here the rows of every block of a level are cut into
// PANELS of C rows (C a multiple of 32, chosen per level, up to 1024).  At set-up the C x C lower-triangular diagonal
// block of every panel is inverted explicitly (k_dp_invert; the factor is strongly diagonally dominant: solving through
// inverted diagonal blocks of up to 1024 rows differs from substitution by 4e-16, tests/test_host.py numerics gate).
// At solve time all blocks of the level advance in lock step, one panel per HOP, on the whole GPU:
//     phase 1:  t_k = start_k - sum over own-block entries left of the panel  L[j,c] x_c     (8 lanes per row)
//     phase 2:  x_k = Inv_k t_k                                                              (one warp per row)
// with one grid barrier behind each phase.  The start vector (right-hand side minus the entries of other, already
// solved blocks) is formed for all rows of the level before the hops start (phase 0, HBM-bound, all rows in parallel).
// The next hop's inverses and sparse rows are prefetched into L2 a hop ahead, so every load behind a barrier is an L2 hit.
//
// Packed inverse of one panel: row-major, rows grouped by 32; every row of group g holds 32 (g + 1) doubles (zeros above
// the diagonal inside the diagonal 32 x 32 tile), so a warp reads a row as (g + 1) coalesced 256-byte segments.
#pragma once

constexpr int DP_THREADS = 1024;             // one CTA per SM (two at most): few participants keep the grid barrier cheap
constexpr uint32_t DP_WARPS = DP_THREADS / 32;
constexpr uint32_t DP_CMAX = 1024;           // largest panel
constexpr uint32_t DP_SMEM = 0;

__host__ __device__ __forceinline__ int64_t dp_row_off(uint32_t i) {   // offset (doubles) of row i inside a packed inverse
  const int64_t g = i >> 5;
  return 512 * g * (g + 1) + (int64_t)(i & 31u) * 32 * (g + 1);
}
__host__ __device__ __forceinline__ int64_t dp_inv_doubles(uint32_t m) {   // size (doubles) of the packed inverse of m rows
  const int64_t G = (m + 31u) >> 5;
  return 512 * G * (G + 1);
}

// ---------------------------------------------------------------------------------------------------------
// set-up: explicit inverse of every panel's diagonal block
// ---------------------------------------------------------------------------------------------------------
// One warp per (panel, slice of 32 columns); lane = column c of the inverse.  X = T^-1 by forward substitution over the
// rows, X[i][c] = (delta_ic - sum_{cc < i} T[i][cc] X[cc][c]) / T[i][i]: every thread reads back only what it wrote itself
// (column c), so there is nothing to synchronise; the row's pattern (col / val) is the same for all lanes.
// rp / col / val: the direction's matrix (rows sorted by column, diagonal last, raw values); far_rp / near_rp: the row
// pointers of the far and near rows -- together they count the entries of row j that lie left of its panel.
__global__ void __launch_bounds__(256) k_dp_invert(const int64_t *__restrict__ rp, const uint32_t *__restrict__ col,
                                                   const double *__restrict__ val, const int64_t *__restrict__ far_rp,
                                                   const int64_t *__restrict__ near_rp, const DpPanel *__restrict__ panels,
                                                   uint32_t npanels, uint32_t nslices, double *inv) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t wpc = blockDim.x >> 5;
  for (uint32_t task = blockIdx.x * wpc + (threadIdx.x >> 5); task < nslices; task += gridDim.x * wpc) {
    int lo = 0, hi = (int)npanels;   // largest panel index with slice0 <= task
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (panels[mid].slice0 <= task) lo = mid; else hi = mid;
    }
    const DpPanel pan = panels[lo];
    const uint32_t c0 = 32u * (task - pan.slice0), c = c0 + lane;
    double *X = inv + pan.inv_off;
    for (uint32_t i = c0; i < pan.m; i++) {
      const uint32_t j = pan.row0 + i;
      const int64_t pd = rp[j + 1] - 1;
      double acc = i == c ? 1.0 : 0.0;
      const uint32_t q = pan.q0 + i;
      for (int64_t p = rp[j] + (far_rp[j + 1] - far_rp[j]) + (near_rp[q + 1] - near_rp[q]); p < pd; p++) {
        const uint32_t cc = col[p] - pan.row0;
        if (cc >= c0) acc = fma(-val[p], X[dp_row_off(cc) + c], acc);
      }
      X[dp_row_off(i) + c] = acc / val[pd];
    }
  }
}

// near-entry span of every panel (bulk prefetch)
__global__ void k_dp_spans(DpPanel *panels, uint32_t npanels, const int64_t *__restrict__ near_rp) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npanels) {
    panels[i].e0 = near_rp[panels[i].q0];
    panels[i].e1 = near_rp[panels[i].q0 + panels[i].m];
  }
}

// ---------------------------------------------------------------------------------------------------------
// solve
// ---------------------------------------------------------------------------------------------------------
struct DpArgs {
  const DpPanel *panels;       // the level's panels, hop-major
  const uint32_t *hop_ptr;     // nhops + 1 offsets into panels
  uint32_t nhops, C;
  uint32_t wpr;                // warps per row of the dense phase: 4 (C >= 512), 2 (C >= 256) or 1
  const double *inv;
  const int64_t *near_rp;      // own-block entries left of the row's panel (compact rows, DpPanel::q0)
  const uint32_t *near_col;
  const double *near_val;
  const int64_t *far_rp;       // entries of other (already solved) blocks: rows in solve space, columns in vector space
  const uint32_t *far_col;
  const double *far_val;
  const double *rhs;           // right-hand side (vector space)
  uint32_t col_min;            // multi-GPU top separators: far columns below col_min are left out ...
  const double *corr;          // ... their sum over all ranks arrives here (indexed by vector index - col_min)
  double *w;                   // start vector (solve space)
  double *out;                 // solution (vector space)
  const double *dotvec;        // nullable
  double *dot_partials;        // one slot per panel (DpPanel::slot)
  uint32_t dot_limit;
  uint32_t N;
  int reversed;
  uint32_t *bar;               // grid-barrier slots of this level, one per CTA (zero at launch)
  unsigned int *abort_g;
  unsigned long long *clk;     // nullable diagnostics (rcg_options.reserved[1] bit 0): cycle profile of CTA 0 into clk[3..9]
  unsigned long long *trace;   // nullable diagnostics (reserved[1] bit 1): [CTA][warp][16] clock64 marks of hop nhops / 2
};

__device__ __forceinline__ void dp_prefetch_bulk(const void *p, uint32_t bytes) {   // p 16-byte aligned, bytes a multiple of 16
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dp_prefetch_line(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Grid barrier without atomics: CTA c publishes the barrier's generation in slot c, warp 0 of every CTA polls all slots
// (lanes over slots).  Nothing is serialised at one L2 address: measured with one atomic counter, 12 ns per arriving CTA.
// Returns false when the wait was abandoned (time-out or another CTA's abort): the kernel then returns, results are
// garbage and the host reports RCG_ERR_CUDA.
__device__ __forceinline__ bool dp_grid_sync(const DpArgs &P, uint32_t gen, volatile uint32_t *dead_s, unsigned long long *tr = nullptr) {
  __syncthreads();
  if (tr && threadIdx.x == 0u) tr[0] = clock64();
  if (threadIdx.x < 32u) {
    if (threadIdx.x == 0u) st_release_gpu(P.bar + blockIdx.x, gen);
    if (tr && threadIdx.x == 0u) tr[1] = clock64();
    uint32_t n = 0;
    long long t0 = 0;
    for (;;) {
      bool ok = true;
      for (uint32_t c = threadIdx.x; c < gridDim.x; c += 32u) {
        uint32_t v;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(P.bar + c) : "memory");
        ok = ok && v >= gen;
      }
      if (__all_sync(0xffffffffu, ok)) break;
      if ((++n & 63u) == 0u) {
        bool dead = false;
        if (threadIdx.x == 0u) {
          if (__ldcg(P.abort_g) != 0u) dead = true;
          const long long now = clock64();
          if (n == 64u) t0 = now;
          else if (now - t0 > BC_TIMEOUT_CYCLES) { atomicCAS(P.abort_g, 0u, 0xD000u); dead = true; }
          if (dead) *dead_s = 1u;
        }
        if (__any_sync(0xffffffffu, dead)) break;
      }
    }
    if (tr && threadIdx.x == 0u) tr[2] = clock64();
    __threadfence();   // (relaxed polls + fence = acquire)
    if (tr && threadIdx.x == 0u) tr[3] = clock64();
  }
  __syncthreads();
  return *dead_s == 0u;
}

// L2 prefetch of everything hop `hop` will read: the packed inverses and the near rows (pointers, columns, values) of
// its panels, all contiguous per panel.  Bulk prefetches of up to 32 KiB, ONE per warp at a time (the instruction is
// warp-uniform: lanes with different addresses are serialised, measured ~40 cycles each), spread over `nwarps`
// participating warps (this warp = widx).
constexpr uint32_t DP_PF = 32768;
__device__ __forceinline__ void dp_prefetch_span(const char *lo, const char *hi, uint32_t piece) {
  lo = reinterpret_cast<const char *>(reinterpret_cast<uintptr_t>(lo) & ~(uintptr_t)15);
  hi = reinterpret_cast<const char *>((reinterpret_cast<uintptr_t>(hi) + 15) & ~(uintptr_t)15);
  const char *q = lo + (size_t)piece * DP_PF;
  if (q < hi) dp_prefetch_bulk(q, (uint32_t)min((ptrdiff_t)DP_PF, hi - q));
}
__device__ __forceinline__ void dp_prefetch_hop(const DpArgs &P, uint32_t hop, uint32_t widx, uint32_t nwarps, uint32_t lane,
                                                bool with_rows) {
  if (lane != 0u) return;
  const uint32_t p0 = P.hop_ptr[hop], np = P.hop_ptr[hop + 1] - p0;
  const uint32_t ipp = (uint32_t)((dp_inv_doubles(P.C) * 8 + DP_PF - 1) / DP_PF);   // inverse: pieces per (full) panel
  for (uint32_t t = widx; t < np * ipp; t += nwarps) {
    const uint32_t pi = t / ipp, pc = t - pi * ipp;
    const DpPanel *pan = P.panels + p0 + pi;
    const char *b = reinterpret_cast<const char *>(P.inv + pan->inv_off);
    dp_prefetch_span(b, b + dp_inv_doubles(pan->m) * 8, pc);
  }
  if (!with_rows) return;
  for (uint32_t t = widx; t < np * 32u; t += nwarps) {   // 32 slots per panel; slot k takes pieces k, k + 32, ... of the spans
    const uint32_t pi = t >> 5;
    const DpPanel *pan = P.panels + p0 + pi;
    const int64_t e0 = pan->e0, e1 = pan->e1;
    const uint32_t q0 = pan->q0, m = pan->m;
    const char *r0 = reinterpret_cast<const char *>(P.near_rp + q0), *r1 = reinterpret_cast<const char *>(P.near_rp + q0 + m + 1);
    const char *c0 = reinterpret_cast<const char *>(P.near_col + e0), *c1 = reinterpret_cast<const char *>(P.near_col + e1);
    const char *v0 = reinterpret_cast<const char *>(P.near_val + e0), *v1 = reinterpret_cast<const char *>(P.near_val + e1);
    const uint32_t nr = (uint32_t)((r1 - r0 + DP_PF - 1 + 32) / DP_PF), nc = (uint32_t)((c1 - c0 + DP_PF - 1 + 32) / DP_PF);
    const uint32_t nv = (uint32_t)((v1 - v0 + DP_PF - 1 + 32) / DP_PF);
    for (uint32_t k = t & 31u; k < nr + nc + nv; k += 32u) {
      if (k < nr) dp_prefetch_span(r0, r1, k);
      else if (k < nr + nc) dp_prefetch_span(c0, c1, k - nr);
      else dp_prefetch_span(v0, v1, k - nr - nc);
    }
  }
}

// ---- sparse phases ---------------------------------------------------------------------------------------------------
// Rows are numbered flat, panel * C + i.  LPR lanes per row, 32 / LPR rows per warp at a time ("batch"); the warp's first
// batch is base = gw * RPW, the next ones follow at a stride of nw * RPW.
// Everything a sparse row needs except x is static: the row pointers, columns and values of a warp's FIRST batch are
// loaded BEFORE the grid barrier in front of the phase (SpPre); behind the barrier only the gather of x, the reduction and
// the update of the start vector remain on the critical path.
struct SpPre {
  int64_t e, e1;               // next entry of this lane, end of the row
  uint32_t j;                  // row (solve space)
  uint32_t c0, c1, c2, c3;     // first four columns of this lane (where the row is shorter: c0 again, with value 0.0)
  double v0, v1, v2, v3;
  double wv;                   // start vector entry
  bool valid;                  // the lane has a row with at least one entry for this lane (else nothing is gathered)
  bool row;                    // the lane has a row
};

// stage A (issued a phase early, 5 registers): the row and its pointers
struct SpRow {
  int64_t e, e1;
  uint32_t j;
  bool row;
};
template <uint32_t LPR>
__device__ __forceinline__ SpRow dp_near_row(const DpArgs &P, const DpPanel *__restrict__ pans, uint32_t total, uint32_t base,
                                             uint32_t lane) {
  SpRow r;
  const uint32_t sub = lane & (LPR - 1u), t = base + lane / LPR;
  const uint32_t pi = t / P.C, i = t - pi * P.C;
  r.row = t < total;
  r.j = 0; r.e = r.e1 = 0;
  if (r.row) {
    const DpPanel *pan = pans + pi;
    r.row = i < pan->m;
    if (r.row) {
      r.j = pan->row0 + i;
      const uint32_t q = pan->q0 + i;
      r.e = P.near_rp[q] + sub;
      r.e1 = P.near_rp[q + 1];
    }
  }
  return r;
}
// stage B (in front of the barrier): the first entries of this lane and the start vector entry
template <uint32_t LPR>
__device__ __forceinline__ SpPre dp_near_pre(const DpArgs &P, const SpRow &r) {
  SpPre s;
  s.row = r.row;
  s.valid = false;
  s.j = r.j; s.e = r.e; s.e1 = r.e1;
  s.c0 = s.c1 = s.c2 = s.c3 = 0u;
  s.v0 = s.v1 = s.v2 = s.v3 = 0.0;
  s.wv = 0.0;
  if (s.row) {
    const int64_t e = r.e, e1 = r.e1;
    if (e < e1) {   // (padding entries gather x[c0] too -- a solved column, finite -- and multiply it by 0.0)
      s.valid = true;
      s.c0 = P.near_col[e]; s.v0 = P.near_val[e];
      s.c1 = s.c2 = s.c3 = s.c0;
      if (e + LPR < e1) { s.c1 = P.near_col[e + LPR]; s.v1 = P.near_val[e + LPR]; }
      if (e + 2 * LPR < e1) { s.c2 = P.near_col[e + 2 * LPR]; s.v2 = P.near_val[e + 2 * LPR]; }
      if (e + 3 * LPR < e1) { s.c3 = P.near_col[e + 3 * LPR]; s.v3 = P.near_val[e + 3 * LPR]; }
    }
    s.e = e + 4 * LPR;
    s.wv = __ldcg(P.w + s.j);   // (written in phase 0, barriers ago)
  }
  return s;
}

// the rest of a row from entry e on (entries e, e + LPR, ...), 4 at a time
template <uint32_t LPR, bool FAR>
__device__ __forceinline__ double dp_row_tail(const DpArgs &P, const uint32_t *__restrict__ col, const double *__restrict__ val,
                                              int64_t e, int64_t e1, double acc) {
  double a0 = acc, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  for (; e + 3 * LPR < e1; e += 4 * LPR) {
    const uint32_t c0 = col[e], c1 = col[e + LPR], c2 = col[e + 2 * LPR], c3 = col[e + 3 * LPR];
    const double v0 = val[e], v1 = val[e + LPR], v2 = val[e + 2 * LPR], v3 = val[e + 3 * LPR];
    if (!FAR || c0 >= P.col_min) a0 = fma(v0, __ldcg(P.out + c0), a0);
    if (!FAR || c1 >= P.col_min) a1 = fma(v1, __ldcg(P.out + c1), a1);
    if (!FAR || c2 >= P.col_min) a2 = fma(v2, __ldcg(P.out + c2), a2);
    if (!FAR || c3 >= P.col_min) a3 = fma(v3, __ldcg(P.out + c3), a3);
  }
  for (; e < e1; e += LPR) {
    const uint32_t c0 = col[e];
    if (!FAR || c0 >= P.col_min) a0 = fma(val[e], __ldcg(P.out + c0), a0);
  }
  return (a0 + a1) + (a2 + a3);
}

template <uint32_t LPR>
__device__ __forceinline__ double dp_lanes_sum(double acc) {
#pragma unroll
  for (uint32_t o = LPR / 2u; o > 0u; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, (int)o);
  return acc;
}

// phase 1 of one hop: start -= near entries * x.  `pre` = the warp's first batch, loaded before the barrier.
template <uint32_t LPR>
__device__ __forceinline__ void dp_near_rows(const DpArgs &P, const DpPanel *__restrict__ pans, uint32_t total, uint32_t gw,
                                             uint32_t nw, uint32_t lane, const SpPre &pre) {
  constexpr uint32_t RPW = 32u / LPR;
  const uint32_t sub = lane & (LPR - 1u);
  if (gw * RPW < total) {   // first batch (warp-uniform)
    double acc = 0.0;
    if (pre.valid) {
      const double x0 = __ldcg(P.out + pre.c0), x1 = __ldcg(P.out + pre.c1), x2 = __ldcg(P.out + pre.c2), x3 = __ldcg(P.out + pre.c3);
      acc = fma(pre.v0, x0, fma(pre.v1, x1, 0.0)) + fma(pre.v2, x2, fma(pre.v3, x3, 0.0));
      if (pre.e < pre.e1) acc = dp_row_tail<LPR, false>(P, P.near_col, P.near_val, pre.e, pre.e1, acc);
    }
    acc = dp_lanes_sum<LPR>(acc);
    if (pre.row && sub == 0u) __stcg(P.w + pre.j, pre.wv - acc);
  }
  for (uint32_t base = gw * RPW + nw * RPW; base < total; base += nw * RPW) {
    const uint32_t t = base + lane / LPR;
    const uint32_t pi = t / P.C, i = t - pi * P.C;
    bool valid = t < total;
    uint32_t j = 0, q = 0;
    if (valid) {
      const DpPanel *pan = pans + pi;
      valid = i < pan->m;
      j = pan->row0 + i;
      q = pan->q0 + i;
    }
    double acc = 0.0;
    if (valid) acc = dp_row_tail<LPR, false>(P, P.near_col, P.near_val, P.near_rp[q] + sub, P.near_rp[q + 1], 0.0);
    acc = dp_lanes_sum<LPR>(acc);
    if (valid && sub == 0u) __stcg(P.w + j, __ldcg(P.w + j) - acc);
  }
}

// phase 0: start[j] = rhs[j] - entries of other (already solved) blocks, all rows of the level, 8 lanes per row
__device__ __forceinline__ void dp_far_rows(const DpArgs &P, uint32_t total, uint32_t gw, uint32_t nw, uint32_t lane) {
  const uint32_t sub = lane & 7u;
  for (uint32_t base = gw * 4u; base < total; base += nw * 4u) {
    const uint32_t t = base + (lane >> 3);
    const uint32_t pi = t / P.C, i = t - pi * P.C;
    bool valid = t < total;
    uint32_t j = 0;
    if (valid) {
      const DpPanel *pan = P.panels + pi;
      valid = i < pan->m;
      j = pan->row0 + i;
    }
    double acc = 0.0;
    if (valid) acc = dp_row_tail<8, true>(P, P.far_col, P.far_val, P.far_rp[j] + sub, P.far_rp[j + 1], 0.0);
    acc = dp_lanes_sum<8>(acc);
    if (valid && sub == 0u) {
      const uint32_t v = P.reversed ? P.N - 1u - j : j;
      double s0 = P.rhs[v];
      if (P.corr) s0 -= P.corr[v - P.col_min];
      __stcg(P.w + j, s0 - acc);
    }
  }
}

// ---- dense phase -----------------------------------------------------------------------------------------------------
// A CTA takes R = 32 / WPR consecutive rows of one panel at a time, WPR warps per row (warp w: row w / WPR, segments
// k, k + WPR, ... of the row with k = w % WPR; a segment = 32 consecutive columns).  The packed inverse is static: the
// (up to 8) segments of the CTA's FIRST task are loaded before the barrier (DnPre); behind it: t -> shared memory, FMAs,
// warp reduction, WPR partial sums per row added in a fixed order.
struct DnPre {
  double v[8];
  uint32_t row0, m, i;         // panel rows [row0, row0 + m), this warp's row i (panel-local)
  uint32_t nload;              // t entries the task's rows read
  bool have;                   // the CTA has a task in this hop
};

template <uint32_t WPR>
__device__ __forceinline__ DnPre dp_dense_pre(const DpArgs &P, const DpPanel *__restrict__ pans, uint32_t ntasks, uint32_t task,
                                              uint32_t warp, uint32_t lane) {
  constexpr uint32_t R = 32u / WPR;
  DnPre d;
#pragma unroll
  for (int u = 0; u < 8; u++) d.v[u] = 0.0;
  d.row0 = d.m = d.i = d.nload = 0u;
  d.have = task < ntasks;
  if (d.have) {
    const uint32_t tpp = P.C / R;
    const uint32_t pi = task / tpp, ck = task - pi * tpp;
    const DpPanel *pan = pans + pi;
    d.row0 = pan->row0; d.m = pan->m;
    d.i = ck * R + warp / WPR;
    d.nload = ((ck * R + R - 1u) | 31u) + 1u;
    if (d.i < d.m) {
      const uint32_t g = d.i >> 5, k = warp % WPR;
      const double *ip = P.inv + pan->inv_off + dp_row_off(d.i) + lane;
#pragma unroll
      for (uint32_t u = 0; u < 8u; u++)
        if (k + WPR * u <= g) d.v[u] = __ldcs(ip + 32u * (k + WPR * u));
    }
  }
  return d;
}

template <uint32_t WPR>
__device__ __forceinline__ void dp_dense_tasks(const DpArgs &P, const DpPanel *__restrict__ pans, uint32_t ntasks, double *tsm,
                                               double *part, uint32_t warp, uint32_t lane, DnPre d) {
  constexpr uint32_t R = 32u / WPR;
  for (uint32_t task = blockIdx.x; task < ntasks; task += gridDim.x) {
    if (task != blockIdx.x) d = dp_dense_pre<WPR>(P, pans, ntasks, task, warp, lane);
    if (threadIdx.x < d.nload) tsm[threadIdx.x] = threadIdx.x < d.m ? __ldcg(P.w + d.row0 + threadIdx.x) : 0.0;
    __syncthreads();
    double acc = 0.0;
    if (d.i < d.m) {
      const uint32_t g = d.i >> 5, k = warp % WPR;
      const double *tp = tsm + lane;
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (uint32_t u = 0; u < 8u; u += 2u) {
        if (k + WPR * u <= g) a0 = fma(d.v[u], tp[32u * (k + WPR * u)], a0);
        if (k + WPR * (u + 1u) <= g) a1 = fma(d.v[u + 1u], tp[32u * (k + WPR * (u + 1u))], a1);
      }
      acc = a0 + a1;
      if (g >= 8u * WPR) {   // (panels of up to 1024 rows have at most 32 segments: never taken with WPR = 4)
        const DpPanel *pan = pans + task / (P.C / R);
        const double *ip = P.inv + pan->inv_off + dp_row_off(d.i) + lane;
        for (uint32_t sgm = k + 8u * WPR; sgm <= g; sgm += WPR) acc = fma(__ldcs(ip + 32u * sgm), tp[32u * sgm], acc);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0u) part[warp] = acc;
    __syncthreads();
    if (threadIdx.x < R) {
      const uint32_t i = (d.i - warp / WPR) + threadIdx.x;   // (d.i - warp / WPR = first row of the task)
      if (i < d.m) {
        double x = part[threadIdx.x * WPR];
#pragma unroll
        for (uint32_t k = 1; k < WPR; k++) x += part[threadIdx.x * WPR + k];
        const uint32_t j = d.row0 + i;
        __stcg(P.out + (P.reversed ? P.N - 1u - j : j), x);
      }
    }
  }
}

__global__ void __launch_bounds__(DP_THREADS, 1) k_dp_solve(const DpArgs P) {
  __shared__ __align__(16) double tsm[DP_CMAX];
  __shared__ double part[DP_WARPS];
  __shared__ uint32_t dead_s;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  // consecutive work items go to different SMs: warp w of CTA c is global warp w * gridDim + c
  const uint32_t gw = warp * gridDim.x + blockIdx.x, nw = gridDim.x * DP_WARPS;
  if (threadIdx.x == 0) dead_s = 0u;
  uint32_t gen = 0;
  const bool prof = P.clk != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  long long c_sync = 0, c_p1 = 0, c_p2 = 0, c_p0 = 0, c_t = 0, c_start = 0;
  if (prof) c_start = c_t = clock64();
  auto lap = [&](long long &acc) {
    if (prof) {
      const long long now = clock64();
      acc += now - c_t;
      c_t = now;
    }
  };
  const uint32_t R = 32u / P.wpr, tpp = P.C / R;   // rows per dense task, dense tasks per panel
  auto dense_pre = [&](uint32_t hop) -> DnPre {
    const uint32_t p0 = P.hop_ptr[hop], ntasks = (P.hop_ptr[hop + 1] - p0) * tpp;
    return P.wpr == 4u   ? dp_dense_pre<4>(P, P.panels + p0, ntasks, blockIdx.x, warp, lane)
           : P.wpr == 2u ? dp_dense_pre<2>(P, P.panels + p0, ntasks, blockIdx.x, warp, lane)
                         : dp_dense_pre<1>(P, P.panels + p0, ntasks, blockIdx.x, warp, lane);
  };
  // L2 prefetch of a later hop: by the CTAs that have no dense task in this hop when there are such, else by all
  auto prefetch = [&](uint32_t hop, uint32_t busy_ctas, bool with_rows) {
    if (busy_ctas + 4u <= gridDim.x) {   // at least four idle CTAs
      if (blockIdx.x >= busy_ctas)
        dp_prefetch_hop(P, hop, warp * (gridDim.x - busy_ctas) + (blockIdx.x - busy_ctas), (gridDim.x - busy_ctas) * DP_WARPS, lane, with_rows);
    } else {
      dp_prefetch_hop(P, hop, gw, nw, lane, with_rows);
    }
  };
  prefetch(0, 0, false);
  if (P.nhops > 1u) prefetch(1, 0, true);
  __syncthreads();

  // sum over a panel's rows of out * dotvec (fused r.z of the backward solve), one warp per panel, fixed order
  auto panel_dot = [&](const DpPanel *pan) {
    double d0 = 0.0, d1 = 0.0;
    uint32_t i = lane;
    for (; i + 32u < pan->m; i += 64u) {
      const uint32_t j0 = pan->row0 + i, j1 = j0 + 32u;
      const uint32_t v0 = P.reversed ? P.N - 1u - j0 : j0, v1 = P.reversed ? P.N - 1u - j1 : j1;
      if (v0 < P.dot_limit) d0 = fma(__ldcg(P.out + v0), P.dotvec[v0], d0);
      if (v1 < P.dot_limit) d1 = fma(__ldcg(P.out + v1), P.dotvec[v1], d1);
    }
    if (i < pan->m) {
      const uint32_t j0 = pan->row0 + i;
      const uint32_t v0 = P.reversed ? P.N - 1u - j0 : j0;
      if (v0 < P.dot_limit) d0 = fma(__ldcg(P.out + v0), P.dotvec[v0], d0);
    }
    d0 = warp_sum(d0 + d1);
    if (lane == 0u) P.dot_partials[pan->slot] = d0;
  };

  // ---- phase 0: start vector of all rows of the level ----------------------------------------------------------------
  dp_far_rows(P, P.hop_ptr[P.nhops] * P.C, gw, nw, lane);
  DnPre dn = dense_pre(0);
  lap(c_p0);
  if (!dp_grid_sync(P, ++gen, &dead_s)) return;
  lap(c_sync);
  for (uint32_t hop = 0; hop < P.nhops; hop++) {
    const uint32_t p0 = P.hop_ptr[hop], np = P.hop_ptr[hop + 1] - p0;
    unsigned long long *tr = (P.trace && hop == P.nhops / 2u) ? P.trace + ((size_t)blockIdx.x * 32u + warp) * 16u : nullptr;
    auto mark = [&](int k) {
      if (tr && lane == 0u) tr[k] = clock64();
    };
    mark(0);
    // (phase 1 of the next hop: own-block entries left of its panels; a whole warp per row when there are warps to
    //  spare.  Its static loads are issued early: row pointers before the dense tasks, first entries after them.)
    const uint32_t p1 = hop + 1u < P.nhops ? P.hop_ptr[hop + 1] : 0u;
    const uint32_t total = hop + 1u < P.nhops ? (P.hop_ptr[hop + 2] - p1) * P.C : 0u;
    const bool wide = total <= nw;
    const SpRow sr = wide ? dp_near_row<32>(P, P.panels + p1, total, gw, lane) : dp_near_row<8>(P, P.panels + p1, total, gw * 4u, lane);
    mark(1);
    // ---- phase 2: x = Inv t ------------------------------------------------------------------------------------
    if (P.wpr == 4u) dp_dense_tasks<4>(P, P.panels + p0, np * tpp, tsm, part, warp, lane, dn);
    else if (P.wpr == 2u) dp_dense_tasks<2>(P, P.panels + p0, np * tpp, tsm, part, warp, lane, dn);
    else dp_dense_tasks<1>(P, P.panels + p0, np * tpp, tsm, part, warp, lane, dn);
    mark(2);
    if (hop + 1u == P.nhops) {
      lap(c_p2);
      break;
    }
    const SpPre sp = wide ? dp_near_pre<32>(P, sr) : dp_near_pre<8>(P, sr);
    mark(3);
    if (hop + 2u < P.nhops) prefetch(hop + 2u, min(np * tpp, gridDim.x), true);
    mark(4);
    lap(c_p2);
    if (!dp_grid_sync(P, ++gen, &dead_s, tr ? tr + 8 : nullptr)) return;
    mark(5);
    lap(c_sync);
    if (wide) dp_near_rows<32>(P, P.panels + p1, total, gw, nw, lane, sp);
    else dp_near_rows<8>(P, P.panels + p1, total, gw, nw, lane, sp);
    mark(6);
    // fused dot of this hop's panels (their rows are final), warps taken from the far end of the grid
    if (P.dot_partials)
      for (uint32_t d = nw - 1u - gw; d < np; d += nw) panel_dot(P.panels + p0 + d);
    dn = dense_pre(hop + 1u);
    mark(7);
    lap(c_p1);
    if (!dp_grid_sync(P, ++gen, &dead_s, tr ? tr + 12 : nullptr)) return;
    lap(c_sync);
  }
  if (P.dot_partials) {   // the last hop's panels
    if (!dp_grid_sync(P, ++gen, &dead_s)) return;
    const uint32_t q0 = P.hop_ptr[P.nhops - 1u], nq = P.hop_ptr[P.nhops] - q0;
    for (uint32_t d = gw; d < nq; d += nw) panel_dot(P.panels + q0 + d);
  }
  if (prof) {
    P.clk[3] = (unsigned long long)(clock64() - c_start);
    P.clk[4] = (unsigned long long)c_sync;
    P.clk[5] = (unsigned long long)c_p0;
    P.clk[6] = (unsigned long long)c_p1;
    P.clk[7] = (unsigned long long)c_p2;
    P.clk[8] = P.nhops;
    P.clk[9] = gridDim.x;
  }
}

// ---------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------
// Panel rows of a level from its shape: time ~ hops * t_hop + bytes / bandwidth with hops = max_rows / C and
// bytes = 4 C per row, minimal at C = sqrt(t_hop * bandwidth * max_rows / (4 * rows)); t_hop * bandwidth ~ 2 us * 5 TB/s.
inline uint32_t dp_choose_panel(uint32_t max_rows, int64_t rows, int forced) {
  uint32_t C;
  if (forced > 0) C = (uint32_t)forced;
  else C = (uint32_t)std::lround(std::sqrt(1.0e7 * (double)max_rows / (4.0 * (double)std::max<int64_t>(1, rows))) / 32.0) * 32u;
  C = std::min(C, ((max_rows + 31u) / 32u) * 32u);
  return std::max(32u, std::min(DP_CMAX, (C / 32u) * 32u));
}

// Panels, hop lists and inverses of the direction's dense-panel levels.  `comb` as in rcg_build_blocked (still alive).
int dp_build(rcg_handle *h, DirectionDev &d, const CsrDev &comb) {
  BlockedDev &B = d.bc;
  DenseDev &D = B.dp;
  D.levels.assign(d.groups.size(), DpLevel());
  std::vector<DpPanel> panels;
  std::vector<uint32_t> hop_ptr;
  int64_t inv_doubles = 0;
  uint32_t slices = 0;
  for (size_t gi = 0; gi < d.groups.size(); gi++) {
    if (!B.levels[gi].dp) continue;
    const GroupHost &G = d.groups[gi];
    DpLevel &L = D.levels[gi];
    L.on = true;
    L.C = B.blocks_host[G.first].pad[0];
    L.nhops = (G.max_rows + L.C - 1u) / L.C;
    L.panel0 = (uint32_t)panels.size();
    L.hop0 = (uint32_t)hop_ptr.size();
    const int64_t inv0 = inv_doubles;
    for (uint32_t hop = 0; hop < L.nhops; hop++) {
      hop_ptr.push_back((uint32_t)panels.size() - L.panel0);
      for (int bi = G.first; bi < G.first + G.count; bi++) {
        const BcBlock &bd = B.blocks_host[bi];
        const uint32_t r0 = bd.lo + hop * L.C;
        if (r0 >= bd.hi) continue;
        DpPanel p;
        p.row0 = r0;
        p.m = std::min(L.C, bd.hi - r0);
        p.slice0 = slices;
        p.slot = B.nblocks + (uint32_t)panels.size();
        p.inv_off = inv_doubles;
        p.q0 = D.q0_of_block[bi] + hop * L.C;
        p.pad = 0; p.e0 = p.e1 = 0;
        slices += (p.m + 31u) / 32u;
        inv_doubles += dp_inv_doubles(p.m);
        panels.push_back(p);
      }
    }
    hop_ptr.push_back((uint32_t)panels.size() - L.panel0);
    L.npanels = (uint32_t)panels.size() - L.panel0;
    L.inv_bytes = (inv_doubles - inv0) * 8;
  }
  D.npanels = (uint32_t)panels.size();
  D.inv_doubles = inv_doubles;
  if (panels.empty()) return RCG_OK;
  RCG_CUDA(h, cudaMalloc(&D.panels, sizeof(DpPanel) * panels.size()));
  RCG_CUDA(h, cudaMalloc(&D.hop_ptr, sizeof(uint32_t) * hop_ptr.size()));
  RCG_CUDA(h, cudaMalloc(&D.inv, sizeof(double) * (size_t)inv_doubles + 256));
  int per_sm = 0;
  RCG_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_dp_solve, DP_THREADS, DP_SMEM));
  D.max_ctas = std::max(1, per_sm) * h->sm_count;
  RCG_CUDA(h, cudaMalloc(&D.bar, sizeof(uint32_t) * d.groups.size() * (size_t)D.max_ctas));
  RCG_CUDA(h, cudaMemcpyAsync(D.panels, panels.data(), sizeof(DpPanel) * panels.size(), cudaMemcpyHostToDevice, h->stream));
  RCG_CUDA(h, cudaMemcpyAsync(D.hop_ptr, hop_ptr.data(), sizeof(uint32_t) * hop_ptr.size(), cudaMemcpyHostToDevice, h->stream));
  RCG_CUDA(h, cudaMemsetAsync(D.inv, 0, sizeof(double) * (size_t)inv_doubles + 256, h->stream));
  RCG_CUDA(h, cudaMemsetAsync(D.bar, 0, sizeof(uint32_t) * d.groups.size() * (size_t)D.max_ctas, h->stream));
  k_dp_spans<<<(D.npanels + 255) / 256, 256, 0, h->stream>>>(D.panels, D.npanels, D.near.rowptr);
  const int igrid = (int)std::min<int64_t>(((int64_t)slices + 7) / 8, (int64_t)h->sm_count * 8);
  k_dp_invert<<<std::max(1, igrid), 256, 0, h->stream>>>(comb.rowptr, comb.col, comb.val, B.far.rowptr, D.near.rowptr, D.panels,
                                                        D.npanels, slices, D.inv);
  h->stats.kernel_launches += 2;
  RCG_CUDA(h, cudaGetLastError());
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));   // (panels / hop_ptr are host vectors of this scope)
  D.on = true;
  return RCG_OK;
}

int dp_launch(rcg_handle *h, BlockedDev &B, const BcArgs &a, size_t gi) {
  const DenseDev &D = B.dp;
  const DpLevel &L = D.levels[gi];
  DpArgs p;
  memset(&p, 0, sizeof(p));
  p.panels = D.panels + L.panel0;
  p.hop_ptr = D.hop_ptr + L.hop0;
  p.nhops = L.nhops; p.C = L.C;
  p.wpr = L.C >= 512u ? 4u : L.C >= 256u ? 2u : 1u;
  p.inv = D.inv;
  p.near_rp = D.near.rowptr; p.near_col = D.near.col; p.near_val = D.near.val;
  p.far_rp = a.far_rp; p.far_col = a.far_col; p.far_val = a.far_val;
  p.rhs = a.rhs; p.col_min = a.col_min; p.corr = a.corr;
  p.w = a.w; p.out = a.out;
  p.dotvec = a.dotvec; p.dot_partials = a.dot_partials; p.dot_limit = a.dot_limit;
  p.N = a.N; p.reversed = a.reversed;
  p.bar = D.bar + gi * (size_t)D.max_ctas;
  p.abort_g = a.abort_g;
  p.clk = (a.dbg & 1u) ? a.clk : nullptr;
  p.trace = (a.dbg & 2u) && h->clk_probe ? h->clk_probe + 16 : nullptr;
  // grid: every CTA takes part in every barrier: one CTA per SM, and no more CTAs than phase 0 (4 rows per warp at a
  // time) can use
  const int64_t tasks = ((int64_t)L.npanels * L.C + 4 * DP_WARPS - 1) / (4 * DP_WARPS);
  int64_t cap = std::min<int64_t>(D.max_ctas, h->sm_count);
  if (const char *e = getenv("RCG_DP_CTAS_PER_SM"))   // tuning experiments: resident CTAs per SM that take part
    if (atoi(e) > 0) cap = std::min<int64_t>(D.max_ctas, (int64_t)atoi(e) * h->sm_count);
  const uint32_t grid = (uint32_t)std::max<int64_t>(1, std::min<int64_t>(cap, tasks));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(DP_THREADS);
  cfg.dynamicSmemBytes = DP_SMEM;
  cfg.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = (h->opt.reserved[6] & 1) ? 0 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k_dp_solve, p);
  if (e != cudaSuccess && cfg.numAttrs == 1) {   // (see rcg_launch_blocked: the grid never exceeds what is co-resident)
    cudaGetLastError();
    cfg.numAttrs = 0;
    e = cudaLaunchKernelEx(&cfg, k_dp_solve, p);
  }
  RCG_CUDA(h, e);
  h->stats.kernel_launches += 1;
  return RCG_OK;
}
