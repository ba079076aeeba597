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
// The start vector (right-hand side minus the entries of other, already solved blocks) is formed for all rows of the
// level before the hops start (k_dp_pre: HBM-bound, all rows in parallel).  There are no grid barriers: every value
// that crosses CTAs carries its own readiness (see "solve" below).  The inverses and sparse rows of the next hops are
// prefetched into L2 three hops ahead, so the loads on the dependency chain are L2 hits.
//
// Packed inverse of one panel: row-major, rows grouped by 32; every row of group g holds 32 (g + 1) doubles (zeros above
// the diagonal inside the diagonal 32 x 32 tile), so a warp reads a row as (g + 1) coalesced 256-byte segments.
#pragma once

constexpr int DP_THREADS = 512;              // one CTA per SM, 128 registers per thread: the static loads of the next step stay in flight
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
    // (the row's bounds are loaded one row ahead; the entries of a row eight at a time, all loads of a batch in flight
    //  together: a thread has no other memory-level parallelism here.  The order of the additions is the row's order.)
    auto bounds = [&](uint32_t i, int64_t &p0, int64_t &pd) {
      const uint32_t j = pan.row0 + i, q = pan.q0 + i;
      pd = rp[j + 1] - 1;
      p0 = rp[j] + (far_rp[j + 1] - far_rp[j]) + (near_rp[q + 1] - near_rp[q]);
    };
    int64_t p0 = 0, pd = 0, np0 = 0, npd = 0;
    if (c0 < pan.m) bounds(c0, p0, pd);
    for (uint32_t i = c0; i < pan.m; i++) {
      if (i + 1u < pan.m) bounds(i + 1u, np0, npd);
      double acc = i == c ? 1.0 : 0.0;
      for (int64_t p = p0; p < pd; p += 8) {
        uint32_t cc[8];
        double vv[8], xx[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          cc[u] = 0u; vv[u] = 0.0;
          if (p + u < pd) { cc[u] = col[p + u] - pan.row0; vv[u] = val[p + u]; }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) xx[u] = (p + u < pd && cc[u] >= c0) ? X[dp_row_off(cc[u]) + c] : 0.0;
#pragma unroll
        for (int u = 0; u < 8; u++)
          if (p + u < pd && cc[u] >= c0) acc = fma(-vv[u], xx[u], acc);
      }
      X[dp_row_off(i) + c] = acc / val[pd];
      p0 = np0; pd = npd;
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
// No grid barriers (measured: 3 000 cycles each for the fences and the polling alone, two per hop).  Every value that
// crosses CTAs carries its own readiness: the three vectors below are filled with a sentinel (a NaN payload no
// computation produces) before the launch, producers store the value with one 8-byte store, consumers re-load a word
// until it is not the sentinel.  No fences are needed because no consumer infers anything about OTHER words.
//   t0[q]  start vector of row q after phase 0 (rows of hop 0: written straight to t1)
//   t1[q]  final right-hand side of the row's panel solve (t0 minus the near entries)
//   out[]  the solution (vector space), rows of this level
// Every warp works through its rows / its CTA's tasks in hop order; all CTAs are co-resident (cooperative launch), so
// every wait is on work that an earlier step of some resident CTA produces: no cycles.  Waits are bounded like all
// waits of the blocked solve (time-out -> device-wide abort word -> RCG_ERR_CUDA).
struct DpArgs {
  const DpPanel *panels;       // the level's panels, hop-major
  const uint32_t *hop_ptr;     // nhops + 1 offsets into panels
  uint32_t nhops, C;
  uint32_t wpr;                // warps per row of the dense phase: 2 (C >= 512) or 1
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
  double *t0, *t1;             // compact rows (DpPanel::q0 + i), sentinel-filled before the launch
  double *out;                 // solution (vector space); this level's rows sentinel-filled before the launch
  const double *dotvec;        // nullable
  double *dot_partials;        // one slot per panel (DpPanel::slot)
  uint32_t dot_limit;
  uint32_t N;
  int reversed;
  unsigned int *abort_g;
  unsigned long long *clk;     // nullable diagnostics (rcg_options.reserved[1] bit 0): cycle profile of CTA 0 into clk[3..9]
  unsigned long long *trace;   // nullable diagnostics (reserved[1] bit 1): [CTA][warp][16] globaltimer marks of hop nhops / 2
};

constexpr unsigned long long DP_SENT = 0xFFF8DEADBEEF0001ull;
__device__ __forceinline__ unsigned long long dp_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool dp_is_sent(double v) { return (unsigned long long)__double_as_longlong(v) == DP_SENT; }
__device__ __forceinline__ double dp_ld(const double *p) {   // L2-coherent load (never a stale L1 line)
  double v;
  asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
// Publishing store.  An atomic exchange is performed at L2 at once; a plain store may sit in the SM's write buffer until
// something flushes it, and every consumer's wait would include that time.
__device__ __forceinline__ void dp_st(double *p, double v) {
  unsigned long long old;
  asm volatile("atom.relaxed.gpu.global.exch.b64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(__double_as_longlong(v)) : "memory");
}

// Bounded waiting: called after a poll that found the sentinel; true = give up (results are garbage, the host reports it).
struct DpSpin {
  uint32_t n = 0;
  long long t0 = 0;
  bool dead = false;
};
__device__ __forceinline__ bool dp_spin_fail(const DpArgs &P, DpSpin &s) {
  if (s.dead) return true;
  if ((++s.n & 255u) == 0u) {
    if (__ldcg(P.abort_g) != 0u) { s.dead = true; return true; }
    const long long now = clock64();
    if (s.n == 256u) s.t0 = now;
    else if (now - s.t0 > BC_TIMEOUT_CYCLES) {
      atomicCAS(P.abort_g, 0u, 0xD000u);
      s.dead = true;
      return true;
    }
  }
  return false;
}
__device__ __forceinline__ double dp_wait(const DpArgs &P, const double *p, DpSpin &s) {   // re-load until the value is there
  double v = dp_ld(p);
  while (dp_is_sent(v)) {
    if (dp_spin_fail(P, s)) break;
    v = dp_ld(p);
  }
  s.n = 0;
  return v;
}

__device__ __forceinline__ void dp_prefetch_bulk(const void *p, uint32_t bytes) {   // p 16-byte aligned, bytes a multiple of 16
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// L2 prefetch of everything hop `hop` will read: the packed inverses and the near rows (pointers, columns, values) of
// its panels, all contiguous per panel.  Bulk prefetches of up to 8 KiB, at most ONE per warp and call site iteration
// (the instruction is warp-uniform: lanes with different addresses are serialised), spread evenly over all warps of the
// grid -- an SM that issues hundreds of KiB of prefetches stalls its own memory pipe for thousands of cycles (measured).
constexpr uint32_t DP_PF = 8192;
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
  for (uint32_t t = nwarps - 1u - widx; t < np * 256u; t += nwarps) {   // 256 slots per panel; slot k: pieces k, k + 256, ...
    const uint32_t pi = t >> 8;
    const DpPanel *pan = P.panels + p0 + pi;
    const int64_t e0 = pan->e0, e1 = pan->e1;
    const uint32_t q0 = pan->q0, m = pan->m;
    const char *r0 = reinterpret_cast<const char *>(P.near_rp + q0), *r1 = reinterpret_cast<const char *>(P.near_rp + q0 + m + 1);
    const char *c0 = reinterpret_cast<const char *>(P.near_col + e0), *c1 = reinterpret_cast<const char *>(P.near_col + e1);
    const char *v0 = reinterpret_cast<const char *>(P.near_val + e0), *v1 = reinterpret_cast<const char *>(P.near_val + e1);
    const uint32_t nr = (uint32_t)((r1 - r0 + DP_PF - 1 + 32) / DP_PF), nc = (uint32_t)((c1 - c0 + DP_PF - 1 + 32) / DP_PF);
    const uint32_t nv = (uint32_t)((v1 - v0 + DP_PF - 1 + 32) / DP_PF);
    for (uint32_t k = t & 255u; k < nr + nc + nv; k += 256u) {
      if (k < nr) dp_prefetch_span(r0, r1, k);
      else if (k < nr + nc) dp_prefetch_span(c0, c1, k - nr);
      else dp_prefetch_span(v0, v1, k - nr - nc);
    }
  }
}

// ---- sparse phases ---------------------------------------------------------------------------------------------------
// Rows are numbered flat, panel * C + i.  LPR lanes per row, 32 / LPR rows per warp at a time ("batch"); the warp's first
// batch is base = gw * RPW, the next ones follow at a stride of nw * RPW.

// the rest of a row from entry e on (entries e, e + LPR, ...), 4 at a time.  WAIT: the columns may still be unsolved.
template <uint32_t LPR, bool FAR, bool WAIT>
__device__ __forceinline__ double dp_row_tail(const DpArgs &P, const uint32_t *__restrict__ col, const double *__restrict__ val,
                                              int64_t e, int64_t e1, double acc, DpSpin &sp) {
  double a0 = acc, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  for (; e + 3 * LPR < e1; e += 4 * LPR) {
    const uint32_t c0 = col[e], c1 = col[e + LPR], c2 = col[e + 2 * LPR], c3 = col[e + 3 * LPR];
    const double v0 = val[e], v1 = val[e + LPR], v2 = val[e + 2 * LPR], v3 = val[e + 3 * LPR];
    double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
    if (!FAR || c0 >= P.col_min) x0 = dp_ld(P.out + c0);
    if (!FAR || c1 >= P.col_min) x1 = dp_ld(P.out + c1);
    if (!FAR || c2 >= P.col_min) x2 = dp_ld(P.out + c2);
    if (!FAR || c3 >= P.col_min) x3 = dp_ld(P.out + c3);
    if (WAIT) {
      if (dp_is_sent(x0)) x0 = dp_wait(P, P.out + c0, sp);
      if (dp_is_sent(x1)) x1 = dp_wait(P, P.out + c1, sp);
      if (dp_is_sent(x2)) x2 = dp_wait(P, P.out + c2, sp);
      if (dp_is_sent(x3)) x3 = dp_wait(P, P.out + c3, sp);
    }
    a0 = fma(v0, x0, a0);
    a1 = fma(v1, x1, a1);
    a2 = fma(v2, x2, a2);
    a3 = fma(v3, x3, a3);
  }
  for (; e < e1; e += LPR) {
    const uint32_t c0 = col[e];
    if (!FAR || c0 >= P.col_min) {
      double x0 = dp_ld(P.out + c0);
      if (WAIT && dp_is_sent(x0)) x0 = dp_wait(P, P.out + c0, sp);
      a0 = fma(val[e], x0, a0);
    }
  }
  return (a0 + a1) + (a2 + a3);
}

template <uint32_t LPR>
__device__ __forceinline__ double dp_lanes_sum(double acc) {
#pragma unroll
  for (uint32_t o = LPR / 2u; o > 0u; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, (int)o);
  return acc;
}

// phase 1 of one hop: t1 = t0 - near entries * x, for this warp's rows of the hop.  Everything but x is static (t0 was
// produced long ago), so the loads of a warp's FIRST batch are issued early, in two stages, and overlap the waits of the
// steps in between:  NearA (row pointers; issued a hop ahead)  ->  NearB (first four entries per lane, t0; issued
// before the gathers can succeed).
struct NearA {
  int64_t e, e1;
  uint32_t q;
  bool row;
};
constexpr uint32_t DP_NPRE = 8;   // entries per lane that are preloaded (256 per row with a warp per row)
struct NearB {
  int64_t e, e1;
  uint32_t q;
  uint32_t c[DP_NPRE];         // first n (1..DP_NPRE) columns of this lane
  uint32_t n;
  double v[DP_NPRE];
  double t0v;                  // start vector entry (may still be the sentinel)
  uint32_t c_last;             // the row's last (newest) column: the word the row's first lane polls before anyone gathers
  bool row, any;               // the lane has a row / at least one entry of it
  bool row_any;                // the row has entries at all
};
template <uint32_t LPR>
__device__ __forceinline__ NearA dp_near_a(const DpArgs &P, const DpPanel *__restrict__ pans, uint32_t total, uint32_t base,
                                           uint32_t lane) {
  NearA r;
  const uint32_t sub = lane & (LPR - 1u), t = base + lane / LPR;
  const uint32_t pi = t / P.C, i = t - pi * P.C;
  r.row = t < total;
  r.q = 0; r.e = r.e1 = 0;
  if (r.row) {
    const DpPanel *pan = pans + pi;
    r.row = i < pan->m;
    if (r.row) {
      r.q = pan->q0 + i;
      r.e = P.near_rp[r.q] + sub;
      r.e1 = P.near_rp[r.q + 1];
    }
  }
  return r;
}
template <uint32_t LPR>
__device__ __forceinline__ NearB dp_near_b(const DpArgs &P, const NearA &r, uint32_t lane) {
  NearB s;
  s.row = r.row; s.any = false; s.row_any = false;
  s.c_last = 0u; s.n = 0u;
  s.q = r.q; s.e = r.e; s.e1 = r.e1;
#pragma unroll
  for (uint32_t u = 0; u < DP_NPRE; u++) { s.c[u] = 0u; s.v[u] = 0.0; }
  s.t0v = 0.0;
  if (s.row) {
    const int64_t e = r.e, e1 = r.e1;
    s.any = e < e1;
#pragma unroll
    for (uint32_t u = 0; u < DP_NPRE; u++)   // (no use of a loaded value in here: the loads only issue)
      if (e + (int64_t)(u * LPR) < e1) { s.c[u] = P.near_col[e + u * LPR]; s.v[u] = P.near_val[e + u * LPR]; s.n = u + 1u; }
    s.e = e + DP_NPRE * LPR;
    s.row_any = e1 > e - (int64_t)(lane & (LPR - 1u));   // (e - sub = the row's first entry)
    if (s.row_any) s.c_last = P.near_col[e1 - 1];
    s.t0v = dp_ld(P.t0 + s.q);
  }
  return s;
}
template <uint32_t LPR>
__device__ __forceinline__ void dp_near_rows(const DpArgs &P, const DpPanel *__restrict__ pans, uint32_t total, uint32_t gw,
                                             uint32_t nw, uint32_t lane, const NearB &nb, DpSpin &sp, unsigned long long *tr) {
  constexpr uint32_t RPW = 32u / LPR;
  const uint32_t sub = lane & (LPR - 1u);
  if (gw * RPW < total) {   // first batch (warp-uniform), from the preloaded entries
    // The row's first lane polls ONE word -- the row's newest column -- before anyone gathers: scattered gathers are one
    // L2 request per lane, and thousands of warps re-polling all of theirs saturate the L2 request queues.
    if (nb.row_any && sub == 0u) dp_wait(P, P.out + nb.c_last, sp);
    __syncwarp();
    double acc = 0.0;
    if (tr && lane == 0u) tr[0] = dp_gtime();
    if (nb.any) {
      // (entries the lane does not have gather x[c[0]] again -- it is waited for anyway -- and carry the value 0.0)
      uint32_t cc[DP_NPRE];
      double x[DP_NPRE];
#pragma unroll
      for (uint32_t u = 0; u < DP_NPRE; u++) cc[u] = u < nb.n ? nb.c[u] : nb.c[0];
#pragma unroll
      for (uint32_t u = 0; u < DP_NPRE; u++) x[u] = dp_ld(P.out + cc[u]);
      for (;;) {   // (all pending words per round trip)
        bool pend = false;
#pragma unroll
        for (uint32_t u = 0; u < DP_NPRE; u++) pend = pend || dp_is_sent(x[u]);
        if (!pend || dp_spin_fail(P, sp)) break;
#pragma unroll
        for (uint32_t u = 0; u < DP_NPRE; u++)
          if (dp_is_sent(x[u])) x[u] = dp_ld(P.out + cc[u]);
      }
      sp.n = 0;
      if (tr && lane == 0u) tr[1] = dp_gtime();
      double b0 = 0.0, b1 = 0.0;
#pragma unroll
      for (uint32_t u = 0; u < DP_NPRE; u += 2u) {
        b0 = fma(nb.v[u], x[u], b0);
        b1 = fma(nb.v[u + 1u], x[u + 1u], b1);
      }
      acc = b0 + b1;
      if (nb.e < nb.e1) acc = dp_row_tail<LPR, false, true>(P, P.near_col, P.near_val, nb.e, nb.e1, acc, sp);
    }
    if (tr && lane == 0u) tr[2] = dp_gtime();
    acc = dp_lanes_sum<LPR>(acc);
    if (nb.row && sub == 0u) {
      double t0v = nb.t0v;
      if (dp_is_sent(t0v)) t0v = dp_wait(P, P.t0 + nb.q, sp);
      dp_st(P.t1 + nb.q, t0v - acc);
    }
    if (tr && lane == 0u) tr[3] = dp_gtime();
  }
  for (uint32_t base = gw * RPW + nw * RPW; base < total; base += nw * RPW) {
    const NearA r = dp_near_a<LPR>(P, pans, total, base, lane);
    if (r.row && sub == 0u && r.e1 > r.e) dp_wait(P, P.out + P.near_col[r.e1 - 1], sp);
    __syncwarp();
    double acc = 0.0;
    if (r.row && r.e < r.e1) acc = dp_row_tail<LPR, false, true>(P, P.near_col, P.near_val, r.e, r.e1, 0.0, sp);
    acc = dp_lanes_sum<LPR>(acc);
    if (r.row && sub == 0u) dp_st(P.t1 + r.q, dp_wait(P, P.t0 + r.q, sp) - acc);
  }
}

// phase 0: t0[q] (t1[q] for the rows of hop 0) = rhs[j] - entries of other (already solved) blocks, all rows of the
// level, 8 lanes per row
__device__ __forceinline__ void dp_far_rows(const DpArgs &P, uint32_t total, uint32_t hop0_panels, uint32_t gw, uint32_t nw,
                                            uint32_t lane, DpSpin &sp) {
  const uint32_t sub = lane & 7u;
  for (uint32_t base = gw * 4u; base < total; base += nw * 4u) {
    const uint32_t t = base + (lane >> 3);
    const uint32_t pi = t / P.C, i = t - pi * P.C;
    bool valid = t < total;
    uint32_t j = 0, q = 0;
    if (valid) {
      const DpPanel *pan = P.panels + pi;
      valid = i < pan->m;
      j = pan->row0 + i;
      q = pan->q0 + i;
    }
    double acc = 0.0;
    if (valid) acc = dp_row_tail<8, true, false>(P, P.far_col, P.far_val, P.far_rp[j] + sub, P.far_rp[j + 1], 0.0, sp);
    acc = dp_lanes_sum<8>(acc);
    if (valid && sub == 0u) {
      const uint32_t v = P.reversed ? P.N - 1u - j : j;
      double s0 = P.rhs[v];
      if (P.corr) s0 -= P.corr[v - P.col_min];
      const double sent = __longlong_as_double((long long)DP_SENT);
      P.out[v] = sent;   // (plain stores: the kernel ends before k_dp_solve starts)
      if (pi < hop0_panels) P.t1[q] = s0 - acc;
      else { P.t0[q] = s0 - acc; P.t1[q] = sent; }
    }
  }
}

// Start vector of all rows of the level (HBM-bound, all rows in parallel, 64 warps per SM), and the sentinels of the
// words k_dp_solve waits for.
__global__ void __launch_bounds__(256) k_dp_pre(const DpArgs P) {
  const uint32_t lane = threadIdx.x & 31u;
  DpSpin sp;
  dp_far_rows(P, P.hop_ptr[P.nhops] * P.C, P.hop_ptr[1], blockIdx.x * 8u + (threadIdx.x >> 5), gridDim.x * 8u, lane, sp);
}

// ---- dense phase -----------------------------------------------------------------------------------------------------
// A CTA (16 warps) takes R = 16 / WPR consecutive rows of one panel at a time, WPR warps per row (warp w: row w / WPR,
// segments k, k + WPR, ... of the row with k = w % WPR; a segment = 32 consecutive columns; at most 16 segments per warp:
// WPR = 2 for C >= 512).  The packed inverse is static: the segments of a task are loaded while the PREVIOUS task (or the
// previous step of the hop) still waits and computes (DenseB); then: wait for t1 -> shared memory, FMAs, warp reduction,
// WPR partial sums per row added in a fixed order.
constexpr uint32_t DP_SEG = 16;
struct DenseB {
  double v[DP_SEG];
  uint32_t row0, m, q0;        // panel rows [row0, row0 + m), compact index of its first row
  uint32_t ck;                 // the task's chunk of R rows inside the panel
  bool have;                   // there is such a task
};
template <uint32_t WPR>
__device__ __forceinline__ DenseB dp_dense_b(const DpArgs &P, const DpPanel *__restrict__ pans, uint32_t ntasks, uint32_t task,
                                             uint32_t warp, uint32_t lane) {
  constexpr uint32_t R = DP_WARPS / WPR;
  DenseB d;
#pragma unroll
  for (int u = 0; u < (int)DP_SEG; u++) d.v[u] = 0.0;
  d.row0 = d.m = d.q0 = d.ck = 0u;
  d.have = task < ntasks;
  if (d.have) {
    const uint32_t tpp = P.C / R;
    const uint32_t pi = task / tpp;
    d.ck = task - pi * tpp;
    const DpPanel *pan = pans + pi;
    d.row0 = pan->row0; d.m = pan->m; d.q0 = pan->q0;
    const uint32_t i = d.ck * R + warp / WPR;
    if (i < d.m) {
      const uint32_t g = i >> 5, k = warp % WPR;
      const double *ip = P.inv + pan->inv_off + dp_row_off(i) + lane;
#pragma unroll
      for (uint32_t u = 0; u < DP_SEG; u++)
        if (k + WPR * u <= g) d.v[u] = __ldcs(ip + 32u * (k + WPR * u));
    }
  }
  return d;
}
template <uint32_t WPR>
__device__ __forceinline__ void dp_dense_tasks(const DpArgs &P, const DpPanel *__restrict__ pans, uint32_t ntasks, double *tsm,
                                               double *part, uint32_t warp, uint32_t lane, DenseB d, DpSpin &sp,
                                               unsigned long long *tr) {
  constexpr uint32_t R = DP_WARPS / WPR;
  for (uint32_t task = blockIdx.x; task < ntasks; task += gridDim.x) {
    // the CTA's next task of this hop: its loads are in flight while this one waits and computes
    const DenseB nx = dp_dense_b<WPR>(P, pans, ntasks, task + gridDim.x, warp, lane);
    if (d.ck * R < d.m) {   // (CTA-uniform; a panel shorter than C has empty chunks)
      const uint32_t i = d.ck * R + warp / WPR, k = warp % WPR, g = i >> 5;
      const uint32_t nload = ((d.ck * R + R - 1u) | 31u) + 1u;   // t entries the task's rows read
      if (tr && lane == 0u) tr[4] = dp_gtime();
      {   // (nload <= 1024: at most two words per thread, both loads in flight before either is waited for)
        const uint32_t qa = threadIdx.x, qb = threadIdx.x + DP_THREADS;
        double ta = 0.0, tb = 0.0;
        if (qa < d.m && qa < nload) ta = dp_ld(P.t1 + d.q0 + qa);
        if (qb < d.m && qb < nload) tb = dp_ld(P.t1 + d.q0 + qb);
        if (dp_is_sent(ta)) ta = dp_wait(P, P.t1 + d.q0 + qa, sp);
        if (dp_is_sent(tb)) tb = dp_wait(P, P.t1 + d.q0 + qb, sp);
        if (qa < nload) tsm[qa] = ta;
        if (qb < nload) tsm[qb] = tb;
      }
      if (tr && lane == 0u) tr[5] = dp_gtime();
      __syncthreads();
      if (tr && lane == 0u) tr[6] = dp_gtime();
      double acc = 0.0;
      if (i < d.m) {
        const double *tp = tsm + lane;
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (uint32_t u = 0; u < DP_SEG; u += 2u) {
          if (k + WPR * u <= g) a0 = fma(d.v[u], tp[32u * (k + WPR * u)], a0);
          if (k + WPR * (u + 1u) <= g) a1 = fma(d.v[u + 1u], tp[32u * (k + WPR * (u + 1u))], a1);
        }
        acc = a0 + a1;
      }
      acc = warp_sum(acc);
      if (lane == 0u) part[warp] = acc;
      __syncthreads();
      if (threadIdx.x < R) {
        const uint32_t ii = d.ck * R + threadIdx.x;
        if (ii < d.m) {
          double x = part[threadIdx.x * WPR];
#pragma unroll
          for (uint32_t kk = 1; kk < WPR; kk++) x += part[threadIdx.x * WPR + kk];
          const uint32_t j = d.row0 + ii;
          dp_st(P.out + (P.reversed ? P.N - 1u - j : j), x);
        }
      }
      if (tr && lane == 0u) tr[7] = dp_gtime();
    }
    d = nx;
  }
}

__global__ void __launch_bounds__(DP_THREADS, 1) k_dp_solve(const DpArgs P) {
  __shared__ __align__(16) double tsm[DP_CMAX];
  __shared__ double part[DP_WARPS];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  // consecutive work items go to different SMs: warp w of CTA c is global warp w * gridDim + c
  const uint32_t gw = warp * gridDim.x + blockIdx.x, nw = gridDim.x * DP_WARPS;
  DpSpin sp;
  const bool prof = P.clk != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  long long c_p1 = 0, c_p2 = 0, c_p0 = 0, c_pf = 0, c_t = 0, c_start = 0;
  if (prof) c_start = c_t = clock64();
  auto lap = [&](long long &acc) {
    if (prof) {
      const long long now = clock64();
      acc += now - c_t;
      c_t = now;
    }
  };
  const uint32_t R = DP_WARPS / P.wpr, tpp = P.C / R;   // rows per dense task, dense tasks per panel
  // the near phase of hop h takes a whole warp per row when there are warps to spare, else 8 lanes per row
  auto rows_of = [&](uint32_t hop) { return (P.hop_ptr[hop + 1] - P.hop_ptr[hop]) * P.C; };
  auto near_a = [&](uint32_t hop) -> NearA {
    if (hop >= P.nhops) return NearA{0, 0, 0u, false};
    const uint32_t total = rows_of(hop);
    return total <= nw ? dp_near_a<32>(P, P.panels + P.hop_ptr[hop], total, gw, lane)
                       : dp_near_a<8>(P, P.panels + P.hop_ptr[hop], total, gw * 4u, lane);
  };
  auto dense_b = [&](uint32_t hop) -> DenseB {
    const uint32_t p0 = P.hop_ptr[hop], ntasks = (P.hop_ptr[hop + 1] - p0) * tpp;
    return P.wpr == 2u ? dp_dense_b<2>(P, P.panels + p0, ntasks, blockIdx.x, warp, lane)
                       : dp_dense_b<1>(P, P.panels + p0, ntasks, blockIdx.x, warp, lane);
  };
  for (uint32_t hh = 0; hh < 3u && hh < P.nhops; hh++) dp_prefetch_hop(P, hh, gw, nw, lane, hh > 0u);
  lap(c_pf);
  NearA na = near_a(1);
  DenseB db = dense_b(0);

  lap(c_p0);   // (phase 0, the start vector of all rows of the level: k_dp_pre)
  for (uint32_t hop = 0; hop < P.nhops; hop++) {
    const uint32_t p0 = P.hop_ptr[hop], np = P.hop_ptr[hop + 1] - p0;
    unsigned long long *tr = (P.trace && hop == P.nhops / 2u) ? P.trace + ((size_t)blockIdx.x * 32u + warp) * 16u : nullptr;   // ([CTA][32][16]: 16 of the 32 warp slots used)
    unsigned long long *tr2 = (P.trace && hop == P.nhops / 2u + 1u) ? P.trace + ((size_t)blockIdx.x * 32u + warp) * 16u + 8u : nullptr;
    if (tr2) tr = tr2;   // (the hop after the traced one: its dense marks go to the second half)
    // ---- phase 2 of this hop: x = Inv t -----------------------------------------------------------------------------
    if (P.wpr == 2u) dp_dense_tasks<2>(P, P.panels + p0, np * tpp, tsm, part, warp, lane, db, sp, tr);
    else dp_dense_tasks<1>(P, P.panels + p0, np * tpp, tsm, part, warp, lane, db, sp, tr);
    if (tr2) tr = nullptr;
    lap(c_p2);
    if (hop + 1u == P.nhops) break;
    // ---- static loads of the next hop (near entries, inverse), then its phase 1 ------------------------------------------
    const uint32_t p1 = P.hop_ptr[hop + 1], total = rows_of(hop + 1u);
    const bool wide = total <= nw;
    const NearB nb = wide ? dp_near_b<32>(P, na, lane) : dp_near_b<8>(P, na, lane);
    db = dense_b(hop + 1u);
    // L2 prefetch three hops ahead by ONE warp per CTA, paced by the solve: not before the first row of this hop is
    // solved (a CTA without work would otherwise run ahead and push the inverses of all hops through L2 at once; one
    // polling warp per CTA: thousands of warps polling one word queue up at its L2 slice, measured 17k cycles per hop)
    if (hop + 3u < P.nhops && warp == DP_WARPS - 1u) {
      if (lane == 0u) {
        const uint32_t j = P.panels[p0].row0;
        dp_wait(P, P.out + (P.reversed ? P.N - 1u - j : j), sp);
      }
      dp_prefetch_hop(P, hop + 3u, blockIdx.x, gridDim.x, lane, true);
    }
    lap(c_pf);
    if (wide) dp_near_rows<32>(P, P.panels + p1, total, gw, nw, lane, nb, sp, tr);
    else dp_near_rows<8>(P, P.panels + p1, total, gw, nw, lane, nb, sp, tr);
    na = near_a(hop + 2u);
    lap(c_p1);
  }
  // ---- fused r.z of the backward solve: sum over a panel's rows of out * dotvec, one warp per panel, fixed order --------
  if (P.dot_partials) {
    const uint32_t npan = P.hop_ptr[P.nhops];
    for (uint32_t d = gw; d < npan; d += nw) {
      const DpPanel *pan = P.panels + d;
      double d0 = 0.0, d1 = 0.0;
      for (uint32_t i = lane; i < pan->m; i += 64u) {
        const uint32_t j0 = pan->row0 + i, j1 = j0 + 32u;
        const uint32_t v0 = P.reversed ? P.N - 1u - j0 : j0, v1 = P.reversed ? P.N - 1u - j1 : j1;
        const bool h1 = i + 32u < pan->m;
        double x0 = dp_ld(P.out + v0), x1 = h1 ? dp_ld(P.out + v1) : 0.0;
        if (dp_is_sent(x0)) x0 = dp_wait(P, P.out + v0, sp);
        if (h1 && dp_is_sent(x1)) x1 = dp_wait(P, P.out + v1, sp);
        if (v0 < P.dot_limit) d0 = fma(x0, P.dotvec[v0], d0);
        if (h1 && v1 < P.dot_limit) d1 = fma(x1, P.dotvec[v1], d1);
      }
      d0 = warp_sum(d0 + d1);
      if (lane == 0u) P.dot_partials[pan->slot] = d0;
    }
  }
  if (prof) {
    P.clk[3] = (unsigned long long)(clock64() - c_start);
    P.clk[4] = (unsigned long long)c_pf;
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
inline uint32_t dp_choose_panel(uint32_t max_rows, int64_t rows, int forced, uint32_t blocks) {
  uint32_t C;
  if (forced > 0) C = (uint32_t)forced;
  else {
    C = (uint32_t)std::lround(std::sqrt(1.0e7 * (double)max_rows / (4.0 * (double)std::max<int64_t>(1, rows))) / 32.0) * 32u;
    // with several panels per hop the dense phase runs best with one warp per row (rows of at most 16 segments, 16 rows
    // per CTA task): measured at 256^3, 8 blocks of 15 691 rows, C = 576 (two warps per row, 576 tasks per hop): 15.7 us per hop
    if (blocks >= 8u && C >= 512u) C = 480u;
  }
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
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  const auto t_start = std::chrono::steady_clock::now();
  RCG_CUDA(h, cudaMalloc(&D.panels, sizeof(DpPanel) * panels.size()));
  RCG_CUDA(h, cudaMalloc(&D.hop_ptr, sizeof(uint32_t) * hop_ptr.size()));
  RCG_CUDA(h, cudaMalloc(&D.inv, sizeof(double) * (size_t)inv_doubles + 256));
  int per_sm = 0;
  RCG_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_dp_solve, DP_THREADS, DP_SMEM));
  D.max_ctas = std::max(1, per_sm) * h->sm_count;
  RCG_CUDA(h, cudaMalloc(&D.t0, sizeof(double) * ((size_t)D.nrows + 4)));
  RCG_CUDA(h, cudaMalloc(&D.t1, sizeof(double) * ((size_t)D.nrows + 4)));
  RCG_CUDA(h, cudaMemcpyAsync(D.panels, panels.data(), sizeof(DpPanel) * panels.size(), cudaMemcpyHostToDevice, h->stream));
  RCG_CUDA(h, cudaMemcpyAsync(D.hop_ptr, hop_ptr.data(), sizeof(uint32_t) * hop_ptr.size(), cudaMemcpyHostToDevice, h->stream));
  RCG_CUDA(h, cudaMemsetAsync(D.inv, 0, sizeof(double) * (size_t)inv_doubles + 256, h->stream));
  k_dp_spans<<<(D.npanels + 255) / 256, 256, 0, h->stream>>>(D.panels, D.npanels, D.near.rowptr);
  const int igrid = (int)std::min<int64_t>(((int64_t)slices + 7) / 8, (int64_t)h->sm_count * 8);
  k_dp_invert<<<std::max(1, igrid), 256, 0, h->stream>>>(comb.rowptr, comb.col, comb.val, B.far.rowptr, D.near.rowptr, D.panels,
                                                        D.npanels, slices, D.inv);
  h->stats.kernel_launches += 2;
  RCG_CUDA(h, cudaGetLastError());
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));   // (panels / hop_ptr are host vectors of this scope)
  if (getenv("RCG_TIMING")) fprintf(stderr, "[rcg] dp_build: %u panels, %lld rows, %.1f MB of inverses, %.1f ms (allocations, k_dp_invert)\n",
                                    D.npanels, (long long)D.nrows, (double)inv_doubles * 8e-6,
                                    1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count());
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
  p.wpr = L.C >= 512u ? 2u : 1u;
  p.inv = D.inv;
  p.near_rp = D.near.rowptr; p.near_col = D.near.col; p.near_val = D.near.val;
  p.far_rp = a.far_rp; p.far_col = a.far_col; p.far_val = a.far_val;
  p.rhs = a.rhs; p.col_min = a.col_min; p.corr = a.corr;
  p.t0 = D.t0; p.t1 = D.t1; p.out = a.out;
  p.dotvec = a.dotvec; p.dot_partials = a.dot_partials; p.dot_limit = a.dot_limit;
  p.N = a.N; p.reversed = a.reversed;
  p.abort_g = a.abort_g;
  p.clk = (a.dbg & 1u) ? a.clk : nullptr;
  p.trace = (a.dbg & 2u) && h->clk_probe ? h->clk_probe + 16 : nullptr;
  if (p.trace) RCG_CUDA(h, cudaMemsetAsync(p.trace, 0, sizeof(unsigned long long) * RCG_DP_TRACE_WORDS, h->stream));   // (the last launch's marks remain)
  // start vector (right-hand side minus the entries of other blocks) and sentinels of this level's rows
  {
    const int64_t batches = ((int64_t)L.npanels * L.C + 31) / 32;   // 32 rows per CTA at a time
    k_dp_pre<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>(batches, (int64_t)h->sm_count * 8)), 256, 0, h->stream>>>(p);
    h->stats.kernel_launches += 1;
  }
  // grid: one CTA per SM, and no more CTAs than the first hop's near phase (4 rows per warp at a time) can use
  const int64_t tasks = std::max<int64_t>(((int64_t)a.nblocks * L.C + 4 * DP_WARPS - 1) / (4 * DP_WARPS), ((int64_t)a.nblocks * L.C + 7) / 8);
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
