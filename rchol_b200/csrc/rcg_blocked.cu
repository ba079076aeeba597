// rchol_b200 -- blocked-inverse sparse triangular solve with the rchol factor (sm_100a): the default replacement of
// the two mkl_sparse_d_trsv calls of /root/reference/c++/util/pcg.cpp:141-159.
//
// Why: the reference keeps the natural order inside the nested-dissection leaves (find_separator.cpp:116-118), so the
// dependency DAG of a leaf is ~0.2 x rows deep (5 rows per level): a row-by-row (or level-by-level) substitution is a
// chain of ~450 000 dependent hops per leaf at 256^3 / T=8.  Here the rows of a block are cut into chunks of 32 and
// the 32x32 lower-triangular diagonal block of every chunk is inverted densely at set-up, so the chain advances one
// CHUNK per hop (x_k = Winv_k * t_k, one shared-memory mat-vec by one warp), 6-7x fewer and not much longer hops.
//
// One launch per tree level.  CTA roles inside the launch (grid = groups x (1 + helpers) <= SM count, co-resident):
//   chain CTA (one per group, walks the blocks assigned to the group one after the other)
//     warp 0        critical warp : per chunk  t = t' - recent entries ; x = Winv t ; window <- x ; prog++
//                                   ALONE on its scheduler: warps 4, 8, 12 only take part in the CTA barriers (a polling
//                                   warp on the same scheduler costs the critical warp its issue slots)
//     warps 5-7, 9-11, 13-15  near helpers : chunk k -> helper k % 9, run ahead of the critical warp:
//                                   t' = start - early entries (jagged diagonals) - late entries (ELL)
//     warp 1 / 2    TMA producers : stream blob A (Winv + recent) / blob B (early + late) into two staging rings
//     warp 3        publisher     : copies solved chunks window -> out[] (vector space), fused dot product, and publishes
//                                   the block's progress counter with release semantics
//   far CTAs (`helpers` per group): start[j] = rhs[j] - sum over far entries (other blocks, or >= Dfar chunks back in
//                                   the own block); tile by tile (8 chunks), each tile as soon as the chain's published
//                                   progress covers the tile's newest far column; tile flag released to the chain CTA.
// Every spin wait is bounded (clock64 time-out -> abort flag, all waits of all CTAs then fall through): corrupt input
// or a scheduling accident cannot hang the GPU; the host turns the flag into an error.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "rcg_device.cuh"

namespace {

constexpr int BC_THREADS = 512;           // 16 warps: critical (alone on its scheduler), 2 producers, publisher, 9 near helpers
constexpr uint32_t BC_NH = 9;            // near helpers of the single-critical-warp variant (the split variant has 8)
constexpr uint32_t BC_SCR = 576;          // scratch doubles: [0,32) t vector(s), [32, 32+4*128) partial sums of the split chain
constexpr uint32_t BC_TR = 32;            // ring of t' vectors between helpers and the critical warp
constexpr uint32_t BC_TILE = 8;           // chunks per far tile of the leaf blocks (separator blocks: BlockedDev::tile_sep)
constexpr uint32_t BC_WBYTES = 1024 * 8;  // Winv, full 32x32 (zeros above the diagonal): [column pair p][row] double2
constexpr uint32_t BC_RBATCH = 3072;      // one batch of 8 recent slots: [32 rows][4 pairs] double2 values, then
                                          // [32 rows][2 halves] uint4 byte offsets into the window (row-major: the
                                          // split chain reads 8 rows x 64 B = 512 contiguous bytes per warp, no bank
                                          // conflicts; measured 16 -> 4 wavefronts per load, profiles/r01_bc_solve_ncu.md)
constexpr uint32_t BC_AHDR = 16, BC_BHDR = 80;
constexpr long long BC_TIMEOUT_CYCLES = 3000000000ll;   // ~1.5 s at 1.965 GHz
constexpr int BC_SMEM_MAX = 232448 - 1024;

__host__ __device__ __forceinline__ uint32_t r16(uint32_t v) { return (v + 15u) & ~15u; }
// A lone warp is bound by instruction issue: Winv is stored in full so that the mat-vec is 16 LDS.128 without
// predicates (ptxas turns predicated shared loads into divergent branches), conflict-free, immediate offsets only.
// byte offset of the double2 {Winv[row][2p], Winv[row][2p+1]} inside the W part of blob A
__host__ __device__ __forceinline__ uint32_t w_pair_off(uint32_t p, uint32_t row) { return 512u * p + 16u * row; }
__host__ __device__ __forceinline__ uint32_t rec_batches(uint32_t nslots) { return nslots <= 8u ? 1u : (nslots + 7u) / 8u; }

// ---- folded layout (chain_mode 5, rcg_fold.cuh) ---------------------------------------------------------------------
// The recent entries (chunks k-1 .. k-Kr) are folded with the inverse of the diagonal block at set-up:
//     x_k = Winv_k (t'_k - L_rec x_rec) = u_k - M_k x_rec ,   u_k = Winv_k t'_k ,   M_k = Winv_k L_rec  (32 x ncol, dense)
// so the chain's hop is ONE dense panel apply by one warp; the mat-vec with Winv_k moves off the chain to the near helper.
// Blob A: 16 B header {batches, rows, columns, 0} | TAIL = the 4 batches of the newest 16 columns: 4 x 4 window byte
// offsets (u32), 4 x 1024 B of panel values, batch-wise [column pair q][row] double2 | BODY = the older batches (an even
// number): offsets, then values.  Column order is ascending over body-then-tail; padding columns come FIRST (value 0,
// offset of the zero slot behind the window), so the tail always holds the newest columns at fixed offsets.
// Blob B: as before, then Winv_k as a packed lower triangle.
constexpr uint32_t FC_MINB = 4;              // tail batches (register-resident in the chain warp)
constexpr uint32_t FC_TAILB = 16u + 1040u * FC_MINB;   // header + tail: byte offset of the body
constexpr uint32_t FC_KRMAX = 8;             // fold depth limit (bitmap of 32*Kr candidate columns)
constexpr uint32_t FC_COLCAP = 32;           // panel columns per chunk beyond the previous chunk's (chunk_fold_depth)
constexpr uint32_t FC_WPACK = 4352;          // packed Winv: pair p holds rows 2p..31 -> 16 * sum(32 - 2p) bytes
__host__ __device__ __forceinline__ uint32_t fold_batches(uint32_t ncol) {   // tail + an even number of body batches
  return ncol <= 4u * FC_MINB ? FC_MINB : FC_MINB + 2u * ((ncol - 4u * FC_MINB + 7u) / 8u);
}
__host__ __device__ __forceinline__ uint32_t fold_bytesA(uint32_t ncb) { return 16u + 1040u * ncb; }
// byte offset of the double2 {Winv[row][2p], Winv[row][2p+1]} (row >= 2p) inside the packed triangle
__host__ __device__ __forceinline__ uint32_t wp_pair_off(uint32_t p, uint32_t row) { return 16u * (p * (33u - p) + row - 2u * p); }

// ---------------------------------------------------------------------------------------------------------
// memory-model primitives
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_acquire_cta_s(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_cta_s(uint32_t saddr, uint32_t v) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t *p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_volatile_u32(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_volatile_u32(uint32_t saddr, uint32_t v) {
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t saddr) {
  uint32_t v;
  asm volatile("{\n\t.reg .u16 h;\n\tld.shared.u16 h, [%1];\n\tcvt.u32.u16 %0, h;\n\t}" : "=r"(v) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ void lds_u16_if(uint32_t &v, uint32_t saddr, bool p) {
  asm volatile("{\n\t.reg .pred q;\n\t.reg .u16 h;\n\tsetp.ne.u32 q, %2, 0;\n\tmov.u16 h, 0;\n\t@q ld.shared.u16 h, [%1];\n\tcvt.u32.u16 %0, h;\n\t}"
               : "=r"(v)
               : "r"(saddr), "r"((uint32_t)p)
               : "memory");
}
__device__ __forceinline__ void sts_f64(uint32_t saddr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(saddr), "d"(v) : "memory");
}
// arrive that cannot issue before the registers a and b (results of shared-memory loads) have arrived
__device__ __forceinline__ void mbar_arrive_after(uint64_t *bar, uint32_t a, double b) {
  asm volatile(
      "{\n\t.reg .b32 lo, hi, z;\n\t"
      "mov.b64 {lo, hi}, %2;\n\t"
      "or.b32 z, lo, %1;\n\t"
      "and.b32 z, z, 0;\n\t"
      "add.u32 z, z, %0;\n\t"
      "mbarrier.arrive.shared::cta.b64 _, [z];\n\t}"
      ::"r"(smem_u32(bar)), "r"(a), "d"(b)
      : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0u;
}

__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {   // non-blocking
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0u;
}

// Bounded waiting.  guard_poll() is called after every failed poll; it returns true when the wait must be abandoned
// (this CTA or another one timed out).  Results are then garbage and the host reports RCG_ERR_CUDA.
struct Guard {
  uint32_t abort_s;          // shared-space address of the CTA's abort word
  unsigned int *abort_g;     // device-wide abort word
  uint32_t n;
  long long t0;
};
__device__ __forceinline__ bool guard_poll(Guard &G, uint32_t code) {
  ++G.n;
  if ((G.n & 7u) == 1u && lds_volatile_u32(G.abort_s) != 0u) return true;   // (first failed poll, then every 8th)
  if ((G.n & 63u) == 0u) {
    if (__ldcg(G.abort_g) != 0u) { sts_volatile_u32(G.abort_s, 1u); return true; }
    const long long now = clock64();
    if (G.n == 64u) G.t0 = now;
    else if (now - G.t0 > BC_TIMEOUT_CYCLES) {
      atomicCAS(G.abort_g, 0u, code);
      sts_volatile_u32(G.abort_s, 1u);
      return true;
    }
  }
  return false;
}
#define BC_WAIT(cond, code, sleep_ns)                    \
  do {                                                   \
    G.n = 0;                                             \
    while (!(cond)) {                                    \
      if (guard_poll(G, (code))) break;                  \
      if ((sleep_ns) > 0) __nanosleep(sleep_ns);         \
    }                                                    \
  } while (0)

// ---------------------------------------------------------------------------------------------------------
// set-up kernels
// ---------------------------------------------------------------------------------------------------------
struct BcGeom {               // blocks in ascending solve order (device arrays of nb+1 entries)
  const uint32_t *bounds, *chunk0, *tile0, *dfar;   // dfar: per block (leaves keep a large window, separators a small one)
  const uint32_t *tile;                             // chunks per far tile, per block
  const uint32_t *eblk;                             // early/late distance E, per block
  const uint32_t *krblk;                            // recent / fold depth Kr, per block (0 for warp-per-block levels)
  const uint32_t *wb;                               // 1: block of a warp-per-block level (k_wb_solve): blob A is a bare header
  const uint32_t *dpc;                              // > 0: block of a dense-panel level (rcg_dense.cuh), panel rows C: no blobs, the
                                                    // far row holds the entries of other blocks, the NEAR row (DenseDev::near, compact
  const uint32_t *dpq0;                             // row index dpq0[b] + row - lo) the own-block entries left of the row's panel
  int nb;
  uint32_t Kr, E;
  uint32_t fold;                                    // 1: folded layout (panels in blob A, Winv in blob B)
};

__device__ __forceinline__ int find_le(const uint32_t *__restrict__ a, int n, uint32_t v) {   // largest i < n with a[i] <= v
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] <= v) lo = mid; else hi = mid;
  }
  return lo;
}

// Row j (sorted by column, diagonal last) splits into contiguous runs:
//   [s, p_far) far | [p_far, p_early) early | [p_early, p_late) late | [p_late, p_rec) recent | [p_rec, p_diag) own chunk | p_diag
struct RowSplit { int64_t s, p_far, p_early, p_late, p_rec, p_diag; };

__device__ __forceinline__ RowSplit split_row(const int64_t *__restrict__ rp, const uint32_t *__restrict__ col, uint32_t j,
                                              uint32_t blo, uint32_t k, uint32_t Kr, uint32_t Dfar, uint32_t E) {
  RowSplit r;
  r.s = rp[j];
  r.p_diag = rp[j + 1] - 1;
  const int ik = (int)k;
  const uint32_t c_far = blo + 32u * (uint32_t)max(0, ik + 1 - (int)Dfar);
  const uint32_t c_early = blo + 32u * (uint32_t)max(0, ik - (int)E);
  const uint32_t c_late = blo + 32u * (uint32_t)max(0, ik - (int)Kr);
  const uint32_t c_rec = blo + 32u * k;
  int64_t p = r.s;
  while (p < r.p_diag && col[p] < c_far) p++;
  r.p_far = p;
  while (p < r.p_diag && col[p] < c_early) p++;
  r.p_early = p;
  while (p < r.p_diag && col[p] < c_late) p++;
  r.p_late = p;
  while (p < r.p_diag && col[p] < c_rec) p++;
  r.p_rec = p;
  return r;
}

__device__ __forceinline__ uint32_t warp_max_u32(uint32_t v) { return __reduce_max_sync(0xffffffffu, v); }
__device__ __forceinline__ uint32_t warp_add_u32(uint32_t v) { return __reduce_add_sync(0xffffffffu, v); }

// Fold depth of one chunk (warp-collective, folded layout): the largest D <= Kr such that the chunk's rows reference at
// most FC_COLCAP distinct columns in the D previous chunks (D >= 1 whenever there is a previous chunk).  It bounds the
// panel (blob A <= 16 + 1040 * 8 bytes: many staging slots, short chain hops); entries older than D chunks stay with the
// near helper as "late" entries.  The helper reads D from its blob's header.  ncol = columns of the panel.
__device__ __forceinline__ uint32_t chunk_fold_depth(const int64_t *__restrict__ rp, const uint32_t *__restrict__ col, uint32_t j,
                                                     bool valid, uint32_t blo, uint32_t k, uint32_t Kr, uint32_t Dfar, uint32_t E,
                                                     uint32_t *bm, uint32_t lane, uint32_t &ncol) {
  const uint32_t ng = min(k, Kr);
  if (lane < FC_KRMAX) bm[lane] = 0u;
  __syncwarp();
  if (valid && ng) {
    const RowSplit r = split_row(rp, col, j, blo, k, Kr, Dfar, E);
    const uint32_t c_late = blo + 32u * (k - ng);
    for (int64_t p = r.p_late; p < r.p_rec; p++) {
      const uint32_t lc = col[p] - c_late;
      atomicOr(&bm[lc >> 5], 1u << (lc & 31u));
    }
  }
  __syncwarp();
  uint32_t cum = 0, D = 0;
  for (uint32_t d = 1; d <= ng; d++) {   // word ng - d holds the columns of the chunk at distance d
    const uint32_t c = (uint32_t)__popc(bm[ng - d]);
    if (d > 1u && cum + c > FC_COLCAP) break;
    cum += c;
    D = d;
  }
  __syncwarp();
  ncol = cum;
  return D;
}

// first position p in [s, e) with col[p] >= c (rows are sorted by column)
__device__ __forceinline__ int64_t lower_bound_col(const uint32_t *__restrict__ col, int64_t s, int64_t e, uint32_t c) {
  while (s < e) {
    const int64_t mid = (s + e) >> 1;
    if (col[mid] < c) s = mid + 1; else e = mid;
  }
  return s;
}

// warp per chunk: blob sizes, far row lengths, far-tile requirements
__global__ void __launch_bounds__(256) k_bc_count(const int64_t *__restrict__ rp, const uint32_t *__restrict__ col, BcGeom g,
                                                  uint32_t nchunks, int64_t *__restrict__ sizeA, int64_t *__restrict__ sizeB,
                                                  int64_t *__restrict__ far_cnt, uint32_t *__restrict__ tile_need, int *err,
                                                  int64_t *__restrict__ near_cnt) {
  __shared__ uint32_t bm_all[8][FC_KRMAX];
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t wpc = blockDim.x >> 5;
  uint32_t *bm = bm_all[threadIdx.x >> 5];
  for (uint32_t gc = blockIdx.x * wpc + (threadIdx.x >> 5); gc < nchunks; gc += gridDim.x * wpc) {
    const int b = find_le(g.chunk0, g.nb, gc);
    const uint32_t k = gc - g.chunk0[b], blo = g.bounds[b], bhi = g.bounds[b + 1];
    const uint32_t j = blo + 32u * k + lane;
    uint32_t n_rec = 0, n_late = 0, n_early = 0, need = 0;
    uint32_t ncol = 0, Dk = g.krblk[b];
    if (g.dpc[b]) {   // dense-panel block: other blocks -> far row, own block left of the panel -> near row, the rest is inverted
      if (j < bhi) {
        const int64_t s = rp[j], pd = rp[j + 1] - 1;
        if (pd < s || col[pd] != j) atomicExch(err, 1);
        else {
          const int64_t a = lower_bound_col(col, s, pd, blo);
          far_cnt[j] = a - s;
          near_cnt[g.dpq0[b] + (j - blo)] = lower_bound_col(col, a, pd, blo + ((j - blo) / g.dpc[b]) * g.dpc[b]) - a;
        }
      }
      if (lane == 0) { sizeA[gc] = 0; sizeB[gc] = 0; }
      continue;
    }
    if (g.fold && !g.wb[b]) Dk = chunk_fold_depth(rp, col, j, j < bhi, blo, k, g.krblk[b], g.dfar[b], g.eblk[b], bm, lane, ncol);
    if (j < bhi) {
      const RowSplit r = split_row(rp, col, j, blo, k, Dk, g.dfar[b], g.eblk[b]);
      if (col[r.p_diag] != j) atomicExch(err, 1);
      far_cnt[j] = r.p_far - r.s;
      n_early = (uint32_t)(r.p_early - r.p_far);
      n_late = (uint32_t)(r.p_late - r.p_early);
      n_rec = (uint32_t)(r.p_rec - r.p_late);
      if (r.p_far > r.s) {
        const uint32_t c = col[r.p_far - 1];
        if (c >= blo) need = ((c - blo) >> 5) + 1u;
      }
    }
    const uint32_t nslots = warp_max_u32(n_rec), nl = warp_max_u32(n_late);
    const uint32_t ne_max = warp_max_u32(n_early), ne_tot = warp_add_u32(n_early);
    need = warp_max_u32(need);
    if (lane == 0) {
      const uint32_t bbytes = BC_BHDR + r16(ne_max) + r16(8u * ne_tot) + r16(2u * ne_tot) + 320u * nl;
      sizeA[gc] = g.wb[b] ? 16 : g.fold ? (int64_t)fold_bytesA(fold_batches(ncol)) : (int64_t)(BC_AHDR + BC_WBYTES + BC_RBATCH * rec_batches(nslots));
      sizeB[gc] = (int64_t)(bbytes + ((g.fold || g.wb[b]) ? FC_WPACK : 0u));
      if (need) atomicMax(&tile_need[g.tile0[b] + k / g.tile[b]], need);
    }
  }
}

// warp per chunk: writes blob A (Winv, recent ELL), blob B (early jagged diagonals, late ELL) and the far CSR rows
__global__ void __launch_bounds__(128) k_bc_fill(const int64_t *__restrict__ rp, const uint32_t *__restrict__ col,
                                                 const double *__restrict__ val, BcGeom g, uint32_t nchunks, uint32_t N,
                                                 int reversed, const int64_t *__restrict__ offA, const int64_t *__restrict__ offB,
                                                 unsigned char *__restrict__ blobA, unsigned char *__restrict__ blobB,
                                                 const int64_t *__restrict__ far_rp, uint32_t *__restrict__ far_col,
                                                 double *__restrict__ far_val, uint32_t *__restrict__ far_split,
                                                 const int64_t *__restrict__ near_rp, uint32_t *__restrict__ near_col,
                                                 double *__restrict__ near_val) {
  __shared__ double Wm_all[4][32][33];
  __shared__ uint32_t bm_all[4][FC_KRMAX];
  const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double(*Wm)[33] = Wm_all[wib];
  const uint32_t wpc = blockDim.x >> 5;
  for (uint32_t gc = blockIdx.x * wpc + wib; gc < nchunks; gc += gridDim.x * wpc) {
    const int b = find_le(g.chunk0, g.nb, gc);
    const uint32_t k = gc - g.chunk0[b], blo = g.bounds[b], bhi = g.bounds[b + 1];
    const uint32_t j = blo + 32u * k + lane;
    const bool valid = j < bhi;
    if (g.dpc[b]) {   // dense-panel block: the far and near rows only (their lengths were counted by k_bc_count)
      if (valid) {
        int64_t o = far_rp[j], p = rp[j];
        const int64_t n = far_rp[j + 1] - o;
        for (const int64_t pe = p + n; p < pe; p++, o++) {
          far_col[o] = reversed ? N - 1u - col[p] : col[p];
          far_val[o] = val[p];
        }
        far_split[j] = (uint32_t)n;
        const uint32_t q = g.dpq0[b] + (j - blo);
        o = near_rp[q];
        for (const int64_t pe = p + (near_rp[q + 1] - o); p < pe; p++, o++) {
          near_col[o] = reversed ? N - 1u - col[p] : col[p];
          near_val[o] = val[p];
        }
      }
      continue;
    }
    const uint32_t nr = min(32u, bhi - (blo + 32u * k));
    RowSplit r;
    r.s = r.p_far = r.p_early = r.p_late = r.p_rec = r.p_diag = 0;
    const uint32_t wmask = 32u * g.dfar[b] - 1u;
    uint32_t ncol_fold = 0, Dk = g.krblk[b];
    if (g.fold && !g.wb[b]) Dk = chunk_fold_depth(rp, col, j, valid, blo, k, g.krblk[b], g.dfar[b], g.eblk[b], bm_all[wib], lane, ncol_fold);
    if (valid) r = split_row(rp, col, j, blo, k, Dk, g.dfar[b], g.eblk[b]);
    const uint32_t n_early = (uint32_t)(r.p_early - r.p_far), n_late = (uint32_t)(r.p_late - r.p_early);
    const uint32_t n_rec = (uint32_t)(r.p_rec - r.p_late), n_diag = (uint32_t)(r.p_diag - r.p_rec);
    const uint32_t nslots = warp_max_u32(n_rec), nl = warp_max_u32(n_late);
    const uint32_t ne_max = warp_max_u32(n_early), ne_tot = warp_add_u32(n_early);
    unsigned char *A = blobA + offA[gc];
    unsigned char *B = blobB + offB[gc];

    // ---- far rows (vector-space columns, raw values) ---------------------------------------------------
    if (valid) {
      int64_t o = far_rp[j];
      uint32_t n_other = 0;   // leading entries whose column belongs to another block (sorted rows: they come first)
      for (int64_t p = r.s; p < r.p_far; p++, o++) {
        const uint32_t c = col[p];
        n_other += c < blo ? 1u : 0u;
        far_col[o] = reversed ? N - 1u - c : c;
        far_val[o] = val[p];
      }
      far_split[j] = n_other;
    }
    // ---- Winv: row i of the inverse of the chunk's diagonal block, all columns in parallel ----------------
    const double dii_mine = valid ? val[r.p_diag] : 1.0;
    for (uint32_t i = 0; i < 32u; i++) {
      const uint32_t ni = __shfl_sync(0xffffffffu, n_diag, (int)i);
      const long long p0 = __shfl_sync(0xffffffffu, (long long)r.p_rec, (int)i);
      const double dii = __shfl_sync(0xffffffffu, dii_mine, (int)i);
      double s = lane == i ? 1.0 : 0.0;
      if (i < nr) {
        for (uint32_t e = 0; e < ni; e++) {
          const uint32_t m = col[p0 + e] - blo - 32u * k;   // < i
          s = fma(-val[p0 + e], Wm[m][lane], s);
        }
        s = s / dii;
      }
      Wm[i][lane] = s;
      __syncwarp();
    }
    if (g.wb[b]) {
      // ---- warp-per-block level: no chain, no panel; the warp that owns the block applies everything ---------------
      if (lane == 0) {
        uint32_t *hd = reinterpret_cast<uint32_t *>(A);
        hd[0] = 0; hd[1] = nr; hd[2] = 0; hd[3] = 0;
      }
    } else if (!g.fold) {
      unsigned char *Wp = A + BC_AHDR;
      for (uint32_t pp = 0; pp < 16u; pp++) {
        double *dst = reinterpret_cast<double *>(Wp + w_pair_off(pp, lane));
        dst[0] = Wm[lane][2u * pp];
        dst[1] = Wm[lane][2u * pp + 1u];
      }
      if (lane == 0) {
        uint32_t *hd = reinterpret_cast<uint32_t *>(A);
        hd[0] = rec_batches(nslots); hd[1] = nr; hd[2] = nslots; hd[3] = 0;
      }
      __syncwarp();
      // ---- recent entries: batches of 8 slots; values [32 rows][4 pairs] double2, window byte offsets
      //      [32 rows][2 halves] uint4.  Padding: value 0, offset of the slot behind the window that always holds 0.0
      const uint32_t nbt = rec_batches(nslots);
      for (uint32_t bt = 0; bt < nbt; bt++) {
        unsigned char *R = A + BC_AHDR + BC_WBYTES + (size_t)BC_RBATCH * bt;
        for (uint32_t u = 0; u < 8u; u++) {
          const uint32_t sidx = 8u * bt + u;
          const bool have = sidx < n_rec;
          reinterpret_cast<double *>(R + 64u * lane + 16u * (u >> 1))[u & 1u] = have ? val[r.p_late + sidx] : 0.0;
          reinterpret_cast<uint32_t *>(R + 2048u + 32u * lane + 16u * (u >> 2))[u & 3u] =
              8u * (have ? ((col[r.p_late + sidx] - blo) & wmask) : (wmask + 1u));
        }
      }
    } else {
      // ---- folded panel M = Winv * L_rec, one dense column per distinct recent column (ascending) ----------
      const uint32_t c_late = blo + 32u * (uint32_t)max(0, (int)k - (int)Dk);
      uint32_t *bm = bm_all[wib];
      if (lane < FC_KRMAX) bm[lane] = 0u;
      __syncwarp();
      if (valid)
        for (int64_t p = r.p_late; p < r.p_rec; p++) {
          const uint32_t lc = col[p] - c_late;
          atomicOr(&bm[lc >> 5], 1u << (lc & 31u));
        }
      __syncwarp();
      const uint32_t ncol = warp_add_u32(lane < FC_KRMAX ? (uint32_t)__popc(bm[lane]) : 0u);
      const uint32_t ncb = fold_batches(ncol);
      // column slot ci (0 .. 4 ncb - 1, padding first): batch = ci / 4; body batches come first in column order
      const uint32_t nbody = ncb - FC_MINB, npad = 4u * ncb - ncol;
      auto off_ptr = [&](uint32_t ci_) -> uint32_t * {
        const uint32_t bt = ci_ >> 2;
        unsigned char *o = bt < nbody ? A + FC_TAILB + 16u * bt : A + 16u + 16u * (bt - nbody);
        return reinterpret_cast<uint32_t *>(o) + (ci_ & 3u);
      };
      auto val_ptr = [&](uint32_t ci_, uint32_t row_) -> double * {
        const uint32_t bt = ci_ >> 2;
        unsigned char *v = bt < nbody ? A + FC_TAILB + 16u * nbody + 1024u * bt : A + 16u + 16u * FC_MINB + 1024u * (bt - nbody);
        return reinterpret_cast<double *>(v + 512u * ((ci_ >> 1) & 1u) + 16u * row_) + (ci_ & 1u);
      };
      if (lane == 0) {
        uint32_t *hd = reinterpret_cast<uint32_t *>(A);
        hd[0] = ncb; hd[1] = nr; hd[2] = ncol; hd[3] = 0;
      }
      for (uint32_t i = lane; i < npad; i += 32u) *off_ptr(i) = 8u * (wmask + 1u);   // padding columns: the zero slot
      __syncwarp();
      int64_t pcur = r.p_late;   // this row's next recent entry (rows are sorted by column)
      uint32_t ci = npad;
      for (uint32_t wd = 0; wd < FC_KRMAX; wd++) {
        uint32_t bits = bm[wd];
        while (bits) {
          const uint32_t bpos = (uint32_t)__ffs((int)bits) - 1u;
          bits &= bits - 1u;
          const uint32_t c = c_late + 32u * wd + bpos;
          double lval = 0.0;
          if (valid && pcur < r.p_rec && col[pcur] == c) { lval = val[pcur]; pcur++; }
          uint32_t nz = __ballot_sync(0xffffffffu, lval != 0.0);
          double m = 0.0;
          while (nz) {
            const int i = __ffs((int)nz) - 1;
            nz &= nz - 1u;
            m = fma(Wm[lane][i], __shfl_sync(0xffffffffu, lval, i), m);   // Winv[row][i] * L[i][c]
          }
          *val_ptr(ci, lane) = m;
          if (lane == 0) *off_ptr(ci) = 8u * ((c - blo) & wmask);
          ci++;
        }
      }
    }
    // ---- blob B ----------------------------------------------------------------------------------------
    {
      uint32_t rank = 0;
      for (uint32_t l = 0; l < 32u; l++) {
        const uint32_t o = __shfl_sync(0xffffffffu, n_early, (int)l);
        rank += (o > n_early || (o == n_early && l < lane)) ? 1u : 0u;
      }
      if (lane == 0) {
        uint32_t *hd = reinterpret_cast<uint32_t *>(B);
        hd[0] = ne_max; hd[1] = ne_tot; hd[2] = nl; hd[3] = Dk;   // Dk: fold depth of the chunk (the helper's late wait)
      }
      B[16u + rank] = (unsigned char)lane;   // perm: sorted position -> row
      B[48u + lane] = (unsigned char)rank;   // rank: row -> sorted position
      unsigned char *cnt = B + BC_BHDR;
      double *ev = reinterpret_cast<double *>(B + BC_BHDR + r16(ne_max));
      uint16_t *ec = reinterpret_cast<uint16_t *>(B + BC_BHDR + r16(ne_max) + r16(8u * ne_tot));
      uint32_t base = 0;
      for (uint32_t s = 0; s < ne_max; s++) {
        const uint32_t cs = (uint32_t)__popc(__ballot_sync(0xffffffffu, n_early > s));
        if (lane == 0) cnt[s] = (unsigned char)cs;
        if (s < n_early) {
          ev[base + rank] = val[r.p_far + s];
          ec[base + rank] = (uint16_t)((col[r.p_far + s] - blo) & wmask);
        }
        base += cs;
      }
      double *lv = reinterpret_cast<double *>(B + BC_BHDR + r16(ne_max) + r16(8u * ne_tot) + r16(2u * ne_tot));
      uint16_t *lc = reinterpret_cast<uint16_t *>(reinterpret_cast<unsigned char *>(lv) + 256u * nl);
      for (uint32_t s = 0; s < nl; s++) {
        const bool have = s < n_late;
        lv[s * 32u + lane] = have ? val[r.p_early + s] : 0.0;
        lc[s * 32u + lane] = have ? (uint16_t)((col[r.p_early + s] - blo) & wmask) : (uint16_t)(wmask + 1u);
      }
      if (g.fold || g.wb[b]) {   // Winv, packed lower triangle, behind the late entries (near helper / the block's warp)
        unsigned char *Wq = reinterpret_cast<unsigned char *>(lv) + 320u * nl;
        for (uint32_t pp = 0; pp < 16u; pp++)
          if (lane >= 2u * pp) {
            double *dst = reinterpret_cast<double *>(Wq + wp_pair_off(pp, lane));
            dst[0] = Wm[lane][2u * pp];
            dst[1] = Wm[lane][2u * pp + 1u];
          }
      }
    }
    __syncwarp();
  }
}

__global__ void k_gather_i64(const int64_t *__restrict__ src, const uint32_t *__restrict__ idx, int n, int64_t *__restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

// ---------------------------------------------------------------------------------------------------------
// the solve kernel
// ---------------------------------------------------------------------------------------------------------
struct BcArgs {
  const BcBlock *blocks;       // blocks of this level
  uint32_t nblocks, ngroups, helpers;
  const int64_t *offA, *offB;
  const unsigned char *blobA, *blobB;
  const int64_t *far_rp;
  const uint32_t *far_col;
  const double *far_val;
  const uint32_t *tile_need;
  const uint32_t *far_split;   // per row: leading far entries whose column belongs to another block
  uint32_t *tileflag, *gprog;
  double *w;
  const double *rhs;
  double *out;
  const double *dotvec;        // nullable
  double *dot_partials;        // nullable: one slot per block (gidx)
  uint32_t dot_limit;
  uint32_t N;
  int reversed;
  uint32_t Kr, E, Dfar, W;     // W = 32*Dfar window rows (power of two)
  uint32_t SA, SB, capA, capB;
  uint32_t tile;               // chunks per far tile of this level's blocks
  uint32_t far_lpr;            // lanes per far row, pass 0 (entries of other blocks): 8 or 32
  uint32_t far_lpr2;           // lanes per far row, pass 1 (entries >= window back in the own block): 8 or 32
  uint32_t col_min;            // multi-GPU top separators: far columns below col_min are left out ...
  const double *corr;          // ... their sum over all ranks arrives here (indexed by vector index - col_min)
  unsigned int *abort_g;
  uint32_t *ticket;            // warp-per-block levels: next block of the level (zeroed with the flags before the launch)
  unsigned long long *clk;     // nullable diagnostics
  uint32_t dbg;                // bit 0: cycle profile of the critical warp and of near helper 0 (CTA 0) into clk[3..12]
};

// CW = 1: one critical warp (first version, kept selectable: chain_mode 3).  CW = 4: the chain's per-chunk work is split
// over four warps, one per scheduler (see "split chain" below) -- a lone warp is bound by instruction issue.
// PROF: cycle counters of the critical warp(s) and of near helper 0 (dbg bit 0); a separate instantiation so that the
// production chain loop stays straight-line code (a lone warp pays a fetch bubble for every taken branch).
template <int CW, bool PROF>
__global__ void __launch_bounds__(BC_THREADS, 1) k_bc_solve(const BcArgs P) {
  constexpr uint32_t W_PRODA = CW == 1 ? 1u : 5u, W_PRODB = CW == 1 ? 2u : 6u, W_PUB = CW == 1 ? 3u : 7u;
  constexpr uint32_t NH = CW == 1 ? BC_NH : 8u;
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t per = 1u + P.helpers;
  const uint32_t grp = blockIdx.x / per, role = blockIdx.x % per;

  // shared-memory carve-up (chain CTAs; far CTAs only use the control words)
  double *win = reinterpret_cast<double *>(smem);
  double *tprime = win + P.W + 16;   // win[W] holds 0.0: the slot padding entries point at
  double *scratch = tprime + BC_TR * 32u;
  unsigned char *ringA = reinterpret_cast<unsigned char *>(scratch + BC_SCR);
  unsigned char *ringB = ringA + (size_t)P.SA * P.capA;
  uint64_t *fullA = reinterpret_cast<uint64_t *>(ringB + (size_t)P.SB * P.capB);
  uint64_t *emptyA = fullA + P.SA;
  uint64_t *fullB = emptyA + P.SA;
  uint64_t *emptyB = fullB + P.SB;
  unsigned long long *ptrB = reinterpret_cast<unsigned long long *>(emptyB + P.SB);
  uint32_t *seqB = reinterpret_cast<uint32_t *>(ptrB + P.SB);   // running chunk index whose blob the slot holds
  uint32_t *tready = seqB + P.SB;
  uint32_t *ctl = tready + BC_TR;   // [0] prog: solved chunks of the current block, [1] published chunks, [2] abort

  if (threadIdx.x < P.SA) { mbar_init(fullA + threadIdx.x, 1); mbar_init(emptyA + threadIdx.x, CW); }
  if (threadIdx.x < BC_SCR) scratch[threadIdx.x] = 0.0;
  if (threadIdx.x < P.SB) { mbar_init(fullB + threadIdx.x, 1); mbar_init(emptyB + threadIdx.x, 1); seqB[threadIdx.x] = 0xFFFFFFFFu; }
  if (threadIdx.x == 0) { ctl[0] = 0; ctl[1] = 0; ctl[2] = 0; }
  if (threadIdx.x < 16) win[P.W + threadIdx.x] = 0.0;
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  Guard G;
  G.abort_s = smem_u32(ctl + 2);
  G.abort_g = P.abort_g;
  G.n = 0;
  G.t0 = 0;
  const uint32_t prog_s = smem_u32(ctl), pub_s = smem_u32(ctl + 1);

  if (role > 0) {
    // =========================== far CTA: start vector of the chain, tile by tile ===========================
    // Pass 1 (never waits): entries whose column belongs to another, already solved block -- the bulk of a separator's
    // far part.  Tiles without far entries inside the own block are complete after it and are released at once.
    // Pass 2: entries >= Dfar chunks back in the own block, tile by tile as the chain's published progress allows.
    const uint32_t hid = role - 1u;
    // lanes per row (8, or 32 for the long rows of separators) and rows per warp, per pass.  Pass 1 sits on the chain's
    // critical path (its tile can start only when the chain is Dfar chunks away and must finish before the chain arrives;
    // measured at 256^3: 16 rows per warp one after the other, ~2 us of dependent HBM round trips each = 32 us per
    // tile, which bounded the separator levels at 2.2 us per chunk), so it uses few lanes per row and many rows at once.
    const uint32_t lpr_pass[2] = {P.far_lpr, P.far_lpr2};
    for (uint32_t bi = grp; bi < P.nblocks; bi += P.ngroups) {
      const BcBlock b = P.blocks[bi];
      const uint32_t nch = (b.hi - b.lo + 31u) >> 5, ntile = (nch + P.tile - 1u) / P.tile;
      // own tiles t_i = hid + i*helpers; pass 1 of tile i runs after pass 0 of tile i + LA: the other-block work of the
      // next tiles is done before this CTA waits for the chain
      constexpr uint32_t LA = 2;
      const uint32_t nown = hid < ntile ? (ntile - hid + P.helpers - 1u) / P.helpers : 0u;
      for (uint32_t it = 0; it < nown + LA; it++) {
        for (uint32_t pass = 0; pass < 2u; pass++) {
          if (pass == 0u ? it >= nown : it < LA) continue;
          const uint32_t t = hid + (pass == 0u ? it : it - LA) * P.helpers;
          const uint32_t need = P.tile_need[b.tile0 + t];
          if (pass == 1u) {
            if (need == 0u) continue;
            if (threadIdx.x == 0) BC_WAIT(ld_acquire_gpu(P.gprog + b.gidx) >= need, 0x100u, 200);
            __syncthreads();
          }
          const uint32_t lpr = lpr_pass[pass], rpw = 32u / lpr, sub = lane & (lpr - 1u);
          const uint32_t r0 = b.lo + t * (32u * P.tile), r1 = min(b.hi, r0 + 32u * P.tile);
          for (uint32_t base = r0 + warp * rpw; base < r1; base += (BC_THREADS / 32) * rpw) {
            const uint32_t j = base + lane / lpr;
            const bool valid = j < r1;
            double acc = 0.0;
            if (valid) {
              const int64_t es = P.far_rp[j] + P.far_split[j];   // [rp, es) other blocks, [es, rp1) own block
              const int64_t e0 = pass == 0u ? P.far_rp[j] : es, e1 = pass == 0u ? es : P.far_rp[j + 1];
              double acc1 = 0.0;
              int64_t e = e0 + sub;
              for (; e + lpr < e1; e += 2u * lpr) {
                const uint32_t c = P.far_col[e], c2 = P.far_col[e + lpr];
                if (c >= P.col_min) acc = fma(P.far_val[e], __ldcg(P.out + c), acc);
                if (c2 >= P.col_min) acc1 = fma(P.far_val[e + lpr], __ldcg(P.out + c2), acc1);
              }
              if (e < e1) {
                const uint32_t c = P.far_col[e];
                if (c >= P.col_min) acc = fma(P.far_val[e], __ldcg(P.out + c), acc);
              }
              acc += acc1;
            }
            if (lpr == 32u) {
              acc += __shfl_xor_sync(0xffffffffu, acc, 16);
              acc += __shfl_xor_sync(0xffffffffu, acc, 8);
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            if (valid && sub == 0u) {
              double s;
              if (pass == 0u) {
                const uint32_t i = P.reversed ? P.N - 1u - j : j;
                s = P.rhs[i];
                if (P.corr) s -= P.corr[i - P.col_min];
              } else {
                s = __ldcg(P.w + j);   // written by this very thread in pass 0
              }
              __stcg(P.w + j, s - acc);
            }
          }
          if (pass == 1u || need == 0u) {
            __syncthreads();
            if (threadIdx.x == 0) {
              __threadfence();
              st_release_gpu(P.tileflag + b.tile0 + t, 1u);
            }
          }
        }
      }
    }
    return;
  }

  // ================================= chain CTA ============================================================
  const uint32_t wmask = P.W - 1u;
  const uint32_t win_s = smem_u32(win), tp_s = smem_u32(tprime), sc_s = smem_u32(scratch), trdy_s = smem_u32(tready);
  uint32_t ia0 = 0;   // chunks of the blocks this CTA has finished: running index of the staging rings
  long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // profiling (dbg bit 0), critical warp of CTA 0: cycles waiting for blob A / for t' /
                                          // recent entries / chunks / mat-vec / store + release
  long long ph[5] = {0, 0, 0, 0, 0};   // near helper 0 of CTA 0: tile flag + start vector, blob B, early, wait for prog, late
  long long clk0 = 0;
  if (P.clk && blockIdx.x == 0 && threadIdx.x == 0) clk0 = clock64();

  for (uint32_t bi = grp; bi < P.nblocks; bi += P.ngroups) {
    const BcBlock b = P.blocks[bi];
    const uint32_t nch = (b.hi - b.lo + 31u) >> 5;
    if (threadIdx.x < BC_TR) tready[threadIdx.x] = 0u;
    if (threadIdx.x == 0) { ctl[0] = 0u; ctl[1] = 0u; }
    __syncthreads();

    if (CW == 1 && warp == 0) {
      // ------------------------------ critical warp ------------------------------------------------------
      // Measured (scripts/ubench/mv.cu): a lone warp pays ~240 cycles for the 32x32 mat-vec and ~500 for the whole
      // chunk when nothing else gets in its way; every exposed shared-memory round trip adds 50-100.  So the loop is
      // software-pipelined: the loads of chunk k+1 that do not depend on the chain (half of Winv, the first batch of
      // recent entries) are issued right after the mat-vec FMAs of chunk k, and their latency hides behind the
      // reduction, the window store, the release and the poll of the next t'.
      const bool prof = PROF && (P.dbg & 1u) != 0u && blockIdx.x == 0;
      constexpr uint32_t WPRE = 6;   // column pairs of Winv held in registers ahead of time
      uint32_t slot = ia0 % P.SA, par = (ia0 / P.SA) & 1u;
      double w[2 * WPRE];
      uint32_t o0, o1, o2, o3, o4, o5, o6, o7, nbt;
      double v0, v1, v2, v3, v4, v5, v6, v7;
      uint32_t w_s, r_s;
#define BC_PRELOAD()                                                                                                      \
      do {                                                                                                                \
        const uint32_t a_s_ = smem_u32(ringA + (size_t)slot * P.capA);                                                    \
        w_s = a_s_ + BC_AHDR + 16u * lane;                                                                                \
        r_s = a_s_ + BC_AHDR + BC_WBYTES + 64u * lane; /* row-major recent batch: 64 B of values per row */                                                                                           \
        _Pragma("unroll") for (uint32_t pp = 0; pp < WPRE; pp++)                                                          \
          asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(w[2 * pp]), "=d"(w[2 * pp + 1]) : "r"(w_s + 512u * pp) : "memory"); \
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(o0), "=r"(o1), "=r"(o2), "=r"(o3) : "r"(r_s + 2048u - 32u * lane) : "memory"); \
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(o4), "=r"(o5), "=r"(o6), "=r"(o7) : "r"(r_s + 2048u - 32u * lane + 16u) : "memory"); \
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(r_s) : "memory");                      \
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v2), "=d"(v3) : "r"(r_s + 16u) : "memory");               \
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v4), "=d"(v5) : "r"(r_s + 32u) : "memory");              \
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v6), "=d"(v7) : "r"(r_s + 48u) : "memory");              \
        nbt = lds_u32(a_s_);                                                                                              \
      } while (0)
      if (nch > 0) {
        BC_WAIT(mbar_try(fullA + slot, par), 0x200u, 0);
        BC_PRELOAD();
      }
      for (uint32_t k = 0; k < nch; k++) {
        const uint32_t ts = k & (BC_TR - 1u);
        long long c1 = 0;
        if (prof) c1 = clock64();
        // t' and its ready flag are read together; the value is only used when the flag (read first) was set
        uint32_t rdy = ld_acquire_cta_s(trdy_s + 4u * ts);
        double t = lds_f64(tp_s + 8u * (ts * 32u + lane));
        if (rdy != k + 1u) {
          BC_WAIT(ld_acquire_cta_s(trdy_s + 4u * ts) == k + 1u, 0x300u, 0);
          t = lds_f64(tp_s + 8u * (ts * 32u + lane));
          if (prof) pc[6] += 1;
        }
        long long c2 = 0;
        if (prof) c2 = clock64();
        double t1, t2, t3;
        {
          const double x0 = lds_f64(win_s + o0), x1 = lds_f64(win_s + o1), x2 = lds_f64(win_s + o2), x3 = lds_f64(win_s + o3);
          const double x4 = lds_f64(win_s + o4), x5 = lds_f64(win_s + o5), x6 = lds_f64(win_s + o6), x7 = lds_f64(win_s + o7);
          t = fma(-v0, x0, t);
          t1 = -v1 * x1;
          t2 = -v2 * x2;
          t3 = -v3 * x3;
          t = fma(-v4, x4, t);
          t1 = fma(-v5, x5, t1);
          t2 = fma(-v6, x6, t2);
          t3 = fma(-v7, x7, t3);
        }
        for (uint32_t bt = 1; bt < nbt; bt++) {   // more than 8 recent slots
          const uint32_t q_s = r_s + BC_RBATCH * bt;
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(o0), "=r"(o1), "=r"(o2), "=r"(o3) : "r"(q_s + 2048u - 32u * lane) : "memory");
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(o4), "=r"(o5), "=r"(o6), "=r"(o7) : "r"(q_s + 2048u - 32u * lane + 16u) : "memory");
          asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(q_s) : "memory");
          asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v2), "=d"(v3) : "r"(q_s + 16u) : "memory");
          asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v4), "=d"(v5) : "r"(q_s + 32u) : "memory");
          asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v6), "=d"(v7) : "r"(q_s + 48u) : "memory");
          const double x0 = lds_f64(win_s + o0), x1 = lds_f64(win_s + o1), x2 = lds_f64(win_s + o2), x3 = lds_f64(win_s + o3);
          const double x4 = lds_f64(win_s + o4), x5 = lds_f64(win_s + o5), x6 = lds_f64(win_s + o6), x7 = lds_f64(win_s + o7);
          t = fma(-v0, x0, t);
          t1 = fma(-v1, x1, t1);
          t2 = fma(-v2, x2, t2);
          t3 = fma(-v3, x3, t3);
          t = fma(-v4, x4, t);
          t1 = fma(-v5, x5, t1);
          t2 = fma(-v6, x6, t2);
          t3 = fma(-v7, x7, t3);
        }
        t = (t + t1) + (t2 + t3);
        sts_f64(sc_s + 8u * lane, t);
        __syncwarp();
        long long c3 = 0;
        if (prof) c3 = clock64();
        // staging slot of the next chunk: test its barrier now (non-blocking), the answer is needed a mat-vec later
        uint32_t nslot = slot + 1u, npar = par;
        if (nslot == P.SA) { nslot = 0; npar ^= 1u; }
        const bool more = k + 1u < nch;
        bool a_ready = more && mbar_test(fullA + nslot, npar);
        double a0, a1, a2, a3;
        {
          double x0, x1, y0, y1;
          asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x0), "=d"(x1) : "r"(sc_s) : "memory");
          a0 = w[0] * x0;
          a1 = w[1] * x1;
          asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x0), "=d"(x1) : "r"(sc_s + 16u) : "memory");
          a2 = w[2] * x0;
          a3 = w[3] * x1;
#pragma unroll
          for (uint32_t pp = 2; pp < WPRE; pp += 2u) {
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x0), "=d"(x1) : "r"(sc_s + 16u * pp) : "memory");
            a0 = fma(w[2 * pp], x0, a0);
            a1 = fma(w[2 * pp + 1], x1, a1);
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x0), "=d"(x1) : "r"(sc_s + 16u * pp + 16u) : "memory");
            a2 = fma(w[2 * pp + 2], x0, a2);
            a3 = fma(w[2 * pp + 3], x1, a3);
          }
#pragma unroll
          for (uint32_t pp = WPRE; pp < 16u; pp += 2u) {
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(y0), "=d"(y1) : "r"(w_s + 512u * pp) : "memory");
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x0), "=d"(x1) : "r"(sc_s + 16u * pp) : "memory");
            a0 = fma(y0, x0, a0);
            a1 = fma(y1, x1, a1);
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(y0), "=d"(y1) : "r"(w_s + 512u * pp + 512u) : "memory");
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x0), "=d"(x1) : "r"(sc_s + 16u * pp + 16u) : "memory");
            a2 = fma(y0, x0, a2);
            a3 = fma(y1, x1, a3);
          }
        }
        long long c4 = 0;
        if (prof) c4 = clock64();
        // next chunk's chain-independent loads (same registers: the mat-vec above has issued its last use of them)
        const uint32_t oslot = slot;
        if (more) {
          slot = nslot;
          par = npar;
          if (!a_ready) {
            long long c0 = 0;
            if (prof) c0 = clock64();
            BC_WAIT(mbar_try(fullA + slot, par), 0x200u, 0);
            if (prof) pc[0] += clock64() - c0;
          }
          BC_PRELOAD();
        }
        const double x = (a0 + a1) + (a2 + a3);
        sts_f64(win_s + 8u * ((32u * k + lane) & wmask), x);
        __syncwarp();
        if (lane == 0) {
          st_release_cta_s(prog_s, k + 1u);
          mbar_arrive(emptyA + oslot);
        }
        if (prof) { pc[1] += c2 - c1; pc[2] += c3 - c2; pc[3] += 1; pc[4] += c4 - c3; pc[5] += clock64() - c4; }
      }
#undef BC_PRELOAD
    } else if (CW == 4 && warp < 4u) {
      // ------------------------------ split chain: four critical warps ------------------------------------
      // Warp cw (one per scheduler) owns rows 8cw..8cw+7 of the chunk for the recent entries (four lanes per row, two of
      // the eight slots each) and columns 8cw..8cw+7 of Winv for the mat-vec (lane = row of the result):
      //   gather x of the two previous chunks  ->  t[row] = t' - recent  (4-lane shuffle reduction)
      //   -> t[8cw..8cw+7] broadcast through 64 bytes of shared memory -> 8 DFMA with the warp's slice of Winv
      //   -> partial[k & 3][cw][lane]  -> named barrier (4 critical warps + scribe)
      // x_k is never materialised by the critical warps: a gather of x_k[c] reads the four partial sums and adds them in
      // the same order as the scribe does when it writes the window.  Four partial buffers per warp indexed by the chunk
      // number mod 4, laid out [warp][chunk mod 4][row] so that the low 10 bits of the window byte offset the set-up
      // stored with the entry ARE the offset inside a warp's partials (one LOP3 per gather: the lone warp is issue-bound);
      // chunk k reads the buffers of k-1 and k-2 while a faster warp may already write that of k+1; one barrier per chunk.
      // Measured (ncu source page, profiles/): a lone warp pays ~4 cycles per dependent instruction, a fetch bubble per
      // taken branch and 30-130 cycles per shared-memory round trip, so the loop is straight-line code, every wait is a
      // compact bounded spin, and the loads of chunk k+1 that do not depend on the chain (Winv slice, recent values and
      // offsets, t' and its flag, the staging barrier's phase) are issued before the barrier of chunk k.
      const bool prof = PROF && (P.dbg & 1u) != 0u && blockIdx.x == 0 && warp == 0u;
      const uint32_t cw = warp, rowg = 8u * cw + (lane >> 2), sub = lane & 3u;
      const uint32_t part_s = sc_s + 256u, tb_s = sc_s + 64u * cw;
      const uint32_t pw_s = part_s + 1024u * cw + 8u * lane;   // this lane's slot in buffer 0 of this warp's partial sums
      const uint32_t rec_off = 64u * rowg + 16u * sub, off_off = 2048u + 32u * rowg + 8u * sub;
      const uint32_t fullA_s = smem_u32(fullA), emptyA_s = smem_u32(emptyA), ringA_s = smem_u32(ringA);
      uint32_t slot = ia0 % P.SA, par = (ia0 / P.SA) & 1u;
      double wa[8];
      double rv0 = 0.0, rv1 = 0.0, tpv = 0.0;
      uint32_t ro0 = 0u, ro1 = 0u, nbt = 1u, a_s = 0u, tpf = 0u, a_ok = 0u;
      uint32_t spins = 0;
      // bounded spin without the guard's bookkeeping (the time-out is counted in polls: ~0.5 s)
#define BC4_SPIN(cond_, code_)                                                                                            \
      while (__builtin_expect(!(cond_), 0)) {                                                                             \
        if (++spins > (1u << 24)) { atomicCAS(P.abort_g, 0u, (code_)); sts_volatile_u32(G.abort_s, 1u); break; }          \
      }
#define BC4_MBAR_TEST(ok_, bar_s_, par_)                                                                                  \
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" \
                   : "=r"(ok_) : "r"(bar_s_), "r"(par_) : "memory")
      // recent values / offsets / batch count of the chunk staged in `slot`
#define BC4_LOAD_REC()                                                                                                    \
      do {                                                                                                                \
        a_s = ringA_s + slot * P.capA;                                                                                    \
        const uint32_t r_s_ = a_s + BC_AHDR + BC_WBYTES;                                                                  \
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(rv0), "=d"(rv1) : "r"(r_s_ + rec_off) : "memory");         \
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ro0), "=r"(ro1) : "r"(r_s_ + off_off) : "memory");         \
        nbt = lds_u32(a_s);                                                                                               \
      } while (0)
#define BC4_LOAD_W(w_)                                                                                                    \
      do {                                                                                                                \
        const uint32_t w_s_ = a_s + BC_AHDR + 2048u * cw + 16u * lane;                                                    \
        _Pragma("unroll") for (uint32_t pp = 0; pp < 4u; pp++)                                                            \
          asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(w_[2 * pp]), "=d"(w_[2 * pp + 1]) : "r"(w_s_ + 512u * pp) : "memory"); \
      } while (0)
      // flag of t' of chunk k_, not waited for: it is tested at the top of the next chunk, and the value is loaded there
      // through an address that depends on the flag (so neither the compiler nor ptxas can hoist it above the flag)
#define BC4_LOAD_TP(k_)                                                                                                   \
      do {                                                                                                                \
        tpf = lds_volatile_u32(trdy_s + 4u * ((k_) & (BC_TR - 1u)));                                                      \
      } while (0)
      // gather x[c] of one of the two previous chunks from the partial sums (buffer = chunk number mod 4)
      // The chain's per-chunk latency is the sum of what a lone, issue-bound warp does between two barriers (ncu source
      // page, profiles/r01_bc_solve_ncu.md), so the order of the statements below is the schedule: every shared-memory
      // load is issued as early as its address allows (gathers, Winv slice, t'), the staging-ring bookkeeping sits where
      // the warp waits for the t broadcast anyway, and nothing but two LOP3 separates the barrier from the first gather.
#define BC4_GATHER_LD(p_, ro_)                                                                                            \
      do {                                                                                                                \
        const uint32_t g_ = part_s + ((ro_) & 0x3F8u);   /* [warp][chunk mod 4][row]: the window offset's low bits */    \
        p_[0] = lds_f64(g_); p_[1] = lds_f64(g_ + 1024u); p_[2] = lds_f64(g_ + 2048u); p_[3] = lds_f64(g_ + 3072u);       \
      } while (0)
#define BC4_CHUNK(MORE_)                                                                                                  \
      do {                                                                                                                \
        constexpr bool more_ = MORE_;                                                                                     \
        long long c1_ = 0;                                                                                                \
        if (prof) c1_ = clock64();                                                                                        \
        double pa_[4], pb_[4];                                                                                            \
        BC4_GATHER_LD(pa_, ro0);                                                                                          \
        BC4_GATHER_LD(pb_, ro1);                                                                                          \
        BC4_LOAD_W(wa);   /* behind the gathers in the shared-memory queue, needed a reduction later */                   \
        if (__builtin_expect(tpf != k + 1u, 0)) {                                                                         \
          BC4_SPIN((tpf = ld_acquire_cta_s(trdy_s + 4u * (k & (BC_TR - 1u)))) == k + 1u, 0x300u);                         \
          if (prof) pc[6] += 1;                                                                                           \
        }                                                                                                                 \
        /* t' through an address that depends on the flag; in flight while the gathers are summed */                      \
        tpv = lds_f64(tp_s + 8u * ((k & (BC_TR - 1u)) * 32u + rowg) + ((tpf ^ (k + 1u)) & 0x7u) * 8u);                    \
        double acc_;                                                                                                      \
        {                                                                                                                 \
          const double x0_ = (pa_[0] + pa_[1]) + (pa_[2] + pa_[3]);                                                       \
          const double x1_ = (pb_[0] + pb_[1]) + (pb_[2] + pb_[3]);                                                       \
          acc_ = -rv0 * x0_;                                                                                              \
          acc_ = fma(-rv1, x1_, acc_);                                                                                    \
        }                                                                                                                 \
        if (__builtin_expect(nbt > 1u, 0))                                                                                \
        _Pragma("unroll 1") for (uint32_t bt_ = 1; bt_ < nbt; bt_++) { /* more than 8 recent slots */                     \
          const uint32_t q_s_ = a_s + BC_AHDR + BC_WBYTES + BC_RBATCH * bt_;                                              \
          double v0_, v1_;                                                                                                \
          uint32_t o0_, o1_;                                                                                              \
          asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0_), "=d"(v1_) : "r"(q_s_ + rec_off) : "memory");       \
          asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(o0_), "=r"(o1_) : "r"(q_s_ + off_off) : "memory");       \
          BC4_GATHER_LD(pa_, o0_);                                                                                        \
          BC4_GATHER_LD(pb_, o1_);                                                                                        \
          acc_ = fma(-v0_, (pa_[0] + pa_[1]) + (pa_[2] + pa_[3]), acc_);                                                  \
          acc_ = fma(-v1_, (pb_[0] + pb_[1]) + (pb_[2] + pb_[3]), acc_);                                                  \
        }                                                                                                                 \
        acc_ += sub == 0u ? tpv : 0.0;                                                                                    \
        long long c2_ = 0;                                                                                                \
        if (prof) { c2_ = clock64() + (acc_ == 1.25e-300 ? 1 : 0); pc[1] += c2_ - c1_; }                                  \
        acc_ += __shfl_xor_sync(0xffffffffu, acc_, 1);                                                                    \
        acc_ += __shfl_xor_sync(0xffffffffu, acc_, 2);                                                                    \
        long long c3_ = 0;                                                                                                \
        if (prof) { c3_ = clock64() + (acc_ == 1.25e-300 ? 1 : 0); pc[2] += c3_ - c2_; }                                  \
        if (sub == 0u) sts_f64(tb_s + 2u * lane, acc_);                                                                   \
        __syncwarp();                                                                                                     \
        double t0_, t1_, t2_, t3_, t4_, t5_, t6_, t7_;                                                                    \
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(t0_), "=d"(t1_) : "r"(tb_s) : "memory");                   \
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(t2_), "=d"(t3_) : "r"(tb_s + 16u) : "memory");             \
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(t4_), "=d"(t5_) : "r"(tb_s + 32u) : "memory");             \
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(t6_), "=d"(t7_) : "r"(tb_s + 48u) : "memory");             \
        /* staging slot of the next chunk: ring bookkeeping and phase test while the t broadcast is in flight */          \
        const uint32_t oslot_ = slot;                                                                                     \
        if (more_) { if (++slot == P.SA) { slot = 0; par ^= 1u; } BC4_MBAR_TEST(a_ok, fullA_s + 8u * slot, par); }         \
        long long c4_ = 0;                                                                                                \
        if (prof) { c4_ = clock64(); pc[4] += c4_ - c3_; }                                                                \
        double a0_ = wa[0] * t0_, a1_ = wa[1] * t1_;                                                                      \
        a0_ = fma(wa[2], t2_, a0_);                                                                                       \
        a1_ = fma(wa[3], t3_, a1_);                                                                                       \
        a0_ = fma(wa[4], t4_, a0_);                                                                                       \
        a1_ = fma(wa[5], t5_, a1_);                                                                                       \
        a0_ = fma(wa[6], t6_, a0_);                                                                                       \
        a1_ = fma(wa[7], t7_, a1_);                                                                                       \
        const double a_ = a0_ + a1_;                                                                                      \
        sts_f64(pw_s + ((k & 3u) << 8), a_);                                                                              \
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(emptyA_s + 8u * oslot_) : "memory"); \
        /* chain-independent loads of the next chunk: in flight across the barrier */                                     \
        if (more_) {                                                                                                      \
          BC4_LOAD_TP(k + 1u);                                                                                            \
          if (__builtin_expect(a_ok == 0u, 0)) {                                                                          \
            long long c0_ = 0;                                                                                            \
            if (prof) c0_ = clock64();                                                                                    \
            BC4_SPIN(mbar_try(fullA + slot, par), 0x200u);                                                                \
            if (prof) pc[0] += clock64() - c0_;                                                                           \
          }                                                                                                               \
          BC4_LOAD_REC();                                                                                                 \
        }                                                                                                                 \
        long long c5_ = 0;                                                                                                \
        if (prof) { c5_ = clock64() + (a_ == 1.25e-300 ? 1 : 0); pc[5] += c5_ - c4_; }                                    \
        asm volatile("bar.sync 1, 160;" ::: "memory");                                                                    \
        if (prof) { pc[7] += clock64() - c5_; pc[3] += 1; }                                                               \
      } while (0)
      if (nch > 0) {
        BC4_SPIN(mbar_try(fullA + slot, par), 0x200u);
        BC4_LOAD_REC();
        BC4_LOAD_TP(0u);
      }
      uint32_t k = 0;
      for (; k + 1u < nch; k++) BC4_CHUNK(true);
      if (k < nch) BC4_CHUNK(false);
#undef BC4_CHUNK
#undef BC4_GATHER_LD
#undef BC4_LOAD_TP
#undef BC4_LOAD_W
#undef BC4_LOAD_REC
#undef BC4_MBAR_TEST
#undef BC4_SPIN
    } else if (CW == 4 && warp == 4u) {
      // ------------------------------ scribe: window <- x, progress counter -------------------------------
      const uint32_t pr_s = sc_s + 256u + 8u * lane;
      for (uint32_t k = 0; k < nch; k++) {
        asm volatile("bar.sync 1, 160;" ::: "memory");
        const uint32_t q_s = pr_s + ((k & 3u) << 8);
        const double p0 = lds_f64(q_s), p1 = lds_f64(q_s + 1024u), p2 = lds_f64(q_s + 2048u), p3 = lds_f64(q_s + 3072u);
        sts_f64(win_s + 8u * ((32u * k + lane) & wmask), (p0 + p1) + (p2 + p3));
        __syncwarp();
        if (lane == 0) st_release_cta_s(prog_s, k + 1u);
      }
    } else if (CW == 1 ? ((warp & 3u) != 0u && warp > 3u) : warp >= 8u) {
      // ------------------------------ near helpers -------------------------------------------------------
      // CW = 1: warps 5,6,7, 9,10,11, 13,14,15 ; CW = 4: warps 8..15
      const uint32_t hidx = CW == 1 ? ((warp >> 2) - 1u) * 3u + (warp & 3u) - 1u : warp - 8u;
      const bool hprof = PROF && (P.dbg & 1u) != 0u && blockIdx.x == 0 && hidx == 0u;
      // The start vector of a chunk (written by the far CTAs, tile flag + L2 round trips) is fetched one own chunk ahead:
      // the flag of the next own chunk's tile is read at the top of the iteration (the value comes back while the helper
      // works), tested before the helper starts to wait for the chain, and the start vector is loaded there.
      uint32_t tiles_known = 0;
      double t0n = 0.0;
      if (hidx < nch) {
        BC_WAIT(ld_acquire_gpu(P.tileflag + b.tile0 + hidx / P.tile) != 0u, 0x400u, 100);
        tiles_known = hidx / P.tile + 1u;
        const uint32_t j = b.lo + 32u * hidx + lane;
        t0n = j < b.hi ? __ldcg(P.w + j) : 0.0;
      }
      for (uint32_t k = hidx; k < nch; k += NH) {
        const uint32_t i = ia0 + k, slot = i % P.SB;
        long long h0 = 0;
        if (hprof) h0 = clock64();
        const double t0 = t0n;
        const uint32_t kn = k + NH, tilen = kn / P.tile;
        uint32_t fln = 1u;
        if (kn < nch && tilen >= tiles_known) fln = ld_acquire_gpu(P.tileflag + b.tile0 + tilen);
        long long h1 = 0;
        if (hprof) h1 = clock64();
        // The slot's sequence word is tested BEFORE the barrier's parity: helpers take chunks out of order, so with fewer
        // slots than helpers a helper can be two uses ahead of its slot, where the parity test alone is already true.
        BC_WAIT(lds_volatile_u32(smem_u32(seqB + slot)) == i && mbar_try(fullB + slot, (i / P.SB) & 1u), 0x500u, 200);
        const unsigned char *bp = reinterpret_cast<const unsigned char *>(ptrB[slot]);
        const uint32_t *hd = reinterpret_cast<const uint32_t *>(bp);
        const bool skip = (P.dbg & 6u) != 0u;   // timing experiments only (results are wrong)
        const uint32_t ne_max = skip ? 0u : hd[0], ne_tot = skip ? 0u : hd[1], nl = skip ? 0u : hd[2];
        const uint32_t perm = bp[16u + lane], rank = bp[48u + lane];
        const unsigned char *cnt = bp + BC_BHDR;
        const double *ev = reinterpret_cast<const double *>(bp + BC_BHDR + r16(ne_max));
        const uint16_t *ec = reinterpret_cast<const uint16_t *>(bp + BC_BHDR + r16(ne_max) + r16(8u * ne_tot));
        const double *lv = reinterpret_cast<const double *>(bp + BC_BHDR + r16(ne_max) + r16(8u * ne_tot) + r16(2u * ne_tot));
        const uint16_t *lc = reinterpret_cast<const uint16_t *>(reinterpret_cast<const unsigned char *>(lv) + 256u * nl);
        // early entries: columns in chunks <= k-E-1.  Four jagged diagonals per trip (their counts are one aligned
        // 32-bit load; the padding of the count array is zero), loads independent, two FMA chains.
        const uint32_t need1 = k > P.E ? k - P.E : 0u;
        long long h2 = 0;
        if (hprof) h2 = clock64();
        if (!(P.dbg & 32u)) BC_WAIT(ld_acquire_cta_s(prog_s) >= need1, 0x600u, 200);
        double ts = __shfl_sync(0xffffffffu, t0, (int)perm), ts1 = 0.0;
        uint32_t base = 0;
        for (uint32_t s = 0; s < ne_max; s += 4u) {
          const uint32_t c4 = *reinterpret_cast<const uint32_t *>(cnt + s);
          const uint32_t n0 = c4 & 255u, n1 = (c4 >> 8) & 255u, n2 = (c4 >> 16) & 255u, n3 = c4 >> 24;
          const uint32_t b1 = base + n0, b2 = b1 + n1, b3 = b2 + n2;
          double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0, x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
          if (lane < n0) { v0 = ev[base + lane]; x0 = win[ec[base + lane]]; }
          if (lane < n1) { v1 = ev[b1 + lane]; x1 = win[ec[b1 + lane]]; }
          if (lane < n2) { v2 = ev[b2 + lane]; x2 = win[ec[b2 + lane]]; }
          if (lane < n3) { v3 = ev[b3 + lane]; x3 = win[ec[b3 + lane]]; }
          ts = fma(-v0, x0, ts);
          ts1 = fma(-v1, x1, ts1);
          ts = fma(-v2, x2, ts);
          ts1 = fma(-v3, x3, ts1);
          base = b3 + n3;
        }
        ts += ts1;
        long long hj = 0;
        if (hprof) hj = clock64() + (ts == 1.25e-300 ? 1 : 0);   // (end of the jagged early entries)
        double t = __shfl_sync(0xffffffffu, ts, (int)rank);
        // late entries: columns in chunks k-E .. k-Kr-1.  Values and window slots of the first LB slots are loaded
        // before the wait; the t' slot and the window slot of chunk k must be free as well.
        constexpr uint32_t LB = 16;
        uint32_t lcr[LB];
        double lvr[LB];
#pragma unroll
        for (uint32_t u = 0; u < LB; u++) {
          lcr[u] = 0;
          lvr[u] = 0.0;
          if (u < nl) { lcr[u] = lc[u * 32u + lane]; lvr[u] = lv[u * 32u + lane]; }
        }
        // the staging slot is free once the late entries are in registers (the arrive depends on the last of them)
        const bool slot_done = nl <= LB;
        if (slot_done && lane == 0) mbar_arrive_after(emptyB + slot, lcr[LB - 1u], lvr[LB - 1u]);
        if (kn < nch) {
          if (fln == 0u) BC_WAIT(ld_acquire_gpu(P.tileflag + b.tile0 + tilen) != 0u, 0x400u, 100);
          tiles_known = tilen + 1u;
          const uint32_t jn = b.lo + 32u * kn + lane;
          t0n = jn < b.hi ? __ldcg(P.w + jn) : 0.0;
        }
        uint32_t need2 = k > P.Kr ? k - P.Kr : 0u;
        if (P.dbg & 32u) need2 = 0u;   // timing experiment only: free-running chain (results are wrong)
        if (k + 1u > BC_TR) need2 = max(need2, k + 1u - BC_TR);
        long long h3 = 0;
        if (hprof) h3 = clock64();
        // (polling costs shared-memory issue slots the critical warp needs: sleep while the chain is two or more
        //  chunks away, spin only for the last one)
        {
          uint32_t pnow = 0;
          G.n = 0;
          while ((pnow = ld_acquire_cta_s(prog_s)) < need2) {
            if (guard_poll(G, 0x700u)) break;
            if (need2 - pnow > 1u || (P.dbg & 16u)) __nanosleep(300);
          }
        }
        if (k >= P.Dfar && !(P.dbg & 32u)) BC_WAIT(ld_acquire_cta_s(pub_s) >= k - P.Dfar + 1u, 0x800u, 100);
        long long h4 = 0;
        if (hprof) h4 = clock64();
        {
          double xv[LB], q1 = 0.0, q2 = 0.0, q3 = 0.0;
#pragma unroll
          for (uint32_t u = 0; u < LB; u++) {
            xv[u] = 0.0;
            lds_f64_if(xv[u], win_s + 8u * lcr[u], u < nl);
          }
#pragma unroll
          for (uint32_t u = 0; u < LB; u += 4u) {
            t = fma(-lvr[u], xv[u], t);
            q1 = fma(-lvr[u + 1u], xv[u + 1u], q1);
            q2 = fma(-lvr[u + 2u], xv[u + 2u], q2);
            q3 = fma(-lvr[u + 3u], xv[u + 3u], q3);
          }
          for (uint32_t s0 = LB; s0 < nl; s0 += 8u) {   // long late rows (separator blocks): batches of 8, loads first
            uint32_t cc[8];
            double vv[8], xx[8];
#pragma unroll
            for (uint32_t u = 0; u < 8u; u++) {
              const bool have = s0 + u < nl;
              cc[u] = have ? lc[(s0 + u) * 32u + lane] : P.W;
              vv[u] = have ? lv[(s0 + u) * 32u + lane] : 0.0;
            }
#pragma unroll
            for (uint32_t u = 0; u < 8u; u++) xx[u] = win[cc[u]];
#pragma unroll
            for (uint32_t u = 0; u < 8u; u += 4u) {
              t = fma(-vv[u], xx[u], t);
              q1 = fma(-vv[u + 1u], xx[u + 1u], q1);
              q2 = fma(-vv[u + 2u], xx[u + 2u], q2);
              q3 = fma(-vv[u + 3u], xx[u + 3u], q3);
            }
          }
          t = (t + q1) + (q2 + q3);
        }
        const uint32_t tsl = k & (BC_TR - 1u);
        sts_f64(tp_s + 8u * (tsl * 32u + lane), t);
        __syncwarp();
        if (lane == 0) {
          st_release_cta_s(trdy_s + 4u * tsl, k + 1u);
          if (!slot_done) mbar_arrive(emptyB + slot);
        }
        // ph[0]: tile flag / start vector of this chunk + late-entry loads and the NEXT own chunk's tile flag and start vector;
        // ph[2]: wait for the chain to reach k-E + jagged early entries
        if (hprof) { ph[0] += (h1 - h0) + (h3 - hj); ph[1] += h2 - h1; ph[2] += hj - h2; ph[3] += h4 - h3; ph[4] += clock64() - h4; }
      }
    } else if (warp == W_PRODA) {
      // ------------------------------ TMA producer, ring A -----------------------------------------------
      // blob offsets of 32 chunks per load, the next batch in flight while this one is issued
      int64_t O0n = 0, O1n = 0;
      if (nch > 0) {
        const uint32_t gl = b.chunk0 + min(lane, nch - 1u);
        O0n = P.offA[gl]; O1n = P.offA[gl + 1];
      }
      for (uint32_t base = 0; base < nch; base += 32u) {
        const int64_t O0 = O0n, O1 = O1n;
        if (base + 32u < nch) {
          const uint32_t gl = b.chunk0 + min(base + 32u + lane, nch - 1u);
          O0n = P.offA[gl]; O1n = P.offA[gl + 1];
        }
        for (uint32_t l = 0; l < 32u && base + l < nch; l++) {
          const uint32_t i = ia0 + base + l, slot = i % P.SA, use = i / P.SA;
          const int64_t e0 = __shfl_sync(0xffffffffu, O0, (int)l), e1 = __shfl_sync(0xffffffffu, O1, (int)l);
          if (use > 0u) BC_WAIT(mbar_try(emptyA + slot, (use - 1u) & 1u), 0x900u, 20);
          if (lane == 0) {
            const uint32_t bytes = (uint32_t)(e1 - e0);
            mbar_expect_tx(fullA + slot, bytes);
            bulk_g2s(ringA + (size_t)slot * P.capA, P.blobA + e0, bytes, fullA + slot);
          }
          __syncwarp();
        }
      }
    } else if (warp == W_PRODB) {
      // ------------------------------ TMA producer, ring B -----------------------------------------------
      // blob offsets of 32 chunks per load, the next batch in flight while this one is issued
      int64_t O0n = 0, O1n = 0;
      if (nch > 0) {
        const uint32_t gl = b.chunk0 + min(lane, nch - 1u);
        O0n = P.offB[gl]; O1n = P.offB[gl + 1];
      }
      for (uint32_t base = 0; base < nch; base += 32u) {
        const int64_t O0 = O0n, O1 = O1n;
        if (base + 32u < nch) {
          const uint32_t gl = b.chunk0 + min(base + 32u + lane, nch - 1u);
          O0n = P.offB[gl]; O1n = P.offB[gl + 1];
        }
        for (uint32_t l = 0; l < 32u && base + l < nch; l++) {
          const uint32_t i = ia0 + base + l, slot = i % P.SB, use = i / P.SB;
          const int64_t e0 = __shfl_sync(0xffffffffu, O0, (int)l), e1 = __shfl_sync(0xffffffffu, O1, (int)l);
          if (use > 0u) BC_WAIT(mbar_try(emptyB + slot, (use - 1u) & 1u), 0xA00u, 20);
          if (lane == 0) {
            const uint32_t bytes = (uint32_t)(e1 - e0);
            if (e1 - e0 <= (int64_t)P.capB && !(P.dbg & 4u)) {
              unsigned char *dst = ringB + (size_t)slot * P.capB;
              ptrB[slot] = (unsigned long long)dst;
              sts_volatile_u32(smem_u32(seqB + slot), i);
              mbar_expect_tx(fullB + slot, bytes);
              bulk_g2s(dst, P.blobB + e0, bytes, fullB + slot);
            } else {   // does not fit a staging slot: the helper reads it from HBM
              ptrB[slot] = (unsigned long long)(P.blobB + e0);
              sts_volatile_u32(smem_u32(seqB + slot), i);
              mbar_arrive(fullB + slot);
            }
          }
          __syncwarp();
        }
      }
    } else if (warp == W_PUB) {
      // ------------------------------ publisher ----------------------------------------------------------
      uint32_t done = 0;
      double dot = 0.0;
      while (done < nch) {
        uint32_t p = done;
        BC_WAIT((p = ld_acquire_cta_s(prog_s)) > done, 0xB00u, 400);
        if (p <= done) break;   // aborted
        for (uint32_t k = done; k < p; k += 4u) {
          double x[4], dv[4];
          uint32_t idx[4];
          bool ok[4];
#pragma unroll
          for (uint32_t u = 0; u < 4u; u++) {
            const uint32_t j = b.lo + 32u * (k + u) + lane;
            ok[u] = (k + u < p) && j < b.hi;
            idx[u] = P.reversed ? P.N - 1u - j : j;
            x[u] = lds_f64(win_s + 8u * ((32u * (k + u) + lane) & wmask));
            dv[u] = (ok[u] && P.dotvec && idx[u] < P.dot_limit) ? P.dotvec[idx[u]] : 0.0;
          }
#pragma unroll
          for (uint32_t u = 0; u < 4u; u++) {
            if (ok[u]) {
              P.out[idx[u]] = x[u];
              dot = fma(x[u], dv[u], dot);
            }
          }
        }
        __syncwarp();
        if (lane == 0) {
          __threadfence();
          st_release_gpu(P.gprog + b.gidx, p);
          st_release_cta_s(pub_s, p);
        }
        done = p;
      }
      dot = warp_sum(dot);
      if (lane == 0 && P.dot_partials) P.dot_partials[b.gidx] = dot;
    }
    __syncthreads();
    ia0 += nch;
  }
  if (P.clk && blockIdx.x == 0 && threadIdx.x == 0) {
    P.clk[0] = (unsigned long long)(clock64() - clk0);
    if (P.dbg & 1u) {
      for (int q = 0; q < 4; q++) P.clk[3 + q] = (unsigned long long)pc[q];
      P.clk[13] = (unsigned long long)pc[4];
      P.clk[14] = (unsigned long long)pc[5];
      P.clk[7] = (unsigned long long)pc[7];    // split chain: cycles in the per-chunk barrier
      P.clk[15] = (unsigned long long)pc[6];   // failed polls of t'
    }
  }
  if (P.clk && (P.dbg & 1u) && blockIdx.x == 0 && threadIdx.x == (CW == 1 ? 160u : 256u))
    for (int q = 0; q < 5; q++) P.clk[8 + q] = (unsigned long long)ph[q];
}

// multi-GPU forward solve: contribution of the rank's own subtree to the right-hand sides of the top separators,
// sbuf[i - n_sub] = sum over far entries with column < n_sub of L[j,c] out[c]   (then summed over the ranks by NCCL)
__global__ void __launch_bounds__(256) k_bc_couple(const int64_t *__restrict__ rp, const uint32_t *__restrict__ col,
                                                   const double *__restrict__ val, const BcBlock *__restrict__ blocks,
                                                   const double *__restrict__ out, double *__restrict__ sbuf, uint32_t n_sub) {
  const BcBlock b = blocks[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const uint32_t wpc = blockDim.x >> 5;
  for (uint32_t v = b.lo + blockIdx.x * wpc + (threadIdx.x >> 5); v < b.hi; v += gridDim.x * wpc) {
    double acc = 0.0;
    const int64_t e = rp[v + 1];
    for (int64_t k = rp[v] + lane; k < e; k += 32) {
      const uint32_t c = col[k];
      if (c < n_sub) acc = fma(val[k], out[c], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) sbuf[v - n_sub] = acc;   // forward direction: vector index == solve index
  }
}

uint32_t floor_pow2_u32(uint32_t v) {
  uint32_t p = 1;
  while ((p << 1) <= v && (p << 1) != 0) p <<= 1;
  return p;
}

// dense-panel levels: tree levels with at most this many blocks (measured at 256^3 / T = 4096, profiles/r02_levels_256.md:
// faster than the chains up to 64 blocks per level, slower than the warp-per-block kernels from 128 on)
constexpr uint32_t DP_MAX_BLOCKS_DEFAULT = 64u;
#include "rcg_cluster.cuh"
#include "rcg_fold.cuh"
#include "rcg_dense.cuh"

}  // namespace

bool rcg_use_blocked(const rcg_handle *h) {
  const int m = h->opt.chain_mode;
  return !h->opt.chain_generic && (m == 0 || m == 3 || m == 4 || m == 5 || m == 6);
}

void rcg_free_blocked(BlockedDev &b) {
  cudaFree(b.offA); cudaFree(b.offB); cudaFree(b.blobA); cudaFree(b.blobB);
  rcg_free_csr(b.far);
  cudaFree(b.tile_need); cudaFree(b.flags); cudaFree(b.w); cudaFree(b.blocks); cudaFree(b.far_split);
  cudaFree(b.dp.panels); cudaFree(b.dp.hop_ptr); cudaFree(b.dp.inv); cudaFree(b.dp.t0); cudaFree(b.dp.t1);
  rcg_free_csr(b.dp.near);
  cudaFree(b.cl.wslab); cudaFree(b.cl.blobN); cudaFree(b.cl.offN); cudaFree(b.cl.c0); cudaFree(b.cl.prog4);
  b = BlockedDev();
}

// `comb`: the direction's lower-triangular matrix in its solve index space, rows sorted by column, diagonal last, raw
// values.  Builds the blocked layout and frees `comb`.  bounds/depth describe the blocks in ascending solve order.
int rcg_build_blocked(rcg_handle *h, DirectionDev &d, CsrDev &comb, const std::vector<uint32_t> &bounds,
                      const std::vector<int> &depth, int max_depth, bool root_first) {
  BlockedDev &B = d.bc;
  const uint32_t N = (uint32_t)h->N;
  const int nb = (int)bounds.size() - 1;
  RcgPhases ph;
  // ---- thresholds ------------------------------------------------------------------------------------
  // (the split chain keeps the solution of two chunks in its partial-sum buffers: recent distance at most 2)
  // chain_mode 0 (default): round-1 blocked chain (four critical warps) for the tree levels with few blocks -- measured 15 %
  // faster per chunk than the folded chain -- and warp-per-block launches for the levels with many blocks; 5: folded chain
  // + warp-per-block levels; 6: round-1 chain for every level
  B.fold = h->opt.chain_mode == 5;
  const bool wb_allowed = h->opt.chain_mode == 0 || h->opt.chain_mode == 5;
  B.Kr = h->opt.reserved[3] > 0 ? (uint32_t)std::min(h->opt.reserved[3], h->opt.chain_mode == 3 ? 4 : 2) : 2u;
  // folded chain: Kr is the fold depth (chunks whose entries become dense panel columns); it is also the slack, in hops,
  // that the near helper has to deliver u_k after the chain solved chunk k-Kr-1
  if (B.fold) B.Kr = h->opt.reserved[3] > 0 ? (uint32_t)std::min<int>(h->opt.reserved[3], (int)FC_KRMAX) : 3u;
  // chunks k-E .. k-Kr-1 are the "late" class (gathered by the near helper AFTER the chain reached chunk k-Kr);
  // older chunks inside the window are "early" (gathered before).  reserved[5] overrides (tuning experiments).
  {
    const int e_leaf = h->opt.reserved[5] & 0xFF, e_sep = (h->opt.reserved[5] >> 8) & 0xFF;   // reserved[5]: leaves | separators << 8
    B.E = e_leaf > 0 ? (uint32_t)std::min(255, std::max<int>(e_leaf, (int)B.Kr + 1)) : 16u;   // (E >= window: no early class)
    // separator blocks are nearly dense next to the diagonal: with E = 16 their late ELL part has 35-50 slots, more than
    // a helper keeps in registers, so the helper holds its staging slot through the late phase; E = 6 keeps it short
    B.E_sep = e_sep > 0 ? (uint32_t)std::min(255, std::max<int>(e_sep, (int)B.Kr + 1)) : std::min(B.E, 6u);
  }
  uint32_t win_rows = h->opt.chain_window > 0 ? (uint32_t)h->opt.chain_window : 4096u;
  B.Dfar = std::min(512u, std::max(32u, floor_pow2_u32(std::max(32u, win_rows) / 32u)));
  // Separator blocks are small, dense and have every far CTA of the launch to themselves: a short window sends most of
  // their entries to the far CTAs and keeps their chain blobs small enough for many staging slots.
  const uint32_t sep_rows = h->opt.reserved[4] > 0 ? (uint32_t)h->opt.reserved[4] : 1024u;
  // (at least 32 chunks: the far CTAs' hand-off takes 10-20 us -- published progress, two dependent L2 gathers, tile flag --
  //  and can only start when the chain is a window away; measured at 256^3 / T=512: a 512-row window doubles the root
  //  separator's time, a 256-row window makes it 9x slower)
  B.Dfar_sep = std::min(B.Dfar, std::max(32u, floor_pow2_u32(std::max(32u, sep_rows) / 32u)));
  if (B.fold) B.Kr = std::min(B.Kr, B.Dfar_sep - 1u);
  B.E_sep = std::max(B.Kr, std::min(B.E_sep, B.Dfar_sep - 1u));   // Kr <= E <= window - 1
  // ---- chunk / tile numbering ------------------------------------------------------------------------
  // Far tiles: the unit in which the far CTAs hand the start vector to the chain.  A tile's in-block pass can start when
  // the chain is a window away from it and is a chain of dependent HBM round trips (measured 26-50 us for 8 chunks of a
  // separator), so the short-window separator blocks use small tiles: more tiles in flight, each one sweep of one CTA.
  const int tile_opt = (h->opt.reserved[9] >> 8) & 0xFF;   // reserved[9] bits 8-15: chunks per far tile of the separator blocks
  B.tile_sep = tile_opt > 0 ? (uint32_t)std::min(8, tile_opt) : 1u;
  std::vector<uint32_t> chunk0(nb + 1, 0), tile0(nb + 1, 0), dfar(nb + 1, B.Dfar), tilesz(nb + 1, BC_TILE), eblk(nb + 1, B.E);
  for (int b = 0; b < nb; b++) dfar[b] = (depth[b] == max_depth) ? B.Dfar : B.Dfar_sep;
  for (int b = 0; b < nb; b++) tilesz[b] = (depth[b] == max_depth) ? BC_TILE : B.tile_sep;
  for (int b = 0; b < nb; b++) eblk[b] = (depth[b] == max_depth) ? B.E : B.E_sep;
  // Warp-per-block levels (k_wb_solve, rcg_fold.cuh): a tree level with many blocks has enough independent chains to fill
  // the GPU with one WARP per block -- no chain CTA, no handshakes; every in-window entry is applied by the block's warp
  // (Kr = 0: nothing folded, E = 0: one jagged class), older entries of the own block and the other blocks stay far.
  std::vector<uint32_t> krblk(nb + 1, B.Kr), wbblk(nb + 1, 0);
  {
    const int wb_opt = (h->opt.reserved[9] >> 16) & 0xFFFF;   // reserved[9] bits 16-31: blocks per level from which the
    B.wb_min = !wb_allowed || wb_opt == 0xFFFF ? 0u : wb_opt > 0 ? (uint32_t)wb_opt : 64u;   // level is warp-per-block (0xFFFF: never)
    // window of a warp-per-block level: 2048 rows.  Older entries of the own block are gathered from HBM/L2 by the block's
    // warp, one dependent round trip after the other (measured at 256^3 / T=512 with a 1024-row window: the plane
    // neighbours of a 32^3 leaf sit ~1024 rows back, 40 % of the factor was "far" and a chunk took 9600 cycles)
    B.Dfar_wb = 64u;
    std::vector<int> per_depth(max_depth + 2, 0);
    std::vector<uint32_t> nch_depth(max_depth + 2, 0);   // longest block of a level, in chunks
    for (int b = 0; b < nb; b++)
      if (bounds[b + 1] > bounds[b]) {
        per_depth[depth[b]]++;
        nch_depth[depth[b]] = std::max(nch_depth[depth[b]], (bounds[b + 1] - bounds[b] + 31u) / 32u);
      }
    for (int b = 0; b < nb; b++)
      if (B.wb_min > 0 && (uint32_t)per_depth[depth[b]] >= B.wb_min) {
        // in-window entries as ONE class: jagged diagonals (E = 0; rows sorted by length, no padding) by default -- measured
        // at 256^3 / T=4096: 2.16 ms per leaf level against 2.91 ms for ELL (E = window, reserved[2] = 1: lane = row, padded)
        // window of the level: the whole block when it is short (no far entries inside the own block at all), at most
        // Dfar_wb chunks
        // (leaves keep the natural order of a 3-D box: the plane neighbours sit rows^(2/3) back, two planes are kept;
        //  separator blocks are dense and short: the whole block, at most 128 chunks)
        uint32_t want = nch_depth[depth[b]], cap = 128u;
        if (depth[b] == max_depth) {
          want = (uint32_t)(2.0 * std::pow(32.0 * nch_depth[depth[b]], 2.0 / 3.0) / 32.0) + 1u;
          cap = B.Dfar_wb;
        }
        uint32_t dw = 4u;
        while (dw < want && dw < cap) dw <<= 1;
        wbblk[b] = 1; krblk[b] = 0; eblk[b] = h->opt.reserved[2] == 1 ? dw : 0u; dfar[b] = dw; tilesz[b] = B.tile_sep;
      }
  }
  // Dense-panel levels (k_dp_solve, rcg_dense.cuh): every separator level whose longest block has at least dp_min rows is
  // solved in lock step on the whole GPU through explicitly inverted C x C diagonal panels (C per level).
  // reserved[6] bits 8-15: 0 = default (128 rows), 255 = never, else rows / 32; bits 16-31: panel rows (0 = from the level's
  // shape, dp_choose_panel); bit 1: the leaf level too.
  std::vector<uint32_t> dpcblk(nb + 1, 0), dpq0(nb + 1, 0);
  int64_t dp_rows_total = 0;
  {
    const int dp_opt = (h->opt.reserved[6] >> 8) & 0xFF, dp_c = (h->opt.reserved[6] >> 16) & 0xFFFF;
    const bool dp_leaf = (h->opt.reserved[6] & 2) != 0;
    // bits 2-7: levels with more than 2^(v-1) blocks stay on the chain / warp-per-block kernels (0 = default, 63 = no limit):
    // the lock-step hops pay off where a level has FEW long blocks; many short blocks are throughput work (k_wb_solve)
    const int dp_mb = (h->opt.reserved[6] >> 2) & 0x3F;
    uint32_t dp_max_blocks = dp_mb == 0 ? DP_MAX_BLOCKS_DEFAULT : dp_mb >= 32 ? 0xFFFFFFFFu : (1u << (dp_mb - 1));
    if (const char *e = getenv("RCG_DP_MAX_BLOCKS")) if (atoi(e) > 0) dp_max_blocks = (uint32_t)atoi(e);   // tuning experiments
    std::vector<uint32_t> cnt(max_depth + 2, 0);
    for (int b = 0; b < nb; b++) if (bounds[b + 1] > bounds[b]) cnt[depth[b]]++;
    const uint32_t dp_min = !wb_allowed || dp_opt == 255 ? 0xFFFFFFFFu : dp_opt > 0 ? 32u * (uint32_t)dp_opt : 128u;
    std::vector<uint32_t> mx(max_depth + 2, 0);
    std::vector<int64_t> sum(max_depth + 2, 0);
    for (int b = 0; b < nb; b++) {
      mx[depth[b]] = std::max(mx[depth[b]], bounds[b + 1] - bounds[b]);
      sum[depth[b]] += bounds[b + 1] - bounds[b];
    }
    for (int b = 0; b < nb; b++) {
      const int dd = depth[b];
      if ((dd == max_depth && !dp_leaf) || mx[dd] < dp_min || bounds[b + 1] == bounds[b] || cnt[dd] > dp_max_blocks) continue;
      dpcblk[b] = dp_choose_panel(mx[dd], sum[dd], dp_c, cnt[dd]);
      dpq0[b] = (uint32_t)dp_rows_total;
      dp_rows_total += bounds[b + 1] - bounds[b];
      wbblk[b] = 0; krblk[b] = 0; dfar[b] = dpcblk[b]; tilesz[b] = B.tile_sep;
    }
  }
  for (int b = 0; b < nb; b++) {
    const uint32_t nch = (bounds[b + 1] - bounds[b] + 31u) / 32u;
    chunk0[b + 1] = chunk0[b] + nch;
    tile0[b + 1] = tile0[b] + (nch + tilesz[b] - 1u) / tilesz[b];
  }
  B.nchunks = chunk0[nb];
  B.ntiles = tile0[nb];
  uint32_t *dgeom = nullptr;   // bounds | chunk0 | tile0 | dfar | tile size | E | Kr | warp-per-block flag | panel rows | first compact row
  RCG_CUDA(h, cudaMalloc(&dgeom, sizeof(uint32_t) * 10 * (nb + 1)));
  RCG_CUDA(h, cudaMemcpyAsync(dgeom + 9 * (nb + 1), dpq0.data(), sizeof(uint32_t) * (nb + 1), cudaMemcpyHostToDevice, h->stream));
  RCG_CUDA(h, cudaMemcpyAsync(dgeom + 8 * (nb + 1), dpcblk.data(), sizeof(uint32_t) * (nb + 1), cudaMemcpyHostToDevice, h->stream));
  RCG_CUDA(h, cudaMemcpyAsync(dgeom + 7 * (nb + 1), wbblk.data(), sizeof(uint32_t) * (nb + 1), cudaMemcpyHostToDevice, h->stream));
  RCG_CUDA(h, cudaMemcpyAsync(dgeom + 6 * (nb + 1), krblk.data(), sizeof(uint32_t) * (nb + 1), cudaMemcpyHostToDevice, h->stream));
  RCG_CUDA(h, cudaMemcpyAsync(dgeom + 5 * (nb + 1), eblk.data(), sizeof(uint32_t) * (nb + 1), cudaMemcpyHostToDevice, h->stream));
  RCG_CUDA(h, cudaMemcpyAsync(dgeom + 4 * (nb + 1), tilesz.data(), sizeof(uint32_t) * (nb + 1), cudaMemcpyHostToDevice, h->stream));
  RCG_CUDA(h, cudaMemcpyAsync(dgeom + 3 * (nb + 1), dfar.data(), sizeof(uint32_t) * (nb + 1), cudaMemcpyHostToDevice, h->stream));
  RCG_CUDA(h, cudaMemcpyAsync(dgeom, bounds.data(), sizeof(uint32_t) * (nb + 1), cudaMemcpyHostToDevice, h->stream));
  RCG_CUDA(h, cudaMemcpyAsync(dgeom + (nb + 1), chunk0.data(), sizeof(uint32_t) * (nb + 1), cudaMemcpyHostToDevice, h->stream));
  RCG_CUDA(h, cudaMemcpyAsync(dgeom + 2 * (nb + 1), tile0.data(), sizeof(uint32_t) * (nb + 1), cudaMemcpyHostToDevice, h->stream));
  BcGeom g;
  g.bounds = dgeom; g.chunk0 = dgeom + (nb + 1); g.tile0 = dgeom + 2 * (nb + 1);
  g.dfar = dgeom + 3 * (nb + 1);
  g.tile = dgeom + 4 * (nb + 1);
  g.eblk = dgeom + 5 * (nb + 1);
  g.krblk = dgeom + 6 * (nb + 1);
  g.wb = dgeom + 7 * (nb + 1);
  g.dpc = dgeom + 8 * (nb + 1);
  g.dpq0 = dgeom + 9 * (nb + 1);
  g.nb = nb; g.Kr = B.Kr; g.E = B.E;
  g.fold = B.fold ? 1u : 0u;

  // ---- sizes ---------------------------------------------------------------------------------------------
  int *derr = nullptr;
  RCG_CUDA(h, cudaMalloc(&derr, sizeof(int)));
  RCG_CUDA(h, cudaMemsetAsync(derr, 0, sizeof(int), h->stream));
  RCG_CUDA(h, cudaMalloc(&B.offA, sizeof(int64_t) * ((size_t)B.nchunks + 1)));
  RCG_CUDA(h, cudaMalloc(&B.offB, sizeof(int64_t) * ((size_t)B.nchunks + 1)));
  RCG_CUDA(h, cudaMalloc(&B.far.rowptr, sizeof(int64_t) * ((size_t)N + 1)));
  RCG_CUDA(h, cudaMalloc(&B.tile_need, sizeof(uint32_t) * std::max(1u, B.ntiles)));
  RCG_CUDA(h, cudaMemsetAsync(B.offA, 0, sizeof(int64_t) * ((size_t)B.nchunks + 1), h->stream));
  RCG_CUDA(h, cudaMemsetAsync(B.offB, 0, sizeof(int64_t) * ((size_t)B.nchunks + 1), h->stream));
  RCG_CUDA(h, cudaMemsetAsync(B.far.rowptr, 0, sizeof(int64_t) * ((size_t)N + 1), h->stream));
  RCG_CUDA(h, cudaMemsetAsync(B.tile_need, 0, sizeof(uint32_t) * std::max(1u, B.ntiles), h->stream));
  B.dp.nrows = dp_rows_total;
  RCG_CUDA(h, cudaMalloc(&B.dp.near.rowptr, sizeof(int64_t) * ((size_t)dp_rows_total + 4)));
  RCG_CUDA(h, cudaMemsetAsync(B.dp.near.rowptr, 0, sizeof(int64_t) * ((size_t)dp_rows_total + 4), h->stream));
  const auto t_count = std::chrono::steady_clock::now();
  const int cgrid = (int)std::min<int64_t>(((int64_t)B.nchunks + 7) / 8, (int64_t)h->sm_count * 16);
  k_bc_count<<<std::max(1, cgrid), 256, 0, h->stream>>>(comb.rowptr, comb.col, g, B.nchunks, B.offA, B.offB, B.far.rowptr,
                                                       B.tile_need, derr, B.dp.near.rowptr);
  h->stats.kernel_launches += 1;
  int herr = 0;
  RCG_CUDA(h, cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  RCG_CUDA(h, cudaFree(derr));
  if (getenv("RCG_TIMING")) fprintf(stderr, "[rcg] k_bc_count: %.1f ms\n", 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t_count).count());
  if (herr) {
    cudaFree(dgeom);
    h->err = "factor row without a trailing diagonal after transposition (G is not triangular)";
    return RCG_ERR_STRUCTURE;
  }
  ph.mark(h->stream, "  layout: geometry, k_bc_count");
  RCG_TRY(rcg_exclusive_scan(h, B.offA, (int64_t)B.nchunks + 1));
  RCG_TRY(rcg_exclusive_scan(h, B.offB, (int64_t)B.nchunks + 1));
  RCG_TRY(rcg_exclusive_scan(h, B.far.rowptr, (int64_t)N + 1));
  RCG_TRY(rcg_exclusive_scan(h, B.dp.near.rowptr, dp_rows_total + 1));
  RCG_CUDA(h, cudaMemcpy(&B.dp.near.nnz, B.dp.near.rowptr + dp_rows_total, sizeof(int64_t), cudaMemcpyDeviceToHost));
  RCG_CUDA(h, cudaMalloc(&B.dp.near.col, sizeof(uint32_t) * (size_t)(B.dp.near.nnz + 32)));
  RCG_CUDA(h, cudaMalloc(&B.dp.near.val, sizeof(double) * (size_t)(B.dp.near.nnz + 16)));
  int64_t far_total = 0;
  RCG_CUDA(h, cudaMemcpy(&B.bytesA, B.offA + B.nchunks, sizeof(int64_t), cudaMemcpyDeviceToHost));
  RCG_CUDA(h, cudaMemcpy(&B.bytesB, B.offB + B.nchunks, sizeof(int64_t), cudaMemcpyDeviceToHost));
  RCG_CUDA(h, cudaMemcpy(&far_total, B.far.rowptr + N, sizeof(int64_t), cudaMemcpyDeviceToHost));
  B.far.nnz = far_total;
  RCG_CUDA(h, cudaMalloc(&B.blobA, (size_t)B.bytesA + 256));
  RCG_CUDA(h, cudaMalloc(&B.blobB, (size_t)B.bytesB + 256));
  RCG_CUDA(h, cudaMemsetAsync(B.blobA, 0, (size_t)B.bytesA + 256, h->stream));
  RCG_CUDA(h, cudaMemsetAsync(B.blobB, 0, (size_t)B.bytesB + 256, h->stream));
  RCG_CUDA(h, cudaMalloc(&B.far_split, sizeof(uint32_t) * ((size_t)N + 1)));
  RCG_CUDA(h, cudaMemsetAsync(B.far_split, 0, sizeof(uint32_t) * ((size_t)N + 1), h->stream));
  RCG_CUDA(h, cudaMalloc(&B.far.col, sizeof(uint32_t) * (size_t)(far_total + 8)));
  RCG_CUDA(h, cudaMalloc(&B.far.val, sizeof(double) * (size_t)(far_total + 8)));
  ph.mark(h->stream, "  layout: scans, allocations, memsets");
  const int fgrid = (int)std::min<int64_t>(((int64_t)B.nchunks + 3) / 4, (int64_t)h->sm_count * 16);
  k_bc_fill<<<std::max(1, fgrid), 128, 0, h->stream>>>(comb.rowptr, comb.col, comb.val, g, B.nchunks, N, d.reversed ? 1 : 0,
                                                      B.offA, B.offB, B.blobA, B.blobB, B.far.rowptr, B.far.col, B.far.val, B.far_split,
                                                      B.dp.near.rowptr, B.dp.near.col, B.dp.near.val);
  h->stats.kernel_launches += 1;
  RCG_CUDA(h, cudaGetLastError());
  if (getenv("RCG_TIMING")) {
    const auto t_fill = std::chrono::steady_clock::now();
    RCG_CUDA(h, cudaStreamSynchronize(h->stream));
    fprintf(stderr, "[rcg] k_bc_fill (+ blob memsets): %.1f ms after the launch\n", 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t_fill).count());
  }

  ph.mark(h->stream, "  layout: k_bc_fill");
  // ---- levels (dependency groups) and their blocks -----------------------------------------------------------
  d.groups.clear();
  B.blocks_host.clear();
  std::vector<int> src_block;
  for (int gl = 0; gl <= max_depth; gl++) {
    const int want = root_first ? gl : max_depth - gl;
    GroupHost G;
    G.depth = want;
    G.first = (int)B.blocks_host.size();
    for (int b = 0; b < nb; b++) {
      if (depth[b] != want || bounds[b + 1] == bounds[b]) continue;
      BcBlock bd;
      memset(&bd, 0, sizeof(bd));
      bd.lo = bounds[b]; bd.hi = bounds[b + 1]; bd.chunk0 = chunk0[b]; bd.tile0 = tile0[b];
      bd.gidx = (uint32_t)B.blocks_host.size();
      bd.pad[0] = dfar[b];
      bd.pad[1] = tilesz[b];
      bd.pad[2] = eblk[b] | (krblk[b] << 8) | (wbblk[b] << 16) | ((dpcblk[b] ? 1u : 0u) << 17);   // E | Kr << 8 | warp-per-block << 16 | dense-panel << 17 (pad[0] = panel rows)
      B.blocks_host.push_back(bd);
      src_block.push_back(b);
      G.count++;
      G.max_rows = std::max(G.max_rows, bd.hi - bd.lo);
      G.rows += bd.hi - bd.lo;
    }
    if (G.count > 0) {
      if ((B.blocks_host[G.first].pad[2] >> 16) & 1u) {   // warp-per-block level: longest block first (dynamic hand-out)
        std::vector<size_t> ord(G.count);
        for (int q = 0; q < G.count; q++) ord[q] = q;
        std::stable_sort(ord.begin(), ord.end(), [&](size_t x, size_t y) {
          const BcBlock &bx = B.blocks_host[G.first + x], &by = B.blocks_host[G.first + y];
          return bx.hi - bx.lo > by.hi - by.lo;
        });
        std::vector<BcBlock> tmpb(G.count);
        std::vector<int> tmps(G.count);
        for (int q = 0; q < G.count; q++) { tmpb[q] = B.blocks_host[G.first + ord[q]]; tmps[q] = src_block[G.first + ord[q]]; }
        for (int q = 0; q < G.count; q++) {
          tmpb[q].gidx = (uint32_t)(G.first + q);
          B.blocks_host[G.first + q] = tmpb[q];
          src_block[G.first + q] = tmps[q];
        }
      }
      d.groups.push_back(G);
    }
  }
  B.nblocks = (uint32_t)B.blocks_host.size();
  B.dp.q0_of_block.assign(B.blocks_host.size(), 0);
  for (size_t q = 0; q < B.blocks_host.size(); q++) B.dp.q0_of_block[q] = dpq0[src_block[q]];
  RCG_CUDA(h, cudaMalloc(&B.blocks, sizeof(BcBlock) * std::max<size_t>(1, B.blocks_host.size())));
  RCG_CUDA(h, cudaMemcpyAsync(B.blocks, B.blocks_host.data(), sizeof(BcBlock) * B.blocks_host.size(), cudaMemcpyHostToDevice,
                              h->stream));
  RCG_CUDA(h, cudaMalloc(&B.flags, sizeof(uint32_t) * ((size_t)B.ntiles + B.nblocks + 4)));
  RCG_CUDA(h, cudaMemsetAsync(B.flags, 0, sizeof(uint32_t) * ((size_t)B.ntiles + B.nblocks + 4), h->stream));
  RCG_CUDA(h, cudaMalloc(&B.w, sizeof(double) * ((size_t)N + 4)));
  RCG_CUDA(h, cudaMemsetAsync(B.w, 0, sizeof(double) * ((size_t)N + 4), h->stream));

  // ---- per-level staging plan from the blob sizes ------------------------------------------------------------
  std::vector<int64_t> hA((size_t)B.nchunks + 1), hB((size_t)B.nchunks + 1);
  RCG_CUDA(h, cudaMemcpyAsync(hA.data(), B.offA, sizeof(int64_t) * hA.size(), cudaMemcpyDeviceToHost, h->stream));
  RCG_CUDA(h, cudaMemcpyAsync(hB.data(), B.offB, sizeof(int64_t) * hB.size(), cudaMemcpyDeviceToHost, h->stream));
  // far entries per block (bench statistics)
  std::vector<int64_t> far_ends(2 * (size_t)B.nblocks), row_ends(2 * (size_t)B.nblocks);
  {
    std::vector<uint32_t> idx;
    for (const BcBlock &bd : B.blocks_host) { idx.push_back(bd.lo); idx.push_back(bd.hi); }
    uint32_t *didx = nullptr;
    int64_t *dout = nullptr;
    RCG_CUDA(h, cudaMalloc(&didx, sizeof(uint32_t) * std::max<size_t>(1, idx.size())));
    RCG_CUDA(h, cudaMalloc(&dout, sizeof(int64_t) * std::max<size_t>(1, idx.size())));
    RCG_CUDA(h, cudaMemcpyAsync(didx, idx.data(), sizeof(uint32_t) * idx.size(), cudaMemcpyHostToDevice, h->stream));
    k_gather_i64<<<((int)idx.size() + 255) / 256, 256, 0, h->stream>>>(B.far.rowptr, didx, (int)idx.size(), dout);
    RCG_CUDA(h, cudaMemcpyAsync(far_ends.data(), dout, sizeof(int64_t) * idx.size(), cudaMemcpyDeviceToHost, h->stream));
    RCG_CUDA(h, cudaStreamSynchronize(h->stream));
    k_gather_i64<<<((int)idx.size() + 255) / 256, 256, 0, h->stream>>>(comb.rowptr, didx, (int)idx.size(), dout);
    RCG_CUDA(h, cudaMemcpyAsync(row_ends.data(), dout, sizeof(int64_t) * idx.size(), cudaMemcpyDeviceToHost, h->stream));
    RCG_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(didx); cudaFree(dout);
  }
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  B.levels.clear();
  for (GroupHost &G : d.groups) {
    const uint32_t W = 32u * B.blocks_host[G.first].pad[0];
    int64_t maxA = 0, maxB = 0, sumB = 0, nchl = 0;
    for (int bi = G.first; bi < G.first + G.count; bi++) {
      const BcBlock &bd = B.blocks_host[bi];
      const uint32_t nch = (bd.hi - bd.lo + 31u) / 32u;
      for (uint32_t c = bd.chunk0; c < bd.chunk0 + nch; c++) {
        maxA = std::max(maxA, hA[c + 1] - hA[c]);
        maxB = std::max(maxB, hB[c + 1] - hB[c]);
        sumB += hB[c + 1] - hB[c];
      }
      nchl += nch;
      const int64_t far_n = far_ends[2 * (size_t)bi + 1] - far_ends[2 * (size_t)bi];
      G.ext_nnz += far_n;                                                                   // entries done by the far CTAs
      G.loc_nnz += (row_ends[2 * (size_t)bi + 1] - row_ends[2 * (size_t)bi]) - far_n;       // entries done by the chain CTA
      G.blob_bytes += (hA[bd.chunk0 + nch] - hA[bd.chunk0]) + (hB[bd.chunk0 + nch] - hB[bd.chunk0]);
    }
    G.max_stage = (uint32_t)maxA;
    BcLevel L;
    if ((B.blocks_host[G.first].pad[2] >> 17) & 1u) {   // dense-panel level: no blobs, no rings (dp_build below)
      L.dp = true;
      L.Dfar = B.blocks_host[G.first].pad[0];
      L.groups = (uint32_t)G.count;
      B.levels.push_back(L);
      continue;
    }
    if ((B.blocks_host[G.first].pad[2] >> 16) & 1u) {   // warp-per-block level: per-warp window + scratch, no rings
      L.wb = true;
      L.Dfar = B.blocks_host[G.first].pad[0];
      const uint32_t Wwb = 32u * L.Dfar;
      // warps per CTA and staging buffers per warp: as many warps as possible (their turns hide each other's latencies;
      // measured at 256^3 / T=4096: one warp per scheduler needs ~5400 cycles per chunk), then two buffers if they fit
      L.capB = (uint32_t)((maxB + 127) & ~127ll);   // every blob of the level is staged
      uint32_t wbw = 1u, nbuf = 1u;
      if ((int64_t)wb_warp_bytes(Wwb, L.capB, 1u) > (int64_t)BC_SMEM_MAX) {
        cudaFree(dgeom);
        h->err = "blocked solve: a chunk of a warp-per-block level does not fit shared memory (raise wb_min: rcg_options.reserved[9])";
        return RCG_ERR_INVALID;
      }
      {
        const uint32_t want = (uint32_t)std::min<int64_t>(WB_WARPS, std::max<int64_t>(1, (G.count + h->sm_count - 1) / h->sm_count));
        const uint32_t fit1 = (uint32_t)((int64_t)BC_SMEM_MAX / wb_warp_bytes(Wwb, L.capB, 1u));
        const uint32_t fit2 = (uint32_t)((int64_t)BC_SMEM_MAX / wb_warp_bytes(Wwb, L.capB, 2u));
        if (fit2 >= want || fit2 >= fit1) { nbuf = 2u; wbw = std::max(1u, std::min(want, fit2)); }
        else { nbuf = 1u; wbw = std::max(1u, std::min(want, fit1)); }
      }
      L.SA = wbw;   // (warp-per-block levels have no rings: the fields carry the warps per CTA and the buffers per warp)
      L.SB = nbuf;
      L.smem = (size_t)wbw * wb_warp_bytes(Wwb, L.capB, nbuf);
      L.wbocc = std::max<uint32_t>(1u, (uint32_t)((int64_t)BC_SMEM_MAX / (int64_t)(L.smem + 1024)));   // resident CTAs per SM
      L.groups = (uint32_t)((G.count + wbw - 1) / wbw);
      B.levels.push_back(L);
      continue;
    }
    const int64_t meanB = nchl ? sumB / nchl : 0;
    L.capA = (uint32_t)((maxA + 127) & ~127ll);
    // staging slot of ring B: a multiple of the level's mean blob (reserved[7], in quarters; default 3x); larger blobs
    // are read from HBM by their helper.  More, smaller slots put more chunks in flight (TMA latency cover).
    // (folded chain: the blobs carry the packed inverse, 4.3 KB each, so the slot is a smaller multiple of the mean: 1.5x)
    const int64_t capq = h->opt.reserved[7] > 0 ? h->opt.reserved[7] : (B.fold ? 6 : 12);
    int64_t capB = std::min<int64_t>(maxB, std::max<int64_t>(capq * meanB / 4, 6144));
    capB = std::min<int64_t>(capB, 24576);
    L.capB = (uint32_t)((capB + 127) & ~127ll);
    const int64_t fixed = (int64_t)W * 8 + 128 + BC_TR * 32 * 8 + (B.fold ? FC_SCR : BC_SCR) * 8 + BC_TR * 4 + 64 + 1024;
    int64_t avail = (int64_t)BC_SMEM_MAX - fixed;
    // ring A feeds one consumer (latency cover), ring B feeds BC_NH helpers that hold their slot while they work
    // (folded chain: the panel is capped at 8.3 KB per chunk -- chunk_fold_depth -- and a bulk copy from HBM takes 1-2 us,
    //  i.e. several hops: up to 12 slots)
    // (folded chain: three chain warps take turns and each preloads its next chunk, 3 ahead: at least 4 slots, 7 wanted)
    const int64_t SA_min = B.fold ? 4 : 3;
    int64_t SA = std::max<int64_t>(B.fold ? 7 : 3, std::min<int64_t>(B.fold ? 12 : 10, (avail * (B.fold ? 35 : 40) / 100) / L.capA));
    if (h->opt.reserved[8] > 0) SA = std::max<int64_t>(B.fold ? SA_min : 2, std::min<int64_t>(12, h->opt.reserved[8]));   // tuning experiments
    const int64_t NHELP = B.fold ? FC_NH : BC_NH;
    int64_t SB = std::max<int64_t>(3, std::min<int64_t>(2 * NHELP + 4, (avail - SA * L.capA) / (L.capB + 24)));
    while (SA > SA_min && SA * L.capA + SB * (L.capB + 24) + SA * 16 > avail) SA--;
    while (SB > 3 && SA * L.capA + SB * (L.capB + 24) + SA * 16 > avail) SB--;
    if (SA * L.capA + SB * (L.capB + 24) + SA * 16 > avail) {
      cudaFree(dgeom);
      h->err = "blocked solve: staging slots do not fit shared memory (window too large for this factor)";
      return RCG_ERR_INVALID;
    }
    L.SA = (uint32_t)SA; L.SB = (uint32_t)SB;
    L.smem = (size_t)(fixed - 1024 + SA * L.capA + SB * L.capB + (2 * SA + 2 * SB) * 8 + SB * 12);
    const uint32_t sms = (uint32_t)h->sm_count;
    L.Dfar = B.blocks_host[G.first].pad[0];
    L.groups = std::max(1u, std::min<uint32_t>((uint32_t)G.count, sms / 2u));
    L.helpers = std::max(1u, sms / L.groups - 1u);
    B.levels.push_back(L);
  }
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(dgeom);
  if (h->opt.chain_mode == 4) {   // cluster chain for the leaf level (128-row chunks, rcg_cluster.cuh)
    const int rc = cl_build(h, d, comb, max_depth);
    if (rc != RCG_OK) { rcg_free_csr(comb); return rc; }
  }
  ph.mark(h->stream, "  layout: levels, staging plan");
  {
    const int rc = dp_build(h, d, comb);
    if (rc != RCG_OK) { rcg_free_csr(comb); return rc; }
  }
  ph.mark(h->stream, "  layout: dp_build");
  rcg_free_csr(comb);
  ph.mark(h->stream, "  layout: free of the direction's CSR");
  if (!h->abort_flag) {
    RCG_CUDA(h, cudaMalloc(&h->abort_flag, sizeof(unsigned int) * 4));
    RCG_CUDA(h, cudaMemset(h->abort_flag, 0, sizeof(unsigned int) * 4));
  }
  B.on = true;
  return RCG_OK;
}

int rcg_check_abort(rcg_handle *h) {
  if (!h->abort_flag) return RCG_OK;
  unsigned int f = 0;
  RCG_CUDA(h, cudaMemcpyAsync(&f, h->abort_flag, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  if (f) {
    char buf[160];
    snprintf(buf, sizeof(buf), "blocked triangular solve: a dependency wait timed out (code 0x%x); results are invalid", f);
    h->err = buf;
    cudaMemsetAsync(h->abort_flag, 0, sizeof(unsigned int), h->stream);
    return RCG_ERR_CUDA;
  }
  return RCG_OK;
}

int rcg_launch_blocked(rcg_handle *h, DirectionDev &d, const double *rhs, double *out, const double *dotvec, int only_group,
                       int only_kernel) {
  BlockedDev &B = d.bc;
  if (!h->smem_optin_blocked) {   // per handle = per device: the attribute belongs to the device's instance of the kernel
    RCG_CUDA(h, cudaFuncSetAttribute(k_bc_solve<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BC_SMEM_MAX));
    RCG_CUDA(h, cudaFuncSetAttribute(k_bc_solve<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BC_SMEM_MAX));
    RCG_CUDA(h, cudaFuncSetAttribute(k_bc_solve<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BC_SMEM_MAX));
    RCG_CUDA(h, cudaFuncSetAttribute(k_bc_solve<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BC_SMEM_MAX));
    RCG_CUDA(h, cudaFuncSetAttribute(k_fc_solve<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BC_SMEM_MAX));
    RCG_CUDA(h, cudaFuncSetAttribute(k_fc_solve<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BC_SMEM_MAX));
    RCG_CUDA(h, cudaFuncSetAttribute(k_wb_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, BC_SMEM_MAX));
    h->smem_optin_blocked = true;
  }
  if (only_kernel > 0) return RCG_OK;   // the level kernel is the only kernel of a group
  RCG_CUDA(h, cudaMemsetAsync(B.flags, 0, sizeof(uint32_t) * ((size_t)B.ntiles + B.nblocks), h->stream));
  if (B.cl.on) RCG_CUDA(h, cudaMemsetAsync(B.cl.prog4, 0, sizeof(uint32_t) * 4 * (size_t)B.nblocks, h->stream));
  double *rz_part = h->partials + 2 * (size_t)h->partial_cap;
  const bool dist_fwd = h->dist.on && !d.reversed && h->N > h->dist.n_sub;
  const uint32_t dot_limit = h->dist.on ? h->dist.dot_limit : 0xFFFFFFFFu;
  bool coupled = false;
  for (size_t gi = 0; gi < d.groups.size(); gi++) {
    if (only_group >= 0 && (int)gi != only_group) continue;
    const GroupHost &G = d.groups[gi];
    const BcLevel &L = B.levels[gi];
    const bool top = dist_fwd && G.depth < h->dist.top_depth;
    if (top && !coupled) {
      for (const GroupHost &t : d.groups) {
        if (t.depth >= h->dist.top_depth) continue;
        const uint32_t gx = std::max(1u, std::min<uint32_t>((t.max_rows + 7u) / 8u, (uint32_t)h->sm_count * 8u / (uint32_t)t.count));
        dim3 grid(gx, (unsigned)t.count);
        k_bc_couple<<<grid, 256, 0, h->stream>>>(B.far.rowptr, B.far.col, B.far.val, B.blocks + t.first, out, h->dist.sbuf,
                                                h->dist.n_sub);
        h->stats.kernel_launches += 1;
      }
      RCG_CUDA(h, cudaGetLastError());
      RCG_TRY(rcg_allreduce_sum(h, h->dist.sbuf, h->N - h->dist.n_sub));
      coupled = true;
    }
    BcArgs a;
    memset(&a, 0, sizeof(a));
    a.blocks = B.blocks + G.first;
    a.nblocks = (uint32_t)G.count; a.ngroups = L.groups; a.helpers = L.helpers;
    a.offA = B.offA; a.offB = B.offB; a.blobA = B.blobA; a.blobB = B.blobB;
    a.far_rp = B.far.rowptr; a.far_col = B.far.col; a.far_val = B.far.val;
    a.tile_need = B.tile_need;
    a.far_split = B.far_split;
    a.tileflag = B.flags; a.gprog = B.flags + B.ntiles;
    a.w = B.w; a.rhs = rhs; a.out = out;
    a.dotvec = dotvec; a.dot_partials = dotvec ? rz_part : nullptr; a.dot_limit = dot_limit;
    a.N = (uint32_t)h->N; a.reversed = d.reversed ? 1 : 0;
    a.Kr = (B.blocks_host[G.first].pad[2] >> 8) & 0xFFu; a.E = B.blocks_host[G.first].pad[2] & 0xFFu; a.Dfar = L.Dfar; a.W = 32u * L.Dfar;
    a.SA = L.SA; a.SB = L.SB; a.capA = L.capA; a.capB = L.capB;
    a.tile = B.blocks_host[G.first].pad[1];
    a.far_lpr = (G.rows > 0 && G.ext_nnz / G.rows > 64) ? 32u : 8u;
    a.far_lpr2 = (h->opt.reserved[9] & 0xFF) == 32 ? 32u : 8u;   // reserved[9] bits 0-7: lanes per row of the in-block far pass
    a.col_min = top ? h->dist.n_sub : 0u;
    a.corr = top ? h->dist.sbuf : nullptr;
    a.abort_g = h->abort_flag;
    a.clk = h->clk_probe;
    a.dbg = (uint32_t)h->opt.reserved[1];
    if (L.dp) {   // dense-panel level: one launch (start vector of all rows, then the lock-step hops)
      RCG_TRY(dp_launch(h, B, a, gi));
      continue;
    }
    if (L.wb) {   // warp-per-block level: fully parallel pre-pass over the entries of other blocks, then one warp per block
      // pre-pass grid: (CTAs per block, blocks) -- about 16 CTAs per SM in total, a CTA takes 128 rows at a time
      const uint32_t pre_y = std::min<uint32_t>((uint32_t)G.count, 32768u);
      const uint32_t pre_x = std::max(1u, std::min<uint32_t>((G.max_rows + 127u) / 128u, ((uint32_t)h->sm_count * 16u + pre_y - 1u) / pre_y));
      a.ticket = B.flags + B.ntiles + B.nblocks + gi % 4;   // (four spare words behind the flags: consecutive levels never share one)
      RCG_CUDA(h, cudaMemsetAsync(a.ticket, 0, sizeof(uint32_t), h->stream));
      BcArgs ap = a;
      if (gi == 0 && !h->dist.on) ap.far_lpr2 = 0xFFFFFFFFu;   // no block is solved before the direction's first level: start = rhs
      k_wb_pre<<<dim3(pre_x, pre_y), 256, 0, h->stream>>>(ap);
      k_wb_solve<<<std::min<uint32_t>(L.groups, (uint32_t)h->sm_count * L.wbocc), L.SA * 32, L.smem, h->stream>>>(a);
      RCG_CUDA(h, cudaGetLastError());
      h->stats.kernel_launches += 2;
      continue;
    }
    if (B.cl.on && B.cl.level_on[gi]) {   // leaf level on the cluster chain
      RCG_TRY(cl_launch(h, B, a, G, gi));
      continue;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(L.groups * (1u + L.helpers));
    cfg.blockDim = dim3(B.fold ? FC_THREADS : BC_THREADS);
    cfg.dynamicSmemBytes = L.smem;
    cfg.stream = h->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = (h->opt.reserved[6] & 1) ? 0 : 1;   // reserved[6] = 1: plain launch (the grid never exceeds the SM count)
    const bool split = h->opt.chain_mode != 3;   // default: four critical warps
    void (*kern)(const BcArgs) = (a.dbg & 1u) ? (split ? k_bc_solve<4, true> : k_bc_solve<1, true>)
                                              : (split ? k_bc_solve<4, false> : k_bc_solve<1, false>);
    if (B.fold) kern = (a.dbg & 1u) ? k_fc_solve<true> : k_fc_solve<false>;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a);
    if (e != cudaSuccess && cfg.numAttrs == 1) {
      // co-residency is what the cooperative attribute guarantees; a context that cannot give it (or cannot capture it)
      // still runs the launch correctly when the GPU is otherwise idle, because the grid never exceeds the SM count
      cudaGetLastError();
      cfg.numAttrs = 0;
      e = cudaLaunchKernelEx(&cfg, kern, a);
    }
    RCG_CUDA(h, e);
    h->stats.kernel_launches += 1;
  }
  RCG_CUDA(h, cudaGetLastError());
  return RCG_OK;
}
