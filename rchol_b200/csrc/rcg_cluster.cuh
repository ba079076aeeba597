// rchol_b200 -- cluster chain (chain_mode 4): the blocked-inverse triangular solve of rcg_blocked.cu with 128-row chunks
// whose dense inverse is split row-wise over a thread-block cluster of 4 CTAs (measured hop cost:
// scripts/ubench/hop.cu, profiles/r01_ubench_hop_cluster.txt).  Included by rcg_blocked.cu inside its anonymous
// namespace: shares the memory-model primitives, Guard/BC_WAIT, BcArgs, the far tiles / start vector / progress words.
//
//   x_K = Winv_K ( start_K - near entries )          K = 128-row chunk of a block, Winv_K dense 128x128 (set-up)
//
//   chain CTA q of the cluster (rows 32q..32q+31 of every chunk), 256 compute threads + producer warp + publisher warp:
//     before x_{K-1} has arrived:  slab of Winv_K -> registers; "old" near entries (columns older than chunk K-1, inside
//                                  the window) gathered from the CTA's own copy of the solution window, jagged diagonals
//     after:                       "recent" entries (chunk K-1) -> t (all 128 rows, redundantly in every CTA) ->
//                                  slab mat-vec (thread = row x column group) -> 32 solved rows -> st.async into the window
//                                  of all four CTAs (data + mbarrier complete_tx, DSMEM)
//   far CTAs: identical to k_bc_solve's (tile flags, start vector P.w, progress words in 32-row units).
constexpr uint32_t CL_S = 4;                 // CTAs per cluster
constexpr uint32_t CL_NT = 256;              // compute threads
constexpr uint32_t CL_THREADS = 576;         // + producer warp (8) + publisher warp (9) + 8 helper warps (old entries, one hop ahead)
constexpr uint32_t CL_WCHUNK = 81920;        // bytes of the four slabs of one chunk: 8 + 16 + 24 + 32 KiB
constexpr uint32_t CL_WSTAGE = 32768;        // slab part of a staging slot
constexpr uint32_t CL_NHDR = 1552;           // near blob: 16 B header, perm_old[512] (u16), n_old[512] (u8)
constexpr uint32_t CL_RB = 6;                // recent entries per lane held in registers
constexpr uint32_t CL_SV = 1040;             // start vector of a chunk inside a staging slot: 130 doubles (16-byte aligned source)
__host__ __device__ __forceinline__ uint32_t cl_slab_off(uint32_t q) { return 4096u * q * (q + 1u); }

struct ClGeom {               // cluster-solved blocks in ascending chunk order
  const uint32_t *lo, *hi, *dfar, *c0;   // c0: nb+1 prefix of 128-row chunks
  int nb;
  uint32_t ring;
};

struct ClRow {                // row j of a chunk: [s,p_far) far | [p_far,p_old) old | [p_old,p_rec) recent | [p_rec,p_diag) own | diag
  int64_t s, p_far, p_old, p_rec, p_diag;
};

__device__ __forceinline__ ClRow cl_split(const int64_t *__restrict__ rp, const uint32_t *__restrict__ col, uint32_t j,
                                          uint32_t blo, uint32_t K, uint32_t Dfar) {
  ClRow r;
  r.s = rp[j];
  r.p_diag = rp[j + 1] - 1;
  const int k32 = (int)((j - blo) >> 5);
  const uint32_t c_far = blo + 32u * (uint32_t)max(0, k32 + 1 - (int)Dfar);   // same rule as split_row(): the far CSR is shared
  const uint32_t c_own = blo + 128u * K;
  const uint32_t c_rec = K > 0u ? c_own - 128u : c_own;
  int64_t p = r.s;
  while (p < r.p_diag && col[p] < c_far) p++;
  r.p_far = p;
  while (p < r.p_diag && col[p] < c_rec) p++;
  r.p_old = p;
  while (p < r.p_diag && col[p] < c_own) p++;
  r.p_rec = p;
  return r;
}

// Jagged-diagonal bookkeeping of one chunk (CTA of 128 threads, thread = row).  Old entries: 512 lane-rows (row, quarter),
// quarter h takes entries h, h+4, ...  Lane-rows are sorted by descending count; `rank` is the sorted position; diagonal s
// holds the s-th entry of every lane-row with more than s entries.  The solve assigns sorted positions t and 511-t to
// helper thread t (longest with shortest).
struct ClJds {
  uint32_t n_old[4], rank_old[4];
  uint32_t n_rec[2], rank_rec;        // recent entries: rows sorted by count, half h of a row takes entries h, h+2, ...
  uint32_t nd_old, nd_rec, tot_old, tot_rec;
};
__device__ __forceinline__ ClJds cl_jds(uint32_t cnt_old, uint32_t cnt_rec, uint32_t *s_cnt /*512*/, uint32_t *s_red /*4*/) {
  const uint32_t i = threadIdx.x;
  ClJds J;
  if (i < 4u) s_red[i] = 0u;
#pragma unroll
  for (uint32_t h = 0; h < 4u; h++) {
    J.n_old[h] = (cnt_old + 3u - h) >> 2;
    s_cnt[128u * h + i] = J.n_old[h];
    J.rank_old[h] = 0u;
  }
  __syncthreads();
  for (uint32_t l = 0; l < 512u; l++) {
    const uint32_t o = s_cnt[l];
#pragma unroll
    for (uint32_t h = 0; h < 4u; h++) J.rank_old[h] += (o > J.n_old[h] || (o == J.n_old[h] && l < 128u * h + i)) ? 1u : 0u;
  }
  atomicMax(&s_red[0], J.n_old[0]);
  atomicMax(&s_red[1], (cnt_rec + 1u) >> 1);
  atomicAdd(&s_red[2], cnt_old);
  atomicAdd(&s_red[3], cnt_rec);
  __syncthreads();
  J.nd_old = s_red[0]; J.nd_rec = s_red[1]; J.tot_old = s_red[2]; J.tot_rec = s_red[3];
  __syncthreads();
  s_cnt[i] = cnt_rec;
  __syncthreads();
  J.n_rec[0] = (cnt_rec + 1u) >> 1;
  J.n_rec[1] = cnt_rec >> 1;
  J.rank_rec = 0u;
  for (uint32_t l = 0; l < 128u; l++) {
    const uint32_t o = s_cnt[l];
    J.rank_rec += (o > cnt_rec || (o == cnt_rec && l < i)) ? 1u : 0u;
  }
  __syncthreads();
  return J;
}
// near blob: header | old: bases[nd_old] u16, values, window slots | recent: perm[128] u8, counts[256] u8 (2*rank + half),
// bases[2][nd_rec] u16, values, window slots
__host__ __device__ __forceinline__ uint32_t cl_near_bytes(uint32_t nd_old, uint32_t nd_rec, uint32_t tot_old, uint32_t tot_rec) {
  return CL_NHDR + r16(2u * nd_old) + r16(8u * tot_old) + r16(2u * tot_old) + 384u + r16(4u * nd_rec) + r16(8u * tot_rec) + r16(2u * tot_rec);
}

// CTA (128 threads) per chunk: size of the near blob
__global__ void __launch_bounds__(128) k_cl_count(const int64_t *__restrict__ rp, const uint32_t *__restrict__ col, ClGeom g,
                                                  uint32_t nchunks, int64_t *__restrict__ sizeN, int *err) {
  __shared__ uint32_t s_cnt[512], s_red[4];
  for (uint32_t gc = blockIdx.x; gc < nchunks; gc += gridDim.x) {
    const int b = find_le(g.c0, g.nb, gc);
    const uint32_t K = gc - g.c0[b], blo = g.lo[b], bhi = g.hi[b];
    const uint32_t j = blo + 128u * K + threadIdx.x;
    uint32_t cnt_old = 0, cnt_rec = 0;
    if (j < bhi) {
      const ClRow r = cl_split(rp, col, j, blo, K, g.dfar[b]);
      cnt_old = (uint32_t)(r.p_old - r.p_far);
      cnt_rec = (uint32_t)(r.p_rec - r.p_old);
    }
    if (cnt_old > 1020u || cnt_rec > 510u) atomicExch(err, 2);
    const ClJds J = cl_jds(cnt_old, cnt_rec, s_cnt, s_red);
    if (threadIdx.x == 0) {
      if (J.tot_old > 60000u || J.tot_rec > 60000u) atomicExch(err, 2);   // diagonal bases are 16-bit
      sizeN[gc] = (int64_t)cl_near_bytes(J.nd_old, J.nd_rec, J.tot_old, J.tot_rec);
    }
  }
}

// CTA (128 threads) per chunk: dense inverse of the 128x128 diagonal block -> four row slabs; near blob
constexpr uint32_t CL_OWNCAP = 6144;   // staged own-block entries of a chunk (more: read from global memory)
// W | oval | dval | ooff (130 words) | s_cnt | s_red | m_old | m_rec | ocolb
constexpr int CL_FILL_SMEM = 128 * 128 * 8 + CL_OWNCAP * 8 + 128 * 8 + 130 * 4 + 512 * 4 + 16 + 512 * 2 + 512 * 2 + CL_OWNCAP;
__global__ void __launch_bounds__(128) k_cl_fill(const int64_t *__restrict__ rp, const uint32_t *__restrict__ col,
                                                 const double *__restrict__ val, ClGeom g, uint32_t nchunks,
                                                 const int64_t *__restrict__ offN, unsigned char *__restrict__ wslab,
                                                 unsigned char *__restrict__ blobN) {
  extern __shared__ __align__(16) unsigned char fsm[];
  double *W = reinterpret_cast<double *>(fsm);                         // [128][128] row-major
  double *oval = W + 128 * 128;                                        // [CL_OWNCAP] staged own-block values
  double *dval = oval + CL_OWNCAP;                                     // [128] diagonal
  uint32_t *ooff = reinterpret_cast<uint32_t *>(dval + 128);           // [129] prefix of own-entry counts (130 words reserved)
  uint32_t *s_cnt = ooff + 130;                                        // [512]
  uint32_t *s_red = s_cnt + 512;                                       // [4]
  uint16_t *m_old = reinterpret_cast<uint16_t *>(s_red + 4);           // [512] diagonal lengths, then bases
  uint16_t *m_rec = m_old + 512;                                       // [512]
  unsigned char *ocolb = reinterpret_cast<unsigned char *>(m_rec + 512);   // [CL_OWNCAP] staged own-block columns (0..127)
  const uint32_t i = threadIdx.x;
  for (uint32_t gc = blockIdx.x; gc < nchunks; gc += gridDim.x) {
    const int b = find_le(g.c0, g.nb, gc);
    const uint32_t K = gc - g.c0[b], blo = g.lo[b], bhi = g.hi[b];
    const uint32_t row0 = blo + 128u * K;
    const uint32_t j = row0 + i;
    const bool valid = j < bhi;
    ClRow r;
    r.s = r.p_far = r.p_old = r.p_rec = r.p_diag = 0;
    if (valid) r = cl_split(rp, col, j, blo, K, g.dfar[b]);
    const uint32_t cnt_old = (uint32_t)(r.p_old - r.p_far), cnt_rec = (uint32_t)(r.p_rec - r.p_old);
    const uint32_t cnt_own = (uint32_t)(r.p_diag - r.p_rec);
    // ---- own entries staged in shared memory ------------------------------------------------------------
    ooff[i + 1u] = cnt_own;
    dval[i] = valid ? val[r.p_diag] : 1.0;
    if (i == 0u) ooff[0] = 0u;
    __syncthreads();
    if (i == 0u)
      for (uint32_t l = 0; l < 128u; l++) ooff[l + 1u] += ooff[l];
    __syncthreads();
    const bool staged = ooff[128] <= CL_OWNCAP;
    if (staged)
      for (uint32_t e = 0; e < cnt_own; e++) {
        oval[ooff[i] + e] = val[r.p_rec + e];
        ocolb[ooff[i] + e] = (unsigned char)(col[r.p_rec + e] - row0);
      }
    __syncthreads();
    // ---- W = inverse of the diagonal block: row l for all columns i in parallel ---------------------------
    for (uint32_t l = 0; l < 128u; l++) {
      double s = l == i ? 1.0 : 0.0;
      if (i <= l) {
        if (staged) {
          for (uint32_t e = ooff[l]; e < ooff[l + 1u]; e++) s = fma(-oval[e], W[(uint32_t)ocolb[e] * 128u + i], s);
        } else if (ooff[l + 1u] > ooff[l]) {
          const int64_t p0 = rp[row0 + l + 1u] - 1 - (int64_t)(ooff[l + 1u] - ooff[l]);
          for (uint32_t e = 0; e < ooff[l + 1u] - ooff[l]; e++) s = fma(-val[p0 + e], W[(col[p0 + e] - row0) * 128u + i], s);
        }
        s = s / dval[l];
      } else {
        s = 0.0;
      }
      W[l * 128u + i] = s;
      __syncthreads();
    }
    // ---- slabs: rank q, [column c < 32(q+1)][row r < 32] ---------------------------------------------------
    {
      double *dst = reinterpret_cast<double *>(wslab + (size_t)gc * CL_WCHUNK);
      for (uint32_t q = 0; q < CL_S; q++) {
        double *sl = dst + cl_slab_off(q) / 8u;
        const uint32_t n = 32u * (q + 1u) * 32u;
        for (uint32_t e = i; e < n; e += 128u) sl[e] = W[(32u * q + (e & 31u)) * 128u + (e >> 5)];
      }
    }
    // ---- near blob ----------------------------------------------------------------------------------------
    const ClJds J = cl_jds(cnt_old, cnt_rec, s_cnt, s_red);
    unsigned char *Nb = blobN + offN[gc];
    for (uint32_t e = i; e < 512u; e += 128u) { m_old[e] = 0; m_rec[e] = 0; }
    __syncthreads();
    // diagonal lengths (shared-memory atomics on 32-bit words: two u16 per word, counts stay below 65536)
    {
      uint32_t *mo32 = reinterpret_cast<uint32_t *>(m_old);
      for (uint32_t h = 0; h < 4u; h++)
        for (uint32_t s = 0; s < J.n_old[h]; s++) atomicAdd(&mo32[s >> 1], (s & 1u) ? 65536u : 1u);
      uint32_t *mr32 = reinterpret_cast<uint32_t *>(m_rec);   // [2][256]: half h, diagonal s
      for (uint32_t h = 0; h < 2u; h++)
        for (uint32_t s = 0; s < J.n_rec[h]; s++) atomicAdd(&mr32[(256u * h + s) >> 1], (s & 1u) ? 65536u : 1u);
    }
    __syncthreads();
    if (i == 0u) {   // lengths -> bases (exclusive prefix)
      uint32_t a = 0;
      for (uint32_t s = 0; s < J.nd_old; s++) { const uint32_t m = m_old[s]; m_old[s] = (uint16_t)a; a += m; }
      a = 0;
      for (uint32_t h = 0; h < 2u; h++)
        for (uint32_t s = 0; s < J.nd_rec; s++) { const uint32_t m = m_rec[256u * h + s]; m_rec[256u * h + s] = (uint16_t)a; a += m; }
      uint32_t *hd = reinterpret_cast<uint32_t *>(Nb);
      hd[0] = J.nd_old; hd[1] = J.nd_rec; hd[2] = J.tot_old; hd[3] = J.tot_rec;
    }
    __syncthreads();
    {
      uint16_t *g_bold = reinterpret_cast<uint16_t *>(Nb + CL_NHDR);   // first entry of every diagonal
      for (uint32_t s2 = i; s2 < J.nd_old; s2 += 128u) g_bold[s2] = m_old[s2];
      uint16_t *perm_old = reinterpret_cast<uint16_t *>(Nb + 16u);
      unsigned char *n_old = Nb + 1040u;
      unsigned char *p = Nb + CL_NHDR + r16(2u * J.nd_old);
      double *v_old = reinterpret_cast<double *>(p);
      p += r16(8u * J.tot_old);
      uint16_t *c_old = reinterpret_cast<uint16_t *>(p);
      p += r16(2u * J.tot_old);
      unsigned char *perm_rec = p, *n_rec = p + 128u;                    // rank -> row; counts indexed 2*rank + half
      uint16_t *g_brec = reinterpret_cast<uint16_t *>(p + 384u);           // [2][nd_rec]
      p += 384u + r16(4u * J.nd_rec);
      double *v_rec = reinterpret_cast<double *>(p);
      uint16_t *c_rec = reinterpret_cast<uint16_t *>(p + r16(8u * J.tot_rec));
      for (uint32_t h = 0; h < 4u; h++) {
        perm_old[J.rank_old[h]] = (uint16_t)(i + 128u * h);   // lane-row id: row + 128*quarter
        n_old[J.rank_old[h]] = (unsigned char)J.n_old[h];
        for (uint32_t s = 0; s < J.n_old[h]; s++) {
          const int64_t e = r.p_far + h + 4u * s;
          const uint32_t pos = (uint32_t)m_old[s] + J.rank_old[h];
          v_old[pos] = val[e];
          c_old[pos] = (uint16_t)((col[e] - blo) % g.ring);
        }
      }
      perm_rec[J.rank_rec] = (unsigned char)i;
      for (uint32_t h = 0; h < 2u; h++) {
        n_rec[2u * J.rank_rec + h] = (unsigned char)J.n_rec[h];
        for (uint32_t s = i; s < J.nd_rec; s += 128u) g_brec[h * J.nd_rec + s] = m_rec[256u * h + s];
        for (uint32_t s = 0; s < J.n_rec[h]; s++) {
          const int64_t e = r.p_old + h + 2u * s;
          const uint32_t pos = (uint32_t)m_rec[256u * h + s] + J.rank_rec;
          v_rec[pos] = val[e];
          c_rec[pos] = (uint16_t)((col[e] - blo) % g.ring);
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// cluster primitives
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cl_mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ uint32_t cl_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cl_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool cl_mbar_try_cluster(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void cl_st_async(uint32_t remote_addr, double v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr),
               "l"(__double_as_longlong(v)), "r"(remote_bar)
               : "memory");
}

struct ClArgs {
  BcArgs f;                       // blocks / far tiles / start vector: same meaning as in k_bc_solve
  const unsigned char *wslab, *blobN;
  const int64_t *offN;
  const uint32_t *c0;             // per block (gidx): first 128-row chunk
  uint32_t *prog4;                // per block (gidx) and cluster rank: published hops (the far CTAs wait for all four)
  uint32_t ring;                  // rows of the window ring
  uint32_t smem_total, capN;      // dynamic shared memory of the launch, bytes of the near part of a staging slot
};

__device__ __forceinline__ bool cl_prog_reached(const uint32_t *p4, uint32_t hops) {
  return ld_acquire_gpu(p4) >= hops && ld_acquire_gpu(p4 + 1) >= hops && ld_acquire_gpu(p4 + 2) >= hops && ld_acquire_gpu(p4 + 3) >= hops;
}

// far CTA of the cluster launch: the far role of k_bc_solve (kept in step with it), any number of warps; the chain's
// progress is one word per cluster rank, counted in hops of 128 rows
__device__ __forceinline__ void cl_far_role(const BcArgs &P, const uint32_t *prog4, uint32_t grp, uint32_t hid, Guard &G) {
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const uint32_t lpr_pass[2] = {P.far_lpr, P.far_lpr2};
  for (uint32_t bi = grp; bi < P.nblocks; bi += P.ngroups) {
    const BcBlock b = P.blocks[bi];
    const uint32_t nch = (b.hi - b.lo + 31u) >> 5, ntile = (nch + P.tile - 1u) / P.tile;
    constexpr uint32_t LA = 2;
    const uint32_t nown = hid < ntile ? (ntile - hid + P.helpers - 1u) / P.helpers : 0u;
    for (uint32_t it = 0; it < nown + LA; it++) {
      for (uint32_t pass = 0; pass < 2u; pass++) {
        if (pass == 0u ? it >= nown : it < LA) continue;
        const uint32_t t = hid + (pass == 0u ? it : it - LA) * P.helpers;
        const uint32_t need = P.tile_need[b.tile0 + t];
        if (pass == 1u) {
          if (need == 0u) continue;
          if (threadIdx.x == 0) BC_WAIT(cl_prog_reached(prog4 + 4u * b.gidx, (need + 3u) >> 2), 0x1100u, 200);
          __syncthreads();
        }
        const uint32_t lpr = lpr_pass[pass], rpw = 32u / lpr, sub = lane & (lpr - 1u);
        const uint32_t r0 = b.lo + t * (32u * P.tile), r1 = min(b.hi, r0 + 32u * P.tile);
        for (uint32_t base = r0 + warp * rpw; base < r1; base += nwarp * rpw) {
          const uint32_t j = base + lane / lpr;
          const bool valid = j < r1;
          double acc = 0.0;
          if (valid) {
            const int64_t es = P.far_rp[j] + P.far_split[j];
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
              s = __ldcg(P.w + j);
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
}

__global__ void __launch_bounds__(CL_THREADS, 1) k_cl_solve(const ClArgs A) {
  const BcArgs &P = A.f;
  extern __shared__ __align__(128) unsigned char smem[];
  // ---- shared-memory carve-up ------------------------------------------------------------------------------
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem);            // [0,2) x, [2,4) told ready, [4,4+NS) full, [4+NS,4+2NS) empty
  uint32_t *ctl = reinterpret_cast<uint32_t *>(smem + 128);       // [0] hops landed in the window, [1] hops published, [2] abort
  double *dotbuf = reinterpret_cast<double *>(smem + 192);        // [4] fused dot product of the four ranks (rank 0's copy is used)
  double *told = reinterpret_cast<double *>(smem + 256);          // [2][512] partial sums of the old entries per lane-row
  double *tfull = told + 1024;                                    // [128] t of the chunk
  double *part = tfull + 128;                                     // [8][32] partial sums of the mat-vec
  double *dummy = part + 256;                                     // [128] target of the block-start round
  double *win = dummy + 128;                                      // [ring] solution window (every CTA holds all of it); win[ring] = 0
  unsigned char *stages = reinterpret_cast<unsigned char *>(win + A.ring + 16);
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const uint32_t ring_hops = A.ring >> 7;   // a window slot is reused after this many hops

  Guard G;
  G.abort_s = smem_u32(ctl + 2);
  G.abort_g = P.abort_g;
  G.n = 0;
  G.t0 = 0;
  if (tid == 0) { ctl[0] = 0u; ctl[1] = 0u; ctl[2] = 0u; }
  if (tid < 16u) win[A.ring + tid] = 0.0;

  if (blockIdx.x >= CL_S * P.ngroups) {
    // ======================================== far CTA ==========================================================
    __syncthreads();
    const uint32_t f = blockIdx.x - CL_S * P.ngroups;
    const uint32_t hid = f / P.ngroups, grp = f % P.ngroups;
    if (hid < P.helpers) cl_far_role(P, A.prog4, grp, hid, G);
    return;
  }

  // ========================================== chain CTA ==========================================================
  const uint32_t q = cl_ctarank();
  const uint32_t grp = blockIdx.x / CL_S;
  // staging slots of this rank: its slab is 8(q+1) KiB, so the low ranks get more slots than rank 3
  const uint32_t wpart = 8192u * (q + 1u);
  const uint32_t stage_bytes = wpart + CL_SV + A.capN;
  const uint32_t NS = min(6u, (A.smem_total - (uint32_t)((const unsigned char *)stages - smem)) / stage_bytes);
  uint64_t *tbar = bars + 2, *fullb = bars + 4, *emptyb = bars + 4 + NS;
  const uint32_t xbar0 = smem_u32(&bars[0]);
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&tbar[0], 8);
    mbar_init(&tbar[1], 8);
    for (uint32_t s = 0; s < NS; s++) {
      mbar_init(&fullb[s], 1);
      mbar_init(&emptyb[s], 16);   // 8 compute warps + 8 helper warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(&bars[0], 1024);   // rounds 0 and 1
    mbar_expect_tx(&bars[1], 1024);
  }
  __syncthreads();
  cl_cluster_sync();

  uint32_t step = 0;   // exchange rounds before this block (round n: bars[n & 1], phase (n >> 1) & 1); compute threads count on
  uint32_t hs = 0;     // hops before this block (staging ring / told buffers); every role counts alike
  const uint32_t ncol = 4u * (q + 1u);   // columns of the slab per column group (8 groups)
  const uint32_t wbytes = 8192u * (q + 1u);
  const uint32_t prof_cta = (P.dbg & 2u) ? 3u : 0u;   // dbg bit 1: profile rank 3 of cluster 0 instead of rank 0
  const bool prof_on = (P.dbg & 1u) != 0u && P.clk != nullptr && blockIdx.x == prof_cta;
  const bool prof = prof_on && tid == 0;
  long long ph[5] = {0, 0, 0, 0, 0};   // helper thread 0: stage wait, x wait, gathers; producer: tile flags, free slot
  long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // cycles of thread 0: stage wait, loads, told wait, x wait, recent, mat-vec, reduce+send, hops
  // remote addresses of the exchange: this thread's window slot and the x barriers in all four CTAs
  uint32_t rbar[CL_S], rwin[CL_S];
#pragma unroll
  for (uint32_t p = 0; p < CL_S; p++) { rbar[p] = cl_mapa(xbar0, p); rwin[p] = cl_mapa(smem_u32(win), p); }

  for (uint32_t bi = grp; bi < P.nblocks; bi += P.ngroups) {
    const BcBlock b = P.blocks[bi];
    const uint32_t rows = b.hi - b.lo, nhop = (rows + 127u) >> 7, nch32 = (rows + 31u) >> 5;
    const uint32_t c0 = A.c0[b.gidx];

    if (tid < CL_NT) {
      // ------------------------------------------ compute threads ------------------------------------------------
      const uint32_t g = warp, r = lane;
      const uint32_t rrank = tid >> 1, half = tid & 1u;   // recent entries: lane pair per row, rows in sorted order
      if (tid < 32u) {   // block-start round: nothing to wait for, but every round has the same shape
        const uint32_t off = (uint32_t)((const unsigned char *)&dummy[32u * q + tid] - (const unsigned char *)win);
#pragma unroll
        for (uint32_t p = 0; p < CL_S; p++) cl_st_async(rwin[p] + off, 0.0, rbar[p] + 8u * (step & 1u));
      }
      uint32_t pos = 32u * q + tid;   // window slot of this thread's row of the current chunk (threads 0..31)
      for (uint32_t K = 0; K < nhop; K++) {
        const uint32_t hp = hs + K;
        long long c_0 = 0, c_1 = 0, c_2 = 0, c_3 = 0, c_4 = 0, c_5 = 0, c_6 = 0;
        if (prof) c_0 = clock64();
        const uint32_t slot = hp % NS;
        const unsigned char *St = stages + (size_t)slot * stage_bytes;
        BC_WAIT(mbar_try(&fullb[slot], (hp / NS) & 1u), 0x1200u, 0);
        if (prof) c_1 = clock64();
        const double *sv = reinterpret_cast<const double *>(St + wpart);   // start vector, staged by the producer
        const unsigned char *Nb = St + wpart + CL_SV;
        const uint32_t nd_old = reinterpret_cast<const uint32_t *>(Nb)[0], nd_rec = reinterpret_cast<const uint32_t *>(Nb)[1];
        const uint32_t tot_old = reinterpret_cast<const uint32_t *>(Nb)[2], tot_rec = reinterpret_cast<const uint32_t *>(Nb)[3];
        const unsigned char *prec = Nb + CL_NHDR + r16(2u * nd_old) + r16(8u * tot_old) + r16(2u * tot_old);
        const uint32_t row = prec[rrank], nrec = prec[128u + tid];
        // start value of the row (even lane of the pair)
        double wst = 0.0;
        {
          const uint32_t j0 = b.lo + 128u * K;
          if (half == 0u && j0 + row < b.hi) wst = sv[(j0 & 1u) + row];
        }
        // slab of Winv_K: this thread's columns [g*ncol, (g+1)*ncol), row r
        double wreg[16];
        {
          const double *Ws = reinterpret_cast<const double *>(St) + (size_t)(g * ncol) * 32u + r;
#pragma unroll
          for (uint32_t i = 0; i < 16u; i++) wreg[i] = i < ncol ? Ws[i * 32u] : 0.0;
        }
        // recent entries of the lane pair's row: values and window slots into registers
        const uint16_t *b_rec = reinterpret_cast<const uint16_t *>(prec + 384u) + half * nd_rec;
        const double *v_rec = reinterpret_cast<const double *>(prec + 384u + r16(4u * nd_rec)) + rrank;
        const uint16_t *c_rec = reinterpret_cast<const uint16_t *>(prec + 384u + r16(4u * nd_rec) + r16(8u * tot_rec)) + rrank;
        double rv[CL_RB];
        uint32_t rk[CL_RB];
#pragma unroll
        for (uint32_t u = 0; u < CL_RB; u++) {
          rv[u] = 0.0;
          rk[u] = A.ring;
          if (u < nrec) { const uint32_t e = b_rec[u]; rv[u] = v_rec[e]; rk[u] = c_rec[e]; }
        }
        if (prof) c_2 = clock64();
        // old entries of this chunk (helper warps, one hop ahead)
        BC_WAIT(mbar_try(&tbar[hp & 1u], (hp >> 1) & 1u), 0x1900u, 0);
        const double *to = told + 512u * (hp & 1u);
        double tb = wst - ((to[row] + to[128u + row]) + (to[256u + row] + to[384u + row]));   // (odd lane: not used)
        // own rows of chunk K - ring_hops must have left the window (publisher of this CTA) before they are overwritten
        if (tid == 0u && K >= ring_hops) BC_WAIT(ld_acquire_cta_s(smem_u32(ctl + 1)) + ring_hops > K, 0x1800u, 0);
        if (prof) c_3 = clock64();
        // x of chunk K-1 (or the block-start round)
        BC_WAIT(cl_mbar_try_cluster(xbar0 + 8u * (step & 1u), (step >> 1) & 1u), 0x1300u, 0);
        if (tid == CL_NT - 1u) mbar_expect_tx(&bars[step & 1u], 1024);   // round step + 2
        if (tid == 0u) st_release_cta_s(smem_u32(ctl), K);
        step++;
        if (prof) c_4 = clock64();
        {
          double xv[CL_RB];
#pragma unroll
          for (uint32_t u = 0; u < CL_RB; u++) xv[u] = win[rk[u]];
          double a0 = 0.0, a1 = 0.0;
#pragma unroll
          for (uint32_t u = 0; u < CL_RB; u += 2u) {
            a0 = fma(rv[u], xv[u], a0);
            a1 = fma(rv[u + 1u], xv[u + 1u], a1);
          }
          for (uint32_t u = CL_RB; u < nrec; u++) { const uint32_t e = b_rec[u]; a0 = fma(v_rec[e], win[c_rec[e]], a0); }
          a0 += a1;
          a0 += __shfl_xor_sync(0xffffffffu, a0, 1);
          if (half == 0u) tfull[row] = tb - a0;
        }
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (prof) c_5 = clock64();
        {
          const double2 *tv = reinterpret_cast<const double2 *>(tfull + g * ncol);
          double a0 = 0.0, a1 = 0.0;
#pragma unroll
          for (uint32_t i = 0; i < 16u; i += 2u) {
            if (i < ncol) {
              const double2 t2 = tv[i >> 1];
              a0 = fma(wreg[i], t2.x, a0);
              a1 = fma(wreg[i + 1u], t2.y, a1);
            }
          }
          part[g * 32u + r] = a0 + a1;
        }
        __syncwarp();
        if (lane == 0u) mbar_arrive(&emptyb[slot]);   // every read of the staging slot by this warp is done
        asm volatile("bar.sync 3, 256;" ::: "memory");
        if (prof) c_6 = clock64();
        if (tid < 32u) {
          const double x = ((part[tid] + part[32u + tid]) + (part[64u + tid] + part[96u + tid])) +
                           ((part[128u + tid] + part[160u + tid]) + (part[192u + tid] + part[224u + tid]));
          const uint32_t nb8 = 8u * (step & 1u);
#pragma unroll
          for (uint32_t p = 0; p < CL_S; p++) cl_st_async(rwin[p] + 8u * pos, x, rbar[p] + nb8);
          pos += 128u;
          if (pos >= A.ring) pos -= A.ring;
        }
        if (prof) {
          pc[0] += c_1 - c_0; pc[1] += c_2 - c_1; pc[2] += c_3 - c_2; pc[3] += c_4 - c_3; pc[4] += c_5 - c_4; pc[5] += c_6 - c_5;
          pc[6] += clock64() - c_6; pc[7] += 1;
        }
      }
      // the last chunk's rows must have landed before the block is finished
      BC_WAIT(cl_mbar_try_cluster(xbar0 + 8u * (step & 1u), (step >> 1) & 1u), 0x1400u, 0);
      if (tid == CL_NT - 1u) mbar_expect_tx(&bars[step & 1u], 1024);
      if (tid == 0u) st_release_cta_s(smem_u32(ctl), nhop);
      step++;
    } else if (warp == 8u) {
      // ------------------------------------------ TMA producer -------------------------------------------------------
      uint32_t hp = hs;
      uint32_t tiles_known = 0;
      const uint32_t ntile = (nch32 + P.tile - 1u) / P.tile;
      for (uint32_t base = 0; base < nhop; base += 31u) {
        // blob offsets of 31 chunks (32 boundaries) with one coalesced load
        const int64_t o_mine = A.offN[c0 + min(base + lane, nhop)];
        for (uint32_t l = 0; l < 31u && base + l < nhop; l++, hp++) {
          const uint32_t K = base + l;
          const uint32_t slot = hp % NS, use = hp / NS;
          const int64_t o0 = __shfl_sync(0xffffffffu, o_mine, (int)l), o1 = __shfl_sync(0xffffffffu, o_mine, (int)l + 1);
          // the start vector of the chunk's rows must be complete: far tiles covering 32-row chunks 4K .. 4K+3
          const uint32_t tl = min(4u * K + 3u, nch32 - 1u) / P.tile;
          long long q0 = 0;
          if (prof_on) q0 = clock64();
          while (tiles_known <= tl) {   // lanes look at consecutive tile flags, the leading run of set flags is taken
            const uint32_t t = tiles_known + lane;
            const bool set = t < ntile && ld_acquire_gpu(P.tileflag + b.tile0 + t) != 0u;
            const uint32_t run = (uint32_t)__ffs((int)~__ballot_sync(0xffffffffu, set)) - 1u;   // all 32 set: ffs(0) - 1 wraps
            tiles_known += (run > 32u) ? 32u : run;
            if (tiles_known <= tl) {
              if (guard_poll(G, 0x1600u)) break;
              __nanosleep(100);
            }
          }
          G.n = 0;
          long long q1 = 0;
          if (prof_on) q1 = clock64();
          if (lane == 0u) {
            if (use > 0u) BC_WAIT(mbar_try(&emptyb[slot], (use - 1u) & 1u), 0x1500u, 20);
            if (prof_on) { ph[3] += q1 - q0; ph[4] += clock64() - q1; }
            const uint32_t nbytes = (uint32_t)(o1 - o0);
            unsigned char *St = stages + (size_t)slot * stage_bytes;
            // start vector of the chunk's rows: written by the far CTAs (generic proxy, released with the tile flags that
            // were acquired above), read here through the async proxy; 16-byte aligned source, even number of doubles
            const uint32_t j0 = b.lo + 128u * K, al = j0 & 1u;
            const uint32_t svbytes = 8u * ((min(128u, b.hi - j0) + al + 1u) & ~1u);
            asm volatile("fence.proxy.async.global;" ::: "memory");
            mbar_expect_tx(&fullb[slot], wbytes + nbytes + svbytes);
            bulk_g2s(St, A.wslab + (size_t)(c0 + K) * CL_WCHUNK + cl_slab_off(q), wbytes, &fullb[slot]);
            bulk_g2s(St + wpart, P.w + (j0 - al), svbytes, &fullb[slot]);
            bulk_g2s(St + wpart + CL_SV, A.blobN + o0, nbytes, &fullb[slot]);
          }
          __syncwarp();
        }
      }
    } else if (warp == 9u) {
      // ------------------------------------------ publisher: this rank's 32 rows of every chunk ------------------------
      uint32_t done = 0;
      double dot = 0.0;
      while (done < nhop) {
        uint32_t p = done;
        BC_WAIT((p = ld_acquire_cta_s(smem_u32(ctl))) > done, 0x1700u, 100);
        if (p <= done) break;   // aborted
        for (uint32_t k = done; k < p; k += 8u) {   // up to 8 hops at once: loads first
          double x[8], dv[8];
          uint32_t idx[8];
          bool ok[8];
#pragma unroll
          for (uint32_t u = 0; u < 8u; u++) {
            const uint32_t ro = 128u * (k + u) + 32u * q + lane;
            const uint32_t j = b.lo + ro;
            ok[u] = (k + u < p) && j < b.hi;
            idx[u] = P.reversed ? P.N - 1u - j : j;
            x[u] = ok[u] ? win[ro % A.ring] : 0.0;
            dv[u] = (ok[u] && P.dotvec && idx[u] < P.dot_limit) ? P.dotvec[idx[u]] : 0.0;
          }
#pragma unroll
          for (uint32_t u = 0; u < 8u; u++) {
            if (ok[u]) {
              P.out[idx[u]] = x[u];
              dot = fma(x[u], dv[u], dot);
            }
          }
        }
        __syncwarp();
        if (lane == 0u) {
          st_release_cta_s(smem_u32(ctl + 1), p);   // window slots of these hops may be overwritten
          __threadfence();
          st_release_gpu(A.prog4 + 4u * b.gidx + q, p);
        }
        done = p;
      }
      dot = warp_sum(dot);
      if (lane == 0u) {   // into rank 0's shared memory; summed in rank order after the cluster barrier below
        const uint32_t dst = cl_mapa(smem_u32(&dotbuf[q]), 0u);
        asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(dst), "d"(dot) : "memory");
      }
    } else {
      // ------------------------------------------ helpers: old near entries, one hop ahead ---------------------------
      // x of chunk K-2 is the newest value an old entry of chunk K can read: it has landed when round step + K - 1 is
      // complete (round `step` is the block-start round), i.e. while the compute threads still work on chunk K-1
      const uint32_t ht = tid - 320u;   // sorted lane-rows ht (long) and 511 - ht (short)
      for (uint32_t K = 0; K < nhop; K++) {
        const uint32_t hp = hs + K;
        const uint32_t slot = hp % NS;
        const unsigned char *Nb = stages + (size_t)slot * stage_bytes + wpart + CL_SV;
        long long h0 = 0, h1 = 0, h2 = 0;
        if (prof_on) h0 = clock64();
        BC_WAIT(mbar_try(&fullb[slot], (hp / NS) & 1u), 0x1A00u, 0);
        if (prof_on) h1 = clock64();
        const uint32_t nd_old = reinterpret_cast<const uint32_t *>(Nb)[0], tot_old = reinterpret_cast<const uint32_t *>(Nb)[2];
        const uint16_t *b_old = reinterpret_cast<const uint16_t *>(Nb + CL_NHDR);
        const double *v_old = reinterpret_cast<const double *>(Nb + CL_NHDR + r16(2u * nd_old));
        const uint16_t *c_old = reinterpret_cast<const uint16_t *>(Nb + CL_NHDR + r16(2u * nd_old) + r16(8u * tot_old));
        const uint16_t *perm = reinterpret_cast<const uint16_t *>(Nb + 16u);
        const uint32_t lrA = perm[ht], lrB = perm[511u - ht];
        const uint32_t nA = Nb[1040u + ht], nB = Nb[1040u + 511u - ht];
        constexpr uint32_t HB = 8;   // entries of the long lane-row held in registers before x arrives
        double hv[HB];
        uint32_t hk[HB];
#pragma unroll
        for (uint32_t u = 0; u < HB; u++) {
          hv[u] = 0.0;
          hk[u] = A.ring;
          if (u < nA) { const uint32_t e = b_old[u] + ht; hv[u] = v_old[e]; hk[u] = c_old[e]; }
        }
        double bv = 0.0;
        uint32_t bk = A.ring;
        if (nB > 0u) { const uint32_t e = b_old[0] + 511u - ht; bv = v_old[e]; bk = c_old[e]; }
        if (prof_on) h2 = clock64();
        if (K >= 1u) {
          const uint32_t rn = step + K - 1u;
          BC_WAIT(cl_mbar_try_cluster(xbar0 + 8u * (rn & 1u), (rn >> 1) & 1u), 0x1B00u, 0);
        }
        long long h3 = 0;
        if (prof_on) h3 = clock64();
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        {
          double xv[HB];
#pragma unroll
          for (uint32_t u = 0; u < HB; u++) xv[u] = win[hk[u]];
#pragma unroll
          for (uint32_t u = 0; u < HB; u += 4u) {
            a0 = fma(hv[u], xv[u], a0);
            a1 = fma(hv[u + 1u], xv[u + 1u], a1);
            a2 = fma(hv[u + 2u], xv[u + 2u], a2);
            a3 = fma(hv[u + 3u], xv[u + 3u], a3);
          }
        }
        double bsum = bv * win[bk];
        for (uint32_t s = HB; s < nA; s++) {
          const uint32_t e0 = b_old[s] + ht;
          a0 = fma(v_old[e0], win[c_old[e0]], a0);
        }
        for (uint32_t s2 = 1; s2 < nB; s2++) {
          const uint32_t e0 = b_old[s2] + 511u - ht;
          bsum = fma(v_old[e0], win[c_old[e0]], bsum);
        }
        told[512u * (hp & 1u) + lrA] = (a0 + a1) + (a2 + a3);
        told[512u * (hp & 1u) + lrB] = bsum;
        __syncwarp();
        if (lane == 0u) {
          mbar_arrive(&tbar[hp & 1u]);
          mbar_arrive(&emptyb[slot]);
        }
        if (prof_on) { ph[0] += h1 - h0; ph[1] += h3 - h2; ph[2] += (h2 - h1) + (clock64() - h3); }
      }
      step += nhop + 1u;
    }
    hs += nhop;
    __syncthreads();
    cl_cluster_sync();
    if (tid == 0) {
      ctl[0] = 0u;
      ctl[1] = 0u;
      if (q == 0u && P.dot_partials) P.dot_partials[b.gidx] = ((dotbuf[0] + dotbuf[1]) + dotbuf[2]) + dotbuf[3];
    }
    __syncthreads();
    cl_cluster_sync();   // dotbuf of rank 0 is free again
  }
  if (prof) {
    for (int i = 0; i < 8; i++) P.clk[3 + i] = (unsigned long long)pc[i];
    P.clk[2] = NS;
  }
  if (prof_on && tid == 320u)
    for (int i = 0; i < 3; i++) P.clk[11 + i] = (unsigned long long)ph[i];
  if (prof_on && tid == 256u) { P.clk[14] = (unsigned long long)ph[3]; P.clk[15] = (unsigned long long)ph[4]; }
  __syncthreads();
  cl_cluster_sync();   // nobody leaves while a peer may still write into its shared memory
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
constexpr int64_t CL_FIXED_SMEM = 256 + 8 * (1024 + 128 + 256 + 128);   // barriers, control words, told / tfull / part / dummy

static size_t cl_smem_bytes(uint32_t ring, uint32_t NS, uint32_t capN) {
  return (size_t)CL_FIXED_SMEM + ((size_t)ring + 16) * 8 + (size_t)NS * (CL_WSTAGE + CL_SV + capN);
}

// Builds the cluster layout of the leaf-level blocks (called from rcg_build_blocked while `comb` is alive).
static int cl_build(rcg_handle *h, DirectionDev &d, const CsrDev &comb, int max_depth) {
  BlockedDev &B = d.bc;
  ClusterDev &C = B.cl;
  C = ClusterDev();
  C.level_on.assign(d.groups.size(), 0);
  C.level_cap.assign(d.groups.size(), 0);
  C.ring = 32u * B.Dfar + 128u;
  if (C.ring > 65535u || B.Dfar < 8u) return RCG_OK;
  std::vector<uint32_t> lo, hi, dfar, c0(1, 0), c0_of_block(B.blocks_host.size(), 0xFFFFFFFFu);
  std::vector<int> level_of;   // cluster block -> level
  for (size_t gi = 0; gi < d.groups.size(); gi++) {
    const GroupHost &G = d.groups[gi];
    if (G.depth != max_depth) continue;
    for (int bi = G.first; bi < G.first + G.count; bi++) {
      const BcBlock &bd = B.blocks_host[bi];
      if (bd.pad[0] != B.Dfar) continue;
      c0_of_block[bd.gidx] = c0.back();
      lo.push_back(bd.lo); hi.push_back(bd.hi); dfar.push_back(bd.pad[0]);
      c0.push_back(c0.back() + (bd.hi - bd.lo + 127u) / 128u);
      level_of.push_back((int)gi);
    }
    C.level_on[gi] = 1;
  }
  const int nb = (int)lo.size();
  C.nchunks = c0.back();
  if (nb == 0 || C.nchunks == 0) { C.level_on.assign(d.groups.size(), 0); return RCG_OK; }
  uint32_t *dgeom = nullptr;   // lo | hi | dfar | c0
  RCG_CUDA(h, cudaMalloc(&dgeom, sizeof(uint32_t) * (4 * (size_t)nb + 1)));
  RCG_CUDA(h, cudaMemcpyAsync(dgeom, lo.data(), sizeof(uint32_t) * nb, cudaMemcpyHostToDevice, h->stream));
  RCG_CUDA(h, cudaMemcpyAsync(dgeom + nb, hi.data(), sizeof(uint32_t) * nb, cudaMemcpyHostToDevice, h->stream));
  RCG_CUDA(h, cudaMemcpyAsync(dgeom + 2 * nb, dfar.data(), sizeof(uint32_t) * nb, cudaMemcpyHostToDevice, h->stream));
  RCG_CUDA(h, cudaMemcpyAsync(dgeom + 3 * nb, c0.data(), sizeof(uint32_t) * (nb + 1), cudaMemcpyHostToDevice, h->stream));
  ClGeom g;
  g.lo = dgeom; g.hi = dgeom + nb; g.dfar = dgeom + 2 * nb; g.c0 = dgeom + 3 * nb; g.nb = nb; g.ring = C.ring;
  int *derr = nullptr;
  RCG_CUDA(h, cudaMalloc(&derr, sizeof(int)));
  RCG_CUDA(h, cudaMemsetAsync(derr, 0, sizeof(int), h->stream));
  RCG_CUDA(h, cudaMalloc(&C.offN, sizeof(int64_t) * ((size_t)C.nchunks + 1)));
  RCG_CUDA(h, cudaMemsetAsync(C.offN, 0, sizeof(int64_t) * ((size_t)C.nchunks + 1), h->stream));
  const int grid = (int)std::min<int64_t>((int64_t)C.nchunks, (int64_t)h->sm_count * 8);
  k_cl_count<<<grid, 128, 0, h->stream>>>(comb.rowptr, comb.col, g, C.nchunks, C.offN, derr);
  h->stats.kernel_launches += 1;
  int herr = 0;
  RCG_CUDA(h, cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  RCG_CUDA(h, cudaFree(derr));
  if (herr) {   // a row with more near entries than a lane-row can count: this factor stays on the 32-row chain
    cudaFree(dgeom); cudaFree(C.offN);
    C = ClusterDev();
    C.level_on.assign(d.groups.size(), 0);
    C.level_cap.assign(d.groups.size(), 0);
    return RCG_OK;
  }
  // largest near blob per level (before the scan turns sizes into offsets)
  std::vector<int64_t> hN((size_t)C.nchunks + 1);
  RCG_CUDA(h, cudaMemcpy(hN.data(), C.offN, sizeof(int64_t) * hN.size(), cudaMemcpyDeviceToHost));
  for (int b = 0; b < nb; b++)
    for (uint32_t c = c0[b]; c < c0[b + 1]; c++)
      C.level_cap[level_of[b]] = std::max<uint32_t>(C.level_cap[level_of[b]], (uint32_t)hN[c]);
  RCG_TRY(rcg_exclusive_scan(h, C.offN, (int64_t)C.nchunks + 1));
  RCG_CUDA(h, cudaMemcpy(&C.bytesN, C.offN + C.nchunks, sizeof(int64_t), cudaMemcpyDeviceToHost));
  RCG_CUDA(h, cudaMalloc(&C.blobN, (size_t)C.bytesN + 256));
  RCG_CUDA(h, cudaMemsetAsync(C.blobN, 0, (size_t)C.bytesN + 256, h->stream));
  RCG_CUDA(h, cudaMalloc(&C.wslab, (size_t)C.nchunks * CL_WCHUNK));
  RCG_CUDA(h, cudaFuncSetAttribute(k_cl_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, CL_FILL_SMEM));
  k_cl_fill<<<std::min<int>((int)C.nchunks, h->sm_count * 4), 128, CL_FILL_SMEM, h->stream>>>(comb.rowptr, comb.col, comb.val, g, C.nchunks,
                                                                                         C.offN, C.wslab, C.blobN);
  h->stats.kernel_launches += 1;
  RCG_CUDA(h, cudaGetLastError());
  RCG_CUDA(h, cudaMalloc(&C.c0, sizeof(uint32_t) * std::max<size_t>(1, c0_of_block.size())));
  RCG_CUDA(h, cudaMalloc(&C.prog4, sizeof(uint32_t) * 4 * std::max<size_t>(1, c0_of_block.size())));
  RCG_CUDA(h, cudaMemcpyAsync(C.c0, c0_of_block.data(), sizeof(uint32_t) * c0_of_block.size(), cudaMemcpyHostToDevice, h->stream));
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(dgeom);
  // staging plan: at least two slots per CTA
  for (size_t gi = 0; gi < d.groups.size(); gi++) {
    if (!C.level_on[gi]) continue;
    C.level_cap[gi] = (C.level_cap[gi] + 127u) & ~127u;
    if (cl_smem_bytes(C.ring, 2, C.level_cap[gi]) > (size_t)BC_SMEM_MAX) C.level_on[gi] = 0;
  }
  C.on = true;
  return RCG_OK;
}

static int cl_launch(rcg_handle *h, BlockedDev &B, BcArgs a, const GroupHost &G, size_t gi) {
  ClusterDev &C = B.cl;
  if (!(h->smem_optin_mask & (1u << 9))) {   // per handle = per device
    RCG_CUDA(h, cudaFuncSetAttribute(k_cl_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, BC_SMEM_MAX));
    h->smem_optin_mask |= 1u << 9;
  }
  ClArgs A;
  memset(&A, 0, sizeof(A));
  A.wslab = C.wslab; A.blobN = C.blobN; A.offN = C.offN; A.c0 = C.c0; A.prog4 = C.prog4;
  A.ring = C.ring;
  A.capN = C.level_cap[gi];
  const size_t smem = (size_t)BC_SMEM_MAX;   // every rank fills it with as many staging slots as fit (at least two)
  A.smem_total = (uint32_t)smem;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(CL_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL_S; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (C.max_clusters == 0) {
    int nc = 0;
    cfg.gridDim = dim3(CL_S * 32);
    if (cudaOccupancyMaxActiveClusters(&nc, k_cl_solve, &cfg) != cudaSuccess || nc <= 0) { cudaGetLastError(); nc = 24; }
    C.max_clusters = std::min(nc, h->sm_count / (int)CL_S);
  }
  const uint32_t cmax = (uint32_t)std::max(2, C.max_clusters);
  const uint32_t ngroups = std::max(1u, std::min<uint32_t>((uint32_t)G.count, cmax / 2u));
  const uint32_t far_clusters = cmax - ngroups;
  a.ngroups = ngroups;
  a.helpers = std::max(1u, CL_S * far_clusters / ngroups);
  A.f = a;
  cfg.gridDim = dim3(CL_S * (ngroups + far_clusters));
  RCG_CUDA(h, cudaLaunchKernelEx(&cfg, k_cl_solve, A));
  h->stats.kernel_launches += 1;
  return RCG_OK;
}
