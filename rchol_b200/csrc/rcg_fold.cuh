// rchol_b200 -- folded chain (chain_mode 5): the blocked-inverse triangular solve of rcg_blocked.cu with the recent entries
// folded into dense panels at set-up, so that the chain's hop is ONE panel apply by ONE warp:
//
//     x_k = u_k - M_k x_rec          M_k = Winv_k L_rec   (32 x ncol, columns = the distinct recent columns, set-up)
//     u_k = Winv_k t'_k              t'_k = start_k - early - late entries                     (near helper, off the chain)
//
// The old chain did  t = t' - L_rec x_rec (sparse gather, shuffle reduction)  THEN  x = Winv t (dense mat-vec, four warps,
// one named barrier): two dependent stages, ~790 cycles per 32 rows.  Here the only dependent work per hop is: 16 shared-
// memory loads of x (addresses and panel values already in registers), 4 dependent DFMA, 2 DADD, one store.
// Included by rcg_blocked.cu inside its anonymous namespace (shares Guard/BC_WAIT, BcArgs, the memory-model primitives,
// the far tiles / start vector / progress words).
//
// CTA = 12 warps (384 threads, 168 registers each):
//   warps 0,1,2   chain warps, one per scheduler 0..2.  Chunk k belongs to chain warp k % 3.  A lone warp needs ~700 cycles
//                 for the ~110 instructions of a chunk (measured), but only the x loads, 4 dependent DFMA, 2 DADD and two
//                 stores depend on the previous chunk: with three warps taking turns, the bookkeeping of chunk k (staging
//                 barrier, next tail into registers, Winv t') overlaps the critical sections of k+1, k+2.  They sit on
//                 DIFFERENT schedulers: a warp that spins for its turn on the scheduler of the warp in its critical
//                 section takes every other issue slot from it (measured: 1095 cycles per hop with all three on scheduler 0).
//   warp 3        publisher: window -> out[], fused dot product, progress published with release semantics
//   warp 4 / 8    TMA producers of ring A (panels + Winv) / ring B (early + late entries); they sleep in mbarrier waits
//   warps 5,6,7,9,10,11   near helpers: chunk k -> helper k % 6, running up to 6 chunks ahead
constexpr int FC_THREADS = 384;
constexpr uint32_t FC_NC = 3;
constexpr uint32_t FC_NH = 6;
constexpr uint32_t FC_SCR = FC_NH * 32u;   // scratch doubles: one t vector per helper (broadcast for the Winv mat-vec)

__device__ __forceinline__ void mbar_arrive_after3(uint64_t *bar, uint32_t a, double b, double c) {
  asm volatile(
      "{\n\t.reg .b32 lo, hi, lo2, hi2, z;\n\t"
      "mov.b64 {lo, hi}, %2;\n\t"
      "mov.b64 {lo2, hi2}, %3;\n\t"
      "or.b32 z, lo, %1;\n\t"
      "or.b32 z, z, lo2;\n\t"
      "and.b32 z, z, 0;\n\t"
      "add.u32 z, z, %0;\n\t"
      "mbarrier.arrive.shared::cta.b64 _, [z];\n\t}"
      ::"r"(smem_u32(bar)), "r"(a), "d"(b), "d"(c)
      : "memory");
}

__device__ __forceinline__ uint32_t mbar_test_s(uint32_t bar_s, uint32_t par) {   // non-blocking, 32-bit shared address
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar_s), "r"(par) : "memory");
  return ok;
}
__device__ __forceinline__ uint32_t mbar_try_s(uint32_t bar_s, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar_s), "r"(par) : "memory");
  return ok;
}

// PROF: cycle counters of the chain warp and of near helper 0 (dbg bit 0) -- a separate instantiation.
template <bool PROF>
__global__ void __launch_bounds__(FC_THREADS, 1) k_fc_solve(const BcArgs P) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t per = 1u + P.helpers;
  const uint32_t grp = blockIdx.x / per, role = blockIdx.x % per;

  double *win = reinterpret_cast<double *>(smem);
  double *uring = win + P.W + 16;   // win[W] holds 0.0: the slot padding columns / entries point at
  double *scratch = uring + BC_TR * 32u;
  unsigned char *ringA = reinterpret_cast<unsigned char *>(scratch + FC_SCR);
  unsigned char *ringB = ringA + (size_t)P.SA * P.capA;
  uint64_t *fullA = reinterpret_cast<uint64_t *>(ringB + (size_t)P.SB * P.capB);
  uint64_t *emptyA = fullA + P.SA;
  uint64_t *fullB = emptyA + P.SA;
  uint64_t *emptyB = fullB + P.SB;
  unsigned long long *ptrB = reinterpret_cast<unsigned long long *>(emptyB + P.SB);
  uint32_t *seqB = reinterpret_cast<uint32_t *>(ptrB + P.SB);
  uint32_t *tready = seqB + P.SB;
  uint32_t *ctl = tready + BC_TR;   // [0] prog: solved chunks of the current block, [1] published chunks, [2] abort

  if (threadIdx.x < P.SA) mbar_init(fullA + threadIdx.x, 1);   // (ring A has no empty barriers: the producer polls prog)
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
    // (same protocol as k_bc_solve: pass 0 = entries of other, already solved blocks, never waits; pass 1 = entries
    //  >= Dfar chunks back in the own block, tile by tile as the chain's published progress allows)
    const uint32_t hid = role - 1u;
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
            if (threadIdx.x == 0) BC_WAIT(ld_acquire_gpu(P.gprog + b.gidx) >= need, 0x100u, 200);
            __syncthreads();
          }
          const uint32_t lpr = lpr_pass[pass], rpw = 32u / lpr, sub = lane & (lpr - 1u);
          const uint32_t r0 = b.lo + t * (32u * P.tile), r1 = min(b.hi, r0 + 32u * P.tile);
          for (uint32_t base = r0 + warp * rpw; base < r1; base += (FC_THREADS / 32) * rpw) {
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
  const uint32_t win_s = smem_u32(win), u_s = smem_u32(uring), sc_s = smem_u32(scratch), trdy_s = smem_u32(tready);
  uint32_t ia0 = 0;   // chunks of the blocks this CTA has finished: running index of the staging rings
  long long pc[4] = {0, 0, 0, 0};   // chain warp of CTA 0 (dbg bit 0): cycles waiting for ring A / for u, chunks, failed polls of u
  long long ph[5] = {0, 0, 0, 0, 0};   // near helper 0 of CTA 0: start vector, blob B, early, wait for the chain, late + mat-vec
  long long clk0 = 0;
  if (P.clk && blockIdx.x == 0 && threadIdx.x == 0) clk0 = clock64();

  for (uint32_t bi = grp; bi < P.nblocks; bi += P.ngroups) {
    const BcBlock b = P.blocks[bi];
    const uint32_t nch = (b.hi - b.lo + 31u) >> 5;
    if (threadIdx.x < BC_TR) tready[threadIdx.x] = 0u;
    if (threadIdx.x == 0) { ctl[0] = 0u; ctl[1] = 0u; }
    __syncthreads();

    if (warp < FC_NC) {
      // ------------------------------ chain warps (0, 1, 2) ----------------------------------------------
      const uint32_t cw = warp;   // this warp's turn: chunks cw, cw + 3, ...
      // A lone warp is bound by the number of instructions between two hops (measured: 4-8 cycles per dependent
      // instruction, 150 for a synchronous mbarrier test, 160-410 for a divergent branch, a MEMBAR for st.release), so:
      // 32-bit shared addresses only, no divergent branch (lane 0's stores are predicated), the progress word is a
      // volatile store behind the window store (same warp, same LSU queue: program order), the staging barrier of
      // chunk k+2 is tested at the top of chunk k and its answer is consumed a chunk later, and two register sets
      // (E / O) hold the tail of the next chunk (fixed offsets in the blob) while this one's x loads are in flight.
      // The loop body is ONE copy of straight-line code of ~90 instructions (1.5 KB): a lone warp has nobody to hide
      // its instruction fetches behind, and a body larger than the scheduler's L0 instruction cache (two unrolled
      // copies were 6.7 KB) costs more than everything else in it.
      const bool prof = PROF && (P.dbg & 1u) != 0u && blockIdx.x == 0 && warp == 0u;
      uint32_t ringA_s = smem_u32(ringA), ringA_e = ringA_s + P.SA * P.capA, capA = P.capA;
      uint32_t fullA_s = smem_u32(fullA), fullA_e = fullA_s + 8u * P.SA;
      uint32_t wm8 = 8u * wmask, lane16 = 16u * lane;
      uint32_t xst_s = win_s + 8u * lane;            // window slot of this lane's row: + ((256 k) & wm8)
      uint32_t ul_s = u_s + 8u * lane;
      // opaque to the compiler: kept in registers instead of being re-derived from the constant bank / special
      // registers inside the loop (the lone warp pays for every instruction)
      asm volatile("" : "+r"(ringA_s), "+r"(ringA_e), "+r"(capA), "+r"(fullA_s), "+r"(fullA_e), "+r"(wm8));
      asm volatile("" : "+r"(xst_s), "+r"(ul_s), "+r"(lane16));
      const uint32_t step8 = 8u * FC_NC, stepA = FC_NC * capA;
      const bool look2 = P.SA >= 2u * FC_NC + 1u;
      const uint32_t i0 = ia0 + cw, slot0 = i0 % P.SA;     // ring index / staging slot of this warp's first chunk
      const uint32_t par0 = (i0 / P.SA) & 1u;
      uint32_t par1 = par0;                                // parity of the staging barrier of this warp's next chunk (k + 3) ...
      uint32_t bar1 = fullA_s + 8u * slot0 + step8;        // ... its address ...
      uint32_t as_n = ringA_s + slot0 * capA + stepA;      // ... and its staging slot
      if (bar1 >= fullA_e) { bar1 -= 8u * P.SA; as_n -= P.SA * capA; par1 ^= 1u; }
      uint32_t ok1 = 0u;                                   // "chunk k+3 is staged", asked one own chunk ahead
      uint32_t o[16];                                      // tail columns of chunk k: window byte offsets of x
      double mt[16];                                       // tail panel values
      uint32_t ncb = FC_MINB, as_c = ringA_s + slot0 * capA;
      uint32_t tpf = 0u, spins = 0u;
      double un = 0.0;                                     // u of this warp's next chunk, loaded behind its flag
#define FC_SPIN(cond_, code_)                                                                                             \
      while (__builtin_expect(!(cond_), 0)) {                                                                             \
        if (++spins > (1u << 24)) { atomicCAS(P.abort_g, 0u, (code_)); sts_volatile_u32(G.abort_s, 1u); break; }          \
      }
#define FC_LOAD_OFFS(AS_)                                                                                                 \
      _Pragma("unroll") for (uint32_t q_ = 0; q_ < FC_MINB; q_++)                                                         \
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"                                                           \
                     : "=r"(o[4 * q_]), "=r"(o[4 * q_ + 1]), "=r"(o[4 * q_ + 2]), "=r"(o[4 * q_ + 3]) : "r"((AS_) + 16u + 16u * q_) : "memory")
#define FC_LOAD_VALS(AS_)                                                                                                 \
      _Pragma("unroll") for (uint32_t q_ = 0; q_ < FC_MINB; q_++) {                                                       \
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(mt[4 * q_]), "=d"(mt[4 * q_ + 1]) : "r"((AS_) + 80u + lane16 + 1024u * q_) : "memory"); \
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(mt[4 * q_ + 2]), "=d"(mt[4 * q_ + 3]) : "r"((AS_) + 592u + lane16 + 1024u * q_) : "memory"); \
      }
      if (cw < nch) {
        FC_SPIN(mbar_try_s(fullA_s + 8u * slot0, par0) != 0u, 0x200u);
        FC_LOAD_OFFS(as_c);
        FC_LOAD_VALS(as_c);
        ncb = lds_u32(as_c);
        tpf = lds_volatile_u32(trdy_s + 4u * cw);
        un = lds_f64(ul_s + (cw << 8));
        if (cw + FC_NC < nch && P.SA >= FC_NC + 1u + cw) ok1 = mbar_test_s(bar1, par1);
      }
#pragma unroll 1
      for (uint32_t k = cw; k < nch; k += FC_NC) {
        long long c0 = 0;
        if (prof) c0 = clock64();
        // staging barrier of this warp's chunk after next (k + 6): asked now, answered (ok2) while this chunk computes
        uint32_t bar2 = bar1 + step8, par2 = par1, as_2 = as_n + stepA;
        if (bar2 >= fullA_e) { bar2 -= 8u * P.SA; as_2 -= P.SA * capA; par2 ^= 1u; }
        // (only when the slot's previous tenant, chunk k + 6 - SA, is a finished chunk: with fewer than 7 slots the parity
        //  of an older phase would answer for it)
        const uint32_t ok2 = look2 ? mbar_test_s(bar2, par2) : 0u;
        if (__builtin_expect(tpf != k + 1u, 0)) {
          FC_SPIN((tpf = ld_acquire_cta_s(trdy_s + 4u * (k & (BC_TR - 1u)))) == k + 1u, 0x300u);
          un = lds_f64(ul_s + ((k & (BC_TR - 1u)) << 8));
          if (prof) { pc[3] += 1; pc[0] += clock64() - c0; }
        }
        double a0 = un, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        // this warp's turn: the chain has to have solved chunk k-1 (everything above was bookkeeping ahead of it)
        if (k > 0u) {
          long long c2 = 0;
          if (prof) c2 = clock64();
          uint32_t pnow;
          while ((pnow = lds_volatile_u32(prog_s)) < k) {   // two or more hops away: sleep; the last hop: spin
            if (k - pnow > 1u) __nanosleep(64);
            if (++spins > (1u << 24)) { atomicCAS(P.abort_g, 0u, 0xD00u); sts_volatile_u32(G.abort_s, 1u); break; }
          }
          if (prof) pc[1] += clock64() - c2;
        }
        // body: the older batches of a wide panel, two per trip (eight independent gathers in flight)
        if (__builtin_expect(ncb != FC_MINB, 0)) {
          const uint32_t nb = ncb - FC_MINB;
          const uint32_t ob = as_c + FC_TAILB, vb = ob + 16u * nb + lane16;
#pragma unroll 1
          for (uint32_t bb = 0; bb < nb; bb += 2u) {
            uint32_t q[8];
            double m[8], y[8];
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]) : "r"(ob + 16u * bb) : "memory");
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]) : "r"(ob + 16u * bb + 16u) : "memory");
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(m[0]), "=d"(m[1]) : "r"(vb + 1024u * bb) : "memory");
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(m[2]), "=d"(m[3]) : "r"(vb + 1024u * bb + 512u) : "memory");
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(m[4]), "=d"(m[5]) : "r"(vb + 1024u * bb + 1024u) : "memory");
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(m[6]), "=d"(m[7]) : "r"(vb + 1024u * bb + 1536u) : "memory");
#pragma unroll
            for (uint32_t i = 0; i < 8u; i++) y[i] = lds_f64(win_s + q[i]);
            a0 = fma(-m[0], y[0], a0);
            a1 = fma(-m[1], y[1], a1);
            a2 = fma(-m[2], y[2], a2);
            a3 = fma(-m[3], y[3], a3);
            a0 = fma(-m[4], y[4], a0);
            a1 = fma(-m[5], y[5], a1);
            a2 = fma(-m[6], y[6], a2);
            a3 = fma(-m[7], y[7], a3);
          }
        }
        // tail: x of the newest 16 columns (the only loads that depend on the previous hop)
        double x[16];
#pragma unroll
        for (uint32_t i = 0; i < 16u; i++) x[i] = lds_f64(win_s + o[i]);
        // chain-independent loads of the next chunk, in flight behind the x loads: offsets (into the registers the
        // x loads have just read), batch count, flag of u and u itself
        const bool more = k + FC_NC < nch;
        if (more) {
          if (__builtin_expect(ok1 == 0u, 0)) {
            FC_SPIN((ok1 = mbar_try_s(bar1, par1)) != 0u, 0x200u);
          }
          FC_LOAD_OFFS(as_n);
          ncb = lds_u32(as_n);
          // flag of u_{k+1}, then the value: shared-memory loads of one warp complete in order, so a value read behind a
          // set flag is the published one; a clear flag is polled at the top of the next chunk
          tpf = lds_volatile_u32(trdy_s + 4u * ((k + FC_NC) & (BC_TR - 1u)));
          un = lds_f64(ul_s + (((k + FC_NC) & (BC_TR - 1u)) << 8));
        }
#pragma unroll
        for (uint32_t i = 0; i < 16u; i += 4u) {
          a0 = fma(-mt[i], x[i], a0);
          a1 = fma(-mt[i + 1u], x[i + 1u], a1);
          a2 = fma(-mt[i + 2u], x[i + 2u], a2);
          a3 = fma(-mt[i + 3u], x[i + 3u], a3);
        }
        // panel values of the next chunk's tail, into the registers the FMAs above have just read
        if (more) { FC_LOAD_VALS(as_n); }
        const double xk = (a0 + a1) + (a2 + a3);
        sts_f64(xst_s + ((k << 8) & wm8), xk);
        // progress word: a volatile store behind the window store, by ALL lanes (same word, same value: no divergent
        // branch).  It also frees the staging slot: the producer of ring A polls it.
        sts_volatile_u32(prog_s, k + 1u);
        as_c = as_n; as_n = as_2; bar1 = bar2; par1 = par2; ok1 = ok2;
        if (prof) pc[2] += 1;
      }
#undef FC_LOAD_VALS
#undef FC_LOAD_OFFS
#undef FC_SPIN
    } else if (warp == 4u) {
      // ------------------------------ TMA producer, ring A -----------------------------------------------
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
          // the slot's previous tenant is chunk base + l - SA of this block (earlier blocks are finished: block barrier);
          // the chain warp has read it completely when it publishes that chunk's progress
          (void)use;
          if (base + l >= P.SA) BC_WAIT(ld_acquire_cta_s(prog_s) >= base + l - P.SA + 1u, 0x900u, 100);
          if (lane == 0) {
            const uint32_t bytes = (uint32_t)(e1 - e0);
            mbar_expect_tx(fullA + slot, bytes);
            bulk_g2s(ringA + (size_t)slot * P.capA, P.blobA + e0, bytes, fullA + slot);
          }
          __syncwarp();
        }
      }
    } else if (warp == 8u) {
      // ------------------------------ TMA producer, ring B -----------------------------------------------
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
            if (e1 - e0 <= (int64_t)P.capB) {
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
    } else if (warp == 3u) {
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
    } else {
      // ------------------------------ near helpers: warps 5,6,7,9,10,11 ----------------------------------
      const uint32_t hidx = warp - 5u - (warp > 8u ? 1u : 0u);
      const bool hprof = PROF && (P.dbg & 1u) != 0u && blockIdx.x == 0 && hidx == 0u;
      const uint32_t hs_s = sc_s + 256u * hidx;
      uint32_t tiles_known = 0;
      double t0n = 0.0;
      if (hidx < nch) {
        BC_WAIT(ld_acquire_gpu(P.tileflag + b.tile0 + hidx / P.tile) != 0u, 0x400u, 100);
        tiles_known = hidx / P.tile + 1u;
        const uint32_t j = b.lo + 32u * hidx + lane;
        t0n = j < b.hi ? __ldcg(P.w + j) : 0.0;
      }
      for (uint32_t k = hidx; k < nch; k += FC_NH) {
        const uint32_t i = ia0 + k, slot = i % P.SB;
        long long h0 = 0;
        if (hprof) h0 = clock64();
        const double t0 = t0n;
        const uint32_t kn = k + FC_NH, tilen = kn / P.tile;
        uint32_t fln = 1u;
        if (kn < nch && tilen >= tiles_known) fln = ld_acquire_gpu(P.tileflag + b.tile0 + tilen);
        long long h1 = 0;
        if (hprof) h1 = clock64();
        BC_WAIT(lds_volatile_u32(smem_u32(seqB + slot)) == i && mbar_try(fullB + slot, (i / P.SB) & 1u), 0x500u, 200);
        const unsigned char *bp = reinterpret_cast<const unsigned char *>(ptrB[slot]);
        const uint32_t *hd = reinterpret_cast<const uint32_t *>(bp);
        const uint32_t ne_max = hd[0], ne_tot = hd[1], nl = hd[2], Dk = hd[3];
        const uint32_t perm = bp[16u + lane], rank = bp[48u + lane];
        const unsigned char *cnt = bp + BC_BHDR;
        const double *ev = reinterpret_cast<const double *>(bp + BC_BHDR + r16(ne_max));
        const uint16_t *ec = reinterpret_cast<const uint16_t *>(bp + BC_BHDR + r16(ne_max) + r16(8u * ne_tot));
        const double *lv = reinterpret_cast<const double *>(bp + BC_BHDR + r16(ne_max) + r16(8u * ne_tot) + r16(2u * ne_tot));
        const uint16_t *lc = reinterpret_cast<const uint16_t *>(reinterpret_cast<const unsigned char *>(lv) + 256u * nl);
        const unsigned char *wq = reinterpret_cast<const unsigned char *>(lv) + 320u * nl;   // packed Winv
        // early entries: columns in chunks <= k-E-1 (jagged diagonals, four per trip)
        const uint32_t need1 = k > P.E ? k - P.E : 0u;
        long long h2 = 0;
        if (hprof) h2 = clock64();
        BC_WAIT(ld_acquire_cta_s(prog_s) >= need1, 0x600u, 200);
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
        if (hprof) hj = clock64() + (ts == 1.25e-300 ? 1 : 0);
        double t = __shfl_sync(0xffffffffu, ts, (int)rank);
        // late entries (chunks k-E .. k-Kr-1): values and window slots of the first LB slots, and the packed Winv of the
        // chunk, are loaded BEFORE the wait for the chain; the staging slot is free once they are in registers
        constexpr uint32_t LB = 8;
        uint32_t lcr[LB];
        double lvr[LB];
#pragma unroll
        for (uint32_t u = 0; u < LB; u++) {
          lcr[u] = 0;
          lvr[u] = 0.0;
          if (u < nl) { lcr[u] = lc[u * 32u + lane]; lvr[u] = lv[u * 32u + lane]; }
        }
        double wv[32];
#pragma unroll
        for (uint32_t pp = 0; pp < 16u; pp++) {
          const uint32_t rr = lane >= 2u * pp ? lane - 2u * pp : 0u;   // rows above the diagonal pair: clamped, then zeroed
          const double2 w2 = *reinterpret_cast<const double2 *>(wq + 16u * (pp * (33u - pp) + rr));
          wv[2 * pp] = lane >= 2u * pp ? w2.x : 0.0;
          wv[2 * pp + 1] = lane >= 2u * pp ? w2.y : 0.0;
        }
        const bool slot_done = nl <= LB;
        if (slot_done && lane == 0) mbar_arrive_after3(emptyB + slot, lcr[LB - 1u], lvr[LB - 1u], wv[0] + wv[31]);
        if (kn < nch) {
          if (fln == 0u) BC_WAIT(ld_acquire_gpu(P.tileflag + b.tile0 + tilen) != 0u, 0x400u, 100);
          tiles_known = tilen + 1u;
          const uint32_t jn = b.lo + 32u * kn + lane;
          t0n = jn < b.hi ? __ldcg(P.w + jn) : 0.0;
        }
        uint32_t need2 = k > Dk ? k - Dk : 0u;   // Dk: this chunk's fold depth (blob header)
        if (k + 1u > BC_TR) need2 = max(need2, k + 1u - BC_TR);
        long long h3 = 0;
        if (hprof) h3 = clock64();
        {
          uint32_t pnow = 0;
          G.n = 0;
          while ((pnow = ld_acquire_cta_s(prog_s)) < need2) {
            if (guard_poll(G, 0x700u)) break;
            if (need2 - pnow > 1u) __nanosleep(200);
          }
        }
        if (k >= P.Dfar) BC_WAIT(ld_acquire_cta_s(pub_s) >= k - P.Dfar + 1u, 0x800u, 100);
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
        // u = Winv t : t broadcast through the helper's scratch row, lane = row of the result
        sts_f64(hs_s + 8u * lane, t);
        __syncwarp();
        double u0 = 0.0, u1 = 0.0, u2 = 0.0, u3 = 0.0;
#pragma unroll
        for (uint32_t pp = 0; pp < 16u; pp += 2u) {
          double ta, tb, tc, td;
          asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(ta), "=d"(tb) : "r"(hs_s + 16u * pp) : "memory");
          asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(tc), "=d"(td) : "r"(hs_s + 16u * pp + 16u) : "memory");
          u0 = fma(wv[2 * pp], ta, u0);
          u1 = fma(wv[2 * pp + 1], tb, u1);
          u2 = fma(wv[2 * pp + 2], tc, u2);
          u3 = fma(wv[2 * pp + 3], td, u3);
        }
        const double uk = (u0 + u1) + (u2 + u3);
        const uint32_t tsl = k & (BC_TR - 1u);
        sts_f64(u_s + 8u * (tsl * 32u + lane), uk);
        __syncwarp();
        if (lane == 0) {
          st_release_cta_s(trdy_s + 4u * tsl, k + 1u);
          if (!slot_done) mbar_arrive(emptyB + slot);
        }
        if (hprof) { ph[0] += (h1 - h0) + (h3 - hj); ph[1] += h2 - h1; ph[2] += hj - h2; ph[3] += h4 - h3; ph[4] += clock64() - h4; }
      }
    }
    __syncthreads();
    ia0 += nch;
  }
  if (P.clk && blockIdx.x == 0 && threadIdx.x == 0) {
    P.clk[0] = (unsigned long long)(clock64() - clk0);
    if (P.dbg & 1u) {
      P.clk[3] = (unsigned long long)pc[0];    // cycles waiting for ring A
      P.clk[4] = (unsigned long long)pc[1];    // cycles waiting for u
      P.clk[6] = (unsigned long long)pc[2];    // chunks
      P.clk[15] = (unsigned long long)pc[3];   // failed first polls of u
    }
  }
  if (P.clk && (P.dbg & 1u) && blockIdx.x == 0 && threadIdx.x == 160u)   // near helper 0 = warp 5
    for (int q = 0; q < 5; q++) P.clk[8 + q] = (unsigned long long)ph[q];
}

// ---------------------------------------------------------------------------------------------------------
// Warp-per-block levels.  A tree level with many blocks (hundreds of leaves, the lower separator levels) has more
// independent dependency chains than the GPU has schedulers, so nothing is gained by making one chain fast: ONE WARP owns
// a block and walks its chunks in order,  x_k = Winv_k (start_k - in-window entries),  with the solution window of the
// block in the warp's own shared memory -- no flags, no staging rings, no roles.  Throughput comes from the number of
// resident warps (4 per CTA, several CTAs per SM).
//   k_wb_pre   : start[j] = rhs[j] - (entries of other, already solved blocks), all rows of the level in parallel
//   k_wb_solve : per chunk: start - far entries of the own block (>= window back, read back through L2) - in-window
//                entries (blob B: jagged diagonals + ELL) -> mat-vec with the packed inverse -> window, out[], fused dot
// ---------------------------------------------------------------------------------------------------------
constexpr int WB_WARPS = 4;

// grid (x, blocks): the CTAs of one block share its rows in units of 128 (8 warps x 4 rows x 4 row batches per trip).
// A row is a dependent chain of loads (far row pointer / split -> columns, values -> x), so a thread keeps the chains of
// FOUR rows in flight (measured with one: 1.18 ms for the 1.3 GB of the backward leaf level at 256^3 / T = 4096, 1.1 TB/s;
// ncu launch list, profiles/r02_launches_T4096.csv).  P.far_lpr2 == 0xFFFFFFFF: no block is solved before this level
// (forward leaves, backward root): the start vector is the right-hand side.
__global__ void __launch_bounds__(256) k_wb_pre(const BcArgs P) {
  constexpr int U = 4;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t sub = lane & 7u;   // 8 lanes per row, 4 rows per warp, 32 rows per CTA and batch
  const bool no_far = P.far_lpr2 == 0xFFFFFFFFu && P.corr == nullptr;
  for (uint32_t bi = blockIdx.y; bi < P.nblocks; bi += gridDim.y) {
    const BcBlock b = P.blocks[bi];
    for (uint32_t base = b.lo + blockIdx.x * (32u * U); base < b.hi; base += gridDim.x * (32u * U)) {
      uint32_t j[U];
      int64_t e[U], e1[U];
      double acc[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        j[u] = base + 32u * (uint32_t)u + 4u * warp + (lane >> 3);
        e[u] = e1[u] = 0;
        acc[u] = 0.0;
        if (j[u] < b.hi && !no_far) {
          e[u] = P.far_rp[j[u]];
          e1[u] = e[u] + P.far_split[j[u]];   // [e, e1): entries of other blocks
          e[u] += sub;
        }
      }
      bool more = true;
      while (more) {
        uint32_t c[U];
        double v[U];
#pragma unroll
        for (int u = 0; u < U; u++)
          if (e[u] < e1[u]) { c[u] = P.far_col[e[u]]; v[u] = P.far_val[e[u]]; }
        more = false;
#pragma unroll
        for (int u = 0; u < U; u++)
          if (e[u] < e1[u]) {
            if (c[u] >= P.col_min) acc[u] = fma(v[u], __ldcg(P.out + c[u]), acc[u]);
            e[u] += 8;
            more = more || e[u] < e1[u];
          }
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], 4);
        acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], 2);
        acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], 1);
      }
#pragma unroll
      for (int u = 0; u < U; u++)
        if (j[u] < b.hi && sub == 0u) {
          const uint32_t i = P.reversed ? P.N - 1u - j[u] : j[u];
          double s = P.rhs[i];
          if (P.corr) s -= P.corr[i - P.col_min];
          __stcg(P.w + j[u], s - acc[u]);
        }
    }
  }
}

// Per-warp shared memory: window (W + 16 doubles) | t vector (32 doubles) | nbuf staging buffers of capB bytes | two mbarriers.
__host__ __device__ __forceinline__ uint32_t wb_warp_bytes(uint32_t W, uint32_t capB, uint32_t nbuf) {
  return (W + 48u) * 8u + nbuf * capB + 16u;
}

// P.SB = staging buffers per warp: 2 = the next chunk's blob is copied while this one is processed; 1 = the copy starts when
// this chunk's blob has been read (more warps fit an SM, and their turns hide the copy).
__global__ void __launch_bounds__(WB_WARPS * 32, 1) k_wb_solve(const BcArgs P) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t nbuf = P.SB;
  unsigned char *mine = smem + (size_t)warp * wb_warp_bytes(P.W, P.capB, nbuf);
  double *win = reinterpret_cast<double *>(mine);
  double *scr = win + P.W + 16u;
  unsigned char *buf = reinterpret_cast<unsigned char *>(scr + 32u);
  uint64_t *full = reinterpret_cast<uint64_t *>(buf + (size_t)nbuf * P.capB);
  const uint32_t wmask = P.W - 1u;
  const uint32_t scr_s = smem_u32(scr), full_s = smem_u32(full), buf_s = smem_u32(buf), win_s = smem_u32(win);
  if (lane < 16u) win[P.W + lane] = 0.0;
  if (lane < 2u) mbar_init(full + lane, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  uint32_t phase = 0u;   // bit q: parity of the next phase of staging barrier q
  // Blocks are handed out dynamically, longest first (the level's blocks are sorted by length at set-up; P.gprog[-1] ... the
  // word behind the tile flags ... is the level's ticket counter, zeroed with the flags before every launch): leaves of a
  // METIS partition differ 2x in length, and with a static assignment the longest block of a CTA set the pace.
  for (;;) {
    uint32_t bi = 0;
    if (lane == 0u) bi = atomicAdd(P.ticket, 1u);
    bi = __shfl_sync(0xffffffffu, bi, 0);
    if (bi >= P.nblocks) break;
    const BcBlock b = P.blocks[bi];
    const uint32_t nch = (b.hi - b.lo + 31u) >> 5;
    double dot = 0.0;
    // blob offsets: one coalesced load per 32 chunks (this group and the next one), shuffles afterwards -- no
    // global load of an offset on the per-chunk path
    int64_t og0 = 0, og1 = 0;
    if (nch > 0u) {
      og0 = P.offB[b.chunk0 + min(lane, nch)];
      og1 = P.offB[b.chunk0 + min(32u + lane, nch)];
    }
    uint32_t ogbase = 0;
#define WB_OFF(idx_) ((idx_) - ogbase < 32u ? __shfl_sync(0xffffffffu, og0, (int)((idx_) - ogbase)) \
                                            : __shfl_sync(0xffffffffu, og1, (int)((idx_) - ogbase - 32u)))
    // the first chunks' blobs towards L2 (the staging copy below then hits L2 instead of HBM)
    if (nch > 0u) {
      const int64_t p0 = WB_OFF(0u), p1 = WB_OFF(min(nch, 3u));
      if (lane == 0u) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(P.blobB + p0), "r"((uint32_t)(p1 - p0)) : "memory");
    }
    double wn = 0.0;          // start vector of the next chunk (pre-pass result), loaded one chunk ahead
    int64_t frn = 0, fe1n = 0;   // its far entries of the own block: [frn + fsn, fe1n).  The sum is formed when it is
    uint32_t fsn = 0;            // USED, a chunk later: an add right behind the loads stalls the warp for an HBM round
                                 // trip (measured: 31 % of the kernel's stall samples on that one instruction)
    if (nch > 0u && b.lo + lane < b.hi) {
      wn = __ldcg(P.w + b.lo + lane);
      frn = P.far_rp[b.lo + lane];
      fsn = P.far_split[b.lo + lane];
      fe1n = P.far_rp[b.lo + lane + 1u];
    }
    // blob of chunk kk -> staging buffer kk & 1 (bulk copy, completion on the buffer's barrier); blobs larger than a
    // buffer are read from HBM directly
    int64_t o0 = nch > 0u ? WB_OFF(0u) : 0, o1 = nch > 0u ? WB_OFF(1u) : 0;
    if (nch > 0u && lane == 0u && o1 - o0 <= (int64_t)P.capB) {
      mbar_expect_tx(full, (uint32_t)(o1 - o0));
      bulk_g2s(buf, P.blobB + o0, (uint32_t)(o1 - o0), full);
    }
    for (uint32_t k = 0; k < nch; k++) {
      const uint32_t j = b.lo + 32u * k + lane;
      const bool valid = j < b.hi;
      const uint32_t i = P.reversed ? P.N - 1u - j : j;
      const uint32_t cur = nbuf == 2u ? (k & 1u) : 0u, nxt = nbuf == 2u ? (cur ^ 1u) : 0u;
      // next chunk's blob into the other buffer (every lane has finished reading it: __syncwarp at the end of chunk k-1)
      if ((k & 31u) == 0u && k > 0u) {   // next group of 32 offsets
        ogbase = k;
        og0 = og1;
        og1 = P.offB[b.chunk0 + min(k + 32u + lane, nch)];
      }
      if (k + 1u < nch) {
        o0 = o1;
        o1 = WB_OFF(k + 2u);
        const int64_t q0 = WB_OFF(min(k + 3u, nch)), q1 = WB_OFF(min(k + 4u, nch));
        if (lane == 0u) {
          if (nbuf == 2u && o1 - o0 <= (int64_t)P.capB) {
            mbar_expect_tx(full + nxt, (uint32_t)(o1 - o0));
            bulk_g2s(buf + (size_t)nxt * P.capB, P.blobB + o0, (uint32_t)(o1 - o0), full + nxt);
          }
          if (q1 > q0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(P.blobB + q0), "r"((uint32_t)(q1 - q0)) : "memory");   // chunk k+3 towards L2
        }
      }
      // start vector of the pre-pass minus the far entries of the own block (>= a window back; read back through L2):
      // four independent gathers per trip
      double t0 = wn;
      {
        const int64_t e1 = fe1n;
        int64_t e = frn + (int64_t)fsn;
        const uint32_t jn = j + 32u;
        if (k + 1u < nch && jn < b.hi) {   // next chunk's start vector and far range: in flight during this chunk
          wn = __ldcg(P.w + jn);
          frn = P.far_rp[jn];
          fsn = P.far_split[jn];
          fe1n = P.far_rp[jn + 1u];
        } else {
          wn = 0.0; frn = 0; fsn = 0; fe1n = 0;
        }
        double t1 = 0.0, t2 = 0.0, t3 = 0.0;
        for (; e < e1; e += 4) {
          uint32_t c[4];
          double v[4], x[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const bool have = e + u < e1;
            c[u] = have ? P.far_col[e + u] : 0u;
            v[u] = have ? P.far_val[e + u] : 0.0;
          }
#pragma unroll
          for (int u = 0; u < 4; u++) x[u] = v[u] != 0.0 ? __ldcg(P.out + c[u]) : 0.0;
          t0 = fma(-v[0], x[0], t0);
          t1 = fma(-v[1], x[1], t1);
          t2 = fma(-v[2], x[2], t2);
          t3 = fma(-v[3], x[3], t3);
        }
        t0 = (t0 + t1) + (t2 + t3);
      }
      // the chunk's blob: always staged (the level's plan sizes the buffers for its largest blob); every access below is a
      // shared-space load through a 32-bit address -- generic loads cost an address conversion each and twice the latency
      {
        uint32_t spins = 0;
        while (!mbar_try_s(full_s + 8u * cur, (phase >> cur) & 1u)) {
          if (++spins > (1u << 22)) { atomicCAS(P.abort_g, 0u, 0xC00u); break; }
        }
        phase ^= 1u << cur;
      }
      const uint32_t bp_s = buf_s + cur * P.capB;
      uint32_t ne_max, ne_tot, nl, hd3;
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(ne_max), "=r"(ne_tot), "=r"(nl), "=r"(hd3) : "r"(bp_s) : "memory");
      uint32_t perm, rank;
      asm volatile("ld.shared.u8 %0, [%1];" : "=r"(perm) : "r"(bp_s + 16u + lane) : "memory");
      asm volatile("ld.shared.u8 %0, [%1];" : "=r"(rank) : "r"(bp_s + 48u + lane) : "memory");
      const uint32_t cnt_s = bp_s + BC_BHDR;
      const uint32_t ev_s = cnt_s + r16(ne_max), ec_s = ev_s + r16(8u * ne_tot);
      const uint32_t lv_s = ec_s + r16(2u * ne_tot), lc_s = lv_s + 256u * nl, wq_s = lv_s + 320u * nl;
      // in-window entries, jagged diagonals (rows sorted by length), four per trip; lanes beyond a diagonal's length are
      // predicated off (no branch)
      double ts = __shfl_sync(0xffffffffu, t0, (int)perm), ts1 = 0.0;
      uint32_t base = 0;
      for (uint32_t s = 0; s < ne_max; s += 4u) {
        const uint32_t c4 = lds_u32(cnt_s + s);
        const uint32_t n0 = c4 & 255u, n1 = (c4 >> 8) & 255u, n2 = (c4 >> 16) & 255u, n3 = c4 >> 24;
        const uint32_t b0 = base + lane, b1 = b0 + n0, b2 = b1 + n1, b3 = b2 + n2;
        double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0, x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
        uint32_t e0 = 0, e1 = 0, e2 = 0, e3 = 0;
        lds_u16_if(e0, ec_s + 2u * b0, lane < n0);
        lds_u16_if(e1, ec_s + 2u * b1, lane < n1);
        lds_u16_if(e2, ec_s + 2u * b2, lane < n2);
        lds_u16_if(e3, ec_s + 2u * b3, lane < n3);
        lds_f64_if(v0, ev_s + 8u * b0, lane < n0);
        lds_f64_if(v1, ev_s + 8u * b1, lane < n1);
        lds_f64_if(v2, ev_s + 8u * b2, lane < n2);
        lds_f64_if(v3, ev_s + 8u * b3, lane < n3);
        lds_f64_if(x0, win_s + 8u * e0, lane < n0);
        lds_f64_if(x1, win_s + 8u * e1, lane < n1);
        lds_f64_if(x2, win_s + 8u * e2, lane < n2);
        lds_f64_if(x3, win_s + 8u * e3, lane < n3);
        ts = fma(-v0, x0, ts);
        ts1 = fma(-v1, x1, ts1);
        ts = fma(-v2, x2, ts);
        ts1 = fma(-v3, x3, ts1);
        base += n0 + n1 + n2 + n3;
      }
      ts += ts1;
      double t = __shfl_sync(0xffffffffu, ts, (int)rank);
      if (nl) {   // ELL class (lane = row; padding slots point at the zero slot): eight independent gathers per trip
        double q1 = 0.0, q2 = 0.0, q3 = 0.0;
        for (uint32_t s0 = 0; s0 < nl; s0 += 8u) {
          uint32_t cc[8];
          double vv[8], xx[8];
#pragma unroll
          for (uint32_t u = 0; u < 8u; u++) {
            uint32_t cu = 0;
            vv[u] = 0.0;
            lds_u16_if(cu, lc_s + 64u * (s0 + u) + 2u * lane, s0 + u < nl);
            lds_f64_if(vv[u], lv_s + 256u * (s0 + u) + 8u * lane, s0 + u < nl);
            cc[u] = s0 + u < nl ? cu : P.W;   // slots beyond the chunk's count: the zero slot behind the window
          }
#pragma unroll
          for (uint32_t u = 0; u < 8u; u++) xx[u] = lds_f64(win_s + 8u * cc[u]);
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
      // x = Winv t : t broadcast through the warp's scratch row, lane = row of the result; the packed triangle is read
      // with a clamped row (lanes above the diagonal pair get a zero factor)
      sts_f64(scr_s + 8u * lane, t);
      __syncwarp();
      double u0 = 0.0, u1 = 0.0, u2 = 0.0, u3 = 0.0;
#pragma unroll
      for (uint32_t pp = 0; pp < 16u; pp += 2u) {
        double ta, tb, tc, td, wa, wb, wc, wd;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(ta), "=d"(tb) : "r"(scr_s + 16u * pp) : "memory");
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(tc), "=d"(td) : "r"(scr_s + 16u * pp + 16u) : "memory");
        const uint32_t r0 = lane >= 2u * pp ? lane - 2u * pp : 0u, r1 = lane >= 2u * pp + 2u ? lane - 2u * pp - 2u : 0u;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(wa), "=d"(wb) : "r"(wq_s + 16u * (pp * (33u - pp) + r0)) : "memory");
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(wc), "=d"(wd) : "r"(wq_s + 16u * ((pp + 1u) * (32u - pp) + r1)) : "memory");
        const double f0 = lane >= 2u * pp ? 1.0 : 0.0, f1 = lane >= 2u * pp + 2u ? 1.0 : 0.0;
        u0 = fma(wa * f0, ta, u0);
        u1 = fma(wb * f0, tb, u1);
        u2 = fma(wc * f1, tc, u2);
        u3 = fma(wd * f1, td, u3);
      }
      (void)hd3;
      const double x = (u0 + u1) + (u2 + u3);
      if (nbuf == 1u && k + 1u < nch) {   // one buffer: every lane has read this chunk's blob; copy the next one
        __syncwarp();
        if (lane == 0u && o1 - o0 <= (int64_t)P.capB) {
          mbar_expect_tx(full, (uint32_t)(o1 - o0));
          bulk_g2s(buf, P.blobB + o0, (uint32_t)(o1 - o0), full);
        }
      }
      sts_f64(win_s + 8u * ((32u * k + lane) & wmask), x);
      if (valid) {
        __stcg(P.out + i, x);
        if (P.dotvec && i < P.dot_limit) dot = fma(x, P.dotvec[i], dot);
      }
      __syncwarp();
    }
    dot = warp_sum(dot);
    if (lane == 0 && P.dot_partials) P.dot_partials[b.gidx] = dot;
#undef WB_OFF
  }
}
