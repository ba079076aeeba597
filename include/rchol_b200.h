/* rchol_b200 -- C ABI of the B200-native PCG solve phase for the rchol preconditioner.
 *
 * This is the drop-in boundary for the reference's solve-phase hot path
 *     pcg(const SparseCSR &A, const std::vector<double> &b, double tol, int maxit,
 *         const SparseCSR &G, std::vector<double> &x, double &relres, int &itr)
 * (/root/reference/c++/util/pcg.hpp:13-16, implemented in c++/util/pcg.cpp:14-159 on top of Intel MKL).
 * Plain pointers and sizes only; every matrix is passed exactly as the reference's SparseCSR holds it
 * (c++/sparse.hpp:10-31): zero-based CSR, `uint64_t` (size_t) rowPtr[N+1] / colIdx[nnz], `double` val[nnz].
 * Host arrays are borrowed for the duration of the call and never retained (the reference deep-copies into
 * MKL handles, pcg.cpp:31-54); all device memory is owned by the handle and released by rcg_destroy.
 *
 * There is NO CPU fallback: every entry point that computes fails with RCG_ERR_CUDA when no sm_100 device
 * is usable.  All functions return 0 (RCG_OK) on success; rcg_last_error() describes the last failure.
 * A handle is not thread-safe; calls are synchronous (like the reference's pcg).
 */
#ifndef RCHOL_B200_H
#define RCHOL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rcg_handle rcg_handle;

enum {
  RCG_OK = 0,
  RCG_ERR_CUDA = 1,        /* CUDA runtime / driver error, or no usable device          */
  RCG_ERR_INVALID = 2,     /* bad argument (NULL pointer, size mismatch, bad partition)  */
  RCG_ERR_STATE = 3,       /* call order: matrix or factor not set yet                   */
  RCG_ERR_STRUCTURE = 4,   /* G is not upper triangular with a positive leading diagonal,
                              or it violates the nested-dissection block structure       */
  RCG_ERR_NOMEM = 5
};

/* Direction selectors of rcg_trsv, named after the two MKL calls they replace. */
enum {
  RCG_TRSV_FORWARD = 0,    /* y = U^{-T} rhs : mkl_sparse_d_trsv(TRANSPOSE, ...)      pcg.cpp:151 */
  RCG_TRSV_BACKWARD = 1    /* z = U^{-1} rhs : mkl_sparse_d_trsv(NON_TRANSPOSE, ...)  pcg.cpp:155 */
};

/* Tunables (all have defaults; see DESIGN.md "SpTRSV"). */
typedef struct rcg_options {
  int chain_threads;       /* threads per CTA of the block-local sync-free triangular solve (0 = default) */
  int chain_window;        /* blocked solve: rows of the solution window (32*Dfar, default 4096); level-space kernels:
                              chunk rows C of the window (segment = 2C rows; 0 = default 2048)                */
  int use_graph;           /* 1 = replay one CUDA graph per PCG iteration (default), 0 = plain launches   */
  int spmv_lanes;          /* lanes per row of the CSR SpMV (0 = choose from the row-length histogram)     */
  int chain_generic;       /* 1 = force the non-pipelined fallback kernel of the triangular solve (testing)  */
  int chain_mode;          /* 0 (default): separator levels with at most 64 blocks as dense-panel levels (k_dp_pre + k_dp_solve,
                              rcg_dense.cuh: explicitly inverted C x C diagonal panels, all blocks of the level in lock step on the
                              whole GPU; reserved[6]), other tree levels with few blocks on the blocked-inverse chain (k_bc_solve:
                              four critical warps, one named barrier per 32-row chunk), tree levels with many blocks (>= 64,
                              reserved[9]) one warp per block
                              (k_wb_pre + k_wb_solve, rcg_fold.cuh); 5 = the same with the few-block levels on the folded chain
                              (k_fc_solve: recent entries folded into dense panels at set-up, three chain warps taking turns);
                              6 = the blocked-inverse chain for every level (round 1), 3 = the same with one critical warp, 4 = the
                              same with the leaf level on the cluster chain (128-row chunks, 4 CTAs per leaf, DSMEM exchange),
                              1 / 2 = level-space sync-free kernels (rcg_trisolve.cu).  Modes other than 0 are kept for A/B
                              measurements.                                                                                     */
  int reserved[10];        /* [0] helper back-off ns, [1] timing-experiment bits, [2] TMA producer warps (default 2),
                              [3] blocked solve: recent chunk distance Kr (default 2), [4] window rows of the separator blocks
                              (default 1024), [5] blocked solve: distance E (chunks) that separates the "early" from the "late" entries:
                              bits 0-7 leaf blocks (default 16), bits 8-15 separator blocks (default 6), [6] bit 0: plain (non-cooperative) launch; dense-panel levels: bit 1 = the leaf
                              level too, bits 2-7 = levels with more than 2^(v-1) blocks are not dense-panel levels (0 = default 64 blocks,
                              >= 32 = no limit), bits 8-15 = rows / 32 of a level's longest block from which the level is a dense-panel
                              level (0 = default 128 rows, 255 = never), bits 16-31 = panel rows C (0 = from the level's shape), [7] staging slot of the helpers' ring in quarters of
                              the mean blob (default 12), [8] staging slots of the chain's ring (0 = automatic), [9] bits 0-7: lanes per row of the
                              far CTAs' in-block pass (8 default, 32); bits 8-15: chunks per far tile of the separator blocks
                              (default 1; leaves use 8) */
} rcg_options;

/* Per-handle measurements, all device-side times from CUDA events on the handle's own stream. */
typedef struct rcg_stats {
  uint64_t N, nnzA, nnzG;
  uint64_t n_blocks;               /* nested-dissection blocks (2T-1), 1 without a partition              */
  uint64_t tree_levels;            /* log2(T)+1                                                            */
  double upload_ms;                /* host->device copies of A, G (wall clock)                             */
  double analysis_ms;              /* device set-up: index narrowing, transpose of G, schedules            */
  double solve_ms;                 /* last solve: the iterations only, CUDA events                         */
  double total_ms;                 /* last rcg_pcg: wall clock of the whole call                           */
  double trsv_ms, spmv_ms, blas1_ms; /* last rcg_profile_iteration(): per-iteration split, CUDA events     */
  uint64_t kernel_launches;        /* kernels launched by this handle since creation                       */
  uint64_t launches_per_iteration; /* kernels in one PCG iteration                                         */
  uint64_t h2d_bytes, d2h_bytes;   /* bytes copied by the last rcg_pcg call (and by set_A/set_G for h2d)   */
  uint64_t device_bytes;           /* device memory currently owned by the handle                          */
  double reserved[8];
} rcg_stats;

/* ---- life cycle -------------------------------------------------------------------------------------- */
int rcg_create(rcg_handle **out, int device);                 /* device = CUDA ordinal                    */
int rcg_create_with_options(rcg_handle **out, int device, const rcg_options *opt);
int rcg_destroy(rcg_handle *h);
const char *rcg_last_error(const rcg_handle *h);              /* h may be NULL: last creation error        */
const char *rcg_version(void);

/* ---- inputs (replaces pcg::create_sparse, pcg.cpp:31-54) ----------------------------------------------- */
/* A: the (permuted) system matrix, general CSR. */
int rcg_set_matrix(rcg_handle *h, uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val);
/* Value-only refresh of A: a new matrix with the SAME sparsity pattern as the one set by rcg_set_matrix (the reference's
 * reuse flow -- python/rchol/rchol.py:25-40 keeps perm / part for "a new matrix with the same sparsity",
 * python/ex_reuse_partition.py).  val = nnz values in the order of the colIdx given to rcg_set_matrix.  The structure on
 * the device and the captured iteration stay; 8 B per entry are uploaded.  RCG_ERR_STATE after rcg_set_matrix_permuted
 * (the device re-sorted the rows), RCG_ERR_INVALID when nnz differs. */
int rcg_update_matrix_values(rcg_handle *h, uint64_t nnz, const double *val);
/* G: CSR of the upper-triangular factor U as returned by rchol(...) (rchol_lap.cpp:146-149): every row sorted,
 * diagonal first and positive.  `part` (may be NULL, npart = 0) = block boundaries of the reference's
 * nested-dissection layout in permuted index space (rchol_parallel.cpp:64-70 `result_idx`, ground vertex
 * dropped): npart = 2T entries, part[0] = 0, part[npart-1] = N, blocks in post-order
 * [left subtree..., right subtree..., separator].  Without it (the reference's stock pcg signature, pcg.hpp:13-16) the
 * blocks are recovered from G itself (rcg_detect_blocks below); a factor without that structure is solved as one block. */
int rcg_set_factor(rcg_handle *h, uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val,
                   const uint64_t *part, uint64_t npart);

/* General form of rcg_set_factor: `nblocks` consecutive blocks with boundaries bounds[0..nblocks] (bounds[0] = 0,
 * bounds[nblocks] = N) and a tree depth per block; a row of a block may only couple to its own block and to blocks of
 * SMALLER depth.  Used by the multi-GPU layout below. */
int rcg_set_factor_blocks(rcg_handle *h, uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val,
                          const uint64_t *bounds, const int32_t *depth, uint64_t nblocks);

/* ---- the permutation steps either side of the path, on the device --------------------------
 * rcg_set_matrix_permuted: A is given in its ORIGINAL ordering together with the permutation P that rchol(A,G,P,threads)
 * returned; the handle then holds B = A(P,P) with every row re-sorted by column, exactly what the reference's
 * reorder(A, P, B) builds on the host (c++/util/util.cpp:16-57: B row i = A row P[i], column c -> inverse(P)[c]), and
 * keeps P.  P must be a permutation of 0..N-1 (checked; RCG_ERR_INVALID otherwise).
 * rcg_set_permutation: only records P (the caller's A is already permuted).
 * rcg_permute_vector:   xp[i] = x[P[i]]   (util.hpp:147-155 reorder(x, P, xp))
 * rcg_unpermute_vector: x[P[i]] = xp[i]   (python/ex_laplace_parallel.py:31-32  y[p] = x)
 * rcg_pcg_original: the solve with b and x in the ORIGINAL ordering (b is permuted and x un-permuted on the device).
 * rcg_get_matrix: downloads the handle's matrix in SparseCSR form (rowPtr N+1, colIdx/val nnz entries; testing). */
int rcg_set_matrix_permuted(rcg_handle *h, uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val,
                            const uint64_t *P);
int rcg_set_permutation(rcg_handle *h, uint64_t N, const uint64_t *P);
int rcg_permute_vector(rcg_handle *h, const double *x_host, double *xp_host);
int rcg_unpermute_vector(rcg_handle *h, const double *xp_host, double *x_host);
int rcg_pcg_original(rcg_handle *h, const double *b_host, double tol, int maxit, double *x_host, double *relres, int *itr);
int rcg_get_matrix(rcg_handle *h, uint64_t *rowPtr, uint64_t *colIdx, double *val);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink ----------------------------------------
 * Rank r of 2^g owns the depth-g subtree r of the reference's nested-dissection tree (rchol_parallel.cpp:62-70) and
 * a replica of the 2^g - 1 separators above it.  The LOCAL index space of a rank is
 *     [ rows of the own subtree (n_sub of them, in their global order) | rows of the top separators (global order) ]
 * and the caller passes the local matrices in that index space: A_local = rows (subtree + top) x columns (subtree +
 * top) of A, with the [top, top] block present on rank 0 only (the top rows of A p are summed over the ranks);
 * G_local = the same rows and columns of U, via rcg_set_factor_blocks.  Vectors are local ([subtree | top], the top
 * part identical on every rank).  rcg_nccl_unique_id produces the 128-byte NCCL id on one rank (the caller
 * broadcasts it); rcg_dist_init must be called before the matrices are set; `top_depth` = g.  Afterwards rcg_spmv,
 * rcg_trsv, rcg_precond and rcg_pcg* work on the distributed problem (all ranks must call them together). */
int rcg_nccl_unique_id(void *out128);
int rcg_dist_init(rcg_handle *h, int nranks, int rank, const void *unique_id128, uint64_t n_sub, int top_depth);
int rcg_dist_finalize(rcg_handle *h);

/* ---- the kernels of the path, exposed one by one so that parity can be tested in isolation ------------- */
int rcg_spmv(rcg_handle *h, const double *x_host, double *y_host);                 /* pcg.cpp:130-138 */
int rcg_trsv(rcg_handle *h, int which, const double *rhs_host, double *out_host);  /* pcg.cpp:151 / :155 */
int rcg_precond(rcg_handle *h, const double *r_host, double *z_host);              /* pcg.cpp:141-159 */

/* ---- the solve (replaces pcg::iteration, pcg.cpp:57-127) ----------------------------------------------- */
/* Zero initial guess; iterate while ||r||_2 > tol * ||b||_2 and it < maxit; relres = ||A x - b|| / ||b|| (true
 * residual); *itr = completed iterations.  b_host and x_host are host arrays of length N. */
int rcg_pcg(rcg_handle *h, const double *b_host, double tol, int maxit, double *x_host, double *relres, int *itr);

/* Same solve with the right-hand side already resident (rcg_set_rhs) and the solution left on the device
 * (rcg_get_solution): the timed region of the "inputs resident in HBM" benchmark. */
int rcg_set_rhs(rcg_handle *h, const double *b_host);
int rcg_pcg_resident(rcg_handle *h, double tol, int maxit, double *relres, int *itr);
int rcg_get_solution(rcg_handle *h, double *x_host);
/* Optional residual history of the last solve: ||r_k|| / ||b|| at every loop test (k = 0..itr). */
int rcg_get_history(rcg_handle *h, double *hist, int capacity, int *count);

/* One-shot form = exactly what the reference constructor does (wrap A, wrap G, iterate, destroy). */
int rcg_pcg_oneshot(int device, uint64_t N, const uint64_t *ArowPtr, const uint64_t *AcolIdx, const double *Aval,
                    const double *b, double tol, int maxit, const uint64_t *GrowPtr, const uint64_t *GcolIdx,
                    const double *Gval, const uint64_t *part, uint64_t npart, double *x, double *relres, int *itr,
                    rcg_stats *stats_or_null);

/* ---- measurement ------------------------------------------------------------------------------------- */
int rcg_get_stats(rcg_handle *h, rcg_stats *out);
/* Runs `reps` PCG iterations' worth of kernels on the resident vectors without the convergence test and fills
 * trsv_ms / spmv_ms / blas1_ms (averages per iteration) -- the live CUDA-event measurement bench.py reports. */
int rcg_profile_iteration(rcg_handle *h, int reps);
/* Times `reps` launches of a single phase with CUDA events: 0 = SpMV, 1 = forward solve, 2 = backward solve,
 * 3 = fused vector updates.  *avg_ms receives the average duration of one phase execution. */
int rcg_time_phase(rcg_handle *h, int phase, int reps, double *avg_ms);

/* Dependency groups of a triangular solve (one per tree level; direction RCG_TRSV_FORWARD / RCG_TRSV_BACKWARD).
 * rcg_get_group_count returns the number of groups through *count.  rcg_get_group_info fills info[0..5] =
 * {blocks, rows, entries handled by the chain CTAs, entries handled as a row-gather SpMV (far CTAs / pre-pass kernel),
 * largest block rows, blocked solve: bytes of the level's chain blobs (level-space kernels: largest staging group)}.
 * rcg_time_group times `reps` launches of ONE kernel of one group with CUDA events: kernel 0 = the level's solve kernel
 * (blocked solve: k_bc_solve, the only kernel of a level), 1 / 2 = pre-pass / scatter kernels of the level-space path;
 * *avg_ms = average duration of one launch (0 if the group has no such kernel). */
int rcg_get_group_count(rcg_handle *h, int direction, int *count);
int rcg_get_group_info(rcg_handle *h, int direction, int group, uint64_t *info6);
int rcg_time_group(rcg_handle *h, int direction, int group, int kernel, int reps, double *avg_ms);

/* ---- diagnostics ------------------------------------------------------------------------------------- */
/* One triangular solve with per-row tracing of the dependency-chain kernel: trace_host receives 4 uint32 per row
 * in solve index order {finish cycle, polling-loop trips, start cycle, cta*1024+thread} (per-SM cycle counters). */
int rcg_debug_trace(rcg_handle *h, int which, const double *rhs_host, double *out_host, uint32_t *trace_host);

/* Layout of the blocked triangular solve (DESIGN.md "SpTRSV"), for the layout tests: info[0..11] = {active, chunks,
 * far tiles, blocks, bytes of blob A, bytes of blob B, far entries, Kr, E, Dfar, N, levels}; rcg_debug_blocked_copy copies
 * one device array to the host: 0 offA, 1 offB, 2 blobA, 3 blobB, 4 far rowptr, 5 far col, 6 far val, 7 tile_need,
 * 8 blocks (8 x uint32 each: lo, hi, chunk0, tile0, gidx, pad), 9 per-level plan (10 x uint64 each). */
int rcg_debug_blocked_info(rcg_handle *h, int direction, uint64_t *info16);
/* Host only: the row-length histogram the SpMV plan is chosen from (BASELINE north star: "chosen per row-length histogram";
 * replaces nothing in the reference, whose mkl_sparse_d_mv at pcg.cpp:137 runs without optimize hints).  entries5[k] = entries in
 * rows of length <= 2, 3..5, 6..12, 13..24, > 24; *lanes = lanes per row (2 ... 32): the smallest whose bucket edge covers
 * 90 % of the entries.  rcg_options.spmv_lanes > 0 overrides the choice. */
int rcg_spmv_row_histogram(uint64_t N, const uint64_t *rowPtr, uint64_t *entries5, int *lanes);

/* Blocks that rcg_set_factor derives from G when no `part` is given (the stock signature of the reference's pcg,
   /root/reference/c++/util/pcg.hpp:13-16, carries none).  Host-only, no device needed: bounds_out gets *nblocks + 1
   boundaries, depth_out *nblocks tree depths; returns RCG_ERR_INVALID when more than `cap` blocks would be written and
   *nblocks = 0 when the factor is solved as one block. */
int rcg_detect_blocks(uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, uint64_t *bounds_out, int32_t *depth_out,
                      uint64_t cap, uint64_t *nblocks);
int rcg_debug_dp_trace(rcg_handle *h, uint64_t *out, uint64_t nwords);   /* per-warp time marks of one hop of the last dense-panel launch (reserved[1] bit 1): [CTA][32 warps][16] */
int rcg_debug_counters(rcg_handle *h, uint64_t *out16);   /* raw cycle counters of the last chain kernel (rcg_options.reserved[1] bit 0) */
int rcg_debug_blocked_copy(rcg_handle *h, int direction, int what, void *dst, uint64_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* RCHOL_B200_H */
