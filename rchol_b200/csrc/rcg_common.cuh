// rchol_b200 -- internal declarations shared by the CUDA translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <string>
#include <vector>

#include "../../include/rchol_b200.h"

#define RCG_SM_COUNT_FALLBACK 148
#define RCG_DP_TRACE_WORDS (160 * 32 * 16)   // diagnostics: [CTA][warp][mark] of one hop of k_dp_solve

// ---------------------------------------------------------------------------------------------------------
// Device data layout (all arrays live in HBM for the life of the handle)
// ---------------------------------------------------------------------------------------------------------

// General CSR (the system matrix A): 64-bit row pointers, 32-bit column indices, fp64 values.
struct CsrDev {
  int64_t *rowptr = nullptr;   // N+1
  uint32_t *col = nullptr;     // nnz
  double *val = nullptr;       // nnz
  int64_t nnz = 0;
};

// A lower-triangular "solve matrix" in its own solve index space:
//   forward  solve (U^T y = r): L = U^T, solve index = permuted row index
//   backward solve (U z = y)  : J U J (J = index reversal), solve index = N-1-row
// split by the nested-dissection block structure into two CSR matrices with sorted rows:
//   loc : entries whose column lies in the row's own block (col < row) followed by the diagonal slot,
//         which holds 1/diag -- the dependency-chain part, solved by k_tri_chain;
//   ext : entries whose column lies in another (already solved) block -- a pure SpMV, done by k_tri_external.
// col/val arrays are padded by 8 entries so that 16-byte-granular bulk copies never run off the end.
struct LowerTriDev {
  CsrDev loc;
  CsrDev ext;
};

struct BlockDesc {             // rows [lo, hi) of one block (a window-sized segment of a nested-dissection block)
  uint32_t lo, hi;
  uint32_t grp0;               // index of the block's first 32-row group in DirectionDev::grp_mask
  uint32_t pad;
};

// One dependency group of blocks: all blocks of one tree depth.  Blocks of group g depend only on blocks of
// groups < g (forward: leaves first; backward: root separator first).
struct GroupHost {
  int depth = 0;               // tree depth of the group's blocks (0 = root separator)
  int first = 0;               // index of the group's first block in the direction's block array
  int count = 0;
  uint32_t max_rows = 0;
  int64_t ext_nnz = 0;         // external entries of the group's rows (0 => no pre-pass, chain starts from rhs)
  int64_t loc_nnz = 0;
  int64_t rows = 0;
  uint32_t max_stage = 0;      // largest number of local entries in an aligned 32-row staging group
  int64_t blob_bytes = 0;      // blocked solve: bytes of the level's chain blobs (Winv + recent + late + early entries)
};

// ---------------------------------------------------------------------------------------------------------
// Blocked-inverse chain layout (the default triangular solve, rcg_blocked.cu).  Inside every nested-dissection block
// the rows keep their solve order and are cut into CHUNKS of 32 consecutive rows.  For chunk k of a block:
//     x_k = Winv_k ( rhs_k - sum over all entries outside the chunk's diagonal block  L[j,c] x_c )
// where Winv_k is the dense inverse of the chunk's 32x32 lower-triangular diagonal block (computed once at set-up;
// exact-arithmetic equivalent of the row-by-row substitution).  The dependency chain of a block is then one step per
// CHUNK instead of one step per DAG level.  Entries outside the diagonal block are classified by their chunk distance
// d = chunk(row) - chunk(col) inside the block:
//     1 <= d <= Kr     "recent"  applied by the critical warp of the block's chain CTA        (blob A, ELL)
//     Kr < d <= E      "late"    applied by a helper warp of the chain CTA, low latency       (blob B, ELL)
//     E  < d <  Dfar   "early"   applied by the same helper warp ahead of time                (blob B, jagged diagonals)
//     d >= Dfar, and entries of other (already solved) blocks: "far" -- a row-gather SpMV done by OTHER CTAs of the
//                      same launch, which hand the start vector to the chain CTA through HBM/L2 with release/acquire flags.
// ---------------------------------------------------------------------------------------------------------
struct BcBlock {               // one nested-dissection block (device copy; ordered by tree level)
  uint32_t lo, hi;             // rows [lo, hi) in the direction's solve index space
  uint32_t chunk0;             // global index of the block's first chunk
  uint32_t tile0;              // global index of the block's first far tile (8 chunks)
  uint32_t gidx;               // index of the block in the direction: progress counter, dot-partial slot
  uint32_t pad[3];             // pad[0] = window of the block in chunks (Dfar), pad[1] = chunks per far tile, pad[2] = early/late distance E
};

struct BcLevel {               // per tree level: shared-memory plan of the launch
  uint32_t capA = 0, capB = 0; // staging slot sizes in bytes
  uint32_t SA = 0, SB = 0;     // staging slots
  uint32_t groups = 0;         // chain CTAs of the launch
  uint32_t helpers = 0;        // far CTAs per chain CTA
  uint32_t Dfar = 0;           // window of the level's blocks in chunks
  bool wb = false;             // warp-per-block level: k_wb_pre + k_wb_solve instead of the chain kernel
  bool dp = false;             // dense-panel level: k_wb_pre + k_dp_solve (rcg_dense.cuh); see DenseDev
  uint32_t wbocc = 1;          // ... resident CTAs per SM (persistent grid: the warps take blocks from a ticket counter)
  size_t smem = 0;
};

// Cluster chain (chain_mode 4, rcg_cluster.cuh): the leaf blocks' chain advances 128 rows per hop; the dense inverse of a
// 128x128 diagonal block is split row-wise over a thread-block cluster of 4 CTAs and the solved rows are exchanged
// through distributed shared memory.  Far entries (tile flags, start vector, progress words) are shared with the
// 32-row layout above, so the far CTAs are the same.
struct ClusterDev {
  bool on = false;
  uint32_t nchunks = 0;                  // 128-row chunks of all cluster-solved blocks
  unsigned char *wslab = nullptr;        // nchunks x 80 KiB: four row slabs of the dense inverse (lower trapezoids)
  unsigned char *blobN = nullptr;        // near entries (inside the window, outside the diagonal block), jagged diagonals
  int64_t *offN = nullptr;               // nchunks+1 byte offsets into blobN
  uint32_t *c0 = nullptr;                // per block of the direction (gidx): first 128-row chunk, 0xFFFFFFFF = not cluster-solved
  uint32_t *prog4 = nullptr;             // per block (gidx) x cluster rank: published hops of the current solve
  int64_t bytesN = 0;
  std::vector<int> level_on;             // parallel to DirectionDev::groups: 1 = the level is launched as k_cl_solve
  std::vector<uint32_t> level_cap;       // largest near blob of the level (bytes)
  uint32_t ring = 0;                     // rows of the solution window ring (32*Dfar + 128)
  int max_clusters = 0;                  // co-resident clusters of 4 CTAs (cudaOccupancyMaxActiveClusters)
};

// Dense-panel levels (rcg_dense.cuh): the rows of every block of the level are cut into PANELS of C rows (C per level, a
// multiple of 32, up to 1024); the inverse of each panel's C x C lower-triangular diagonal block is computed at set-up and
// stored as a packed lower triangle.  All blocks of the level advance in lock step, one panel per HOP:
//     t_k = start_k - (entries of the own block left of the panel) x        sparse, all SMs       | grid barrier
//     x_k = Inv_k t_k                                                       dense mat-vec, all SMs | grid barrier
// so the dependency chain of a 65 536-row separator is 64 hops of the whole GPU instead of 2048 hops of one CTA.
struct DpPanel {               // one panel (device); the level's panels are ordered hop-major
  uint32_t row0, m;            // rows [row0, row0 + m) in the direction's solve index space, m <= C
  uint32_t slice0;             // set-up: global index of the panel's first 32-column slice (inversion tasks)
  uint32_t slot;               // dot-partial slot of the panel
  int64_t inv_off;             // offset (doubles) of the packed inverse in DenseDev::inv
  uint32_t q0;                 // compact index of the panel's first row in DenseDev::near
  uint32_t pad;
  int64_t e0, e1;              // near entries of the panel's rows: [e0, e1) (contiguous: bulk L2 prefetch a hop ahead)
};

struct DpLevel {               // host: one entry per tree level of the direction (parallel to DirectionDev::groups)
  bool on = false;
  uint32_t C = 0;              // panel rows of the level
  uint32_t nhops = 0;          // panels of the level's longest block
  uint32_t panel0 = 0;         // first panel of the level in DenseDev::panels
  uint32_t hop0 = 0;           // first entry of the level in DenseDev::hop_ptr (nhops + 1 entries, relative to panel0)
  uint32_t npanels = 0;
  int64_t inv_bytes = 0;
};

struct DenseDev {
  bool on = false;
  DpPanel *panels = nullptr;
  uint32_t *hop_ptr = nullptr;
  double *inv = nullptr;
  CsrDev near;                           // own-block entries left of the row's panel; rows = compact index of the dense-panel rows
                                         // (level by level, block by block), columns in vector space, raw values
  std::vector<uint32_t> q0_of_block;     // set-up: compact index of the first row of every block (by gidx), 0 if not dense-panel
  int64_t nrows = 0;                     // rows of all dense-panel levels
  double *t0 = nullptr, *t1 = nullptr;   // per compact row: start vector after the entries of other blocks / after the near entries
  int64_t inv_doubles = 0;
  uint32_t npanels = 0;
  int max_ctas = 0;                      // co-resident CTAs of k_dp_solve (occupancy x SMs)
  std::vector<DpLevel> levels;
};

struct BlockedDev {
  ClusterDev cl;
  DenseDev dp;
  bool on = false;
  bool fold = false;                     // folded layout (chain_mode 5, rcg_fold.cuh): dense panels in blob A, packed Winv in blob B
  uint32_t Kr = 2, E = 16, Dfar = 128;   // chunk-distance thresholds (see above); window = 32*Dfar rows (leaf blocks)
  uint32_t Dfar_sep = 32;                // window of the separator blocks
  uint32_t E_sep = 6;                    // early/late distance of the separator blocks (leaves: E)
  uint32_t tile_sep = 1;                 // chunks per far tile of the separator blocks (leaves: 8)
  uint32_t wb_min = 0;                   // tree levels with at least this many blocks are solved warp-per-block (0: none)
  uint32_t Dfar_wb = 32;                 // window (chunks) of the warp-per-block levels
  uint32_t nchunks = 0, ntiles = 0, nblocks = 0;
  int64_t *offA = nullptr, *offB = nullptr;          // nchunks+1 byte offsets into the blobs
  unsigned char *blobA = nullptr, *blobB = nullptr;
  int64_t bytesA = 0, bytesB = 0;
  CsrDev far;                                        // rows in solve space, columns in vector space, raw values
  uint32_t *far_split = nullptr;                     // per row: leading far entries that belong to other blocks
  uint32_t *tile_need = nullptr;                     // per far tile: chunks of the own block that must be published first
  uint32_t *flags = nullptr;                         // [0, ntiles) tile-ready flags, [ntiles, ntiles+nblocks) block progress
  double *w = nullptr;                               // N+2: start vector of the chain (rhs - far part), solve space
  BcBlock *blocks = nullptr;                         // device, ordered like DirectionDev::groups
  std::vector<BcBlock> blocks_host;
  std::vector<BcLevel> levels;                       // parallel to DirectionDev::groups
};

struct DirectionDev {          // one solve direction
  LowerTriDev M;               // rows in LEVEL SPACE: inside every block sorted by (DAG level, row); loc columns are
                               // block-relative level-space positions, ext columns are vector-space indices
  uint32_t *vecidx = nullptr;  // N: vector-space index of level-space row v (folds the backward solve's reversal)
  double *w = nullptr;         // N+2: level-space work vector (start vector in, solution out)
  uint32_t *grp_mask = nullptr; // per 32-row group of a block: bit l set <=> row l of the group starts a new DAG level
  BlockDesc *blocks = nullptr; // device, ordered by group
  std::vector<BlockDesc> blocks_host;
  std::vector<GroupHost> groups;
  bool reversed = false;       // vector index = N-1-solve index
  BlockedDev bc;               // blocked-inverse chain layout (default); M / vecidx / w / grp_mask are then unused
};

// Scalars of the PCG recurrences, resident on the device (pcg.cpp:82-110 keeps them on the host).
struct PcgScalars {
  double rz;        // r.z of the current iteration
  double rz_prev;   // r.z of the previous iteration  (the reference recomputes prev_r.prev_cond, pcg.cpp:94)
  double pq;        // p.q
  double pr;        // p.r
  double rr;        // r.r after the update (loop test of the next iteration, pcg.cpp:82)
  double bb;        // b.b
  int it;           // completed iterations
  int pad;
};

// Multi-GPU layout (one process per GPU): the local index space of a rank is [rows of its own depth-g subtree |
// rows of the replicated top separators]; NCCL sums the three things that cross subtree boundaries.
struct DistState {
  bool on = false;
  int rank = 0, nranks = 1;
  uint32_t n_sub = 0;          // local rows [0, n_sub) = own subtree, [n_sub, N) = replicated top separators
  int top_depth = 0;           // groups with depth < top_depth are the replicated top separators
  void *comm = nullptr;        // ncclComm_t
  double *sbuf = nullptr;      // N - n_sub doubles: subtree -> top-separator coupling of the forward solve
  uint32_t dot_limit = 0;      // rows counted in dot products on this rank (top rows count on rank 0 only)
};

struct rcg_handle {
  int device = 0;
  DistState dist;
  int sm_count = RCG_SM_COUNT_FALLBACK;
  cudaStream_t stream = nullptr;
  rcg_options opt{};
  std::string err;

  uint64_t N = 0;
  bool haveA = false, haveG = false, haveB = false;
  CsrDev A;
  int spmv_lanes = 8;
  DirectionDev fwd, bwd;
  uint64_t nnzG = 0;
  int n_blocks = 1, tree_levels = 1;

  // work vectors (N doubles each)
  double *b = nullptr, *x = nullptr, *r = nullptr, *p = nullptr, *q = nullptr, *y = nullptr, *z = nullptr;
  double *io = nullptr;                 // staging vector for the single-kernel entry points
  PcgScalars *scal = nullptr;           // device
  double *partials = nullptr;           // reduction scratch: [0,P) slot A, [P,2P) slot B, then rz partials
  unsigned int *counters = nullptr;     // last-block-done counters
  int partial_cap = 0;                  // P
  int rz_slots = 0;                     // per-CTA partials of the backward solve's fused r.z
  int reduce_grid = 0;                  // CTAs of the grid-stride vector kernels

  uint32_t *trace = nullptr;                 // diagnostics: per-row timing trace of the chain kernel (rcg_debug_trace)
  unsigned long long *clk_probe = nullptr;   // device {cycles, ns} written by CTA 0 of the chain kernel
  unsigned int *abort_flag = nullptr;        // device: non-zero when a dependency wait of the blocked solve timed out
  bool smem_optin_blocked = false;           // large dynamic shared memory enabled for this handle's device (rcg_blocked.cu)
  uint32_t smem_optin_mask = 0;              // same for the opt-in paths: bit MODE = k_tri_chain<MODE>, bit 8 = k_tri_chain_lv, bit 9 = k_cl_solve
  cudaGraphExec_t iter_graph = nullptr;
  std::vector<double> history;

  double dbg_nbatch = 0;
  rcg_stats stats{};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;

  uint32_t *perm = nullptr;   // device copy of the caller's permutation P (rcg_set_permutation / rcg_set_matrix_permuted)
  bool a_resorted = false;    // A was built on the device from the ORIGINAL ordering (rcg_set_matrix_permuted): its entries are
                              // not in the caller's order, so a value-only refresh does not apply

  // two pinned staging buffers of the host->device upload (rcg_setup.cu), allocated at the first upload
  void *stage_buf[2] = {nullptr, nullptr};
  cudaEvent_t stage_ev[2] = {nullptr, nullptr};
};

// ---------------------------------------------------------------------------------------------------------
// device memory
// ---------------------------------------------------------------------------------------------------------
// The set-up allocates and frees tens of GB in a dozen pieces per call (CSR copies of G, its transpose, the blocked
// layouts); cudaMalloc / cudaFree of GB-sized pieces map and unmap physical memory and took 25 ... 3000 ms of a 30 ms
// analysis from one call to the next (measured at 128^3; at 256^3 1400 ms of the second one-shot solve, also with
// cudaMallocAsync and an unlimited release threshold).  All device allocations of the library therefore go through a
// small caching allocator (rcg_pool.cu): freed blocks are kept and handed to the next request of the same size -- the
// reference's one-shot `pcg(...)` constructor creates and destroys a handle per solve and repeats the same sizes.
// RCG_POOL=0 in the environment restores plain cudaMalloc / cudaFree.
cudaError_t rcg_pool_malloc(void **p, size_t bytes);
cudaError_t rcg_pool_free(void *p);
#ifndef RCG_POOL_IMPL
#define cudaMalloc(p, s) rcg_pool_malloc((void **)(p), (s))
#define cudaFree(p) rcg_pool_free((void *)(p))
#endif

// Phase times of the set-up on stderr when RCG_TIMING is set in the environment (development aid; syncs the stream).
struct RcgPhases {
  bool on;
  double t;
  explicit RcgPhases() : on(getenv("RCG_TIMING") != nullptr), t(now()) {}
  static double now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
  }
  void mark(cudaStream_t s, const char *what) {
    if (!on) return;
    cudaStreamSynchronize(s);
    const double n = now();
    fprintf(stderr, "[rcg]   %-44s %8.1f ms\n", what, n - t);
    t = n;
  }
};

// ---------------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------------
#define RCG_CUDA(h, expr)                                                                         \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      (h)->err = std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" +     \
                 std::to_string(__LINE__) + ")";                                                  \
      return RCG_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)

#define RCG_TRY(expr)                \
  do {                               \
    int _rc = (expr);                \
    if (_rc != RCG_OK) return _rc;   \
  } while (0)

// ---------------------------------------------------------------------------------------------------------
// entry points implemented in the other translation units
// ---------------------------------------------------------------------------------------------------------
// rcg_setup.cu
int rcg_setup_matrix(rcg_handle *h, uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val);
int rcg_setup_permutation(rcg_handle *h, uint64_t N, const uint64_t *P);
int rcg_setup_matrix_permuted(rcg_handle *h, uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val,
                              const uint64_t *P);
int rcg_download_matrix(rcg_handle *h, uint64_t *rowPtr, uint64_t *colIdx, double *val);
int rcg_refresh_matrix_values(rcg_handle *h, uint64_t nnz, const double *val);
int rcg_apply_permutation(rcg_handle *h, const double *src, double *dst, bool inverse);   // device vectors
int rcg_setup_factor(rcg_handle *h, uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val,
                     const uint64_t *part, uint64_t npart);
int rcg_setup_factor_blocks(rcg_handle *h, uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val,
                            const uint64_t *bounds, const int32_t *depth, uint64_t nblocks);
void rcg_release_stage_buffer(void *pinned);
void rcg_free_direction(DirectionDev &d);
void rcg_free_csr(CsrDev &a);

int rcg_exclusive_scan(rcg_handle *h, int64_t *data, int64_t n);
// rcg_blocked.cu
bool rcg_use_blocked(const rcg_handle *h);
int rcg_build_blocked(rcg_handle *h, DirectionDev &d, CsrDev &comb, const std::vector<uint32_t> &bounds,
                      const std::vector<int> &depth, int max_depth, bool root_first);
int rcg_launch_blocked(rcg_handle *h, DirectionDev &d, const double *rhs, double *out, const double *dotvec,
                       int only_group, int only_kernel);
void rcg_free_blocked(BlockedDev &b);
int rcg_check_abort(rcg_handle *h);

// rcg_kernels.cu  (all launches go to h->stream and bump h->stats.kernel_launches)
int rcg_launch_spmv(rcg_handle *h, const double *x, double *y, const double *dot_r /*nullable*/, bool with_dots);
// rcg_trisolve.cu
int rcg_launch_trisolve(rcg_handle *h, DirectionDev &d, const double *rhs, double *out, const double *dotvec,
                        int only_group = -1, int only_kernel = -1);
int rcg_post_slots(const rcg_handle *h, const DirectionDev &d, std::vector<int> *first_slot);
int rcg_compute_levels(rcg_handle *h, const CsrDev &loc, const BlockDesc *blocks_dev, const std::vector<GroupHost> &groups,
                       double *w);
int rcg_launch_p_update(rcg_handle *h);
int rcg_launch_xr_update(rcg_handle *h);
int rcg_launch_init_solve(rcg_handle *h);
int rcg_launch_sum_rz(rcg_handle *h);
int rcg_launch_dots_pq_pr(rcg_handle *h);
// rcg_dist.cu
int rcg_allreduce_sum(rcg_handle *h, double *dev_ptr, size_t count);
extern "C" int rcg_dist_finalize(rcg_handle *h);
int rcg_launch_residual_norm(rcg_handle *h, double *out_host_norm2);
