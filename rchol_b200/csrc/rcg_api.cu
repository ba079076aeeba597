// rchol_b200 -- C ABI (include/rchol_b200.h) and the PCG driver.
// The iteration restates /root/reference/c++/util/pcg.cpp:57-127 with the scalars resident on the device and the
// BLAS-1 calls fused into three vector kernels; one CUDA graph per iteration, one 8-byte read-back per iteration
// for the loop test `||r|| > tol ||b|| && it < maxit` (pcg.cpp:82).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>

#include "rcg_common.cuh"

namespace {

std::string g_create_error;

double wall_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

int ensure_vectors(rcg_handle *h) {
  if (h->b) return RCG_OK;
  // two doubles of padding: the chain kernel stages 16-byte aligned slices of these vectors with bulk copies
  const size_t bytes = sizeof(double) * (h->N + 2);
  double **vecs[] = {&h->b, &h->x, &h->r, &h->p, &h->q, &h->y, &h->z, &h->io};
  for (double **v : vecs) {
    RCG_CUDA(h, cudaMalloc(v, bytes));
    RCG_CUDA(h, cudaMemsetAsync(*v, 0, bytes, h->stream));
  }
  h->reduce_grid = h->sm_count * 8;
  if (h->dist.on) {
    if (h->dist.n_sub > h->N) { h->err = "n_sub exceeds the local dimension"; return RCG_ERR_INVALID; }
    h->dist.dot_limit = h->dist.rank == 0 ? (uint32_t)h->N : h->dist.n_sub;
    if (!h->dist.sbuf) RCG_CUDA(h, cudaMalloc(&h->dist.sbuf, sizeof(double) * (h->N - h->dist.n_sub + 2)));
  }
  h->partial_cap = h->reduce_grid;
  h->rz_slots = h->haveG ? rcg_post_slots(h, h->bwd, nullptr) : 0;
  RCG_CUDA(h, cudaMalloc(&h->partials, sizeof(double) * (2 * (size_t)h->partial_cap + (size_t)h->rz_slots + 8)));
  RCG_CUDA(h, cudaMemsetAsync(h->partials, 0, sizeof(double) * (2 * (size_t)h->partial_cap + (size_t)h->rz_slots + 8), h->stream));
  RCG_CUDA(h, cudaMalloc(&h->counters, sizeof(unsigned int) * 8));
  RCG_CUDA(h, cudaMemsetAsync(h->counters, 0, sizeof(unsigned int) * 8, h->stream));
  // 16 counters + the per-warp time marks of one hop of k_dp_solve (RCG_DP_TRACE_WORDS, rcg_debug_dp_trace)
  RCG_CUDA(h, cudaMalloc(&h->clk_probe, sizeof(unsigned long long) * (16 + RCG_DP_TRACE_WORDS)));
  RCG_CUDA(h, cudaMemsetAsync(h->clk_probe, 0, sizeof(unsigned long long) * (16 + RCG_DP_TRACE_WORDS), h->stream));
  if (!h->abort_flag) {
    RCG_CUDA(h, cudaMalloc(&h->abort_flag, sizeof(unsigned int) * 4));
    RCG_CUDA(h, cudaMemsetAsync(h->abort_flag, 0, sizeof(unsigned int) * 4, h->stream));
  }
  RCG_CUDA(h, cudaMalloc(&h->scal, sizeof(PcgScalars)));
  RCG_CUDA(h, cudaMemsetAsync(h->scal, 0, sizeof(PcgScalars), h->stream));
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  return RCG_OK;
}

void free_vectors(rcg_handle *h) {
  double **vecs[] = {&h->b, &h->x, &h->r, &h->p, &h->q, &h->y, &h->z, &h->io};
  for (double **v : vecs) { cudaFree(*v); *v = nullptr; }
  cudaFree(h->partials); h->partials = nullptr;
  cudaFree(h->counters); h->counters = nullptr;
  cudaFree(h->scal); h->scal = nullptr;
  cudaFree(h->clk_probe); h->clk_probe = nullptr;
  cudaFree(h->abort_flag); h->abort_flag = nullptr;
  if (h->iter_graph) { cudaGraphExecDestroy(h->iter_graph); h->iter_graph = nullptr; }
  if (h->dist.sbuf) { cudaFree(h->dist.sbuf); h->dist.sbuf = nullptr; }
  h->haveB = false;
}

int require(rcg_handle *h, bool needA, bool needG) {
  if (!h) return RCG_ERR_INVALID;
  if ((needA && !h->haveA) || (needG && !h->haveG)) {
    h->err = needA && !h->haveA ? "rcg_set_matrix has not been called" : "rcg_set_factor has not been called";
    return RCG_ERR_STATE;
  }
  return ensure_vectors(h);
}

// kernels of one PCG iteration, in stream order (pcg.cpp:85-110)
int enqueue_iteration(rcg_handle *h) {
  RCG_TRY(rcg_launch_trisolve(h, h->fwd, h->r, h->y, nullptr));        // y = U^{-T} r          :151
  RCG_TRY(rcg_launch_trisolve(h, h->bwd, h->y, h->z, h->r));           // z = U^{-1} y, r.z     :155, :93
  RCG_TRY(rcg_launch_p_update(h));                                      // p = z + beta p        :89-96
  if (h->dist.on) {   // multi-GPU: the top-separator rows of q are summed over the ranks before the dot products
    RCG_TRY(rcg_launch_spmv(h, h->p, h->q, nullptr, false));
    if (h->N > h->dist.n_sub) RCG_TRY(rcg_allreduce_sum(h, h->q + h->dist.n_sub, h->N - h->dist.n_sub));
    RCG_TRY(rcg_launch_dots_pq_pr(h));
  } else {
    RCG_TRY(rcg_launch_spmv(h, h->p, h->q, h->r, true));                // q = A p, p.q, p.r     :100-102
  }
  RCG_TRY(rcg_launch_xr_update(h));                                     // x, r, r.r, it++       :103-110
  return RCG_OK;
}

int build_graph(rcg_handle *h) {
  if (h->iter_graph || !h->opt.use_graph) return RCG_OK;
  cudaGraph_t graph = nullptr;
  uint64_t before = h->stats.kernel_launches;
  RCG_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
  int rc = enqueue_iteration(h);
  cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
  h->stats.launches_per_iteration = h->stats.kernel_launches - before;
  h->stats.kernel_launches = before;   // capture does not launch
  if (rc != RCG_OK) return rc;
  RCG_CUDA(h, e);
  RCG_CUDA(h, cudaGraphInstantiate(&h->iter_graph, graph, 0));
  RCG_CUDA(h, cudaGraphDestroy(graph));
  return RCG_OK;
}

int run_iteration(rcg_handle *h) {
  if (h->opt.use_graph) {
    RCG_TRY(build_graph(h));
    RCG_CUDA(h, cudaGraphLaunch(h->iter_graph, h->stream));
    h->stats.kernel_launches += h->stats.launches_per_iteration;
    return RCG_OK;
  }
  uint64_t before = h->stats.kernel_launches;
  RCG_TRY(enqueue_iteration(h));
  h->stats.launches_per_iteration = h->stats.kernel_launches - before;
  return RCG_OK;
}

// The solve on resident vectors: b on the device, x left on the device.
int solve_resident(rcg_handle *h, double tol, int maxit, double *relres, int *itr) {
  RCG_TRY(require(h, true, true));
  if (!h->haveB) { h->err = "no right-hand side: call rcg_set_rhs or rcg_pcg"; return RCG_ERR_STATE; }
  if (h->opt.use_graph) RCG_TRY(build_graph(h));
  h->history.clear();
  RCG_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  RCG_TRY(rcg_launch_init_solve(h));
  double two[2];   // {rr, bb}: contiguous in PcgScalars
  RCG_CUDA(h, cudaMemcpyAsync(two, &h->scal->rr, sizeof(double) * 2, cudaMemcpyDeviceToHost, h->stream));
  // scal->rr is not written by init: the loop test of iteration 0 uses r = b
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  const double bb = two[1];
  const double nb = std::sqrt(bb);
  double rr = bb;
  int it = 0;
  uint64_t d2h = 16;
  for (;;) {
    const double nr = std::sqrt(rr);
    h->history.push_back(nb > 0 ? nr / nb : nr);
    if (!(nr > nb * tol && it < maxit)) break;       // pcg.cpp:82
    RCG_TRY(run_iteration(h));
    RCG_CUDA(h, cudaMemcpyAsync(&rr, &h->scal->rr, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    RCG_CUDA(h, cudaStreamSynchronize(h->stream));
    d2h += 8;
    it++;
  }
  RCG_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  double res2 = 0.0;
  RCG_TRY(rcg_launch_residual_norm(h, &res2));       // pcg.cpp:116-118
  d2h += 8;
  float ms = 0.f;
  RCG_CUDA(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->stats.solve_ms = ms;
  h->stats.d2h_bytes = d2h;
  RCG_TRY(rcg_check_abort(h));
  if (relres) *relres = std::sqrt(res2) / nb;
  if (itr) *itr = it;
  return RCG_OK;
}

}  // namespace

extern "C" {

const char *rcg_version(void) { return "rchol_b200 0.1 (sm_100a)"; }

const char *rcg_last_error(const rcg_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int rcg_create_with_options(rcg_handle **out, int device, const rcg_options *opt) {
  if (!out) return RCG_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) +
                     " -- rchol_b200 has no CPU fallback";
    return RCG_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) {
    g_create_error = "device ordinal out of range";
    return RCG_ERR_INVALID;
  }
  cudaDeviceProp prop;
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
    return RCG_ERR_CUDA;
  }
  if (prop.major != 10) {
    g_create_error = "rchol_b200 is built for sm_100a (B200) only; device is sm_" + std::to_string(prop.major) +
                     std::to_string(prop.minor);
    return RCG_ERR_CUDA;
  }
  rcg_handle *h = new (std::nothrow) rcg_handle();
  if (!h) return RCG_ERR_NOMEM;
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  if (opt) h->opt = *opt;
  else { memset(&h->opt, 0, sizeof(h->opt)); h->opt.use_graph = 1; }
  if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaEventCreate(&h->ev0)) != cudaSuccess || (e = cudaEventCreate(&h->ev1)) != cudaSuccess) {
    g_create_error = std::string("stream/event creation: ") + cudaGetErrorString(e);
    delete h;
    return RCG_ERR_CUDA;
  }
  memset(&h->stats, 0, sizeof(h->stats));
  *out = h;
  return RCG_OK;
}

int rcg_create(rcg_handle **out, int device) { return rcg_create_with_options(out, device, nullptr); }

int rcg_destroy(rcg_handle *h) {
  if (!h) return RCG_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  rcg_dist_finalize(h);
  free_vectors(h);
  if (h->haveA) rcg_free_csr(h->A);
  if (h->haveG) { rcg_free_direction(h->fwd); rcg_free_direction(h->bwd); }
  cudaFree(h->perm);
  for (int i = 0; i < 2; i++) {
    rcg_release_stage_buffer(h->stage_buf[i]);
    if (h->stage_ev[i]) cudaEventDestroy(h->stage_ev[i]);
  }
  cudaEventDestroy(h->ev0);
  cudaEventDestroy(h->ev1);
  cudaStreamDestroy(h->stream);
  delete h;
  return RCG_OK;
}

int rcg_set_matrix(rcg_handle *h, uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val) {
  if (!h) return RCG_ERR_INVALID;
  RCG_CUDA(h, cudaSetDevice(h->device));
  if (h->b && N != h->N) free_vectors(h);
  // the captured iteration holds A's device pointers and the SpMV's lane template: a new A invalidates it
  if (h->iter_graph) { cudaGraphExecDestroy(h->iter_graph); h->iter_graph = nullptr; }
  return rcg_setup_matrix(h, N, rowPtr, colIdx, val);
}

int rcg_set_matrix_permuted(rcg_handle *h, uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val,
                            const uint64_t *P) {
  if (!h) return RCG_ERR_INVALID;
  RCG_CUDA(h, cudaSetDevice(h->device));
  if (h->b && N != h->N) free_vectors(h);
  if (h->iter_graph) { cudaGraphExecDestroy(h->iter_graph); h->iter_graph = nullptr; }
  return rcg_setup_matrix_permuted(h, N, rowPtr, colIdx, val, P);
}

int rcg_update_matrix_values(rcg_handle *h, uint64_t nnz, const double *val) {
  if (!h) return RCG_ERR_INVALID;
  RCG_CUDA(h, cudaSetDevice(h->device));
  return rcg_refresh_matrix_values(h, nnz, val);   // same device arrays: the captured iteration stays valid
}

int rcg_set_permutation(rcg_handle *h, uint64_t N, const uint64_t *P) {
  if (!h) return RCG_ERR_INVALID;
  RCG_CUDA(h, cudaSetDevice(h->device));
  return rcg_setup_permutation(h, N, P);
}

int rcg_get_matrix(rcg_handle *h, uint64_t *rowPtr, uint64_t *colIdx, double *val) {
  if (!h) return RCG_ERR_INVALID;
  RCG_CUDA(h, cudaSetDevice(h->device));
  return rcg_download_matrix(h, rowPtr, colIdx, val);
}

// xp[i] = x[P[i]] (inverse = false) or x[P[i]] = xp[i] (inverse = true), host vectors through the device
static int permute_host_vector(rcg_handle *h, const double *in_host, double *out_host, bool inverse) {
  RCG_TRY(require(h, false, false));
  if (!in_host || !out_host) { h->err = "null vector"; return RCG_ERR_INVALID; }
  if (!h->perm) { h->err = "no permutation set (rcg_set_permutation / rcg_set_matrix_permuted)"; return RCG_ERR_STATE; }
  RCG_CUDA(h, cudaSetDevice(h->device));
  const size_t bytes = sizeof(double) * h->N;
  RCG_CUDA(h, cudaMemcpyAsync(h->io, in_host, bytes, cudaMemcpyHostToDevice, h->stream));
  RCG_TRY(rcg_apply_permutation(h, h->io, h->q, inverse));   // q is scratch outside a solve
  RCG_CUDA(h, cudaMemcpyAsync(out_host, h->q, bytes, cudaMemcpyDeviceToHost, h->stream));
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  return RCG_OK;
}

int rcg_permute_vector(rcg_handle *h, const double *x_host, double *xp_host) {
  if (!h) return RCG_ERR_INVALID;
  return permute_host_vector(h, x_host, xp_host, false);
}

int rcg_unpermute_vector(rcg_handle *h, const double *xp_host, double *x_host) {
  if (!h) return RCG_ERR_INVALID;
  return permute_host_vector(h, xp_host, x_host, true);
}

int rcg_set_factor(rcg_handle *h, uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val,
                   const uint64_t *part, uint64_t npart) {
  if (!h) return RCG_ERR_INVALID;
  RCG_CUDA(h, cudaSetDevice(h->device));
  free_vectors(h);   // block count (partials) and the captured graph depend on the factor
  return rcg_setup_factor(h, N, rowPtr, colIdx, val, part, npart);
}

int rcg_set_factor_blocks(rcg_handle *h, uint64_t N, const uint64_t *rowPtr, const uint64_t *colIdx, const double *val,
                          const uint64_t *bounds, const int32_t *depth, uint64_t nblocks) {
  if (!h) return RCG_ERR_INVALID;
  RCG_CUDA(h, cudaSetDevice(h->device));
  free_vectors(h);
  return rcg_setup_factor_blocks(h, N, rowPtr, colIdx, val, bounds, depth, nblocks);
}

int rcg_spmv(rcg_handle *h, const double *x_host, double *y_host) {
  RCG_TRY(require(h, true, false));
  if (!x_host || !y_host) { h->err = "null vector"; return RCG_ERR_INVALID; }
  RCG_CUDA(h, cudaSetDevice(h->device));
  const size_t bytes = sizeof(double) * h->N;
  RCG_CUDA(h, cudaMemcpyAsync(h->io, x_host, bytes, cudaMemcpyHostToDevice, h->stream));
  RCG_TRY(rcg_launch_spmv(h, h->io, h->q, nullptr, false));
  if (h->dist.on && h->N > h->dist.n_sub) RCG_TRY(rcg_allreduce_sum(h, h->q + h->dist.n_sub, h->N - h->dist.n_sub));
  RCG_CUDA(h, cudaMemcpyAsync(y_host, h->q, bytes, cudaMemcpyDeviceToHost, h->stream));
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  return RCG_OK;
}

int rcg_trsv(rcg_handle *h, int which, const double *rhs_host, double *out_host) {
  RCG_TRY(require(h, false, true));
  if (!rhs_host || !out_host) { h->err = "null vector"; return RCG_ERR_INVALID; }
  if (which != RCG_TRSV_FORWARD && which != RCG_TRSV_BACKWARD) { h->err = "bad direction"; return RCG_ERR_INVALID; }
  RCG_CUDA(h, cudaSetDevice(h->device));
  const size_t bytes = sizeof(double) * h->N;
  RCG_CUDA(h, cudaMemcpyAsync(h->io, rhs_host, bytes, cudaMemcpyHostToDevice, h->stream));
  RCG_TRY(rcg_launch_trisolve(h, which == RCG_TRSV_FORWARD ? h->fwd : h->bwd, h->io, h->y, nullptr));
  RCG_CUDA(h, cudaMemcpyAsync(out_host, h->y, bytes, cudaMemcpyDeviceToHost, h->stream));
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  return rcg_check_abort(h);
}

// Diagnostics: runs one triangular solve with per-row tracing; trace_host receives 4 uint32 per row (solve index
// space): {finish cycle, polling-loop trips, start cycle, cta*1024+thread}.  Cycle counters are per SM.
int rcg_debug_trace(rcg_handle *h, int which, const double *rhs_host, double *out_host, uint32_t *trace_host) {
  RCG_TRY(require(h, false, true));
  RCG_CUDA(h, cudaSetDevice(h->device));
  const size_t bytes = sizeof(double) * h->N;
  RCG_CUDA(h, cudaMalloc(&h->trace, sizeof(uint32_t) * 4 * h->N));
  RCG_CUDA(h, cudaMemsetAsync(h->trace, 0, sizeof(uint32_t) * 4 * h->N, h->stream));
  RCG_CUDA(h, cudaMemcpyAsync(h->io, rhs_host, bytes, cudaMemcpyHostToDevice, h->stream));
  int rc = rcg_launch_trisolve(h, which == RCG_TRSV_FORWARD ? h->fwd : h->bwd, h->io, h->y, nullptr);
  if (rc == RCG_OK) {
    cudaMemcpyAsync(out_host, h->y, bytes, cudaMemcpyDeviceToHost, h->stream);
    cudaMemcpyAsync(trace_host, h->trace, sizeof(uint32_t) * 4 * h->N, cudaMemcpyDeviceToHost, h->stream);
  }
  cudaStreamSynchronize(h->stream);
  cudaFree(h->trace);
  h->trace = nullptr;
  return rc;
}

int rcg_precond(rcg_handle *h, const double *r_host, double *z_host) {
  RCG_TRY(require(h, false, true));
  if (!r_host || !z_host) { h->err = "null vector"; return RCG_ERR_INVALID; }
  RCG_CUDA(h, cudaSetDevice(h->device));
  const size_t bytes = sizeof(double) * h->N;
  RCG_CUDA(h, cudaMemcpyAsync(h->io, r_host, bytes, cudaMemcpyHostToDevice, h->stream));
  RCG_TRY(rcg_launch_trisolve(h, h->fwd, h->io, h->y, nullptr));
  RCG_TRY(rcg_launch_trisolve(h, h->bwd, h->y, h->z, nullptr));
  RCG_CUDA(h, cudaMemcpyAsync(z_host, h->z, bytes, cudaMemcpyDeviceToHost, h->stream));
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  return rcg_check_abort(h);
}

int rcg_set_rhs(rcg_handle *h, const double *b_host) {
  RCG_TRY(require(h, false, false));
  if (!b_host) { h->err = "null vector"; return RCG_ERR_INVALID; }
  if (h->N == 0) { h->err = "set the matrix or the factor first"; return RCG_ERR_STATE; }
  RCG_CUDA(h, cudaSetDevice(h->device));
  RCG_CUDA(h, cudaMemcpyAsync(h->b, b_host, sizeof(double) * h->N, cudaMemcpyHostToDevice, h->stream));
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  h->haveB = true;
  return RCG_OK;
}

int rcg_pcg_resident(rcg_handle *h, double tol, int maxit, double *relres, int *itr) {
  if (!h) return RCG_ERR_INVALID;
  RCG_CUDA(h, cudaSetDevice(h->device));
  double t0 = wall_ms();
  int rc = solve_resident(h, tol, maxit, relres, itr);
  h->stats.total_ms = wall_ms() - t0;
  return rc;
}

int rcg_get_solution(rcg_handle *h, double *x_host) {
  RCG_TRY(require(h, false, false));
  if (!x_host) { h->err = "null vector"; return RCG_ERR_INVALID; }
  RCG_CUDA(h, cudaSetDevice(h->device));
  RCG_CUDA(h, cudaMemcpyAsync(x_host, h->x, sizeof(double) * h->N, cudaMemcpyDeviceToHost, h->stream));
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  return RCG_OK;
}

int rcg_pcg(rcg_handle *h, const double *b_host, double tol, int maxit, double *x_host, double *relres, int *itr) {
  if (!h) return RCG_ERR_INVALID;
  if (!b_host || !x_host) { h->err = "null vector"; return RCG_ERR_INVALID; }
  double t0 = wall_ms();
  RCG_TRY(rcg_set_rhs(h, b_host));
  RCG_TRY(solve_resident(h, tol, maxit, relres, itr));
  RCG_TRY(rcg_get_solution(h, x_host));
  h->stats.total_ms = wall_ms() - t0;
  h->stats.d2h_bytes += sizeof(double) * h->N;
  return RCG_OK;
}

// b and x in the ORIGINAL ordering: bperm[i] = b[P[i]] (util.hpp:147-155), solve, y[P[i]] = x[i]
// (python/ex_laplace_parallel.py:31-32), both on the device
int rcg_pcg_original(rcg_handle *h, const double *b_host, double tol, int maxit, double *x_host, double *relres, int *itr) {
  if (!h) return RCG_ERR_INVALID;
  if (!b_host || !x_host) { h->err = "null vector"; return RCG_ERR_INVALID; }
  RCG_TRY(require(h, true, true));
  if (!h->perm) { h->err = "no permutation set (rcg_set_permutation / rcg_set_matrix_permuted)"; return RCG_ERR_STATE; }
  if (h->dist.on) { h->err = "rcg_pcg_original is a single-GPU entry point"; return RCG_ERR_STATE; }
  RCG_CUDA(h, cudaSetDevice(h->device));
  double t0 = wall_ms();
  const size_t bytes = sizeof(double) * h->N;
  RCG_CUDA(h, cudaMemcpyAsync(h->io, b_host, bytes, cudaMemcpyHostToDevice, h->stream));
  RCG_TRY(rcg_apply_permutation(h, h->io, h->b, false));
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  h->haveB = true;
  RCG_TRY(solve_resident(h, tol, maxit, relres, itr));
  RCG_TRY(rcg_apply_permutation(h, h->x, h->io, true));
  RCG_CUDA(h, cudaMemcpyAsync(x_host, h->io, bytes, cudaMemcpyDeviceToHost, h->stream));
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  h->stats.total_ms = wall_ms() - t0;
  h->stats.d2h_bytes += bytes;
  return RCG_OK;
}

int rcg_get_history(rcg_handle *h, double *hist, int capacity, int *count) {
  if (!h) return RCG_ERR_INVALID;
  int n = (int)h->history.size();
  if (count) *count = n;
  if (hist)
    for (int i = 0; i < n && i < capacity; i++) hist[i] = h->history[i];
  return RCG_OK;
}

int rcg_pcg_oneshot(int device, uint64_t N, const uint64_t *ArowPtr, const uint64_t *AcolIdx, const double *Aval,
                    const double *b, double tol, int maxit, const uint64_t *GrowPtr, const uint64_t *GcolIdx,
                    const double *Gval, const uint64_t *part, uint64_t npart, double *x, double *relres, int *itr,
                    rcg_stats *stats_or_null) {
  rcg_handle *h = nullptr;
  double t0 = wall_ms();
  int rc = rcg_create(&h, device);
  if (rc != RCG_OK) return rc;
  const double t1 = wall_ms();
  rc = rcg_set_matrix(h, N, ArowPtr, AcolIdx, Aval);
  const double t2 = wall_ms();
  if (rc == RCG_OK) rc = rcg_set_factor(h, N, GrowPtr, GcolIdx, Gval, part, npart);
  const double t3 = wall_ms();
  if (rc == RCG_OK) rc = rcg_pcg(h, b, tol, maxit, x, relres, itr);
  const double t4 = wall_ms();
  if (rc != RCG_OK) g_create_error = h->err;
  if (stats_or_null) {
    rcg_get_stats(h, stats_or_null);
    stats_or_null->total_ms = wall_ms() - t0;
  }
  rcg_destroy(h);
  if (getenv("RCG_TIMING"))
    fprintf(stderr, "[rcg] one-shot: create %.1f ms, set_matrix %.1f, set_factor %.1f, pcg %.1f, destroy %.1f\n", t1 - t0, t2 - t1, t3 - t2,
            t4 - t3, wall_ms() - t4);
  return rc;
}

int rcg_get_stats(rcg_handle *h, rcg_stats *out) {
  if (!h || !out) return RCG_ERR_INVALID;
  *out = h->stats;
  out->N = h->N;
  out->nnzA = h->haveA ? (uint64_t)h->A.nnz : 0;
  out->nnzG = h->nnzG;
  out->n_blocks = (uint64_t)h->n_blocks;
  out->tree_levels = (uint64_t)h->tree_levels;
  size_t dev = 0;
  if (h->haveA) dev += sizeof(int64_t) * (h->N + 1) + (size_t)h->A.nnz * 12;
  if (h->haveG)
    for (const DirectionDev *d : {&h->fwd, &h->bwd}) {
      const BlockedDev &B = d->bc;
      if (B.on) {   // blocked layouts: chain blobs, far entries, start vector, flags, block table
        dev += (size_t)B.bytesA + (size_t)B.bytesB + 512 + 2 * sizeof(int64_t) * ((size_t)B.nchunks + 1);
        dev += sizeof(int64_t) * (h->N + 1) + (size_t)B.far.nnz * 12 + sizeof(uint32_t) * (h->N + 1);
        dev += sizeof(double) * (h->N + 4) + sizeof(uint32_t) * ((size_t)B.ntiles * 2 + B.nblocks + 4) + sizeof(BcBlock) * B.nblocks;
        if (B.dp.on)   // dense-panel levels: packed inverses, near rows, panel table
          dev += sizeof(double) * (size_t)B.dp.inv_doubles + (size_t)B.dp.near.nnz * 12 + sizeof(int64_t) * ((size_t)B.dp.nrows + 4) +
                 sizeof(DpPanel) * B.dp.npanels;
      } else {      // level-space layout
        dev += sizeof(int64_t) * (2 * h->N + 1) + (size_t)h->nnzG * 12;
      }
    }
  if (h->b) dev += sizeof(double) * h->N * 8;
  out->device_bytes = dev;
  if (h->clk_probe) {
    unsigned long long ck[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpy(ck, h->clk_probe, sizeof(ck), cudaMemcpyDeviceToHost) == cudaSuccess && ck[1] > 0)
      out->reserved[0] = (double)ck[0] / (double)ck[1] * 1000.0;   // SM MHz seen by the last chain kernel
    out->reserved[1] = (double)ck[2];                              // 1 + row of a dependency-wait time-out (0 = none)
    for (int i = 0; i < 4; i++) out->reserved[4 + i] = (double)ck[3 + i];   // diagnostics of the critical warp (cycles)
    h->dbg_nbatch = (double)ck[7];
  }
  return RCG_OK;
}

int rcg_time_phase(rcg_handle *h, int phase, int reps, double *avg_ms) {
  RCG_TRY(require(h, true, true));
  if (reps < 1 || !avg_ms) { h->err = "bad arguments"; return RCG_ERR_INVALID; }
  RCG_CUDA(h, cudaSetDevice(h->device));
  auto once = [&]() -> int {
    switch (phase) {
      case 0: return rcg_launch_spmv(h, h->p, h->q, h->r, true);
      case 1: return rcg_launch_trisolve(h, h->fwd, h->r, h->y, nullptr);
      case 2: return rcg_launch_trisolve(h, h->bwd, h->y, h->z, h->r);
      case 3: RCG_TRY(rcg_launch_p_update(h)); return rcg_launch_xr_update(h);
      default: h->err = "bad phase"; return RCG_ERR_INVALID;
    }
  };
  RCG_TRY(once());   // warm-up
  RCG_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  for (int i = 0; i < reps; i++) RCG_TRY(once());
  RCG_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  RCG_CUDA(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  *avg_ms = ms / reps;
  return RCG_OK;
}

int rcg_get_group_count(rcg_handle *h, int direction, int *count) {
  if (!h || !count) return RCG_ERR_INVALID;
  if (!h->haveG) { h->err = "rcg_set_factor has not been called"; return RCG_ERR_STATE; }
  *count = (int)(direction == RCG_TRSV_FORWARD ? h->fwd : h->bwd).groups.size();
  return RCG_OK;
}

int rcg_get_group_info(rcg_handle *h, int direction, int group, uint64_t *info6) {
  if (!h || !info6) return RCG_ERR_INVALID;
  if (!h->haveG) { h->err = "rcg_set_factor has not been called"; return RCG_ERR_STATE; }
  const DirectionDev &d = direction == RCG_TRSV_FORWARD ? h->fwd : h->bwd;
  if (group < 0 || group >= (int)d.groups.size()) { h->err = "group out of range"; return RCG_ERR_INVALID; }
  const GroupHost &g = d.groups[group];
  info6[0] = (uint64_t)g.count; info6[1] = (uint64_t)g.rows; info6[2] = (uint64_t)g.loc_nnz;
  info6[3] = (uint64_t)g.ext_nnz; info6[4] = g.max_rows; info6[5] = d.bc.on ? (uint64_t)g.blob_bytes : g.max_stage;
  return RCG_OK;
}

int rcg_time_group(rcg_handle *h, int direction, int group, int kernel, int reps, double *avg_ms) {
  RCG_TRY(require(h, false, true));
  if (reps < 1 || !avg_ms || kernel < 0 || kernel > 2) { h->err = "bad arguments"; return RCG_ERR_INVALID; }
  DirectionDev &d = direction == RCG_TRSV_FORWARD ? h->fwd : h->bwd;
  if (group < 0 || group >= (int)d.groups.size()) { h->err = "group out of range"; return RCG_ERR_INVALID; }
  RCG_CUDA(h, cudaSetDevice(h->device));
  *avg_ms = 0.0;
  const double *rhs = direction == RCG_TRSV_FORWARD ? h->r : h->y;
  double *out = direction == RCG_TRSV_FORWARD ? h->y : h->z;
  RCG_TRY(rcg_launch_trisolve(h, d, rhs, out, nullptr, group, kernel));   // warm-up
  RCG_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  for (int i = 0; i < reps; i++) RCG_TRY(rcg_launch_trisolve(h, d, rhs, out, nullptr, group, kernel));
  RCG_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  RCG_CUDA(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  *avg_ms = ms / reps;
  return RCG_OK;
}

// Raw diagnostic counters of the last chain kernel (16 x uint64; meaning depends on the kernel, see rcg_blocked.cu).
int rcg_debug_counters(rcg_handle *h, uint64_t *out16) {
  if (!h || !out16) return RCG_ERR_INVALID;
  memset(out16, 0, sizeof(uint64_t) * 16);
  if (!h->clk_probe) return RCG_OK;
  RCG_CUDA(h, cudaSetDevice(h->device));
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  RCG_CUDA(h, cudaMemcpy(out16, h->clk_probe, sizeof(uint64_t) * 16, cudaMemcpyDeviceToHost));
  return RCG_OK;
}

// Time marks of the last traced k_dp_solve launch (rcg_options.reserved[1] bit 1): [CTA][warp][16] clock64 values.
int rcg_debug_dp_trace(rcg_handle *h, uint64_t *out, uint64_t nwords) {
  if (!h || !out) return RCG_ERR_INVALID;
  if (!h->clk_probe) { memset(out, 0, sizeof(uint64_t) * nwords); return RCG_OK; }
  RCG_CUDA(h, cudaSetDevice(h->device));
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  RCG_CUDA(h, cudaMemcpy(out, h->clk_probe + 16, sizeof(uint64_t) * std::min<uint64_t>(nwords, RCG_DP_TRACE_WORDS), cudaMemcpyDeviceToHost));
  return RCG_OK;
}

// Diagnostics of the blocked solve layout (tests): sizes, then raw copies of the device arrays.
int rcg_debug_blocked_info(rcg_handle *h, int direction, uint64_t *info16) {
  if (!h || !info16) return RCG_ERR_INVALID;
  if (!h->haveG) { h->err = "rcg_set_factor has not been called"; return RCG_ERR_STATE; }
  const BlockedDev &B = (direction == RCG_TRSV_FORWARD ? h->fwd : h->bwd).bc;
  memset(info16, 0, sizeof(uint64_t) * 16);
  info16[0] = B.on ? 1 : 0;
  info16[1] = B.nchunks; info16[2] = B.ntiles; info16[3] = B.nblocks;
  info16[4] = (uint64_t)B.bytesA; info16[5] = (uint64_t)B.bytesB; info16[6] = (uint64_t)B.far.nnz;
  info16[7] = B.Kr; info16[8] = B.E; info16[9] = B.Dfar; info16[10] = h->N;
  info16[11] = B.levels.size();
  info16[12] = B.Dfar_sep;
  info16[13] = (uint64_t)B.tile_sep | ((uint64_t)B.wb_min << 16) | ((uint64_t)B.Dfar_wb << 40);
  info16[14] = B.E_sep;
  {   // levels launched on the cluster chain (chain_mode 4)
    uint64_t n = 0;
    if (B.cl.on) for (int v : B.cl.level_on) n += v ? 1 : 0;
    info16[15] = n | ((uint64_t)(B.fold ? 1 : 0) << 32);   // bit 32: folded layout (chain_mode 5)
  }
  return RCG_OK;
}

int rcg_debug_blocked_copy(rcg_handle *h, int direction, int what, void *dst, uint64_t bytes) {
  if (!h || !dst) return RCG_ERR_INVALID;
  if (!h->haveG) { h->err = "rcg_set_factor has not been called"; return RCG_ERR_STATE; }
  const DirectionDev &d = direction == RCG_TRSV_FORWARD ? h->fwd : h->bwd;
  const BlockedDev &B = d.bc;
  if (!B.on) { h->err = "blocked layout is not active"; return RCG_ERR_STATE; }
  RCG_CUDA(h, cudaSetDevice(h->device));
  const void *src = nullptr;
  uint64_t have = 0;
  switch (what) {
    case 0: src = B.offA; have = sizeof(int64_t) * ((uint64_t)B.nchunks + 1); break;
    case 1: src = B.offB; have = sizeof(int64_t) * ((uint64_t)B.nchunks + 1); break;
    case 2: src = B.blobA; have = (uint64_t)B.bytesA; break;
    case 3: src = B.blobB; have = (uint64_t)B.bytesB; break;
    case 4: src = B.far.rowptr; have = sizeof(int64_t) * (h->N + 1); break;
    case 5: src = B.far.col; have = sizeof(uint32_t) * (uint64_t)B.far.nnz; break;
    case 6: src = B.far.val; have = sizeof(double) * (uint64_t)B.far.nnz; break;
    case 7: src = B.tile_need; have = sizeof(uint32_t) * (uint64_t)B.ntiles; break;
    case 8: src = B.blocks; have = sizeof(BcBlock) * (uint64_t)B.nblocks; break;
    case 9: {   // per level: {depth, first, count, capA, capB, SA, SB, groups, helpers, smem}
      std::vector<uint64_t> v;
      for (size_t i = 0; i < B.levels.size(); i++) {
        const GroupHost &G = d.groups[i];
        const BcLevel &L = B.levels[i];
        const uint64_t row[10] = {(uint64_t)G.depth, (uint64_t)G.first, (uint64_t)G.count, L.capA, L.capB, L.SA, L.SB, L.groups, L.helpers, (uint64_t)L.smem};
        v.insert(v.end(), row, row + 10);
      }
      if (bytes < v.size() * 8) { h->err = "buffer too small"; return RCG_ERR_INVALID; }
      memcpy(dst, v.data(), v.size() * 8);
      return RCG_OK;
    }
    default: h->err = "bad selector"; return RCG_ERR_INVALID;
  }
  if (bytes < have) { h->err = "buffer too small"; return RCG_ERR_INVALID; }
  if (have) RCG_CUDA(h, cudaMemcpy(dst, src, have, cudaMemcpyDeviceToHost));
  return RCG_OK;
}

int rcg_profile_iteration(rcg_handle *h, int reps) {
  double a = 0, f = 0, b = 0, v = 0;
  RCG_TRY(rcg_time_phase(h, 0, reps, &a));
  RCG_TRY(rcg_time_phase(h, 1, reps, &f));
  RCG_TRY(rcg_time_phase(h, 2, reps, &b));
  RCG_TRY(rcg_time_phase(h, 3, reps, &v));
  h->stats.spmv_ms = a;
  h->stats.trsv_ms = f + b;
  h->stats.blas1_ms = v;
  return RCG_OK;
}

}  // extern "C"
