// rchol_b200 -- the kernels of the PCG iteration (sm_100a).  fp64 throughout; nothing here is a dense
// contraction, so no tensor-core path exists (BASELINE.json north_star); the bound is HBM bandwidth for SpMV and
// the vector updates, and the dependency chain of the factor for the triangular solves.
//
//   k_spmv<LPR>        q = A p (+ fused p.q and p.r)            replaces mkl_sparse_d_mv      pcg.cpp:130-138
//   (triangular solves: rcg_trisolve.cu)
//   k_p_update         p = z + (r.z / r_prev.z_prev) p          replaces ddot x2, dscal, daxpy pcg.cpp:89-96
//   k_xr_update        x += a p ; r -= a q ; r.r                replaces ddot x2, daxpy x2, dcopy x2, dnrm2
//                                                                                            pcg.cpp:101-108,82
#include <cstdio>

#include "rcg_device.cuh"

namespace {

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
  // U row batches per trip: the loads of a row are a dependent chain (row pointers -> columns / values -> x), so a warp
  // with one batch in flight moves 4 rows per three memory round trips (measured 2.1 TB/s); U independent chains overlap.
  // The order in which a thread adds up its rows is unchanged (bit-identical p.q / p.r).
  constexpr int U = 4;
  const uint32_t stride = gridDim.x * rows_per_cta;
  for (uint32_t base = blockIdx.x * rows_per_cta; base < N; base += U * stride) {
    uint32_t row[U];
    int64_t k[U], e[U];
    double s[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint64_t r64 = (uint64_t)base + (uint64_t)u * stride + threadIdx.x / LPR;
      const bool in = r64 < N;
      row[u] = in ? (uint32_t)r64 : 0xFFFFFFFFu;
      k[u] = e[u] = 0;
      s[u] = 0.0;
      if (in) {
        k[u] = rowptr[row[u]] + sub;
        e[u] = rowptr[row[u] + 1];
      }
    }
    bool more = true;
    while (more) {
      uint32_t c[U];
      double a[U];
#pragma unroll
      for (int u = 0; u < U; u++)
        if (k[u] < e[u]) { c[u] = col[k[u]]; a[u] = val[k[u]]; }
      more = false;
#pragma unroll
      for (int u = 0; u < U; u++)
        if (k[u] < e[u]) {
          s[u] = fma(a[u], __ldg(&x[c[u]]), s[u]);
          k[u] += LPR;
          more = more || k[u] < e[u];
        }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
#pragma unroll
      for (int o = LPR >> 1; o > 0; o >>= 1) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
    }
#pragma unroll
    for (int u = 0; u < U; u++)
      if (row[u] != 0xFFFFFFFFu && sub == 0) {
        y[row[u]] = s[u];
        if (DOTS) {
          const double xr = x[row[u]];
          pq = fma(xr, s[u], pq);
          pr = fma(xr, r[row[u]], pr);
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
// fused vector kernels
// ---------------------------------------------------------------------------------------------------------
// r = b, x = 0, p = 0 ; bb = rr = b.b ; it = 0        (pcg.cpp:67-81; x0 = 0 is assumed by the reference)
__global__ void __launch_bounds__(256) k_init_solve(const double *__restrict__ b, double *__restrict__ x,
                                                    double *__restrict__ r, double *__restrict__ p, uint32_t N,
                                                    uint32_t dot_limit, double *partials, int pstride,
                                                    unsigned int *counter, PcgScalars *scal) {
  __shared__ double sm[32];
  double s = 0.0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const double bi = b[i];
    r[i] = bi;
    x[i] = 0.0;
    p[i] = 0.0;
    if (i < dot_limit) s = fma(bi, bi, s);
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

// rz = sum of the backward solve's per-CTA partials (k_tri_post), in index order
__global__ void __launch_bounds__(256) k_sum_rz(const double *__restrict__ rz_partials, int n_partials, PcgScalars *scal) {
  __shared__ double sm[32];
  const double rz = sum_partials(rz_partials, n_partials, sm);
  if (threadIdx.x == 0) scal->rz = rz;
}

// beta = rz / rz_prev (0 in the first iteration, where the reference copies z into p, pcg.cpp:87-90); p = z + beta p.
__global__ void __launch_bounds__(256) k_p_update(const double *__restrict__ z, double *__restrict__ p, uint32_t N,
                                                  const PcgScalars *__restrict__ scal) {
  const int it = scal->it;
  const double beta = it == 0 ? 0.0 : scal->rz / scal->rz_prev;
  if (it == 0) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) p[i] = z[i];
  } else {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
      p[i] = fma(beta, p[i], z[i]);
  }
}

// p.q and p.r over the first `limit` rows (multi-GPU: after q's top rows have been all-reduced)
__global__ void __launch_bounds__(256) k_dots_pq_pr(const double *__restrict__ p, const double *__restrict__ q,
                                                    const double *__restrict__ r, uint32_t limit, double *partials,
                                                    int pstride, unsigned int *counter, PcgScalars *scal) {
  __shared__ double sm[32];
  double pq = 0.0, pr = 0.0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < limit; i += gridDim.x * blockDim.x) {
    const double pi = p[i];
    pq = fma(pi, q[i], pq);
    pr = fma(pi, r[i], pr);
  }
  double v[2] = {pq, pr};
  double *const out[2] = {&scal->pq, &scal->pr};
  publish_and_finalize<2>(v, partials, pstride, counter, out, sm);
}

// alpha = (p.r)/(p.q) ; x += alpha p ; r -= alpha q ; rr = r.r ; it++ ; rz_prev = rz      (pcg.cpp:101-110)
__global__ void __launch_bounds__(256) k_xr_update(const double *__restrict__ p, const double *__restrict__ q,
                                                   double *__restrict__ x, double *__restrict__ r, uint32_t N,
                                                   uint32_t dot_limit, double *partials, int pstride,
                                                   unsigned int *counter, PcgScalars *scal) {
  __shared__ double sm[32];
  const double alpha = scal->pr / scal->pq;
  double s = 0.0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const double pi = p[i];
    x[i] = fma(alpha, pi, x[i]);
    const double ri = fma(-alpha, q[i], r[i]);
    r[i] = ri;
    if (i < dot_limit) s = fma(ri, ri, s);
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
                                                       uint32_t N /* rows counted */, double *partials, int pstride,
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

int rcg_launch_init_solve(rcg_handle *h) {
  const uint32_t lim = h->dist.on ? h->dist.dot_limit : (uint32_t)h->N;
  k_init_solve<<<h->reduce_grid, 256, 0, h->stream>>>(h->b, h->x, h->r, h->p, (uint32_t)h->N, lim, h->partials,
                                                     h->partial_cap, h->counters + 1, h->scal);
  h->stats.kernel_launches += 1;
  RCG_CUDA(h, cudaGetLastError());
  if (h->dist.on) RCG_TRY(rcg_allreduce_sum(h, &h->scal->bb, 1));
  return RCG_OK;
}

int rcg_launch_sum_rz(rcg_handle *h) {
  const double *rz_part = h->partials + 2 * (size_t)h->partial_cap;
  k_sum_rz<<<1, 256, 0, h->stream>>>(rz_part, h->rz_slots, h->scal);
  h->stats.kernel_launches += 1;
  RCG_CUDA(h, cudaGetLastError());
  if (h->dist.on) RCG_TRY(rcg_allreduce_sum(h, &h->scal->rz, 1));
  return RCG_OK;
}

int rcg_launch_p_update(rcg_handle *h) {
  RCG_TRY(rcg_launch_sum_rz(h));
  k_p_update<<<h->reduce_grid, 256, 0, h->stream>>>(h->z, h->p, (uint32_t)h->N, h->scal);
  h->stats.kernel_launches += 1;
  RCG_CUDA(h, cudaGetLastError());
  return RCG_OK;
}

int rcg_launch_dots_pq_pr(rcg_handle *h) {
  k_dots_pq_pr<<<h->reduce_grid, 256, 0, h->stream>>>(h->p, h->q, h->r, h->dist.dot_limit, h->partials, h->partial_cap,
                                                     h->counters + 4, h->scal);
  h->stats.kernel_launches += 1;
  RCG_CUDA(h, cudaGetLastError());
  return rcg_allreduce_sum(h, &h->scal->pq, 2);   // pq and pr are adjacent
}

int rcg_launch_xr_update(rcg_handle *h) {
  const uint32_t lim = h->dist.on ? h->dist.dot_limit : (uint32_t)h->N;
  k_xr_update<<<h->reduce_grid, 256, 0, h->stream>>>(h->p, h->q, h->x, h->r, (uint32_t)h->N, lim, h->partials,
                                                    h->partial_cap, h->counters + 2, h->scal);
  h->stats.kernel_launches += 1;
  RCG_CUDA(h, cudaGetLastError());
  if (h->dist.on) RCG_TRY(rcg_allreduce_sum(h, &h->scal->rr, 1));
  k_advance<<<1, 1, 0, h->stream>>>(h->scal);
  h->stats.kernel_launches += 1;
  RCG_CUDA(h, cudaGetLastError());
  return RCG_OK;
}

int rcg_launch_residual_norm(rcg_handle *h, double *out_host_norm2) {
  // q = A x ; ||q - b||^2 -> scal->pq is reused as the output slot
  RCG_TRY(rcg_launch_spmv(h, h->x, h->q, nullptr, false));
  if (h->dist.on && h->N > h->dist.n_sub) RCG_TRY(rcg_allreduce_sum(h, h->q + h->dist.n_sub, h->N - h->dist.n_sub));
  const uint32_t lim = h->dist.on ? h->dist.dot_limit : (uint32_t)h->N;
  k_residual_norm<<<h->reduce_grid, 256, 0, h->stream>>>(h->q, h->b, lim, h->partials, h->partial_cap,
                                                        h->counters + 3, &h->scal->pq);
  h->stats.kernel_launches += 1;
  RCG_CUDA(h, cudaGetLastError());
  if (h->dist.on) RCG_TRY(rcg_allreduce_sum(h, &h->scal->pq, 1));
  RCG_CUDA(h, cudaMemcpyAsync(out_host_norm2, &h->scal->pq, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  RCG_CUDA(h, cudaStreamSynchronize(h->stream));
  return RCG_OK;
}
