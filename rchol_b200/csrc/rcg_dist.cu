// rchol_b200 -- multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch.
// The library resolves NCCL at run time (dlopen of libnccl.so.2: in a process that has already imported torch this
// is torch's bundled NCCL, otherwise the system one), so the single-GPU path has no NCCL dependency.
// Collectives per PCG iteration: one all-reduce of the subtree -> top-separator coupling (forward
// solve), one of the SpMV's top-separator rows, and three scalar all-reduces (r.z; p.q, p.r; r.r).
#include <dlfcn.h>

#include <cstring>

#include "rcg_common.cuh"

namespace {

typedef struct { char internal[128]; } nccl_unique_id;
typedef int (*fn_get_unique_id)(nccl_unique_id *);
typedef int (*fn_comm_init_rank)(void **, int, nccl_unique_id, int);
typedef int (*fn_comm_destroy)(void *);
typedef int (*fn_all_reduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef const char *(*fn_error_string)(int);

struct NcclApi {
  void *lib = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_all_reduce all_reduce = nullptr;
  fn_error_string error_string = nullptr;
} g_nccl;

constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;   // ncclDataType_t / ncclRedOp_t values (nccl.h)

bool load_nccl(std::string &err) {
  if (g_nccl.lib) return true;
  void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
  g_nccl.get_unique_id = (fn_get_unique_id)dlsym(lib, "ncclGetUniqueId");
  g_nccl.comm_init_rank = (fn_comm_init_rank)dlsym(lib, "ncclCommInitRank");
  g_nccl.comm_destroy = (fn_comm_destroy)dlsym(lib, "ncclCommDestroy");
  g_nccl.all_reduce = (fn_all_reduce)dlsym(lib, "ncclAllReduce");
  g_nccl.error_string = (fn_error_string)dlsym(lib, "ncclGetErrorString");
  if (!g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.comm_destroy || !g_nccl.all_reduce) {
    err = "libnccl.so.2 lacks a required symbol";
    return false;
  }
  g_nccl.lib = lib;
  return true;
}

}  // namespace

int rcg_allreduce_sum(rcg_handle *h, double *dev_ptr, size_t count) {
  if (!h->dist.on || h->dist.nranks == 1 || count == 0) return RCG_OK;
  const int rc = g_nccl.all_reduce(dev_ptr, dev_ptr, count, NCCL_FLOAT64, NCCL_SUM, h->dist.comm, h->stream);
  if (rc != 0) {
    h->err = std::string("ncclAllReduce: ") + (g_nccl.error_string ? g_nccl.error_string(rc) : "error");
    return RCG_ERR_CUDA;
  }
  return RCG_OK;
}

extern "C" {

int rcg_nccl_unique_id(void *out128) {
  std::string err;
  if (!out128 || !load_nccl(err)) return RCG_ERR_CUDA;
  nccl_unique_id id;
  if (g_nccl.get_unique_id(&id) != 0) return RCG_ERR_CUDA;
  memcpy(out128, &id, sizeof(id));
  return RCG_OK;
}

int rcg_dist_init(rcg_handle *h, int nranks, int rank, const void *unique_id128, uint64_t n_sub, int top_depth) {
  if (!h || nranks < 1 || rank < 0 || rank >= nranks || !unique_id128) return RCG_ERR_INVALID;
  if (n_sub >= 0xFFFFFFFFull) { h->err = "rcg_dist_init: n_sub does not fit 32 bits"; return RCG_ERR_INVALID; }
  // call order: before the matrices / vectors exist (their sizes depend on the layout) and only once per handle
  if (h->dist.comm || h->haveA || h->haveG || h->b) {
    h->err = "rcg_dist_init must be the first call on a handle (before rcg_set_matrix / rcg_set_factor), once";
    return RCG_ERR_STATE;
  }
  if (!load_nccl(h->err)) return RCG_ERR_CUDA;
  RCG_CUDA(h, cudaSetDevice(h->device));
  nccl_unique_id id;
  memcpy(&id, unique_id128, sizeof(id));
  void *comm = nullptr;
  const int rc = g_nccl.comm_init_rank(&comm, nranks, id, rank);
  if (rc != 0) {
    h->err = std::string("ncclCommInitRank: ") + (g_nccl.error_string ? g_nccl.error_string(rc) : "error");
    return RCG_ERR_CUDA;
  }
  h->dist.on = true;
  h->dist.rank = rank;
  h->dist.nranks = nranks;
  h->dist.comm = comm;
  h->dist.n_sub = (uint32_t)n_sub;
  h->dist.top_depth = top_depth;
  h->opt.use_graph = 0;   // collectives are enqueued between the kernels; no graph capture in this mode
  return RCG_OK;
}

int rcg_dist_finalize(rcg_handle *h) {
  if (!h) return RCG_ERR_INVALID;
  if (h->dist.comm) { g_nccl.comm_destroy(h->dist.comm); h->dist.comm = nullptr; }
  if (h->dist.sbuf) { cudaFree(h->dist.sbuf); h->dist.sbuf = nullptr; }
  h->dist.on = false;
  return RCG_OK;
}

}  // extern "C"
