// Row-sharded evaluation (BASELINE config 5): every rank holds all chains and N/G rows of X; after
// each gradient evaluation the [chains, ld] gradient and the [chains] log-likelihoods are summed over the ranks.
// The reference has no counterpart (host multiprocessing only).
//
// The collective is NCCL's (NVLink 5 / NVSwitch; NVLS reduces inside the switch when available) and it is enqueued
// from the C driver, on the context's stream, as ONE grouped call per evaluation: ncclGroupStart, all-reduce of the
// fp32 gradient, all-reduce of the fp64 statistics, ncclGroupEnd -- NCCL fuses a group into a single launch.  (Round
// 1 issued two separate torch.distributed collectives from a Python callback per evaluation.)
//
// libnccl is resolved at run time with dlopen/dlsym: libbhmc.so has no link-time dependency on it, single-GPU users
// never load it, and inside a PyTorch process the copy torch already loaded (libnccl.so.2) is the one that is used.
#include <dlfcn.h>
#include <nccl.h>  // types and enums only; no symbol of libnccl is referenced at link time

#include "internal.cuh"

namespace bhmc {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  bool ok = false;
};

static const NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.ok ? &api : nullptr;
  tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the host process (PyTorch) already uses
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    set_error("libnccl.so.2 not found (%s): row-sharded runs need NCCL", dlerror());
    return nullptr;
  }
#define BHMC_SYM(field, name)                                                  \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));           \
  if (!api.field) {                                                            \
    set_error("libnccl: symbol %s not found", name);                           \
    return nullptr;                                                            \
  }
  BHMC_SYM(GetUniqueId, "ncclGetUniqueId")
  BHMC_SYM(CommInitRank, "ncclCommInitRank")
  BHMC_SYM(CommDestroy, "ncclCommDestroy")
  BHMC_SYM(AllReduce, "ncclAllReduce")
  BHMC_SYM(GroupStart, "ncclGroupStart")
  BHMC_SYM(GroupEnd, "ncclGroupEnd")
  BHMC_SYM(GetErrorString, "ncclGetErrorString")
  BHMC_SYM(GetVersion, "ncclGetVersion")
#undef BHMC_SYM
  api.ok = true;
  return &api;
}

#define BHMC_NCCL_OK(api, expr)                                                                  \
  do {                                                                                           \
    ncclResult_t r__ = (expr);                                                                   \
    if (r__ != ncclSuccess) {                                                                    \
      set_error("%s failed: %s (%s:%d)", #expr, (api)->GetErrorString(r__), __FILE__, __LINE__); \
      return BHMC_ERR_CUDA;                                                                      \
    }                                                                                            \
  } while (0)

}  // namespace bhmc

struct bhmc_comm {
  bhmc_ctx* ctx = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  bool owned = false;
};

using namespace bhmc;

extern "C" {

int bhmc_comm_unique_id(uint8_t* id_out) {
  BHMC_CHECK_ARG(id_out, "id_out is NULL");
  static_assert(sizeof(ncclUniqueId) == BHMC_COMM_ID_BYTES, "ncclUniqueId size");
  const NcclApi* api = nccl_api();
  if (!api) return BHMC_ERR_UNSUPPORTED;
  ncclUniqueId id;
  BHMC_NCCL_OK(api, api->GetUniqueId(&id));
  memcpy(id_out, &id, sizeof(id));
  return BHMC_OK;
}

int bhmc_comm_create(bhmc_ctx* ctx, const uint8_t* id, int32_t rank, int32_t world, bhmc_comm** out) {
  BHMC_CHECK_ARG(ctx && id && out && world >= 1 && rank >= 0 && rank < world, "bad argument");
  const NcclApi* api = nccl_api();
  if (!api) return BHMC_ERR_UNSUPPORTED;
  BHMC_CUDA_OK(cudaSetDevice(ctx->device));
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  auto* c = new (std::nothrow) bhmc_comm();
  if (!c) return BHMC_ERR_NOMEM;
  c->ctx = ctx;
  c->rank = rank;
  c->world = world;
  c->owned = true;
  ncclResult_t r = api->CommInitRank(&c->comm, world, uid, rank);
  if (r != ncclSuccess) {
    set_error("ncclCommInitRank(rank %d of %d) failed: %s", rank, world, api->GetErrorString(r));
    delete c;
    return BHMC_ERR_CUDA;
  }
  *out = c;
  return BHMC_OK;
}

int bhmc_comm_wrap(bhmc_ctx* ctx, void* nccl_comm, int32_t rank, int32_t world, bhmc_comm** out) {
  BHMC_CHECK_ARG(ctx && nccl_comm && out && world >= 1 && rank >= 0 && rank < world, "bad argument");
  if (!nccl_api()) return BHMC_ERR_UNSUPPORTED;
  auto* c = new (std::nothrow) bhmc_comm();
  if (!c) return BHMC_ERR_NOMEM;
  c->ctx = ctx;
  c->comm = (ncclComm_t)nccl_comm;
  c->rank = rank;
  c->world = world;
  c->owned = false;
  *out = c;
  return BHMC_OK;
}

int bhmc_comm_destroy(bhmc_comm* c) {
  if (!c) return BHMC_OK;
  const NcclApi* api = nccl_api();
  if (api && c->owned && c->comm) {
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    api->CommDestroy(c->comm);
  }
  delete c;
  return BHMC_OK;
}

int32_t bhmc_comm_world(const bhmc_comm* c) { return c ? c->world : 0; }

int bhmc_nccl_version(void) {
  const NcclApi* api = nccl_api();
  int v = 0;
  if (!api || api->GetVersion(&v) != ncclSuccess) return 0;
  return v;
}

int bhmc_allreduce_grad(bhmc_ctx* ctx, void* nccl_comm, float* g_dev, int64_t g_count, double* stat_dev, int32_t n_stat) {
  BHMC_CHECK_ARG(ctx && nccl_comm, "NULL argument");
  const NcclApi* api = nccl_api();
  if (!api) return BHMC_ERR_UNSUPPORTED;
  ncclComm_t comm = (ncclComm_t)nccl_comm;
  const bool two = g_dev && g_count > 0 && stat_dev && n_stat > 0;
  if (two) BHMC_NCCL_OK(api, api->GroupStart());
  if (g_dev && g_count > 0)
    BHMC_NCCL_OK(api, api->AllReduce(g_dev, g_dev, (size_t)g_count, ncclFloat32, ncclSum, comm, ctx->stream));
  if (stat_dev && n_stat > 0)
    BHMC_NCCL_OK(api, api->AllReduce(stat_dev, stat_dev, (size_t)n_stat, ncclFloat64, ncclSum, comm, ctx->stream));
  if (two) BHMC_NCCL_OK(api, api->GroupEnd());
  ctx->launches++;  // one fused NCCL kernel
  return BHMC_OK;
}

int bhmc_comm_allreduce(bhmc_comm* c, float* g_dev, int64_t g_count, double* stat_dev, int32_t n_stat) {
  BHMC_CHECK_ARG(c, "comm is NULL");
  if (c->world == 1) return BHMC_OK;
  return bhmc_allreduce_grad(c->ctx, c->comm, g_dev, g_count, stat_dev, n_stat);
}

}  // extern "C"
