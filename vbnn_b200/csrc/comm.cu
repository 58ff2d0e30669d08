// comm.cu -- data-parallel gradient exchange (new functionality: the reference is single-GPU,
// SURVEY.md section 2.2).  One rank per GPU; the {gW, gS, gb} arena of every layer is summed
// with one NCCL allreduce over NVLink 5 / NVSwitch.  libnccl is bound at run time with dlopen so
// that single-GPU users (and the LuaJIT host) need no NCCL at link time; when torch has already
// loaded its bundled libnccl.so.2 the same copy is reused.
#include <dlfcn.h>
#include <string.h>

#include "state.h"

namespace vbnn {

namespace {

typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void* ncclComm_t_;
enum { kNcclFloat32 = 7, kNcclSum = 0 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId_t*) = nullptr;
  int (*CommInitRank)(ncclComm_t_*, int, ncclUniqueId_t, int) = nullptr;
  int (*CommDestroy)(ncclComm_t_) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t_, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

NcclApi* nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) return nullptr;
  api.GetUniqueId = (int (*)(ncclUniqueId_t*))dlsym(api.handle, "ncclGetUniqueId");
  api.CommInitRank = (int (*)(ncclComm_t_*, int, ncclUniqueId_t, int))dlsym(api.handle, "ncclCommInitRank");
  api.CommDestroy = (int (*)(ncclComm_t_))dlsym(api.handle, "ncclCommDestroy");
  api.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t_, cudaStream_t))dlsym(api.handle, "ncclAllReduce");
  api.GetErrorString = (const char* (*)(int))dlsym(api.handle, "ncclGetErrorString");
  if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce) {
    dlclose(api.handle);
    api.handle = nullptr;
    return nullptr;
  }
  return &api;
}

#define VB_NCCL(api, expr)                                                                \
  do {                                                                                    \
    int _r = (expr);                                                                      \
    if (_r != 0) {                                                                        \
      set_error("%s failed: %s", #expr, (api)->GetErrorString ? (api)->GetErrorString(_r) : "?"); \
      return VBNN_E_NCCL;                                                                 \
    }                                                                                     \
  } while (0)

}  // namespace

int comm_allreduce_internal(vbnn_ctx* ctx, float* buf, size_t count, cudaStream_t st) {
  if (ctx->nranks <= 1) return VBNN_OK;
  NcclApi* api = nccl();
  VB_CHECK(api && ctx->nccl_comm, VBNN_E_NCCL, "allreduce without an initialised communicator");
  VB_NCCL(api, api->AllReduce(buf, buf, count, kNcclFloat32, kNcclSum, (ncclComm_t_)ctx->nccl_comm, st));
  return VBNN_OK;
}

}  // namespace vbnn

using namespace vbnn;

extern "C" int vbnn_comm_unique_id(void* id128) {
  VB_CHECK(id128, VBNN_E_INVALID, "null id buffer");
  NcclApi* api = nccl();
  VB_CHECK(api, VBNN_E_NCCL, "libnccl.so.2 not found");
  ncclUniqueId_t id;
  VB_NCCL(api, api->GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return VBNN_OK;
}

extern "C" int vbnn_comm_init(vbnn_ctx* ctx, const void* id128, int rank, int nranks) {
  VB_CHECK(ctx && id128 && nranks >= 1 && rank >= 0 && rank < nranks, VBNN_E_INVALID, "vbnn_comm_init: bad argument");
  VB_CHECK(ctx->nccl_comm == nullptr, VBNN_E_STATE, "communicator already initialised");
  NcclApi* api = nccl();
  VB_CHECK(api, VBNN_E_NCCL, "libnccl.so.2 not found");
  VB_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId_t id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t_ comm = nullptr;
  VB_NCCL(api, api->CommInitRank(&comm, nranks, id, rank));
  ctx->nccl_comm = comm;
  ctx->rank = rank;
  ctx->nranks = nranks;
  if (!ctx->comm_stream) VB_CUDA(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
  return VBNN_OK;
}

extern "C" int vbnn_comm_destroy(vbnn_ctx* ctx) {
  if (!ctx || !ctx->nccl_comm) return VBNN_OK;
  NcclApi* api = nccl();
  if (api) api->CommDestroy((ncclComm_t_)ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
  ctx->rank = 0;
  ctx->nranks = 1;
  if (ctx->comm_stream) { cudaStreamDestroy(ctx->comm_stream); ctx->comm_stream = nullptr; }
  return VBNN_OK;
}

extern "C" int vbnn_comm_allreduce(vbnn_ctx* ctx, float* buf_dev, size_t count) {
  VB_CHECK(ctx && buf_dev, VBNN_E_INVALID, "null argument");
  return comm_allreduce_internal(ctx, buf_dev, count, ctx->stream);
}
