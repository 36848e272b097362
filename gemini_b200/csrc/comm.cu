// Multi-GPU exchange of the MSM: one process per GPU, the SRS split by contiguous point range, every
// rank runs the whole pipeline on its range and the "all-reduce of partial G1 accumulators" is ONE ncclAllGather of the
// 192-byte XYZZ partials on the library's own stream followed by world-1 curve additions on every rank (NCCL has no
// curve-addition reduction operator).  No torch, no host staging: the collective is stream-ordered between the bucket
// reduction and the normalisation, so the whole sharded MSM is timed by CUDA events like the single-GPU one.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): a single-GPU caller never needs the library, and inside a
// torch process the loader hands back the NCCL build torch already mapped (same SONAME), so there is one NCCL per process.
#include <dlfcn.h>
#include <string.h>

#include <nccl.h>

#include "common.cuh"
#include "g1.cuh"
#include "msm.cuh"

struct gm_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  void* d_gather = nullptr;   // world * 256 B staging of the all-gathers
  size_t gather_bytes = 0;
};

namespace gm {

namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

int nccl_load() {
  std::lock_guard<std::mutex> guard(g_nccl_mu);
  if (g_nccl.handle) return GM_OK;
  const char* names[] = {getenv("GM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* nm : names) {
    if (!nm || !*nm) continue;
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    set_error("NCCL not found (dlopen libnccl.so.2: %s)", dlerror());
    return GM_ERR_STATE;
  }
  NcclApi a;
  a.handle = h;
  a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(h, "ncclAllGather"));
  a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(dlsym(h, "ncclGetVersion"));
  if (!a.GetUniqueId || !a.CommInitRank || !a.AllGather || !a.CommDestroy || !a.GetErrorString) {
    set_error("libnccl lacks a required symbol");
    return GM_ERR_STATE;
  }
  g_nccl = a;
  return GM_OK;
}
}  // namespace

#define GM_NCCL(expr)                                                                               \
  do {                                                                                              \
    ncclResult_t _r = (expr);                                                                       \
    if (_r != ncclSuccess) {                                                                        \
      ::gm::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r));      \
      return GM_ERR_CUDA;                                                                           \
    }                                                                                               \
  } while (0)

void comm_destroy(gm_ctx* ctx) {
  gm_comm* c = ctx->comm;
  if (!c) return;
  if (c->comm) g_nccl.CommDestroy(c->comm);
  if (c->d_gather) cudaFree(c->d_gather);
  delete c;
  ctx->comm = nullptr;
}

int comm_world(const gm_ctx* ctx) { return ctx->comm ? ctx->comm->world : 1; }
int comm_rank(const gm_ctx* ctx) { return ctx->comm ? ctx->comm->rank : 0; }

// every rank contributes `bytes` (<= 256) from d_send; returns the device staging that holds world * bytes
int comm_allgather_dev(gm_ctx* ctx, const void* d_send, size_t bytes, void** out_d_all) {
  gm_comm* c = ctx->comm;
  if (!c) { set_error("no communicator: call gm_comm_init first"); return GM_ERR_STATE; }
  if (bytes > 256) { set_error("all-gather payload larger than 256 bytes"); return GM_ERR_ARG; }
  GM_NCCL(g_nccl.AllGather(d_send, c->d_gather, bytes, ncclUint8, c->comm, ctx->stream));
  *out_d_all = c->d_gather;
  return GM_OK;
}

}  // namespace gm

using namespace gm;

extern "C" {

int gm_comm_unique_id(uint8_t out_id[GM_COMM_ID_BYTES]) {
  GM_ARG(out_id, "NULL argument");
  static_assert(sizeof(ncclUniqueId) == GM_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  GM_TRY(nccl_load());
  ncclUniqueId id;
  GM_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(out_id, &id, sizeof(id));
  return GM_OK;
}

int gm_comm_init(gm_ctx* ctx, const uint8_t id[GM_COMM_ID_BYTES], int rank, int world) {
  GM_ARG(ctx && id, "NULL argument");
  GM_ARG(world >= 1 && rank >= 0 && rank < world, "rank / world out of range");
  GM_ENTER(ctx);
  if (ctx->comm) { set_error("the context already has a communicator"); return GM_ERR_STATE; }
  GM_TRY(nccl_load());
  gm_comm* c = new (std::nothrow) gm_comm();
  if (!c) return GM_ERR_OOM;
  c->rank = rank;
  c->world = world;
  c->gather_bytes = (size_t)world * 256;
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, uid, rank);
  cudaError_t e = r == ncclSuccess ? cudaMalloc(&c->d_gather, c->gather_bytes + 256) : cudaSuccess;
  if (r != ncclSuccess || e != cudaSuccess) {
    set_error("gm_comm_init: %s", r != ncclSuccess ? g_nccl.GetErrorString(r) : cudaGetErrorString(e));
    if (c->comm) g_nccl.CommDestroy(c->comm);
    delete c;
    return GM_ERR_CUDA;
  }
  ctx->comm = c;
  // first collective: NCCL sets up its channels here, outside any timed region
  return gm_comm_barrier(ctx);
}

int gm_comm_rank(const gm_ctx* ctx) { return ctx ? comm_rank(ctx) : 0; }
int gm_comm_world(const gm_ctx* ctx) { return ctx ? comm_world(ctx) : 1; }
int gm_comm_nccl_version(void) {
  int v = 0;
  if (nccl_load() == GM_OK && g_nccl.GetVersion) g_nccl.GetVersion(&v);
  return v;
}

int gm_comm_barrier(gm_ctx* ctx) {
  GM_ARG(ctx, "ctx is NULL");
  GM_ENTER(ctx);
  gm_comm* c = ctx->comm;
  if (!c) return GM_OK;
  void* all = nullptr;
  uint8_t* send = reinterpret_cast<uint8_t*>(c->d_gather) + c->gather_bytes;
  GM_TRY(comm_allgather_dev(ctx, send, 8, &all));
  GM_CUDA(cudaStreamSynchronize(ctx->stream));
  return GM_OK;
}

int gm_comm_allgather(gm_ctx* ctx, const void* send, size_t bytes, void* recv_all) {
  GM_ARG(ctx && send && recv_all && bytes > 0 && bytes <= 256, "bad argument (1..256 bytes per rank)");
  GM_ENTER(ctx);
  gm_comm* c = ctx->comm;
  if (!c) { memcpy(recv_all, send, bytes); return GM_OK; }
  uint8_t* d_send = reinterpret_cast<uint8_t*>(c->d_gather) + c->gather_bytes;
  GM_CUDA(cudaMemcpyAsync(d_send, send, bytes, cudaMemcpyHostToDevice, ctx->stream));
  void* all = nullptr;
  GM_TRY(comm_allgather_dev(ctx, d_send, bytes, &all));
  GM_CUDA(cudaMemcpyAsync(recv_all, all, bytes * (size_t)c->world, cudaMemcpyDeviceToHost, ctx->stream));
  GM_CUDA(cudaStreamSynchronize(ctx->stream));
  return GM_OK;
}

}  // extern "C"
