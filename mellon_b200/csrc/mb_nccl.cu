// NCCL plumbing: one rank per process / GPU.  libnccl is opened at run time (dlopen) so the
// library loads on a box without NCCL and so that a process that already loaded torch's
// bundled libnccl.so.2 shares that copy instead of pulling in a second one.
#include <dlfcn.h>

#include "mb_common.cuh"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct NcclApi {
  void* handle;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char* (*GetErrorString)(ncclResult_t);
};

NcclApi g_nccl = {};

int load_nccl() {
  if (g_nccl.handle) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* nm : names) {
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  const char* env = getenv("MELLON_B200_NCCL_LIB");
  if (!h && env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  MB_CHECK(h, "could not dlopen libnccl.so.2 (%s); set MELLON_B200_NCCL_LIB to its path", dlerror());
#define SYM(field, name)                                               \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name)); \
  MB_CHECK(g_nccl.field, "libnccl is missing symbol %s", name);
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllReduce, "ncclAllReduce")
  SYM(AllGather, "ncclAllGather")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(Broadcast, "ncclBroadcast")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl.handle = h;
  return 0;
}

#define MB_NCCL(call)                                                                     \
  do {                                                                                    \
    ncclResult_t _r = (call);                                                             \
    if (_r != 0) {                                                                        \
      mb_set_error("%s:%d NCCL error %d: %s", __FILE__, __LINE__, _r, g_nccl.GetErrorString(_r)); \
      return -3;                                                                          \
    }                                                                                     \
  } while (0)

}  // namespace

extern "C" int mb_comm_unique_id(unsigned char* out128) {
  MB_CHECK(out128, "mb_comm_unique_id: null output");
  MB_TRY(load_nccl());
  ncclUniqueId id;
  MB_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(out128, id.internal, 128);
  return 0;
}

extern "C" int mb_comm_init(mb_ctx* ctx, const unsigned char* id128, int rank, int world) {
  MB_CHECK(ctx && id128, "mb_comm_init: null argument");
  MB_CHECK(world >= 1 && rank >= 0 && rank < world, "mb_comm_init: rank %d of %d", rank, world);
  MB_CHECK(ctx->comm == nullptr, "mb_comm_init: communicator already attached");
  MB_CUDA(cudaSetDevice(ctx->device));
  ctx->rank = rank;
  ctx->world = world;
  if (world == 1) return 0;
  MB_TRY(load_nccl());
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  ncclComm_t comm;
  MB_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
  ctx->comm = comm;
  return 0;
}

extern "C" int mb_comm_destroy(mb_ctx* ctx) {
  if (!ctx || !ctx->comm) return 0;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  g_nccl.CommDestroy(reinterpret_cast<ncclComm_t>(ctx->comm));
  ctx->comm = nullptr;
  ctx->rank = 0;
  ctx->world = 1;
  return 0;
}

extern "C" int mb_comm_solo(mb_ctx* ctx, int on) {
  MB_CHECK(ctx, "mb_comm_solo: null ctx");
  ctx->solo = on != 0;
  return 0;
}

extern "C" int mb_comm_info(mb_ctx* ctx, int* rank, int* world) {
  MB_CHECK(ctx, "mb_comm_info: null ctx");
  if (rank) *rank = ctx->rank;
  if (world) *world = ctx->world;
  return 0;
}

int mb_allreduce_raw(mb_ctx* ctx, double* p, int64_t count) {
  if (!ctx->comm || ctx->world == 1 || ctx->solo || count == 0) return 0;
  MB_NCCL(g_nccl.AllReduce(p, p, (size_t)count, ncclFloat64, ncclSum, reinterpret_cast<ncclComm_t>(ctx->comm),
                           ctx->stream));
  return 0;
}

int mb_sendrecv_raw(mb_ctx* ctx, const double* send, double* recv, int64_t count, int peer) {
  MB_CHECK(ctx->comm && peer >= 0 && peer < ctx->world && peer != ctx->rank, "mb_sendrecv_raw: bad peer %d", peer);
  if (count == 0) return 0;
  ncclComm_t comm = reinterpret_cast<ncclComm_t>(ctx->comm);
  MB_NCCL(g_nccl.GroupStart());
  ncclResult_t a = g_nccl.Send(send, (size_t)count, ncclFloat64, peer, comm, ctx->stream);
  ncclResult_t b = g_nccl.Recv(recv, (size_t)count, ncclFloat64, peer, comm, ctx->stream);
  MB_NCCL(g_nccl.GroupEnd());
  MB_NCCL(a);
  MB_NCCL(b);
  return 0;
}

int mb_bcast_raw(mb_ctx* ctx, double* p, int64_t count, int root) {
  if (!ctx->comm || ctx->world == 1 || count == 0) return 0;
  MB_NCCL(g_nccl.Broadcast(p, p, (size_t)count, ncclFloat64, root, reinterpret_cast<ncclComm_t>(ctx->comm), ctx->stream));
  return 0;
}

extern "C" int mb_comm_allreduce(mb_ctx* ctx, mb_mat* a) {
  MB_CHECK(ctx && a, "mb_comm_allreduce: null argument");
  MB_CUDA(cudaSetDevice(ctx->device));
  return mb_allreduce_raw(ctx, a->p, a->rows * a->cols);
}

extern "C" int mb_comm_allgather(mb_ctx* ctx, const mb_mat* a, mb_mat* out) {
  MB_CHECK(ctx && a && out, "mb_comm_allgather: null argument");
  MB_CHECK(out->rows == a->rows * ctx->world && out->cols == a->cols,
           "mb_comm_allgather: output must be (%lld, %lld)", (long long)(a->rows * ctx->world),
           (long long)a->cols);
  MB_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->comm || ctx->world == 1) return mb_mat_copy(ctx, a, out);
  MB_NCCL(g_nccl.AllGather(a->p, out->p, (size_t)(a->rows * a->cols), ncclFloat64,
                           reinterpret_cast<ncclComm_t>(ctx->comm), ctx->stream));
  return 0;
}
