// Multi-GPU plumbing for the slab-partitioned PCG: NCCL over NVLink is used for
// exactly two things per iteration -- the exchange-add of the interface face layers after the
// SpMV and the all-reduce of the dot products.
//
// NCCL is NOT a link-time dependency: the process already holds the NCCL that PyTorch loaded
// (torch bundles its own libnccl.so.2); linking a second copy with the same SONAME breaks
// whichever is loaded later.  The entry points are resolved with dlopen/dlsym from the library
// that is already resident (or from an explicit path).
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only
#include <string.h>

#include "common.cuh"

struct NcclApi {
  void* handle;
  decltype(&ncclGetUniqueId) GetUniqueId;
  decltype(&ncclCommInitRank) CommInitRank;
  decltype(&ncclCommDestroy) CommDestroy;
  decltype(&ncclAllReduce) AllReduce;
  decltype(&ncclSend) Send;
  decltype(&ncclRecv) Recv;
  decltype(&ncclGroupStart) GroupStart;
  decltype(&ncclGroupEnd) GroupEnd;
  decltype(&ncclGetErrorString) GetErrorString;
};
static NcclApi g_nccl = {};

static int nccl_load(const char* path) {
  if (g_nccl.handle) return MIMPY_OK;
  void* h = nullptr;
  if (path && path[0]) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // already resident?
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    mimpy_set_error("cannot load NCCL: %s", dlerror());
    return MIMPY_E_NCCL;
  }
#define RESOLVE(field, name)                                          \
  g_nccl.field = (decltype(g_nccl.field))dlsym(h, name);              \
  if (!g_nccl.field) {                                                \
    mimpy_set_error("NCCL symbol %s not found", name);                \
    return MIMPY_E_NCCL;                                              \
  }
  RESOLVE(GetUniqueId, "ncclGetUniqueId");
  RESOLVE(CommInitRank, "ncclCommInitRank");
  RESOLVE(CommDestroy, "ncclCommDestroy");
  RESOLVE(AllReduce, "ncclAllReduce");
  RESOLVE(Send, "ncclSend");
  RESOLVE(Recv, "ncclRecv");
  RESOLVE(GroupStart, "ncclGroupStart");
  RESOLVE(GroupEnd, "ncclGroupEnd");
  RESOLVE(GetErrorString, "ncclGetErrorString");
#undef RESOLVE
  g_nccl.handle = h;
  return MIMPY_OK;
}

#define NCCL_TRY(expr)                                                              \
  do {                                                                              \
    ncclResult_t _r = (expr);                                                       \
    if (_r != ncclSuccess) {                                                        \
      mimpy_set_error("%s failed: %s", #expr, g_nccl.GetErrorString(_r));           \
      return MIMPY_E_NCCL;                                                          \
    }                                                                               \
  } while (0)

// in-stream sum of `count` doubles over all ranks (no-op for a single rank)
int mimpy_allreduce_f64(mimpy_ctx* ctx, double* buf, int count) {
  if (ctx->world <= 1) return MIMPY_OK;
  if (!ctx->nccl_comm || !g_nccl.handle) {
    mimpy_set_error("world size %d but no NCCL communicator attached", ctx->world);
    return MIMPY_E_NCCL;
  }
  NCCL_TRY(g_nccl.AllReduce(buf, buf, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  return MIMPY_OK;
}

__global__ void k_pack(int n, const int* __restrict__ idx, const double* __restrict__ v, double* __restrict__ buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) buf[i] = v[idx[i]];
}

__global__ void k_unpack_add(int n, const int* __restrict__ idx, const double* __restrict__ buf, double* __restrict__ v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[idx[i]] += buf[i];
}

// v[idx_lo] and v[idx_hi] hold this rank's PARTIAL sums on the interface layers shared with the
// lower / upper neighbour; afterwards both ranks hold own + neighbour (a + b is commutative, so the
// two copies are bit-identical).
int mimpy_halo_exchange_add(mimpy_ctx* ctx, double* v) {
  const mimpy_halo& h = ctx->halo;
  if (ctx->world <= 1 || (h.n_lo == 0 && h.n_hi == 0)) return MIMPY_OK;
  cudaStream_t st = ctx->stream;
  double* send_lo = h.buf;
  double* recv_lo = h.buf + h.n_lo;
  double* send_hi = h.buf + 2 * (size_t)h.n_lo;
  double* recv_hi = send_hi + h.n_hi;
  if (h.n_lo) k_pack<<<(h.n_lo + 255) / 256, 256, 0, st>>>(h.n_lo, h.idx_lo, v, send_lo);
  if (h.n_hi) k_pack<<<(h.n_hi + 255) / 256, 256, 0, st>>>(h.n_hi, h.idx_hi, v, send_hi);
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  NCCL_TRY(g_nccl.GroupStart());
  if (h.n_lo) {
    NCCL_TRY(g_nccl.Send(send_lo, h.n_lo, ncclDouble, h.rank_lo, comm, st));
    NCCL_TRY(g_nccl.Recv(recv_lo, h.n_lo, ncclDouble, h.rank_lo, comm, st));
  }
  if (h.n_hi) {
    NCCL_TRY(g_nccl.Send(send_hi, h.n_hi, ncclDouble, h.rank_hi, comm, st));
    NCCL_TRY(g_nccl.Recv(recv_hi, h.n_hi, ncclDouble, h.rank_hi, comm, st));
  }
  NCCL_TRY(g_nccl.GroupEnd());
  if (h.n_lo) k_unpack_add<<<(h.n_lo + 255) / 256, 256, 0, st>>>(h.n_lo, h.idx_lo, recv_lo, v);
  if (h.n_hi) k_unpack_add<<<(h.n_hi + 255) / 256, 256, 0, st>>>(h.n_hi, h.idx_hi, recv_hi, v);
  KERNEL_CHECK();
  return MIMPY_OK;
}

extern "C" {

int mimpy_b200_nccl_load(const char* path) { return nccl_load(path); }

int mimpy_b200_nccl_unique_id(void* out128) {
  REQUIRE(out128, "NULL argument");
  int rc = nccl_load(nullptr);
  if (rc) return rc;
  ncclUniqueId id;
  NCCL_TRY(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out128, &id, 128);
  return MIMPY_OK;
}

int mimpy_b200_nccl_init(mimpy_ctx* ctx, const void* id128, int32_t rank, int32_t world) {
  REQUIRE(ctx && id128, "NULL argument");
  REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank/world");
  int rc = nccl_load(nullptr);
  if (rc) return rc;
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t comm = nullptr;
  CUDA_TRY(cudaSetDevice(ctx->device));
  NCCL_TRY(g_nccl.CommInitRank(&comm, world, id, rank));
  ctx->nccl_comm = comm;
  ctx->rank = rank;
  ctx->world = world;
  return MIMPY_OK;
}

int mimpy_b200_set_halo(mimpy_ctx* ctx, int32_t n_owned, int32_t n_lo, const int32_t* idx_lo,
                        int32_t rank_lo, int32_t n_hi, const int32_t* idx_hi, int32_t rank_hi,
                        double* buf) {
  REQUIRE(ctx, "ctx is NULL");
  REQUIRE((n_lo == 0 || idx_lo) && (n_hi == 0 || idx_hi), "index arrays missing");
  REQUIRE((n_lo + n_hi == 0) || buf, "exchange buffer missing");
  ctx->halo.n_owned = n_owned;
  ctx->halo.n_lo = n_lo;
  ctx->halo.idx_lo = idx_lo;
  ctx->halo.rank_lo = rank_lo;
  ctx->halo.n_hi = n_hi;
  ctx->halo.idx_hi = idx_hi;
  ctx->halo.rank_hi = rank_hi;
  ctx->halo.buf = buf;
  return MIMPY_OK;
}

int mimpy_b200_halo_exchange_add(mimpy_ctx* ctx, double* v) {
  REQUIRE(ctx && v, "NULL argument");
  return mimpy_halo_exchange_add(ctx, v);
}

int mimpy_b200_allreduce_sum(mimpy_ctx* ctx, double* buf, int32_t count) {
  REQUIRE(ctx && buf, "NULL argument");
  return mimpy_allreduce_f64(ctx, buf, count);
}

}  // extern "C"
