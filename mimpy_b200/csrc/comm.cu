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

// ---------------------------------------------------------------------------------------------
// NVLink peer-memory path.  The two collectives of a PCG iteration move a few bytes (3 doubles)
// and 512 KiB (one face layer): what they cost is LATENCY.  Every rank owns one symmetric buffer
// (cudaMalloc + cudaIpc handles); a rank PUSHES its contribution straight into its peers' buffers
// over NVLink (st.global on the mapped peer pointer), fences (system scope) and raises an epoch
// flag there; the receiver spins on its LOCAL copy of the flag.  Buffers are double-buffered by
// epoch parity: a rank can run at most one exchange ahead of its partner (it needs the partner's
// push to finish the current one), so parity e&1 is never overwritten while still being read.
// All waits are bounded and raise ctx error flag 1 instead of hanging.
//   layout of the symmetric buffer:  [0, 1024)    reduce slots  [parity][rank] x 32 B (3 doubles + epoch)
//                                    [1024, 2048) halo flags    [dir][parity] x u64
//                                    [4096, 16384) gather slots [parity][rank] x 528 B (64 doubles + epoch)
//                                    [16384, ...) halo data     [dir][parity][cap] doubles
// dir 0 = data arriving from the LOWER neighbour, dir 1 = from the UPPER neighbour.
// ---------------------------------------------------------------------------------------------
#define PEER_RED_OFF 0
#define PEER_FLAG_OFF 1024
#define PEER_GATHER_OFF 4096
#define PEER_GATHER_SLOT 528
#define PEER_GATHER_MAX 64   /* doubles per rank */
#define PEER_DATA_OFF 16384
#define PEER_SPIN_MAX (1u << 24)
// large all-gather (agglomerated multigrid levels, auxmg.cu): behind the halo data,
//   [parity][source rank] x PEER_BIGG_CAP doubles; flags [parity][source rank] x u64 at PEER_BIGG_FLAG_OFF
#define PEER_BIGG_CAP 32768
#define PEER_BIGG_FLAG_OFF 2048
#define PEER_BIGG_BYTES (sizeof(double) * 2 * 8 * (size_t)PEER_BIGG_CAP)
__device__ __host__ __forceinline__ size_t peer_bigg_off(size_t halo_cap) { return PEER_DATA_OFF + sizeof(double) * 4 * halo_cap; }

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// one warp; lane r talks to rank r.  Sum in rank order => identical bits on every rank.
__global__ void k_peer_allreduce(mimpy_peer pr, double* buf, int count, unsigned int* error_flag) {
  const int r = threadIdx.x;
  const unsigned long long e = pr.epoch[0] + 1;
  const int par = (int)(e & 1);
  double mine[3] = {0., 0., 0.};
  for (int k = 0; k < count; ++k) mine[k] = buf[k];
  double got[3] = {0., 0., 0.};
  if (r < pr.world) {
    char* dst = pr.peers[r] + PEER_RED_OFF + ((par * 8 + pr.rank) * 32);
    for (int k = 0; k < count; ++k)
      st_volatile_u64(reinterpret_cast<unsigned long long*>(dst) + k, (unsigned long long)__double_as_longlong(mine[k]));
    __threadfence_system();
    st_volatile_u64(reinterpret_cast<unsigned long long*>(dst) + 3, e);
    const char* src = pr.local + PEER_RED_OFF + ((par * 8 + r) * 32);
    unsigned int spins = 0;
    while (ld_volatile_u64(reinterpret_cast<const unsigned long long*>(src) + 3) != e && ++spins < PEER_SPIN_MAX) {
    }
    if (spins >= PEER_SPIN_MAX) atomicExch(error_flag, 1u);
    __threadfence_system();
    for (int k = 0; k < count; ++k)
      got[k] = __longlong_as_double((long long)ld_volatile_u64(reinterpret_cast<const unsigned long long*>(src) + k));
  }
  for (int k = 0; k < count; ++k) {
    double tot = 0.;
    for (int q = 0; q < pr.world; ++q) tot += __shfl_sync(0xffffffffu, got[k], q);
    if (r == 0) buf[k] = tot;
  }
  if (r == 0) pr.epoch[0] = e;
}

// pack the two interface layers and push them into the neighbours' buffers
__global__ void k_peer_halo_push(mimpy_peer pr, mimpy_halo h, const double* __restrict__ v, unsigned int* counter) {
  const unsigned long long e = pr.epoch[1] + 1;
  const int par = (int)(e & 1);
  const size_t cap = pr.halo_cap;
  const int total = h.n_lo + h.n_hi;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (i < h.n_lo) {  // to the lower neighbour: arrives there as "from the UPPER neighbour" (dir 1)
      double* dst = reinterpret_cast<double*>(pr.peers[h.rank_lo] + PEER_DATA_OFF) + (size_t)(1 * 2 + par) * cap;
      dst[i] = v[h.idx_lo[i]];
    } else {           // to the upper neighbour: dir 0 there
      const int k = i - h.n_lo;
      double* dst = reinterpret_cast<double*>(pr.peers[h.rank_hi] + PEER_DATA_OFF) + (size_t)(0 * 2 + par) * cap;
      dst[k] = v[h.idx_hi[k]];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(counter, 1u);
    if (done == gridDim.x - 1) {  // every block's stores are fenced: raise the flags
      __threadfence_system();
      if (h.n_lo)
        st_volatile_u64(reinterpret_cast<unsigned long long*>(pr.peers[h.rank_lo] + PEER_FLAG_OFF) + (1 * 2 + par), e);
      if (h.n_hi)
        st_volatile_u64(reinterpret_cast<unsigned long long*>(pr.peers[h.rank_hi] + PEER_FLAG_OFF) + (0 * 2 + par), e);
      *counter = 0u;
    }
  }
}

__global__ void k_peer_halo_add(mimpy_peer pr, mimpy_halo h, double* __restrict__ v, unsigned int* counter,
                                unsigned int* error_flag) {
  const unsigned long long e = pr.epoch[1] + 1;
  const int par = (int)(e & 1);
  const size_t cap = pr.halo_cap;
  if (threadIdx.x == 0) {
    const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(pr.local + PEER_FLAG_OFF);
    unsigned int spins = 0;
    if (h.n_lo)
      while (ld_volatile_u64(flags + (0 * 2 + par)) != e && ++spins < PEER_SPIN_MAX) {
      }
    if (h.n_hi)
      while (ld_volatile_u64(flags + (1 * 2 + par)) != e && ++spins < PEER_SPIN_MAX) {
      }
    if (spins >= PEER_SPIN_MAX) atomicExch(error_flag, 1u);
    __threadfence_system();
  }
  __syncthreads();
  const double* from_lo = reinterpret_cast<const double*>(pr.local + PEER_DATA_OFF) + (size_t)(0 * 2 + par) * cap;
  const double* from_hi = reinterpret_cast<const double*>(pr.local + PEER_DATA_OFF) + (size_t)(1 * 2 + par) * cap;
  const int total = h.n_lo + h.n_hi;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (i < h.n_lo) {
      v[h.idx_lo[i]] += __longlong_as_double((long long)ld_volatile_u64(reinterpret_cast<const unsigned long long*>(from_lo) + i));
    } else {
      const int k = i - h.n_lo;
      v[h.idx_hi[k]] += __longlong_as_double((long long)ld_volatile_u64(reinterpret_cast<const unsigned long long*>(from_hi) + k));
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(counter, 1u);
    if (done == gridDim.x - 1) {
      pr.epoch[1] = e;
      *counter = 0u;
    }
  }
}

// in-stream sum of `count` doubles over all ranks (no-op for a single rank)
int mimpy_allreduce_f64(mimpy_ctx* ctx, double* buf, int count) {
  if (ctx->world <= 1) return MIMPY_OK;
  if (ctx->peer.enabled && count <= 3) {
    k_peer_allreduce<<<1, 32, 0, ctx->stream>>>(ctx->peer, buf, count, ctx->counters + 9);
    KERNEL_CHECK();
    return MIMPY_OK;
  }
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
  if (ctx->peer.enabled && (size_t)h.n_lo <= ctx->peer.halo_cap && (size_t)h.n_hi <= ctx->peer.halo_cap) {
    const int total = h.n_lo + h.n_hi;
    const int blocks = min((total + 255) / 256, 2 * ctx->sm_count);
    k_peer_halo_push<<<blocks, 256, 0, st>>>(ctx->peer, h, v, ctx->counters + 10);
    k_peer_halo_add<<<blocks, 256, 0, st>>>(ctx->peer, h, v, ctx->counters + 11, ctx->counters + 9);
    KERNEL_CHECK();
    return MIMPY_OK;
  }
  if (!ctx->nccl_comm || !g_nccl.handle) {
    mimpy_set_error("halo exchange: world size %d, the interface (%d / %d rows) does not fit the peer buffers (%zu) and no "
                    "NCCL communicator is attached", ctx->world, h.n_lo, h.n_hi, ctx->peer.halo_cap);
    return MIMPY_E_NCCL;
  }
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

// all-gather of n <= 64 doubles per rank (coarsest level of auxmg.cu): thread (r, i) stores value i of
// this rank into rank r's slot [parity][this rank]; out[r * n + i] = value i of rank r
__global__ void __launch_bounds__(512)
k_peer_allgather(mimpy_peer pr, const double* __restrict__ mine, int n, double* __restrict__ out,
                 unsigned int* error_flag) {
  const unsigned long long e = pr.epoch[2] + 1;
  const int par = (int)(e & 1), t = threadIdx.x;
  const int r = t / n, i = t - r * n;
  if (r < pr.world) {
    char* dst = pr.peers[r] + PEER_GATHER_OFF + (size_t)(par * 8 + pr.rank) * PEER_GATHER_SLOT;
    st_volatile_u64(reinterpret_cast<unsigned long long*>(dst) + i, (unsigned long long)__double_as_longlong(mine[i]));
  }
  __threadfence_system();
  __syncthreads();
  if (t < pr.world) {
    char* dst = pr.peers[t] + PEER_GATHER_OFF + (size_t)(par * 8 + pr.rank) * PEER_GATHER_SLOT;
    st_volatile_u64(reinterpret_cast<unsigned long long*>(dst) + PEER_GATHER_MAX, e);
    const char* src = pr.local + PEER_GATHER_OFF + (size_t)(par * 8 + t) * PEER_GATHER_SLOT;
    unsigned int spins = 0;
    while (ld_volatile_u64(reinterpret_cast<const unsigned long long*>(src) + PEER_GATHER_MAX) != e && ++spins < PEER_SPIN_MAX) {
    }
    if (spins >= PEER_SPIN_MAX) atomicExch(error_flag, 1u);
    __threadfence_system();
  }
  __syncthreads();
  if (r < pr.world) {
    const char* src = pr.local + PEER_GATHER_OFF + (size_t)(par * 8 + r) * PEER_GATHER_SLOT;
    out[t] = __longlong_as_double((long long)ld_volatile_u64(reinterpret_cast<const unsigned long long*>(src) + i));
  }
  if (t == 0) pr.epoch[2] = e;
}

// all-gather of up to PEER_BIGG_CAP doubles per rank: block b pushes this rank's values into rank b's slot
// [parity][this rank] over NVLink and raises its flag there, then waits for rank b's values in the local slot
// [parity][b] and copies them out -- one kernel, no block depends on another.  Double-buffered by epoch
// parity like the other peer collectives (a rank can be at most one gather ahead of its peers).
__global__ void __launch_bounds__(1024)
k_peer_allgather_big(mimpy_peer pr, const double* __restrict__ mine, int n, double* __restrict__ out,
                     unsigned int* counter, unsigned int* error_flag) {
  const unsigned long long e = pr.epoch[3] + 1;
  const int par = (int)(e & 1), b = blockIdx.x, t = threadIdx.x;
  const size_t off = peer_bigg_off(pr.halo_cap);
  double* dst = reinterpret_cast<double*>(pr.peers[b] + off) + (size_t)(par * 8 + pr.rank) * PEER_BIGG_CAP;
  for (int i = t; i < n; i += blockDim.x) dst[i] = mine[i];
  __threadfence_system();
  __syncthreads();
  if (t == 0) {
    st_volatile_u64(reinterpret_cast<unsigned long long*>(pr.peers[b] + PEER_BIGG_FLAG_OFF) + par * 8 + pr.rank, e);
    const unsigned long long* flag = reinterpret_cast<const unsigned long long*>(pr.local + PEER_BIGG_FLAG_OFF) + par * 8 + b;
    unsigned int spins = 0;
    while (ld_volatile_u64(flag) != e && ++spins < PEER_SPIN_MAX) {
    }
    if (spins >= PEER_SPIN_MAX) atomicExch(error_flag, 1u);
    __threadfence_system();
  }
  __syncthreads();
  const unsigned long long* src = reinterpret_cast<const unsigned long long*>(pr.local + off) + (size_t)(par * 8 + b) * PEER_BIGG_CAP;
  for (int i = t; i < n; i += blockDim.x) out[(size_t)b * n + i] = __longlong_as_double((long long)ld_volatile_u64(src + i));
  __syncthreads();
  if (t == 0) {
    __threadfence();
    if (atomicAdd(counter, 1u) == gridDim.x - 1) {  // the last block closes the epoch
      pr.epoch[3] = e;
      *counter = 0u;
    }
  }
}

__global__ void k_gather_pad(int N, int n, int rank, const double* __restrict__ mine, double* __restrict__ out) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < N) out[g] = (g / n == rank) ? mine[g - rank * n] : 0.;
}

// out[r * n + i] = mine[i] of rank r.  Peer memory for small n, otherwise a zero-padded all-reduce.
int mimpy_allgather_f64(mimpy_ctx* ctx, const double* mine, int n, double* out) {
  if (n <= 0) return MIMPY_OK;
  if (ctx->world <= 1) {
    CUDA_TRY(cudaMemcpyAsync(out, mine, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
    return MIMPY_OK;
  }
  if (ctx->peer.enabled && n <= PEER_GATHER_MAX && ctx->world * n <= 512) {
    k_peer_allgather<<<1, 512, 0, ctx->stream>>>(ctx->peer, mine, n, out, ctx->counters + 9);
    KERNEL_CHECK();
    return MIMPY_OK;
  }
  if (ctx->peer.enabled && ctx->peer.bigg && n <= PEER_BIGG_CAP) {
    k_peer_allgather_big<<<ctx->world, 1024, 0, ctx->stream>>>(ctx->peer, mine, n, out, ctx->counters + 13, ctx->counters + 9);
    KERNEL_CHECK();
    return MIMPY_OK;
  }
  const int N = ctx->world * n;
  k_gather_pad<<<(N + 255) / 256, 256, 0, ctx->stream>>>(N, n, ctx->rank, mine, out);
  KERNEL_CHECK();
  return mimpy_allreduce_f64(ctx, out, N);
}

// ---------------------------------------------------------------------------------------------
// Ghost-layer exchange of a z-slab partition (auxmg.cu): n doubles starting at v + send_lo go to
// the lower neighbour (rank - 1) and arrive there at v + recv_hi; n doubles at v + send_hi go to the
// upper neighbour (rank + 1) and arrive at its v + recv_lo.  Plain copies (no sum).  Same buffers,
// flags and epoch as the halo exchange-add: exchanges on one stream are strictly sequential.
// ---------------------------------------------------------------------------------------------
__global__ void k_peer_layer_push(mimpy_peer pr, const double* __restrict__ v, mimpy_xchg x, int has_lo, int has_hi,
                                  unsigned int* counter) {
  const unsigned long long e = pr.epoch[1] + 1;
  const int par = (int)(e & 1);
  const size_t cap = pr.halo_cap;
  const int n_lo = has_lo ? x.n : 0, total = n_lo + (has_hi ? x.n : 0);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (i < n_lo) {
      double* dst = reinterpret_cast<double*>(pr.peers[pr.rank - 1] + PEER_DATA_OFF) + (size_t)(1 * 2 + par) * cap;
      dst[i] = v[x.send_lo + i];
    } else {
      const int k = i - n_lo;
      double* dst = reinterpret_cast<double*>(pr.peers[pr.rank + 1] + PEER_DATA_OFF) + (size_t)(0 * 2 + par) * cap;
      dst[k] = v[x.send_hi + k];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(counter, 1u);
    if (done == gridDim.x - 1) {
      __threadfence_system();
      if (has_lo)
        st_volatile_u64(reinterpret_cast<unsigned long long*>(pr.peers[pr.rank - 1] + PEER_FLAG_OFF) + (1 * 2 + par), e);
      if (has_hi)
        st_volatile_u64(reinterpret_cast<unsigned long long*>(pr.peers[pr.rank + 1] + PEER_FLAG_OFF) + (0 * 2 + par), e);
      *counter = 0u;
    }
  }
}

__global__ void k_peer_layer_pull(mimpy_peer pr, double* __restrict__ v, mimpy_xchg x, int has_lo, int has_hi,
                                  unsigned int* counter, unsigned int* error_flag) {
  const unsigned long long e = pr.epoch[1] + 1;
  const int par = (int)(e & 1);
  const size_t cap = pr.halo_cap;
  if (threadIdx.x == 0) {
    const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(pr.local + PEER_FLAG_OFF);
    unsigned int spins = 0;
    if (has_lo)
      while (ld_volatile_u64(flags + (0 * 2 + par)) != e && ++spins < PEER_SPIN_MAX) {
      }
    if (has_hi)
      while (ld_volatile_u64(flags + (1 * 2 + par)) != e && ++spins < PEER_SPIN_MAX) {
      }
    if (spins >= PEER_SPIN_MAX) atomicExch(error_flag, 1u);
    __threadfence_system();
  }
  __syncthreads();
  const unsigned long long* from_lo =
      reinterpret_cast<const unsigned long long*>(reinterpret_cast<const double*>(pr.local + PEER_DATA_OFF) + (size_t)(0 * 2 + par) * cap);
  const unsigned long long* from_hi =
      reinterpret_cast<const unsigned long long*>(reinterpret_cast<const double*>(pr.local + PEER_DATA_OFF) + (size_t)(1 * 2 + par) * cap);
  const int n_lo = has_lo ? x.n : 0, total = n_lo + (has_hi ? x.n : 0);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (i < n_lo) v[x.recv_lo + i] = __longlong_as_double((long long)ld_volatile_u64(from_lo + i));
    else v[x.recv_hi + (i - n_lo)] = __longlong_as_double((long long)ld_volatile_u64(from_hi + (i - n_lo)));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(counter, 1u);
    if (done == gridDim.x - 1) {
      pr.epoch[1] = e;
      *counter = 0u;
    }
  }
}

int mimpy_layer_exchange(mimpy_ctx* ctx, double* v, const mimpy_xchg& x) {
  if (ctx->world <= 1 || x.n <= 0) return MIMPY_OK;
  const int has_lo = ctx->rank > 0, has_hi = ctx->rank < ctx->world - 1;
  cudaStream_t st = ctx->stream;
  if (ctx->peer.enabled && (size_t)x.n <= ctx->peer.halo_cap) {
    const int total = x.n * (has_lo + has_hi);
    const int blocks = max(1, min((total + 255) / 256, 2 * ctx->sm_count));
    k_peer_layer_push<<<blocks, 256, 0, st>>>(ctx->peer, v, x, has_lo, has_hi, ctx->counters + 10);
    k_peer_layer_pull<<<blocks, 256, 0, st>>>(ctx->peer, v, x, has_lo, has_hi, ctx->counters + 11, ctx->counters + 9);
    KERNEL_CHECK();
    return MIMPY_OK;
  }
  if (!ctx->nccl_comm || !g_nccl.handle) {
    mimpy_set_error("world size %d but no NCCL communicator attached", ctx->world);
    return MIMPY_E_NCCL;
  }
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  NCCL_TRY(g_nccl.GroupStart());
  if (has_lo) {
    NCCL_TRY(g_nccl.Send(v + x.send_lo, x.n, ncclDouble, ctx->rank - 1, comm, st));
    NCCL_TRY(g_nccl.Recv(v + x.recv_lo, x.n, ncclDouble, ctx->rank - 1, comm, st));
  }
  if (has_hi) {
    NCCL_TRY(g_nccl.Send(v + x.send_hi, x.n, ncclDouble, ctx->rank + 1, comm, st));
    NCCL_TRY(g_nccl.Recv(v + x.recv_hi, x.n, ncclDouble, ctx->rank + 1, comm, st));
  }
  NCCL_TRY(g_nccl.GroupEnd());
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

// NVLink peer-memory set-up: every rank allocates its symmetric buffer and exports a cudaIpc
// handle (64 bytes); the caller all-gathers the handles (e.g. torch.distributed) and attaches.
int mimpy_b200_peer_alloc(mimpy_ctx* ctx, int64_t halo_cap_doubles, void* handle64_host) {
  REQUIRE(ctx && handle64_host && halo_cap_doubles >= 0, "bad argument");
  CUDA_TRY(cudaSetDevice(ctx->device));
  const size_t bytes = peer_bigg_off((size_t)halo_cap_doubles) + PEER_BIGG_BYTES;
  char* p = nullptr;
  CUDA_TRY(cudaMalloc(&p, bytes));
  CUDA_TRY(cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, p));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64_host, &h, 64);
  ctx->peer.local = p;
  ctx->peer.halo_cap = (size_t)halo_cap_doubles;
  ctx->peer.bigg = 1;
  return MIMPY_OK;
}

int mimpy_b200_peer_attach(mimpy_ctx* ctx, int32_t rank, int32_t world, const void* handles_host) {
  REQUIRE(ctx && handles_host && ctx->peer.local, "peer_alloc must be called first");
  REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, "peer path supports up to 8 ranks of one node");
  CUDA_TRY(cudaSetDevice(ctx->device));
  char* ptrs[8] = {nullptr};
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      ptrs[r] = ctx->peer.local;
    } else {
      cudaIpcMemHandle_t h;
      memcpy(&h, (const char*)handles_host + 64 * r, 64);
      void* q = nullptr;
      CUDA_TRY(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess));
      ptrs[r] = (char*)q;
    }
  }
  char** dev_ptrs = nullptr;
  CUDA_TRY(cudaMalloc(&dev_ptrs, sizeof(char*) * 8));
  CUDA_TRY(cudaMemcpy(dev_ptrs, ptrs, sizeof(char*) * 8, cudaMemcpyHostToDevice));
  unsigned long long* epoch = nullptr;
  CUDA_TRY(cudaMalloc(&epoch, sizeof(unsigned long long) * 4));
  CUDA_TRY(cudaMemset(epoch, 0, sizeof(unsigned long long) * 4));
  ctx->peer.peers = dev_ptrs;
  ctx->peer.epoch = epoch;
  ctx->peer.rank = rank;
  ctx->peer.world = world;
  ctx->peer.enabled = 1;
  ctx->rank = rank;
  ctx->world = world;
  return MIMPY_OK;
}

int mimpy_b200_peer_disable(mimpy_ctx* ctx) {
  REQUIRE(ctx, "ctx is NULL");
  ctx->peer.enabled = 0;
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
