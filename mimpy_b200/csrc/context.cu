// Context, error reporting and scratch management for libmimpy_b200.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void mimpy_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int mimpy_scratch(mimpy_ctx* ctx, size_t bytes, void** out) {
  if (bytes > ctx->scratch_bytes) {
    // stream-ordered: earlier kernels may still use the old block
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (ctx->scratch) CUDA_TRY(cudaFree(ctx->scratch));
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    size_t want = bytes + bytes / 4 + 4096;
    CUDA_TRY(cudaMalloc(&ctx->scratch, want));
    ctx->scratch_bytes = want;
  }
  *out = ctx->scratch;
  return MIMPY_OK;
}

// Synchronises the stream and reads the device error flag (counters[9]) that the bounded waits of
// the kernels raise instead of hanging: a wait that ran out is an ERROR, never a silent wrong result.
int mimpy_check_device_flag(mimpy_ctx* ctx) {
  unsigned int* host = reinterpret_cast<unsigned int*>(ctx->host_scalars + 32);
  CUDA_TRY(cudaMemcpyAsync(host, ctx->counters + 9, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (*host != 0u) {
    CUDA_TRY(cudaMemsetAsync(ctx->counters + 9, 0, sizeof(unsigned int), ctx->stream));
    mimpy_set_error("a bounded device-side wait timed out (peer-memory exchange with rank neighbours, or the "
                    "cross-CTA hand-over of the assembly kernel): results are invalid");
    return MIMPY_E_CUDA;
  }
  return MIMPY_OK;
}

extern "C" {

int mimpy_b200_version(void) { return 100; }

const char* mimpy_b200_last_error(void) { return g_err; }

int mimpy_b200_create(int device, void* stream, mimpy_ctx** out) {
  REQUIRE(out != nullptr, "out is NULL");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    mimpy_set_error("no CUDA device available (%s); libmimpy_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return MIMPY_E_CUDA;
  }
  REQUIRE(device >= 0 && device < count, "device index out of range");
  CUDA_TRY(cudaSetDevice(device));
  mimpy_ctx* c = new mimpy_ctx();
  memset(c, 0, sizeof(*c));
  c->device = device;
  c->stream = (cudaStream_t)stream;
  c->world = 1;
  CUDA_TRY(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
  CUDA_TRY(cudaMalloc(&c->partials, sizeof(double) * 16384));
  CUDA_TRY(cudaMalloc(&c->counters, sizeof(unsigned int) * 64));
  CUDA_TRY(cudaMemset(c->counters, 0, sizeof(unsigned int) * 64));
  CUDA_TRY(cudaMalloc(&c->scalars, sizeof(double) * 64));
  CUDA_TRY(cudaMemset(c->scalars, 0, sizeof(double) * 64));
  CUDA_TRY(cudaMallocHost(&c->host_scalars, sizeof(double) * 64));
  CUDA_TRY(cudaEventCreate(&c->ev0));
  CUDA_TRY(cudaEventCreate(&c->ev1));
  *out = c;
  return MIMPY_OK;
}

int mimpy_b200_destroy(mimpy_ctx* ctx) {
  if (!ctx) return MIMPY_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->diag_swap) cudaFree(ctx->diag_swap);
  cudaFree(ctx->partials);
  cudaFree(ctx->counters);
  cudaFree(ctx->scalars);
  cudaFreeHost(ctx->host_scalars);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  delete ctx;
  return MIMPY_OK;
}

int mimpy_b200_set_stream(mimpy_ctx* ctx, void* stream) {
  REQUIRE(ctx != nullptr, "ctx is NULL");
  ctx->stream = (cudaStream_t)stream;
  return MIMPY_OK;
}

int mimpy_b200_sync(mimpy_ctx* ctx) {
  REQUIRE(ctx != nullptr, "ctx is NULL");
  return mimpy_check_device_flag(ctx);
}

int mimpy_b200_set_comm(mimpy_ctx* ctx, void* nccl_comm, int32_t rank, int32_t world) {
  REQUIRE(ctx != nullptr, "ctx is NULL");
  ctx->nccl_comm = nccl_comm;
  ctx->rank = rank;
  ctx->world = world;
  return MIMPY_OK;
}

}  // extern "C"
