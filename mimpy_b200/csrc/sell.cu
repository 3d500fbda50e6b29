// Sliced-ELLPACK (SELL-32) layout of the face system.  Rows of A have ~11 entries (hexes); in CSR a
// row is 88+44 bytes, far below a warp's 32x8-byte coalescing width, so a CSR SpMV either wastes
// lanes (k lanes per row) or coalescing (lane per row).  In SELL-32 the 32 rows of a slice are
// stored column-major, padded to the longest row of the slice: lane r of a warp owns row r of the
// slice and every load of the warp is one contiguous 256-byte (values) / 128-byte (columns) line.
//   entry k of row g lives at  sell_ptr[g >> 5] + 32*k + (g & 31)
// Padding entries have value 0 and the row's own index as column.  The pattern is derived once
// from the CSR pattern of the M block (m_indptr/m_indices); hybrid_setup writes values straight
// into this layout (no conversion pass).
#include <cub/device/device_scan.cuh>

#include "common.cuh"

__global__ void k_sell_width(int n, int nslices, const int* __restrict__ indptr, int* __restrict__ slice_entries) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslices) return;
  int w = 0;
  const int r1 = min(n, 32 * s + 32);
  for (int r = 32 * s; r < r1; ++r) w = max(w, indptr[r + 1] - indptr[r]);
  slice_entries[s] = 32 * w;
}

__global__ void k_sell_fill(int n, const int* __restrict__ indptr, const int* __restrict__ indices,
                            const int* __restrict__ sell_ptr, int* __restrict__ sell_cols) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = row >> 5, lane = row & 31;
  if (32 * s >= n) return;
  const int base = sell_ptr[s];
  const int w = (sell_ptr[s + 1] - base) >> 5;
  if (row < n) {
    const int b = indptr[row], len = indptr[row + 1] - b;
    for (int k = 0; k < w; ++k) sell_cols[base + 32 * k + lane] = (k < len) ? indices[b + k] : row;
  } else {  // rows past the end of the last slice
    for (int k = 0; k < w; ++k) sell_cols[base + 32 * k + lane] = n - 1;
  }
}

extern "C" {

// sell_ptr: (n+31)/32 + 1 ints; returns the padded entry count through *total_host (synchronises)
int mimpy_b200_sell_sizes(mimpy_ctx* ctx, int32_t n, const int32_t* indptr, int32_t* sell_ptr,
                          int64_t* total_host) {
  REQUIRE(ctx && indptr && sell_ptr && total_host, "NULL argument");
  REQUIRE(n > 0, "empty matrix");
  const int nslices = (n + 31) / 32;
  cudaStream_t st = ctx->stream;
  int* tmp = nullptr;
  CUDA_TRY(cudaMalloc(&tmp, sizeof(int) * (size_t)(nslices + 1)));
  cudaMemsetAsync(tmp, 0, sizeof(int) * (size_t)(nslices + 1), st);
  k_sell_width<<<(nslices + 255) / 256, 256, 0, st>>>(n, nslices, indptr, tmp);
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, tmp, sell_ptr, nslices + 1, st);
  void* scratch = nullptr;
  int rc = mimpy_scratch(ctx, bytes, &scratch);
  if (rc == MIMPY_OK) {
    cub::DeviceScan::ExclusiveSum(scratch, bytes, tmp, sell_ptr, nslices + 1, st);
    int total = 0;
    cudaMemcpyAsync(&total, sell_ptr + nslices, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) {
      mimpy_set_error("sell_sizes: %s", cudaGetErrorString(e));
      rc = MIMPY_E_CUDA;
    } else if (total < 0) {
      mimpy_set_error("padded face system exceeds int32 indexing");
      rc = MIMPY_E_UNSUPPORTED;
    }
    *total_host = total;
  }
  cudaFree(tmp);
  return rc;
}

int mimpy_b200_sell_fill(mimpy_ctx* ctx, int32_t n, const int32_t* indptr, const int32_t* indices,
                         const int32_t* sell_ptr, int32_t* sell_cols) {
  REQUIRE(ctx && indptr && indices && sell_ptr && sell_cols, "NULL argument");
  const int nslices = (n + 31) / 32;
  k_sell_fill<<<(32 * nslices + 255) / 256, 256, 0, ctx->stream>>>(n, indptr, indices, sell_ptr, sell_cols);
  KERNEL_CHECK();
  return MIMPY_OK;
}

}  // extern "C"
