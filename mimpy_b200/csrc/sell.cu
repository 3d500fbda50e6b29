// Sliced-ELLPACK (SELL-32) layout of the face system.  Rows of A have ~11 entries (hexes); in CSR a
// row is 88+44 bytes, far below a warp's 32x8-byte coalescing width, so a CSR SpMV either wastes
// lanes (k lanes per row) or coalescing (lane per row).  In SELL-32 the 32 rows of a slice are
// stored column-major, padded to the longest row of the slice: lane r of a warp owns row r of the
// slice and every load of the warp is one contiguous 256-byte (values) / 128-byte (columns) line.
//   entry k of row g lives at  sell_ptr[g >> 5] + 32*k + (g & 31)
// Padding entries have value 0 and the row's own index as column.  The pattern is derived once
// from the CSR pattern of the M block (m_indptr/m_indices); hybrid_setup writes values straight
// into this layout (no conversion pass).
#include <stdlib.h>

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

// ---- column patterns ------------------------------------------------------------------------------
// On a structured mesh the columns of almost every row are the row index plus one of a handful of offset
// vectors (one per face type): instead of 4 bytes per ENTRY (44 B of the 152 B a row costs the SpMV) a row
// stores a one-byte pattern id, and the offset vectors of the 254 most frequent patterns live in shared
// memory.  Id 255 is the escape: that row reads its columns from sell_cols as before (rows next to faces
// that are not flux DOFs have offsets that depend on their position: ~1 % of the rows of a 256^3 box).
// Built with a device hash set keyed by a 64-bit hash of the offsets; every row is VERIFIED against the
// table entry it is given, so a hash collision only costs that row the escape.
#define PAT_SLOTS 65536
#define PAT_EMPTY 0xFFFFFFFFFFFFFFFFull
#define PAT_ESCAPE 255
#define PAT_TABLE 254

__device__ __forceinline__ unsigned long long pat_hash(const int* d, int w) {
  unsigned long long h = 0x9E3779B97F4A7C15ull ^ (unsigned long long)w;
  for (int k = 0; k < w; ++k) {
    h ^= (unsigned long long)(unsigned)d[k] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xFF51AFD7ED558CCDull;
    h ^= h >> 33;
  }
  return h == PAT_EMPTY ? 0ull : h;
}

__device__ __forceinline__ int pat_row_deltas(int row, const int* __restrict__ sell_ptr, const int* __restrict__ cols, int* d) {
  const int s = row >> 5, base = sell_ptr[s];
  const int w = (sell_ptr[s + 1] - base) >> 5;
  if (w > MIMPY_SELL_PAT_W) return -1;
  for (int k = 0; k < w; ++k) d[k] = cols[base + 32 * k + (row & 31)] - row;
  return w;
}

// returns the slot of hash h, inserting it when `insert`; -1 when absent / the set is full
__device__ __forceinline__ int pat_find(unsigned long long* keys, unsigned long long h, bool insert, bool* inserted) {
  unsigned slot = (unsigned)(h % PAT_SLOTS);
  for (int probe = 0; probe < 64; ++probe, slot = (slot + 1) % PAT_SLOTS) {
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(keys + slot);
    if (cur == PAT_EMPTY) {
      if (!insert) return -1;
      cur = atomicCAS(keys + slot, PAT_EMPTY, h);
      if (cur == PAT_EMPTY) {
        *inserted = true;
        return (int)slot;
      }
    }
    if (cur == h) return (int)slot;
  }
  return -1;
}

// candidates: the patterns of a SAMPLE of the rows (every stride-th).  A frequent pattern is in any sample;
// the rows next to non-DOF faces, whose offsets depend on their position (the whole bottom cell layer of
// a box with a no-flow floor: 200 000 different vectors at 256^3), would otherwise fill the set first.
__global__ void k_pat_insert(int n_pad, int stride, const int* __restrict__ sell_ptr, const int* __restrict__ cols,
                             unsigned long long* __restrict__ keys, int* __restrict__ rep) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * stride;
  if (row >= n_pad) return;
  int d[MIMPY_SELL_PAT_W];
  const int w = pat_row_deltas((int)row, sell_ptr, cols, d);
  if (w < 0) return;
  bool inserted = false;
  const int slot = pat_find(keys, pat_hash(d, w), true, &inserted);
  if (slot >= 0 && inserted) rep[slot] = (int)row;  // the inserting row is the pattern's representative
}

// how many rows have each candidate pattern
__global__ void k_pat_count(int n_pad, const int* __restrict__ sell_ptr, const int* __restrict__ cols,
                            unsigned long long* __restrict__ keys, unsigned int* __restrict__ count) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_pad) return;
  int d[MIMPY_SELL_PAT_W];
  const int w = pat_row_deltas(row, sell_ptr, cols, d);
  if (w < 0) return;
  bool inserted = false;
  const int slot = pat_find(keys, pat_hash(d, w), false, &inserted);
  if (slot < 0) return;
  // one count per warp and slot: the lanes of a warp mostly share three or four patterns
  const unsigned peers = __match_any_sync(__activemask(), slot);
  if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(count + slot, (unsigned)__popc(peers));
}

// the PAT_TABLE most frequent patterns get ids 0.. in order of frequency (one block)
__global__ void __launch_bounds__(1024)
k_pat_select(const int* __restrict__ sell_ptr, const int* __restrict__ cols, const int* __restrict__ rep,
             unsigned int* __restrict__ count, int* __restrict__ slot_id, int* __restrict__ table, int* __restrict__ npat) {
  __shared__ unsigned long long best[32];
  __shared__ unsigned long long winner;
  const int t = threadIdx.x;
  for (int s = t; s < PAT_SLOTS; s += 1024) slot_id[s] = -1;
  __syncthreads();
  int n = 0;
  for (; n < PAT_TABLE; ++n) {
    unsigned long long mine = 0;  // (count << 32) | (~slot): ties go to the lower slot
    for (int s = t; s < PAT_SLOTS; s += 1024) {
      const unsigned c = count[s];
      if (c) {
        const unsigned long long key = ((unsigned long long)c << 32) | (unsigned)(PAT_SLOTS - 1 - s);
        mine = key > mine ? key : mine;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, mine, o);
      mine = other > mine ? other : mine;
    }
    if ((t & 31) == 0) best[t >> 5] = mine;
    __syncthreads();
    if (t < 32) {
      unsigned long long v = best[t];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other > v ? other : v;
      }
      if (t == 0) winner = v;
    }
    __syncthreads();
    if (winner == 0) break;
    const int slot = PAT_SLOTS - 1 - (int)(unsigned)(winner & 0xffffffffull);
    if (t == 0) {
      slot_id[slot] = n;
      count[slot] = 0;
      int d[MIMPY_SELL_PAT_W];
      const int w = pat_row_deltas(rep[slot], sell_ptr, cols, d);
      for (int k = 0; k < MIMPY_SELL_PAT_W; ++k) table[n * MIMPY_SELL_PAT_W + k] = (k < w) ? d[k] : 0;
    }
    __syncthreads();
  }
  if (t == 0) *npat = n;
}

__global__ void k_pat_assign(int n_pad, const int* __restrict__ sell_ptr, const int* __restrict__ cols,
                             unsigned long long* __restrict__ keys, const int* __restrict__ slot_id,
                             const int* __restrict__ table, uint8_t* __restrict__ pat_id, unsigned int* __restrict__ escaped) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_pad) return;
  int d[MIMPY_SELL_PAT_W];
  const int w = pat_row_deltas(row, sell_ptr, cols, d);
  int id = -1;
  if (w >= 0) {
    bool dummy = false;
    const int slot = pat_find(keys, pat_hash(d, w), false, &dummy);
    if (slot >= 0) id = slot_id[slot];
    for (int k = 0; id >= 0 && k < w; ++k)
      if (table[id * MIMPY_SELL_PAT_W + k] != d[k]) id = -1;   // verified: a hash collision only escapes
  }
  pat_id[row] = (uint8_t)(id >= 0 ? id : PAT_ESCAPE);
  if (id < 0) atomicAdd(escaped, 1u);
}

extern "C" {

// pat_id: 32 * number of slices bytes; pat_table: 255 * MIMPY_SELL_PAT_W ints.  *count_host = number of
// table patterns, 0 when patterns do not pay (more than a quarter of the rows would escape).  Synchronises.
int mimpy_b200_sell_patterns(mimpy_ctx* ctx, int32_t n, const int32_t* sell_ptr, const int32_t* sell_cols,
                             uint8_t* pat_id, int32_t* pat_table, int32_t* count_host) {
  REQUIRE(ctx && sell_ptr && sell_cols && pat_id && pat_table && count_host, "NULL argument");
  *count_host = 0;
  if (n <= 0) return MIMPY_OK;
  const int n_pad = (n + 31) / 32 * 32;
  cudaStream_t st = ctx->stream;
  void* scratch = nullptr;
  const size_t bytes = sizeof(unsigned long long) * PAT_SLOTS + sizeof(int) * (3 * PAT_SLOTS + 2);
  int rc = mimpy_scratch(ctx, bytes, &scratch);
  if (rc) return rc;
  unsigned long long* keys = (unsigned long long*)scratch;
  int* rep = (int*)(keys + PAT_SLOTS);
  int* slot_id = rep + PAT_SLOTS;
  unsigned int* count = (unsigned int*)(slot_id + PAT_SLOTS);
  int* npat = (int*)(count + PAT_SLOTS);
  unsigned int* escaped = (unsigned int*)(npat + 1);
  CUDA_TRY(cudaMemsetAsync(keys, 0xFF, sizeof(unsigned long long) * PAT_SLOTS, st));
  CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int) * (PAT_SLOTS + 2), st));
  const int T = 256, blocks = (n_pad + T - 1) / T;
  const int stride = n_pad > 100000 ? n_pad / 100000 : 1;
  const int sampled = (n_pad + stride - 1) / stride;
  k_pat_insert<<<(sampled + T - 1) / T, T, 0, st>>>(n_pad, stride, sell_ptr, sell_cols, keys, rep);
  k_pat_count<<<blocks, T, 0, st>>>(n_pad, sell_ptr, sell_cols, keys, count);
  k_pat_select<<<1, 1024, 0, st>>>(sell_ptr, sell_cols, rep, count, slot_id, pat_table, npat);
  k_pat_assign<<<blocks, T, 0, st>>>(n_pad, sell_ptr, sell_cols, keys, slot_id, pat_table, pat_id, escaped);
  KERNEL_CHECK();
  int host[2] = {0, 0};
  CUDA_TRY(cudaMemcpyAsync(host, npat, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (getenv("MIMPY_B200_DEBUG")) fprintf(stderr, "sell_patterns: n=%d table patterns=%d escaped rows=%u\n", n, host[0], (unsigned)host[1]);
  *count_host = ((unsigned)host[1] <= (unsigned)n_pad / 4) ? host[0] : 0;
  return MIMPY_OK;
}

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
