// Device scalars of the PCG solvers and the ordered two-stage reduction they share (pcg.cu, auxmg.cu).
#pragma once
#include "common.cuh"

// S_RDR = r . (w_J D^-1 r) and S_BE = b~ . e~ (cell space): the two parts of r . z of the auxiliary-space preconditioner,
// known before z itself is formed (pcg.cu, fused update of p)
enum { S_RZ = 0, S_PAP = 1, S_RZNEW = 2, S_RR = 3, S_RDR = 4, S_BNORM = 5, S_BE = 6, S_TMP = 8 };

#define RED_THREADS 256

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  double t = 0.;
  if (w == 0) {
    t = (l < (blockDim.x >> 5)) ? sh[l] : 0.;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  return t;  // valid in warp 0
}

// last block to arrive sums the per-block partials (NV values each) in index order
template <int NV>
__device__ __forceinline__ void finish_reduction(const double (&v)[NV], double* partials,
                                                 unsigned int* counter, double* dst0, double* sh) {
  __shared__ bool is_last;
  double tot[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) tot[k] = block_sum(v[k], sh);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) partials[(size_t)k * gridDim.x + blockIdx.x] = tot[k];
    __threadfence();
    const unsigned int t = atomicAdd(counter, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double a = 0.;
      for (int b = threadIdx.x; b < gridDim.x; b += blockDim.x)
        a += __ldcg(partials + (size_t)k * gridDim.x + b);
      a = block_sum(a, sh);
      if (threadIdx.x == 0) dst0[k] = a;
    }
    if (threadIdx.x == 0) *counter = 0u;
  }
}

