// Jacobi-preconditioned conjugate gradients on the SPD face system (replaces the SuperLU call of
// MFD.solve, mfd.py:1349-1358; the reference's own Schur-complement CG, mfd.py:1370-1428 and
// twophase.py:31-86, is the design precedent).
//
// One iteration = three kernels, all scalars stay on the device:
//   K1  Ap = A p                   + partial sums of p.Ap      -> S_PAP
//   K2  x += a p ; r -= a Ap       + partial sums of r.D^-1 r, r.r -> S_RZNEW, S_RR   (a = S_RZ/S_PAP)
//   K3  p = D^-1 r + b p           (b = S_RZNEW/S_RZ) ; then S_RZ <- S_RZNEW
// Reductions are two-stage and ordered (per-block partials, last block sums them in index order),
// so a solve is bit-reproducible for a fixed grid.  With more than one rank the three scalars are
// all-reduced over NCCL between the kernels.  check_every iterations are replayed as one CUDA
// graph; the host reads ||r||^2 only at those points.
//
// Algorithmic bytes per iteration (n rows, nnz non-zeros): SpMV 12 nnz + 4(n+1) + 16 n,
// K2 48 n + 8 n (D^-1), K3 8 n (r) + 8 n (D^-1) + 16 n (p)  ->  12 nnz + 108 n  (SURVEY 8d).
#include "common.cuh"


#include <stdio.h>
#include <stdlib.h>

#include "reduce.cuh"

// ---- SpMV: LPR lanes per row (rows of the face system have ~11 entries) --------------------
template <int LPR, bool DOT>
__global__ void __launch_bounds__(RED_THREADS)
k_spmv(int n, const int* __restrict__ indptr, const int* __restrict__ indices,
       const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y,
       double* partials, unsigned int* counter, double* scal) {
  __shared__ double sh[32];
  const int lane = threadIdx.x % LPR;
  const int rows_per_block = RED_THREADS / LPR;
  double acc_dot = 0.;
  for (long long row0 = (long long)blockIdx.x * rows_per_block; row0 < n;
       row0 += (long long)gridDim.x * rows_per_block) {
    const int row = (int)row0 + threadIdx.x / LPR;
    double s = 0.;
    if (row < n) {
      const int b = __ldg(indptr + row), e = __ldg(indptr + row + 1);
      for (int k = b + lane; k < e; k += LPR) s += __ldg(vals + k) * __ldg(x + __ldg(indices + k));
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (row < n && lane == 0) {
      y[row] = s;
      if (DOT) acc_dot += s * __ldg(x + row);
    }
  }
  if (DOT) {
    const double v[1] = {acc_dot};
    finish_reduction<1>(v, partials, counter, scal + S_PAP, sh);
  }
}

__global__ void __launch_bounds__(RED_THREADS)
k_update_xr(int n, int n_owned, const double* __restrict__ p, const double* __restrict__ ap,
            const double* __restrict__ dinv, double* __restrict__ x, double* __restrict__ r,
            double* partials, unsigned int* counter, double* scal) {
  __shared__ double sh[32];
  const double pap = scal[S_PAP];
  const double alpha = (pap != 0.) ? scal[S_RZ] / pap : 0.;
  double rz = 0., rr = 0.;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    x[i] += alpha * p[i];
    const double ri = r[i] - alpha * ap[i];
    r[i] = ri;
    if (i < n_owned) {  // ghost copies of interface rows are counted by their owner rank
      rz += ri * (__ldg(dinv + i) * ri);
      rr += ri * ri;
    }
  }
  const double v[2] = {rz, rr};
  finish_reduction<2>(v, partials, counter, scal + S_RZNEW, sh);
}

__global__ void __launch_bounds__(RED_THREADS)
k_update_p(int n, const double* __restrict__ r, const double* __restrict__ dinv,
           double* __restrict__ p, unsigned int* counter, double* scal) {
  const double rz = scal[S_RZ];
  const double beta = (rz != 0.) ? scal[S_RZNEW] / rz : 0.;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    p[i] = __ldg(dinv + i) * r[i] + beta * p[i];
  // the last block to finish rotates the scalar (every block has read S_RZ by then)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(counter, 1u);
    if (t == gridDim.x - 1) {
      scal[S_RZ] = scal[S_RZNEW];
      *counter = 0u;
    }
  }
}

// ---- variants for a preconditioner that is not a diagonal (auxmg.cu): z is a vector of its own ----
// x += a p ; r -= a Ap ; S_RR = r.r
__global__ void __launch_bounds__(RED_THREADS)
k_update_xr_gen(int n, int n_owned, const double* __restrict__ p, const double* __restrict__ ap,
                double* __restrict__ x, double* __restrict__ r, double* partials, unsigned int* counter,
                double* scal) {
  __shared__ double sh[32];
  const double pap = scal[S_PAP];
  const double alpha = (pap != 0.) ? scal[S_RZ] / pap : 0.;
  double rr = 0.;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    x[i] += alpha * p[i];
    const double ri = r[i] - alpha * ap[i];
    r[i] = ri;
    if (i < n_owned) rr += ri * ri;
  }
  const double v[1] = {rr};
  finish_reduction<1>(v, partials, counter, scal + S_RR, sh);
}

// the same for the fused update of p (mimpy_auxmg_apply_p): also S_RDR = r . (w_J D^-1 r) with the FP32 copy of the
// Jacobi diagonal that the prolongation will use
__global__ void __launch_bounds__(RED_THREADS)
k_update_xr_aux(int n, const double* __restrict__ p, const double* __restrict__ ap, const float* __restrict__ dinv32,
                double wj, double* __restrict__ x, double* __restrict__ r, double* partials, unsigned int* counter,
                double* scal) {
  __shared__ double sh[32];
  const double pap = scal[S_PAP];
  const double alpha = (pap != 0.) ? scal[S_RZ] / pap : 0.;
  double rr = 0., rdr = 0.;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    x[i] += alpha * p[i];
    const double ri = r[i] - alpha * ap[i];
    r[i] = ri;
    rr += ri * ri;
    rdr += ri * (wj * (double)__ldg(dinv32 + i) * ri);
  }
  const double v[2] = {rr, rdr};
  finish_reduction<2>(v, partials, counter, scal + S_RR, sh);  // S_RR, S_RDR are neighbours
}

// S_RDR of the initial residual
__global__ void __launch_bounds__(RED_THREADS)
k_rdr(int n, const double* __restrict__ r, const float* __restrict__ dinv32, double wj, double* partials,
      unsigned int* counter, double* scal) {
  __shared__ double sh[32];
  double rdr = 0.;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double ri = r[i];
    rdr += ri * (wj * (double)__ldg(dinv32 + i) * ri);
  }
  const double v[1] = {rdr};
  finish_reduction<1>(v, partials, counter, scal + S_RDR, sh);
}

// p = z + b p (b = S_RZNEW / S_RZ, or 0 when `first`), then S_RZ <- S_RZNEW
__global__ void __launch_bounds__(RED_THREADS)
k_update_p_gen(int n, const double* __restrict__ z, double* __restrict__ p, unsigned int* counter,
               double* scal, int first) {
  const double rz = scal[S_RZ];
  const double beta = (first || rz == 0.) ? 0. : scal[S_RZNEW] / rz;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    p[i] = z[i] + beta * p[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(counter, 1u);
    if (t == gridDim.x - 1) {
      scal[S_RZ] = scal[S_RZNEW];
      *counter = 0u;
    }
  }
}

// r = b - A x (given Ax in ap), p = D^-1 r, S_RZ = r.D^-1 r, S_RR = r.r, S_BNORM = b.b
__global__ void __launch_bounds__(RED_THREADS)
k_init(int n, int n_owned, const double* __restrict__ b, const double* __restrict__ ax,
       const double* __restrict__ dinv, double* __restrict__ r, double* __restrict__ p,
       double* partials, unsigned int* counter, double* scal) {
  __shared__ double sh[32];
  double rz = 0., rr = 0., bb = 0.;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const double bi = b[i];
    const double ri = bi - ax[i];
    const double zi = __ldg(dinv + i) * ri;
    r[i] = ri;
    p[i] = zi;
    if (i < n_owned) {
      rz += ri * zi;
      rr += ri * ri;
      bb += bi * bi;
    }
  }
  const double v[3] = {rz, rr, bb};
  finish_reduction<3>(v, partials, counter, scal + S_TMP, sh);
}

__global__ void k_init_scalars(double* scal) {
  scal[S_RZ] = scal[S_TMP];
  scal[S_RR] = scal[S_TMP + 1];
  scal[S_BNORM] = scal[S_TMP + 2];
  scal[S_RZNEW] = scal[S_TMP];
}

// ---- SpMV, SELL-32 layout (sell.cu): lane r of a warp owns row r of a 32-row slice; every load of
// the warp is one contiguous line (256 B of values, 128 B of columns), no cross-lane reduction ----
template <bool DOT>
__global__ void __launch_bounds__(RED_THREADS)
k_spmv_sell(int n, const int* __restrict__ sell_ptr, const int* __restrict__ cols,
            const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y,
            double* partials, unsigned int* counter, double* scal) {
  __shared__ double sh[32];
  const long long n_pad = ((long long)n + 31) & ~31LL;
  double acc_dot = 0.;
  for (long long row = (long long)blockIdx.x * RED_THREADS + threadIdx.x; row < n_pad;
       row += (long long)gridDim.x * RED_THREADS) {
    const int s = (int)(row >> 5);
    const int base = __ldg(sell_ptr + s);
    const int w = (__ldg(sell_ptr + s + 1) - base) >> 5;
    const double* v = vals + base + (row & 31);
    const int* c = cols + base + (row & 31);
    double a0 = 0., a1 = 0.;
    int k = 0;
    for (; k + 3 < w; k += 4) {  // 4 independent value/column/x chains in flight
      const int c0 = __ldg(c + 32 * k), c1 = __ldg(c + 32 * (k + 1));
      const int c2 = __ldg(c + 32 * (k + 2)), c3 = __ldg(c + 32 * (k + 3));
      const double v0 = __ldg(v + 32 * k), v1 = __ldg(v + 32 * (k + 1));
      const double v2 = __ldg(v + 32 * (k + 2)), v3 = __ldg(v + 32 * (k + 3));
      a0 += v0 * __ldg(x + c0);
      a1 += v1 * __ldg(x + c1);
      a0 += v2 * __ldg(x + c2);
      a1 += v3 * __ldg(x + c3);
    }
    for (; k < w; ++k) a0 += __ldg(v + 32 * k) * __ldg(x + __ldg(c + 32 * k));
    const double acc = a0 + a1;
    if (row < n) {
      y[row] = acc;
      if (DOT) acc_dot += acc * __ldg(x + row);
    }
  }
  if (DOT) {
    const double vv[1] = {acc_dot};
    finish_reduction<1>(vv, partials, counter, scal + S_PAP, sh);
  }
}

// SELL-32 with column patterns (sell.cu): column = row + table[pattern of the row][k]; the table sits in
// shared memory with an odd pitch (rows of a warp mix a few patterns: no bank conflict between them)
#define PAT_PITCH (MIMPY_SELL_PAT_W + 1)
template <bool DOT>
__global__ void __launch_bounds__(RED_THREADS)
k_spmv_sell_pat(int n, const int* __restrict__ sell_ptr, const uint8_t* __restrict__ pat_id,
                const int* __restrict__ pat_table, int npat, const int* __restrict__ cols,
                const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y,
                double* partials, unsigned int* counter, double* scal) {
  __shared__ double sh[32];
  __shared__ int tab[255 * PAT_PITCH];
  for (int e = threadIdx.x; e < npat * MIMPY_SELL_PAT_W; e += RED_THREADS)
    tab[(e / MIMPY_SELL_PAT_W) * PAT_PITCH + e % MIMPY_SELL_PAT_W] = pat_table[e];
  __syncthreads();
  const long long n_pad = ((long long)n + 31) & ~31LL;
  double acc_dot = 0.;
  for (long long row = (long long)blockIdx.x * RED_THREADS + threadIdx.x; row < n_pad;
       row += (long long)gridDim.x * RED_THREADS) {
    const int s = (int)(row >> 5);
    const int base = __ldg(sell_ptr + s);
    const int w = (__ldg(sell_ptr + s + 1) - base) >> 5;
    const double* v = vals + base + (row & 31);
    const int pid = (int)__ldg(pat_id + row);
    double a0 = 0., a1 = 0.;
    int k = 0;
    if (pid != 255) {
      const int* off = tab + pid * PAT_PITCH;
      const double* xr = x + row;
      for (; k + 3 < w; k += 4) {  // 4 independent value / x chains in flight
        const double v0 = __ldg(v + 32 * k), v1 = __ldg(v + 32 * (k + 1));
        const double v2 = __ldg(v + 32 * (k + 2)), v3 = __ldg(v + 32 * (k + 3));
        a0 += v0 * __ldg(xr + off[k]);
        a1 += v1 * __ldg(xr + off[k + 1]);
        a0 += v2 * __ldg(xr + off[k + 2]);
        a1 += v3 * __ldg(xr + off[k + 3]);
      }
      for (; k < w; ++k) a0 += __ldg(v + 32 * k) * __ldg(xr + off[k]);
    } else {  // escape: a row whose offsets are not in the table reads its columns (same summation order)
      const int* c = cols + base + (row & 31);
      for (; k + 3 < w; k += 4) {
        const double v0 = __ldg(v + 32 * k), v1 = __ldg(v + 32 * (k + 1));
        const double v2 = __ldg(v + 32 * (k + 2)), v3 = __ldg(v + 32 * (k + 3));
        a0 += v0 * __ldg(x + __ldg(c + 32 * k));
        a1 += v1 * __ldg(x + __ldg(c + 32 * (k + 1)));
        a0 += v2 * __ldg(x + __ldg(c + 32 * (k + 2)));
        a1 += v3 * __ldg(x + __ldg(c + 32 * (k + 3)));
      }
      for (; k < w; ++k) a0 += __ldg(v + 32 * k) * __ldg(x + __ldg(c + 32 * k));
    }
    const double acc = a0 + a1;
    if (row < n) {
      y[row] = acc;
      if (DOT) acc_dot += acc * __ldg(x + row);
    }
  }
  if (DOT) {
    const double vv[1] = {acc_dot};
    finish_reduction<1>(vv, partials, counter, scal + S_PAP, sh);
  }
}

// the face system in either layout
struct SpmvMat {
  int sell;          // 0: CSR (ptr = indptr), 1: SELL-32 (ptr = sell_ptr)
  const int* ptr;
  const int* idx;
  const double* vals;
  const uint8_t* pat_id;   // SELL only, nullable: column patterns instead of idx
  const int* pat_table;
  int npat;
};

static void launch_spmv(mimpy_ctx* ctx, int grid, int n, const SpmvMat& A, const double* x, double* y,
                        bool dot) {
  cudaStream_t st = ctx->stream;
  if (A.sell && A.pat_id && A.npat > 0) {
    if (dot)
      k_spmv_sell_pat<true><<<grid, RED_THREADS, 0, st>>>(n, A.ptr, A.pat_id, A.pat_table, A.npat, A.idx, A.vals, x, y,
                                                          ctx->partials, ctx->counters + 0, ctx->scalars);
    else
      k_spmv_sell_pat<false><<<grid, RED_THREADS, 0, st>>>(n, A.ptr, A.pat_id, A.pat_table, A.npat, A.idx, A.vals, x, y,
                                                           nullptr, nullptr, nullptr);
  } else if (A.sell) {
    if (dot)
      k_spmv_sell<true><<<grid, RED_THREADS, 0, st>>>(n, A.ptr, A.idx, A.vals, x, y, ctx->partials,
                                                      ctx->counters + 0, ctx->scalars);
    else
      k_spmv_sell<false><<<grid, RED_THREADS, 0, st>>>(n, A.ptr, A.idx, A.vals, x, y, nullptr, nullptr, nullptr);
  } else {
    if (dot)
      k_spmv<8, true><<<grid, RED_THREADS, 0, st>>>(n, A.ptr, A.idx, A.vals, x, y, ctx->partials,
                                                    ctx->counters + 0, ctx->scalars);
    else
      k_spmv<8, false><<<grid, RED_THREADS, 0, st>>>(n, A.ptr, A.idx, A.vals, x, y, nullptr, nullptr, nullptr);
  }
}

static int allreduce(mimpy_ctx* ctx, double* buf, int count) { return mimpy_allreduce_f64(ctx, buf, count); }

// z = B r with the auxiliary-space preconditioner, completed on interface rows, r.z all-reduced
static int apply_aux(mimpy_ctx* ctx, mimpy_auxmg* aux, int n, const double* dinv, const double* r, double* z,
                     int reduce_count = 2) {
  const int n_owned = (ctx->world > 1 && ctx->halo.n_owned > 0) ? ctx->halo.n_owned : n;
  int rc = mimpy_auxmg_apply(ctx, aux, dinv, r, z, ctx->scalars + S_RZNEW, n_owned);
  if (rc) return rc;
  rc = allreduce(ctx, ctx->scalars + S_RZNEW, reduce_count);  // r.z (and r.r of k_update_xr_gen)
  if (rc) return rc;
  return mimpy_halo_exchange_add(ctx, z);
}

static int launch_iteration(mimpy_ctx* ctx, int grid, int n, const SpmvMat& A, const double* dinv,
                            double* x, double* r, double* p, double* ap, mimpy_auxmg* aux = nullptr,
                            double* z = nullptr, const mimpy_aux_fused* fz = nullptr) {
  cudaStream_t st = ctx->stream;
  if (aux && fz) {  // single rank: three kernels around the V-cycle, z never stored
    launch_spmv(ctx, grid, n, A, p, ap, true);
    k_update_xr_aux<<<grid, RED_THREADS, 0, st>>>(n, p, ap, fz->dinv32, fz->wj, x, r, ctx->partials, ctx->counters + 1,
                                                  ctx->scalars);
    return mimpy_auxmg_apply_p(ctx, aux, r, p, 0);
  }
  if (aux) {
    launch_spmv(ctx, grid, n, A, p, ap, true);
    int rc = allreduce(ctx, ctx->scalars + S_PAP, 1);
    if (rc) return rc;
    rc = mimpy_halo_exchange_add(ctx, ap);
    if (rc) return rc;
    const int n_owned = (ctx->world > 1 && ctx->halo.n_owned > 0) ? ctx->halo.n_owned : n;
    k_update_xr_gen<<<grid, RED_THREADS, 0, st>>>(n, n_owned, p, ap, x, r, ctx->partials, ctx->counters + 1,
                                                  ctx->scalars);
    rc = apply_aux(ctx, aux, n, dinv, r, z);
    if (rc) return rc;
    k_update_p_gen<<<grid, RED_THREADS, 0, st>>>(n, z, p, ctx->counters + 2, ctx->scalars, 0);
    return MIMPY_OK;
  }
  launch_spmv(ctx, grid, n, A, p, ap, true);
  // p.Ap = sum over ranks of p_loc.(A_loc p_loc): the fused dot over ALL local rows of the
  // still-partial interface values is already this rank's exact share (sub-assembled operator)
  int rc = allreduce(ctx, ctx->scalars + S_PAP, 1);
  if (rc) return rc;
  rc = mimpy_halo_exchange_add(ctx, ap);  // interface rows of Ap: own + neighbour
  if (rc) return rc;
  const int n_owned = (ctx->world > 1 && ctx->halo.n_owned > 0) ? ctx->halo.n_owned : n;
  k_update_xr<<<grid, RED_THREADS, 0, st>>>(n, n_owned, p, ap, dinv, x, r, ctx->partials, ctx->counters + 1,
                                            ctx->scalars);
  rc = allreduce(ctx, ctx->scalars + S_RZNEW, 2);
  if (rc) return rc;
  k_update_p<<<grid, RED_THREADS, 0, st>>>(n, r, dinv, p, ctx->counters + 2, ctx->scalars);
  return MIMPY_OK;
}

extern "C" {

int mimpy_b200_spmv(mimpy_ctx* ctx, int32_t n, const int32_t* indptr, const int32_t* indices,
                    const double* vals, const double* x, double* y) {
  REQUIRE(ctx && indptr && indices && vals && x && y, "NULL argument");
  if (n <= 0) return MIMPY_OK;
  const int rpb = RED_THREADS / 8;
  long long want = ((long long)n + rpb - 1) / rpb;
  const int grid = (int)(want < (long long)ctx->sm_count * 8 ? want : (long long)ctx->sm_count * 8);
  k_spmv<8, false><<<grid, RED_THREADS, 0, ctx->stream>>>(n, indptr, indices, vals, x, y, nullptr,
                                                          nullptr, nullptr);
  KERNEL_CHECK();
  return MIMPY_OK;
}

int mimpy_b200_spmv_sell(mimpy_ctx* ctx, int32_t n, const int32_t* sell_ptr, const int32_t* sell_cols,
                         const double* sell_vals, const double* x, double* y) {
  REQUIRE(ctx && sell_ptr && sell_cols && sell_vals && x && y, "NULL argument");
  if (n <= 0) return MIMPY_OK;
  long long want = ((long long)n + RED_THREADS - 1) / RED_THREADS;
  const int grid = (int)(want < (long long)ctx->sm_count * 8 ? want : (long long)ctx->sm_count * 8);
  SpmvMat A = {1, sell_ptr, sell_cols, sell_vals, nullptr, nullptr, 0};
  launch_spmv(ctx, grid, n, A, x, y, false);
  KERNEL_CHECK();
  return MIMPY_OK;
}

int mimpy_b200_spmv_sell_pat(mimpy_ctx* ctx, int32_t n, const int32_t* sell_ptr, const int32_t* sell_cols,
                             const uint8_t* pat_id, const int32_t* pat_table, int32_t npat, const double* sell_vals,
                             const double* x, double* y) {
  REQUIRE(ctx && sell_ptr && sell_cols && pat_id && pat_table && sell_vals && x && y, "NULL argument");
  REQUIRE(npat > 0 && npat <= 254, "pattern count out of range");
  if (n <= 0) return MIMPY_OK;
  long long want = ((long long)n + RED_THREADS - 1) / RED_THREADS;
  const int grid = (int)(want < (long long)ctx->sm_count * 8 ? want : (long long)ctx->sm_count * 8);
  SpmvMat A = {1, sell_ptr, sell_cols, sell_vals, pat_id, pat_table, npat};
  launch_spmv(ctx, grid, n, A, x, y, false);
  KERNEL_CHECK();
  return MIMPY_OK;
}

}  // extern "C"

static int pcg_core(mimpy_ctx* ctx, int32_t n, const SpmvMat& A, const double* diag_inv,
                    const double* b, double* x, double* work, mimpy_pcg_info* info,
                    mimpy_auxmg* aux = nullptr) {
  REQUIRE(ctx && A.ptr && A.idx && A.vals && diag_inv && b && x && work && info, "NULL argument");
  REQUIRE(n > 0, "empty system");
  REQUIRE(info->maxit > 0, "maxit must be positive");
  cudaStream_t st = ctx->stream;
  double* r = work;
  double* p = work + (size_t)n;
  double* ap = work + 2 * (size_t)n;
  double* z = work + 3 * (size_t)n;  // used with `aux` only
  const int grid = ctx->sm_count * 8;  // persistent-style: 8 resident CTAs of 256 threads per SM
  const int check = info->check_every > 0 ? info->check_every : 25;

  CUDA_TRY(cudaEventRecord(ctx->ev0, st));
  // r0 = b - A x0
  launch_spmv(ctx, grid, n, A, x, ap, false);
  int rc = mimpy_halo_exchange_add(ctx, ap);
  if (rc) return rc;
  const int n_owned = (ctx->world > 1 && ctx->halo.n_owned > 0) ? ctx->halo.n_owned : n;
  k_init<<<grid, RED_THREADS, 0, st>>>(n, n_owned, b, ap, diag_inv, r, p, ctx->partials, ctx->counters + 3,
                                       ctx->scalars);
  rc = allreduce(ctx, ctx->scalars + S_TMP, 3);
  if (rc) return rc;
  k_init_scalars<<<1, 1, 0, st>>>(ctx->scalars);
  KERNEL_CHECK();
  mimpy_aux_fused fz_store;
  const mimpy_aux_fused* fz = (aux && mimpy_auxmg_fused(ctx, aux, diag_inv, n, &fz_store)) ? &fz_store : nullptr;
  if (aux && fz) {  // p0 = B r0 with the fused update
    k_rdr<<<grid, RED_THREADS, 0, st>>>(n, r, fz->dinv32, fz->wj, ctx->partials, ctx->counters + 3, ctx->scalars);
    rc = mimpy_auxmg_apply_p(ctx, aux, r, p, 1);
    if (rc) return rc;
  } else if (aux) {  // replace the Jacobi start by z0 = B r0, p0 = z0
    rc = apply_aux(ctx, aux, n, diag_inv, r, z, 1);
    if (rc) return rc;
    k_update_p_gen<<<grid, RED_THREADS, 0, st>>>(n, z, p, ctx->counters + 2, ctx->scalars, 1);
    KERNEL_CHECK();
  }
  CUDA_TRY(cudaMemcpyAsync(ctx->host_scalars, ctx->scalars, sizeof(double) * 8,
                           cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  const double bnorm = sqrt(ctx->host_scalars[S_BNORM]);
  double rnorm = sqrt(ctx->host_scalars[S_RR]);
  const double rz0 = ctx->host_scalars[S_RZ];
  double rz = rz0;
  const bool energy = info->norm_kind == 1;
  const double target = fmax(info->rtol * (info->ref_norm > 0. ? info->ref_norm : bnorm), info->atol);
  const double target_rz = info->rtol * info->rtol * rz0;
  info->rhs_norm = bnorm;
  info->rz0 = rz0;
  int it = 0;
  // true while the stopping test is not met
  auto running = [&]() { return energy ? (rz > target_rz && rnorm > info->atol) : (rnorm > target); };

  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  bool use_graph = info->use_graph && st != nullptr && st != cudaStreamLegacy &&
                   st != cudaStreamPerThread;
  if (use_graph && running()) {
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      int crc = MIMPY_OK;
      for (int k = 0; k < check && crc == MIMPY_OK; ++k)
        crc = launch_iteration(ctx, grid, n, A, diag_inv, x, r, p, ap, aux, z, fz);
      cudaError_t e = cudaStreamEndCapture(st, &graph);
      if (crc != MIMPY_OK || e != cudaSuccess || graph == nullptr ||
          cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        graph = nullptr;
        gexec = nullptr;
        use_graph = false;
        cudaGetLastError();
      }
    } else {
      use_graph = false;
      cudaGetLastError();
    }
  }

  // Between two checks `check` iterations run as one graph.  Near the end that granularity is wasted work (up to
  // check - 1 iterations of ~3 ms each at 256^3 against ~30 us for a check): from the contraction observed over the last
  // interval the number of iterations still needed is estimated, and when it is below `check` exactly that many
  // single-iteration graphs are launched before the next check.  Deterministic (a function of the all-reduced residuals).
  cudaGraph_t graph1 = nullptr;
  cudaGraphExec_t gexec1 = nullptr;
  bool single_ok = use_graph && gexec != nullptr && check > 1 && !getenv("MIMPY_B200_PCG_FIXED_CHECKS");
  const bool trace = getenv("MIMPY_B200_PCG_TRACE") != nullptr;
  double prev_measure = -1.;
  int prev_it = 0;
  auto measure = [&]() { return energy ? sqrt(fabs(rz)) : rnorm; };
  auto goal = [&]() { return energy ? sqrt(target_rz) : target; };
  while (running() && it < info->maxit) {
    int todo = check;
    if (single_ok && prev_measure > 0. && measure() > 0. && measure() < prev_measure && it > prev_it) {
      const double rate = log(measure() / prev_measure) / (double)(it - prev_it);  // < 0
      const double need = ceil(log(goal() / measure()) / rate);
      if (need >= 1. && need < (double)check) todo = (int)need;
    }
    prev_measure = measure();
    prev_it = it;
    if (todo < check && !gexec1) {
      if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        const int crc = launch_iteration(ctx, grid, n, A, diag_inv, x, r, p, ap, aux, z, fz);
        const cudaError_t e = cudaStreamEndCapture(st, &graph1);
        if (crc != MIMPY_OK || e != cudaSuccess || graph1 == nullptr ||
            cudaGraphInstantiate(&gexec1, graph1, 0) != cudaSuccess) {
          if (graph1) cudaGraphDestroy(graph1);
          graph1 = nullptr;
          gexec1 = nullptr;
          cudaGetLastError();
        }
      } else {
        cudaGetLastError();
      }
      if (!gexec1) {
        single_ok = false;
        todo = check;
      }
    }
    if (todo < check && gexec1) {
      for (int k = 0; k < todo; ++k) CUDA_TRY(cudaGraphLaunch(gexec1, st));
    } else if (use_graph && gexec) {
      CUDA_TRY(cudaGraphLaunch(gexec, st));
    } else {
      for (int k = 0; k < check; ++k) {
        rc = launch_iteration(ctx, grid, n, A, diag_inv, x, r, p, ap, aux, z, fz);
        if (rc) return rc;
      }
      KERNEL_CHECK();
    }
    it += todo;
    if (trace) fprintf(stderr, "pcg: it %d (+%d) residual before %.3e target %.3e\n", it, todo, prev_measure, goal());
    CUDA_TRY(cudaMemcpyAsync(ctx->host_scalars, ctx->scalars, sizeof(double) * 8,
                             cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    rnorm = sqrt(ctx->host_scalars[S_RR]);
    rz = ctx->host_scalars[S_RZ];
    if (!(rnorm == rnorm)) break;  // NaN: matrix not SPD
  }
  if (gexec1) cudaGraphExecDestroy(gexec1);
  if (graph1) cudaGraphDestroy(graph1);
  CUDA_TRY(cudaEventRecord(ctx->ev1, st));
  CUDA_TRY(cudaEventSynchronize(ctx->ev1));
  CUDA_TRY(cudaEventElapsedTime(&info->ms_total, ctx->ev0, ctx->ev1));
  if (ctx->world > 1) {  // a peer-memory wait that ran out must not pass as a result
    rc = mimpy_check_device_flag(ctx);
    if (rc) {
      if (gexec) cudaGraphExecDestroy(gexec);
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
  }
  if (gexec) cudaGraphExecDestroy(gexec);
  if (graph) cudaGraphDestroy(graph);
  info->iterations = it;
  info->residual = rnorm;
  info->rz = rz;
  if (!(rnorm == rnorm)) {
    mimpy_set_error("pcg: residual became NaN (face system not SPD?)");
    return MIMPY_E_NOCONV;
  }
  if (running()) {
    mimpy_set_error("pcg: %d iterations, residual %.3e > target %.3e", it, energy ? sqrt(fabs(rz)) : rnorm,
                    energy ? sqrt(target_rz) : target);
    return MIMPY_E_NOCONV;
  }
  return MIMPY_OK;
}

extern "C" {

int mimpy_b200_pcg(mimpy_ctx* ctx, int32_t n, const int32_t* indptr, const int32_t* indices,
                   const double* vals, const double* diag_inv, const double* b, double* x,
                   double* work, mimpy_pcg_info* info) {
  SpmvMat A = {0, indptr, indices, vals, nullptr, nullptr, 0};
  return pcg_core(ctx, n, A, diag_inv, b, x, work, info);
}

int mimpy_b200_pcg_sell(mimpy_ctx* ctx, int32_t n, const int32_t* sell_ptr, const int32_t* sell_cols,
                        const double* sell_vals, const double* diag_inv, const double* b, double* x,
                        double* work, mimpy_pcg_info* info) {
  REQUIRE(info, "NULL argument");
  SpmvMat A = {1, sell_ptr, sell_cols, sell_vals, info->sell_pat_id, info->sell_pat_table, info->sell_npat};
  return pcg_core(ctx, n, A, diag_inv, b, x, work, info);
}

int mimpy_b200_pcg_aux(mimpy_ctx* ctx, int32_t n, const int32_t* ptr, const int32_t* idx,
                       const double* vals, int32_t sell, const double* diag_inv, void* auxmg,
                       const double* b, double* x, double* work, mimpy_pcg_info* info) {
  REQUIRE(auxmg, "NULL preconditioner handle");
  REQUIRE(info, "NULL argument");
  SpmvMat A = {sell ? 1 : 0, ptr, idx, vals, sell ? info->sell_pat_id : nullptr, sell ? info->sell_pat_table : nullptr,
               sell ? info->sell_npat : 0};
  return pcg_core(ctx, n, A, diag_inv, b, x, work, info, static_cast<mimpy_auxmg*>(auxmg));
}

}  // extern "C"
