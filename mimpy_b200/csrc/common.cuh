// Shared helpers for libmimpy_b200 (sm_100a). FP64 CUDA-core math only: the path works on
// 3xn / nxn blocks with n = faces per cell (6 for hexes), far too small for tcgen05 tiles.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "mimpy_b200.h"

// interface layers of a z-slab partition: positions (in the local face-system
// numbering) of the faces shared with the lower / upper neighbour rank
struct mimpy_halo {
  int n_owned;          // rows [0, n_owned) are counted in dot products (the rest are ghost copies)
  int n_lo, n_hi;
  const int* idx_lo;
  const int* idx_hi;
  int rank_lo, rank_hi;
  double* buf;          // 2*(n_lo + n_hi) doubles
};

// NVLink peer-memory state (comm.cu): this rank's symmetric buffer and its peers' mapped copies
struct mimpy_peer {
  int enabled, rank, world;
  char* local;
  char** peers;                 // device array of `world` pointers (peers[rank] == local)
  unsigned long long* epoch;    // device: [0] all-reduce, [1] halo, [2] small all-gather, [3] large all-gather
  size_t halo_cap;              // doubles per halo buffer
  int bigg;                     // 1: the buffer has the large all-gather region behind the halo data
};

struct mimpy_ctx {
  mimpy_halo halo;
  mimpy_peer peer;
  int device;
  cudaStream_t stream;
  int sm_count;
  void* scratch;         // ctx-owned temporary storage (CUB temp, reduction partials)
  size_t scratch_bytes;
  double* diag_swap;     // per-face exchange slots of the staged brick assembly kernel; all-sentinel
  size_t diag_swap_len;  //   between launches (the kernel restores every slot it used)
  // a launch over a RANGE of cell layers leaves the halves of its outermost z-face layers in their slots
  // until the neighbouring range arrives: layer -> 1 (left by the cells below) / 2 (by the cells above)
  int swap_dirty[8][2];
  int swap_dirty_n;
  int swap_L;            // faces per layer of the mesh those entries refer to
  // plan binding of the face records (assemble_brick3.cu): record array and symbolic serial last bound
  const void* bound_geom;
  long long bound_serial;
  unsigned b3_ticket;    // value of the brick ticket counter (counters[12]) before the next launch
  // per-face row bases of the face system last written by mimpy_face_rowbase (plan serial, layout, buffer)
  long long rowbase_serial;
  const void* rowbase_sell;
  const void* rowbase_out;
  int rowbase_nf;
  double* partials;      // reduction partials (device)
  unsigned int* counters;  // last-block counters (device)
  double* scalars;       // device scalars for PCG (alpha, beta, rz, ...)
  double* host_scalars;  // pinned
  void* nccl_comm;
  int rank, world;
  cudaEvent_t ev0, ev1;
};

void mimpy_set_error(const char* fmt, ...);
// uniform-face-count fast path of the numeric assembly (assemble_uniform.cu)
int mimpy_numeric_uniform(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym,
                          int method, int flags, const double* multipliers, double alpha,
                          const double* c_diag, double* vals, int layer_begin = 0, int layer_end = -1);
// assemble_brick3.cu: face-system set-up on the pipelined brick kernel (structured hex meshes)
int mimpy_face_rowbase(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym, const int* sell_ptr, int* out);
int mimpy_hybrid_setup_brick(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym, int method, int flags,
                             const double* multipliers, double alpha, const double* c_diag, const int* face_rowbase,
                             int row_stride, double* w_e, double* tau, double* a_vals, double* a_diag, int invert_diag);
// assemble_uniform.cu: exchange slots of the two-contributor diagonals, clean (all-sentinel) for the launch to come
int mimpy_swap_prepare(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, bool ranged, int layer_begin, int layer_end);
int mimpy_scratch(mimpy_ctx* ctx, size_t bytes, void** out);
int mimpy_check_device_flag(mimpy_ctx* ctx);  // stream sync + error if a bounded device wait ran out
// auxmg.cu: z = B r (auxiliary-space multigrid + Jacobi), this rank's share of r.z -> *rz_dst
struct mimpy_auxmg;
int mimpy_auxmg_apply(mimpy_ctx* ctx, mimpy_auxmg* h, const double* dinv, const double* r, double* z,
                      double* rz_dst, int n_owned);
// single rank, FP32 levels: the preconditioner application fused with the update of the search direction,
// p = B r + beta p, z never stored (r . z = S_RDR + b . e is known before the prolongation).  mimpy_auxmg_fused
// returns 1 when the handle supports it and prepares the FP32 copy of the Jacobi diagonal.
struct mimpy_aux_fused {
  const float* dinv32;
  double wj;
};
int mimpy_auxmg_fused(mimpy_ctx* ctx, mimpy_auxmg* h, const double* dinv, int n, mimpy_aux_fused* out);
int mimpy_auxmg_apply_p(mimpy_ctx* ctx, mimpy_auxmg* h, const double* r, double* p, int first);
// comm.cu
int mimpy_allreduce_f64(mimpy_ctx* ctx, double* buf, int count);
int mimpy_halo_exchange_add(mimpy_ctx* ctx, double* v);
int mimpy_allgather_f64(mimpy_ctx* ctx, const double* mine, int n, double* out);  // out[r*n+i] = mine[i] of rank r
// ghost layers of a z-slab partition: v[send_lo .. +n) -> rank-1 (lands at its recv_hi), v[send_hi .. +n)
// -> rank+1 (lands at its recv_lo); plain copies
struct mimpy_xchg {
  int n;
  size_t send_lo, send_hi, recv_lo, recv_hi;
};
int mimpy_layer_exchange(mimpy_ctx* ctx, double* v, const mimpy_xchg& x);

// ---- the packed face record -----------------------------------------------------------------
// face_geom[f] = four 16-byte chunks  c0 = (area, n_x)  c1 = (n_y, n_z)  c2 = (x_f, y_f)  c3 = (z_f, bind)
// stored in a 2-bit XOR swizzle keyed by the face id: chunk c lives at position c ^ ((f >> 1) & 3).
// A run of consecutive records that a bulk copy (cp.async.bulk) drops into shared memory unchanged can
// then be read with one 128-bit load per lane and chunk without bank conflicts: the records of a cell's
// z-, y-, x- faces are 192 bytes apart (assemble_brick3.cu).  `bind` is the plan binding written by
// mimpy_bind_plan: low word = CSR offset of the face's row in the saddle matrix (-1: Neumann face, no
// row), high word = which faces of the two adjacent cells are flux DOFs (bits 0-5: the cell that lists f
// as one of its x+/y+/z+ faces, bits 8-13: the cell that lists it as x-/y-/z-).
__device__ __host__ __forceinline__ int rec_pos(long long f, int c) { return c ^ (int)((f >> 1) & 3); }
__device__ __forceinline__ const double2* rec_chunk(const double* face_geom, size_t f, int c) {
  return reinterpret_cast<const double2*>(face_geom + 8 * f) + rec_pos((long long)f, c);
}
__device__ __forceinline__ double rec_area(const double* face_geom, size_t f) {
  return __ldg(reinterpret_cast<const double*>(rec_chunk(face_geom, f, 0)));
}
__device__ __forceinline__ void rec_store(double* face_geom, size_t f, double area, const double* n, const double* c) {
  double2* g = reinterpret_cast<double2*>(face_geom + 8 * f);
  g[rec_pos((long long)f, 0)] = make_double2(area, n[0]);
  g[rec_pos((long long)f, 1)] = make_double2(n[1], n[2]);
  g[rec_pos((long long)f, 2)] = make_double2(c[0], c[1]);
  g[rec_pos((long long)f, 3)] = make_double2(c[2], 0.);
}

#define CUDA_TRY(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      mimpy_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                 \
                      cudaGetErrorString(_e));                                           \
      return MIMPY_E_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define KERNEL_CHECK() CUDA_TRY(cudaGetLastError())

#define REQUIRE(cond, msg)                                   \
  do {                                                       \
    if (!(cond)) {                                           \
      mimpy_set_error("%s: %s", __func__, msg);              \
      return MIMPY_E_INVALID;                                \
    }                                                        \
  } while (0)

static inline int mimpy_group_size(int max_faces) {
  if (max_faces <= 8) return 8;
  if (max_faces <= 16) return 16;
  return 32;
}

// ---- small dense helpers (device) -------------------------------------------------------
__device__ __forceinline__ void inv3(const double a[9], double out[9]) {
  const double c00 = a[4] * a[8] - a[5] * a[7];
  const double c01 = a[3] * a[8] - a[5] * a[6];
  const double c02 = a[3] * a[7] - a[4] * a[6];
  const double det = a[0] * c00 - a[1] * c01 + a[2] * c02;
  const double id = 1.0 / det;
  out[0] = c00 * id;
  out[1] = (a[2] * a[7] - a[1] * a[8]) * id;
  out[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  out[3] = -c01 * id;
  out[4] = (a[0] * a[8] - a[2] * a[6]) * id;
  out[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  out[6] = c02 * id;
  out[7] = (a[1] * a[6] - a[0] * a[7]) * id;
  out[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}

// Cholesky of a symmetric 3x3 given as (g00,g10,g11,g20,g21,g22); returns the inverse
// diagonal (i0,i1,i2) and the strict lower part (l10,l20,l21) of L.
struct Chol3 {
  double i0, i1, i2, l10, l20, l21;
};
__device__ __forceinline__ Chol3 chol3(double g00, double g10, double g11, double g20, double g21,
                                       double g22) {
  Chol3 c;
  const double l00 = sqrt(g00);
  c.i0 = 1.0 / l00;
  c.l10 = g10 * c.i0;
  c.l20 = g20 * c.i0;
  const double l11 = sqrt(g11 - c.l10 * c.l10);
  c.i1 = 1.0 / l11;
  c.l21 = (g21 - c.l20 * c.l10) * c.i1;
  const double l22 = sqrt(g22 - c.l20 * c.l20 - c.l21 * c.l21);
  c.i2 = 1.0 / l22;
  return c;
}
// solve L q = b in place
__device__ __forceinline__ void chol3_fwd(const Chol3& c, double& b0, double& b1, double& b2) {
  b0 = b0 * c.i0;
  b1 = (b1 - c.l10 * b0) * c.i1;
  b2 = (b2 - c.l20 * b0 - c.l21 * b1) * c.i2;
}

// sum over the G lanes of a lane group (G a power of two, groups aligned inside the warp)
template <int G>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// in-place Gauss-Jordan inverse of the n x n matrix A (shared memory, leading dim G+1), performed
// by the G lanes of a group, lane j owning column j.  No pivoting: the blocks are SPD.
template <int G>
__device__ __forceinline__ void group_invert(double (*A)[G + 1], int n, int j, int nw) {
  for (int k = 0; k < nw; ++k) {
    double pinv = 0., akj = 0.;
    const bool act = (k < n) && (j < n);
    if (act) {
      pinv = 1.0 / A[k][k];
      akj = A[k][j] * pinv;
    }
    __syncwarp();
    if (act && j != k) {
      for (int i = 0; i < n; ++i) {
        if (i == k) continue;
        A[i][j] -= A[i][k] * akj;
      }
      A[k][j] = akj;
    }
    __syncwarp();
    if (act && j == k) {
      for (int i = 0; i < n; ++i) {
        if (i == k) continue;
        A[i][k] = -A[i][k] * pinv;
      }
      A[k][k] = pinv;
    }
    __syncwarp();
  }
}

template <int G>
__device__ __forceinline__ int warp_max_n(int n) {
  int nw = n;
#pragma unroll
  for (int o = 16; o >= G; o >>= 1) nw = max(nw, __shfl_xor_sync(0xffffffffu, nw, o));
  return nw;
}

