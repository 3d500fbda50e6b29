// Exact algebraic hybridisation of the saddle-point system of MFD.build_lhs (mfd.py:851-914)
//
//     M v - DIV^T p = b_v ,   DIV v + C p = b_p            (solved by SuperLU in mfd.py:1349-1358)
//
// Every flux DOF g(f) gets a multiplier lambda_f.  With u_E = sigma o v|_E (outward fluxes of cell
// E), a = face areas, W_E = (mult_E M_E restricted to non-Neumann faces)^-1:
//     u_E = W_E (a p_E - a o lambda_E + sigma o beta_E),   a^T u_E + C_E p_E = b_p[E]
// (beta carries b_v of interior faces, given to the lower-index cell).  Eliminating u_E, p_E
// cell by cell leaves the SPD face system  A lambda = h,
//     A = sum_E P_E^T (a W_E a - r_E c_E^T / S_E) P_E ,  r = (aWa) 1, c = (aWa)^T 1, S = 1^T(aWa)1 + C_E
// whose pattern is the pattern of M (m_indptr / m_indices).  Boundary flux DOFs (one cell) have
// lambda prescribed: lambda_f = -sigma b_v / a_f (= the Dirichlet pressure), identity rows in A.
// The recovered [v ; p] is the solution of the original system (up to the CG tolerance).
#include <stdlib.h>

#include "common.cuh"
#include "cell_core.cuh"

struct HybArgs {
  const int* face_to_flux;
  const int* face_cell;
  const uint8_t* face_flag;  // nullable; bit0/bit1: interface face of a slab partition (see header)
  const int* m_indptr;
  const int* sell_ptr;  // non-NULL: the face-system values are laid out in SELL-32 (sell.cu)

  // address of entry k of row g, and the distance between consecutive entries of a row
  __device__ __forceinline__ int row_base(int g) const {
    return sell_ptr ? sell_ptr[g >> 5] + (g & 31) : m_indptr[g];
  }
  __device__ __forceinline__ int row_stride() const { return sell_ptr ? 32 : 1; }
  const int* m_ptr;
  const uint8_t* pos_m;
  int flux_dof;
  double alpha;
  const double* c_diag;
};

template <int G>
__global__ void __launch_bounds__(128)
k_hybrid_setup(mimpy_mesh_view m, HybArgs h, const double* __restrict__ m_e,
               const double* __restrict__ multipliers, double* __restrict__ w_e,
               double* __restrict__ a_vals) {
  constexpr int CPB = 128 / G;
  __shared__ double As[CPB][G][G + 1];
  __shared__ double s_a[CPB][G];
  __shared__ double s_r[CPB][G];
  __shared__ int s_row[CPB][G];
  __shared__ int s_flag[CPB][G];  // bit0: dropped (Neumann), bit1: boundary (single cell)
  const int slot = threadIdx.x / G;
  const int j = threadIdx.x % G;
  const int cell = blockIdx.x * CPB + slot;
  const bool cell_ok = cell < m.nc;
  const int base = cell_ok ? m.cell_ptr[cell] : 0;
  const int n = cell_ok ? m.cell_ptr[cell + 1] - base : 0;
  const bool valid = j < n;
  const int nw = warp_max_n<G>(n);
  int gj = -1, flag = 1;
  double area = 0.;
  if (valid) {
    const int f = m.cell_faces[base + j];
    area = rec_area(m.face_geom, (size_t)f);
    gj = h.face_to_flux[f];
    const bool iface = h.face_flag && (h.face_flag[f] & 3);  // the other cell lives on another rank
    flag = (gj < 0 ? 1 : 0) | ((h.face_cell[2 * (size_t)f + 1] < 0 && !iface) ? 2 : 0);
  }
  s_a[slot][j] = area;
  s_flag[slot][j] = flag;
  s_row[slot][j] = gj >= 0 ? h.row_base(gj) : -1;
  const int stride = h.row_stride();
  __syncwarp();
  const double mult = (multipliers && cell_ok) ? __ldg(multipliers + cell) : 1.;
  const int boff = cell_ok ? h.m_ptr[cell] : 0;
  for (int i = 0; i < n; ++i) {
    if (valid) {
      const bool drop = (s_flag[slot][i] & 1) || (flag & 1);
      As[slot][i][j] = drop ? (i == j ? 1. : 0.) : mult * m_e[boff + (size_t)i * n + j];
    }
  }
  __syncwarp();
  group_invert<G>(As[slot], n, j, nw);
  // masked inverse, W stored for rhs / recovery; scaled block aWa in place
  for (int i = 0; i < n; ++i) {
    if (valid) {
      const bool drop = (s_flag[slot][i] & 1) || (flag & 1);
      const double w = drop ? 0. : As[slot][i][j];
      w_e[boff + (size_t)i * n + j] = w;
      As[slot][i][j] = s_a[slot][i] * w * area;
    }
  }
  __syncwarp();
  double rj = 0., cj = 0.;  // row sum of row j, column sum of column j
  if (valid) {
    for (int k = 0; k < n; ++k) {
      rj += As[slot][j][k];
      cj += As[slot][k][j];
    }
  }
  s_r[slot][j] = rj;
  const double cd = cell_ok ? (h.c_diag ? __ldg(h.c_diag + cell) : h.alpha * __ldg(m.cell_volume + cell)) : 0.;
  const double S = group_sum<G>(cj) + cd;
  __syncwarp();
  for (int i = 0; i < n; ++i) {
    const int rs = s_row[slot][i];
    if (valid && rs >= 0 && gj >= 0) {
      const bool bnd = (s_flag[slot][i] & 2) || (flag & 2);
      double v = As[slot][i][j] - s_r[slot][i] * cj / S;
      if (bnd) v = (i == j) ? 1. : 0.;
      const unsigned pos = h.pos_m[boff + (size_t)i * n + j];
      if (i == j)
        atomicAdd(a_vals + rs + (int)pos * stride, v);
      else
        a_vals[rs + (int)pos * stride] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fused set-up for hexahedral cells (6 faces, methods 0/1/6): one thread per cell builds M_E in
// registers (cell_core.cuh), inverts it in a private column of a shared-memory tile and scatters
// the face-system contributions -- the M_E blocks never go to memory.  Replaces k_local_matrices +
// k_hybrid_setup<8> (9.2 + 11.5 ms at 256^3).  W_E leaves through the tile, so the stores of a CTA
// are one contiguous 36 KB piece of w_e.
// ---------------------------------------------------------------------------------------------
#define HX_THREADS 128
#define HX_LD (HX_THREADS + 1)

template <int METHOD>
__global__ void __launch_bounds__(HX_THREADS, 3)
k_hybrid_setup_hex(mimpy_mesh_view m, HybArgs h, int k_unity, const double* __restrict__ multipliers,
                   double* __restrict__ w_e, double* __restrict__ a_vals) {
  constexpr int NF = 6;
  extern __shared__ double tile[];  // [36][HX_LD]: entry e of the cell of thread t at tile[e * HX_LD + t]
  const int t = threadIdx.x;
  const long long cell0 = (long long)blockIdx.x * HX_THREADS;
  const long long cell = cell0 + t;
  const bool ok = cell < m.nc;
  double* A = tile + t;
  int gj[NF], rowb[NF];
  double area[NF];
  unsigned flagbits = 0;  // bit i: dropped (Neumann), bit 8+i: boundary (single cell)
  double cd = 0.;
  if (ok) {
    double K[9], xc[3], sgn[NF];
    double2 rec[NF][4];
    int fid[NF];
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      fid[i] = m.cell_faces[cell * NF + i];
      sgn[i] = (double)m.cell_sigma[cell * NF + i];
#pragma unroll
      for (int q = 0; q < 4; ++q) rec[i][q] = __ldg(rec_chunk(m.face_geom, (size_t)fid[i], q));
      area[i] = rec[i][0].x;
      gj[i] = h.face_to_flux[fid[i]];
      const bool iface = h.face_flag && (h.face_flag[fid[i]] & 3);
      if (gj[i] < 0) flagbits |= 1u << i;
      if (h.face_cell[2 * (size_t)fid[i] + 1] < 0 && !iface) flagbits |= 1u << (8 + i);
      rowb[i] = gj[i] >= 0 ? h.row_base(gj[i]) : -1;
    }
#pragma unroll
    for (int q = 0; q < 9; ++q) K[q] = k_unity ? ((q % 4 == 0) ? 1. : 0.) : __ldg(m.cell_k + cell * 9 + q);
#pragma unroll
    for (int q = 0; q < 3; ++q) xc[q] = __ldg(m.cell_centroid + cell * 3 + q);
    const double vol = __ldg(m.cell_volume + cell);
    const double mult = multipliers ? __ldg(multipliers + cell) : 1.;
    cd = h.c_diag ? __ldg(h.c_diag + cell) : h.alpha * vol;
    // mult * M_E (raw, i.e. acting on the outward fluxes u = sigma o v), masked like k_hybrid_setup
    cell_entries<NF, METHOD>(K, xc, vol, mult, rec, sgn, [&](int i, int j, double v) {
      const bool drop = ((flagbits >> i) | (flagbits >> j)) & 1u;
      A[(i * NF + j) * HX_LD] = drop ? (i == j ? 1. : 0.) : sgn[i] * sgn[j] * v;
    });
    // in-place Gauss-Jordan without pivoting, same elimination order as group_invert (common.cuh)
#pragma unroll 1
    for (int k = 0; k < NF; ++k) {
      const double pinv = 1.0 / A[(k * NF + k) * HX_LD];
#pragma unroll
      for (int j = 0; j < NF; ++j) {
        if (j == k) continue;
        const double akj = A[(k * NF + j) * HX_LD] * pinv;
#pragma unroll
        for (int i = 0; i < NF; ++i) {
          if (i == k) continue;
          A[(i * NF + j) * HX_LD] -= A[(i * NF + k) * HX_LD] * akj;
        }
        A[(k * NF + j) * HX_LD] = akj;
      }
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        if (i == k) continue;
        A[(i * NF + k) * HX_LD] = -A[(i * NF + k) * HX_LD] * pinv;
      }
      A[(k * NF + k) * HX_LD] = pinv;
    }
    // W = masked inverse (stays in the tile for the coalesced copy-out); aWa, r, c, S in registers
    double rsum[NF], csum[NF];
#pragma unroll
    for (int i = 0; i < NF; ++i) rsum[i] = csum[i] = 0.;
#pragma unroll
    for (int i = 0; i < NF; ++i)
#pragma unroll
      for (int j = 0; j < NF; ++j) {
        const bool drop = ((flagbits >> i) | (flagbits >> j)) & 1u;
        const double w = drop ? 0. : A[(i * NF + j) * HX_LD];
        A[(i * NF + j) * HX_LD] = w;
        const double awa = area[i] * w * area[j];
        rsum[i] += awa;
        csum[j] += awa;
      }
    double S = cd;
#pragma unroll
    for (int j = 0; j < NF; ++j) S += csum[j];
    const int stride = h.row_stride();
    const uint8_t* pm = h.pos_m + cell * (NF * NF);
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      if (rowb[i] < 0) continue;
#pragma unroll
      for (int j = 0; j < NF; ++j) {
        if (gj[j] < 0) continue;
        const bool bnd = ((flagbits >> (8 + i)) | (flagbits >> (8 + j))) & 1u;
        double v = area[i] * A[(i * NF + j) * HX_LD] * area[j] - rsum[i] * csum[j] / S;
        if (bnd) v = (i == j) ? 1. : 0.;
        const unsigned pos = pm[i * NF + j];
        if (i == j) atomicAdd(a_vals + rowb[i] + (int)pos * stride, v);
        else a_vals[rowb[i] + (int)pos * stride] = v;
      }
    }
  }
  __syncthreads();
  // copy-out of W: the CTA's cells are one contiguous piece of w_e (36 doubles per cell)
  const long long ncta = min((long long)HX_THREADS, (long long)m.nc - cell0);
  for (int q = t; q < ncta * NF * NF; q += HX_THREADS) {
    const int c = q / (NF * NF), e = q - c * (NF * NF);
    w_e[cell0 * (NF * NF) + q] = tile[e * HX_LD + c];
  }
}

// INVERT = false: the diagonal itself (inverted by k_invert after the interface exchange of a slab partition)
template <bool INVERT>
__global__ void k_diag_inv(int F, HybArgs h, const uint8_t* __restrict__ diag_pos,
                           const double* __restrict__ a_vals, double* __restrict__ dinv) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= F) return;
  const double d = a_vals[h.row_base(g) + (int)diag_pos[g] * h.row_stride()];
  dinv[g] = INVERT ? 1.0 / d : d;
}

__global__ void k_invert(int F, double* __restrict__ d) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < F) d[g] = 1.0 / d[g];
}

__global__ void k_zero_diag_m(int F, HybArgs h, const uint8_t* __restrict__ diag_pos,
                              double* __restrict__ vals) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= F) return;
  vals[h.row_base(g) + (int)diag_pos[g] * h.row_stride()] = 0.;
}

// MODE 0: h (face-system right-hand side) ; MODE 1: recover [v ; p] from lambda
template <int G, int MODE>
__global__ void __launch_bounds__(128)
k_hybrid_local(mimpy_mesh_view m, HybArgs h, const double* __restrict__ w_e,
               const double* __restrict__ rhs, const double* __restrict__ lambda,
               double* __restrict__ out) {
  constexpr int CPB = 128 / G;
  __shared__ double Ws[CPB][G][G + 1];
  __shared__ double s_a[CPB][G];
  __shared__ double s_t[CPB][G];
  const int slot = threadIdx.x / G;
  const int j = threadIdx.x % G;
  const int cell = blockIdx.x * CPB + slot;
  const bool cell_ok = cell < m.nc;
  const int base = cell_ok ? m.cell_ptr[cell] : 0;
  const int n = cell_ok ? m.cell_ptr[cell + 1] - base : 0;
  const bool valid = j < n;
  const int F = h.flux_dof;
  int gj = -1;
  double area = 0., sg = 0., t = 0., lam = 0.;
  bool bnd = false, owner = false, iface = false;
  if (valid) {
    const int f = m.cell_faces[base + j];
    sg = (double)m.cell_sigma[base + j];
    area = rec_area(m.face_geom, (size_t)f);
    gj = h.face_to_flux[f];
    const unsigned ff = h.face_flag ? (h.face_flag[f] & 3) : 0;
    bnd = h.face_cell[2 * (size_t)f + 1] < 0 && !ff;
    owner = ff ? (ff & 1) : (h.face_cell[2 * (size_t)f] == cell);
    iface = ff != 0;
    if (gj >= 0) {
      const double bv = rhs[gj];
      double beta = 0.;
      if (bnd)
        lam = -sg * bv / area;
      else {
        if (MODE == 1) lam = lambda[gj];
        if (owner) beta = bv;
      }
      t = -area * lam + sg * beta;
    }
  }
  s_a[slot][j] = area;
  s_t[slot][j] = t;
  const int boff = cell_ok ? h.m_ptr[cell] : 0;
  for (int i = 0; i < n; ++i)
    if (valid) Ws[slot][i][j] = w_e[boff + (size_t)i * n + j];
  __syncwarp();
  double wa = 0., wt = 0.;
  if (valid) {
    for (int k = 0; k < n; ++k) {
      const double w = Ws[slot][j][k];
      wa += w * s_a[slot][k];
      wt += w * s_t[slot][k];
    }
  }
  const double cd = cell_ok ? (h.c_diag ? __ldg(h.c_diag + cell) : h.alpha * __ldg(m.cell_volume + cell)) : 0.;
  const double S = group_sum<G>(area * wa) + cd;
  const double awt = group_sum<G>(area * wt);
  const double bp = cell_ok ? rhs[F + cell] : 0.;
  const double p = (bp - awt) / S;
  const double u = p * wa + wt;
  if (MODE == 0) {
    if (valid && gj >= 0) {
      if (bnd)
        out[gj] = lam;
      else
        atomicAdd(out + gj, area * u);  // two cells on a zeroed slot: order-free
    }
  } else {
    if (j == 0 && cell_ok) out[F + cell] = p;
    // the lower-index cell reports the face flux; on a slab interface the only local cell does
    if (valid && gj >= 0 && (owner || iface)) out[gj] = sg * u;
  }
}

static HybArgs make_hyb(const mimpy_symbolic* sym, double alpha, const double* c_diag) {
  HybArgs h;
  h.face_to_flux = sym->face_to_flux;
  h.face_cell = sym->face_cell;
  h.face_flag = sym->face_flag;
  h.m_indptr = sym->m_indptr;
  h.sell_ptr = nullptr;
  h.m_ptr = sym->m_ptr;
  h.pos_m = sym->pos_m;
  h.flux_dof = sym->flux_dof;
  h.alpha = alpha;
  h.c_diag = c_diag;
  return h;
}

#define DISPATCH_G(G, CALL8, CALL16, CALL32) \
  do {                                       \
    if ((G) == 8) { CALL8; }                 \
    else if ((G) == 16) { CALL16; }          \
    else { CALL32; }                         \
  } while (0)

extern "C" {

int mimpy_b200_hybrid_setup(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym,
                            const double* m_e, const double* multipliers, double alpha,
                            const double* c_diag, double* w_e, double* a_vals, double* a_diag_inv,
                            const int32_t* sell_ptr) {
  REQUIRE(ctx && mesh && sym && m_e && w_e && a_vals && a_diag_inv, "NULL argument");
  REQUIRE(mesh->max_faces >= 1 && mesh->max_faces <= MIMPY_MAX_CELL_FACES, "max_faces out of range");
  if (mesh->nc <= 0) return MIMPY_OK;
  const int F = sym->flux_dof;
  cudaStream_t st = ctx->stream;
  HybArgs h = make_hyb(sym, alpha, c_diag);
  h.sell_ptr = sell_ptr;
  if (F > 0) {
    k_zero_diag_m<<<(F + 255) / 256, 256, 0, st>>>(F, h, sym->diag_pos, a_vals);
    KERNEL_CHECK();
  }
  const int G = mimpy_group_size(mesh->max_faces);
  const int blocks = (mesh->nc + 128 / G - 1) / (128 / G);
  DISPATCH_G(G, (k_hybrid_setup<8><<<blocks, 128, 0, st>>>(*mesh, h, m_e, multipliers, w_e, a_vals)),
             (k_hybrid_setup<16><<<blocks, 128, 0, st>>>(*mesh, h, m_e, multipliers, w_e, a_vals)),
             (k_hybrid_setup<32><<<blocks, 128, 0, st>>>(*mesh, h, m_e, multipliers, w_e, a_vals)));
  KERNEL_CHECK();
  if (F > 0) {
    k_diag_inv<false><<<(F + 255) / 256, 256, 0, st>>>(F, h, sym->diag_pos, a_vals, a_diag_inv);
    KERNEL_CHECK();
    // slab partition: the diagonal of an interface row is the sum of both ranks' halves
    int rc = mimpy_halo_exchange_add(ctx, a_diag_inv);
    if (rc) return rc;
    k_invert<<<(F + 255) / 256, 256, 0, st>>>(F, a_diag_inv);
    KERNEL_CHECK();
  }
  return MIMPY_OK;
}

int mimpy_b200_hybrid_setup_fused(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym,
                                  int32_t method, int32_t flags, const double* multipliers, double alpha,
                                  const double* c_diag, double* w_e, double* a_vals, double* a_diag_inv,
                                  const int32_t* sell_ptr, int32_t* face_rowbase_work, double* tau_out,
                                  int32_t* tau_written_host) {
  REQUIRE(ctx && mesh && sym && w_e && a_vals && a_diag_inv, "NULL argument");
  if (tau_written_host) *tau_written_host = 0;
  if (mesh->uniform_faces != 6 || (method != 0 && method != 1 && method != 6)) return MIMPY_E_UNSUPPORTED;
  if (mesh->nc <= 0) return MIMPY_OK;
  const int F = sym->flux_dof;
  cudaStream_t st = ctx->stream;
  HybArgs h = make_hyb(sym, alpha, c_diag);
  h.sell_ptr = sell_ptr;
  // structured hex meshes: the pipelined brick kernel (bulk-copy staging, closed-form positions; the two halves
  // of a diagonal entry are added by a finishing pass that also writes the Jacobi diagonal: no slot is zeroed
  // first and nothing is added atomically)
  bool done = false;
  if (face_rowbase_work && !getenv("MIMPY_B200_SETUP_OLD")) {
    int rc = mimpy_face_rowbase(ctx, mesh, sym, sell_ptr, face_rowbase_work);
    if (rc) return rc;
    rc = mimpy_hybrid_setup_brick(ctx, mesh, sym, method, flags, multipliers, alpha, c_diag, face_rowbase_work,
                                  sell_ptr ? 32 : 1, w_e, tau_out, a_vals, a_diag_inv, ctx->world > 1 ? 0 : 1);
    if (rc == MIMPY_OK) done = true;
    else if (rc != MIMPY_E_UNSUPPORTED) return rc;
  }
  if (done) {
    if (F > 0) {
      if (ctx->world > 1) {  // slab partition: the diagonal of an interface row is the sum of both ranks' halves
        int rc = mimpy_halo_exchange_add(ctx, a_diag_inv);
        if (rc) return rc;
        k_invert<<<(F + 255) / 256, 256, 0, st>>>(F, a_diag_inv);
        KERNEL_CHECK();
      }
    }
    if (tau_written_host) *tau_written_host = tau_out ? 1 : 0;
    return MIMPY_OK;
  }
  if (F > 0) {
    k_zero_diag_m<<<(F + 255) / 256, 256, 0, st>>>(F, h, sym->diag_pos, a_vals);
    KERNEL_CHECK();
  }
  const size_t smem = sizeof(double) * 36 * HX_LD;
  const int blocks = (mesh->nc + HX_THREADS - 1) / HX_THREADS;
  const int ku = flags & MIMPY_FLAG_K_UNITY;
#define HX_LAUNCH(M)                                                                                     \
  do {                                                                                                   \
    CUDA_TRY(cudaFuncSetAttribute(k_hybrid_setup_hex<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_hybrid_setup_hex<M><<<blocks, HX_THREADS, smem, st>>>(*mesh, h, ku, multipliers, w_e, a_vals);       \
  } while (0)
  if (method == 0) HX_LAUNCH(0);
  else if (method == 1) HX_LAUNCH(1);
  else HX_LAUNCH(6);
#undef HX_LAUNCH
  KERNEL_CHECK();
  if (F > 0) {
    k_diag_inv<false><<<(F + 255) / 256, 256, 0, st>>>(F, h, sym->diag_pos, a_vals, a_diag_inv);
    KERNEL_CHECK();
    int rc = mimpy_halo_exchange_add(ctx, a_diag_inv);
    if (rc) return rc;
    k_invert<<<(F + 255) / 256, 256, 0, st>>>(F, a_diag_inv);
    KERNEL_CHECK();
  }
  return MIMPY_OK;
}

int mimpy_b200_hybrid_rhs(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym,
                          const double* w_e, double alpha, const double* c_diag, const double* rhs,
                          double* cell_tmp, double* hvec) {
  (void)cell_tmp;
  REQUIRE(ctx && mesh && sym && w_e && rhs && hvec, "NULL argument");
  if (mesh->nc <= 0) return MIMPY_OK;
  cudaStream_t st = ctx->stream;
  CUDA_TRY(cudaMemsetAsync(hvec, 0, sizeof(double) * (size_t)sym->flux_dof, st));
  const int G = mimpy_group_size(mesh->max_faces);
  const int blocks = (mesh->nc + 128 / G - 1) / (128 / G);
  HybArgs h = make_hyb(sym, alpha, c_diag);
  DISPATCH_G(G, (k_hybrid_local<8, 0><<<blocks, 128, 0, st>>>(*mesh, h, w_e, rhs, nullptr, hvec)),
             (k_hybrid_local<16, 0><<<blocks, 128, 0, st>>>(*mesh, h, w_e, rhs, nullptr, hvec)),
             (k_hybrid_local<32, 0><<<blocks, 128, 0, st>>>(*mesh, h, w_e, rhs, nullptr, hvec)));
  KERNEL_CHECK();
  return mimpy_halo_exchange_add(ctx, hvec);  // interface rows: own + neighbour (no-op on one rank)
}

int mimpy_b200_hybrid_recover(mimpy_ctx* ctx, const mimpy_mesh_view* mesh,
                              const mimpy_symbolic* sym, const double* w_e, double alpha,
                              const double* c_diag, const double* rhs, const double* lambda,
                              double* solution) {
  REQUIRE(ctx && mesh && sym && w_e && rhs && lambda && solution, "NULL argument");
  if (mesh->nc <= 0) return MIMPY_OK;
  cudaStream_t st = ctx->stream;
  const int G = mimpy_group_size(mesh->max_faces);
  const int blocks = (mesh->nc + 128 / G - 1) / (128 / G);
  HybArgs h = make_hyb(sym, alpha, c_diag);
  DISPATCH_G(G, (k_hybrid_local<8, 1><<<blocks, 128, 0, st>>>(*mesh, h, w_e, rhs, lambda, solution)),
             (k_hybrid_local<16, 1><<<blocks, 128, 0, st>>>(*mesh, h, w_e, rhs, lambda, solution)),
             (k_hybrid_local<32, 1><<<blocks, 128, 0, st>>>(*mesh, h, w_e, rhs, lambda, solution)));
  KERNEL_CHECK();
  return MIMPY_OK;
}

}  // extern "C"
