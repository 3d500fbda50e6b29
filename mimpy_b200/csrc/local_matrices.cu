// Batched local inner-product matrices M_E (mfd.py:248-483) and the fused numeric assembly of
// the saddle-point CSR values (mfd.py:545-599, 203-246, 745-774, 729-743).
//
// Mapping: a group of G lanes (8/16/32, G >= faces per cell) owns one cell, lane j = local face
// j.  Per-face vectors R_j, N_j (and the orthonormalised Q_j) are staged in shared memory so that
// every lane can form its column M_E[:, j]; for a fixed row i the G lanes then write one
// contiguous row segment of the block / of CSR row g(f_i) with a single store instruction.
//
//   R[f,:] = a_f (x_f - x_E)                                   mfd.py:274-281
//   N[f,:] = s_f K n_f                                         mfd.py:307-309
//   M0     = R (R^T N)^-1 R^T                                  mfd.py:377
//   P      = I - Q Q^T,  Q = orth(N)   (== C_E C_E^T of the SVD null-space basis, mfd.py:318-320;
//            Q from CholeskyQR2 so the projector is accurate to O(eps) for anisotropic K)
//   method 0..7: stabilisation                                 mfd.py:379-468
#include "common.cuh"
#include "cell_core.cuh"

#define SM_PAD 1

template <int G>
struct CellSmem {
  double r[G][3];
  double n[G][3];
  double q[G][3];
  double d[G];      // per-row coefficient (method 6) / scratch
  double sg[G];     // orientation as double
  int row[G];       // CSR row start of flux row g(f_i), -1 when dropped
};

// method 4 only: W_E per cell (G x (G+1) doubles), inverted in place -- dynamic shared memory so
// that the other methods keep their occupancy
extern __shared__ __align__(16) unsigned char lm_dyn_smem[];

struct ScatterArgs {
  const int* face_to_flux;
  const int* face_cell;
  const int* indptr;
  const int* m_ptr;
  const uint8_t* pos_m;
  const uint8_t* pos_dt;
  const uint8_t* pos_d;
  const double* multipliers;
  const double* c_diag;
  double alpha;
  double* vals;
  int flux_dof;
};

// MT: the construction method at compile time (0, 1, 6: the methods of the BASELINE configs), or -1 = the run-time
// argument (all eight methods in one body: more registers, dead branches)
template <int G, bool SCATTER, int MT>
// Resident CTAs per SM: the kernel waits on its gather chains (cell_ptr -> cell_faces -> record / face_to_flux -> indptr), so
// occupancy pays -- measured on 128^3 hexes, method 1: 4 CTAs (126 registers) 1.85 ms, 5 (94) 1.55 ms, 6 (80) 1.32 ms, 8 (64,
// spills) 1.26 ms.  Methods 0 and 1 fit 80 registers without spilling; the others get 5 CTAs.
#ifndef LM_MINB
#define LM_MINB(MT) (((MT) == 0 || (MT) == 1) ? 6 : 5)
#endif
__global__ void __launch_bounds__(128, LM_MINB(MT))
k_local_matrices(mimpy_mesh_view m, int method_rt, int flags, const int* __restrict__ m_ptr,
                 double* __restrict__ m_e, double* __restrict__ diagonality,
                 double* __restrict__ consistency, ScatterArgs sc) {
  const int method = MT >= 0 ? MT : method_rt;
  constexpr int CPB = 128 / G;
  __shared__ CellSmem<G> sm[CPB];
  const int slot = threadIdx.x / G;
  const int j = threadIdx.x % G;
  const int cell = blockIdx.x * CPB + slot;
  const bool cell_ok = cell < m.nc;
  const int base = cell_ok ? m.cell_ptr[cell] : 0;
  const int n = cell_ok ? m.cell_ptr[cell + 1] - base : 0;
  const bool valid = j < n;
  const bool k_unity = flags & 1;
  CellSmem<G>& s = sm[slot];

  // ---- gather: face record, cell data ---------------------------------------------------
  double area = 0., nrm[3] = {0., 0., 0.}, xf[3] = {0., 0., 0.}, sg = 0.;
  int f = -1;
  if (valid) {
    f = m.cell_faces[base + j];
    sg = (double)m.cell_sigma[base + j];
    const double2 a0 = __ldg(rec_chunk(m.face_geom, (size_t)f, 0)), a1 = __ldg(rec_chunk(m.face_geom, (size_t)f, 1));
    const double2 a2 = __ldg(rec_chunk(m.face_geom, (size_t)f, 2)), a3 = __ldg(rec_chunk(m.face_geom, (size_t)f, 3));
    area = a0.x;
    nrm[0] = a0.y;
    nrm[1] = a1.x;
    nrm[2] = a1.y;
    xf[0] = a2.x;
    xf[1] = a2.y;
    xf[2] = a3.x;
  }
  double K[9] = {1., 0., 0., 0., 1., 0., 0., 0., 1.};
  double xc[3] = {0., 0., 0.}, vol = 1.;
  if (cell_ok) {
    if (!k_unity) {
#pragma unroll
      for (int t = 0; t < 9; ++t) K[t] = __ldg(m.cell_k + 9 * (size_t)cell + t);
    }
#pragma unroll
    for (int t = 0; t < 3; ++t) xc[t] = __ldg(m.cell_centroid + 3 * (size_t)cell + t);
    vol = __ldg(m.cell_volume + cell);
  }
  double r[3], nn[3];
#pragma unroll
  for (int t = 0; t < 3; ++t) r[t] = area * (xf[t] - xc[t]);
#pragma unroll
  for (int t = 0; t < 3; ++t)
    nn[t] = sg * (K[3 * t] * nrm[0] + K[3 * t + 1] * nrm[1] + K[3 * t + 2] * nrm[2]);
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    s.r[j][t] = r[t];
    s.n[j][t] = nn[t];
  }
  s.sg[j] = sg;
  __syncwarp();

  // ---- 3x3 cell quantities (every lane redundantly; SIMT makes this free per warp) --------
  double T[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double g00 = 0, g10 = 0, g11 = 0, g20 = 0, g21 = 0, g22 = 0;
  double rr0[3] = {0, 0, 0};  // first row of R^T R (method 0)
  double rr11 = 0, rr21 = 0, rr22 = 0;  // rest of R^T R (method 4)
  for (int i = 0; i < n; ++i) {
    const double ri0 = s.r[i][0], ri1 = s.r[i][1], ri2 = s.r[i][2];
    const double ni0 = s.n[i][0], ni1 = s.n[i][1], ni2 = s.n[i][2];
    T[0] += ri0 * ni0; T[1] += ri0 * ni1; T[2] += ri0 * ni2;
    T[3] += ri1 * ni0; T[4] += ri1 * ni1; T[5] += ri1 * ni2;
    T[6] += ri2 * ni0; T[7] += ri2 * ni1; T[8] += ri2 * ni2;
    g00 += ni0 * ni0; g10 += ni1 * ni0; g11 += ni1 * ni1;
    g20 += ni2 * ni0; g21 += ni2 * ni1; g22 += ni2 * ni2;
    rr0[0] += ri0 * ri0; rr0[1] += ri0 * ri1; rr0[2] += ri0 * ri2;
    rr11 += ri1 * ri1; rr21 += ri2 * ri1; rr22 += ri2 * ri2;
  }
  double Ti[9];
  if (n > 0) {
    inv3(T, Ti);
  } else {
#pragma unroll
    for (int t = 0; t < 9; ++t) Ti[t] = 0.;
  }
  // y_j = T^-1 r_j  =>  M0[i][j] = r_i . y_j
  double y[3];
#pragma unroll
  for (int t = 0; t < 3; ++t) y[t] = Ti[3 * t] * r[0] + Ti[3 * t + 1] * r[1] + Ti[3 * t + 2] * r[2];

  // ---- projector: CholeskyQR2 of N (n x 3) ------------------------------------------------
  // method 4 (mfd.py:338-355, 420-421) needs D D^T = I - Q_R Q_R^T with Q_R = orth(R) instead
  const bool q_from_r = (method == 4);
  double q[3] = {q_from_r ? r[0] : nn[0], q_from_r ? r[1] : nn[1], q_from_r ? r[2] : nn[2]};
  const bool need_p = (method <= 4) || (method == 6);
  if (need_p) {
    if (q_from_r) {
      g00 = rr0[0]; g10 = rr0[1]; g11 = rr11; g20 = rr0[2]; g21 = rr21; g22 = rr22;
    }
    if (n > 0) {
      Chol3 c1 = chol3r(g00, g10, g11, g20, g21, g22);
      chol3_fwd(c1, q[0], q[1], q[2]);
    }
#pragma unroll
    for (int t = 0; t < 3; ++t) s.q[j][t] = valid ? q[t] : 0.;
    __syncwarp();
    double h00 = 0, h10 = 0, h11 = 0, h20 = 0, h21 = 0, h22 = 0;
    for (int i = 0; i < n; ++i) {
      const double q0 = s.q[i][0], q1 = s.q[i][1], q2 = s.q[i][2];
      h00 += q0 * q0; h10 += q1 * q0; h11 += q1 * q1;
      h20 += q2 * q0; h21 += q2 * q1; h22 += q2 * q2;
    }
    __syncwarp();
    if (n > 0) {
      Chol3 c2 = chol3r(h00, h10, h11, h20, h21, h22);
      chol3_fwd(c2, q[0], q[1], q[2]);
    }
#pragma unroll
    for (int t = 0; t < 3; ++t) s.q[j][t] = valid ? q[t] : 0.;
  }

  // ---- stabilisation coefficient(s) -------------------------------------------------------
  double u = 0.;          // scalar coefficient of P (methods 0,1,2,3)
  double drow = 0.;       // per-face coefficient (methods 5,6,7), staged in s.d
  if (method == 0) {      // u = [(R^T R) T^-1]_00                       mfd.py:380-382
    u = rr0[0] * Ti[0] + rr0[1] * Ti[3] + rr0[2] * Ti[6];
  } else if (method == 1) {  // u = |E| / tr K                              mfd.py:387
    u = vol / (K[0] + K[4] + K[8]);
  } else if (method == 2 || method == 6) {
    double Ki[9];
    inv3(K, Ki);
    if (method == 2) {    // u = sum_f d^T K^-1 d / (dim^2 n_E)         mfd.py:395-402
      double dd[3] = {xf[0] - xc[0], xf[1] - xc[1], xf[2] - xc[2]};
      double e = 0.;
#pragma unroll
      for (int a = 0; a < 3; ++a)
        e += dd[a] * (Ki[3 * a] * dd[0] + Ki[3 * a + 1] * dd[1] + Ki[3 * a + 2] * dd[2]);
      e = valid ? e : 0.;
      u = group_sum<G>(e) / 9. / (double)(n > 0 ? n : 1);
    } else {              // d_f = (s n)^T K^-1 (s n) a_f |x_f - x_E|   mfd.py:446-453
      double e = 0.;
#pragma unroll
      for (int a = 0; a < 3; ++a)
        e += nrm[a] * (Ki[3 * a] * nrm[0] + Ki[3 * a + 1] * nrm[1] + Ki[3 * a + 2] * nrm[2]);
      const double dx = xf[0] - xc[0], dy = xf[1] - xc[1], dz = xf[2] - xc[2];
      drow = e * area * sqrt(dx * dx + dy * dy + dz * dz);
    }
  } else if (method == 3) {  // lambda = tr(M0)/2                           mfd.py:409-413
    double e = valid ? (r[0] * y[0] + r[1] * y[1] + r[2] * y[2]) : 0.;
    u = .5 * group_sum<G>(e);
  } else if (method == 5) {  // diag R[i,m]/N[i,m], m = argmax |R[i,:]|     mfd.py:425-436
    int mi = 0;
    double cur = fabs(r[0]);
    if (fabs(r[1]) > cur) { mi = 1; cur = fabs(r[1]); }
    if (fabs(r[2]) > cur) { mi = 2; }
    drow = valid ? r[mi] / nn[mi] : 0.;
  } else if (method == 7) {  // diag |R_i| a_i / K00                        mfd.py:467
    drow = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) * area / K[0];
  }
  s.d[j] = drow;

  // ---- scatter bookkeeping ----------------------------------------------------------------
  int gj = -1;
  double mult = 1.;
  bool shared_face = false;
  if (SCATTER) {
    if (valid) {
      gj = sc.face_to_flux[f];
      shared_face = sc.face_cell[2 * (size_t)f + 1] >= 0;
    }
    s.row[j] = (gj >= 0) ? sc.indptr[gj] : -1;
    if (sc.multipliers && cell_ok) mult = __ldg(sc.multipliers + cell);
  }
  __syncwarp();

  // ---- column j of M_E, row by row ----------------------------------------------------------
  const int boff = cell_ok ? m_ptr[cell] : 0;
  double fro2 = 0., off2 = 0., cons2 = 0.;
  const bool want_cons = consistency != nullptr;
  // warp-uniform trip count: groups of one warp may own cells with different face counts and
  // the group reductions below use full-warp shuffles
  int nw = n;
#pragma unroll
  for (int o = 16; o >= G; o >>= 1) nw = max(nw, __shfl_xor_sync(0xffffffffu, nw, o));
  double (*wbuf)[G + 1] = reinterpret_cast<double (*)[G + 1]>(lm_dyn_smem) + (size_t)slot * G;
  if (method == 4) {
    // W_E = (N (N^T R)^-1 N^T + D D^T) tr(K)/|E| ;  M_E = W_E^-1          mfd.py:349-353, 421
    double z[3];  // (N^T R)^-1 N_j = T^-T N_j
#pragma unroll
    for (int t = 0; t < 3; ++t) z[t] = Ti[t] * nn[0] + Ti[3 + t] * nn[1] + Ti[6 + t] * nn[2];
    const double scale = (K[0] + K[4] + K[8]) / vol;
    for (int i = 0; i < n; ++i) {
      if (valid) {
        const double w1 = s.n[i][0] * z[0] + s.n[i][1] * z[1] + s.n[i][2] * z[2];
        const double ddt = ((i == j) ? 1. : 0.) - (s.q[i][0] * q[0] + s.q[i][1] * q[1] + s.q[i][2] * q[2]);
        wbuf[i][j] = (w1 + ddt) * scale;
      }
    }
    __syncwarp();
    group_invert<G>(wbuf, n, j, nw);
  }
  for (int i = 0; i < nw; ++i) {
    const bool row_ok = i < n;
    double v;
    if (method == 4) {
      v = (row_ok && valid) ? wbuf[i][j] : 0.;
    } else if (method == 5 || method == 7) {
      v = (i == j) ? s.d[i] : 0.;
    } else {
      const double m0 = s.r[i][0] * y[0] + s.r[i][1] * y[1] + s.r[i][2] * y[2];
      const double pij = ((i == j) ? 1. : 0.) - (s.q[i][0] * q[0] + s.q[i][1] * q[1] + s.q[i][2] * q[2]);
      const double coef = (method == 6) ? s.d[i] : u;
      v = m0 + coef * pij;
    }
    if (!valid || !row_ok) v = 0.;
    if (m_e && valid && row_ok) m_e[boff + (size_t)i * n + j] = v;
    if (diagonality) {
      fro2 += v * v;
      if (i != j) off2 += v * v;
    }
    if (want_cons) {
      // (M N - R)[i,:] = sum_j M[i][j] N_j - R_i
      double e0 = group_sum<G>(v * nn[0]) - (row_ok ? s.r[i][0] : 0.);
      double e1 = group_sum<G>(v * nn[1]) - (row_ok ? s.r[i][1] : 0.);
      double e2 = group_sum<G>(v * nn[2]) - (row_ok ? s.r[i][2] : 0.);
      cons2 += e0 * e0 + e1 * e1 + e2 * e2;
    }
    if (SCATTER) {
      const int rs = row_ok ? s.row[i] : -1;
      if (valid && rs >= 0 && gj >= 0) {
        const unsigned pos = sc.pos_m[boff + (size_t)i * n + j];
        const double val = s.sg[i] * sg * mult * v;
        if (i == j)
          atomicAdd(sc.vals + rs + pos, val);  // <=2 contributors on a zeroed slot: order-free
        else
          sc.vals[rs + pos] = val;
      }
    }
  }
  if (diagonality) {
    fro2 = group_sum<G>(fro2);
    off2 = group_sum<G>(off2);
    if (j == 0 && cell_ok) diagonality[cell] = sqrt(off2) / sqrt(fro2);
  }
  if (want_cons && j == 0 && cell_ok) consistency[cell] = sqrt(cons2);

  if (SCATTER) {
    // DIV / -DIV^T / C                                            mfd.py:232-243, 766-772
    const int F = sc.flux_dof;
    if (valid && gj >= 0) {
      const double e = sg * area;
      sc.vals[s.row[j] + sc.pos_dt[base + j]] = -e;
      sc.vals[sc.indptr[F + cell] + sc.pos_d[base + j]] = e;
    }
    if (j == 0 && cell_ok) {
      const double cd = sc.c_diag ? __ldg(sc.c_diag + cell) : sc.alpha * vol;
      sc.vals[sc.indptr[F + cell + 1] - 1] = cd;
    }
  }
  (void)shared_face;
}

// scatter of caller-provided blocks (same slots as above)
template <int G>
__global__ void __launch_bounds__(128)
k_scatter_blocks(mimpy_mesh_view m, const double* __restrict__ m_e, ScatterArgs sc) {
  constexpr int CPB = 128 / G;
  __shared__ int s_row[CPB][G];
  __shared__ double s_sg[CPB][G];
  const int slot = threadIdx.x / G;
  const int j = threadIdx.x % G;
  const int cell = blockIdx.x * CPB + slot;
  const bool cell_ok = cell < m.nc;
  const int base = cell_ok ? m.cell_ptr[cell] : 0;
  const int n = cell_ok ? m.cell_ptr[cell + 1] - base : 0;
  const bool valid = j < n;
  int f = -1, gj = -1;
  double sg = 0., area = 0.;
  if (valid) {
    f = m.cell_faces[base + j];
    sg = (double)m.cell_sigma[base + j];
    area = rec_area(m.face_geom, (size_t)f);
    gj = sc.face_to_flux[f];
  }
  s_row[slot][j] = (gj >= 0) ? sc.indptr[gj] : -1;
  s_sg[slot][j] = sg;
  __syncwarp();
  const double mult = (sc.multipliers && cell_ok) ? __ldg(sc.multipliers + cell) : 1.;
  const int boff = cell_ok ? sc.m_ptr[cell] : 0;
  for (int i = 0; i < n; ++i) {
    const int rs = s_row[slot][i];
    if (valid && rs >= 0 && gj >= 0) {
      const double v = m_e[boff + (size_t)i * n + j];
      const unsigned pos = sc.pos_m[boff + (size_t)i * n + j];
      const double val = s_sg[slot][i] * sg * mult * v;
      if (i == j)
        atomicAdd(sc.vals + rs + pos, val);
      else
        sc.vals[rs + pos] = val;
    }
  }
  const int F = sc.flux_dof;
  if (valid && gj >= 0) {
    const double e = sg * area;
    sc.vals[s_row[slot][j] + sc.pos_dt[base + j]] = -e;
    sc.vals[sc.indptr[F + cell] + sc.pos_d[base + j]] = e;
  }
  if (j == 0 && cell_ok) {
    const double cd = sc.c_diag ? __ldg(sc.c_diag + cell) : sc.alpha * __ldg(m.cell_volume + cell);
    sc.vals[sc.indptr[F + cell + 1] - 1] = cd;
  }
}

__global__ void k_zero_diag(int F, const int* __restrict__ indptr,
                            const uint8_t* __restrict__ diag_pos, double* __restrict__ vals) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= F) return;
  vals[indptr[g] + diag_pos[g]] = 0.;
}

// COO view in the reference's order (cell-major, (i,j) row-major, Neumann skipped)
__global__ void k_coo_count(int nc, const int* __restrict__ cell_ptr,
                            const int* __restrict__ cell_faces, const int* __restrict__ f2f,
                            long long* __restrict__ cnt) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  int k = 0;
  for (int p = cell_ptr[c]; p < cell_ptr[c + 1]; ++p) k += f2f[cell_faces[p]] >= 0;
  cnt[c] = (long long)k * k;
}

__global__ void k_coo_fill(int nc, const int* __restrict__ cell_ptr,
                           const int* __restrict__ cell_faces, const int8_t* __restrict__ sigma,
                           const int* __restrict__ f2f, const int* __restrict__ m_ptr,
                           const double* __restrict__ m_e, const long long* __restrict__ coo_ptr,
                           int* __restrict__ rows, int* __restrict__ cols,
                           double* __restrict__ vals) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const int base = cell_ptr[c];
  const int n = cell_ptr[c + 1] - base;
  long long o = coo_ptr[c];
  for (int i = 0; i < n; ++i) {
    const int gi = f2f[cell_faces[base + i]];
    if (gi < 0) continue;
    for (int j = 0; j < n; ++j) {
      const int gj = f2f[cell_faces[base + j]];
      if (gj < 0) continue;
      if (rows) rows[o] = gi;
      if (cols) cols[o] = gj;
      if (vals) vals[o] = m_e[m_ptr[c] + (size_t)i * n + j] * (double)sigma[base + i] * (double)sigma[base + j];
      ++o;
    }
  }
}

#include <cub/device/device_scan.cuh>

static int check_mesh(const mimpy_mesh_view* mesh) {
  REQUIRE(mesh, "mesh is NULL");
  REQUIRE(mesh->cell_ptr && mesh->cell_faces && mesh->cell_sigma && mesh->face_geom &&
              mesh->cell_centroid && mesh->cell_volume,
          "mesh view has NULL arrays");
  REQUIRE(mesh->max_faces >= 1 && mesh->max_faces <= MIMPY_MAX_CELL_FACES,
          "cells with more than 32 faces are not supported");
  return MIMPY_OK;
}

static ScatterArgs make_scatter(const mimpy_symbolic* sym, const double* mult, double alpha,
                                const double* c_diag, double* vals) {
  ScatterArgs sc;
  sc.face_to_flux = sym->face_to_flux;
  sc.face_cell = sym->face_cell;
  sc.indptr = sym->indptr;
  sc.m_ptr = sym->m_ptr;
  sc.pos_m = sym->pos_m;
  sc.pos_dt = sym->pos_dt;
  sc.pos_d = sym->pos_d;
  sc.multipliers = mult;
  sc.c_diag = c_diag;
  sc.alpha = alpha;
  sc.vals = vals;
  sc.flux_dof = sym->flux_dof;
  return sc;
}

extern "C" {

int mimpy_b200_mfd_local_matrices(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, int32_t method,
                                  int32_t flags, const int32_t* m_ptr, double* m_e,
                                  double* diagonality, double* consistency) {
  REQUIRE(ctx, "ctx is NULL");
  int rc = check_mesh(mesh);
  if (rc) return rc;
  REQUIRE(method >= 0 && method <= 7, "unknown M_E construction method");
  REQUIRE((flags & 1) || mesh->cell_k, "cell_k is NULL");
  REQUIRE(m_ptr && (m_e || diagonality || consistency), "no output requested");
  if (mesh->nc <= 0) return MIMPY_OK;
  const int G = mimpy_group_size(mesh->max_faces);
  const int cpb = 128 / G;
  const int blocks = (mesh->nc + cpb - 1) / cpb;
  ScatterArgs sc = {};
  const size_t dyn = (method == 4) ? sizeof(double) * (size_t)cpb * G * (G + 1) : 0;
#define LAUNCH_M(GG, MM)                                                                   \
  k_local_matrices<GG, false, MM><<<blocks, 128, dyn, ctx->stream>>>(*mesh, method, flags, m_ptr, m_e, \
                                                                     diagonality, consistency, sc)
#define LAUNCH(GG)                       \
  do {                                   \
    if (method == 0) LAUNCH_M(GG, 0);    \
    else if (method == 1) LAUNCH_M(GG, 1); \
    else if (method == 6) LAUNCH_M(GG, 6); \
    else LAUNCH_M(GG, -1);               \
  } while (0)
  if (G == 8) LAUNCH(8);
  else if (G == 16) LAUNCH(16);
  else LAUNCH(32);
#undef LAUNCH
#undef LAUNCH_M
  KERNEL_CHECK();
  return MIMPY_OK;
}

int mimpy_b200_mfd_numeric(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym,
                           int32_t method, int32_t flags, const double* multipliers, double alpha,
                           const double* c_diag, double* vals, double* m_e_out) {
  REQUIRE(ctx && sym && vals, "NULL argument");
  int rc = check_mesh(mesh);
  if (rc) return rc;
  REQUIRE(method >= 0 && method <= 7, "unsupported M_E construction method");
  REQUIRE((flags & 1) || mesh->cell_k, "cell_k is NULL");
  if (mesh->nc <= 0) return MIMPY_OK;
  if (!m_e_out && !(flags & MIMPY_FLAG_GENERIC_KERNEL)) {
    // uniform-face-count fast path (hexes); same arithmetic, thread-per-cell mapping
    rc = mimpy_numeric_uniform(ctx, mesh, sym, method, flags, multipliers, alpha, c_diag, vals);
    if (rc != MIMPY_E_UNSUPPORTED) return rc;
  }
  const int F = sym->flux_dof;
  if (F > 0) {
    k_zero_diag<<<(F + 255) / 256, 256, 0, ctx->stream>>>(F, sym->indptr, sym->diag_pos, vals);
    KERNEL_CHECK();
  }
  const int G = mimpy_group_size(mesh->max_faces);
  const int cpb = 128 / G;
  const int blocks = (mesh->nc + cpb - 1) / cpb;
  ScatterArgs sc = make_scatter(sym, multipliers, alpha, c_diag, vals);
  const size_t dyn = (method == 4) ? sizeof(double) * (size_t)cpb * G * (G + 1) : 0;
#define LAUNCH_M(GG, MM)                                                                              \
  k_local_matrices<GG, true, MM><<<blocks, 128, dyn, ctx->stream>>>(*mesh, method, flags, sym->m_ptr, \
                                                                    m_e_out, nullptr, nullptr, sc)
#define LAUNCH(GG)                       \
  do {                                   \
    if (method == 0) LAUNCH_M(GG, 0);    \
    else if (method == 1) LAUNCH_M(GG, 1); \
    else if (method == 6) LAUNCH_M(GG, 6); \
    else LAUNCH_M(GG, -1);               \
  } while (0)
  if (G == 8) LAUNCH(8);
  else if (G == 16) LAUNCH(16);
  else LAUNCH(32);
#undef LAUNCH
#undef LAUNCH_M
  KERNEL_CHECK();
  return MIMPY_OK;
}

int mimpy_b200_mfd_numeric_layers(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym,
                                  int32_t method, int32_t flags, const double* multipliers, double alpha,
                                  const double* c_diag, double* vals, int32_t layer_begin, int32_t layer_end) {
  REQUIRE(ctx && sym && vals, "NULL argument");
  int rc = check_mesh(mesh);
  if (rc) return rc;
  REQUIRE((flags & 1) || mesh->cell_k, "cell_k is NULL");
  REQUIRE(layer_begin >= 0 && layer_end > layer_begin && layer_end <= mesh->hex_nz, "bad layer range");
  return mimpy_numeric_uniform(ctx, mesh, sym, method, flags, multipliers, alpha, c_diag, vals, layer_begin, layer_end);
}

int mimpy_b200_mfd_numeric_from_blocks(mimpy_ctx* ctx, const mimpy_mesh_view* mesh,
                                       const mimpy_symbolic* sym, const double* m_e,
                                       const double* multipliers, double alpha,
                                       const double* c_diag, double* vals) {
  REQUIRE(ctx && sym && vals && m_e, "NULL argument");
  int rc = check_mesh(mesh);
  if (rc) return rc;
  if (mesh->nc <= 0) return MIMPY_OK;
  const int F = sym->flux_dof;
  if (F > 0) {
    k_zero_diag<<<(F + 255) / 256, 256, 0, ctx->stream>>>(F, sym->indptr, sym->diag_pos, vals);
    KERNEL_CHECK();
  }
  const int G = mimpy_group_size(mesh->max_faces);
  const int cpb = 128 / G;
  const int blocks = (mesh->nc + cpb - 1) / cpb;
  ScatterArgs sc = make_scatter(sym, multipliers, alpha, c_diag, vals);
  if (G == 8) k_scatter_blocks<8><<<blocks, 128, 0, ctx->stream>>>(*mesh, m_e, sc);
  else if (G == 16) k_scatter_blocks<16><<<blocks, 128, 0, ctx->stream>>>(*mesh, m_e, sc);
  else k_scatter_blocks<32><<<blocks, 128, 0, ctx->stream>>>(*mesh, m_e, sc);
  KERNEL_CHECK();
  return MIMPY_OK;
}

int mimpy_b200_mfd_coo(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym,
                       const double* m_e, int64_t* coo_ptr, int32_t* rows, int32_t* cols,
                       double* vals) {
  REQUIRE(ctx && mesh && sym && coo_ptr, "NULL argument");
  REQUIRE(!vals || m_e, "vals requested without blocks");
  const int nc = mesh->nc;
  if (nc <= 0) return MIMPY_OK;
  const int T = 256;
  long long* cnt = nullptr;
  CUDA_TRY(cudaMalloc(&cnt, sizeof(long long) * (size_t)(nc + 1)));
  cudaMemsetAsync(cnt, 0, sizeof(long long) * (size_t)(nc + 1), ctx->stream);
  k_coo_count<<<(nc + T - 1) / T, T, 0, ctx->stream>>>(nc, mesh->cell_ptr, mesh->cell_faces,
                                                       sym->face_to_flux, cnt);
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, cnt, (long long*)coo_ptr, nc + 1, ctx->stream);
  void* tmp = nullptr;
  int rc = mimpy_scratch(ctx, bytes, &tmp);
  if (rc == MIMPY_OK) {
    cub::DeviceScan::ExclusiveSum(tmp, bytes, cnt, (long long*)coo_ptr, nc + 1, ctx->stream);
    if (rows || cols || vals)
      k_coo_fill<<<(nc + T - 1) / T, T, 0, ctx->stream>>>(
          nc, mesh->cell_ptr, mesh->cell_faces, mesh->cell_sigma, sym->face_to_flux, sym->m_ptr,
          m_e, (const long long*)coo_ptr, rows, cols, vals);
  }
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cudaFree(cnt);
  if (rc) return rc;
  if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) {
    mimpy_set_error("mfd_coo: %s", cudaGetErrorString(e));
    return MIMPY_E_CUDA;
  }
  return MIMPY_OK;
}

}  // extern "C"
