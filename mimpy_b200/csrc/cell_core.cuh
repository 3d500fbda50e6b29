// Register-resident M_E pipeline of one hexahedral (NF-face) cell, shared by the fused numeric
// assembly (assemble_uniform.cu) and the fused set-up of the hybridised face system (hybrid.cu).
// Restates MFD.build_m_e (mfd.py:357-483, methods 0, 1, 6): no redundant work across lanes -- the 3x3
// inverse, the two Cholesky factorisations and their rsqrts are done once per cell.  Orientations are
// folded into the operands:  R~_i = s_i a_i (x_f - x_E),  KN_i = K n_f, so that
//   s_i s_j M_E[i,j] = R~_i . (T^-1 R~_j) + coef_i (d_ij - Q~_i . Q~_j)
// (T = R~^T KN, Q~ = orth(KN) by CholeskyQR2; mfd.py:575-588 applies s_i s_j afterwards).
#pragma once
#include "common.cuh"

// Cholesky with reciprocal square roots (one MUFU-seeded rsqrt per pivot instead of sqrt + div)
__device__ __forceinline__ Chol3 chol3r(double g00, double g10, double g11, double g20, double g21,
                                        double g22) {
  Chol3 c;
  c.i0 = rsqrt(g00);
  c.l10 = g10 * c.i0;
  c.l20 = g20 * c.i0;
  c.i1 = rsqrt(g11 - c.l10 * c.l10);
  c.l21 = (g21 - c.l20 * c.l10) * c.i1;
  c.i2 = rsqrt(g22 - c.l20 * c.l20 - c.l21 * c.l21);
  return c;
}

// Per-cell operands: R~_i = s_i a_i (x_f - x_E), KN_i = K n_f, d_f of method 6.
template <int NF>
struct CellOps {
  double rt[NF][3], kn[NF][3], dcoef[NF];
};

// rec[i] = the packed face record [a, n, x_f, 0]; sgn[i] = orientation.
template <int NF, int METHOD>
__device__ __forceinline__ void cell_prep(const double (&K)[9], const double (&xc)[3],
                                          const double2 (&rec)[NF][4], const double (&sgn)[NF],
                                          CellOps<NF>& o) {
  double Ki[9];
  if (METHOD == 6) inv3(K, Ki);
#pragma unroll
  for (int i = 0; i < NF; ++i) {
    const double area = rec[i][0].x;
    const double nx = rec[i][0].y, ny = rec[i][1].x, nz = rec[i][1].y;
    const double dx = rec[i][2].x - xc[0], dy = rec[i][2].y - xc[1], dz = rec[i][3].x - xc[2];
    const double sa = sgn[i] * area;
    o.rt[i][0] = sa * dx;
    o.rt[i][1] = sa * dy;
    o.rt[i][2] = sa * dz;
    o.kn[i][0] = K[0] * nx + K[1] * ny + K[2] * nz;
    o.kn[i][1] = K[3] * nx + K[4] * ny + K[5] * nz;
    o.kn[i][2] = K[6] * nx + K[7] * ny + K[8] * nz;
    if (METHOD == 6) {  // d_f = n^T K^-1 n a_f |x_f - x_E|            mfd.py:446-453
      const double e = nx * (Ki[0] * nx + Ki[1] * ny + Ki[2] * nz) +
                       ny * (Ki[3] * nx + Ki[4] * ny + Ki[5] * nz) +
                       nz * (Ki[6] * nx + Ki[7] * ny + Ki[8] * nz);
      o.dcoef[i] = e * area * sqrt(dx * dx + dy * dy + dz * dz);
    }
  }
}

// Calls sink(i, j, v) with v = mult * s_i s_j M_E[i,j] for all NF^2 entries, row by row.
// trk = trace(K) (method 1 only).
template <int NF, int METHOD, class Sink>
__device__ __forceinline__ void cell_core(CellOps<NF>& o, double trk, double vol, double mult,
                                          Sink&& sink) {
  double (&rt)[NF][3] = o.rt;
  double (&kn)[NF][3] = o.kn;
  double T[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double g00 = 0, g10 = 0, g11 = 0, g20 = 0, g21 = 0, g22 = 0;
  double rr0 = 0, rr1 = 0, rr2 = 0;
#pragma unroll
  for (int i = 0; i < NF; ++i) {
    T[0] += rt[i][0] * kn[i][0]; T[1] += rt[i][0] * kn[i][1]; T[2] += rt[i][0] * kn[i][2];
    T[3] += rt[i][1] * kn[i][0]; T[4] += rt[i][1] * kn[i][1]; T[5] += rt[i][1] * kn[i][2];
    T[6] += rt[i][2] * kn[i][0]; T[7] += rt[i][2] * kn[i][1]; T[8] += rt[i][2] * kn[i][2];
    g00 += kn[i][0] * kn[i][0]; g10 += kn[i][1] * kn[i][0]; g11 += kn[i][1] * kn[i][1];
    g20 += kn[i][2] * kn[i][0]; g21 += kn[i][2] * kn[i][1]; g22 += kn[i][2] * kn[i][2];
    if (METHOD == 0) {
      rr0 += rt[i][0] * rt[i][0]; rr1 += rt[i][0] * rt[i][1]; rr2 += rt[i][0] * rt[i][2];
    }
  }
  double Ti[9];
  inv3(T, Ti);
  double u = 0.;
  if (METHOD == 0) u = rr0 * Ti[0] + rr1 * Ti[3] + rr2 * Ti[6];      // mfd.py:380-382
  if (METHOD == 1) u = vol / trk;                                     // mfd.py:387
  u *= mult;
  {
    const Chol3 c1 = chol3r(g00, g10, g11, g20, g21, g22);
    double h00 = 0, h10 = 0, h11 = 0, h20 = 0, h21 = 0, h22 = 0;
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      chol3_fwd(c1, kn[i][0], kn[i][1], kn[i][2]);
      h00 += kn[i][0] * kn[i][0]; h10 += kn[i][1] * kn[i][0]; h11 += kn[i][1] * kn[i][1];
      h20 += kn[i][2] * kn[i][0]; h21 += kn[i][2] * kn[i][1]; h22 += kn[i][2] * kn[i][2];
    }
    const Chol3 c2 = chol3r(h00, h10, h11, h20, h21, h22);
#pragma unroll
    for (int i = 0; i < NF; ++i) chol3_fwd(c2, kn[i][0], kn[i][1], kn[i][2]);
  }
#pragma unroll
  for (int i = 0; i < NF; ++i) {
    // row i: R~_i^T (mult T^-1) R~_j  ->  z_i = mult T^-T R~_i, then z_i . R~_j
    double z[3];
#pragma unroll
    for (int q = 0; q < 3; ++q)
      z[q] = mult * (Ti[q] * rt[i][0] + Ti[3 + q] * rt[i][1] + Ti[6 + q] * rt[i][2]);
    const double coef = (METHOD == 6) ? mult * o.dcoef[i] : u;
#pragma unroll
    for (int j = 0; j < NF; ++j) {
      // explicit fma chains: every kernel that instantiates this core rounds identically whatever its
      // sink looks like (the contraction of a*b+c otherwise depends on the surrounding control flow)
      const double m0 = fma(z[2], rt[j][2], fma(z[1], rt[j][1], z[0] * rt[j][0]));
      const double qq = fma(kn[i][2], kn[j][2], fma(kn[i][1], kn[j][1], kn[i][0] * kn[j][0]));
      sink(i, j, fma(coef, ((i == j) ? 1. : 0.) - qq, m0));
    }
  }
}

template <int NF, int METHOD, class Sink>
__device__ __forceinline__ void cell_entries(const double (&K)[9], const double (&xc)[3], double vol,
                                             double mult, const double2 (&rec)[NF][4],
                                             const double (&sgn)[NF], Sink&& sink) {
  CellOps<NF> o;
  cell_prep<NF, METHOD>(K, xc, rec, sgn, o);
  cell_core<NF, METHOD>(o, K[0] + K[4] + K[8], vol, mult, sink);
}

__device__ __forceinline__ unsigned byte_of(const unsigned short* w, int i) {
  return (w[i >> 1] >> ((i & 1) * 8)) & 0xFFu;
}

