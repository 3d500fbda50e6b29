// Fused numeric assembly for meshes whose cells all have NF faces (hexes: NF = 6) -- the
// headline kernel (700 algorithmic bytes per hex cell).
//
// One CTA of 128 threads owns 128 consecutive cells and runs three stages:
//
//  S  staging: every contiguous input of the CTA (K, centroids, volumes, face ids, orientations,
//     the pos_* byte maps, the cell-row starts) is fetched with coalesced, fully unrolled loads
//     into shared memory -- ~40 independent loads in flight per thread, one DRAM latency.
//  A  thread-per-cell: the whole 3xNF / 3x3 pipeline of MFD.build_m_e (mfd.py:357-483, methods 0,
//     1, 6) in registers -- no redundant work across lanes (the 3x3 inverse, the two Cholesky
//     factorisations and their divisions/rsqrts are done once per cell, not once per face).
//     Orientations are folded into the operands:  R~_i = s_i a_i (x_f - x_E),  KN_i = K n_f, so
//     that  s_i s_j M_E[i,j] = R~_i . (T^-1 R~_j) + coef_i (d_ij - Q~_i . Q~_j)  directly
//     (T = R~^T KN, Q~ = orth(KN) by CholeskyQR2; mfd.py:575-588 applies s_i s_j afterwards).
//     The NF^2 signed, scaled entries go to shared memory.
//  B  entry-per-thread: the CTA streams its 128*NF^2 entries out of shared memory; consecutive
//     lanes hold consecutive (i,j) of one cell, i.e. contiguous pieces of CSR row g(f_i).  No
//     global loads are left in this stage (positions and row starts come from shared memory).
//
// Diagonal entries M[g,g] have two contributors (the two cells of face f).  Instead of atomics on
// a pre-zeroed array, the lower-index cell PUBLISHES its contribution in diag_pub[f] and the
// higher-index cell adds it and writes the slot -- a + b in a fixed order, bit-reproducible.  CTAs
// take their cell range from an atomic ticket, so a CTA only ever waits for CTAs that started
// earlier and are resident (the ordering argument of decoupled look-back scans): no deadlock.
#include <stdlib.h>

#include "common.cuh"

#define AU_CELLS 128
#define AU_SENTINEL 0xFFFFFFFFFFFFFFFFull  // memset(0xFF): a NaN no arithmetic produces

struct UniArgs {
  const int* face_to_flux;
  const int* indptr;
  const uint8_t* pos_m;
  const uint8_t* pos_dt;
  const uint8_t* pos_d;
  const int* cell_rowstart;
  const uint8_t* role;  // per (cell, face): 0 sole cell, 1 lower cell (publish), 2 higher (finish)
  const double* multipliers;
  const double* c_diag;
  double alpha;
  double* vals;
  double* diag_pub;
  unsigned int* ticket;
  unsigned int* error_flag;
  int flux_dof;
};

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

template <int NF>
struct UniSmem {
  // phase-A inputs (dead once every thread has read its cell) alias the entry buffer
  union {
    struct {
      double k[AU_CELLS * 9];
      double xc[AU_CELLS * 3];
      double vol[AU_CELLS];
    } in;
    double entries[AU_CELLS * (NF * NF + 1)];  // odd per-cell stride: conflict-free 64-bit stores
  } u;
  int ids[AU_CELLS * NF];       // global face ids
  int rowstart[AU_CELLS * NF];  // indptr[g(f_i)] or -1 (cell_rowstart of the symbolic plan)
  int cellrow[AU_CELLS + 1];    // indptr[F + c]
  unsigned int ticket;
  alignas(16) uint8_t pos_m[AU_CELLS * NF * NF];
  alignas(16) uint8_t pos_dt[AU_CELLS * NF];
  alignas(16) uint8_t pos_d[AU_CELLS * NF];
  alignas(16) uint8_t role[AU_CELLS * NF];
  alignas(16) int8_t sg[AU_CELLS * NF];
};

// COUNT elements (multiple of 128 per pass) from src[first ...] into shared memory: all loads are
// issued before the first store (PER independent requests in flight per thread)
template <typename T, int COUNT>
__device__ __forceinline__ void stage(T* dst, const T* __restrict__ src, long long first, long long limit) {
  constexpr int PER = (COUNT + AU_CELLS - 1) / AU_CELLS;
  T tmp[PER];
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int o = k * AU_CELLS + threadIdx.x;
    const long long g = first + o;
    tmp[k] = (o < COUNT && g < limit) ? src[g] : T(0);
  }
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int o = k * AU_CELLS + threadIdx.x;
    if (o < COUNT) dst[o] = tmp[k];
  }
}

// byte maps as 32-bit words (the CTA's slice starts at a multiple of 128*NF bytes)
template <int BYTES>
__device__ __forceinline__ void stage_bytes(uint8_t* dst, const uint8_t* __restrict__ src, long long first,
                                            long long limit) {
  static_assert(BYTES % 4 == 0, "slice must be a whole number of words");
  if (first + BYTES <= limit && ((reinterpret_cast<uintptr_t>(src) + first) & 3) == 0) {
    stage<unsigned int, BYTES / 4>(reinterpret_cast<unsigned int*>(dst),
                                   reinterpret_cast<const unsigned int*>(src + first), 0, BYTES / 4);
  } else {
    stage<uint8_t, BYTES>(dst, src, first, limit);
  }
}

template <int NF, int METHOD>
__global__ void __launch_bounds__(AU_CELLS, 4)
k_numeric_uniform(mimpy_mesh_view m, int flags, UniArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  UniSmem<NF>& s = *reinterpret_cast<UniSmem<NF>*>(smem_raw);
  const int t = threadIdx.x;
  if (t == 0) s.ticket = atomicAdd(a.ticket, 1u);
  __syncthreads();
  const long long c0 = (long long)s.ticket * AU_CELLS;
  const int nc = m.nc;
  const int F = a.flux_dof;
  const bool k_unity = flags & 1;
  constexpr int NN = NF * NF;
  constexpr int ES = NN + 1;
  const long long cell = c0 + t;
  const bool ok = cell < nc;

  // ---- S0: face ids first, so that the record gathers can start as early as possible ---------
  stage<int, AU_CELLS * NF>(s.ids, m.cell_faces, c0 * NF, (long long)nc * NF);
  __syncthreads();
  int fid[NF];
  double2 rec[NF][4];
#pragma unroll
  for (int i = 0; i < NF; ++i) fid[i] = s.ids[t * NF + i];
#pragma unroll
  for (int i = 0; i < NF; ++i) {
    const double2* p = reinterpret_cast<const double2*>(m.face_geom + 8 * (size_t)(ok ? fid[i] : 0));
#pragma unroll
    for (int q = 0; q < 4; ++q) rec[i][q] = __ldg(p + q);
  }
  // ---- S1: the CTA's contiguous inputs, in flight together with the gathers ---------------------
  if (!k_unity) stage<double, AU_CELLS * 9>(s.u.in.k, m.cell_k, c0 * 9, (long long)nc * 9);
  stage<double, AU_CELLS * 3>(s.u.in.xc, m.cell_centroid, c0 * 3, (long long)nc * 3);
  stage<double, AU_CELLS>(s.u.in.vol, m.cell_volume, c0, nc);
  stage<int, AU_CELLS * NF>(s.rowstart, a.cell_rowstart, c0 * NF, (long long)nc * NF);
  stage_bytes<AU_CELLS * NF>(reinterpret_cast<uint8_t*>(s.sg), reinterpret_cast<const uint8_t*>(m.cell_sigma),
                             c0 * NF, (long long)nc * NF);
  stage_bytes<AU_CELLS * NN>(s.pos_m, a.pos_m, c0 * NN, (long long)nc * NN);
  stage_bytes<AU_CELLS * NF>(s.pos_dt, a.pos_dt, c0 * NF, (long long)nc * NF);
  stage_bytes<AU_CELLS * NF>(s.pos_d, a.pos_d, c0 * NF, (long long)nc * NF);
  stage_bytes<AU_CELLS * NF>(s.role, a.role, c0 * NF, (long long)nc * NF);
  for (int k = t; k <= AU_CELLS; k += AU_CELLS) {
    const long long c = c0 + k;
    s.cellrow[k] = (c <= nc) ? a.indptr[F + c] : 0;
  }
  const double mult = (a.multipliers && ok) ? __ldg(a.multipliers + cell) : 1.;
  const double cdiag = (a.c_diag && ok) ? __ldg(a.c_diag + cell) : 0.;
  __syncthreads();

  double K[9], xc[3], vol;
  double sg[NF];
  if (k_unity) {
#pragma unroll
    for (int q = 0; q < 9; ++q) K[q] = (q % 4 == 0) ? 1. : 0.;
  } else {
#pragma unroll
    for (int q = 0; q < 9; ++q) K[q] = s.u.in.k[t * 9 + q];
  }
#pragma unroll
  for (int q = 0; q < 3; ++q) xc[q] = s.u.in.xc[t * 3 + q];
  vol = s.u.in.vol[t];
#pragma unroll
  for (int i = 0; i < NF; ++i) sg[i] = (double)s.sg[t * NF + i];
  __syncthreads();  // inputs consumed: the entry buffer may now be overwritten

  if (ok) {
    double rt[NF][3], kn[NF][3];
    double dcoef[NF];  // method 6
    double Ki[9];
    if (METHOD == 6) inv3(K, Ki);
    const int crow = s.cellrow[t];
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      const double area = rec[i][0].x;
      const double nx = rec[i][0].y, ny = rec[i][1].x, nz = rec[i][1].y;
      const double dx = rec[i][2].x - xc[0], dy = rec[i][2].y - xc[1], dz = rec[i][3].x - xc[2];
      const double sa = sg[i] * area;
      const int rs = s.rowstart[t * NF + i];
      if (rs >= 0) {  // DIV / -DIV^T                                    mfd.py:232-243
        a.vals[rs + s.pos_dt[t * NF + i]] = -sa;
        a.vals[crow + s.pos_d[t * NF + i]] = sa;
      }
      rt[i][0] = sa * dx;
      rt[i][1] = sa * dy;
      rt[i][2] = sa * dz;
      kn[i][0] = K[0] * nx + K[1] * ny + K[2] * nz;
      kn[i][1] = K[3] * nx + K[4] * ny + K[5] * nz;
      kn[i][2] = K[6] * nx + K[7] * ny + K[8] * nz;
      if (METHOD == 6) {  // d_f = n^T K^-1 n a_f |x_f - x_E|            mfd.py:446-453
        const double e = nx * (Ki[0] * nx + Ki[1] * ny + Ki[2] * nz) +
                         ny * (Ki[3] * nx + Ki[4] * ny + Ki[5] * nz) +
                         nz * (Ki[6] * nx + Ki[7] * ny + Ki[8] * nz);
        dcoef[i] = e * area * sqrt(dx * dx + dy * dy + dz * dz);
      }
    }
    // C diagonal                                                        mfd.py:766-772
    a.vals[s.cellrow[t + 1] - 1] = a.c_diag ? cdiag : a.alpha * vol;
    // ---- T = R~^T KN, G = KN^T KN ------------------------------------------------------------------
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
    if (METHOD == 1) u = vol / (K[0] + K[4] + K[8]);                    // mfd.py:387
    u *= mult;
    // ---- Q~ = orth(KN): CholeskyQR2 -------------------------------------------------------------------
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
    // ---- entries: row i uses R~_i, Q~_i ; column j uses Y~_j = mult T^-1 R~_j, Q~_j ---------------
    double* ent = s.u.entries + t * ES;
#pragma unroll
    for (int j = 0; j < NF; ++j) {
      double y[3];
#pragma unroll
      for (int q = 0; q < 3; ++q)
        y[q] = mult * (Ti[3 * q] * rt[j][0] + Ti[3 * q + 1] * rt[j][1] + Ti[3 * q + 2] * rt[j][2]);
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        const double coef = (METHOD == 6) ? mult * dcoef[i] : u;
        const double m0 = rt[i][0] * y[0] + rt[i][1] * y[1] + rt[i][2] * y[2];
        const double qq = kn[i][0] * kn[j][0] + kn[i][1] * kn[j][1] + kn[i][2] * kn[j][2];
        const double v = m0 + coef * (((i == j) ? 1. : 0.) - qq);
        ent[i * NF + j] = v;
        // lower-index cell of a shared face publishes its diagonal contribution right away
        if (i == j && s.role[t * NF + i] == 1 && s.rowstart[t * NF + i] >= 0) a.diag_pub[fid[i]] = v;
      }
    }
  }
  __syncthreads();

  // ---- B: stream the off-diagonal M entries (no global loads) -------------------------------------
  const long long elimit = (long long)nc * NN - c0 * NN;
#pragma unroll 6
  for (int k = 0; k < NN; ++k) {
    const int e = k * AU_CELLS + t;
    if (e >= elimit) break;
    const int cl = e / NN;
    const int ij = e - cl * NN;
    const int i = ij / NF;
    const int j = ij - i * NF;
    const unsigned pos = s.pos_m[e];
    const int rs = s.rowstart[cl * NF + i];
    if (i != j && pos != 0xFFu && rs >= 0) a.vals[rs + pos] = s.u.entries[cl * ES + ij];
  }
  // ---- diagonals: the finishing cell adds the half published by the lower-index cell ---------------
  // first an optimistic, non-blocking read of all NF candidates (independent loads), then spin on
  // the few that were not there yet
  unsigned long long pub[NF];
#pragma unroll
  for (int k = 0; k < NF; ++k) {
    const int e = k * AU_CELLS + t;
    pub[k] = 0ull;
    if (c0 + e / NF < nc && s.role[e] == 2 && s.rowstart[e] >= 0) {
      const unsigned long long* src = reinterpret_cast<const unsigned long long*>(a.diag_pub + s.ids[e]);
      asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(pub[k]) : "l"(src));
    }
  }
#pragma unroll
  for (int k = 0; k < NF; ++k) {
    const int e = k * AU_CELLS + t;
    const int cl = e / NF;
    const int i = e - cl * NF;
    if (c0 + cl >= nc) break;
    const int rs = s.rowstart[e];
    const unsigned role = s.role[e];
    if (rs < 0 || role == 1) continue;  // Neumann, or the other cell finishes this diagonal
    double d = s.u.entries[cl * ES + i * NF + i];
    if (role == 2) {
      unsigned long long bits = pub[k];
      if (bits == AU_SENTINEL) {
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(a.diag_pub + s.ids[e]);
        unsigned int spins = 0;
        do {
          __nanosleep(64);
          asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(bits) : "l"(src));
        } while (bits == AU_SENTINEL && ++spins < (1u << 22));
        if (bits == AU_SENTINEL) atomicExch(a.error_flag, 1u);
      }
      d = __longlong_as_double((long long)bits) + d;  // lower cell first: fixed order
    }
    a.vals[rs + s.pos_m[cl * NN + i * NF + i]] = d;
  }
}

// ---------------------------------------------------------------------------------------------
// Variant "direct": thread-per-cell end to end.  No shared-memory staging or transposition: every
// input address is a function of the cell index alone, so all loads of a thread are issued up
// front (two dependent levels: face ids -> face records), and each entry is stored straight from
// registers to its CSR slot.  Far fewer instructions and barriers, higher occupancy; costs ~30 %
// more L2 write-sector slots than the row-cooperative stores of the staged variant.
// ---------------------------------------------------------------------------------------------
#define AD_THREADS 128

template <int NF, int METHOD>
#ifndef AD_MINB
#define AD_MINB 2  /* measured on B200: 2 -> 7.0 ms, 3 -> 7.5 ms, 4 -> 9.5 ms at 256^3 (spills) */
#endif
__global__ void __launch_bounds__(AD_THREADS, AD_MINB)
k_numeric_direct(mimpy_mesh_view m, int flags, UniArgs a) {
  __shared__ unsigned int s_ticket;
  const int t = threadIdx.x;
  if (t == 0) s_ticket = atomicAdd(a.ticket, 1u);
  __syncthreads();
  const long long cell = (long long)s_ticket * AD_THREADS + t;
  if (cell >= m.nc) return;
  const int F = a.flux_dof;
  const bool k_unity = flags & 1;
  const bool dbg_no_mstore = flags & 256, dbg_no_gather = flags & 512, dbg_no_wait = flags & 1024;
  constexpr int NN = NF * NF;

  // ---- level-1 loads (addresses depend on the cell index only) ---------------------------------
  int fid[NF], rs[NF];
  {
    const int2* p = reinterpret_cast<const int2*>(m.cell_faces + cell * NF);
    const int2* q = reinterpret_cast<const int2*>(a.cell_rowstart + cell * NF);
#pragma unroll
    for (int i = 0; i < NF / 2; ++i) {
      const int2 v = __ldg(p + i);
      fid[2 * i] = v.x;
      fid[2 * i + 1] = v.y;
      const int2 w = __ldg(q + i);
      rs[2 * i] = w.x;
      rs[2 * i + 1] = w.y;
    }
  }
  double K[9];
  if (k_unity) {
#pragma unroll
    for (int q = 0; q < 9; ++q) K[q] = (q % 4 == 0) ? 1. : 0.;
  } else {
#pragma unroll
    for (int q = 0; q < 9; ++q) K[q] = __ldg(m.cell_k + cell * 9 + q);
  }
  double xc[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) xc[q] = __ldg(m.cell_centroid + cell * 3 + q);
  const double vol = __ldg(m.cell_volume + cell);
  const double mult = a.multipliers ? __ldg(a.multipliers + cell) : 1.;
  const int crow = __ldg(a.indptr + F + cell);
  const int crow_end = __ldg(a.indptr + F + cell + 1);
  unsigned int pm[NN / 4];
  {
    const unsigned int* p = reinterpret_cast<const unsigned int*>(a.pos_m + cell * NN);
#pragma unroll
    for (int q = 0; q < NN / 4; ++q) pm[q] = __ldg(p + q);
  }
  // 6-byte maps: three 16-bit words each (cell*6 is even)
  unsigned short w_sg[NF / 2], w_dt[NF / 2], w_d[NF / 2], w_role[NF / 2];
  {
    const unsigned short* p0 = reinterpret_cast<const unsigned short*>(m.cell_sigma + cell * NF);
    const unsigned short* p1 = reinterpret_cast<const unsigned short*>(a.pos_dt + cell * NF);
    const unsigned short* p2 = reinterpret_cast<const unsigned short*>(a.pos_d + cell * NF);
    const unsigned short* p3 = reinterpret_cast<const unsigned short*>(a.role + cell * NF);
#pragma unroll
    for (int q = 0; q < NF / 2; ++q) {
      w_sg[q] = __ldg(p0 + q);
      w_dt[q] = __ldg(p1 + q);
      w_d[q] = __ldg(p2 + q);
      w_role[q] = __ldg(p3 + q);
    }
  }
  // ---- level-2 loads: face records; optimistic read of the published diagonal halves ------------
  double2 rec[NF][4];
#pragma unroll
  for (int i = 0; i < NF; ++i) {
    const double2* p = reinterpret_cast<const double2*>(m.face_geom + 8 * (size_t)(dbg_no_gather ? (cell % 1024) * 3 + i : fid[i]));
#pragma unroll
    for (int q = 0; q < 4; ++q) rec[i][q] = __ldg(p + q);
  }
  auto byte_of = [](const unsigned short* w, int i) -> unsigned { return (w[i >> 1] >> ((i & 1) * 8)) & 0xFFu; };
  unsigned long long pub[NF];
#pragma unroll
  for (int i = 0; i < NF; ++i) {
    pub[i] = 0ull;
    if (byte_of(w_role, i) == 2 && rs[i] >= 0 && !dbg_no_wait) {
      const unsigned long long* src = reinterpret_cast<const unsigned long long*>(a.diag_pub + fid[i]);
      asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(pub[i]) : "l"(src));
    }
  }

  // ---- geometry -> R~, KN ; DIV / -DIV^T / C stores -------------------------------------------------
  double rt[NF][3], kn[NF][3];
  double dcoef[NF];
  double Ki[9];
  if (METHOD == 6) inv3(K, Ki);
#pragma unroll
  for (int i = 0; i < NF; ++i) {
    const double area = rec[i][0].x;
    const double nx = rec[i][0].y, ny = rec[i][1].x, nz = rec[i][1].y;
    const double dx = rec[i][2].x - xc[0], dy = rec[i][2].y - xc[1], dz = rec[i][3].x - xc[2];
    const double sgn = (double)(signed char)byte_of(w_sg, i);
    const double sa = sgn * area;
    if (rs[i] >= 0) {  // mfd.py:232-243
      a.vals[rs[i] + byte_of(w_dt, i)] = -sa;
      a.vals[crow + byte_of(w_d, i)] = sa;
    }
    rt[i][0] = sa * dx;
    rt[i][1] = sa * dy;
    rt[i][2] = sa * dz;
    kn[i][0] = K[0] * nx + K[1] * ny + K[2] * nz;
    kn[i][1] = K[3] * nx + K[4] * ny + K[5] * nz;
    kn[i][2] = K[6] * nx + K[7] * ny + K[8] * nz;
    if (METHOD == 6) {
      const double e = nx * (Ki[0] * nx + Ki[1] * ny + Ki[2] * nz) +
                       ny * (Ki[3] * nx + Ki[4] * ny + Ki[5] * nz) +
                       nz * (Ki[6] * nx + Ki[7] * ny + Ki[8] * nz);
      dcoef[i] = e * area * sqrt(dx * dx + dy * dy + dz * dz);
    }
  }
  a.vals[crow_end - 1] = a.c_diag ? __ldg(a.c_diag + cell) : a.alpha * vol;  // mfd.py:766-772

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
  if (METHOD == 0) u = rr0 * Ti[0] + rr1 * Ti[3] + rr2 * Ti[6];
  if (METHOD == 1) u = vol / (K[0] + K[4] + K[8]);
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
  // ---- entries, row by row, stored from registers ----------------------------------------------------
  double diag[NF];
#pragma unroll
  for (int i = 0; i < NF; ++i) {
    // row i: R~_i^T (mult T^-1) R~_j  -> z_i = mult T^-T R~_i, then z_i . R~_j
    double z[3];
#pragma unroll
    for (int q = 0; q < 3; ++q)
      z[q] = mult * (Ti[q] * rt[i][0] + Ti[3 + q] * rt[i][1] + Ti[6 + q] * rt[i][2]);
    const double coef = (METHOD == 6) ? mult * dcoef[i] : u;
#pragma unroll
    for (int j = 0; j < NF; ++j) {
      const double m0 = z[0] * rt[j][0] + z[1] * rt[j][1] + z[2] * rt[j][2];
      const double qq = kn[i][0] * kn[j][0] + kn[i][1] * kn[j][1] + kn[i][2] * kn[j][2];
      const double v = m0 + coef * (((i == j) ? 1. : 0.) - qq);
      if (i == j) {
        diag[i] = v;
      } else {
        const unsigned pos = (pm[(i * NF + j) >> 2] >> (((i * NF + j) & 3) * 8)) & 0xFFu;
        if (rs[i] >= 0 && pos != 0xFFu && !dbg_no_mstore) a.vals[rs[i] + pos] = v;
      }
    }
    // the lower-index cell of a shared face publishes its diagonal half as soon as it exists
    if (rs[i] >= 0 && byte_of(w_role, i) == 1) a.diag_pub[fid[i]] = diag[i];
  }
  // ---- diagonals ---------------------------------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < NF; ++i) {
    const unsigned role = byte_of(w_role, i);
    if (rs[i] < 0 || role == 1) continue;
    double d = diag[i];
    if (role == 2 && !dbg_no_wait) {
      unsigned long long bits = pub[i];
      if (bits == AU_SENTINEL) {
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(a.diag_pub + fid[i]);
        unsigned int spins = 0;
        do {
          __nanosleep(64);
          asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(bits) : "l"(src));
        } while (bits == AU_SENTINEL && ++spins < (1u << 22));
        if (bits == AU_SENTINEL) atomicExch(a.error_flag, 1u);
      }
      d = __longlong_as_double((long long)bits) + d;
    }
    const unsigned pos = (pm[(i * NF + i) >> 2] >> (((i * NF + i) & 3) * 8)) & 0xFFu;
    a.vals[rs[i] + pos] = d;
  }
}

static int asm_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MIMPY_B200_ASM_VARIANT");
    v = (e && e[0] == 's') ? 1 : 0;  // "staged" selects the shared-memory variant, default direct
  }
  return v;
}

template <int NF, int METHOD>
static int launch_uniform(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, int flags, const UniArgs& a) {
  if (asm_variant() == 0) {
    const int blocks = (mesh->nc + AD_THREADS - 1) / AD_THREADS;
    k_numeric_direct<NF, METHOD><<<blocks, AD_THREADS, 0, ctx->stream>>>(*mesh, flags, a);
    KERNEL_CHECK();
    return MIMPY_OK;
  }
  const size_t smem = sizeof(UniSmem<NF>);
  static bool configured = false;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(k_numeric_uniform<NF, METHOD>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int blocks = (mesh->nc + AU_CELLS - 1) / AU_CELLS;
  k_numeric_uniform<NF, METHOD><<<blocks, AU_CELLS, smem, ctx->stream>>>(*mesh, flags, a);
  KERNEL_CHECK();
  return MIMPY_OK;
}

// returns MIMPY_E_UNSUPPORTED when the fast path does not apply (caller falls back to the
// generic warp-group kernel of local_matrices.cu -- same results, CUDA as well)
int mimpy_numeric_uniform(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym,
                          int method, int flags, const double* multipliers, double alpha,
                          const double* c_diag, double* vals) {
  if (mesh->uniform_faces != 6 || !sym->role || !sym->cell_rowstart) return MIMPY_E_UNSUPPORTED;
  if (method != 0 && method != 1 && method != 6) return MIMPY_E_UNSUPPORTED;
  void* scratch = nullptr;
  int rc = mimpy_scratch(ctx, sizeof(double) * (size_t)mesh->nf, &scratch);
  if (rc) return rc;
  CUDA_TRY(cudaMemsetAsync(scratch, 0xFF, sizeof(double) * (size_t)mesh->nf, ctx->stream));
  CUDA_TRY(cudaMemsetAsync(ctx->counters + 8, 0, 2 * sizeof(unsigned int), ctx->stream));
  UniArgs a;
  a.face_to_flux = sym->face_to_flux;
  a.indptr = sym->indptr;
  a.pos_m = sym->pos_m;
  a.pos_dt = sym->pos_dt;
  a.pos_d = sym->pos_d;
  a.role = sym->role;
  a.cell_rowstart = sym->cell_rowstart;
  a.multipliers = multipliers;
  a.c_diag = c_diag;
  a.alpha = alpha;
  a.vals = vals;
  a.diag_pub = (double*)scratch;
  a.ticket = ctx->counters + 8;
  a.error_flag = ctx->counters + 9;
  a.flux_dof = sym->flux_dof;
  if (method == 0) return launch_uniform<6, 0>(ctx, mesh, flags, a);
  if (method == 1) return launch_uniform<6, 1>(ctx, mesh, flags, a);
  return launch_uniform<6, 6>(ctx, mesh, flags, a);
}
