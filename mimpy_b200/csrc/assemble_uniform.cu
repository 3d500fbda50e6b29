// Fused numeric assembly for hexahedral meshes (every cell has 6 faces) -- the headline kernel
// (700 algorithmic bytes per hex cell).  Three kernels share one thread-per-cell core
// (cell_core.cuh):
//
//  k_numeric_brick2  (structured hex meshes in HexMesh numbering, even nx; the default): a CTA owns a
//     brick of 8x4x4 cells; inputs staged by cp.async, output image in shared memory laid out like
//     the CSR value array, cross-brick diagonals by a wait-free swap -- see the comment above it.
//
//  k_numeric_brick   (the same bricks for odd nx): per-thread loads; the 464 face rows of the brick
//     are assembled in a shared-memory row buffer of 13 slots per row and written as whole rows.
//
//  k_numeric_direct  (any uniform-6 mesh): each thread stores its 36+13 entries straight from
//     registers to their CSR slots (48 write sectors per cell: the L2 is the busiest unit).
//
// Diagonal entries M[g,g] have two contributors (the two cells of face f).  Inside a brick they are
// added in shared memory in a fixed order.  Across CTAs the two older kernels let the lower-index
// cell PUBLISH its half in diag_pub[f] and the higher-index cell add it (CTAs take their work from
// an atomic ticket, bricks ordered lexicographically, so a CTA only ever waits for CTAs that started
// earlier and are resident; the wait is bounded and raises an error flag instead of hanging).
// k_numeric_brick2 replaces the wait by an atomic swap of the halves.
#include <stdlib.h>

#include "common.cuh"
#include "cell_core.cuh"
#include "assemble_args.cuh"

// bounded wait for the half published by the lower-index cell
__device__ __forceinline__ double wait_published(const double* slot, unsigned long long first,
                                                 unsigned int* error_flag) {
  unsigned long long bits = first;
  if (bits == AU_SENTINEL) {
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(slot);
    unsigned int spins = 0;
    do {
      __nanosleep(64);
      asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(bits) : "l"(src));
    } while (bits == AU_SENTINEL && ++spins < (1u << 22));
    if (bits == AU_SENTINEL) atomicExch(error_flag, 1u);
  }
  return __longlong_as_double((long long)bits);
}

// per-cell inputs whose addresses depend on the cell index only
template <int NF>
struct CellIn {
  int fid[NF], rs[NF];
  double K[9], xc[3], vol, mult, sgn[NF];
  int crow, crow_end;
  unsigned int pm[NF * NF / 4];
  unsigned short w_dt[NF / 2], w_d[NF / 2], w_role[NF / 2];
};

template <int NF>
__device__ __forceinline__ void load_cell(const mimpy_mesh_view& m, const UniArgs& a, long long cell,
                                          bool k_unity, CellIn<NF>& in) {
  const int2* p = reinterpret_cast<const int2*>(m.cell_faces + cell * NF);
  const int2* q = reinterpret_cast<const int2*>(a.cell_rowstart + cell * NF);
#pragma unroll
  for (int i = 0; i < NF / 2; ++i) {
    const int2 v = __ldg(p + i);
    in.fid[2 * i] = v.x;
    in.fid[2 * i + 1] = v.y;
    const int2 w = __ldg(q + i);
    in.rs[2 * i] = w.x;
    in.rs[2 * i + 1] = w.y;
  }
  if (k_unity) {
#pragma unroll
    for (int t = 0; t < 9; ++t) in.K[t] = (t % 4 == 0) ? 1. : 0.;
  } else {
#pragma unroll
    for (int t = 0; t < 9; ++t) in.K[t] = __ldg(m.cell_k + cell * 9 + t);
  }
#pragma unroll
  for (int t = 0; t < 3; ++t) in.xc[t] = __ldg(m.cell_centroid + cell * 3 + t);
  in.vol = __ldg(m.cell_volume + cell);
  in.mult = a.multipliers ? __ldg(a.multipliers + cell) : 1.;
  in.crow = __ldg(a.indptr + a.flux_dof + cell);
  in.crow_end = __ldg(a.indptr + a.flux_dof + cell + 1);
  const unsigned int* pp = reinterpret_cast<const unsigned int*>(a.pos_m + cell * (NF * NF));
#pragma unroll
  for (int t = 0; t < NF * NF / 4; ++t) in.pm[t] = __ldg(pp + t);
  const unsigned short* p0 = reinterpret_cast<const unsigned short*>(m.cell_sigma + cell * NF);
  const unsigned short* p1 = reinterpret_cast<const unsigned short*>(a.pos_dt + cell * NF);
  const unsigned short* p2 = reinterpret_cast<const unsigned short*>(a.pos_d + cell * NF);
  const unsigned short* p3 = reinterpret_cast<const unsigned short*>(a.role + cell * NF);
  unsigned short w_sg[NF / 2];
#pragma unroll
  for (int t = 0; t < NF / 2; ++t) {
    w_sg[t] = __ldg(p0 + t);
    in.w_dt[t] = __ldg(p1 + t);
    in.w_d[t] = __ldg(p2 + t);
    in.w_role[t] = __ldg(p3 + t);
  }
#pragma unroll
  for (int i = 0; i < NF; ++i) in.sgn[i] = (double)(signed char)byte_of(w_sg, i);
}

template <int NF>
__device__ __forceinline__ void load_records(const mimpy_mesh_view& m, const int (&fid)[NF],
                                             double2 (&rec)[NF][4]) {
#pragma unroll
  for (int i = 0; i < NF; ++i) {
#pragma unroll
    for (int q = 0; q < 4; ++q) rec[i][q] = __ldg(rec_chunk(m.face_geom, (size_t)fid[i], q));
  }
}

#define PM_POS(pm, e) (((pm)[(e) >> 2] >> (((e) & 3) * 8)) & 0xFFu)

// ---------------------------------------------------------------------------------------------
// direct: thread-per-cell, stores from registers
// ---------------------------------------------------------------------------------------------
#define AD_THREADS 128
#ifndef AD_MINB
#define AD_MINB 2  /* measured on B200: 2 -> 7.0 ms, 3 -> 7.5 ms, 4 -> 9.5 ms at 256^3 (spills) */
#endif

template <int NF, int METHOD>
__global__ void __launch_bounds__(AD_THREADS, AD_MINB)
k_numeric_direct(mimpy_mesh_view m, int flags, UniArgs a) {
  __shared__ unsigned int s_ticket;
  if (threadIdx.x == 0) s_ticket = atomicAdd(a.ticket, 1u);
  __syncthreads();
  const long long cell = (long long)s_ticket * AD_THREADS + threadIdx.x;
  if (cell >= m.nc) return;
  CellIn<NF> in;
  load_cell<NF>(m, a, cell, flags & 1, in);
  double2 rec[NF][4];
  load_records<NF>(m, in.fid, rec);
  unsigned long long pub[NF];
#pragma unroll
  for (int i = 0; i < NF; ++i) {  // optimistic, non-blocking read of the published halves
    pub[i] = 0ull;
    if (byte_of(in.w_role, i) == 2 && in.rs[i] >= 0) {
      const unsigned long long* src = reinterpret_cast<const unsigned long long*>(a.diag_pub + in.fid[i]);
      asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(pub[i]) : "l"(src));
    }
  }
#pragma unroll
  for (int i = 0; i < NF; ++i) {  // DIV / -DIV^T                       mfd.py:232-243
    const double sa = in.sgn[i] * rec[i][0].x;
    if (in.rs[i] >= 0) {
      a.vals[in.rs[i] + byte_of(in.w_dt, i)] = -sa;
      a.vals[in.crow + byte_of(in.w_d, i)] = sa;
    }
  }
  a.vals[in.crow_end - 1] = a.c_diag ? __ldg(a.c_diag + cell) : a.alpha * in.vol;  // mfd.py:766-772
  double diag[NF];
  cell_entries<NF, METHOD>(in.K, in.xc, in.vol, in.mult, rec, in.sgn, [&](int i, int j, double v) {
    if (i == j) {
      diag[i] = v;
      // the lower-index cell of a shared face publishes its diagonal half as soon as it exists
      if (in.rs[i] >= 0 && byte_of(in.w_role, i) == 1) a.diag_pub[in.fid[i]] = v;
    } else {
      const unsigned pos = PM_POS(in.pm, i * NF + j);
      if (in.rs[i] >= 0 && pos != 0xFFu) a.vals[in.rs[i] + pos] = v;
    }
  });
#pragma unroll
  for (int i = 0; i < NF; ++i) {
    const unsigned role = byte_of(in.w_role, i);
    if (in.rs[i] < 0 || role == 1) continue;
    double d = diag[i];
    if (role == 2) d = wait_published(a.diag_pub + in.fid[i], pub[i], a.error_flag) + d;  // lower first
    a.vals[in.rs[i] + PM_POS(in.pm, i * NF + i)] = d;
  }
}

// ---------------------------------------------------------------------------------------------
// brick: 8x4x4 cells per CTA, face rows assembled in shared memory, whole-row bursts
// ---------------------------------------------------------------------------------------------
struct BrickSmem {
  double row[BR_NFACE * BR_RW];
  int rowstart[BR_NFACE];
  unsigned int ticket;
};

template <int METHOD>
#ifndef BR_MINB
#define BR_MINB 3  /* measured at 256^3 on B200: 2 -> 4.90 ms, 3 -> 4.64 ms, 4 -> 7.98 ms (spills) */
#endif
__global__ void __launch_bounds__(BR_THREADS, BR_MINB)
k_numeric_brick(mimpy_mesh_view m, int flags, UniArgs a) {
  constexpr int NF = 6;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BrickSmem& s = *reinterpret_cast<BrickSmem*>(smem_raw);
  const int t = threadIdx.x;
  if (t == 0) s.ticket = atomicAdd(a.ticket, 1u);
  // sentinel-fill the row buffer: a slot nobody in this brick writes is not stored
  unsigned long long* rowbits = reinterpret_cast<unsigned long long*>(s.row);
  for (int e = t; e < BR_NFACE * BR_RW; e += BR_THREADS) rowbits[e] = AU_SENTINEL;
  for (int e = t; e < BR_NFACE; e += BR_THREADS) s.rowstart[e] = -1;
  __syncthreads();
  const int nbx = (m.hex_nx + BX - 1) / BX, nby = (m.hex_ny + BY - 1) / BY;
  const int brick = (int)s.ticket;  // lexicographic: x fastest, z slowest
  const int bi = brick % nbx, bj = (brick / nbx) % nby, bk = brick / (nbx * nby);
  const int li = t % BX, lj = (t / BX) % BY, lk = t / (BX * BY);
  const int ci = bi * BX + li, cj = bj * BY + lj, ck = bk * BZ + lk;
  const bool ok = ci < m.hex_nx && cj < m.hex_ny && ck < m.hex_nz;
  const long long cell = ci + (long long)m.hex_nx * (cj + (long long)m.hex_ny * ck);

  // local face ids of the cell's faces [z-, y-, x-, x+, y+, z+]
  int lf[NF];
  lf[0] = (lk * BY + lj) * BX + li;
  lf[5] = ((lk + 1) * BY + lj) * BX + li;
  lf[1] = BR_NZF + (lk * (BY + 1) + lj) * BX + li;
  lf[4] = BR_NZF + (lk * (BY + 1) + lj + 1) * BX + li;
  lf[2] = BR_NZF + BR_NYF + (lk * BY + lj) * (BX + 1) + li;
  lf[3] = BR_NZF + BR_NYF + (lk * BY + lj) * (BX + 1) + li + 1;
  // is the other cell of face i inside this brick?  (it exists iff role != 0)
  const bool in_brick[NF] = {lk > 0, lj > 0, li > 0, li < BX - 1, lj < BY - 1, lk < BZ - 1};

  CellIn<NF> in;
  double diag[NF];
  if (ok) {
    load_cell<NF>(m, a, cell, flags & 1, in);
    double2 rec[NF][4];
    load_records<NF>(m, in.fid, rec);
    unsigned long long pub[NF];
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // halves published by lower cells of OTHER bricks
      pub[i] = 0ull;
      if (byte_of(in.w_role, i) == 2 && in.rs[i] >= 0 && !in_brick[i]) {
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(a.diag_pub + in.fid[i]);
        asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(pub[i]) : "l"(src));
      }
    }
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      const double sa = in.sgn[i] * rec[i][0].x;
      if (in.rs[i] >= 0) {
        s.rowstart[lf[i]] = in.rs[i];                             // both cells write the same value
        s.row[lf[i] * BR_RW + byte_of(in.w_dt, i)] = -sa;         // -DIV^T   mfd.py:240-243
        a.vals[in.crow + byte_of(in.w_d, i)] = sa;                // DIV      mfd.py:232-238
      }
    }
    a.vals[in.crow_end - 1] = a.c_diag ? __ldg(a.c_diag + cell) : a.alpha * in.vol;
    cell_entries<NF, METHOD>(in.K, in.xc, in.vol, in.mult, rec, in.sgn, [&](int i, int j, double v) {
      if (i == j) {
        diag[i] = v;
      } else {
        const unsigned pos = PM_POS(in.pm, i * NF + j);
        if (in.rs[i] >= 0 && pos != 0xFFu) s.row[lf[i] * BR_RW + pos] = v;
      }
    });
    // diagonals, step 1: single-cell faces and the LOWER cell of shared faces.  Publish first
    // (faces x+, y+, z+), wait afterwards (faces z-, y-, x-): a brick never delays its successors
    // by its own waiting.
#pragma unroll
    for (int ii = 0; ii < NF; ++ii) {
      const int i = (ii + 3) % NF;  // 3, 4, 5, 0, 1, 2
      if (in.rs[i] < 0) continue;
      const unsigned role = byte_of(in.w_role, i);
      const int slot = lf[i] * BR_RW + PM_POS(in.pm, i * NF + i);
      if (role == 0) s.row[slot] = diag[i];
      if (role == 1) {
        if (in_brick[i]) s.row[slot] = diag[i];           // finished in shared memory by the upper cell
        else a.diag_pub[in.fid[i]] = diag[i];              // the upper cell lives in a later brick
      }
      if (role == 2 && !in_brick[i])                       // lower cell lives in an earlier brick
        s.row[slot] = wait_published(a.diag_pub + in.fid[i], pub[i < 3 ? i : 0], a.error_flag) + diag[i];
    }
  }
  __syncthreads();
  // diagonals, step 2: the UPPER cell adds its half (lower + upper, fixed order)
  if (ok) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (in.rs[i] >= 0 && byte_of(in.w_role, i) == 2 && in_brick[i]) {
        const int slot = lf[i] * BR_RW + PM_POS(in.pm, i * NF + i);
        s.row[slot] = s.row[slot] + diag[i];
      }
    }
  }
  __syncthreads();
  // stream the rows out: 8 rows x 13 slots per pass (104 of the 128 threads), thread = fixed slot
  // of a fixed row-in-pass, so that no index arithmetic is left in the loop; consecutive lanes hold
  // consecutive slots of a row
  if (t < 8 * BR_RW) {
    const int r0 = t / BR_RW, pos = t - r0 * BR_RW;
#pragma unroll 8
    for (int r = r0; r < BR_NFACE; r += 8) {
      const unsigned long long bits = rowbits[r * BR_RW + pos];
      if (bits != AU_SENTINEL) a.vals[s.rowstart[r] + pos] = __longlong_as_double((long long)bits);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// staged brick: same brick, but every input is brought into shared memory by coalesced
// asynchronous copies (cp.async) instead of per-thread strided loads, and the output leaves through an
// image of the CSR value array.  The per-thread
// loads of the kernel above touch one 32-byte sector per lane and instruction (53 sectors per
// cell measured, the L1 data pipe was the busiest unit); here
//   * the 464 face records of the brick are fetched once (not once per adjacent cell), 4 lanes
//     per 64-byte record, into a 16-byte-chunk XOR swizzle that makes the later per-cell
//     128-bit reads conflict free;
//   * K, centroid, volume and the symbolic byte/int maps of the 16 x-runs of 8 cells are copied
//     as contiguous runs;
//   * face ids come from the closed form of the HexMesh numbering (hexmesh.py:161-234), so no
//     load depends on another load.
// The staging area for records/K/centroid/volume ALIASES the output image: operands are pulled into
// registers (R~, KN), then the image is sentinel-filled (a slot nobody in this brick writes is not
// stored), filled by the cells and streamed out with consecutive lanes writing consecutive values.
// Bricks are walked in super-tiles of 4x4x4 so that the two bricks of a shared surface row run
// close in time and their half rows merge in L2.  Diagonals shared with another brick: both cells
// swap their half into a face-indexed exchange slot; the second to arrive writes half + half and
// restores the sentinel -- wait-free, and bit-reproducible because IEEE addition commutes.
// Needs hex_nx even (16-byte alignment of the cell runs).  History and measurements: DESIGN.md 6.
// ---------------------------------------------------------------------------------------------
#define B2_OFF_REC 0
#define B2_OFF_K (BR_NFACE * 64)
#define B2_OFF_XC (B2_OFF_K + BR_THREADS * 72)
#define B2_OFF_VOL (B2_OFF_XC + BR_THREADS * 24)
#define B2_OFF_ROWSTART (B2_OUT_SLOTS * 8)   /* int run_base[16], cell_base[16], extra_rowstart[80] */
#define B2_OFF_CRS (B2_OFF_ROWSTART + (2 * B2_RUNS + B2_EXTRA_ROWS) * 4)
#define B2_OFF_PM (B2_OFF_CRS + BR_THREADS * 24)
#define B2_OFF_SG (B2_OFF_PM + BR_THREADS * 36)
#define B2_OFF_DT (B2_OFF_SG + BR_THREADS * 6)
#define B2_OFF_D (B2_OFF_DT + BR_THREADS * 6)
#define B2_OFF_ROLE (B2_OFF_D + BR_THREADS * 6)
#define B2_SMEM (B2_OFF_ROLE + BR_THREADS * 6)
static_assert(B2_OFF_VOL + BR_THREADS * 8 <= B2_OFF_ROWSTART, "staging area must fit in the row buffer");
static_assert(B2_OFF_ROWSTART % 16 == 0 && B2_OFF_CRS % 16 == 0 && B2_OFF_PM % 16 == 0, "alignment");

template <int BYTES>
__device__ __forceinline__ void cp_async(unsigned dst, const void* src) {
  if (BYTES == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  else if (BYTES == 8)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}

// 8 lanes copy one x-run (nbytes, a multiple of 4; of 16 when vec16) of a per-cell array
template <int MAXBYTES>
__device__ __forceinline__ void stage_run(unsigned dst, const void* src_, int lane8, int nbytes, bool vec16) {
  const char* src = reinterpret_cast<const char*>(src_);
  if (vec16) {
#pragma unroll
    for (int off0 = 0; off0 < MAXBYTES; off0 += 128) {
      const int off = off0 + lane8 * 16;
      if (off < nbytes) cp_async<16>(dst + off, src + off);
    }
  } else {
#pragma unroll
    for (int off0 = 0; off0 < MAXBYTES; off0 += 32) {
      const int off = off0 + lane8 * 4;
      if (off < nbytes) cp_async<4>(dst + off, src + off);
    }
  }
}

__device__ __forceinline__ void stage_record(unsigned sbase, const double* face_geom, int lf, int c, int fid) {
  cp_async<16>(sbase + B2_OFF_REC + lf * 64 + (((c ^ (lf >> 1)) & 3) << 4), face_geom + 8 * (size_t)fid + 2 * rec_pos(fid, c));
}

template <int METHOD>
__global__ void __launch_bounds__(BR_THREADS, BR_MINB)
k_numeric_brick2(mimpy_mesh_view m, int flags, UniArgs a, int bk_begin, int bk_end) {
  constexpr int NF = 6;
  extern __shared__ __align__(128) unsigned char sm[];
  double* row = reinterpret_cast<double*>(sm);
  int* s_runbase = reinterpret_cast<int*>(sm + B2_OFF_ROWSTART);
  int* s_cellbase = s_runbase + B2_RUNS;
  int* s_extra_rs = s_cellbase + B2_RUNS;
  const int t = threadIdx.x;
  if (t < B2_EXTRA_ROWS) s_extra_rs[t] = 0;  // (barriers follow before use)
  const int nx = m.hex_nx, ny = m.hex_ny, nz = m.hex_nz;
  // Bricks are independent of each other; blockIdx walks them in super-tiles of 4x4x4 bricks so
  // that the two bricks sharing a surface row (and its face record) run close together in time
  // and their half rows merge in L2 instead of each making a partial-sector round trip to HBM.
  const int nbx = (nx + BX - 1) / BX, nby = (ny + BY - 1) / BY, nbz = (nz + BZ - 1) / BZ;
  const int nsx = (nbx + 3) >> 2, nsy = (nby + 3) >> 2;
  const int st = (int)(blockIdx.x >> 6), wi = (int)(blockIdx.x & 63);
  const int bi = (st % nsx) * 4 + (wi & 3), bj = ((st / nsx) % nsy) * 4 + ((wi >> 2) & 3),
            bk = bk_begin + (st / (nsx * nsy)) * 4 + (wi >> 4);  // [bk_begin, bk_end): a slab of brick layers
  if (bi >= nbx || bj >= nby || bk >= min(nbz, bk_end)) return;
  const int ci0 = bi * BX, cj0 = bj * BY, ck0 = bk * BZ;
  const int runlen = min(BX, nx - ci0);
  const int li = t % BX, lj = (t / BX) % BY, lk = t / (BX * BY);
  const int ci = ci0 + li, cj = cj0 + lj, ck = ck0 + lk;
  const bool run_ok = cj < ny && ck < nz;          // the x-run (lj, lk) this thread belongs to
  const bool ok = run_ok && ci < nx;
  const int cell0 = ci0 + nx * (cj + ny * ck);     // first cell of the run
  const int cell = cell0 + li;
  const int L = nx * ny + nx * (ny + 1) + (nx + 1) * ny;  // nf < 2^31 (checked by hex_topology)
  const int xrow = 3 * nx + 1;
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
  const bool k_unity = flags & 1;

  // ---- stage: face records, 4 lanes per 64-byte record (HexMesh numbering, hexmesh.py:161-234) ----
  {
    const int c = t & 3, i = (t >> 2) & 7, w = t >> 5;
    const int gi = ci0 + i;
    if (gi < nx) {
      {  // z faces: lj = w, lk = 0..BZ
        const int gj = cj0 + w;
        int fid = ck0 * L + gj * xrow + 3 * gi;
#pragma unroll
        for (int q = 0; q <= BZ; ++q, fid += L) {
          const int gk = ck0 + q;
          if (gj < ny && gk <= nz)
            stage_record(sbase, m.face_geom, (q * BY + w) * BX + i, c, gk < nz ? fid : nz * L + gj * nx + gi);
        }
      }
      {  // y faces: lk = w, lj = 0..BY
        const int gk = ck0 + w;
        int fid = gk * L + cj0 * xrow + 3 * gi + 1;
#pragma unroll
        for (int q = 0; q <= BY; ++q, fid += xrow) {
          const int gj = cj0 + q;
          if (gk < nz && gj <= ny)
            stage_record(sbase, m.face_geom, BR_NZF + (w * (BY + 1) + q) * BX + i, c,
                         gj < ny ? fid : gk * L + ny * xrow + gi);
        }
      }
    }
    if (run_ok) {  // x faces of this thread's run: 9 records = 36 chunks over 8 lanes
      const int base = ck * L + cj * xrow;
#pragma unroll
      for (int q0 = 0; q0 < 4 * (BX + 1); q0 += 8) {
        const int cc = q0 + (t & 7);
        const int q = cc >> 2, gx = ci0 + q;
        if (cc < 4 * (BX + 1) && gx <= nx)
          stage_record(sbase, m.face_geom, BR_NZF + BR_NYF + (t >> 3) * (BX + 1) + q, cc & 3,
                       base + (gx < nx ? 3 * gx + 2 : 3 * nx));
      }
    }
  }
  // ---- stage: per-cell arrays of this thread's x-run (8 lanes per run) ----
  if (run_ok) {
    const int r = t >> 3, l8 = t & 7;
    const bool v16_pm = (nx & 3) == 0 && (runlen & 3) == 0;
    const bool v16_b6 = (nx & 7) == 0 && runlen == BX;
    const size_t c0 = (size_t)cell0;
    if (!k_unity) stage_run<BX * 72>(sbase + B2_OFF_K + r * (BX * 72), m.cell_k + c0 * 9, l8, runlen * 72, true);
    stage_run<BX * 24>(sbase + B2_OFF_XC + r * (BX * 24), m.cell_centroid + c0 * 3, l8, runlen * 24, true);
    stage_run<BX * 8>(sbase + B2_OFF_VOL + r * (BX * 8), m.cell_volume + c0, l8, runlen * 8, true);
    stage_run<BX * 24>(sbase + B2_OFF_CRS + r * (BX * 24), a.cell_rowstart + c0 * 6, l8, runlen * 24, true);
    stage_run<BX * 36>(sbase + B2_OFF_PM + r * (BX * 36), a.pos_m + c0 * 36, l8, runlen * 36, v16_pm);
    stage_run<BX * 6>(sbase + B2_OFF_SG + r * (BX * 6), m.cell_sigma + c0 * 6, l8, runlen * 6, v16_b6);
    stage_run<BX * 6>(sbase + B2_OFF_DT + r * (BX * 6), a.pos_dt + c0 * 6, l8, runlen * 6, v16_b6);
    stage_run<BX * 6>(sbase + B2_OFF_D + r * (BX * 6), a.pos_d + c0 * 6, l8, runlen * 6, v16_b6);
    stage_run<BX * 6>(sbase + B2_OFF_ROLE + r * (BX * 6), a.role + c0 * 6, l8, runlen * 6, v16_b6);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");

  // local face ids of the cell's faces [z-, y-, x-, x+, y+, z+]
  int lf[NF];
  lf[0] = (lk * BY + lj) * BX + li;
  lf[5] = ((lk + 1) * BY + lj) * BX + li;
  lf[1] = BR_NZF + (lk * (BY + 1) + lj) * BX + li;
  lf[4] = BR_NZF + (lk * (BY + 1) + lj + 1) * BX + li;
  lf[2] = BR_NZF + BR_NYF + (lk * BY + lj) * (BX + 1) + li;
  lf[3] = BR_NZF + BR_NYF + (lk * BY + lj) * (BX + 1) + li + 1;
  const bool in_brick[NF] = {lk > 0, lj > 0, li > 0, li < BX - 1, lj < BY - 1, lk < BZ - 1};
  const int cbase = ck * L + cj * xrow + 3 * ci;  // global ids of faces z-, y-, x- are cbase + 0, 1, 2

  int crow = 0, crow_end = 0;
  double mult = 1., cdiag = 0.;
  if (ok) {
    crow = __ldg(a.indptr + a.flux_dof + cell);
    crow_end = __ldg(a.indptr + a.flux_dof + cell + 1);
    if (a.multipliers) mult = __ldg(a.multipliers + cell);
    if (a.c_diag) cdiag = __ldg(a.c_diag + cell);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- pull: operands into registers ----
  CellOps<NF> o;
  int rs[NF];
  unsigned int pm[NF * NF / 4];
  unsigned short w_dt[NF / 2], w_d[NF / 2], w_role[NF / 2];
  double sa[NF], vol = 0., trk = 1.;
  if (ok) {
    double K[9], xc[3], sgn[NF];
    double2 rec[NF][4];
#pragma unroll
    for (int i = 0; i < NF; ++i)
#pragma unroll
      for (int q = 0; q < 4; ++q)
        rec[i][q] = *reinterpret_cast<const double2*>(sm + B2_OFF_REC + lf[i] * 64 + (((q ^ (lf[i] >> 1)) & 3) << 4));
    if (k_unity) {
#pragma unroll
      for (int q = 0; q < 9; ++q) K[q] = (q % 4 == 0) ? 1. : 0.;
    } else {
#pragma unroll
      for (int q = 0; q < 9; ++q) K[q] = *reinterpret_cast<const double*>(sm + B2_OFF_K + t * 72 + q * 8);
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) xc[q] = *reinterpret_cast<const double*>(sm + B2_OFF_XC + t * 24 + q * 8);
    vol = *reinterpret_cast<const double*>(sm + B2_OFF_VOL + t * 8);
    unsigned short w_sg[NF / 2];
#pragma unroll
    for (int q = 0; q < NF / 2; ++q) {
      const int2 v = *reinterpret_cast<const int2*>(sm + B2_OFF_CRS + t * 24 + q * 8);
      rs[2 * q] = v.x;
      rs[2 * q + 1] = v.y;
      w_sg[q] = *reinterpret_cast<const unsigned short*>(sm + B2_OFF_SG + t * 6 + q * 2);
      w_dt[q] = *reinterpret_cast<const unsigned short*>(sm + B2_OFF_DT + t * 6 + q * 2);
      w_d[q] = *reinterpret_cast<const unsigned short*>(sm + B2_OFF_D + t * 6 + q * 2);
      w_role[q] = *reinterpret_cast<const unsigned short*>(sm + B2_OFF_ROLE + t * 6 + q * 2);
    }
#pragma unroll
    for (int q = 0; q < NF * NF / 4; ++q) pm[q] = *reinterpret_cast<const unsigned int*>(sm + B2_OFF_PM + t * 36 + q * 4);
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      sgn[i] = (double)(signed char)byte_of(w_sg, i);
      sa[i] = sgn[i] * rec[i][0].x;
    }
    trk = K[0] + K[4] + K[8];
    cell_prep<NF, METHOD>(K, xc, rec, sgn, o);
  }
  const int run = t >> 3;
  {  // CSR offset of the run's first face row (rows z-, y-, x- of a cell ascend) and first cell row
    int first = 0x7fffffff;
    if (ok) first = rs[0] >= 0 ? rs[0] : rs[1] >= 0 ? rs[1] : rs[2] >= 0 ? rs[2] : 0x7fffffff;
#pragma unroll
    for (int d = 4; d > 0; d >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, d));
    if (li == 0) {
      s_runbase[run] = first;
      s_cellbase[run] = crow;
    }
  }
  __syncthreads();
  // sentinel-fill the output image: a slot nobody in this brick writes is not stored
  {
    const ulonglong2 s2 = make_ulonglong2(AU_SENTINEL, AU_SENTINEL);
    ulonglong2* r2 = reinterpret_cast<ulonglong2*>(sm);
    for (int e = t; e < B2_OUT_SLOTS / 2; e += BR_THREADS) r2[e] = s2;
  }
  __syncthreads();

  double diag[NF];
  unsigned long long other[NF];
  int sb[NF];  // where slot 0 of the face's row lives in the output image
  const int fid[NF] = {cbase, cbase + 1, cbase + 2,
                       ci + 1 < nx ? cbase + 5 : ck * L + cj * xrow + 3 * nx,
                       cj + 1 < ny ? cbase + xrow + 1 : ck * L + ny * xrow + ci,
                       ck + 1 < nz ? cbase + L : nz * L + cj * nx + ci};
  if (ok) {
    const int rb = s_runbase[run];
    sb[0] = run * B2_RUN_SLOTS + rs[0] - rb;
    sb[1] = run * B2_RUN_SLOTS + rs[1] - rb;
    sb[2] = run * B2_RUN_SLOTS + rs[2] - rb;
    if (li < BX - 1 && ci + 1 < nx) {
      sb[3] = run * B2_RUN_SLOTS + rs[3] - rb;
    } else {
      sb[3] = B2_EXTRA0 + run * BR_RW;
      s_extra_rs[run] = rs[3];
    }
    if (lj < BY - 1 && cj + 1 < ny) {
      sb[4] = (run + 1) * B2_RUN_SLOTS + rs[4] - s_runbase[run + 1];
    } else {
      sb[4] = B2_EXTRA0 + (B2_RUNS + lk * BX + li) * BR_RW;
      s_extra_rs[B2_RUNS + lk * BX + li] = rs[4];
    }
    if (lk < BZ - 1 && ck + 1 < nz) {
      sb[5] = (run + BY) * B2_RUN_SLOTS + rs[5] - s_runbase[run + BY];
    } else {
      sb[5] = B2_EXTRA0 + (B2_RUNS + BX * BZ + lj * BX + li) * BR_RW;
      s_extra_rs[B2_RUNS + BX * BZ + lj * BX + li] = rs[5];
    }
    const int cb = B2_CELL0 + run * B2_CELL_SLOTS + (crow - s_cellbase[run]);
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      if (rs[i] >= 0) {
        row[sb[i] + byte_of(w_dt, i)] = -sa[i];                 // -DIV^T   mfd.py:240-243
        row[cb + byte_of(w_d, i)] = sa[i];                      // DIV      mfd.py:232-238
      }
    }
    row[cb + (crow_end - 1 - crow)] = a.c_diag ? cdiag : a.alpha * vol;    // mfd.py:766-772
    cell_core<NF, METHOD>(o, trk, vol, mult, [&](int i, int j, double v) {
      if (i == j) {
        diag[i] = v;
      } else {
        const unsigned pos = PM_POS(pm, i * NF + j);
        if (rs[i] >= 0 && pos != 0xFFu) row[sb[i] + pos] = v;
      }
    });
    // diagonals, step 1: single-cell faces and the LOWER cell of faces inside the brick go to the
    // row buffer.  A face shared with ANOTHER brick: both cells swap their half into diag_pub[f]
    // (all-sentinel between launches); whoever finds the other's half there -- the second to arrive --
    // writes half + half to the CSR slot and restores the sentinel (step 3).  IEEE addition commutes, so the result does not
    // depend on the arrival order: bit-reproducible, and nobody ever waits.
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      other[i] = AU_SENTINEL;
      if (rs[i] < 0) continue;
      const unsigned role = byte_of(w_role, i);
      if (role == 0 || (role == 1 && in_brick[i])) row[sb[i] + PM_POS(pm, i * NF + i)] = diag[i];
      if (role != 0 && !in_brick[i])
        other[i] = atomicExch(reinterpret_cast<unsigned long long*>(a.diag_pub) + fid[i],
                              (unsigned long long)__double_as_longlong(diag[i]));
    }
  }
  __syncthreads();
  if (ok) {  // diagonals, step 2: the UPPER cell adds its half (lower + upper, fixed order)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (rs[i] >= 0 && byte_of(w_role, i) == 2 && in_brick[i]) {
        const int slot = sb[i] + PM_POS(pm, i * NF + i);
        row[slot] = row[slot] + diag[i];
      }
    }
  }
  __syncthreads();
  // stream the image out; consecutive lanes write consecutive CSR values.  A sentinel has an
  // all-ones high word, which no finite value has.
  {  // face rows of the runs: 312 slots = 128 + 128 + 56 per run, addresses fixed per thread
    static_assert(B2_RUN_SLOTS > 2 * BR_THREADS && B2_RUN_SLOTS <= 3 * BR_THREADS, "three passes per run");
    const double* q = row + t;
    double* out = a.vals + t;
#pragma unroll 4
    for (int r = 0; r < B2_RUNS; ++r, q += B2_RUN_SLOTS) {
      const int base = s_runbase[r];
      if (base == 0x7fffffff) continue;
      const double v0 = q[0], v1 = q[BR_THREADS];
      const double v2 = t < B2_RUN_SLOTS - 2 * BR_THREADS ? q[2 * BR_THREADS] : __longlong_as_double(-1ll);
      double* o = out + base;
      if (__double2hiint(v0) != -1) o[0] = v0;
      if (__double2hiint(v1) != -1) o[BR_THREADS] = v1;
      if (__double2hiint(v2) != -1) o[2 * BR_THREADS] = v2;
    }
  }
  {  // cell rows: two runs per pass, 56 of 64 lanes each
    const int half = t >> 6, l = t & 63;
    if (l < B2_CELL_SLOTS) {
#pragma unroll
      for (int r = half; r < B2_RUNS; r += 2) {
        const double v = row[B2_CELL0 + r * B2_CELL_SLOTS + l];
        if (__double2hiint(v) != -1) a.vals[s_cellbase[r] + l] = v;
      }
    }
  }
  if (t < 8 * BR_RW) {  // rows on the brick's upper surfaces: 8 rows x 13 slots per pass
    const int r0 = t / BR_RW, pos = t - r0 * BR_RW;
#pragma unroll
    for (int e = r0; e < B2_EXTRA_ROWS; e += 8) {
      const double v = row[B2_EXTRA0 + e * BR_RW + pos];
      if (__double2hiint(v) != -1) a.vals[s_extra_rs[e] + pos] = v;
    }
  }
  // diagonals, step 3: second arrivals at a face shared with another brick finish the slot
  if (ok) {
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      if (other[i] != AU_SENTINEL) {
        a.vals[rs[i] + PM_POS(pm, i * NF + i)] = __longlong_as_double((long long)other[i]) + diag[i];
        reinterpret_cast<unsigned long long*>(a.diag_pub)[fid[i]] = AU_SENTINEL;  // leave the slot clean
      }
    }
  }
}

// 0 direct, 1 brick, 2 staged brick (cp.async), 3 pipelined brick (bulk copies, assemble_brick3.cu; the
// default where the mesh is a structured hex mesh with even nx)
static int asm_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MIMPY_B200_ASM_VARIANT");
    v = (e && e[0] == 'd') ? 0 : (e && e[0] == 'b') ? 1 : (e && e[0] == '2') ? 2 : 3;
  }
  return v;
}

template <int METHOD>
static int launch_uniform(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, int flags, const UniArgs& a,
                          int bk_begin, int bk_end, bool staged) {
  const bool structured = mesh->hex_nx > 0 && mesh->hex_ny > 0 && mesh->hex_nz > 0 &&
                          (long long)mesh->hex_nx * mesh->hex_ny * mesh->hex_nz == mesh->nc;
  // cudaFuncSetAttribute is per device and cheap: no process-wide "configured" flag (a second device
  // in the same process would never get the opt-in)
  if (staged) {
    CUDA_TRY(cudaFuncSetAttribute(k_numeric_brick2<METHOD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B2_SMEM));
    const int nbz = (mesh->hex_nz + BZ - 1) / BZ;
    if (bk_end > nbz) bk_end = nbz;
    const long long super_tiles = (long long)(((mesh->hex_nx + BX - 1) / BX + 3) / 4) * (((mesh->hex_ny + BY - 1) / BY + 3) / 4) *
                                  ((bk_end - bk_begin + 3) / 4);
    if (super_tiles > 0)
      k_numeric_brick2<METHOD><<<(unsigned)(super_tiles * 64), BR_THREADS, B2_SMEM, ctx->stream>>>(*mesh, flags, a, bk_begin,
                                                                                                   bk_end);
  } else if (structured && asm_variant() >= 1 && !(flags & MIMPY_FLAG_DIRECT_KERNEL)) {
    const size_t smem = sizeof(BrickSmem);
    CUDA_TRY(cudaFuncSetAttribute(k_numeric_brick<METHOD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long bricks = (long long)((mesh->hex_nx + BX - 1) / BX) * ((mesh->hex_ny + BY - 1) / BY) *
                             ((mesh->hex_nz + BZ - 1) / BZ);
    k_numeric_brick<METHOD><<<(unsigned)bricks, BR_THREADS, smem, ctx->stream>>>(*mesh, flags, a);
  } else {
    const int blocks = (mesh->nc + AD_THREADS - 1) / AD_THREADS;
    k_numeric_direct<6, METHOD><<<blocks, AD_THREADS, 0, ctx->stream>>>(*mesh, flags, a);
  }
  KERNEL_CHECK();
  return MIMPY_OK;
}

// Exchange slots of the wait-free diagonal swap: persistent, all-sentinel between FULL assemblies.  A
// launch over a range of cell layers [b, e) leaves the halves of the z faces of layers b and e in their
// slots until the neighbouring range is assembled.  Left there they would be taken for the partner's
// half by a REPEATED or out-of-order launch, so the layers a launch is about to write from the same side
// again are re-sentinelled first (1 = left by the cells below the layer, 2 = by the cells above).
static int swap_prepare(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, bool ranged, int layer_begin, int layer_end) {
  const int nx = mesh->hex_nx, ny = mesh->hex_ny, nz = mesh->hex_nz;
  const int L = nx * ny + nx * (ny + 1) + (nx + 1) * ny;
  if (ctx->diag_swap_len < (size_t)mesh->nf) {
    if (ctx->diag_swap) {
      CUDA_TRY(cudaStreamSynchronize(ctx->stream));
      CUDA_TRY(cudaFree(ctx->diag_swap));
    }
    ctx->diag_swap = nullptr;
    ctx->diag_swap_len = 0;
    CUDA_TRY(cudaMalloc(&ctx->diag_swap, sizeof(double) * (size_t)mesh->nf));
    ctx->diag_swap_len = (size_t)mesh->nf;
    CUDA_TRY(cudaMemsetAsync(ctx->diag_swap, 0xFF, sizeof(double) * (size_t)mesh->nf, ctx->stream));
    ctx->swap_dirty_n = 0;
  }
  if (ctx->swap_dirty_n > 0 && ctx->swap_L != L) {  // leftovers of another mesh
    CUDA_TRY(cudaMemsetAsync(ctx->diag_swap, 0xFF, sizeof(double) * ctx->diag_swap_len, ctx->stream));
    ctx->swap_dirty_n = 0;
  }
  ctx->swap_L = L;
  auto clean_layer = [&](int layer) -> int {
    const size_t first = (size_t)layer * L;
    const size_t count = layer < nz ? (size_t)L : (size_t)nx * ny;
    if (first < ctx->diag_swap_len)
      CUDA_TRY(cudaMemsetAsync(ctx->diag_swap + first, 0xFF, sizeof(double) * (first + count <= ctx->diag_swap_len ? count : ctx->diag_swap_len - first),
                               ctx->stream));
    return MIMPY_OK;
  };
  auto find = [&](int layer) {
    for (int q = 0; q < ctx->swap_dirty_n; ++q)
      if (ctx->swap_dirty[q][0] == layer) return q;
    return -1;
  };
  auto erase = [&](int q) {
    ctx->swap_dirty[q][0] = ctx->swap_dirty[ctx->swap_dirty_n - 1][0];
    ctx->swap_dirty[q][1] = ctx->swap_dirty[ctx->swap_dirty_n - 1][1];
    --ctx->swap_dirty_n;
  };
  if (!ranged) {  // a full assembly completes nothing that is pending: start clean, end clean
    for (int q = 0; q < ctx->swap_dirty_n; ++q) {
      int rc = clean_layer(ctx->swap_dirty[q][0]);
      if (rc) return rc;
    }
    ctx->swap_dirty_n = 0;
    return MIMPY_OK;
  }
  // boundary layers this launch touches from one side only
  const int bnd[2][2] = {{layer_begin, 2}, {layer_end >= nz ? -1 : layer_end, 1}};
  for (int s = 0; s < 2; ++s) {
    const int layer = bnd[s][0], side = bnd[s][1];
    if (layer <= 0) continue;  // the mesh boundary: single-cell faces, no swap
    const int q = find(layer);
    if (q >= 0 && ctx->swap_dirty[q][1] != side) {
      erase(q);  // the halves waiting there are the partners of this launch: completed by it
    } else {
      if (q >= 0) {  // same side again: stale halves of an earlier launch of the same range
        int rc = clean_layer(layer);
        if (rc) return rc;
      } else if (ctx->swap_dirty_n < 8) {
        ctx->swap_dirty[ctx->swap_dirty_n][0] = layer;
        ctx->swap_dirty[ctx->swap_dirty_n][1] = side;
        ++ctx->swap_dirty_n;
      } else {  // more pending layers than we track: clean everything but what this launch completes
        CUDA_TRY(cudaMemsetAsync(ctx->diag_swap, 0xFF, sizeof(double) * ctx->diag_swap_len, ctx->stream));
        ctx->swap_dirty_n = 1;
        ctx->swap_dirty[0][0] = layer;
        ctx->swap_dirty[0][1] = side;
      }
    }
  }
  return MIMPY_OK;
}

// returns MIMPY_E_UNSUPPORTED when the fast path does not apply (caller falls back to the
// generic warp-group kernel of local_matrices.cu -- same results, CUDA as well)
int mimpy_numeric_uniform(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym,
                          int method, int flags, const double* multipliers, double alpha,
                          const double* c_diag, double* vals, int layer_begin, int layer_end) {
  if (mesh->uniform_faces != 6) return MIMPY_E_UNSUPPORTED;
  if (method != 0 && method != 1 && method != 6) return MIMPY_E_UNSUPPORTED;
  const bool pipelined = asm_variant() == 3 && mimpy_brick3_applies(mesh, sym, method, flags);
  if (!pipelined && (!sym->role || !sym->cell_rowstart)) return MIMPY_E_UNSUPPORTED;
  const bool staged = !pipelined && mesh->hex_nx > 0 && mesh->hex_ny > 0 && mesh->hex_nz > 0 &&
                      (long long)mesh->hex_nx * mesh->hex_ny * mesh->hex_nz == mesh->nc && asm_variant() >= 2 &&
                      mesh->hex_nx % 2 == 0 && !(flags & MIMPY_FLAG_DIRECT_KERNEL);
  // a range of cell layers [layer_begin, layer_end) is only served by the (staged / pipelined) brick kernels
  const bool ranged = !(layer_begin == 0 && layer_end < 0);
  if (ranged && ((!staged && !pipelined) || layer_begin % BZ != 0 || (layer_end % BZ != 0 && layer_end != mesh->hex_nz)))
    return MIMPY_E_UNSUPPORTED;
  const int bk_begin = ranged ? layer_begin / BZ : 0;
  const int bk_end = ranged ? (layer_end + BZ - 1) / BZ : 0x3fffffff;
  double* exchange = nullptr;
  if (staged || pipelined) {
    int rc = swap_prepare(ctx, mesh, ranged && !(layer_begin == 0 && layer_end >= mesh->hex_nz), layer_begin, layer_end);
    if (rc) return rc;
    exchange = ctx->diag_swap;
  } else {
    void* scratch = nullptr;
    int rc = mimpy_scratch(ctx, sizeof(double) * (size_t)mesh->nf, &scratch);
    if (rc) return rc;
    CUDA_TRY(cudaMemsetAsync(scratch, 0xFF, sizeof(double) * (size_t)mesh->nf, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(ctx->counters + 8, 0, 2 * sizeof(unsigned int), ctx->stream));
    exchange = (double*)scratch;
  }
  UniArgs a;
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
  a.diag_pub = exchange;
  a.ticket = ctx->counters + 8;
  a.error_flag = ctx->counters + 9;
  a.flux_dof = sym->flux_dof;
  if (pipelined) a.ticket = ctx->counters + 12;  // monotonic across launches (no memset): see B3Geo::ticket_base
  if (pipelined) return mimpy_numeric_brick3(ctx, mesh, sym, method, flags, a, bk_begin, bk_end);
  if (method == 0) return launch_uniform<0>(ctx, mesh, flags, a, bk_begin, bk_end, staged);
  if (method == 1) return launch_uniform<1>(ctx, mesh, flags, a, bk_begin, bk_end, staged);
  return launch_uniform<6>(ctx, mesh, flags, a, bk_begin, bk_end, staged);
}
