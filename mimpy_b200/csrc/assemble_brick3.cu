// Pipelined brick assembly for structured hex meshes (HexMesh numbering, even nx) -- the headline kernel.
// Fuses L1-L4 (MFD.build_m_e, mfd.py:357-483) with A2-A5/A8 (build_m / build_div / build_bottom_right /
// update_m, mfd.py:545-599, 203-246, 745-774, 729-743) like the other kernels of assemble_uniform.cu,
// but is built around the Blackwell/Hopper bulk-copy engine instead of per-thread loads:
//
//   * PERSISTENT CTAs (2 per SM, 128 threads = one brick of 8x4x4 cells at a time) walk the bricks in
//     super-tile order.  The inputs of a brick -- per x-run of 8 cells: the 24 CONSECUTIVE face records of
//     its z-/y-/x- faces plus the x+ record of the last cell (1728 B), K (576 B), centroid (192 B), volume
//     (64 B); plus, per cell row of the brick's upper y / z surface, the record run that holds its y+ / z+
//     faces -- arrive by 72 1-D bulk copies (cp.async.bulk -> SASS UBLKCP) issued by a dedicated PRODUCER
//     warp (the fifth warp of the CTA) and tracked by a full/empty pair of mbarriers (expect_tx).  As soon
//     as the operands of brick n sit in the registers of the four compute warps, the copies for brick n+1
//     are issued, so they fly during the whole compute / stream-out of brick n: no compute thread ever
//     forms a load address, and the staging latency that brick2 exposed (cp.async + wait_group 0) is hidden.
//   * NO symbolic map is read.  The 82 B/cell of pos_m / pos_dt / pos_d / role / cell_rowstart / sigma
//     of the older kernels are closed-form on a structured hex mesh once two facts per face are known:
//     the CSR offset of its row and which faces of its two cells are flux DOFs (non-Neumann).  Both live
//     in the otherwise unused eighth double of the 64-byte face record (`bind`, common.cuh), written
//     once per (mesh, plan) by k_bind_plan_hex: they arrive with the record, for free.  The position of
//     entry (i, j) inside row g(f_i) is a popcount over the merged, id-sorted face list of the two cells.
//   * The output image in shared memory, the cross-brick diagonal swap and the stream-out are those of
//     k_numeric_brick2 (assemble_uniform.cu).
//
// Algorithmic bytes per cell: 296 read (3 records, K, centroid, volume) + 4 (cell row offset) + 368
// written = 668 (+32 B of the old count that no longer exist).  DESIGN.md 3 keeps the 700 B/cell of
// SURVEY 8d as the roofline numerator.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "cell_core.cuh"
#include "assemble_args.cuh"

// ---- shared memory ------------------------------------------------------------------------------
#define B3_REC_RUN (27 * 64)                        /* 27 record slots per x-run: 3q+d, x+ of cell q at 3q+5 */
#define B3_SURF_RUN (22 * 64)                       /* records 1..22 (y+) / 0..21 (z+) of a neighbouring run */
#define B3_OFF_TRASH 0                              /* 16 B nobody reads: target of the stores a cell must not make */
#define B3_OFF_OUT 16
#define B3_OFF_IN ((B3_OFF_OUT + B2_OUT_SLOTS * 8 + 127) / 128 * 128) /* input stage (bulk-copy target)   */
#define B3_IN_REC 0
#define B3_IN_YP (B2_RUNS * B3_REC_RUN)             /* y+ records of the last cell row, one run per lk      */
#define B3_IN_ZP (B3_IN_YP + BZ * B3_SURF_RUN)      /* z+ records of the top cell layer, one run per lj     */
#define B3_IN_K (B3_IN_ZP + BY * B3_SURF_RUN)
#define B3_IN_XC (B3_IN_K + BR_THREADS * 72)
#define B3_IN_VOL (B3_IN_XC + BR_THREADS * 24)
#define B3_IN_BYTES (B3_IN_VOL + BR_THREADS * 8)
#define B3_OFF_SMALL (B3_OFF_IN + B3_IN_BYTES)      /* int run_base[16], cell_base[16], extra_rowstart[80]  */
#define B3_OFF_BAR (B3_OFF_SMALL + (2 * B2_RUNS + B2_EXTRA_ROWS) * 4)
#define B3_OFF_IDS (B3_OFF_BAR + 16)                /* int4 {brick, its (bi,bj,bk) packed, next brick, packed} per parity */
#define B3_SMEM (B3_OFF_IDS + 32)
#define B3_THREADS (2 * BR_THREADS)                 /* compute warpgroup + producer warpgroup (one warp of it works) */
#define B3_REGS_PRODUCER 40                         /* setmaxnreg: 2 CTAs x (128 x 216 + 128 x 40) = 64 K registers  */
#define B3_REGS_COMPUTE 216
static_assert(B3_OFF_IN % 128 == 0 && B3_IN_K % 16 == 0 && B3_IN_XC % 16 == 0 && B3_IN_VOL % 16 == 0, "alignment");
static_assert(B3_OFF_BAR % 8 == 0, "mbarrier alignment");

// ---- mbarrier / bulk copy (PTX ISA 8.x, sm_90+) ----------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// barrier of the four compute warps only (the producer warp runs ahead on its own)
__device__ __forceinline__ void sync_compute() { asm volatile("bar.sync 1, %0;" ::"n"(BR_THREADS) : "memory"); }
// the same barrier with an AND-reduction of a predicate over the compute threads
__device__ __forceinline__ bool sync_compute_and(bool v) {
  unsigned r;
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.u32 q, %1, 0;\n"
      "bar.red.and.pred p, 1, %2, q;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(r)
      : "r"((unsigned)v), "n"(BR_THREADS)
      : "memory");
  return r != 0;
}
// bounded: a transfer that never completes raises the device error flag instead of hanging the GPU
template <bool BACKOFF>
__device__ __forceinline__ bool mbar_wait(unsigned bar, unsigned parity, unsigned int* error_flag) {
  unsigned done = 0, spins = 0;
  do {
    if (BACKOFF && spins) __nanosleep(256);  // the producer waits most of the time: keep it off the issue ports
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done && ++spins < (1u << 24));
  if (!done) atomicExch(error_flag, 2u);
  return done != 0;
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// store v at p + OFF bytes unless v is the sentinel (all-ones high word, which no finite value has); one
// address register pair serves every store of a run (the compiler re-derived it per predicated store)
template <int OFF>
__device__ __forceinline__ void st_unless_sentinel(const double* p, double v) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b32 lo, hi;\n"
      "mov.b64 {lo, hi}, %1;\n"
      "setp.ne.s32 p, hi, -1;\n"
      "@p st.global.f64 [%0 + %2], %1;\n"
      "}\n" ::"l"(p),
      "d"(v), "n"(OFF)
      : "memory");
}
// 1 / x without the slow path of the IEEE division: MUFU seed (~20 bits) and two Newton steps (<= 1 ulp);
// x is a pivot or a sum of an SPD block here, far from the denormal range
__device__ __forceinline__ double rcp_nr(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- plan binding of the face records ---------------------------------------------------------------
// bind = { low word: indptr[g(f)] or -1 ; high word: bits 0-5 "present" (flux DOF) flags of the six faces
// of the cell that lists f at local index >= 3 (its x+/y+/z+ face), bits 8-13 those of the cell that lists
// it at local index < 3; 0 where there is no such cell }
__global__ void __launch_bounds__(256)
k_bind_plan_hex(int nf, const int* __restrict__ face_cell, const int* __restrict__ cell_faces,
                const int* __restrict__ f2f, const int* __restrict__ indptr, double* __restrict__ geom) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  unsigned field[2] = {0u, 0u};
  for (int s = 0; s < 2; ++s) {
    const int c = face_cell[2 * (size_t)f + s];
    if (c < 0) continue;
    unsigned bits = 0;
    int li = -1;
    for (int q = 0; q < 6; ++q) {
      const int fq = cell_faces[6 * (size_t)c + q];
      if (fq == f) li = q;
      if (f2f[fq] >= 0) bits |= 1u << q;
    }
    field[li >= 3 ? 0 : 1] = bits;
  }
  const int g = f2f[f];
  const unsigned lo = (unsigned)(g >= 0 ? indptr[g] : -1);
  const unsigned long long w = (unsigned long long)lo | ((unsigned long long)(field[0] | (field[1] << 8)) << 32);
  double2* chunk = reinterpret_cast<double2*>(geom + 8 * (size_t)f) + rec_pos(f, 3);
  chunk->y = __longlong_as_double((long long)w);
}

// position of face j of the cell among the id-sorted union of the faces of the two cells of face i
// (HexMesh numbering, hexmesh.py:161-234; derivation in DESIGN.md 3): index in the merged list
__device__ __forceinline__ constexpr int merged_index(int i, int j) {
  constexpr int Q[6][6] = {{5, 6, 7, 8, 9, 10},   // z- : this cell is the upper one of a z face
                           {4, 5, 6, 7, 8, 10},   // y-
                           {3, 4, 5, 6, 8, 10},   // x-
                           {0, 1, 2, 5, 7, 9},    // x+ : this cell is the lower one of an x face
                           {0, 1, 2, 3, 5, 9},    // y+
                           {0, 1, 2, 3, 4, 5}};   // z+
  return Q[i][j];
}
// present flags of the merged list of face i; p1 = flags of the lower cell (0: none), p2 = upper cell
__device__ __forceinline__ unsigned merged_present(int i, unsigned p1, unsigned p2) {
  if (i == 0 || i == 5) return (p1 & 0x1Fu) | 0x20u | (((p2 >> 1) & 0x1Fu) << 6);
  if (i == 1 || i == 4)
    return (p1 & 0xFu) | ((p2 & 1u) << 4) | 0x20u | (((p2 >> 2) & 7u) << 6) | (((p1 >> 5) & 1u) << 9) | (((p2 >> 5) & 1u) << 10);
  return (p1 & 7u) | ((p2 & 3u) << 3) | 0x20u | (((p2 >> 3) & 1u) << 6) | (((p1 >> 4) & 1u) << 7) | (((p2 >> 4) & 1u) << 8) |
         (((p1 >> 5) & 1u) << 9) | (((p2 >> 5) & 1u) << 10);
}

struct B3Geo {
  int nx, ny, nz, L, xrow;
  int nbx, nby, nsx, nsy, bk_begin, bk_end;
  int total;  // linear brick indices (super-tile order), some of them outside the mesh
  unsigned ticket_base;  // value of the ticket counter when this launch starts
};

__device__ __forceinline__ bool b3_decode(const B3Geo& g, int b, int& bi, int& bj, int& bk) {
  const int st = b >> 6, wi = b & 63;
  bi = (st % g.nsx) * 4 + (wi & 3);
  bj = ((st / g.nsx) % g.nsy) * 4 + ((wi >> 2) & 3);
  bk = g.bk_begin + (st / (g.nsx * g.nsy)) * 4 + (wi >> 4);
  return bi < g.nbx && bj < g.nby && bk < g.bk_end;
}

// the producer warp issues every bulk copy of brick (bi, bj, bk) into the input stage and arms the barrier
__device__ __forceinline__ void b3_issue(const mimpy_mesh_view& m, const B3Geo& g, bool k_unity, int bi, int bj, int bk,
                                         unsigned in_base, unsigned bar, int lane) {
  const int ci0 = bi * BX, cj0 = bj * BY, ck0 = bk * BZ;
  const int runlen = min(BX, g.nx - ci0);
  const bool row_end = ci0 + runlen >= g.nx;  // the run ends at the mesh boundary: its last x+ face follows directly
  unsigned bytes = 0;
  {  // per x-run: lanes 0-15 the records of the z-/y-/x- faces incl. the x+ face of the last cell, and the
     // volumes; lanes 16-31 K and the centroids
    const int r = lane & 15, lj = r & 3, lk = r >> 2;
    const int cj = cj0 + lj, ck = ck0 + lk;
    if (cj < g.ny && ck < g.nz) {
      const int cbase0 = ck * g.L + cj * g.xrow + 3 * ci0;
      const size_t cell0 = (size_t)ci0 + (size_t)g.nx * ((size_t)cj + (size_t)g.ny * ck);
      if (lane < 16) {
        const unsigned nrec = row_end ? 3 * runlen + 1 : 3 * runlen + 3;
        bulk_g2s(in_base + B3_IN_REC + r * B3_REC_RUN, m.face_geom + 8 * (size_t)cbase0, nrec * 64, bar);
        bulk_g2s(in_base + B3_IN_VOL + r * (BX * 8), m.cell_volume + cell0, runlen * 8, bar);
        bytes += nrec * 64 + runlen * 8;
      } else {
        if (!k_unity) {
          bulk_g2s(in_base + B3_IN_K + r * (BX * 72), m.cell_k + 9 * cell0, runlen * 72, bar);
          bytes += runlen * 72;
        }
        bulk_g2s(in_base + B3_IN_XC + r * (BX * 24), m.cell_centroid + 3 * cell0, runlen * 24, bar);
        bytes += runlen * 24;
      }
    }
  }
  if (lane < 8) {  // y+ records of the last cell row (lanes 0-3: one per lk), z+ records of the top layer (4-7: per lj)
    const int l = lane & 3;
    if (lane < 4) {
      const int ck = ck0 + l, cjl = min(cj0 + BY - 1, g.ny - 1);
      if (ck < g.nz) {
        // inside the mesh they are the y- faces (record 3q+1) of the next cell row; on the boundary a run of their own
        const bool inner = cjl + 1 < g.ny;
        const int f0 = inner ? ck * g.L + (cjl + 1) * g.xrow + 3 * ci0 + 1 : ck * g.L + g.ny * g.xrow + ci0;
        const unsigned nrec = inner ? 3 * runlen - 2 : runlen;
        bulk_g2s(in_base + B3_IN_YP + l * B3_SURF_RUN, m.face_geom + 8 * (size_t)f0, nrec * 64, bar);
        bytes += nrec * 64;
      }
    } else {
      const int cj = cj0 + l, ckl = min(ck0 + BZ - 1, g.nz - 1);
      if (cj < g.ny) {
        const bool inner = ckl + 1 < g.nz;
        const int f0 = inner ? (ckl + 1) * g.L + cj * g.xrow + 3 * ci0 : g.nz * g.L + cj * g.nx + ci0;
        const unsigned nrec = inner ? 3 * runlen - 2 : runlen;
        bulk_g2s(in_base + B3_IN_ZP + l * B3_SURF_RUN, m.face_geom + 8 * (size_t)f0, nrec * 64, bar);
        bytes += nrec * 64;
      }
    }
  }
  bytes = __reduce_add_sync(0xffffffffu, bytes);
  if (lane == 0) mbar_arrive_expect_tx(bar, bytes);
}

#ifndef B3_MINB
#define B3_MINB 2
#endif

// ---- producer warpgroup: gives its registers to the compute warpgroup; one warp runs one brick ahead of
// the compute warps.  Bricks are handed out by an atomic ticket in super-tile order, so the bricks in
// flight on the whole GPU stay a compact window however the CTAs drift apart: the two halves of a surface
// row (and the shared face records) meet in L2.
__device__ __forceinline__ void b3_producer(const mimpy_mesh_view& m, const B3Geo& g, bool k_unity, unsigned int* ticket,
                                            unsigned int* error_flag, int4* s_ids, unsigned in_base, unsigned bar_full,
                                            unsigned bar_empty, int warp, int lane) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(B3_REGS_PRODUCER));
  if (warp != BR_THREADS / 32) return;
  int bi = 0, bj = 0, bk = 0;
  auto next_brick = [&]() {  // next valid brick of the global order, or g.total
    int b;
    do {
      unsigned tk = 0;
      if (lane == 0) tk = atomicAdd(ticket, 1u) - g.ticket_base;
      tk = __shfl_sync(0xffffffffu, tk, 0);
      b = tk < (unsigned)g.total ? (int)tk : g.total;
    } while (b < g.total && !b3_decode(g, b, bi, bj, bk));
    return b;
  };
  int b = next_brick();
  unsigned it = 0;
  while (true) {
    const int cbi = bi, cbj = bj, cbk = bk;
    const int bn = b < g.total ? next_brick() : g.total;  // (overwrites bi, bj, bk)
    // the compute warps get the brick coordinates ready-made: the decode (integer divisions) stays here
    const int4 ids = make_int4(b, cbi | (cbj << 10) | (cbk << 20), bn, bi | (bj << 10) | (bk << 20));
    if (it > 0) {
      // the operands of the previous brick are in registers
      if (!mbar_wait<true>(bar_empty, (it - 1) & 1u, error_flag)) return;
      fence_proxy_async();
    }
    if (lane == 0) s_ids[it & 1u] = ids;
    __syncwarp();
    if (b >= g.total) {  // tell the compute warps that the work is done
      if (lane == 0) mbar_arrive_expect_tx(bar_full, 0);
      return;
    }
    b3_issue(m, g, k_unity, cbi, cbj, cbk, in_base, bar_full, lane);
    ++it;
    b = bn;
  }
}


template <int METHOD>
__global__ void __launch_bounds__(B3_THREADS, B3_MINB)
k_numeric_brick3(mimpy_mesh_view m, int flags, UniArgs a, B3Geo g) {
  constexpr int NF = 6;
  extern __shared__ __align__(128) unsigned char sm[];
  double* row = reinterpret_cast<double*>(sm + B3_OFF_OUT);
  const unsigned char* in = sm + B3_OFF_IN;
  int* s_runbase = reinterpret_cast<int*>(sm + B3_OFF_SMALL);
  int* s_cellbase = s_runbase + B2_RUNS;
  int* s_extra_rs = s_cellbase + B2_RUNS;
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
  const unsigned bar_full = sbase + B3_OFF_BAR, bar_empty = bar_full + 8, in_base = sbase + B3_OFF_IN;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int nx = g.nx, ny = g.ny, nz = g.nz, L = g.L, xrow = g.xrow;
  const bool k_unity = flags & 1;

  int4* s_ids = reinterpret_cast<int4*>(sm + B3_OFF_IDS);
  if (t == 0) {
    mbar_init(bar_full, 1);            // the producer's arrive.expect_tx
    mbar_init(bar_empty, BR_THREADS / 32);  // one arrival per compute warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= BR_THREADS / 32) {
    b3_producer(m, g, k_unity, a.ticket, a.error_flag, s_ids, in_base, bar_full, bar_empty, warp, lane);
    return;
  }

  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(B3_REGS_COMPUTE));
  const int li = t % BX, lj = (t / BX) % BY, lk = t / (BX * BY);
  const int run = t >> 3;
  // per-cell scalars that no bulk copy brings: offset of the cell's CSR row, multiplier, C diagonal
  int crow = 0;
  double mult = 1., cdiag = 0.;
  auto prefetch = [&](int pbi, int pbj, int pbk, int& pcrow, double& pmult, double& pcdiag) {
    const int ci = pbi * BX + li, cj = pbj * BY + lj, ck = pbk * BZ + lk;
    pcrow = 0;
    pmult = 1.;
    pcdiag = 0.;
    if (ci < nx && cj < ny && ck < nz) {
      const size_t cell = (size_t)ci + (size_t)nx * ((size_t)cj + (size_t)ny * ck);
      pcrow = __ldg(a.indptr + a.flux_dof + cell);
      if (a.multipliers) pmult = __ldg(a.multipliers + cell);
      if (a.c_diag) pcdiag = __ldg(a.c_diag + cell);
    }
  };
  unsigned phase = 0;
  bool first_brick = true;
  bool image_regular = false;  // every slot of the image that a regular brick does not write holds the sentinel
  double* const out_t = a.vals + t;

  while (true) {
    if (!mbar_wait<false>(bar_full, phase, a.error_flag)) return;  // (uniform: every compute thread waits on the same phase)
    const int4 ids = s_ids[phase];
    phase ^= 1u;
    if (ids.x >= g.total) return;
    const int bi = ids.y & 1023, bj = (ids.y >> 10) & 1023, bk = ids.y >> 20;
    const int bn = ids.z, nbi = ids.w & 1023, nbj = (ids.w >> 10) & 1023, nbk = ids.w >> 20;
    if (first_brick) {
      prefetch(bi, bj, bk, crow, mult, cdiag);
      first_brick = false;
    }
    const int ci0 = bi * BX, cj0 = bj * BY, ck0 = bk * BZ;
    const int ci = ci0 + li, cj = cj0 + lj, ck = ck0 + lk;
    const bool ok = ci < nx && cj < ny && ck < nz;
    const int cbase = ck * L + cj * xrow + 3 * ci;  // global ids of faces z-, y-, x- are cbase + 0, 1, 2
    const int fid[NF] = {cbase, cbase + 1, cbase + 2,
                         ci + 1 < nx ? cbase + 5 : ck * L + cj * xrow + 3 * nx,
                         cj + 1 < ny ? cbase + xrow + 1 : ck * L + ny * xrow + ci,
                         ck + 1 < nz ? cbase + L : nz * L + cj * nx + ci};
    const bool in_brick[NF] = {lk > 0, lj > 0, li > 0, li < BX - 1, lj < BY - 1, lk < BZ - 1};

    // ---- pull: operands into registers ----
    CellOps<NF> o;
    int rs[NF];
    unsigned mp[NF], me = 0, other_mask = 0;  // other_mask bit i: face i has a second cell
    double sa[NF], vol = 0., trk = 1.;
    bool mine_regular = ok;
    if (ok) {
      double K[9], xc[3], sgn[NF];
      double2 rec[NF][4];
      int off[NF];
      off[0] = B3_IN_REC + run * B3_REC_RUN + (3 * li) * 64;
      off[1] = off[0] + 64;
      off[2] = off[0] + 128;
      off[3] = off[0] + (ci + 1 < nx ? 5 : 3) * 64;  // the mesh-boundary x+ face follows the run directly
      off[4] = (lj < BY - 1 && cj + 1 < ny) ? B3_IN_REC + (run + 1) * B3_REC_RUN + (3 * li + 1) * 64
                                            : B3_IN_YP + lk * B3_SURF_RUN + (cj + 1 < ny ? 3 * li : li) * 64;
      off[5] = (lk < BZ - 1 && ck + 1 < nz) ? B3_IN_REC + (run + BY) * B3_REC_RUN + (3 * li) * 64
                                            : B3_IN_ZP + lj * B3_SURF_RUN + (ck + 1 < nz ? 3 * li : li) * 64;
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        const int key = (fid[i] >> 1) & 3;
#pragma unroll
        for (int q = 0; q < 4; ++q) rec[i][q] = *reinterpret_cast<const double2*>(in + off[i] + ((q ^ key) << 4));
      }
      if (k_unity) {
#pragma unroll
        for (int q = 0; q < 9; ++q) K[q] = (q % 4 == 0) ? 1. : 0.;
      } else {
#pragma unroll
        for (int q = 0; q < 9; ++q) K[q] = *reinterpret_cast<const double*>(in + B3_IN_K + t * 72 + q * 8);
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) xc[q] = *reinterpret_cast<const double*>(in + B3_IN_XC + t * 24 + q * 8);
      vol = *reinterpret_cast<const double*>(in + B3_IN_VOL + t * 8);
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        const unsigned long long w = (unsigned long long)__double_as_longlong(rec[i][3].y);
        rs[i] = (int)(unsigned)(w & 0xffffffffull);
        const unsigned hi = (unsigned)(w >> 32);
        const unsigned p1 = hi & 63u, p2 = (hi >> 8) & 63u;  // lower cell, upper cell of the face
        mp[i] = merged_present(i, p1, p2);
        mine_regular = mine_regular && mp[i] == 0x7FFu;
        if (i == 0) me = p2;
        if ((i < 3 ? p1 : p2) != 0u) other_mask |= 1u << i;
        sgn[i] = i < 3 ? -1. : 1.;  // hexmesh.py:329-342
        sa[i] = sgn[i] * rec[i][0].x;
      }
      trk = K[0] + K[4] + K[8];
      cell_prep<NF, METHOD>(K, xc, rec, sgn, o);
    }
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
    // the input stage is consumed: hand it back to the producer, whose copies for the next brick fly
    // during everything below
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty);
    // A REGULAR brick: 128 cells whose six faces are all shared with a cell whose six faces are all flux
    // DOFs (no boundary, no Neumann face in reach).  Every regular brick writes the same set of image
    // slots, at positions known at compile time.
    const bool regular = sync_compute_and(mine_regular);  // (barrier: run bases are visible)
    if (!(regular && image_regular)) {
      // sentinel-fill the output image: a slot nobody in this brick writes is not stored.  Skipped from
      // one regular brick to the next: what the last one wrote is overwritten, the rest still is sentinel.
      const ulonglong2 s2 = make_ulonglong2(AU_SENTINEL, AU_SENTINEL);
      ulonglong2* r2 = reinterpret_cast<ulonglong2*>(sm + B3_OFF_OUT);
#pragma unroll 4
      for (int e = t; e < B2_OUT_SLOTS / 2; e += BR_THREADS) r2[e] = s2;
      if (t < B2_EXTRA_ROWS) s_extra_rs[t] = 0;
      sync_compute();
    }
    image_regular = regular;

    int crow_n = 0;
    double mult_n = 1., cdiag_n = 0.;
    if (bn < g.total) prefetch(nbi, nbj, nbk, crow_n, mult_n, cdiag_n);

    double diag[NF];
    unsigned long long other[NF];
    int sb[NF];  // where slot 0 of the face's row lives in the output image
    auto emit = [&](auto tag) {
      constexpr bool REG = decltype(tag)::value;
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
      if (REG) {
        // regular brick: every position is its index in the merged list; no Neumann test, no popcount
#pragma unroll
        for (int i = 0; i < NF; ++i) {
          row[sb[i] + 11 + (i < 3 ? 1 : 0)] = -sa[i];              // -DIV^T   mfd.py:240-243
          row[cb + i] = sa[i];                                      // DIV      mfd.py:232-238
        }
        row[cb + 6] = a.c_diag ? cdiag : a.alpha * vol;             // mfd.py:766-772
        cell_core<NF, METHOD>(o, trk, vol, mult, [&](int i, int j, double v) {
          if (i == j)
            diag[i] = v;
          else
            row[sb[i] + merged_index(i, j)] = v;
        });
#pragma unroll
        for (int i = 0; i < NF; ++i) {  // diagonals, step 1 (see below): every face is shared
          other[i] = AU_SENTINEL;
          if (in_brick[i]) {
            if (i >= 3) row[sb[i] + 5] = diag[i];
          } else {
            other[i] = atomicExch(reinterpret_cast<unsigned long long*>(a.diag_pub) + fid[i],
                                  (unsigned long long)__double_as_longlong(diag[i]));
          }
        }
        return;
      }
      // Straight-line code from here to the end of cell_core: a store that must not happen (Neumann row or
      // column) goes to the trash word in front of the image (slot index -2 + ... = `dead`) instead of
      // being branched around, so the 36 entry chains of cell_core form ONE basic block and the
      // scheduler can interleave them (with a branch per entry each DMUL-DFMA-DFMA-DADD-DFMA chain ran
      // alone, exposed: 34 % of all stall samples were fixed-latency waits).
      constexpr int dead = -(B3_OFF_OUT / 8);                      // row[dead] = the trash word
      auto slot = [&](int i, int j, unsigned pos) {                 // image slot of entry (i, j), dead if dropped
        const unsigned need = (1u << i) | (1u << j);                // faces i and j are flux DOFs (rs[i] >= 0)
        return ((me & need) == need) ? sb[i] + (int)pos : dead;
      };
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        // -DIV^T: behind the M entries of the row, the lower-index cell first   mfd.py:240-243
        row[slot(i, i, __popc(mp[i]) + ((i < 3 && ((other_mask >> i) & 1u)) ? 1u : 0u))] = -sa[i];
        const int dslot = cb + __popc(me & ((1u << i) - 1u));       // DIV      mfd.py:232-238
        row[((me >> i) & 1u) ? dslot : dead] = sa[i];
      }
      row[cb + __popc(me)] = a.c_diag ? cdiag : a.alpha * vol;      // mfd.py:766-772
      cell_core<NF, METHOD>(o, trk, vol, mult, [&](int i, int j, double v) {
        if (i == j)
          diag[i] = v;
        else
          row[slot(i, j, __popc(mp[i] & ((1u << merged_index(i, j)) - 1u)))] = v;
      });
      // diagonals, step 1: single-cell faces and the LOWER cell of faces inside the brick go to the
      // image.  A face shared with ANOTHER brick: both cells swap their half into diag_pub[f]
      // (all-sentinel between launches); whoever finds the other's half there -- the second to arrive --
      // writes half + half to the CSR slot and restores the sentinel (step 3).  IEEE addition commutes,
      // so the result does not depend on the arrival order: bit-reproducible, and nobody ever waits.
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        other[i] = AU_SENTINEL;
        if (rs[i] < 0) continue;
        const bool shared = (other_mask >> i) & 1u;
        if (!shared || (i >= 3 && in_brick[i])) row[sb[i] + __popc(mp[i] & 0x1Fu)] = diag[i];
        if (shared && !in_brick[i])
          other[i] = atomicExch(reinterpret_cast<unsigned long long*>(a.diag_pub) + fid[i],
                                (unsigned long long)__double_as_longlong(diag[i]));
      }
    };
    if (regular)
      emit(std::true_type{});
    else if (ok)
      emit(std::false_type{});
    sync_compute();
    if (ok) {  // diagonals, step 2: the UPPER cell adds its half (lower + upper, fixed order)
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (rs[i] >= 0 && ((other_mask >> i) & 1u) && in_brick[i]) {
          const int slot = sb[i] + __popc(mp[i] & 0x1Fu);
          row[slot] = row[slot] + diag[i];
        }
      }
    }
    sync_compute();
    // stream the image out; consecutive lanes write consecutive CSR values.  A sentinel has an
    // all-ones high word, which no finite value has.
    {  // face rows of the runs: 312 slots = 128 + 128 + 56 per run, addresses fixed per thread
      static_assert(B2_RUN_SLOTS > 2 * BR_THREADS && B2_RUN_SLOTS <= 3 * BR_THREADS, "three passes per run");
      const double* q = row + t;
      const bool third = t < B2_RUN_SLOTS - 2 * BR_THREADS;
#pragma unroll 4
      for (int r = 0; r < B2_RUNS; ++r, q += B2_RUN_SLOTS) {
        const int base = s_runbase[r];
        if (base == 0x7fffffff) continue;
        const double v0 = q[0], v1 = q[BR_THREADS];
        const double v2 = third ? q[2 * BR_THREADS] : __longlong_as_double(-1ll);
        const double* op = out_t + base;
        st_unless_sentinel<0>(op, v0);
        st_unless_sentinel<BR_THREADS * 8>(op, v1);
        st_unless_sentinel<2 * BR_THREADS * 8>(op, v2);
      }
    }
    {  // cell rows: two runs per pass, 56 of 64 lanes each
      const int half = t >> 6, l = t & 63;
      if (l < B2_CELL_SLOTS) {
#pragma unroll
        for (int r = half; r < B2_RUNS; r += 2) {
          const double v = row[B2_CELL0 + r * B2_CELL_SLOTS + l];
          st_unless_sentinel<0>(a.vals + s_cellbase[r] + l, v);
        }
      }
    }
    if (t < 8 * BR_RW) {  // rows on the brick's upper surfaces: 8 rows x 13 slots per pass
      const int r0 = t / BR_RW, pos = t - r0 * BR_RW;
#pragma unroll
      for (int e = r0; e < B2_EXTRA_ROWS; e += 8) {
        const double v = row[B2_EXTRA0 + e * BR_RW + pos];
        st_unless_sentinel<0>(a.vals + s_extra_rs[e] + pos, v);
      }
    }
    // diagonals, step 3: second arrivals at a face shared with another brick finish the slot
    if (ok) {
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        if (other[i] != AU_SENTINEL) {
          a.vals[rs[i] + __popc(mp[i] & 0x1Fu)] = __longlong_as_double((long long)other[i]) + diag[i];
          reinterpret_cast<unsigned long long*>(a.diag_pub)[fid[i]] = AU_SENTINEL;  // leave the slot clean
        }
      }
    }
    if (bn >= g.total) break;
    sync_compute();  // every read of the image is done before the next brick refills it
    crow = crow_n;
    mult = mult_n;
    cdiag = cdiag_n;
  }
}

bool mimpy_brick3_applies(const mimpy_mesh_view* mesh, const mimpy_symbolic* sym, int method, int flags) {
  const bool structured = mesh->hex_nx > 0 && mesh->hex_ny > 0 && mesh->hex_nz > 0 &&
                          (long long)mesh->hex_nx * mesh->hex_ny * mesh->hex_nz == mesh->nc;
  if (!structured || mesh->uniform_faces != 6 || mesh->hex_nx % 2 != 0) return false;
  // the producer hands the brick coordinates to the compute warps packed 10 + 10 + 11 bits
  if ((mesh->hex_nx + BX - 1) / BX > 1023 || (mesh->hex_ny + BY - 1) / BY > 1023 || (mesh->hex_nz + BZ - 1) / BZ > 2047) return false;
  if (method != 0 && method != 1 && method != 6) return false;
  if (flags & (MIMPY_FLAG_DIRECT_KERNEL | MIMPY_FLAG_GENERIC_KERNEL)) return false;
  if (!sym->face_cell || !sym->face_to_flux || !sym->indptr || sym->serial == 0) return false;
  // bulk copies need 16-byte aligned sources
  const uintptr_t al = (uintptr_t)mesh->face_geom | (uintptr_t)mesh->cell_k | (uintptr_t)mesh->cell_centroid |
                       (uintptr_t)mesh->cell_volume;
  return (al & 15u) == 0;
}

// Writes the plan binding into the face records unless `sym` is the plan already bound to them.
static int bind_plan(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym) {
  if (ctx->bound_geom == (const void*)mesh->face_geom && ctx->bound_serial == (long long)sym->serial) return MIMPY_OK;
  const int T = 256;
  k_bind_plan_hex<<<(mesh->nf + T - 1) / T, T, 0, ctx->stream>>>(mesh->nf, sym->face_cell, mesh->cell_faces, sym->face_to_flux,
                                                                sym->indptr, const_cast<double*>(mesh->face_geom));
  KERNEL_CHECK();
  ctx->bound_geom = mesh->face_geom;
  ctx->bound_serial = (long long)sym->serial;
  return MIMPY_OK;
}

static int b3_ctas_per_sm();

// ---------------------------------------------------------------------------------------------
// Set-up of the hybridised face system on the same pipeline (replaces k_hybrid_setup_hex of hybrid.cu on
// structured hex meshes): inputs by bulk copy, M_E in registers, its inverse W_E in a private column of a
// shared-memory tile (same elimination order as hybrid.cu), then
//   * W_E leaves as contiguous 2304-byte pieces per x-run (36 doubles per cell),
//   * tau_{E,f} = a_f^2 W_E[f,f] (the half transmissibilities the multigrid preconditioner is built from)
//     leave with it, so auxmg.cu does not read W_E again,
//   * the face-system entries a W a - r c^T / S go straight to their SELL-32 slots: the position inside the
//     row is the same closed form as in the assembly kernel, the row's base comes from a per-face array.
// ---------------------------------------------------------------------------------------------
#define HS_LD (BR_THREADS + 1)
struct SetupArgs {
  const int* sell_rowbase;  // per face: sell_ptr[g >> 5] + (g & 31), -1 on faces that are no flux DOF
  int row_stride;           // 32 (SELL) or 1 (CSR: sell_rowbase then holds m_indptr[g])
  const uint8_t* face_flag; // nullable: slab partition, bit0 top / bit1 bottom interface layer
  int nf;
  double* w_e;
  double* tau;              // nullable: nc * 6
  double* a_vals;
  double* dhalf;            // [nf] half of the diagonal entry that the upper cell of a two-cell face contributes
};
// image of a regular brick: [16 runs][12 = 11 entries + the diagonal halves][32: 24 rows + 8 idle], pitch 392 per
// run (bank-conflict free for the stores of a half-warp: 3 li + 8 lj mod 16 are distinct)
#define HS_IMG_K 32
#define HS_IMG_RUN (12 * HS_IMG_K + 8)
#define HS_OFF_WO ((B3_OFF_OUT + B2_RUNS * HS_IMG_RUN * 8 + 15) / 16 * 16)  /* int wo[9][32]: tile offsets of the W copy-out */
static_assert(36 * HS_LD <= B2_RUNS * HS_IMG_RUN, "the image overlays the inversion tile");
static_assert(HS_OFF_WO + 9 * 32 * 4 <= B3_OFF_IN, "image and copy-out table fit in the image region");
static_assert(B2_RUNS * 8 + 3 * B2_RUNS * 4 <= (2 * B2_RUNS + B2_EXTRA_ROWS) * 4, "small arrays of the set-up kernel fit");
#define HS_OFF_ROWB B3_SMEM                       /* int rowb[2][6][128]: behind everything the assembly kernel has */
#define HS_SMEM (HS_OFF_ROWB + 2 * 6 * BR_THREADS * 4)
static_assert(2 * (HS_SMEM + 1024) <= 232448, "two CTAs per SM");
__device__ __forceinline__ void cp_async4(unsigned dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int METHOD, int STRIDE>
__global__ void __launch_bounds__(B3_THREADS, B3_MINB)
k_hybrid_setup_brick(mimpy_mesh_view m, int flags, UniArgs a, SetupArgs sa_, B3Geo g) {
  constexpr int NF = 6;
  extern __shared__ __align__(128) unsigned char sm[];
  double* tile = reinterpret_cast<double*>(sm + B3_OFF_OUT);  // entry e of the cell of thread t at tile[e * HS_LD + t]
  const unsigned char* in = sm + B3_OFF_IN;
  long long* s_cell0 = reinterpret_cast<long long*>(sm + B3_OFF_SMALL);  // first cell of each x-run (-1: no run)
  int* s_rb0 = reinterpret_cast<int*>(s_cell0 + B2_RUNS);  // per run of a regular brick: row base of its first row,
  int* s_rbl = s_rb0 + B2_RUNS;                            //   of its last row - 23 (the run may straddle a slice),
  int* s_cb0 = s_rbl + B2_RUNS;                            //   id of its first face
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
  const unsigned bar_full = sbase + B3_OFF_BAR, bar_empty = bar_full + 8, in_base = sbase + B3_OFF_IN;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int nx = g.nx, ny = g.ny, nz = g.nz, L = g.L, xrow = g.xrow;
  const bool k_unity = flags & 1;
  int4* s_ids = reinterpret_cast<int4*>(sm + B3_OFF_IDS);
  if (t == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_empty, BR_THREADS / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp >= BR_THREADS / 32) {
    b3_producer(m, g, k_unity, a.ticket, a.error_flag, s_ids, in_base, bar_full, bar_empty, warp, lane);
    return;
  }
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(B3_REGS_COMPUTE));
  const int li = t % BX, lj = (t / BX) % BY, lk = t / (BX * BY);
  const int run = t >> 3;
  // W copy-out of a full x-run (288 doubles, nine passes of a warp): element e = lane + 32 p is entry e % 36 of
  // cell e / 36 -- its offset in the tile, tabulated once
  int* s_wo = reinterpret_cast<int*>(sm + HS_OFF_WO);
  for (int e = t; e < 9 * 32; e += BR_THREADS) s_wo[e] = (e % 36) * HS_LD + e / 36;
  // interface layers of a slab partition are interior faces of the global problem, not boundary faces
  const bool has_lo = sa_.face_flag && (__ldg(sa_.face_flag) & 2);
  const bool has_hi = sa_.face_flag && (__ldg(sa_.face_flag + sa_.nf - 1) & 1);
  double* A = tile + t;
  unsigned phase = 0;
  bool first_brick = true;
  double mult = 1., cdiag = 0.;
  // per-cell scalars that no bulk copy brings.  The six row bases travel by cp.async straight into shared memory
  // (double-buffered by brick parity): held in registers across the elimination they were spilled right behind
  // their loads, and the spill store waited for the data.
  const int* s_rowb = reinterpret_cast<const int*>(sm + HS_OFF_ROWB);
  const unsigned rowb_base = sbase + HS_OFF_ROWB + t * 4;
  auto prefetch = [&](int pbi, int pbj, int pbk, double& pmult, double& pcdiag, unsigned par) {
    const int ci = pbi * BX + li, cj = pbj * BY + lj, ck = pbk * BZ + lk;
    pmult = 1.;
    pcdiag = 0.;
    if (ci < nx && cj < ny && ck < nz) {
      const size_t cell = (size_t)ci + (size_t)nx * ((size_t)cj + (size_t)ny * ck);
      if (a.multipliers) pmult = __ldg(a.multipliers + cell);
      if (a.c_diag) pcdiag = __ldg(a.c_diag + cell);
      const int cb = ck * L + cj * xrow + 3 * ci;
      const unsigned dst = rowb_base + par * (NF * BR_THREADS * 4);
      cp_async4(dst, sa_.sell_rowbase + cb);
      cp_async4(dst + BR_THREADS * 4, sa_.sell_rowbase + cb + 1);
      cp_async4(dst + 2 * BR_THREADS * 4, sa_.sell_rowbase + cb + 2);
      cp_async4(dst + 3 * BR_THREADS * 4, sa_.sell_rowbase + (ci + 1 < nx ? cb + 5 : ck * L + cj * xrow + 3 * nx));
      cp_async4(dst + 4 * BR_THREADS * 4, sa_.sell_rowbase + (cj + 1 < ny ? cb + xrow + 1 : ck * L + ny * xrow + ci));
      cp_async4(dst + 5 * BR_THREADS * 4, sa_.sell_rowbase + (ck + 1 < nz ? cb + L : nz * L + cj * nx + ci));
    }
    cp_async_commit();
  };
  unsigned par = 0;  // parity of the brick in this CTA's sequence: which row-base buffer is its own

  while (true) {
    if (!mbar_wait<false>(bar_full, phase, a.error_flag)) return;
    const int4 ids = s_ids[phase];
    phase ^= 1u;
    if (ids.x >= g.total) return;
    const int bi = ids.y & 1023, bj = (ids.y >> 10) & 1023, bk = ids.y >> 20;
    const int bn = ids.z, nbi = ids.w & 1023, nbj = (ids.w >> 10) & 1023, nbk = ids.w >> 20;
    if (first_brick) {
      prefetch(bi, bj, bk, mult, cdiag, par);
      first_brick = false;
    }
    const int ci0 = bi * BX, cj0 = bj * BY, ck0 = bk * BZ;
    const int ci = ci0 + li, cj = cj0 + lj, ck = ck0 + lk;
    const bool ok = ci < nx && cj < ny && ck < nz;
    const long long cell = (long long)ci + (long long)nx * ((long long)cj + (long long)ny * ck);
    const int cbase = ck * L + cj * xrow + 3 * ci;
    const int fid[NF] = {cbase, cbase + 1, cbase + 2,
                         ci + 1 < nx ? cbase + 5 : ck * L + cj * xrow + 3 * nx,
                         cj + 1 < ny ? cbase + xrow + 1 : ck * L + ny * xrow + ci,
                         ck + 1 < nz ? cbase + L : nz * L + cj * nx + ci};
    if (li == 0) s_cell0[run] = (cj < ny && ck < nz) ? cell : -1;

    // ---- pull ----
    CellOps<NF> o;
    unsigned mp[NF], me = 0, other_mask = 0;
    double area[NF], vol = 0., trk = 1.;
    if (ok) {
      double K[9], xc[3], sgn[NF];
      double2 rec[NF][4];
      int off[NF];
      off[0] = B3_IN_REC + run * B3_REC_RUN + (3 * li) * 64;
      off[1] = off[0] + 64;
      off[2] = off[0] + 128;
      off[3] = off[0] + (ci + 1 < nx ? 5 : 3) * 64;
      off[4] = (lj < BY - 1 && cj + 1 < ny) ? B3_IN_REC + (run + 1) * B3_REC_RUN + (3 * li + 1) * 64
                                            : B3_IN_YP + lk * B3_SURF_RUN + (cj + 1 < ny ? 3 * li : li) * 64;
      off[5] = (lk < BZ - 1 && ck + 1 < nz) ? B3_IN_REC + (run + BY) * B3_REC_RUN + (3 * li) * 64
                                            : B3_IN_ZP + lj * B3_SURF_RUN + (ck + 1 < nz ? 3 * li : li) * 64;
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        const int key = (fid[i] >> 1) & 3;
#pragma unroll
        for (int q = 0; q < 4; ++q) rec[i][q] = *reinterpret_cast<const double2*>(in + off[i] + ((q ^ key) << 4));
      }
      if (k_unity) {
#pragma unroll
        for (int q = 0; q < 9; ++q) K[q] = (q % 4 == 0) ? 1. : 0.;
      } else {
#pragma unroll
        for (int q = 0; q < 9; ++q) K[q] = *reinterpret_cast<const double*>(in + B3_IN_K + t * 72 + q * 8);
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) xc[q] = *reinterpret_cast<const double*>(in + B3_IN_XC + t * 24 + q * 8);
      vol = *reinterpret_cast<const double*>(in + B3_IN_VOL + t * 8);
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        const unsigned hi = (unsigned)((unsigned long long)__double_as_longlong(rec[i][3].y) >> 32);
        const unsigned p1 = hi & 63u, p2 = (hi >> 8) & 63u;
        mp[i] = merged_present(i, p1, p2);
        if (i == 0) me = p2;
        if ((i < 3 ? p1 : p2) != 0u) other_mask |= 1u << i;
        sgn[i] = 1.;
        area[i] = rec[i][0].x;
        if (i < 3) {  // outward normal sigma n_f of the z-, y-, x- faces (hexmesh.py:329-342)
          rec[i][0].y = -rec[i][0].y;
          rec[i][1].x = -rec[i][1].x;
          rec[i][1].y = -rec[i][1].y;
        }
      }
      trk = K[0] + K[4] + K[8];
      cell_prep<NF, METHOD>(K, xc, rec, sgn, o);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty);  // the input stage goes back to the producer

    // this brick's row bases have landed (their copies were issued a whole brick ago); then the next brick's
    cp_async_wait_all();
    double mult_n = 1., cdiag_n = 0.;
    if (bn < g.total) prefetch(nbi, nbj, nbk, mult_n, cdiag_n, par ^ 1u);

    // A REGULAR cell: its six faces are flux DOFs shared with a cell whose six faces are flux DOFs as well (no
    // boundary, no Neumann face in reach).  A warp of regular cells takes the instantiation without masks,
    // selects and popcounts: every row position is its index in the merged list (88 % of the cells at 256^3).
    // A brick of 128 regular cells (SELL layout) sends its face-system entries through an IMAGE in shared memory.
    bool mine_regular = ok;
#pragma unroll
    for (int i = 0; i < NF; ++i) mine_regular = mine_regular && mp[i] == 0x7FFu;
    // (barrier: every read of the tile / image by the previous brick is done)
    const bool brick_regular = sync_compute_and(mine_regular) && STRIDE == 32;
    const bool warp_regular = __all_sync(0xffffffffu, mine_regular);
    double R[NF * NF];
    // ---- part 1: M_E, its inverse W (to the tile), aWa, the half transmissibilities ----
    auto invert = [&](auto tag) {
      constexpr bool REG = decltype(tag)::value;
      double rsum[NF], csum[NF];
      // faces that are no flux DOF (Neumann), and boundary faces (one cell, prescribed multiplier)
      const unsigned dropbits = REG ? 0u : (~me & 63u);
      unsigned bndbits = REG ? 0u : (~other_mask & 63u);
      if (!REG) {
        if (has_lo && ck == 0) bndbits &= ~1u;
        if (has_hi && ck == nz - 1) bndbits &= ~32u;
      }
      // mult * M_E acting on the outward fluxes u = sigma o v: the orientations were folded into the normals
      // (cell_prep got sigma n_f and R_f itself), so cell_core delivers M_E and not sigma_i sigma_j M_E
      cell_core<NF, METHOD>(o, trk, vol, mult, [&](int i, int j, double v) {
        const bool drop = !REG && (((dropbits >> i) | (dropbits >> j)) & 1u);
        R[i * NF + j] = drop ? (i == j ? 1. : 0.) : v;
      });
      // Gauss-Jordan without pivoting in registers, same elimination order as group_invert (common.cuh)
#pragma unroll
      for (int k = 0; k < NF; ++k) {
        const double pinv = rcp_nr(R[k * NF + k]);
#pragma unroll
        for (int j = 0; j < NF; ++j) {
          if (j == k) continue;
          const double akj = R[k * NF + j] * pinv;
#pragma unroll
          for (int i = 0; i < NF; ++i) {
            if (i == k) continue;
            R[i * NF + j] -= R[i * NF + k] * akj;
          }
          R[k * NF + j] = akj;
        }
#pragma unroll
        for (int i = 0; i < NF; ++i) {
          if (i == k) continue;
          R[i * NF + k] = -R[i * NF + k] * pinv;
        }
        R[k * NF + k] = pinv;
      }
#pragma unroll
      for (int i = 0; i < NF; ++i) rsum[i] = csum[i] = 0.;
#pragma unroll
      for (int i = 0; i < NF; ++i)
#pragma unroll
        for (int j = 0; j < NF; ++j) {
          const bool drop = !REG && (((dropbits >> i) | (dropbits >> j)) & 1u);
          const double w = drop ? 0. : R[i * NF + j];
          A[(i * NF + j) * HS_LD] = w;   // W leaves through the tile (coalesced copy-out below)
          const double awa = area[i] * w * area[j];
          R[i * NF + j] = awa;
          rsum[i] += awa;
          csum[j] += awa;
        }
      const double cd = a.c_diag ? cdiag : a.alpha * vol;
      double S = cd;
#pragma unroll
      for (int j = 0; j < NF; ++j) S += csum[j];
      const double inv_s = rcp_nr(S);
      if (sa_.tau) {  // tau_{E,f} = a_f^2 W_E[f,f]: 48 contiguous bytes per cell
        double2* tp = reinterpret_cast<double2*>(sa_.tau + cell * NF);
#pragma unroll
        for (int i = 0; i < NF; i += 2)
          tp[i >> 1] = make_double2((REG || ((me >> i) & 1u)) ? R[i * NF + i] : 0.,
                                    (REG || ((me >> (i + 1)) & 1u)) ? R[(i + 1) * NF + i + 1] : 0.);
      }
      // entry (i, j) of the cell's contribution to the face system: aWa[i,j] - rsum[i] csum[j] / S (identity rows
      // and columns for boundary faces, whose multiplier is prescribed)
#pragma unroll
      for (int j = 0; j < NF; ++j) csum[j] *= inv_s;
#pragma unroll
      for (int i = 0; i < NF; ++i)
#pragma unroll
        for (int j = 0; j < NF; ++j) {
          double v = fma(-rsum[i], csum[j], R[i * NF + j]);
          if (!REG && (((bndbits >> i) | (bndbits >> j)) & 1u)) v = (i == j) ? 1. : 0.;
          R[i * NF + j] = v;
        }
    };
    if (warp_regular)
      invert(std::true_type{});
    else if (ok)
      invert(std::false_type{});
    sync_compute();
    // copy-out of W: an x-run of 8 cells is one contiguous piece of w_e (288 doubles): warp w streams the runs
    // 4w .. 4w+3, nine coalesced passes each; element e = lane + 32 p of a run is entry e % 36 of cell e / 36
    {
      const int runlen = min(BX, nx - ci0);
#pragma unroll 1
      for (int rr = 0; rr < 4; ++rr) {
        const int r = warp * 4 + rr;
        const long long c0 = s_cell0[r];
        if (c0 < 0) continue;
        double* dst = sa_.w_e + c0 * (NF * NF) + lane;
        const double* src = tile + r * BX;
        if (runlen == BX) {
#pragma unroll
          for (int p = 0; p < 9; ++p) dst[32 * p] = src[s_wo[32 * p + lane]];
        } else {
          for (int e = lane; e < runlen * NF * NF; e += 32) dst[e - lane] = src[s_wo[e]];
        }
      }
    }
    // ---- part 2: the face-system entries a W a - r c^T / S ----
    int rowb[NF];  // (read here, not before the elimination: six registers less across it)
#pragma unroll
    for (int i = 0; i < NF; ++i) rowb[i] = s_rowb[(par * NF + i) * BR_THREADS + t];
    auto face_id = [&](int i) {  // global id of face i of the cell (HexMesh numbering, hexmesh.py:161-234)
      return i < 3 ? cbase + i
                   : i == 3 ? (ci + 1 < nx ? cbase + 5 : ck * L + cj * xrow + 3 * nx)
                            : i == 4 ? (cj + 1 < ny ? cbase + xrow + 1 : ck * L + ny * xrow + ci)
                                     : (ck + 1 < nz ? cbase + L : nz * L + cj * nx + ci);
    };
    if (!brick_regular) {
      // straight to memory, row by row: position by popcount over the merged face list (constant for a regular cell)
      auto scatter = [&](auto tag) {
        constexpr bool REG = decltype(tag)::value;
#pragma unroll
        for (int i = 0; i < NF; ++i) {
          if (!REG && rowb[i] < 0) continue;
          double* const rowp = sa_.a_vals + rowb[i];
#pragma unroll
          for (int j = 0; j < NF; ++j) {
            if (!REG && !((me >> j) & 1u)) continue;
            const double v = R[i * NF + j];
            const int pos = REG ? merged_index(i, j) : __popc(mp[i] & ((1u << merged_index(i, j)) - 1u));
            if (i != j) {
              rowp[pos * STRIDE] = v;
            } else if (!REG && !((other_mask >> i) & 1u)) {
              rowp[pos * STRIDE] = v;  // the only local cell of the face
              sa_.dhalf[face_id(i)] = 0.;
            } else if (i >= 3) {
              rowp[pos * STRIDE] = v;  // a face with two cells: the half of the lower cell goes to the slot,
            } else {
              sa_.dhalf[face_id(i)] = v;   // the half of the upper cell to a per-face array; k_diag_finish adds them
            }
          }
        }
      };
      if (warp_regular)
        scatter(std::true_type{});
      else if (ok)
        scatter(std::false_type{});
    } else {
      // Regular brick, SELL-32: entry k of the 24 consecutive face rows of an x-run is ONE contiguous piece of the
      // value array (two where the run straddles a slice), but the 8-byte stores of the scatter above reach it from
      // 36 different instructions: 16 sectors per request, every sector written four times (1.9 of the kernel's
      // 4.3 ms at 256^3).  The entries are laid out in shared memory as the array has them -- image[run][k][row],
      // k = 11: the upper cells' halves of the diagonals, bound for dhalf -- and streamed out with consecutive lanes
      // on consecutive values.  Rows on the brick's upper surfaces belong to other runs / bricks: stored directly.
      sync_compute();  // the copy-out has read the tile, which the image overlays
      double* img = tile;
      const int ib = run * HS_IMG_RUN + 3 * li;
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        const bool inside = i < 3 || (i == 3 ? li < BX - 1 : i == 4 ? lj < BY - 1 : lk < BZ - 1);
        const int tb = i < 3 ? ib + i : i == 3 ? ib + 5 : i == 4 ? ib + HS_IMG_RUN + 1 : ib + BY * HS_IMG_RUN;
        if (inside) {
#pragma unroll
          for (int j = 0; j < NF; ++j) {
            const int k = (i == j && i < 3) ? 11 : merged_index(i, j);
            img[tb + k * HS_IMG_K] = R[i * NF + j];
          }
        } else {
          double* const rowp = sa_.a_vals + rowb[i];
#pragma unroll
          for (int j = 0; j < NF; ++j) rowp[merged_index(i, j) * 32] = R[i * NF + j];
        }
      }
      if (li == 0) {
        s_rb0[run] = rowb[0];
        s_cb0[run] = cbase;
      }
      if (li == BX - 1) s_rbl[run] = rowb[2] - 23;
      sync_compute();
      // stream-out: lane = row of the run, one pass per entry k (lanes 24-31 idle).  A slot whose entry comes from the
      // LOWER cell of its face is only there when that cell belongs to the brick (z-: lk > 0, y-: lj > 0, x-: not the
      // first cell of the run); the merged positions of those entries (and 5, the diagonal) per face type:
      const int q = (lane * 11) >> 5, d = lane - 3 * q;  // row = 3 q + d for lane < 24
      const unsigned lower = d == 0 ? 0x03Fu : d == 1 ? 0x22Fu : 0x2A7u;
#pragma unroll 1
      for (int rr = 0; rr < 4; ++rr) {
        const int r = warp * 4 + rr;
        const bool lower_absent = d == 0 ? (r >> 2) == 0 : d == 1 ? (r & 3) == 0 : q == 0;
        const unsigned there = lane < 24 ? (lower_absent ? ~lower & 0x7FFu : 0x7FFu) : 0u;
        const int rb0 = s_rb0[r];
        double* const dst = sa_.a_vals + ((lane < 32 - (rb0 & 31)) ? rb0 : s_rbl[r]) + lane;
        const double* src = img + r * HS_IMG_RUN + lane;
#pragma unroll
        for (int k = 0; k < 11; ++k)
          if ((there >> k) & 1u) dst[32 * k] = src[HS_IMG_K * k];
        if (lane < 24) sa_.dhalf[s_cb0[r] + lane] = src[HS_IMG_K * 11];
      }
    }
    if (bn >= g.total) break;
    mult = mult_n;
    cdiag = cdiag_n;
    par ^= 1u;
  }
}

// the diagonal of the face system: half of the lower cell (already in its slot) + half of the upper cell; the Jacobi
// diagonal with it (INVERT = false: not inverted yet -- the interface rows of a slab partition are completed first)
template <bool INVERT>
__global__ void __launch_bounds__(256)
k_diag_finish(int nf, const int* __restrict__ f2f, const int* __restrict__ rowbase, const uint8_t* __restrict__ diag_pos,
              int stride, const double* __restrict__ dhalf, double* __restrict__ a_vals, double* __restrict__ dinv) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  const int g = f2f[f];
  if (g < 0) return;
  double* slot = a_vals + rowbase[f] + (int)diag_pos[g] * stride;
  const double d = *slot + dhalf[f];
  *slot = d;
  dinv[g] = INVERT ? 1.0 / d : d;
}

// per-face base of the face-system row: SELL-32 slot of entry 0, or the CSR offset
__global__ void __launch_bounds__(256)
k_face_rowbase(int nf, const int* __restrict__ f2f, const int* __restrict__ sell_ptr, const int* __restrict__ m_indptr,
               int* __restrict__ out) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  const int g = f2f[f];
  out[f] = g < 0 ? -1 : (sell_ptr ? sell_ptr[g >> 5] + (g & 31) : m_indptr[g]);
}

int mimpy_face_rowbase(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym, const int* sell_ptr, int* out) {
  // pattern-only: the same (plan, layout, buffer) as last time needs no second pass
  if (sym->serial != 0 && ctx->rowbase_serial == (long long)sym->serial && ctx->rowbase_sell == (const void*)sell_ptr &&
      ctx->rowbase_out == (const void*)out && ctx->rowbase_nf == mesh->nf)
    return MIMPY_OK;
  ctx->rowbase_serial = (long long)sym->serial;
  ctx->rowbase_sell = sell_ptr;
  ctx->rowbase_out = out;
  ctx->rowbase_nf = mesh->nf;
  k_face_rowbase<<<(mesh->nf + 255) / 256, 256, 0, ctx->stream>>>(mesh->nf, sym->face_to_flux, sell_ptr, sym->m_indptr, out);
  KERNEL_CHECK();
  return MIMPY_OK;
}

template <int METHOD, int STRIDE>
static int launch_setup_brick(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, int flags, const UniArgs& a, const SetupArgs& sa,
                              const B3Geo& g, long long grid) {
  CUDA_TRY(cudaFuncSetAttribute(k_hybrid_setup_brick<METHOD, STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HS_SMEM));
  k_hybrid_setup_brick<METHOD, STRIDE><<<(unsigned)grid, B3_THREADS, HS_SMEM, ctx->stream>>>(*mesh, flags, a, sa, g);
  KERNEL_CHECK();
  return MIMPY_OK;
}

static int b3_geometry(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, int bk_begin, int bk_end, B3Geo& g, long long& grid);

// Face-system set-up on the pipelined brick kernel; MIMPY_E_UNSUPPORTED when it does not apply.
int mimpy_hybrid_setup_brick(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym, int method, int flags,
                             const double* multipliers, double alpha, const double* c_diag, const int* face_rowbase,
                             int row_stride, double* w_e, double* tau, double* a_vals, double* a_diag, int invert_diag) {
  if (!mimpy_brick3_applies(mesh, sym, method, flags) || !face_rowbase || !a_diag || !sym->diag_pos) return MIMPY_E_UNSUPPORTED;
  int rc = bind_plan(ctx, mesh, sym);
  if (rc) return rc;
  B3Geo g;
  long long grid = 0;
  rc = b3_geometry(ctx, mesh, 0, 0x3fffffff, g, grid);
  if (rc) return rc;
  if (grid == 0) return MIMPY_OK;
  if (row_stride != 32 && row_stride != 1) return MIMPY_E_UNSUPPORTED;
  void* dhalf = nullptr;  // per-face scratch of the two-contributor diagonals (every slot is written by the kernel)
  rc = mimpy_scratch(ctx, sizeof(double) * (size_t)mesh->nf, &dhalf);
  if (rc) return rc;
  UniArgs a = {};
  a.multipliers = multipliers;
  a.c_diag = c_diag;
  a.alpha = alpha;
  a.ticket = ctx->counters + 12;
  a.error_flag = ctx->counters + 9;
  a.flux_dof = sym->flux_dof;
  SetupArgs sa;
  sa.sell_rowbase = face_rowbase;
  sa.row_stride = row_stride;
  sa.face_flag = sym->face_flag;
  sa.nf = mesh->nf;
  sa.w_e = w_e;
  sa.tau = tau;
  sa.a_vals = a_vals;
  sa.dhalf = static_cast<double*>(dhalf);
#define SETUP_LAUNCH(M)                                                                        \
  (row_stride == 32 ? launch_setup_brick<M, 32>(ctx, mesh, flags, a, sa, g, grid)               \
                    : launch_setup_brick<M, 1>(ctx, mesh, flags, a, sa, g, grid))
  rc = method == 0 ? SETUP_LAUNCH(0) : method == 1 ? SETUP_LAUNCH(1) : SETUP_LAUNCH(6);
#undef SETUP_LAUNCH
  if (rc) return rc;
  const int blocks = (mesh->nf + 255) / 256;
  if (invert_diag)
    k_diag_finish<true><<<blocks, 256, 0, ctx->stream>>>(mesh->nf, sym->face_to_flux, face_rowbase, sym->diag_pos, row_stride,
                                                        sa.dhalf, a_vals, a_diag);
  else
    k_diag_finish<false><<<blocks, 256, 0, ctx->stream>>>(mesh->nf, sym->face_to_flux, face_rowbase, sym->diag_pos, row_stride,
                                                         sa.dhalf, a_vals, a_diag);
  KERNEL_CHECK();
  return MIMPY_OK;
}

static int b3_ctas_per_sm() {
  static int per_sm = 0;
  if (per_sm == 0) {
    const char* e = getenv("MIMPY_B200_B3_CTAS");
    per_sm = e ? atoi(e) : B3_MINB;
    if (per_sm < 1) per_sm = 1;
  }
  return per_sm;
}

template <int METHOD>
static int launch_brick3(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, int flags, const UniArgs& a, const B3Geo& g) {
  CUDA_TRY(cudaFuncSetAttribute(k_numeric_brick3<METHOD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B3_SMEM));
  long long grid = (long long)ctx->sm_count * b3_ctas_per_sm();
  if (grid > g.total) grid = g.total;
  k_numeric_brick3<METHOD><<<(unsigned)grid, B3_THREADS, B3_SMEM, ctx->stream>>>(*mesh, flags, a, g);
  KERNEL_CHECK();
  return MIMPY_OK;
}

static int b3_geometry(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, int bk_begin, int bk_end, B3Geo& g, long long& grid) {
  g.nx = mesh->hex_nx;
  g.ny = mesh->hex_ny;
  g.nz = mesh->hex_nz;
  g.L = g.nx * g.ny + g.nx * (g.ny + 1) + (g.nx + 1) * g.ny;  // nf < 2^31 (checked by hex_topology)
  g.xrow = 3 * g.nx + 1;
  g.nbx = (g.nx + BX - 1) / BX;
  g.nby = (g.ny + BY - 1) / BY;
  const int nbz = (g.nz + BZ - 1) / BZ;
  g.nsx = (g.nbx + 3) >> 2;
  g.nsy = (g.nby + 3) >> 2;
  g.bk_begin = bk_begin;
  g.bk_end = bk_end < nbz ? bk_end : nbz;
  const long long super_tiles = (long long)g.nsx * g.nsy * ((g.bk_end - g.bk_begin + 3) / 4);
  grid = 0;
  g.total = 0;
  if (super_tiles <= 0) return MIMPY_OK;
  if (super_tiles * 64 > 0x7fffffffLL) return MIMPY_E_UNSUPPORTED;
  g.total = (int)(super_tiles * 64);
  // every CTA's producer draws tickets until it gets one past the end: total + grid draws per launch
  grid = (long long)ctx->sm_count * b3_ctas_per_sm();
  if (grid > g.total) grid = g.total;
  g.ticket_base = ctx->b3_ticket;
  ctx->b3_ticket += (unsigned)g.total + (unsigned)grid;
  return MIMPY_OK;
}

int mimpy_numeric_brick3(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym, int method,
                         int flags, const UniArgs& a, int bk_begin, int bk_end) {
  int rc = bind_plan(ctx, mesh, sym);
  if (rc) return rc;
  B3Geo g;
  long long grid = 0;
  rc = b3_geometry(ctx, mesh, bk_begin, bk_end, g, grid);
  if (rc) return rc;
  if (grid == 0) return MIMPY_OK;
  if (method == 0) return launch_brick3<0>(ctx, mesh, flags, a, g);
  if (method == 1) return launch_brick3<1>(ctx, mesh, flags, a, g);
  return launch_brick3<6>(ctx, mesh, flags, a, g);
}
