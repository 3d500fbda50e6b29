// Fused numeric assembly for hexahedral meshes (every cell has 6 faces): dispatch and the two kernels that
// serve the meshes the headline kernel does not.  All share one thread-per-cell core (cell_core.cuh).
//
//  k_numeric_brick3  (assemble_brick3.cu; structured hex meshes in HexMesh numbering, even nx; the default):
//     persistent CTAs, bulk-copy staging through an mbarrier pipeline, closed-form CSR positions, output image
//     in shared memory laid out like the CSR value array, cross-brick diagonals by a wait-free swap.
//
//  k_numeric_brick   (the same bricks for odd nx): per-thread loads; the 464 face rows of the brick
//     are assembled in a shared-memory row buffer of 13 slots per row and written as whole rows.
//
//  k_numeric_direct  (any uniform-6 mesh): each thread stores its 36+13 entries straight from
//     registers to their CSR slots (48 write sectors per cell: the L2 is the busiest unit).
//
// Diagonal entries M[g,g] have two contributors (the two cells of face f).  Inside a brick they are
// added in shared memory in a fixed order.  Across CTAs the two kernels of this file let the lower-index
// cell PUBLISH its half in diag_pub[f] and the higher-index cell add it (CTAs take their work from
// an atomic ticket, bricks ordered lexicographically, so a CTA only ever waits for CTAs that started
// earlier and are resident; the wait is bounded and raises an error flag instead of hanging).
// k_numeric_brick3 replaces the wait by an atomic swap of the halves.
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

// 0 direct, 1 brick, 3 pipelined brick (bulk copies, assemble_brick3.cu; the default where the mesh is a
// structured hex mesh with even nx).  (Variant 2, the cp.async-staged brick of round 1, is gone: the
// pipelined brick replaces it.)
static int asm_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MIMPY_B200_ASM_VARIANT");
    v = (e && e[0] == 'd') ? 0 : (e && e[0] == 'b') ? 1 : 3;
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
  (void)bk_begin;
  (void)bk_end;
  (void)staged;
  if (structured && asm_variant() >= 1 && !(flags & MIMPY_FLAG_DIRECT_KERNEL)) {
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
int mimpy_swap_prepare(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, bool ranged, int layer_begin, int layer_end) {
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
  const bool staged = false;  // (the cp.async-staged brick of round 1 was replaced by the pipelined brick)
  // a range of cell layers [layer_begin, layer_end) is only served by the (staged / pipelined) brick kernels
  const bool ranged = !(layer_begin == 0 && layer_end < 0);
  if (ranged && ((!staged && !pipelined) || layer_begin % BZ != 0 || (layer_end % BZ != 0 && layer_end != mesh->hex_nz)))
    return MIMPY_E_UNSUPPORTED;
  const int bk_begin = ranged ? layer_begin / BZ : 0;
  const int bk_end = ranged ? (layer_end + BZ - 1) / BZ : 0x3fffffff;
  double* exchange = nullptr;
  if (staged || pipelined) {
    int rc = mimpy_swap_prepare(ctx, mesh, ranged && !(layer_begin == 0 && layer_end >= mesh->hex_nz), layer_begin, layer_end);
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
