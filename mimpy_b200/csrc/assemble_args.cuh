// Arguments and brick geometry shared by the hexahedral numeric-assembly kernels
// (assemble_uniform.cu: direct / brick / staged brick; assemble_brick3.cu: pipelined brick).
#pragma once
#include "common.cuh"

#define AU_SENTINEL 0xFFFFFFFFFFFFFFFFull  // memset(0xFF): a NaN no arithmetic produces

struct UniArgs {
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

// a CTA owns a brick of BX x BY x BZ cells, one thread per cell
#define BX 8
#define BY 4
#define BZ 4
#define BR_THREADS (BX * BY * BZ)
#define BR_NZF (BX * BY * (BZ + 1))
#define BR_NYF (BX * (BY + 1) * BZ)
#define BR_NXF ((BX + 1) * BY * BZ)
#define BR_NFACE (BR_NZF + BR_NYF + BR_NXF)
#define BR_RW 13  // 11 M entries of an interior row + 2 (-DIV^T); longer rows cannot occur on a hex mesh

// Output image of a brick in shared memory, laid out like the CSR value array itself:
//   16 x-runs x 312 slots : the 24 face rows (z-, y-, x- of 8 cells) of a run are CONSECUTIVE CSR rows,
//                           so a run is one contiguous piece of `vals` (<= 24 * 13 values)
//   80 extra rows x 13    : x+ faces of the last cell column, y+ of the last row, z+ of the top layer
//   16 x-runs x 56 slots  : the cell rows (DIV + C diagonal, <= 7 values per cell) of a run
#define B2_RUNS (BY * BZ)
#define B2_RUN_SLOTS (3 * BX * BR_RW)
#define B2_EXTRA_ROWS (BY * BZ + BX * BZ + BX * BY)
#define B2_EXTRA0 (B2_RUNS * B2_RUN_SLOTS)
#define B2_CELL0 (B2_EXTRA0 + B2_EXTRA_ROWS * BR_RW)
#define B2_CELL_SLOTS (BX * 7)
#define B2_OUT_SLOTS (B2_CELL0 + B2_RUNS * B2_CELL_SLOTS)
static_assert(B2_EXTRA0 + B2_EXTRA_ROWS * BR_RW == BR_NFACE * BR_RW, "row image holds every face row of the brick");
static_assert(B2_OUT_SLOTS % 2 == 0, "output image is a whole number of 16-byte words");

int mimpy_numeric_brick3(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym, int method,
                         int flags, const UniArgs& a, int bk_begin, int bk_end);
bool mimpy_brick3_applies(const mimpy_mesh_view* mesh, const mimpy_symbolic* sym, int method, int flags);
