// Symbolic assembly: every index-valued decision of MFD.renumber_for_neumann (mfd.py:141-159),
// build_m (545-599), build_div (203-246), build_bottom_right (745-774) and
// coo_matrix(...).tocsr() (906-908), computed once per (mesh, Neumann set) on the device.
//
// Row g(f) of the M block holds the union of the flux DOFs of the (<=2) cells that list f, so a
// thread per face row can build its sorted column list locally (<= 2*32 candidates) -- no global
// sort of the 36*nc COO triples is needed.
#include <cub/device/device_scan.cuh>

#include <atomic>

#include "common.cuh"

static std::atomic<long long> g_plan_serial{0};

#define MAXC (2 * MIMPY_MAX_CELL_FACES)

__global__ void k_adj_init(int nf, int* __restrict__ face_cell) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  face_cell[2 * (size_t)f] = 0x7fffffff;
  face_cell[2 * (size_t)f + 1] = -1;
}

__global__ void k_adj_mark(int nc, const int* __restrict__ cell_ptr,
                           const int* __restrict__ cell_faces, int* __restrict__ face_cell) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  for (int p = cell_ptr[c]; p < cell_ptr[c + 1]; ++p) {
    const int f = cell_faces[p];
    atomicMin(&face_cell[2 * (size_t)f], c);
    atomicMax(&face_cell[2 * (size_t)f + 1], c);
  }
}

__global__ void k_adj_fix(int nf, int* __restrict__ face_cell, const uint8_t* __restrict__ neumann,
                          int* __restrict__ flag) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  int a = face_cell[2 * (size_t)f], b = face_cell[2 * (size_t)f + 1];
  if (a == 0x7fffffff) a = -1;
  if (b == a) b = -1;
  face_cell[2 * (size_t)f] = a;
  face_cell[2 * (size_t)f + 1] = b;
  flag[f] = neumann[f] ? 0 : 1;
}

// after the exclusive scan of flag: Neumann faces get -1 (mfd.py:151-155)
__global__ void k_flux_fix(int nf, const uint8_t* __restrict__ neumann, int* __restrict__ f2f) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  if (neumann[f]) f2f[f] = -1;
}

// sorted unique flux DOFs of the cells adjacent to face f; returns count
__device__ int row_columns(int f, const int* __restrict__ cell_ptr,
                           const int* __restrict__ cell_faces, const int* __restrict__ f2f,
                           const int* __restrict__ face_cell, int* cols, int* n_dup = nullptr) {
  int n = 0, dups = 0;
  for (int s = 0; s < 2; ++s) {
    const int c = face_cell[2 * (size_t)f + s];
    if (c < 0) continue;
    for (int p = cell_ptr[c]; p < cell_ptr[c + 1]; ++p) {
      const int g = f2f[cell_faces[p]];
      if (g < 0) continue;
      // insertion into the sorted unique list
      int k = n;
      bool dup = false;
      while (k > 0 && cols[k - 1] >= g) {
        if (cols[k - 1] == g) {
          dup = true;
          break;
        }
        --k;
      }
      if (dup) {
        ++dups;
        continue;
      }
      for (int m = n; m > k; --m) cols[m] = cols[m - 1];
      cols[k] = g;
      ++n;
    }
  }
  if (n_dup) *n_dup = dups;
  return n;
}

// row lengths: rows [0,F) = flux rows (M part + one -DIV^T entry per adjacent cell),
// rows [F, F+nc) = cell rows (DIV entries + the C diagonal). len_m: M-only lengths.
__global__ void k_row_len_faces(int nf, const int* __restrict__ cell_ptr,
                                const int* __restrict__ cell_faces, const int* __restrict__ f2f,
                                const int* __restrict__ face_cell, int* __restrict__ len,
                                int* __restrict__ len_m, int* __restrict__ multi_shared) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  const int g = f2f[f];
  if (g < 0) return;
  int cols[MAXC];
  int dups = 0;
  const int n = row_columns(f, cell_ptr, cell_faces, f2f, face_cell, cols, &dups);
  // the two cells of f share f itself; any further common flux face means an entry M[g(f), g(f')] with two
  // contributors, which the numeric kernels (one writer per off-diagonal slot) do not sum
  if (dups > 1) *multi_shared = 1;
  const int ncell = (face_cell[2 * (size_t)f] >= 0) + (face_cell[2 * (size_t)f + 1] >= 0);
  len_m[g] = n;
  len[g] = n + ncell;
}

__global__ void k_row_len_cells(int nc, int F, const int* __restrict__ cell_ptr,
                                const int* __restrict__ cell_faces, const int* __restrict__ f2f,
                                int* __restrict__ len, int* __restrict__ blk) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  int n = 0;
  for (int p = cell_ptr[c]; p < cell_ptr[c + 1]; ++p) n += f2f[cell_faces[p]] >= 0;
  len[F + c] = n + 1;
  const int nf_c = cell_ptr[c + 1] - cell_ptr[c];
  blk[c] = nf_c * nf_c;
}

__global__ void k_fill_faces(int nf, int F, const int* __restrict__ cell_ptr,
                             const int* __restrict__ cell_faces, const int* __restrict__ f2f,
                             const int* __restrict__ face_cell, const int* __restrict__ indptr,
                             int* __restrict__ indices, const int* __restrict__ m_indptr,
                             int* __restrict__ m_indices, uint8_t* __restrict__ diag_pos) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  const int g = f2f[f];
  if (g < 0) return;
  int cols[MAXC];
  const int n = row_columns(f, cell_ptr, cell_faces, f2f, face_cell, cols);
  int* dst = indices + indptr[g];
  int* dm = m_indices + m_indptr[g];
  for (int k = 0; k < n; ++k) {
    dst[k] = cols[k];
    dm[k] = cols[k];
    if (cols[k] == g) diag_pos[g] = (uint8_t)k;
  }
  int k = n;
  const int c0 = face_cell[2 * (size_t)f], c1 = face_cell[2 * (size_t)f + 1];
  if (c0 >= 0) dst[k++] = F + c0;  // c0 < c1 by construction
  if (c1 >= 0) dst[k++] = F + c1;
}

// one thread per (cell, local face i): cell row columns, pos_d, pos_dt and the pos_m row
__global__ void k_fill_cells(int nc, int F, const int* __restrict__ cell_ptr,
                             const int* __restrict__ cell_faces, const int* __restrict__ f2f,
                             const int* __restrict__ face_cell, const int* __restrict__ indptr,
                             int* __restrict__ indices, const int* __restrict__ m_indptr,
                             const int* __restrict__ m_indices, const int* __restrict__ m_ptr,
                             uint8_t* __restrict__ pos_m, uint8_t* __restrict__ pos_dt,
                             uint8_t* __restrict__ pos_d, uint8_t* __restrict__ role,
                             int* __restrict__ cell_rowstart) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= cell_ptr[nc]) return;
  // find the cell of entry t: binary search in cell_ptr
  int lo = 0, hi = nc;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (cell_ptr[mid] <= t)
      lo = mid;
    else
      hi = mid;
  }
  const int c = lo;
  const int base = cell_ptr[c];
  const int n = cell_ptr[c + 1] - base;
  const int i = (int)(t - base);
  const int fi = cell_faces[base + i];
  const int gi = f2f[fi];
  uint8_t* pm = pos_m + m_ptr[c] + (size_t)i * n;
  if (role) {
    const int c_lo = face_cell[2 * (size_t)fi], c_hi = face_cell[2 * (size_t)fi + 1];
    role[base + i] = (c_hi < 0) ? 0 : (c == c_lo ? 1 : 2);
  }
  if (cell_rowstart) cell_rowstart[base + i] = (gi >= 0) ? indptr[gi] : -1;
  if (gi < 0) {
    for (int j = 0; j < n; ++j) pm[j] = 0xFF;
    pos_dt[base + i] = 0xFF;
    pos_d[base + i] = 0xFF;
  } else {
    const int r0 = m_indptr[gi];
    const int rl = m_indptr[gi + 1] - r0;
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const int gj = f2f[cell_faces[base + j]];
      if (gj < 0) {
        pm[j] = 0xFF;
        continue;
      }
      rank += (gj < gi);
      int a = 0, b = rl;  // lower bound of gj in the sorted row
      while (a < b) {
        const int mid = (a + b) >> 1;
        if (m_indices[r0 + mid] < gj)
          a = mid + 1;
        else
          b = mid;
      }
      pm[j] = (uint8_t)a;
    }
    pos_dt[base + i] = (uint8_t)(rl + (face_cell[2 * (size_t)fi] == c ? 0 : 1));
    pos_d[base + i] = (uint8_t)rank;
    indices[indptr[F + c] + rank] = gi;
  }
  if (i == 0) indices[indptr[F + c + 1] - 1] = F + c;
}

static int scan_i32(mimpy_ctx* ctx, const int* in, int* out, int n) {
  size_t bytes = 0;
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, ctx->stream));
  void* tmp = nullptr;
  int rc = mimpy_scratch(ctx, bytes, &tmp);
  if (rc) return rc;
  CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, n, ctx->stream));
  return MIMPY_OK;
}

extern "C" {

int mimpy_b200_mfd_symbolic_sizes(mimpy_ctx* ctx, const mimpy_mesh_view* mesh,
                                  const uint8_t* neumann_mask, mimpy_symbolic* sym) {
  REQUIRE(ctx && mesh && neumann_mask && sym, "NULL argument");
  REQUIRE(sym->face_to_flux && sym->face_cell && sym->indptr && sym->m_indptr && sym->m_ptr,
          "sizes pass needs face_to_flux, face_cell, indptr, m_indptr, m_ptr");
  REQUIRE(mesh->max_faces >= 1 && mesh->max_faces <= MIMPY_MAX_CELL_FACES, "max_faces out of range");
  const int nf = mesh->nf, nc = mesh->nc;
  const int T = 256;
  cudaStream_t st = ctx->stream;
  // temporaries: flag[nf+1], len[nf+nc+1], len_m[nf+1], blk[nc+1] -- carved from one cudaMalloc
  int* tmp = nullptr;
  const size_t n_tmp = (size_t)(nf + 1) + (size_t)(nf + nc + 1) + (size_t)(nf + 1) + (size_t)(nc + 1);
  CUDA_TRY(cudaMalloc(&tmp, n_tmp * sizeof(int)));
  int* flag = tmp;
  int* len = flag + nf + 1;
  int* len_m = len + nf + nc + 1;
  int* blk = len_m + nf + 1;
  int rc = MIMPY_OK;
  do {
    if (cudaMemsetAsync(tmp, 0, n_tmp * sizeof(int), st) != cudaSuccess) { rc = MIMPY_E_CUDA; break; }
    k_adj_init<<<(nf + T - 1) / T, T, 0, st>>>(nf, sym->face_cell);
    k_adj_mark<<<(nc + T - 1) / T, T, 0, st>>>(nc, mesh->cell_ptr, mesh->cell_faces, sym->face_cell);
    k_adj_fix<<<(nf + T - 1) / T, T, 0, st>>>(nf, sym->face_cell, neumann_mask, flag);
    if ((rc = scan_i32(ctx, flag, sym->face_to_flux, nf + 1))) break;  // [nf] = flux_dof
    int F = 0;
    // face_to_flux has nf entries only: scan into a temp when the caller's array is nf long
    // (we scanned nf+1 items into face_to_flux, so require nf+1 capacity -- documented)
    if (cudaMemcpyAsync(&F, sym->face_to_flux + nf, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) { rc = MIMPY_E_CUDA; break; }
    if (cudaStreamSynchronize(st) != cudaSuccess) { rc = MIMPY_E_CUDA; break; }
    k_flux_fix<<<(nf + T - 1) / T, T, 0, st>>>(nf, neumann_mask, sym->face_to_flux);
    k_row_len_faces<<<(nf + T - 1) / T, T, 0, st>>>(nf, mesh->cell_ptr, mesh->cell_faces,
                                                    sym->face_to_flux, sym->face_cell, len, len_m, flag + nf);
    k_row_len_cells<<<(nc + T - 1) / T, T, 0, st>>>(nc, F, mesh->cell_ptr, mesh->cell_faces,
                                                    sym->face_to_flux, len, blk);
    if ((rc = scan_i32(ctx, len, sym->indptr, F + nc + 1))) break;
    if ((rc = scan_i32(ctx, len_m, sym->m_indptr, F + 1))) break;
    if ((rc = scan_i32(ctx, blk, sym->m_ptr, nc + 1))) break;
    int nnz = 0, nnz_m = 0, nb = 0, multi = 0;
    cudaMemcpyAsync(&multi, flag + nf, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&nnz, sym->indptr + F + nc, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&nnz_m, sym->m_indptr + F, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&nb, sym->m_ptr + nc, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) { rc = MIMPY_E_CUDA; break; }
    if (cudaGetLastError() != cudaSuccess) { rc = MIMPY_E_CUDA; break; }
    sym->flux_dof = F;
    sym->nnz = nnz;
    sym->nnz_m = nnz_m;
    sym->n_blocks = nb;
    if (nnz < 0 || nb < 0) {
      mimpy_set_error("saddle matrix exceeds int32 indexing (nnz >= 2^31)");
      rc = MIMPY_E_UNSUPPORTED;
    } else if (multi) {
      mimpy_set_error("two cells of the mesh share more than one face: an off-diagonal entry of M would have two "
                      "contributors, which the assembly kernels do not sum");
      rc = MIMPY_E_UNSUPPORTED;
    }
  } while (0);
  if (rc == MIMPY_E_CUDA) mimpy_set_error("symbolic_sizes: CUDA failure: %s", cudaGetErrorString(cudaGetLastError()));
  cudaFree(tmp);
  return rc;
}

int mimpy_b200_mfd_symbolic_fill(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, mimpy_symbolic* sym) {
  REQUIRE(ctx && mesh && sym, "NULL argument");
  REQUIRE(sym->indices && sym->m_indices && sym->pos_m && sym->pos_dt && sym->pos_d && sym->diag_pos,
          "fill pass needs indices, m_indices, pos_m, pos_dt, pos_d, diag_pos");
  const int nf = mesh->nf, nc = mesh->nc, F = sym->flux_dof;
  const int T = 256;
  cudaStream_t st = ctx->stream;
  sym->serial = ++g_plan_serial;
  k_fill_faces<<<(nf + T - 1) / T, T, 0, st>>>(nf, F, mesh->cell_ptr, mesh->cell_faces,
                                               sym->face_to_flux, sym->face_cell, sym->indptr,
                                               sym->indices, sym->m_indptr, sym->m_indices,
                                               sym->diag_pos);
  KERNEL_CHECK();
  int total = 0;
  CUDA_TRY(cudaMemcpyAsync(&total, mesh->cell_ptr + nc, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (total > 0) {
    const long long blocks = ((long long)total + T - 1) / T;
    k_fill_cells<<<(unsigned)blocks, T, 0, st>>>(nc, F, mesh->cell_ptr, mesh->cell_faces,
                                                 sym->face_to_flux, sym->face_cell, sym->indptr,
                                                 sym->indices, sym->m_indptr, sym->m_indices,
                                                 sym->m_ptr, sym->pos_m, sym->pos_dt, sym->pos_d,
                                                 sym->role, sym->cell_rowstart);
    KERNEL_CHECK();
  }
  return MIMPY_OK;
}

}  // extern "C"
