// Right-hand side of the saddle system: MFD.build_rhs (mfd.py:991-999) with its three parts
// build_rhs_neumann (1035-1041), build_rhs_dirichlet (1209-1215), build_rhs_forcing (1027-1033).
#include "common.cuh"

__global__ void k_rhs_dirichlet(int n, const int* __restrict__ faces, const double* __restrict__ values,
                                const int* __restrict__ f2f, int ndof, double* __restrict__ rhs) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int g = f2f[faces[k]];
  if (g < 0) g += ndof;  // numpy negative index (a Dirichlet value set on a Neumann face)
  rhs[g] = -values[k];
}

__global__ void k_rhs_forcing(int nc, int F, const double* __restrict__ forcing, double* __restrict__ rhs) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  rhs[F + c] += forcing[c];
}

__global__ void k_set_one(double* p, double v) { *p = v; }

extern "C" int mimpy_b200_mfd_rhs(mimpy_ctx* ctx, int32_t nc, int32_t flux_dof,
                                  const int32_t* face_to_flux, int32_t n_dirichlet,
                                  const int32_t* dirichlet_faces, const double* dirichlet_values,
                                  int32_t has_neumann, double last_neumann_value,
                                  const double* forcing, double* rhs) {
  REQUIRE(ctx && face_to_flux && rhs, "NULL argument");
  const int ndof = flux_dof + nc;
  if (ndof <= 0) return MIMPY_OK;
  cudaStream_t st = ctx->stream;
  CUDA_TRY(cudaMemsetAsync(rhs, 0, sizeof(double) * (size_t)ndof, st));
  // every Neumann face has face_to_flux = -1, so each write of mfd.py:1041 lands on rhs[-1];
  // what survives is the value of the last face in insertion order
  if (has_neumann) k_set_one<<<1, 1, 0, st>>>(rhs + ndof - 1, last_neumann_value);
  if (n_dirichlet > 0) {
    REQUIRE(dirichlet_faces && dirichlet_values, "Dirichlet arrays are NULL");
    k_rhs_dirichlet<<<(n_dirichlet + 255) / 256, 256, 0, st>>>(n_dirichlet, dirichlet_faces,
                                                               dirichlet_values, face_to_flux, ndof, rhs);
  }
  if (forcing && nc > 0) k_rhs_forcing<<<(nc + 255) / 256, 256, 0, st>>>(nc, flux_dof, forcing, rhs);
  KERNEL_CHECK();
  return MIMPY_OK;
}
