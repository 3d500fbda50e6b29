// Right-hand side of the saddle system: MFD.build_rhs (mfd.py:991-999) with its three parts
// build_rhs_neumann (1035-1041), build_rhs_dirichlet (1209-1215), build_rhs_forcing (1027-1033);
// and the per-step terms of SinglePhase.start_solving (singlephase.py:256-287).
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

// singlephase.py:263-281: density = ((p - p_ref) c + 1) rho_ref ; multiplier = mu / density ;
// c_entry = c phi / dt |E| ; rhs[F+c] += c_entry p
__global__ void k_singlephase_terms(int nc, int F, double ref_p, double compressibility, double ref_rho,
                                    double mu, double dt, const double* __restrict__ p,
                                    const double* __restrict__ porosity, const double* __restrict__ vol,
                                    double* __restrict__ rhs, double* __restrict__ mult,
                                    double* __restrict__ c_diag) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  double density = -ref_p;
  density += p[c];
  density *= compressibility;
  density += 1.;
  density *= ref_rho;
  mult[c] = mu / density;
  double ce = compressibility;
  ce *= porosity[c];
  ce /= dt;
  ce *= vol[c];
  rhs[F + c] += ce * p[c];
  c_diag[c] = ce;
}

extern "C" {

int mimpy_b200_mfd_rhs(mimpy_ctx* ctx, int32_t nc, int32_t flux_dof, const int32_t* face_to_flux,
                       int32_t n_dirichlet, const int32_t* dirichlet_faces,
                       const double* dirichlet_values, int32_t has_neumann, double last_neumann_value,
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

int mimpy_b200_singlephase_terms(mimpy_ctx* ctx, int32_t nc, int32_t flux_dof, double ref_pressure,
                                 double compressibility, double ref_density, double viscosity,
                                 double dt, const double* pressure, const double* porosity,
                                 const double* cell_volume, double* rhs, double* multipliers,
                                 double* c_diag) {
  REQUIRE(ctx && pressure && porosity && cell_volume && rhs && multipliers && c_diag, "NULL argument");
  if (nc <= 0) return MIMPY_OK;
  k_singlephase_terms<<<(nc + 255) / 256, 256, 0, ctx->stream>>>(nc, flux_dof, ref_pressure, compressibility,
                                                                 ref_density, viscosity, dt, pressure, porosity,
                                                                 cell_volume, rhs, multipliers, c_diag);
  KERNEL_CHECK();
  return MIMPY_OK;
}

}  // extern "C"
