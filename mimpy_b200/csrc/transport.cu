// Explicit upwind update of a transported concentration (reference transport.py:303-361; SURVEY 8f
// rank 2).  The reference builds a list of (face, upwind cell) pairs (find_upwinding_direction,
// :346-361), scatters c into a face array, overrides inflow boundary faces that carry a boundary
// concentration (:308-317), multiplies by the face velocity and applies DIV (:319-321).  Here one
// thread per cell gathers the same quantities: the cell the flux leaves is this cell when
// sigma v > 0 and the other cell of the face otherwise.  Algorithmic bytes per hex cell:
// c, p_prev, p, phi 32 B + cell faces / orientations 30 B + per face (v 8, area 8, bc 8, face_cell 8,
// face_to_flux 4) x 3 = 108 B + 8 B written  ~ 180 B.
#include "common.cuh"

__global__ void __launch_bounds__(256)
k_transport_step(mimpy_mesh_view m, const int* __restrict__ f2f, const int* __restrict__ face_cell,
                 const double* __restrict__ v, const double* __restrict__ c_old,
                 const double* __restrict__ bc_conc, const double* __restrict__ p_prev,
                 const double* __restrict__ p_cur, const double* __restrict__ phi, double ref_p,
                 double compressibility, double ref_rho, double dt, double* __restrict__ c_new) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m.nc) return;
  const int b = m.cell_ptr[c], e = m.cell_ptr[c + 1];
  double div_c = 0.;
  for (int q = b; q < e; ++q) {
    const int f = m.cell_faces[q];
    const int g = f2f[f];
    if (g < 0) continue;  // no-flow face
    const double s = (double)m.cell_sigma[q];
    const double vel = v[g];
    double cu = 0.;
    if (s * vel > 0.) {
      cu = c_old[c];
    } else if (s * vel < 0.) {
      const int c0 = face_cell[2 * (size_t)f], c1 = face_cell[2 * (size_t)f + 1];
      const int other = (c0 == c) ? c1 : c0;
      if (other >= 0) cu = c_old[other];
      else if (bc_conc) {
        const double bv = bc_conc[f];
        if (bv == bv) cu = bv;  // NaN marks "no boundary concentration": stays 0 like current_c_uw
      }
    }
    div_c += s * rec_area(m.face_geom, (size_t)f) * (cu * vel);  // DIV[c, g(f)] = sigma a_f  (mfd.py:232-238)
  }
  // transport.py:323-342 in the reference's operation order (no division by |E|: :344 is commented out)
  const double prev_density = ((p_prev[c] - ref_p) * compressibility + 1.) * ref_rho;
  const double cur_density = ((p_cur[c] - ref_p) * compressibility + 1.) * ref_rho;
  double conc = c_old[c] * (prev_density * phi[c]);
  conc = conc / dt;
  conc = conc - div_c;
  conc = conc * (dt / phi[c]);
  c_new[c] = conc / cur_density;
}

extern "C" int mimpy_b200_transport_step(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym,
                                         const double* velocity, const double* conc_old,
                                         const double* boundary_conc, const double* p_prev,
                                         const double* p_cur, const double* porosity, double ref_pressure,
                                         double compressibility, double ref_density, double dt,
                                         double* conc_new) {
  REQUIRE(ctx && mesh && sym && velocity && conc_old && p_prev && p_cur && porosity && conc_new, "NULL argument");
  REQUIRE(conc_new != conc_old, "the update reads neighbour cells: conc_new must not alias conc_old");
  if (mesh->nc <= 0) return MIMPY_OK;
  k_transport_step<<<(mesh->nc + 255) / 256, 256, 0, ctx->stream>>>(
      *mesh, sym->face_to_flux, sym->face_cell, velocity, conc_old, boundary_conc, p_prev, p_cur, porosity,
      ref_pressure, compressibility, ref_density, dt, conc_new);
  KERNEL_CHECK();
  return MIMPY_OK;
}
