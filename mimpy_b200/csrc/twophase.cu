// Two-phase (IMPES) transport kernels: Corey mobilities (twophase.py:268-344), upwind direction
// (twophase.py:928-958) and the explicit saturation update (twophase.py:960-1092).
#include <stdlib.h>

#include "common.cuh"

__device__ __forceinline__ double powd(double x, double e) {
  if (e == 2.) return x * x;
  if (e == 1.) return x;
  if (e == 3.) return x * x * x;
  return pow(x, e);
}

__device__ __forceinline__ void mobilities(const mimpy_fluid& fl, double sw, double p, double& mw,
                                           double& mo) {
  // sefromsw (twophase.py:317-324), krw/kro with clipping (275-289), mobility (326-344)
  double se = sw - fl.s_wr;
  se /= 1. - fl.s_wr - fl.s_or;
  se = se < 0. ? 0. : (se > 1. ? 1. : se);
  const double krw = fl.k_rw_o * powd(se, fl.n_w);
  const double kro = powd(1.0 - se, fl.n_o);
  mw = krw / fl.mu_w * fl.rho_w * (1. + fl.c_w * p);
  mo = kro / fl.mu_o * fl.rho_o * (1. + fl.c_o * p);
}

__global__ void k_mobility(int nc, mimpy_fluid fl, const double* __restrict__ sw,
                           const double* __restrict__ p, double* __restrict__ inv_tot,
                           double* __restrict__ wm, double* __restrict__ om) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  double mw, mo;
  mobilities(fl, sw[c], p[c], mw, mo);
  if (inv_tot) inv_tot[c] = 1. / (mw + mo);
  if (wm) wm[c] = mw;
  if (om) om[c] = mo;
}

__global__ void k_fill_i32(int n, int* a, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}

// (flux_index, cell) pairs with sigma*u_t[flux_index] > 0, flux_index = -1 on Neumann faces reads
// u_t[-1] = the LAST flux DOF and later overwrites its f_w: the last pair in (cell, local face)
// order wins (twophase.py:933-943, 968-971).
__global__ void k_upwind(int nc, int F, const int* __restrict__ cell_ptr,
                         const int* __restrict__ cell_faces, const int8_t* __restrict__ sigma,
                         const int* __restrict__ f2f, const double* __restrict__ u_t,
                         int* __restrict__ upwind, long long* __restrict__ last_key) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const int base = cell_ptr[c];
  const int n = cell_ptr[c + 1] - base;
  for (int j = 0; j < n; ++j) {
    int g = f2f[cell_faces[base + j]];
    if (g < 0) g = F - 1;
    const double dir = (double)sigma[base + j] * u_t[g];
    if (dir > 0) {
      if (g == F - 1)
        atomicMax(last_key, (long long)c * 64 + j);
      else
        upwind[g] = c;
    }
  }
}

__global__ void k_upwind_last(int F, const long long* __restrict__ last_key, int* __restrict__ upwind) {
  if (*last_key >= 0) upwind[F - 1] = (int)(*last_key >> 6);
}

// wm / om: per-cell mobilities evaluated by the caller (user relperm callables, twophase.py:292-300);
// NULL: Corey curves evaluated here
__global__ void k_frac_flow(int F, mimpy_fluid fl, const int* __restrict__ upwind,
                            const double* __restrict__ sw, const double* __restrict__ p,
                            const double* __restrict__ wm, const double* __restrict__ om,
                            double* __restrict__ f_w) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= F) return;
  const int c = upwind[g];
  if (c < 0) return;
  double mw, mo;
  if (wm) {
    mw = wm[c];
    mo = om[c];
  } else {
    mobilities(fl, sw[c], p[c], mw, mo);
  }
  f_w[g] = mw / (mw + mo);
}

__global__ void k_bnd_inflow(int n_bnd, const int* __restrict__ faces, const int8_t* __restrict__ sg,
                             const double* __restrict__ bnd_fw, const int* __restrict__ f2f,
                             const double* __restrict__ u_t, double* __restrict__ f_w) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_bnd) return;
  const int g = f2f[faces[k]];
  if (g < 0) return;
  if ((double)sg[k] * u_t[g] < 0) f_w[g] = bnd_fw[k];
}

__global__ void k_saturation(int nc, mimpy_fluid fl, double dt, const int* __restrict__ cell_ptr,
                             const int* __restrict__ cell_faces, const int8_t* __restrict__ sigma,
                             const double* __restrict__ face_geom, const int* __restrict__ f2f,
                             const double* __restrict__ u_t, const double* __restrict__ f_w,
                             const double* __restrict__ p, const double* __restrict__ p_prev,
                             const double* __restrict__ porosity, const double* __restrict__ vol,
                             double* __restrict__ s_w) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const int base = cell_ptr[c];
  const int n = cell_ptr[c + 1] - base;
  double div = 0.;
  for (int j = 0; j < n; ++j) {
    const int f = cell_faces[base + j];
    const int g = f2f[f];
    if (g < 0) continue;
    const double e = (double)sigma[base + j] * rec_area(face_geom, (size_t)f);
    div += e * (f_w[g] * u_t[g]);
  }
  // twophase.py:1007-1022
  double a = p_prev[c] * fl.c_w + 1.;
  a *= fl.rho_w;
  a *= s_w[c];
  a /= fl.rho_w;
  a /= 1. + fl.c_w * p[c];
  double b = 1. / porosity[c];
  b /= fl.rho_w;
  b /= (1. + fl.c_w * p[c]);
  b /= vol[c];
  b *= div;
  b *= -dt;
  s_w[c] = a + b;
}

// Fractional flow of every CELL (twophase.py:968-971 evaluates it per upwinded face from the face's upwind cell: the
// same arithmetic on the same operands, once per cell instead of once per face)
__global__ void __launch_bounds__(256)
k_cell_frac_flow(int nc, mimpy_fluid fl, const double* __restrict__ sw, const double* __restrict__ p,
                 const double* __restrict__ wm, const double* __restrict__ om, double* __restrict__ fwc, int F,
                 const int* __restrict__ upwind, double* __restrict__ f_w) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  double mw, mo;
  if (wm) {
    mw = wm[c];
    mo = om[c];
  } else {
    mobilities(fl, sw[c], p[c], mw, mo);
  }
  const double fw = mw / (mw + mo);
  fwc[c] = fw;
  // the last flux DOF takes f_w of its upwind cell before the boundary loop may override it (twophase.py:968-994)
  if (F > 0 && upwind[F - 1] == c) f_w[F - 1] = fw;
}

// Fused form of k_frac_flow + k_saturation: the fractional flow of a face is evaluated from its upwind cell
// on the fly (or, two-pass form, taken from the per-cell array fwc; s_new may then be s_old) (no pass over the faces, no f_w read for upwinded faces); the upwind cell records it in f_w,
// which stays the STATE the reference keeps between steps (a face that is in no upwind pair -- inflow
// boundary, zero flux -- uses the value it had, twophase.py:968-994).  Reads the saturation of the
// previous sub-step (s_old), writes s_new.  The last flux DOF (the u_t[-1] quirk) is prepared separately.
__global__ void __launch_bounds__(256)
k_saturation_fused(int nc, int F, mimpy_fluid fl, double dt, const int* __restrict__ cell_ptr,
                   const int* __restrict__ cell_faces, const int8_t* __restrict__ sigma,
                   const double* __restrict__ face_geom, const int* __restrict__ f2f,
                   const int* __restrict__ upwind, const double* __restrict__ u_t, const double* __restrict__ p,
                   const double* __restrict__ p_prev, const double* __restrict__ porosity,
                   const double* __restrict__ vol, const double* __restrict__ wm, const double* __restrict__ om,
                   const double* __restrict__ fwc, const double* s_old, double* s_new, double* __restrict__ f_w) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const int base = cell_ptr[c];
  const int n = cell_ptr[c + 1] - base;
  double div = 0.;
  for (int j = 0; j < n; ++j) {
    const int f = cell_faces[base + j];
    const int g = f2f[f];
    if (g < 0) continue;
    const int up = upwind[g];
    double fw;
    if (up < 0 || g == F - 1) {
      fw = f_w[g];
    } else {
      if (fwc) {  // two-pass form: the fractional flow of every cell was evaluated once (k_cell_frac_flow)
        fw = fwc[up];
      } else {
        double mw, mo;
        if (wm) {
          mw = wm[up];
          mo = om[up];
        } else {
          mobilities(fl, s_old[up], p[up], mw, mo);
        }
        fw = mw / (mw + mo);
      }
      if (up == c) f_w[g] = fw;
    }
    const double e = (double)sigma[base + j] * rec_area(face_geom, (size_t)f);
    div += e * (fw * u_t[g]);
  }
  // twophase.py:1007-1022
  double a = p_prev[c] * fl.c_w + 1.;
  a *= fl.rho_w;
  a *= s_old[c];
  a /= fl.rho_w;
  a /= 1. + fl.c_w * p[c];
  double b = 1. / porosity[c];
  b /= fl.rho_w;
  b /= (1. + fl.c_w * p[c]);
  b /= vol[c];
  b *= div;
  b *= -dt;
  s_new[c] = a + b;
}

// Default form: the update over the cells with the fractional flow of every cell evaluated once before (k_cell_frac_flow).
// Same arithmetic as k_saturation_fused; what changes is how the gathers are issued.  The face loop of the kernels above
// walks a dependent chain per face (cell_faces -> face_to_flux -> upwind -> f_w of the upwind cell), one face after the
// other: 4 x n_E exposed latencies per thread, 20 % issue utilisation, 45 % of the DRAM peak (ncu, 128^3).  Here the
// faces go through the chain B at a time, stage by stage, so B loads of a stage are in flight together; s_w is updated in
// place (only the own saturation is read).  flux_area: area of the face of every flux DOF (mimpy_b200_flux_areas), 8 B
// instead of the 32-B sector of the packed record; NULL = read the record.
// HEX: structured hex mesh with HexMesh's numbering (hexmesh.py:172-243): the face ids and orientations of a cell are
// closed-form, two levels of the chain (cell_ptr, cell_faces) and 34 B/cell disappear.
struct HexDims {
  int nx, ny, nz;
};
template <int B, int MINB, bool HEX>
__global__ void __launch_bounds__(256, MINB)
k_saturation_cells(int nc, int F, mimpy_fluid fl, double dt, HexDims hd, const int* __restrict__ cell_ptr,
                   const int* __restrict__ cell_faces, const int8_t* __restrict__ sigma,
                   const double* __restrict__ face_geom, const double* __restrict__ flux_area,
                   const int* __restrict__ f2f, const int* __restrict__ upwind, const double* __restrict__ u_t,
                   const double* __restrict__ p, const double* __restrict__ p_prev,
                   const double* __restrict__ porosity, const double* __restrict__ vol,
                   const double* __restrict__ fwc, double* __restrict__ s_w, double* f_w) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  static_assert(!HEX || B == 6, "a hex cell is one batch");
  const int base = HEX ? 0 : __ldg(cell_ptr + c);
  const int n = HEX ? 6 : __ldg(cell_ptr + c + 1) - base;
  const double pc = __ldg(p + c), pp = __ldg(p_prev + c), phi = __ldg(porosity + c), v = __ldg(vol + c);
  const double s_old = s_w[c];
  double div = 0.;
  for (int j0 = 0; j0 < n; j0 += B) {
    int f[B], g[B], up[B];
    double sg[B], a[B], u[B], fw[B];
    if (HEX) {  // cells list [z-, y-, x-, x+, y+, z+] with orientations [-1, -1, -1, 1, 1, 1]   hexmesh.py:329-342
      const int nx = hd.nx, ny = hd.ny, nz = hd.nz;
      const int ci = c % nx, cj = (c / nx) % ny, ck = c / (nx * ny);
      const int xrow = 3 * nx + 1, L = nx * ny + nx * (ny + 1) + (nx + 1) * ny;
      const int cbase = ck * L + cj * xrow + 3 * ci;
      f[0] = cbase;
      f[1] = cbase + 1;
      f[2] = cbase + 2;
      f[3 % B] = ci + 1 < nx ? cbase + 5 : ck * L + cj * xrow + 3 * nx;
      f[4 % B] = cj + 1 < ny ? cbase + xrow + 1 : ck * L + ny * xrow + ci;
      f[5 % B] = ck + 1 < nz ? cbase + L : nz * L + cj * nx + ci;
#pragma unroll
      for (int q = 0; q < B; ++q) sg[q] = q < 3 ? -1. : 1.;
    } else {
#pragma unroll
      for (int q = 0; q < B; ++q) {
        const bool in = j0 + q < n;
        f[q] = in ? __ldg(cell_faces + base + j0 + q) : -1;
        sg[q] = in ? (double)sigma[base + j0 + q] : 0.;
      }
    }
#pragma unroll
    for (int q = 0; q < B; ++q) g[q] = f[q] >= 0 ? __ldg(f2f + f[q]) : -1;
#pragma unroll
    for (int q = 0; q < B; ++q) {
      up[q] = -1;
      u[q] = a[q] = 0.;
      if (g[q] >= 0) {
        up[q] = __ldg(upwind + g[q]);
        u[q] = __ldg(u_t + g[q]);
        a[q] = flux_area ? __ldg(flux_area + g[q]) : rec_area(face_geom, (size_t)f[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < B; ++q) {
      fw[q] = 0.;
      if (g[q] >= 0) {
        const bool state = up[q] < 0 || g[q] == F - 1;  // no upwind pair / the u_t[-1] quirk: the value f_w holds
        fw[q] = state ? f_w[g[q]] : __ldg(fwc + up[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < B; ++q) {
      if (g[q] < 0) continue;
      if (up[q] == c && g[q] != F - 1) f_w[g[q]] = fw[q];  // the upwind cell records the state (twophase.py:968-971)
      div += (sg[q] * a[q]) * (fw[q] * u[q]);
    }
  }
  // twophase.py:1007-1022
  double x = pp * fl.c_w + 1.;
  x *= fl.rho_w;
  x *= s_old;
  x /= fl.rho_w;
  x /= 1. + fl.c_w * pc;
  double y = 1. / phi;
  y /= fl.rho_w;
  y /= (1. + fl.c_w * pc);
  y /= v;
  y *= div;
  y *= -dt;
  s_w[c] = x + y;
}

__global__ void k_flux_areas(int nf, const int* __restrict__ f2f, const double* __restrict__ face_geom,
                             double* __restrict__ out) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  const int g = f2f[f];
  if (g >= 0) out[g] = rec_area(face_geom, (size_t)f);
}

// the last flux DOF: f_w of its upwind cell, before the boundary loop may override it (twophase.py:968-994)
__global__ void k_frac_flow_last(int F, mimpy_fluid fl, const int* __restrict__ upwind, const double* __restrict__ sw,
                                 const double* __restrict__ p, const double* __restrict__ wm,
                                 const double* __restrict__ om, double* __restrict__ f_w) {
  const int c = upwind[F - 1];
  if (c < 0) return;
  double mw, mo;
  if (wm) {
    mw = wm[c];
    mo = om[c];
  } else {
    mobilities(fl, sw[c], p[c], mw, mo);
  }
  f_w[F - 1] = mw / (mw + mo);
}

__global__ void k_rate_wells(int nw, mimpy_fluid fl, double dt, const int* __restrict__ cells,
                             const double* __restrict__ rate_w, const double* __restrict__ p,
                             const double* __restrict__ porosity, const double* __restrict__ vol,
                             double* __restrict__ s_w) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nw) return;
  const int c = cells[w];
  // twophase.py:1024-1035
  double v = dt;
  v /= porosity[c];
  v /= 1. + fl.c_w * p[c];
  v /= fl.rho_w;
  v *= rate_w[w];
  v /= vol[c];
  atomicAdd(s_w + c, v);
}

extern "C" {

int mimpy_b200_twophase_mobility(mimpy_ctx* ctx, int32_t nc, const mimpy_fluid* fluid_host,
                                 const double* s_w, const double* p, double* inv_total_mobility,
                                 double* water_mob, double* oil_mob) {
  REQUIRE(ctx && fluid_host && s_w && p, "NULL argument");
  if (nc <= 0) return MIMPY_OK;
  k_mobility<<<(nc + 255) / 256, 256, 0, ctx->stream>>>(nc, *fluid_host, s_w, p, inv_total_mobility,
                                                        water_mob, oil_mob);
  KERNEL_CHECK();
  return MIMPY_OK;
}

int mimpy_b200_twophase_upwind(mimpy_ctx* ctx, const mimpy_mesh_view* mesh,
                               const mimpy_symbolic* sym, const double* u_t, int32_t* upwind_cell) {
  REQUIRE(ctx && mesh && sym && u_t && upwind_cell, "NULL argument");
  const int F = sym->flux_dof, nc = mesh->nc;
  if (F <= 0 || nc <= 0) return MIMPY_OK;
  cudaStream_t st = ctx->stream;
  long long* key = reinterpret_cast<long long*>(ctx->scalars + 32);
  CUDA_TRY(cudaMemsetAsync(key, 0xFF, sizeof(long long), st));  // -1
  k_fill_i32<<<(F + 255) / 256, 256, 0, st>>>(F, upwind_cell, -1);
  k_upwind<<<(nc + 255) / 256, 256, 0, st>>>(nc, F, mesh->cell_ptr, mesh->cell_faces,
                                             mesh->cell_sigma, sym->face_to_flux, u_t, upwind_cell, key);
  k_upwind_last<<<1, 1, 0, st>>>(F, key, upwind_cell);
  KERNEL_CHECK();
  return MIMPY_OK;
}

int mimpy_b200_flux_areas(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym, double* flux_area) {
  REQUIRE(ctx && mesh && sym && flux_area && mesh->face_geom && sym->face_to_flux, "NULL argument");
  if (mesh->nf <= 0) return MIMPY_OK;
  k_flux_areas<<<(mesh->nf + 255) / 256, 256, 0, ctx->stream>>>(mesh->nf, sym->face_to_flux, mesh->face_geom, flux_area);
  KERNEL_CHECK();
  return MIMPY_OK;
}

int mimpy_b200_twophase_saturation_step(mimpy_ctx* ctx, const mimpy_mesh_view* mesh,
                                        const mimpy_symbolic* sym, const mimpy_fluid* fluid_host,
                                        double dt, const int32_t* upwind_cell, const double* u_t,
                                        const double* p, const double* p_prev,
                                        const double* porosity, int32_t n_bnd,
                                        const int32_t* bnd_faces, const int8_t* bnd_sigma,
                                        const double* bnd_fw, int32_t n_rate_wells,
                                        const int32_t* well_cells, const double* well_rate_water,
                                        double* f_w, double* s_w, const double* water_mob,
                                        const double* oil_mob) {
  return mimpy_b200_twophase_saturation_step2(ctx, mesh, sym, fluid_host, dt, upwind_cell, u_t, p, p_prev, porosity, n_bnd,
                                              bnd_faces, bnd_sigma, bnd_fw, n_rate_wells, well_cells, well_rate_water, f_w,
                                              s_w, water_mob, oil_mob, nullptr);
}

int mimpy_b200_twophase_saturation_step2(mimpy_ctx* ctx, const mimpy_mesh_view* mesh,
                                         const mimpy_symbolic* sym, const mimpy_fluid* fluid_host,
                                         double dt, const int32_t* upwind_cell, const double* u_t,
                                         const double* p, const double* p_prev,
                                         const double* porosity, int32_t n_bnd,
                                         const int32_t* bnd_faces, const int8_t* bnd_sigma,
                                         const double* bnd_fw, int32_t n_rate_wells,
                                         const int32_t* well_cells, const double* well_rate_water,
                                         double* f_w, double* s_w, const double* water_mob,
                                         const double* oil_mob, const double* flux_area) {
  REQUIRE(ctx && mesh && sym && fluid_host && upwind_cell && u_t && p && p_prev && porosity && f_w && s_w,
          "NULL argument");
  const int F = sym->flux_dof, nc = mesh->nc;
  if (nc <= 0) return MIMPY_OK;
  cudaStream_t st = ctx->stream;
  const mimpy_fluid fl = *fluid_host;
  REQUIRE((water_mob == nullptr) == (oil_mob == nullptr), "water_mob and oil_mob must be given together");
  // Three forms of the same update (identical arithmetic per face):
  //  * default, two passes: the fractional flow of every cell once (k_cell_frac_flow, 24 B/cell), then one pass over the
  //    cells that takes f_w of a face from its upwind cell's entry and updates s_w in place -- no pass over the faces,
  //    one division per cell;
  //  * MIMPY_B200_SAT_PERFACE: the reference's formulation, a pass over the faces (k_frac_flow) and one over the cells
  //    (k_saturation), 378 B/cell: 227 us at 128^3;
  //  * MIMPY_B200_SAT_FUSED: single pass, mobilities and a division per (cell, upwinded face): 313 us, division-bound.
  // f_w stays the STATE the reference keeps between steps in all three.
  const int form = getenv("MIMPY_B200_SAT_PERFACE") ? 1 : getenv("MIMPY_B200_SAT_FUSED") ? 2 : 0;
  if (form == 1) {
    if (F > 0) k_frac_flow<<<(F + 255) / 256, 256, 0, st>>>(F, fl, upwind_cell, s_w, p, water_mob, oil_mob, f_w);
    if (n_bnd > 0) {
      REQUIRE(bnd_faces && bnd_sigma && bnd_fw, "boundary inflow arrays are NULL");
      k_bnd_inflow<<<(n_bnd + 255) / 256, 256, 0, st>>>(n_bnd, bnd_faces, bnd_sigma, bnd_fw,
                                                        sym->face_to_flux, u_t, f_w);
    }
    k_saturation<<<(nc + 255) / 256, 256, 0, st>>>(nc, fl, dt, mesh->cell_ptr, mesh->cell_faces,
                                                   mesh->cell_sigma, mesh->face_geom,
                                                   sym->face_to_flux, u_t, f_w, p, p_prev, porosity,
                                                   mesh->cell_volume, s_w);
  } else {
    void* scratch = nullptr;
    int rc = mimpy_scratch(ctx, sizeof(double) * (size_t)nc, &scratch);
    if (rc) return rc;
    double* buf = static_cast<double*>(scratch);  // form 0: per-cell fractional flow; form 2: the new saturation
    if (form == 0)
      k_cell_frac_flow<<<(nc + 255) / 256, 256, 0, st>>>(nc, fl, s_w, p, water_mob, oil_mob, buf, F, upwind_cell, f_w);
    else if (F > 0)
      k_frac_flow_last<<<1, 1, 0, st>>>(F, fl, upwind_cell, s_w, p, water_mob, oil_mob, f_w);
    if (n_bnd > 0) {
      REQUIRE(bnd_faces && bnd_sigma && bnd_fw, "boundary inflow arrays are NULL");
      k_bnd_inflow<<<(n_bnd + 255) / 256, 256, 0, st>>>(n_bnd, bnd_faces, bnd_sigma, bnd_fw,
                                                        sym->face_to_flux, u_t, f_w);
    }
    if (form == 0) {
      const HexDims hd = {mesh->hex_nx, mesh->hex_ny, mesh->hex_nz};
      const bool hex = hd.nx > 0 && hd.ny > 0 && hd.nz > 0 && mesh->uniform_faces == 6 &&
                       (long long)hd.nx * hd.ny * hd.nz == nc && !getenv("MIMPY_B200_SAT_NOHEX");
#define SAT_LAUNCH(HX)                                                                                                         \
  k_saturation_cells<6, 3, HX><<<(nc + 255) / 256, 256, 0, st>>>(nc, F, fl, dt, hd, mesh->cell_ptr, mesh->cell_faces,          \
                                                                 mesh->cell_sigma, mesh->face_geom, flux_area,                  \
                                                                 sym->face_to_flux, upwind_cell, u_t, p, p_prev, porosity,     \
                                                                 mesh->cell_volume, buf, s_w, f_w)
      if (hex) SAT_LAUNCH(true);
      else SAT_LAUNCH(false);
#undef SAT_LAUNCH
    } else {
      k_saturation_fused<<<(nc + 255) / 256, 256, 0, st>>>(nc, F, fl, dt, mesh->cell_ptr, mesh->cell_faces, mesh->cell_sigma,
                                                           mesh->face_geom, sym->face_to_flux, upwind_cell, u_t, p, p_prev,
                                                           porosity, mesh->cell_volume, water_mob, oil_mob, nullptr, s_w, buf,
                                                           f_w);
      CUDA_TRY(cudaMemcpyAsync(s_w, buf, sizeof(double) * (size_t)nc, cudaMemcpyDeviceToDevice, st));
    }
  }
  if (n_rate_wells > 0) {
    REQUIRE(well_cells && well_rate_water, "well arrays are NULL");
    k_rate_wells<<<(n_rate_wells + 255) / 256, 256, 0, st>>>(n_rate_wells, fl, dt, well_cells,
                                                             well_rate_water, p, porosity,
                                                             mesh->cell_volume, s_w);
  }
  KERNEL_CHECK();
  return MIMPY_OK;
}

}  // extern "C"

// Newton set-up of TwoPhase.update_pressure (twophase.py:764-816): the accumulation diagonal
//   C_c = (rho_w s c_w + rho_o c_o (1-s)) phi |E| / dt
// and the cell part of the residual  rhs[F+c] += f1 + f2 - f3  (mass of both phases at time n
// minus its linearisation), evaluated with the pressure of time level n.
__global__ void k_newton_cell_terms(int nc, int F, mimpy_fluid fl, double dt,
                                    const double* __restrict__ s_w, const double* __restrict__ p_n,
                                    const double* __restrict__ porosity, const double* __restrict__ vol,
                                    double* __restrict__ rhs, double* __restrict__ c_diag) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const double s = s_w[c], phi = porosity[c], v = vol[c], p = p_n[c];
  if (c_diag) {
    double cm = fl.rho_w * s;
    cm *= fl.c_w;
    cm += fl.rho_o * (fl.c_o * (1. - s));
    cm *= phi;
    cm *= v;
    cm /= dt;
    c_diag[c] = cm;
  }
  if (rhs) {
    double f1 = 1. * (fl.rho_w * s);
    f1 *= phi / dt;
    f1 *= v;
    double f2 = 1. * fl.rho_o;
    f2 *= 1. - s;
    f2 *= phi / dt;
    f2 *= v;
    double f3 = 0. + fl.rho_w * (1. + fl.c_w * p);
    f3 *= s;
    f3 += fl.rho_o * (1 + fl.c_o * p) * (1. - s);
    f3 *= phi / dt;
    f3 *= v;
    double r = rhs[F + c];
    r += f1;
    r += f2;
    r -= f3;
    rhs[F + c] = r;
  }
}

extern "C" int mimpy_b200_twophase_newton_terms(mimpy_ctx* ctx, int32_t nc, int32_t flux_dof,
                                                const mimpy_fluid* fluid_host, double dt,
                                                const double* s_w, const double* p_n,
                                                const double* porosity, const double* cell_volume,
                                                double* rhs, double* c_diag) {
  REQUIRE(ctx && fluid_host && s_w && p_n && porosity && cell_volume, "NULL argument");
  if (nc <= 0) return MIMPY_OK;
  k_newton_cell_terms<<<(nc + 255) / 256, 256, 0, ctx->stream>>>(nc, flux_dof, *fluid_host, dt, s_w, p_n,
                                                                 porosity, cell_volume, rhs, c_diag);
  KERNEL_CHECK();
  return MIMPY_OK;
}
