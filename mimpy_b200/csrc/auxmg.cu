// Auxiliary-space multigrid preconditioner for the hybridised face system (structured hex meshes).
//
// Jacobi-PCG on the face system A lambda = h needs O(n) iterations on an n^3 mesh (1600 at 256^3).
// The reference has no counterpart (it factorises with SuperLU, mfd.py:1349-1358; its dead
// Schur-complement CG, twophase.py:665-686, preconditions with D diag(M)^-1 D^T + C -- a
// cell-centred two-point-flux operator, which is exactly the auxiliary operator used here):
//
//     B = D^-1  +  Pi  V(A_c)  Pi^T                                  (additive, SPD)
//
//   A_c   cell operator with two-point transmissibilities: every cell E gives each of its flux
//         faces a half transmissibility tau_{E,f} = a_f^2 W_E[f,f] (W_E = the inverse local matrix of
//         hybrid.cu); interior face: t_f = tau_1 tau_2 / (tau_1 + tau_2) couples its two cells,
//         Dirichlet boundary face: tau on the diagonal; plus the C block (alpha |E| or c_diag).
//   Pi    cells -> faces, lambda_f = (tau_1 p_1 + tau_2 p_2) / (tau_1 + tau_2)  (flux continuity);
//         zero on boundary faces (their multiplier is prescribed).
//   V     one V(1,1) cycle of geometric multigrid on the 7-point operator: 2x2x2 aggregation with
//         piecewise-constant transfer (the Galerkin coarse operators stay 7-point: coarse couplings
//         are sums of the fine couplings through the coarse face), damped Jacobi smoothing, coarse
//         correction scaled by 1.5 (over-correction, the usual cure for unsmoothed aggregation; 2 is worse
//         here), dense inverse on the
//         coarsest (<= 64 cells) level.  Every operation is linear and symmetric, so B is a fixed SPD
//         operator and plain CG applies.
//
// Measured on B200 (log-normal K, rtol 1e-8, scripts/exp_precond.py): 200/600/1050/1600 Jacobi
// iterations at 16^3/64^3/128^3/256^3 become 35/40/45/45; with an exact solve of A_c instead of the
// V-cycle a CPU prototype needs 37-42, so the cycle is as good as the auxiliary space allows.
//
// Slab partitions (identical slabs): the hierarchy is shared between the ranks.  Aggregates never
// straddle a rank; every level exchanges one ghost layer of x (before the residual) and of the coarse
// correction (before the prolongation) with its z-neighbours (mimpy_layer_exchange, comm.cu); the
// couplings through the interfaces (tzi / bzi, Galerkin sums of the interface t_f) are applied to
// the ghost values.  Below AUX_AGG cells per rank (default 32768) a level is AGGLOMERATED: its right-hand
// side is all-gathered once (one peer-memory kernel) and every rank runs the rest of the V-cycle
// redundantly on the GLOBAL coarse grid (nx x ny x world*nz), down to a dense inverse of <= 64 cells -- no
// ghost exchange and no waiting on neighbours on those levels, whose kernels are pure launch latency.
// The cycle is the single-GPU cycle up to where the smoothers of the agglomerated levels sit (same
// iteration counts on 1, 2 and 8 GPUs; a rank-local block-Jacobi version needed 85 / 145 instead of 45).
// tau sums and the prolongated correction of interface rows are completed with the same exchange-add as
// the SpMV.  Uneven slabs: rank-local cycle.
#include "reduce.cuh"

#define AUX_MAX_LEVELS 24
#define AUX_COARSEST 64
#define AUX_AGG_DEFAULT 32768 /* a level with at most this many cells per rank is agglomerated (8 GPUs, 256^3: 0.805 / 0.756 / 0.721 ms per iteration with 64 / 4096 / 32768) */
#define AUX_T 256

template <typename T>
struct AuxLevelT {
  int nx, ny, nz, n;
  T *tx, *ty, *tz;       // coupling of cell c with its +x / +y / +z neighbour (0 on the boundary)
  T *dd;                 // extra diagonal: Dirichlet faces, interface faces, C block
  T *dg;                 // full diagonal
  T *b, *x, *x2;         // right-hand side, pre-smoothed iterate, result of the level; x and x2 carry two
                         // ghost layers behind the n cells: [n, n+nx*ny) from the rank below, then from above
  T *tzi, *bzi;          // [nx*ny] coupling of the top / bottom layer cells with the rank above / below
  int has_lo, has_hi;    // 1: the ghost layer is live (slab partition with the levels shared across ranks)
};
typedef AuxLevelT<double> AuxLevel;
// Single rank: the V-cycle runs on FP32 mirrors of the levels (coefficients and vectors; arithmetic in FP64
// registers).  It is a preconditioner -- the Krylov recurrences stay FP64 -- and it is pure bandwidth: 64 instead of
// 136 bytes per cell and cycle at level 0.  Iteration counts are unchanged (256^3: 95 / 95, C2: 1080 / 1110, C4: 35 /
// 35 per step).  Coefficients are scaled by a power of two s_A ~ 1 / tau, the restricted residual by s_b ~ 1 / |r|,
// so that nothing depends on the units of K or of the right-hand side; the cycle is linear, the prolongation undoes
// both.  Slab partitions keep FP64 levels (their ghost-layer exchanges are typed).
typedef AuxLevelT<float> AuxLevel32;

struct mimpy_auxmg {
  int nlev;
  AuxLevel lev[AUX_MAX_LEVELS];
  int nc, nf, flux_dof;
  double* wc;    // [nc*6]  set-up scratch: half transmissibility of cell c at its local face i
  double* fw;    // [nf]    transfer weight of the FIRST cell of face f (face_cell[2f]); the second cell
                 //         of an interior face has 1 - fw (the two weights of Pi sum to one)
  double* tsum;  // [flux_dof] sum of the half transmissibilities of a face
  double* cinv;  // dense inverse of the coarsest operator
  double* slab;  // the one stream-ordered allocation all arrays above are carved from
  int global;    // 1: slab partition with the hierarchy shared between the ranks
  int nglob;     // levels of the agglomerated (global, redundantly computed) part of the hierarchy
  AuxLevel glev[AUX_MAX_LEVELS];  // glev[0] = the last distributed level, gathered: nx x ny x (world * nz)
  double* gpack; // [world * 6 * n_last] set-up gather of the last distributed level's stencil
  int use32;     // 1: FP32 mirrors of the levels are live (single rank)
  AuxLevel32 lev32[AUX_MAX_LEVELS];
  double* sdev;  // device: [0] = s_A (power of two ~ 1 / tau of the first cells)
  float* fw32;   // FP32 copy of fw for the transfers of the FP32 cycle
  float* dinv32; // FP32 copy of the Jacobi diagonal (fused update of p), made per solve
  // cell-block smoother (mimpy_b200_auxmg_cellblocks): instead of D^-1 the face-space term of B is the weighted
  // additive Schwarz sum over the cells, sum_E P_E^T S_E (A_EE)^-1 S_E P_E (A_EE: the block of the face system on
  // the faces of cell E; S_E = diag(sqrt(1/2)) on faces with two cells, 1 on the others)
  float* binv;   // [21 * cpad] upper triangles of s_A^-1 x the weighted block inverses, entry-major (SoA)
  void* ycell;   // [6 * cpad] float: the block solutions of the current residual
  size_t cpad;
  int blocks;    // 0: not allocated, 1: allocated, 2: set up for the current matrix
  double omega, scale;
  double wj;     // weight of the Jacobi term D^-1 against the auxiliary-space term
  const int* cell_faces;
  const int* face_to_flux;
  const int* face_cell;
  const uint8_t* face_flag;
};

// ---------------------------------------------------------------------------------------------
// set-up
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AUX_T)
k_aux_tau(int nc, const int* __restrict__ cell_faces, const int* __restrict__ f2f,
          const int* __restrict__ m_ptr, const double* __restrict__ face_geom,
          const double* __restrict__ w_e, double* __restrict__ wc, double* __restrict__ tsum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const int boff = m_ptr[c];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int f = cell_faces[c * 6 + i];
    const int g = f2f[f];
    double tau = 0.;
    if (g >= 0) {
      const double a = rec_area(face_geom, (size_t)f);
      tau = a * a * w_e[boff + i * 7];
      atomicAdd(tsum + g, tau);  // <= 2 addends on a zeroed slot: order-free
    }
    wc[(size_t)c * 6 + i] = tau;
  }
}

// the same with the half transmissibilities already at hand (written by the set-up kernel)
__global__ void __launch_bounds__(AUX_T)
k_aux_tau_given(int nc, const int* __restrict__ cell_faces, const int* __restrict__ f2f,
                const double* __restrict__ tau_in, double* __restrict__ wc, double* __restrict__ tsum) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= (long long)nc * 6) return;
  const double tau = tau_in[e];
  wc[e] = tau;
  if (tau != 0.) {
    const int g = f2f[cell_faces[e]];
    if (g >= 0) atomicAdd(tsum + g, tau);
  }
}

__global__ void __launch_bounds__(AUX_T)
k_aux_level0(int nc, const int* __restrict__ cell_faces, const int* __restrict__ f2f,
             const int* __restrict__ face_cell, const uint8_t* __restrict__ face_flag,
             const double* __restrict__ tsum, const double* __restrict__ cell_volume, double alpha,
             const double* __restrict__ c_diag, double* __restrict__ wc, double* __restrict__ fw,
             float* __restrict__ fw32, AuxLevel L) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  double dd = c_diag ? c_diag[c] : alpha * cell_volume[c];
  double tp[3] = {0., 0., 0.};
  double t_up = 0., t_dn = 0.;  // couplings through the top / bottom interface of a slab partition
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int f = cell_faces[c * 6 + i];
    const int g = f2f[f];
    const double tau = wc[(size_t)c * 6 + i];
    double w = 0.;
    if (g >= 0 && tau > 0.) {
      const bool iface = face_flag && (face_flag[f] & 3);
      const bool two = face_cell[2 * (size_t)f + 1] >= 0;
      if (!two && !iface) {
        dd += tau;  // boundary face with a prescribed multiplier
      } else {
        const double S = tsum[g];
        const double t = tau * (S - tau) / S;
        w = tau / S;
        if (iface) {
          dd += t;  // neighbour on another rank: block-Jacobi (the shared coarsest level couples them)
          if (face_flag[f] & 1) t_up = t;
          else t_dn = t;
        } else if (i >= 3) {
          tp[i - 3] = t;  // faces x+, y+, z+ (cell face order z-,y-,x-,x+,y+,z+)
        }
      }
    }
    if (face_cell[2 * (size_t)f] == c) {
      fw[f] = w;
      if (fw32) fw32[f] = (float)w;
    }
  }
  L.tx[c] = tp[0];
  L.ty[c] = tp[1];
  L.tz[c] = tp[2];
  L.dd[c] = dd;
  const int layer = L.nx * L.ny;
  if (c >= L.n - layer) L.tzi[c - (L.n - layer)] = t_up;
  if (c < layer) L.bzi[c] = t_dn;
}

// full diagonal; with s_a: the FP32 mirror of the level's operator (scaled by s_A) is written on the way
__global__ void __launch_bounds__(AUX_T) k_aux_diag(AuxLevel L, AuxLevel32 M, const double* __restrict__ s_a) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= L.n) return;
  const int i = c % L.nx, j = (c / L.nx) % L.ny, k = c / (L.nx * L.ny);
  const double tx = L.tx[c], ty = L.ty[c], tz = L.tz[c];
  double d = L.dd[c] + tx + ty + tz;
  if (i > 0) d += L.tx[c - 1];
  if (j > 0) d += L.ty[c - L.nx];
  if (k > 0) d += L.tz[c - L.nx * L.ny];
  d = d > 0. ? d : 1.;
  L.dg[c] = d;
  if (s_a) {
    const double s = s_a[0];
    M.tx[c] = (float)(s * tx);
    M.ty[c] = (float)(s * ty);
    M.tz[c] = (float)(s * tz);
    M.dg[c] = (float)(s * d);
  }
}

// Galerkin coarse operator for piecewise-constant aggregation of 2x2x2 cells
__global__ void __launch_bounds__(AUX_T) k_aux_coarsen(AuxLevel F, AuxLevel C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C.n) return;
  const int I = c % C.nx, J = (c / C.nx) % C.ny, K = c / (C.nx * C.ny);
  const int i0 = 2 * I, j0 = 2 * J, k0 = 2 * K;
  const int i1 = min(i0 + 1, F.nx - 1), j1 = min(j0 + 1, F.ny - 1), k1 = min(k0 + 1, F.nz - 1);
  double tx = 0., ty = 0., tz = 0., dd = 0.;
  for (int k = k0; k <= k1; ++k)
    for (int j = j0; j <= j1; ++j)
      for (int i = i0; i <= i1; ++i) {
        const int f = i + F.nx * (j + F.ny * k);
        dd += F.dd[f];
        if (i == i1) tx += F.tx[f];  // the fine faces through the coarse +x face (0 on the boundary)
        if (j == j1) ty += F.ty[f];
        if (k == k1) tz += F.tz[f];
      }
  C.tx[c] = tx;
  C.ty[c] = ty;
  C.tz[c] = tz;
  C.dd[c] = dd;
  if (K == C.nz - 1) {
    double tu = 0.;
    for (int j = j0; j <= j1; ++j)
      for (int i = i0; i <= i1; ++i) tu += F.tzi[i + F.nx * j];
    C.tzi[I + C.nx * J] = tu;
  }
  if (K == 0) {
    double td = 0.;
    for (int j = j0; j <= j1; ++j)
      for (int i = i0; i <= i1; ++i) td += F.bzi[i + F.nx * j];
    C.bzi[I + C.nx * J] = td;
  }
}

// ---- agglomerated levels of a slab partition ---------------------------------------------------
// own stencil of the last distributed level into slot `rank` of the (zeroed) gather buffer; an all-reduce
// completes the gather.  Slot: tx, ty, tz, dd [n each], tzi, bzi [nx*ny each]
__global__ void __launch_bounds__(AUX_T) k_aux_global_pack(AuxLevel L, int rank, double* __restrict__ gpack) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= L.n) return;
  double* slot = gpack + (size_t)rank * 6 * L.n;
  slot[c] = L.tx[c];
  slot[L.n + c] = L.ty[c];
  slot[2 * L.n + c] = L.tz[c];
  slot[3 * L.n + c] = L.dd[c];
  if (c < L.nx * L.ny) {
    slot[4 * L.n + c] = L.tzi[c];
    slot[5 * L.n + c] = L.bzi[c];
  }
}

// global level 0: rank r's cells are the z-layers [r*nz, (r+1)*nz).  The coupling through a rank interface
// becomes an ordinary tz; the distributed levels keep it in their extra diagonal dd, so it is taken out.
__global__ void __launch_bounds__(AUX_T)
k_aux_global_unpack(int n, int nxy, int world, const double* __restrict__ gpack, AuxLevel G) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G.n) return;
  const int r = g / n, c = g - r * n;
  const double* slot = gpack + (size_t)r * 6 * n;
  double tz = slot[2 * n + c], dd = slot[3 * n + c];
  if (c >= n - nxy) {  // top layer of rank r
    const double tu = slot[4 * n + (c - (n - nxy))];
    dd -= tu;
    if (r + 1 < world) tz = tu;
  }
  if (c < nxy) dd -= slot[5 * n + c];  // bottom layer: coupling with the rank below (its tz carries it)
  G.tx[g] = slot[c];
  G.ty[g] = slot[n + c];
  G.tz[g] = tz;
  G.dd[g] = dd;
  if (g < G.nx * G.ny) {
    G.tzi[g] = 0.;
    G.bzi[g] = 0.;
  }
}

// result of the global level 0 back into the last distributed level: own cells, then the two ghost layers
__global__ void __launch_bounds__(AUX_T)
k_aux_global_extract(int n, int nxy, int rank, int world, const double* __restrict__ gx, double* __restrict__ x2) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n + 2 * nxy) return;
  long long g;
  if (t < n) g = (long long)rank * n + t;
  else if (t < n + nxy) g = rank > 0 ? (long long)rank * n - nxy + (t - n) : -1;
  else g = rank + 1 < world ? (long long)(rank + 1) * n + (t - n - nxy) : -1;
  if (g >= 0) x2[t] = gx[g];
}

// dense inverse of the coarsest operator (n <= AUX_COARSEST), Gauss-Jordan without pivoting (SPD)
__global__ void __launch_bounds__(AUX_T) k_aux_dense_inverse(AuxLevel L, double* __restrict__ cinv) {
  __shared__ double A[AUX_COARSEST][AUX_COARSEST + 1];
  __shared__ double piv_row[AUX_COARSEST], piv_col[AUX_COARSEST];
  const int n = L.n, t = threadIdx.x;
  for (int e = t; e < n * n; e += AUX_T) A[e / n][e % n] = 0.;
  __syncthreads();
  double dmax = 0.;
  for (int c = 0; c < n; ++c) dmax = fmax(dmax, L.dg[c]);
  for (int c = t; c < n; c += AUX_T) {
    const int i = c % L.nx, j = (c / L.nx) % L.ny, k = c / (L.nx * L.ny);
    A[c][c] = L.dg[c] + 1e-13 * dmax;  // keeps a pure-Neumann (singular) operator invertible
    if (i + 1 < L.nx) { A[c][c + 1] = -L.tx[c]; A[c + 1][c] = -L.tx[c]; }
    if (j + 1 < L.ny) { A[c][c + L.nx] = -L.ty[c]; A[c + L.nx][c] = -L.ty[c]; }
    if (k + 1 < L.nz) { A[c][c + L.nx * L.ny] = -L.tz[c]; A[c + L.nx * L.ny][c] = -L.tz[c]; }
  }
  __syncthreads();
  for (int p = 0; p < n; ++p) {  // in-place inversion
    const double piv = 1.0 / A[p][p];
    for (int e = t; e < n; e += AUX_T) {
      piv_row[e] = A[p][e] * piv;
      piv_col[e] = A[e][p];
    }
    __syncthreads();
    for (int e = t; e < n * n; e += AUX_T) {
      const int r = e / n, c = e % n;
      double v;
      if (r == p) v = (c == p) ? piv : piv_row[c];
      else if (c == p) v = -piv_col[r] * piv;
      else v = A[r][c] - piv_col[r] * piv_row[c];
      A[r][c] = v;
    }
    __syncthreads();
  }
  for (int e = t; e < n * n; e += AUX_T) cinv[e] = A[e / n][e % n];
}

// ---------------------------------------------------------------------------------------------
// V-cycle
// ---------------------------------------------------------------------------------------------
// The cycle computes in the storage type of the levels: FP32 arithmetic on the FP32 mirrors (conversions to FP64 run at
// a sixteenth of the FP32 rate and had made the level-0 kernels compute-bound: 0.31 ms for 0.5 GB), FP64 on FP64 levels.
template <typename T>
__device__ __forceinline__ T aux_rcp(T d) { return (T)1 / d; }
template <>
__device__ __forceinline__ float aux_rcp<float>(float d) { return __frcp_rn(d); }

template <typename T>
__global__ void __launch_bounds__(AUX_T) k_aux_presmooth(AuxLevelT<T> L, double omega) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < L.n) L.x[c] = (T)omega * L.b[c] * aux_rcp<T>(L.dg[c]);
}

template <typename T>
__device__ __forceinline__ T aux_apply(const AuxLevelT<T>& L, const T* __restrict__ x, int c,
                                       int i, int j, int k, T xc) {
  T s = L.dg[c] * xc;
  const int sx = 1, sy = L.nx, sz = L.nx * L.ny;
  if (i + 1 < L.nx) s -= L.tx[c] * x[c + sx];
  if (i > 0) s -= L.tx[c - sx] * x[c - sx];
  if (j + 1 < L.ny) s -= L.ty[c] * x[c + sy];
  if (j > 0) s -= L.ty[c - sy] * x[c - sy];
  if (k + 1 < L.nz) s -= L.tz[c] * x[c + sz];
  else if (L.has_hi) s -= L.tzi[i + sy * j] * x[L.n + sz + i + sy * j];  // cell of the rank above
  if (k > 0) s -= L.tz[c - sz] * x[c - sz];
  else if (L.has_lo) s -= L.bzi[i + sy * j] * x[L.n + i + sy * j];       // cell of the rank below
  return s;
}

// coarse right-hand side = sum over the aggregate of the fine residual b - A x
template <typename T>
__global__ void __launch_bounds__(AUX_T) k_aux_restrict(AuxLevelT<T> F, AuxLevelT<T> C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C.n) return;
  const int I = c % C.nx, J = (c / C.nx) % C.ny, K = c / (C.nx * C.ny);
  const int i0 = 2 * I, j0 = 2 * J, k0 = 2 * K;
  const int i1 = min(i0 + 1, F.nx - 1), j1 = min(j0 + 1, F.ny - 1), k1 = min(k0 + 1, F.nz - 1);
  T s = 0;
  for (int k = k0; k <= k1; ++k)
    for (int j = j0; j <= j1; ++j)
      for (int i = i0; i <= i1; ++i) {
        const int f = i + F.nx * (j + F.ny * k);
        s += F.b[f] - aux_apply(F, F.x, f, i, j, k, F.x[f]);
      }
  C.b[c] = s;
}

// x2 = xp + omega D^-1 (b - A xp),  xp = x + scale * P e_coarse   (prolongation fused into the smoother).
// DOT (level 0 of the fused update of p): b . x2 summed on the way (in FP64) -- the cell-space half of r . z.
template <typename T, bool DOT>
__global__ void __launch_bounds__(AUX_T)
k_aux_prolong_smooth(AuxLevelT<T> F, AuxLevelT<T> C, double omega_, const double* __restrict__ scale_dev, double* partials,
                     unsigned int* counter, double* dst) {
  __shared__ double sh[32];
  const T* __restrict__ e = C.x2;
  const T* __restrict__ x = F.x;
  const T omega = (T)omega_, scale = (T)scale_dev[0];
  const int cs_y = C.nx, cs_z = C.nx * C.ny;
  const int sx = 1, sy = F.nx, sz = F.nx * F.ny;
  double acc = 0.;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < F.n; c += gridDim.x * blockDim.x) {
    const int i = c % F.nx, j = (c / F.nx) % F.ny, k = c / (F.nx * F.ny);
    // the coarse cell of a neighbour is the own one, or the next one across the parity boundary
    const int ce = (i >> 1) + cs_y * (j >> 1) + cs_z * (k >> 1);
    const T xc = x[c] + scale * e[ce];
    T s = F.dg[c] * xc;
    if (i + 1 < F.nx) s -= F.tx[c] * (x[c + sx] + scale * e[ce + (i & 1)]);
    if (i > 0) s -= F.tx[c - sx] * (x[c - sx] + scale * e[ce - ((i & 1) ^ 1)]);
    if (j + 1 < F.ny) s -= F.ty[c] * (x[c + sy] + scale * e[ce + (j & 1) * cs_y]);
    if (j > 0) s -= F.ty[c - sy] * (x[c - sy] + scale * e[ce - ((j & 1) ^ 1) * cs_y]);
    const int ij = i + sy * j, cij = (i >> 1) + cs_y * (j >> 1);
    if (k + 1 < F.nz) s -= F.tz[c] * (x[c + sz] + scale * e[ce + (k & 1) * cs_z]);
    else if (F.has_hi) s -= F.tzi[ij] * (x[F.n + sz + ij] + scale * e[C.n + C.nx * C.ny + cij]);
    if (k > 0) s -= F.tz[c - sz] * (x[c - sz] + scale * e[ce - ((k & 1) ^ 1) * cs_z]);
    else if (F.has_lo) s -= F.bzi[ij] * (x[F.n + ij] + scale * e[C.n + cij]);
    const T bc = F.b[c];
    const T out = xc + omega * (bc - s) * aux_rcp<T>(F.dg[c]);
    F.x2[c] = out;
    if (DOT) acc += (double)bc * (double)out;  // (the stored, rounded value: what the prolongation will read)
  }
  if (DOT) {
    const double v[1] = {acc};
    finish_reduction<1>(v, partials, counter, dst, sh);
  }
}

// x2 = inv_scale * cinv b  (cinv: FP64 inverse of the unscaled coarsest operator; inv_scale = 1 / s_A for the FP32 levels)
template <typename T>
__global__ void __launch_bounds__(AUX_T)
k_aux_coarse_solve(AuxLevelT<T> L, const double* __restrict__ cinv, const double* __restrict__ s_a) {
  const int r = threadIdx.x;
  if (r >= L.n) return;
  double s = 0.;
  for (int c = 0; c < L.n; ++c) s += cinv[r * L.n + c] * (double)L.b[c];
  L.x2[r] = (T)(s_a ? s / s_a[0] : s);
}

// Over-correction of the piecewise-constant coarse space, chosen from the anisotropy of the cell operator: sums of the
// couplings of level 0 per direction (deterministic two-stage reduction; all-reduced over the ranks of a partition), then
//   scale = 1.8 - 0.5 log(ratio) / log(30), clamped to [1.3, 1.8],  ratio = largest / smallest directional sum.
// Measured with the cell-block smoother (iterations to a saddle residual of 1e-8): 256^3 with statistically isotropic
// couplings 60 / 50 / 45 at 1.5 / 1.65 / 1.8; C2 (flat cells, k_v = 0.1 k_h: ratio ~60) 245 / 265 / 310 / 350 at 1.3 / 1.5 /
// 1.65 / 1.8.  A prototype over six coefficient fields shows the same trend (isotropic: the larger the better up to 2,
// ratio >= 60: 1.3).
__global__ void __launch_bounds__(RED_THREADS)
k_aux_aniso_sums(AuxLevel L, double* partials, unsigned int* counter, double* dst) {
  __shared__ double sh[32];
  double v[3] = {0., 0., 0.};
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < L.n; c += gridDim.x * blockDim.x) {
    v[0] += L.tx[c];
    v[1] += L.ty[c];
    v[2] += L.tz[c];
  }
  finish_reduction<3>(v, partials, counter, dst, sh);
}
__global__ void k_aux_pick_scale(const double* __restrict__ sums, double fixed, double* __restrict__ scale_out) {
  if (fixed > 0.) {
    scale_out[0] = fixed;
    return;
  }
  double hi = 0., lo = 1e300;
  for (int d = 0; d < 3; ++d) {
    if (sums[d] > 0.) {
      hi = fmax(hi, sums[d]);
      lo = fmin(lo, sums[d]);
    }
  }
  double sc = 1.8;
  if (hi > 0. && lo < 1e300 && hi > lo) sc = 1.8 - 0.5 * log(hi / lo) / log(30.);
  sc = fmin(1.8, fmax(1.3, sc));
  scale_out[0] = floor(sc * 64. + 0.5) / 64.;  // (a few bits only: nothing hangs on the last digits of the sums)
}

// s_A = 2^-e with 2^e ~ the largest of the first half transmissibilities (any representative magnitude will do)
__global__ void k_aux_scale(const double* __restrict__ tau, int n, double* __restrict__ s_a) {
  double m = 0.;
  for (int q = threadIdx.x; q < n; q += 32) m = fmax(m, fabs(tau[q]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (threadIdx.x == 0) s_a[0] = (m > 0. && m < 1e300) ? scalbn(1.0, -ilogb(m)) : 1.0;
}

// ---------------------------------------------------------------------------------------------
// transfers between the face system and the cells
// ---------------------------------------------------------------------------------------------
// HexMesh numbering in closed form (hexmesh.py:161-234): faces of cell (i, j, k)
struct AuxHex {
  int nx, ny, nz, L, xrow;
  __device__ __forceinline__ int cbase(int i, int j, int k) const { return k * L + j * xrow + 3 * i; }  // z-, y-, x- = +0, +1, +2
  __device__ __forceinline__ int xp(int i, int j, int k) const { return i + 1 < nx ? cbase(i, j, k) + 5 : k * L + j * xrow + 3 * nx; }
  __device__ __forceinline__ int yp(int i, int j, int k) const { return j + 1 < ny ? cbase(i, j, k) + xrow + 1 : k * L + ny * xrow + i; }
  __device__ __forceinline__ int zp(int i, int j, int k) const { return k + 1 < nz ? cbase(i, j, k) + L : nz * L + j * nx + i; }
};

// Level 0 of a single-rank structured hex mesh from the half transmissibilities the set-up kernel wrote
// (tau[c*6 + q]): the partner of face q of cell c is local face 5 - q of the neighbouring cell, so neither the
// per-face sums (tsum: 100 M atomics at 256^3), nor a copy of tau, nor an index array is needed.  Same
// arithmetic as k_aux_tau_given + k_aux_level0 (the sum of two addends does not depend on their order).
__global__ void __launch_bounds__(AUX_T)
k_aux_level0_hex(AuxHex hx, const int* __restrict__ f2f, const double* __restrict__ tau,
                 const double* __restrict__ cell_volume, double alpha, const double* __restrict__ c_diag,
                 double* __restrict__ fw, float* __restrict__ fw32, AuxLevel L) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= L.n) return;
  const int sy = hx.nx, sz = hx.nx * hx.ny;
  const int i = c % hx.nx, j = (c / hx.nx) % hx.ny, k = c / sz;
  const int cb = hx.cbase(i, j, k);
  const int f[6] = {cb, cb + 1, cb + 2, hx.xp(i, j, k), hx.yp(i, j, k), hx.zp(i, j, k)};
  const bool partner[6] = {k > 0, j > 0, i > 0, i + 1 < hx.nx, j + 1 < hx.ny, k + 1 < hx.nz};
  const int other[6] = {c - sz, c - sy, c - 1, c + 1, c + sy, c + sz};
  double own[6];
  {
    const double2* t2 = reinterpret_cast<const double2*>(tau + (size_t)c * 6);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const double2 v = t2[q];
      own[2 * q] = v.x;
      own[2 * q + 1] = v.y;
    }
  }
  double dd = c_diag ? c_diag[c] : alpha * cell_volume[c];
  double tp[3] = {0., 0., 0.};
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    const int g = f2f[f[q]];
    double w = 0.;
    if (g >= 0 && own[q] > 0.) {
      if (!partner[q]) {
        dd += own[q];  // boundary face with a prescribed multiplier
      } else {
        const double S = own[q] + tau[(size_t)other[q] * 6 + (5 - q)];
        w = own[q] / S;
        if (q >= 3) tp[q - 3] = own[q] * (S - own[q]) / S;
      }
    }
    if (q >= 3 || !partner[q]) {  // the first (or only) cell of the face
      if (fw32) fw32[f[q]] = (float)w;
      else fw[f[q]] = w;
    }
  }
  L.tx[c] = tp[0];
  L.ty[c] = tp[1];
  L.tz[c] = tp[2];
  L.dd[c] = dd;
  if (c >= L.n - sz) L.tzi[c - (L.n - sz)] = 0.;
  if (c < sz) L.bzi[c] = 0.;
}

// b . x of a level of at most AUX_COARSEST cells (a mesh whose only level is the coarsest one)
__global__ void __launch_bounds__(AUX_T) k_aux_dot32(int n, const float* __restrict__ b, const float* __restrict__ x, double* dst) {
  __shared__ double sh[32];
  double a = 0.;
  for (int c = threadIdx.x; c < n; c += blockDim.x) a += (double)b[c] * (double)x[c];
  a = block_sum(a, sh);
  if (threadIdx.x == 0) dst[0] = a;
}

// power of two ~ 1 / rms(r), from the device scalar ||r||^2 of the PCG (FP32 levels: the restricted residual is O(1))
__device__ __forceinline__ double residual_scale(const double* __restrict__ rr, double n) {
  const double m = sqrt(rr[0] / n);
  return (m > 1e-300 && m < 1e300) ? scalbn(1.0, -ilogb(m)) : 1.0;
}

// b = Pi^T r on the cells.  A cell is the SECOND cell of its z-/y-/x- face when that neighbour exists
// (weight 1 - fw), the first cell of every other face (weight fw); no index array is read.
// BLOCKS: also y_E = S_E (A_EE)^-1 S_E r_E, the cell's term of the block smoother (the six residuals are in registers
// anyway); stored per (local face, cell), scaled like b, and summed over the two cells of a face by k_aux_face_prolong.
template <typename T, bool BLOCKS>
__global__ void __launch_bounds__(AUX_T)
k_aux_face_restrict(AuxHex hx, const int* __restrict__ f2f, const T* __restrict__ fw,
                    const double* __restrict__ r, T* __restrict__ b, const double* __restrict__ rr, double nrows,
                    const float* __restrict__ binv, float* __restrict__ y, size_t cpad, const double* __restrict__ ry) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= hx.nx * hx.ny * hx.nz) return;
  const int i = c % hx.nx, j = (c / hx.nx) % hx.ny, k = c / (hx.nx * hx.ny);
  const int cb = hx.cbase(i, j, k);
  const int f[6] = {cb, cb + 1, cb + 2, hx.xp(i, j, k), hx.yp(i, j, k), hx.zp(i, j, k)};
  const bool second[6] = {k > 0, j > 0, i > 0, false, false, false};
  const double sc = rr ? residual_scale(rr, nrows) : 1.;
  // the block solves run in FP32 whatever the type of the levels (scaled by a power of two ~ 1 / rms(r) of this rank)
  const double scy = (BLOCKS && ry) ? residual_scale(ry, nrows) : 1.;
  double s = 0.;
  float rq[6];
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    rq[q] = 0;
    const int g = f2f[f[q]];
    if (g < 0) continue;
    const double w1 = (double)fw[f[q]];
    const double w = second[q] ? (w1 > 0. ? 1. - w1 : 0.) : w1;
    if (BLOCKS) {
      const double rg = r[g];
      rq[q] = (float)(rg * scy);
      s += w * rg;
    } else if (w != 0.) {
      s += w * r[g];
    }
  }
  b[c] = (T)(s * sc);
  if (BLOCKS) {
    float m[21];
#pragma unroll
    for (int e = 0; e < 21; ++e) m[e] = binv[e * cpad + c];
    float yq[6] = {0, 0, 0, 0, 0, 0};
    int e = 0;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
#pragma unroll
      for (int p2 = q; p2 < 6; ++p2, ++e) {
        yq[q] += m[e] * rq[p2];
        if (p2 != q) yq[p2] += m[e] * rq[q];
      }
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) y[q * cpad + c] = yq[q];
  }
}

// Weighted inverses of the cell blocks of the face system (see mimpy_auxmg::binv).  Off-diagonal entries of a block
// belong to the cell alone (two faces of a hex mesh share at most one cell) and are looked up in the rows of the
// assembled matrix; the diagonal is the ASSEMBLED one (1 / dinv: complete also on the interface rows of a slab
// partition, where the matrix rows hold this rank's half).  A face that is not a flux DOF leaves an identity row.
// Gauss-Jordan without pivoting in registers: the block is a principal submatrix of an SPD matrix.
template <bool SELL>
__global__ void __launch_bounds__(128)
k_aux_cellblocks(AuxHex hx, const int* __restrict__ f2f, const uint8_t* __restrict__ face_flag,
                 const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ vals,
                 const double* __restrict__ dinv, const double* __restrict__ s_a, float* __restrict__ binv, size_t cpad,
                 double w_two) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= hx.nx * hx.ny * hx.nz) return;
  const int i = c % hx.nx, j = (c / hx.nx) % hx.ny, k = c / (hx.nx * hx.ny);
  const int cb = hx.cbase(i, j, k);
  const int f[6] = {cb, cb + 1, cb + 2, hx.xp(i, j, k), hx.yp(i, j, k), hx.zp(i, j, k)};
  const bool inner[6] = {k > 0, j > 0, i > 0, i + 1 < hx.nx, j + 1 < hx.ny, k + 1 < hx.nz};
  int g[6];
  double wt[6];
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    g[q] = f2f[f[q]];
    const bool two = inner[q] || (face_flag && (face_flag[f[q]] & 3));
    wt[q] = two ? w_two : 1.;
  }
  double B[6][6];
#pragma unroll
  for (int q = 0; q < 6; ++q)
#pragma unroll
    for (int p2 = 0; p2 < 6; ++p2) B[q][p2] = (q == p2) ? 1. : 0.;
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    if (g[q] < 0) continue;
    B[q][q] = 1. / dinv[g[q]];
    int base, stride, len;
    if (SELL) {
      const int s0 = ptr[g[q] >> 5];
      base = s0 + (g[q] & 31);
      stride = 32;
      len = (ptr[(g[q] >> 5) + 1] - s0) >> 5;
    } else {
      base = ptr[g[q]];
      stride = 1;
      len = ptr[g[q] + 1] - base;
    }
    for (int e = 0; e < len; ++e) {
      const int col = idx[base + e * stride];
      const double v = vals[base + e * stride];
#pragma unroll
      for (int p2 = q + 1; p2 < 6; ++p2)
        if (col == g[p2]) B[q][p2] = v;   // (a padding entry names its own row: never a partner)
    }
  }
#pragma unroll
  for (int q = 0; q < 6; ++q)
#pragma unroll
    for (int p2 = q + 1; p2 < 6; ++p2) B[p2][q] = B[q][p2];
#pragma unroll
  for (int kk = 0; kk < 6; ++kk) {
    const double pinv = 1. / B[kk][kk];
#pragma unroll
    for (int p2 = 0; p2 < 6; ++p2)
      if (p2 != kk) B[kk][p2] *= pinv;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      if (q == kk) continue;
      const double fqk = B[q][kk];
#pragma unroll
      for (int p2 = 0; p2 < 6; ++p2)
        if (p2 != kk) B[q][p2] -= fqk * B[kk][p2];
      B[q][kk] = -fqk * pinv;
    }
    B[kk][kk] = pinv;
  }
  const double inv_sa = 1. / s_a[0];
  int e = 0;
#pragma unroll
  for (int q = 0; q < 6; ++q)
#pragma unroll
    for (int p2 = q; p2 < 6; ++p2, ++e)  // symmetrised: rounding leaves the two triangles a few ulps apart
      binv[e * cpad + c] = (float)(0.5 * (B[q][p2] + B[p2][q]) * wt[q] * wt[p2] * inv_sa);
}

// z = [owned rows] D^-1 r + Pi e ; partial sums of r.z over ALL local rows (interface rows hold this
// rank's share of z, r is identical on both ranks, so the shares add up to r.z across ranks).  Every
// face is written by exactly one cell: its second cell (the cell whose z-/y-/x- face it is), or its
// only cell -- three consecutive rows per cell.  FP32 levels: e comes back as (s_b / s_A) x the correction.
template <typename T>
__global__ void __launch_bounds__(RED_THREADS)
k_aux_face_prolong(AuxHex hx, int n_owned, double wj, const int* __restrict__ f2f, const T* __restrict__ fw,
                   const T* __restrict__ e, const double* __restrict__ dinv,
                   const double* __restrict__ r, double* __restrict__ z, double* partials,
                   unsigned int* counter, double* dst, const double* __restrict__ rr, double nrows,
                   const double* __restrict__ s_a, const float* __restrict__ y, size_t cpad, const double* __restrict__ s_y,
                   const double* __restrict__ ry) {
  __shared__ double sh[32];
  const int nc = hx.nx * hx.ny * hx.nz, sy = hx.nx, sz = hx.nx * hx.ny;
  const double back = rr ? s_a[0] / residual_scale(rr, nrows) : 1.;
  // block smoother: y holds (s_b / s_A) x the block solutions of both cells of a face (every rank adds its own cells'
  // on interface rows; the exchange-add of z completes them like the prolongated correction)
  const double yback = y ? wj * s_y[0] / (ry ? residual_scale(ry, nrows) : 1.) : 0.;
  double acc = 0.;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x) {
    const int i = c % hx.nx, j = (c / hx.nx) % hx.ny, k = c / sz;
    const int cb = hx.cbase(i, j, k);
    const double ec = back * (double)e[c];
    auto face = [&](int f, int other, int q) {  // other: the first cell of the face, or -1 when this cell is
      const int g = f2f[f];
      if (g < 0) return;
      const double rg = r[g], w1 = (double)fw[f];
      double zz;
      if (y) {
        double ys = (double)y[q * cpad + c];
        if (other >= 0) ys += (double)y[(5 - q) * cpad + other];
        zz = yback * ys;
      } else {
        zz = (g < n_owned) ? wj * dinv[g] * rg : 0.;
      }
      if (other >= 0) {
        if (w1 > 0.) zz += w1 * (back * (double)e[other]) + (1. - w1) * ec;
      } else {
        zz += w1 * ec;
      }
      z[g] = zz;
      acc += rg * zz;
    };
    face(cb, k > 0 ? c - sz : -1, 0);
    face(cb + 1, j > 0 ? c - sy : -1, 1);
    face(cb + 2, i > 0 ? c - 1 : -1, 2);
    if (i + 1 == hx.nx) face(hx.xp(i, j, k), -1, 3);
    if (j + 1 == hx.ny) face(hx.yp(i, j, k), -1, 4);
    if (k + 1 == hx.nz) face(hx.zp(i, j, k), -1, 5);
  }
  const double v[1] = {acc};
  finish_reduction<1>(v, partials, counter, dst, sh);
}

static int aux_free(mimpy_ctx* ctx, mimpy_auxmg* h) {
  if (!h) return MIMPY_OK;
  if (h->slab) cudaFreeAsync(h->slab, ctx->stream);  // stream-ordered: no device synchronisation
  if (h->binv) cudaFreeAsync(h->binv, ctx->stream);
  delete h;
  return MIMPY_OK;
}

static inline int aux_blocks(int n) { return (n + AUX_T - 1) / AUX_T; }

// the cycle on the FP32 levels of a single rank: same kernels, no exchanges
static int aux_apply32(mimpy_ctx* ctx, mimpy_auxmg* h, const AuxHex& hx, const double* dinv, const double* r, double* z,
                       double* rz_dst, int n_owned) {
  cudaStream_t st = ctx->stream;
  AuxLevel32* L = h->lev32;
  const double* rr = ctx->scalars + S_RR;  // ||r||^2 of the PCG: current when this runs (k_init / k_update_xr_gen)
  const double nrows = (double)(h->flux_dof > 0 ? h->flux_dof : 1);
  const bool blocks = h->blocks == 2;
  float* y = blocks ? static_cast<float*>(h->ycell) : nullptr;
  if (blocks)
    k_aux_face_restrict<float, true><<<aux_blocks(h->nc), AUX_T, 0, st>>>(hx, h->face_to_flux, h->fw32, r, L[0].b, rr, nrows,
                                                                         h->binv, y, h->cpad, rr);
  else
    k_aux_face_restrict<float, false><<<aux_blocks(h->nc), AUX_T, 0, st>>>(hx, h->face_to_flux, h->fw32, r, L[0].b, rr, nrows,
                                                                          nullptr, nullptr, 0, nullptr);
  const int last = h->nlev - 1;
  for (int l = 0; l < last; ++l) {
    k_aux_presmooth<float><<<aux_blocks(L[l].n), AUX_T, 0, st>>>(L[l], h->omega);
    k_aux_restrict<float><<<aux_blocks(L[l + 1].n), AUX_T, 0, st>>>(L[l], L[l + 1]);
  }
  k_aux_coarse_solve<float><<<1, AUX_COARSEST, 0, st>>>(L[last], h->cinv, h->sdev);
  for (int l = last - 1; l >= 0; --l)
    k_aux_prolong_smooth<float, false><<<aux_blocks(L[l].n), AUX_T, 0, st>>>(L[l], L[l + 1], h->omega, h->sdev + 1, nullptr, nullptr, nullptr);
  const int grid = ctx->sm_count * 8;
  k_aux_face_prolong<float><<<grid, RED_THREADS, 0, st>>>(hx, n_owned, h->wj, h->face_to_flux, h->fw32, L[0].x2, dinv, r, z,
                                                          ctx->partials, ctx->counters + 4, rz_dst, rr, nrows, h->sdev, y,
                                                          h->cpad, h->sdev, rr);
  KERNEL_CHECK();
  return MIMPY_OK;
}

// p = z + beta p with z = w_J D^-1 r + Pi e formed on the fly (never stored).  r . z = S_RDR + (s_A / s_b^2) S_BE is
// complete before this kernel starts, so beta = r.z / (r.z of the previous iteration) is known to every thread; the last
// block rotates the scalars (every block has read S_RZ by then).
__global__ void __launch_bounds__(RED_THREADS)
k_aux_face_prolong_p(AuxHex hx, int nf, double wj, const int* __restrict__ f2f, const float* __restrict__ fw,
                     const float* __restrict__ e, const float* __restrict__ dinv32, const double* __restrict__ r,
                     double* __restrict__ p, double* scal, unsigned int* counter, double nrows,
                     const double* __restrict__ s_a, int first) {
  const int sy = hx.nx, sz = hx.nx * hx.ny;
  const double sb = residual_scale(scal + S_RR, nrows), sa = s_a[0];
  const double back = sa / sb;
  const double rz_new = scal[S_RDR] + scal[S_BE] * (sa / (sb * sb));
  const double rz_old = scal[S_RZ];
  const double beta = (first || rz_old == 0.) ? 0. : rz_new / rz_old;
  // one thread per FACE (rows of p, r, D^-1 and the weights are then read and written as consecutive runs; a thread per
  // cell touched them with a stride of three rows): the face's cell(s) from the closed-form HexMesh numbering
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
    const int g = f2f[f];
    if (g < 0) continue;
    const int k = f / hx.L, rem = f - k * hx.L;
    int c, other = -1;  // c: the cell whose z-/y-/x- face f is, or the only cell of a face on the upper boundaries
    if (k == hx.nz) {                       // top z level: face of cell (i, j, nz-1)
      c = rem + (hx.nz - 1) * sz;
    } else {
      const int j = rem / hx.xrow, r2 = rem - j * hx.xrow;
      if (j == hx.ny) {                     // last y row of the layer: face of cell (i, ny-1, k)
        c = r2 + (hx.ny - 1) * sy + k * sz;
      } else if (r2 == 3 * hx.nx) {         // x+ boundary face of cell (nx-1, j, k)
        c = hx.nx - 1 + j * sy + k * sz;
      } else {
        const int i = r2 / 3, d = r2 - 3 * i;
        c = i + j * sy + k * sz;
        if (d == 0) other = k > 0 ? c - sz : -1;
        else if (d == 1) other = j > 0 ? c - sy : -1;
        else other = i > 0 ? c - 1 : -1;
      }
    }
    const double w1 = (double)fw[f];
    const double ec = back * (double)e[c];
    double zz = wj * (double)dinv32[g] * r[g];
    if (other >= 0) {
      if (w1 > 0.) zz += w1 * (back * (double)e[other]) + (1. - w1) * ec;
    } else {
      zz += w1 * ec;
    }
    p[g] = first ? zz : zz + beta * p[g];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(counter, 1u);
    if (t == gridDim.x - 1) {
      scal[S_RZNEW] = rz_new;
      scal[S_RZ] = rz_new;
      *counter = 0u;
    }
  }
}

__global__ void __launch_bounds__(256) k_aux_dinv32(int n, const double* __restrict__ dinv, float* __restrict__ out) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) out[g] = (float)dinv[g];
}

int mimpy_auxmg_fused(mimpy_ctx* ctx, mimpy_auxmg* h, const double* dinv, int n, mimpy_aux_fused* out) {
  // opt-in (MIMPY_B200_AUX_FUSED_P): measured 2.82 ms against 2.79 ms per iteration at 256^3 -- the 0.8 GB of z it saves
  // are paid back by the face-indexed gather of e and the second pass over D^-1; kept for smaller caches / other meshes
  if (!h->use32 || ctx->world != 1 || n != h->flux_dof || h->blocks == 2 || !getenv("MIMPY_B200_AUX_FUSED_P")) return 0;
  k_aux_dinv32<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, dinv, h->dinv32);
  if (cudaGetLastError() != cudaSuccess) return 0;
  out->dinv32 = h->dinv32;
  out->wj = h->wj;
  return 1;
}

// p = B r + beta p on the FP32 levels of a single rank (see k_aux_face_prolong_p); S_RR and S_RDR are current
int mimpy_auxmg_apply_p(mimpy_ctx* ctx, mimpy_auxmg* h, const double* r, double* p, int first) {
  cudaStream_t st = ctx->stream;
  AuxLevel32* L = h->lev32;
  AuxHex hx;
  hx.nx = L[0].nx; hx.ny = L[0].ny; hx.nz = L[0].nz;
  hx.L = hx.nx * hx.ny + hx.nx * (hx.ny + 1) + (hx.nx + 1) * hx.ny;
  hx.xrow = 3 * hx.nx + 1;
  const double* rr = ctx->scalars + S_RR;
  const double nrows = (double)(h->flux_dof > 0 ? h->flux_dof : 1);
  const int grid = ctx->sm_count * 8;
  k_aux_face_restrict<float, false><<<aux_blocks(h->nc), AUX_T, 0, st>>>(hx, h->face_to_flux, h->fw32, r, L[0].b, rr, nrows,
                                                                        nullptr, nullptr, 0, nullptr);
  const int last = h->nlev - 1;
  for (int l = 0; l < last; ++l) {
    k_aux_presmooth<float><<<aux_blocks(L[l].n), AUX_T, 0, st>>>(L[l], h->omega);
    k_aux_restrict<float><<<aux_blocks(L[l + 1].n), AUX_T, 0, st>>>(L[l], L[l + 1]);
  }
  k_aux_coarse_solve<float><<<1, AUX_COARSEST, 0, st>>>(L[last], h->cinv, h->sdev);
  for (int l = last - 1; l >= 1; --l)
    k_aux_prolong_smooth<float, false><<<aux_blocks(L[l].n), AUX_T, 0, st>>>(L[l], L[l + 1], h->omega, h->sdev + 1, nullptr, nullptr, nullptr);
  if (last >= 1) {
    // (as many blocks as the reduction has partial slots for: neighbouring cells stay neighbours in time)
    const int dot_grid = aux_blocks(L[0].n) < 8192 ? aux_blocks(L[0].n) : 8192;
    k_aux_prolong_smooth<float, true><<<dot_grid, AUX_T, 0, st>>>(L[0], L[1], h->omega, h->sdev + 1, ctx->partials, ctx->counters + 4,
                                                               ctx->scalars + S_BE);
  } else {  // a mesh of one level: the coarse solve is the whole cycle
    k_aux_dot32<<<1, AUX_T, 0, st>>>(L[0].n, L[0].b, L[0].x2, ctx->scalars + S_BE);
  }
  k_aux_face_prolong_p<<<grid, RED_THREADS, 0, st>>>(hx, h->nf, h->wj, h->face_to_flux, h->fw32, L[0].x2, h->dinv32, r, p, ctx->scalars,
                                                    ctx->counters + 2, nrows, h->sdev, first);
  KERNEL_CHECK();
  return MIMPY_OK;
}

// z = B r, scalar r.z (this rank's share) -> *rz_dst.  Stream-ordered, no host synchronisation.
int mimpy_auxmg_apply(mimpy_ctx* ctx, mimpy_auxmg* h, const double* dinv, const double* r, double* z,
                      double* rz_dst, int n_owned) {
  cudaStream_t st = ctx->stream;
  AuxLevel* L = h->lev;
  AuxHex hx;
  hx.nx = L[0].nx; hx.ny = L[0].ny; hx.nz = L[0].nz;
  hx.L = hx.nx * hx.ny + hx.nx * (hx.ny + 1) + (hx.nx + 1) * hx.ny;
  hx.xrow = 3 * hx.nx + 1;
  if (h->use32) return aux_apply32(ctx, h, hx, dinv, r, z, rz_dst, n_owned);
  const bool blocks = h->blocks == 2;
  float* y = blocks ? static_cast<float*>(h->ycell) : nullptr;
  // scale of the FP32 block solves: this rank's ||r||^2 (k_update_xr_gen's partial sum until apply_aux all-reduces it
  // AFTER this application: the same value in the restriction and in the prolongation)
  const double* ry = ctx->scalars + S_RR;
  const double nrows_y = (double)(h->flux_dof > 0 ? h->flux_dof : 1);
  if (blocks)
    k_aux_face_restrict<double, true><<<aux_blocks(h->nc), AUX_T, 0, st>>>(hx, h->face_to_flux, h->fw, r, L[0].b, nullptr,
                                                                          nrows_y, h->binv, y, h->cpad, ry);
  else
    k_aux_face_restrict<double, false><<<aux_blocks(h->nc), AUX_T, 0, st>>>(hx, h->face_to_flux, h->fw, r, L[0].b, nullptr, 1.,
                                                                           nullptr, nullptr, 0, nullptr);
  const int last = h->nlev - 1;
  auto ghosts = [&](double* v, const AuxLevel& lv) {  // both ghost layers of a level vector
    mimpy_xchg x;
    x.n = lv.nx * lv.ny;
    x.send_lo = 0;
    x.send_hi = (size_t)lv.n - x.n;
    x.recv_lo = (size_t)lv.n;
    x.recv_hi = (size_t)lv.n + x.n;
    return mimpy_layer_exchange(ctx, v, x);
  };
  int rc;
  for (int l = 0; l < last; ++l) {
    k_aux_presmooth<double><<<aux_blocks(L[l].n), AUX_T, 0, st>>>(L[l], h->omega);
    if (h->global && (rc = ghosts(L[l].x, L[l]))) return rc;
    k_aux_restrict<double><<<aux_blocks(L[l + 1].n), AUX_T, 0, st>>>(L[l], L[l + 1]);
  }
  if (h->global) {
    // agglomerated levels: every rank holds the whole coarse grid and cycles on it without communication
    AuxLevel* G = h->glev;
    rc = mimpy_allgather_f64(ctx, L[last].b, L[last].n, G[0].b);
    if (rc) return rc;
    const int glast = h->nglob - 1;
    for (int l = 0; l < glast; ++l) {
      k_aux_presmooth<double><<<aux_blocks(G[l].n), AUX_T, 0, st>>>(G[l], h->omega);
      k_aux_restrict<double><<<aux_blocks(G[l + 1].n), AUX_T, 0, st>>>(G[l], G[l + 1]);
    }
    k_aux_coarse_solve<double><<<1, AUX_COARSEST, 0, st>>>(G[glast], h->cinv, nullptr);
    for (int l = glast - 1; l >= 0; --l)
      k_aux_prolong_smooth<double, false><<<aux_blocks(G[l].n), AUX_T, 0, st>>>(G[l], G[l + 1], h->omega, h->sdev + 1, nullptr, nullptr, nullptr);
    const int nxy = L[last].nx * L[last].ny;
    k_aux_global_extract<<<aux_blocks(L[last].n + 2 * nxy), AUX_T, 0, st>>>(L[last].n, nxy, ctx->rank, ctx->world, G[0].x2,
                                                                        L[last].x2);
  } else {
    k_aux_coarse_solve<double><<<1, AUX_COARSEST, 0, st>>>(L[last], h->cinv, nullptr);
  }
  for (int l = last - 1; l >= 0; --l) {
    k_aux_prolong_smooth<double, false><<<aux_blocks(L[l].n), AUX_T, 0, st>>>(L[l], L[l + 1], h->omega, h->sdev + 1, nullptr, nullptr, nullptr);
    if (h->global && l > 0 && (rc = ghosts(L[l].x2, L[l]))) return rc;
  }
  const int grid = ctx->sm_count * 8;
  k_aux_face_prolong<double><<<grid, RED_THREADS, 0, st>>>(hx, n_owned, h->wj, h->face_to_flux, h->fw, L[0].x2, dinv, r, z,
                                                           ctx->partials, ctx->counters + 4, rz_dst, nullptr, nrows_y, nullptr,
                                                           y, h->cpad, h->sdev, ry);
  KERNEL_CHECK();
  return MIMPY_OK;
}

extern "C" {

int mimpy_b200_auxmg_create(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym,
                            void** handle) {
  REQUIRE(ctx && mesh && sym && handle, "NULL argument");
  REQUIRE(mesh->uniform_faces == 6 && mesh->hex_nx > 0 && mesh->hex_ny > 0 && mesh->hex_nz > 0 &&
              (long long)mesh->hex_nx * mesh->hex_ny * mesh->hex_nz == mesh->nc,
          "the auxiliary-space multigrid needs a structured hex mesh (DeviceMesh.hex)");
  mimpy_auxmg* h = new mimpy_auxmg();
  memset(h, 0, sizeof(*h));
  h->nc = mesh->nc;
  h->nf = mesh->nf;
  h->flux_dof = sym->flux_dof;
  // damping of the Jacobi smoother and over-correction of the piecewise-constant coarse space, measured
  // (scripts/exp_iters.py, rtol 1e-8 on the saddle residual): (0.8, 2.0) 110 iterations at 256^3 and 1600 on C2,
  // (0.9, 1.5) 95 and 1085, (1.0, 1.3) 105 and 960, (0.9, 1.7) 95 and 1255, scale 1.0: 135 at 128^3, 3.0: 180
  h->omega = 0.9;
  h->scale = 0.;  // 0: chosen at set-up from the anisotropy of the cell operator (k_aux_pick_scale)
  h->wj = 1.0;
  if (const char* e = getenv("MIMPY_B200_AUX_WJ")) h->wj = atof(e);
  if (const char* e = getenv("MIMPY_B200_AUX_OMEGA")) h->omega = atof(e);
  if (const char* e = getenv("MIMPY_B200_AUX_SCALE")) h->scale = atof(e);
  h->cell_faces = mesh->cell_faces;
  h->face_to_flux = sym->face_to_flux;
  h->face_cell = sym->face_cell;
  h->face_flag = sym->face_flag;
  int nx = mesh->hex_nx, ny = mesh->hex_ny, nz = mesh->hex_nz;
  size_t total = 0;
  auto pad = [](size_t n) { return (n + 31) & ~(size_t)31; };  // keep every array 256-byte aligned
  // slab partition with identical slabs on every rank: share the hierarchy (see the header)
  int coarsest = AUX_COARSEST;
  h->global = 0;
  if (ctx->world > 1 && ctx->world <= 16 && ctx->nccl_comm != nullptr) {
    const char* env = getenv("MIMPY_B200_AUX_GLOBAL");
    if (!(env && env[0] == '0')) {
      double* d = ctx->scalars + 16;  // world * 3 doubles of scratch (64 scalars are allocated)
      double mine[48];
      for (int q = 0; q < ctx->world * 3; ++q) mine[q] = 0.;
      mine[ctx->rank * 3] = nx; mine[ctx->rank * 3 + 1] = ny; mine[ctx->rank * 3 + 2] = nz;
      CUDA_TRY(cudaMemcpyAsync(d, mine, sizeof(double) * ctx->world * 3, cudaMemcpyHostToDevice, ctx->stream));
      int rc = mimpy_allreduce_f64(ctx, d, ctx->world * 3);
      if (rc) { delete h; return rc; }
      CUDA_TRY(cudaMemcpyAsync(mine, d, sizeof(double) * ctx->world * 3, cudaMemcpyDeviceToHost, ctx->stream));
      CUDA_TRY(cudaStreamSynchronize(ctx->stream));
      bool same = true;
      for (int r = 0; r < ctx->world; ++r)
        same = same && mine[r * 3] == nx && mine[r * 3 + 1] == ny && mine[r * 3 + 2] == nz;
      if (same) {
        h->global = 1;
        const char* agg = getenv("MIMPY_B200_AUX_AGG");
        coarsest = agg ? atoi(agg) : AUX_AGG_DEFAULT;
        if (coarsest < 1) coarsest = 1;
      }
    }
  }
  auto level_doubles = [&](int lx, int ly, int lz) {
    const size_t n = (size_t)lx * ly * lz;
    return 6 * pad(n) + 2 * pad(n + 2 * (size_t)lx * ly) + 2 * pad((size_t)lx * ly);
  };
  auto build_levels = [&](int stop_at) {
    int lx = mesh->hex_nx, ly = mesh->hex_ny, lz = mesh->hex_nz;
    h->nlev = 0;
    total = 0;
    for (;;) {
      if (h->nlev >= AUX_MAX_LEVELS) break;
      AuxLevel& L = h->lev[h->nlev++];
      L.nx = lx; L.ny = ly; L.nz = lz; L.n = lx * ly * lz;
      total += level_doubles(lx, ly, lz);
      if (L.n <= stop_at) break;
      lx = (lx + 1) / 2; ly = (ly + 1) / 2; lz = (lz + 1) / 2;
    }
  };
  build_levels(coarsest);
  if (h->global && (h->lev[h->nlev - 1].n > coarsest || h->nlev < 2)) {
    // a slab that is below the agglomeration threshold already (or a hierarchy that does not get there): the
    // rank-local cycle with its own coarsest level
    h->global = 0;
    build_levels(AUX_COARSEST);
  }
  h->nglob = 0;
  if (h->global) {  // the agglomerated levels: the last distributed level stacked over the ranks, then coarsened
    const AuxLevel& last = h->lev[h->nlev - 1];
    int gx = last.nx, gy = last.ny, gz = last.nz * ctx->world;
    for (;;) {
      if (h->nglob >= AUX_MAX_LEVELS) break;
      AuxLevel& G = h->glev[h->nglob++];
      G.nx = gx; G.ny = gy; G.nz = gz; G.n = gx * gy * gz;
      total += level_doubles(gx, gy, gz);
      if (G.n <= AUX_COARSEST) break;
      gx = (gx + 1) / 2; gy = (gy + 1) / 2; gz = (gz + 1) / 2;
    }
    total += pad((size_t)ctx->world * 6 * last.n);
  }
  const int coarse_n = h->global ? h->glev[h->nglob - 1].n : h->lev[h->nlev - 1].n;
  if (coarse_n > AUX_COARSEST) {
    delete h;
    mimpy_set_error("auxmg_create: too many levels");
    return MIMPY_E_INVALID;
  }
  total += pad((size_t)h->nc * 6) + pad((size_t)h->nf) + pad((size_t)h->flux_dof) +
           pad((size_t)AUX_COARSEST * AUX_COARSEST);
  // FP32 mirrors of the levels (single rank): 4 coefficient arrays and 3 vectors per level, in floats
  h->use32 = 0;
  {
    const char* env = getenv("MIMPY_B200_AUX_FP32");
    if (ctx->world == 1 && !h->global && !(env && env[0] == '0')) h->use32 = 1;
  }
  size_t total32 = 0;  // in doubles
  if (h->use32)
    for (int l = 0; l < h->nlev; ++l) total32 += 7 * pad(((size_t)h->lev[l].n + 1) / 2);
  if (h->use32) total32 += pad(((size_t)h->nf + 1) / 2) + pad(((size_t)h->flux_dof + 1) / 2);
  total += total32 + pad(8);
  // cell-block smoother: storage comes with the first mimpy_b200_auxmg_cellblocks call
  h->blocks = 0;
  h->cpad = pad((size_t)h->nc);
  static bool pool_configured[64] = {false};  // per device
  bool& pool_done = pool_configured[ctx->device & 63];
  if (!pool_done) {  // keep freed blocks in the pool: set-up is repeated every Newton step
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    pool_done = true;
  }
  cudaError_t e = cudaMallocAsync(&h->slab, sizeof(double) * total, ctx->stream);
  if (e != cudaSuccess) {
    delete h;
    mimpy_set_error("auxmg_create: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return MIMPY_E_CUDA;
  }
  double* cur = h->slab;
  auto take = [&](size_t n) { double* p = cur; cur += pad(n); return p; };
  for (int l = 0; l < h->nlev; ++l) {
    AuxLevel& L = h->lev[l];
    L.tx = take(L.n); L.ty = take(L.n); L.tz = take(L.n); L.dd = take(L.n); L.dg = take(L.n);
    L.b = take(L.n);
    L.x = take((size_t)L.n + 2 * (size_t)L.nx * L.ny);
    L.x2 = take((size_t)L.n + 2 * (size_t)L.nx * L.ny);
    L.tzi = take((size_t)L.nx * L.ny);
    L.bzi = take((size_t)L.nx * L.ny);
    L.has_lo = (h->global && ctx->rank > 0) ? 1 : 0;
    L.has_hi = (h->global && ctx->rank < ctx->world - 1) ? 1 : 0;
  }
  for (int l = 0; l < h->nglob; ++l) {
    AuxLevel& G = h->glev[l];
    G.tx = take(G.n); G.ty = take(G.n); G.tz = take(G.n); G.dd = take(G.n); G.dg = take(G.n);
    G.b = take(G.n);
    G.x = take((size_t)G.n + 2 * (size_t)G.nx * G.ny);
    G.x2 = take((size_t)G.n + 2 * (size_t)G.nx * G.ny);
    G.tzi = take((size_t)G.nx * G.ny);
    G.bzi = take((size_t)G.nx * G.ny);
    G.has_lo = G.has_hi = 0;
  }
  h->gpack = h->global ? take((size_t)ctx->world * 6 * h->lev[h->nlev - 1].n) : nullptr;
  h->wc = take((size_t)h->nc * 6);
  h->fw = take((size_t)h->nf);
  h->tsum = take((size_t)h->flux_dof);
  h->cinv = take((size_t)AUX_COARSEST * AUX_COARSEST);
  h->sdev = take(8);
  h->fw32 = nullptr;
  h->dinv32 = nullptr;
  if (h->use32) {
    h->fw32 = reinterpret_cast<float*>(take(((size_t)h->nf + 1) / 2));
    h->dinv32 = reinterpret_cast<float*>(take(((size_t)h->flux_dof + 1) / 2));
    for (int l = 0; l < h->nlev; ++l) {
      const AuxLevel& L = h->lev[l];
      AuxLevel32& M = h->lev32[l];
      M.nx = L.nx; M.ny = L.ny; M.nz = L.nz; M.n = L.n;
      auto takef = [&]() { return reinterpret_cast<float*>(take(((size_t)L.n + 1) / 2)); };
      M.tx = takef(); M.ty = takef(); M.tz = takef(); M.dg = takef();
      M.b = takef(); M.x = takef(); M.x2 = takef();
      M.dd = nullptr; M.tzi = nullptr; M.bzi = nullptr;
      M.has_lo = M.has_hi = 0;
    }
  }
  h->binv = nullptr;
  h->ycell = nullptr;
  *handle = h;
  return MIMPY_OK;
}

int mimpy_b200_auxmg_cellblocks(mimpy_ctx* ctx, void* handle, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym,
                                const int32_t* ptr, const int32_t* idx, const double* vals, int32_t sell,
                                const double* diag_inv) {
  REQUIRE(ctx && handle && mesh && sym && ptr && idx && vals && diag_inv, "NULL argument");
  mimpy_auxmg* h = static_cast<mimpy_auxmg*>(handle);
  REQUIRE(h->nc == mesh->nc && h->nf == mesh->nf && h->flux_dof == sym->flux_dof, "handle/mesh mismatch");
  {
    const char* env = getenv("MIMPY_B200_AUX_BLOCKS");
    if (env && env[0] == '0') return MIMPY_OK;  // switched off: the Jacobi term stays
  }
  if (!h->binv) {  // 21 floats per cell + the block solutions (6 floats per cell)
    const size_t bytes = sizeof(float) * (21 + 6) * h->cpad;
    void* blk = nullptr;
    cudaError_t e = cudaMallocAsync(&blk, bytes, ctx->stream);
    if (e != cudaSuccess) {
      mimpy_set_error("auxmg_cellblocks: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return MIMPY_E_CUDA;
    }
    h->binv = static_cast<float*>(blk);
    h->ycell = h->binv + 21 * h->cpad;
    h->blocks = 1;
  }
  AuxHex hx;
  hx.nx = mesh->hex_nx; hx.ny = mesh->hex_ny; hx.nz = mesh->hex_nz;
  hx.L = hx.nx * hx.ny + hx.nx * (hx.ny + 1) + (hx.nx + 1) * hx.ny;
  hx.xrow = 3 * hx.nx + 1;
  const int grid = (h->nc + 127) / 128;
  // weight of a block on a face with two cells (the two blocks of a face add up): 1/2 measured best (prototype, 24^3 and
  // the C2 shape: 0.5 -> 28 / 90 iterations, 0.75 -> 32 / 92, 1.0 -> 36 / 97)
  double w_two = sqrt(0.5);
  if (const char* e = getenv("MIMPY_B200_AUX_BW")) w_two = sqrt(atof(e));
  if (sell)
    k_aux_cellblocks<true><<<grid, 128, 0, ctx->stream>>>(hx, sym->face_to_flux, sym->face_flag, ptr, idx, vals, diag_inv,
                                                          h->sdev, h->binv, h->cpad, w_two);
  else
    k_aux_cellblocks<false><<<grid, 128, 0, ctx->stream>>>(hx, sym->face_to_flux, sym->face_flag, ptr, idx, vals, diag_inv,
                                                           h->sdev, h->binv, h->cpad, w_two);
  KERNEL_CHECK();
  h->blocks = 2;
  return MIMPY_OK;
}

// number of cells whose permeability tensor is clearly anisotropic: tr(K)^3 / (27 det K) > threshold (1 for an isotropic
// tensor by the AM-GM inequality, 1.54 for eigenvalues 3:1:1); deterministic two-stage sum
__global__ void __launch_bounds__(RED_THREADS)
k_aniso_count(int nc, const double* __restrict__ k, double threshold, double* partials, unsigned int* counter, double* dst) {
  __shared__ double sh[32];
  double v[1] = {0.};
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x) {
    const double* a = k + 9 * (size_t)c;
    const double tr = a[0] + a[4] + a[8];
    const double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) +
                       a[2] * (a[3] * a[7] - a[4] * a[6]);
    const double t = tr * tr * tr / (27. * det);
    if (!(t <= threshold)) v[0] += 1.;  // (NaN / inf count as anisotropic)
  }
  finish_reduction<1>(v, partials, counter, dst, sh);
}

int mimpy_b200_k_anisotropy(mimpy_ctx* ctx, int32_t nc, const double* cell_k, double threshold, double* count_out) {
  REQUIRE(ctx && cell_k && count_out, "NULL argument");
  if (nc <= 0) {
    *count_out = 0.;
    return MIMPY_OK;
  }
  int grid = (nc + RED_THREADS - 1) / RED_THREADS;
  if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
  k_aniso_count<<<grid, RED_THREADS, 0, ctx->stream>>>(nc, cell_k, threshold, ctx->partials, ctx->counters + 6,
                                                       ctx->scalars + 40);
  KERNEL_CHECK();
  if (ctx->world > 1) {  // every rank of a partition must take the same decision
    int rc = mimpy_allreduce_f64(ctx, ctx->scalars + 40, 1);
    if (rc) return rc;
  }
  CUDA_TRY(cudaMemcpyAsync(ctx->host_scalars + 40, ctx->scalars + 40, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  *count_out = ctx->host_scalars[40];
  return MIMPY_OK;
}

int mimpy_b200_auxmg_destroy(mimpy_ctx* ctx, void* handle) {
  REQUIRE(ctx, "NULL context");
  if (!handle) return MIMPY_OK;
  return aux_free(ctx, static_cast<mimpy_auxmg*>(handle));
}

int mimpy_b200_auxmg_setup(mimpy_ctx* ctx, void* handle, const mimpy_mesh_view* mesh,
                           const mimpy_symbolic* sym, const double* w_e, double alpha,
                           const double* c_diag, const double* tau) {
  REQUIRE(ctx && handle && mesh && sym && w_e, "NULL argument");
  mimpy_auxmg* h = static_cast<mimpy_auxmg*>(handle);
  REQUIRE(h->nc == mesh->nc && h->nf == mesh->nf && h->flux_dof == sym->flux_dof, "handle/mesh mismatch");
  cudaStream_t st = ctx->stream;
  h->cell_faces = mesh->cell_faces;
  h->face_to_flux = sym->face_to_flux;
  h->face_cell = sym->face_cell;
  h->face_flag = sym->face_flag;
  int rc = MIMPY_OK;
  if (tau && ctx->world == 1 && !sym->face_flag && ((uintptr_t)tau & 15u) == 0) {
    // one rank, half transmissibilities at hand: closed-form partners, no sums over faces, no copy
    AuxHex hx;
    hx.nx = mesh->hex_nx; hx.ny = mesh->hex_ny; hx.nz = mesh->hex_nz;
    hx.L = hx.nx * hx.ny + hx.nx * (hx.ny + 1) + (hx.nx + 1) * hx.ny;
    hx.xrow = 3 * hx.nx + 1;
    k_aux_level0_hex<<<aux_blocks(h->nc), AUX_T, 0, st>>>(hx, sym->face_to_flux, tau, mesh->cell_volume, alpha, c_diag,
                                                         h->fw, h->use32 ? h->fw32 : nullptr, h->lev[0]);
    KERNEL_CHECK();
  } else {
    CUDA_TRY(cudaMemsetAsync(h->tsum, 0, sizeof(double) * (size_t)(h->flux_dof ? h->flux_dof : 1), st));
    CUDA_TRY(cudaMemsetAsync(h->fw, 0, sizeof(double) * (size_t)h->nf, st));
    if (tau)
      k_aux_tau_given<<<(unsigned)(((long long)h->nc * 6 + AUX_T - 1) / AUX_T), AUX_T, 0, st>>>(h->nc, mesh->cell_faces, sym->face_to_flux,
                                                                                              tau, h->wc, h->tsum);
    else
      k_aux_tau<<<aux_blocks(h->nc), AUX_T, 0, st>>>(h->nc, mesh->cell_faces, sym->face_to_flux, sym->m_ptr,
                                                    mesh->face_geom, w_e, h->wc, h->tsum);
    KERNEL_CHECK();
    rc = mimpy_halo_exchange_add(ctx, h->tsum);  // interface faces: tau of the cell on the other rank
    if (rc) return rc;
    k_aux_level0<<<aux_blocks(h->nc), AUX_T, 0, st>>>(h->nc, mesh->cell_faces, sym->face_to_flux, sym->face_cell,
                                                     sym->face_flag, h->tsum, mesh->cell_volume, alpha, c_diag,
                                                     h->wc, h->fw, h->use32 ? h->fw32 : nullptr, h->lev[0]);
  }
  k_aux_scale<<<1, 32, 0, st>>>(tau ? tau : h->wc, h->nc * 6 < 256 ? h->nc * 6 : 256, h->sdev);
  if (h->blocks == 2) h->blocks = 1;  // the blocks of the previous matrix are stale until auxmg_cellblocks runs again
  for (int l = 0; l < h->nlev; ++l) {
    if (l > 0) k_aux_coarsen<<<aux_blocks(h->lev[l].n), AUX_T, 0, st>>>(h->lev[l - 1], h->lev[l]);
    k_aux_diag<<<aux_blocks(h->lev[l].n), AUX_T, 0, st>>>(h->lev[l], h->lev32[l], h->use32 ? h->sdev : nullptr);
  }
  {  // over-correction factor -> sdev[1]
    const int n0 = h->lev[0].n;
    const int grid = aux_blocks(n0) < ctx->sm_count * 4 ? aux_blocks(n0) : ctx->sm_count * 4;
    if (h->scale <= 0.) {
      k_aux_aniso_sums<<<grid, RED_THREADS, 0, st>>>(h->lev[0], ctx->partials, ctx->counters + 6, h->sdev + 2);
      KERNEL_CHECK();
      if (ctx->world > 1) {
        rc = mimpy_allreduce_f64(ctx, h->sdev + 2, 3);
        if (rc) return rc;
      }
    }
    k_aux_pick_scale<<<1, 1, 0, st>>>(h->sdev + 2, h->scale, h->sdev + 1);
  }
  const AuxLevel& last = h->lev[h->nlev - 1];
  if (h->global) {
    const size_t pack = (size_t)ctx->world * 6 * last.n;
    CUDA_TRY(cudaMemsetAsync(h->gpack, 0, sizeof(double) * pack, st));
    k_aux_global_pack<<<aux_blocks(last.n), AUX_T, 0, st>>>(last, ctx->rank, h->gpack);
    rc = mimpy_allreduce_f64(ctx, h->gpack, (int)pack);
    if (rc) return rc;
    k_aux_global_unpack<<<aux_blocks(h->glev[0].n), AUX_T, 0, st>>>(last.n, last.nx * last.ny, ctx->world, h->gpack, h->glev[0]);
    for (int l = 0; l < h->nglob; ++l) {
      if (l > 0) k_aux_coarsen<<<aux_blocks(h->glev[l].n), AUX_T, 0, st>>>(h->glev[l - 1], h->glev[l]);
      k_aux_diag<<<aux_blocks(h->glev[l].n), AUX_T, 0, st>>>(h->glev[l], AuxLevel32(), nullptr);
    }
    k_aux_dense_inverse<<<1, AUX_T, 0, st>>>(h->glev[h->nglob - 1], h->cinv);
  } else {
    k_aux_dense_inverse<<<1, AUX_T, 0, st>>>(last, h->cinv);
  }
  KERNEL_CHECK();
  return MIMPY_OK;
}

}  // extern "C"
