/*
 * mimpy_b200 -- C-ABI of the B200-native MFD hot path (libmimpy_b200.so).
 *
 * Drop-in boundary: these entry points replace, one for one, the reference's native operator
 * boundary (its three Cython modules) and the Python hot loops of mimpy.mfd.MFD /
 * mimpy.models.TwoPhase that sit directly above it.  File:line citations are relative to the
 * reference tree (ohinai/mimpy).  INTEGRATION.md shows the ctypes stubs a mimpy maintainer
 * would add at the cited call sites.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (MIMPY_E_*); the message for the last
 *     error on the calling thread is returned by mimpy_b200_last_error().
 *   - all array arguments are DEVICE pointers unless the name ends in _host; the caller owns
 *     every buffer, nothing is retained beyond the call (the reference natives have the same
 *     contract: caller-owned, in-place, mfd_cython.pyx:6-10, mesh_cython.pyx:5-17).
 *   - work is enqueued on the context's stream; functions that return sizes synchronise.
 *   - indices are int32 (reference: array('i') / np.dtype('i'), mfd.py:557-558), values are
 *     FP64, orientations are int8 (+1/-1).
 *   - one host thread per context.
 */
#ifndef MIMPY_B200_H
#define MIMPY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIMPY_OK 0
#define MIMPY_E_INVALID (-1) /* bad argument                               */
#define MIMPY_E_CUDA (-2)    /* a CUDA runtime call failed                 */
#define MIMPY_E_UNSUPPORTED (-3)
#define MIMPY_E_NOCONV (-4)  /* solver hit maxit before reaching tolerance */
#define MIMPY_E_NCCL (-5)

#define MIMPY_MAX_CELL_FACES 32 /* one warp lane per face of a cell */

/* flags of mfd_local_matrices / mfd_numeric */
#define MIMPY_FLAG_K_UNITY 1        /* N_E = s n (mfd.py:300-304)                          */
#define MIMPY_FLAG_GENERIC_KERNEL 2 /* force the generic warp-group kernel (tests, polyhedra) */
#define MIMPY_FLAG_DIRECT_KERNEL 4  /* hex meshes: thread-per-cell direct stores instead of bricks */

typedef struct mimpy_ctx mimpy_ctx;

int mimpy_b200_version(void);
const char* mimpy_b200_last_error(void);
/* stream: a cudaStream_t (may be NULL = legacy default stream). */
int mimpy_b200_create(int device, void* stream, mimpy_ctx** out);
int mimpy_b200_destroy(mimpy_ctx* ctx);
int mimpy_b200_set_stream(mimpy_ctx* ctx, void* stream);
int mimpy_b200_sync(mimpy_ctx* ctx);

/* ---------------------------------------------------------------------------------------
 * SoA mesh view (device).  face_geom is the packed per-face record
 *   [area, n_x | n_y, n_z | x_f, y_f | z_f, bind]   (8 doubles = 64 B = two 32-B sectors)
 * gathered once per (cell, face) by the M_E kernels.  Written ONLY by mimpy_b200_geom_hex_faces /
 * mimpy_b200_face_pack: its four 16-byte chunks are stored in an XOR swizzle keyed by the face id (chunk c
 * at position c ^ ((f >> 1) & 3), csrc/common.cuh) so that runs of records bulk-copied into shared
 * memory are read without bank conflicts; `bind` belongs to the library (row offset + flux-DOF flags of
 * the symbolic plan last used with these records; the array is therefore written by the numeric assembly
 * of structured hex meshes although the view declares it const).  cell_* arrays keep the reference's own
 * layouts (mesh.py:246-255: cell_volume[nc], cell_real_centroid[nc,3], cell_k[nc,9]).
 * ------------------------------------------------------------------------------------- */
typedef struct {
  int32_t nc;
  int32_t nf;
  int32_t max_faces;          /* max faces per cell (<= MIMPY_MAX_CELL_FACES)             */
  int32_t uniform_faces;      /* >0: every cell has exactly this many faces (hex: 6)      */
  int32_t hex_nx, hex_ny, hex_nz; /* >0: structured hex mesh with HexMesh's cell/face numbering
                                 (hexmesh.py:172-243) and that many cells per axis; enables the
                                 brick-tiled assembly kernel.  0 for any other mesh.           */
  int32_t reserved0;
  const int32_t* cell_ptr;    /* nc+1 offsets into cell_faces / cell_sigma                */
  const int32_t* cell_faces;  /* global face ids, reference order (mesh.py:1034-1042)     */
  const int8_t* cell_sigma;   /* cell_normal_orientation (mesh.py:1044-1056)              */
  const double* face_geom;    /* nf*8                                                     */
  const double* cell_k;       /* nc*9 row-major (mesh.py:1163-1173)                       */
  const double* cell_centroid;/* nc*3 (real, or shifted when the mesh uses shifted ones)  */
  const double* cell_volume;  /* nc                                                       */
} mimpy_mesh_view;

/* -- (0) G2: structured hex topology ----------------------------------------------------------
 * replaces the Python loops of HexMesh.build_mesh / _build_faces (hexmesh.py:161-234, 269-350):
 * points (i dx, j dy, k dz), the interleaved z/y/x face numbering, face point lists, cells
 * [z-,y-,x-,x+,y+,z+] with orientations [-1,-1,-1,1,1,1].  Any output may be NULL.
 * Sizes: points 3*(cx+1)(cy+1)(cz+1); face_points 4*nf; cell_faces/cell_sigma 6*cx*cy*cz. */
int mimpy_b200_hex_topology(mimpy_ctx* ctx, int32_t cx, int32_t cy, int32_t cz, double dx, double dy,
                            double dz, double* points, int32_t* face_points, int32_t* cell_faces,
                            int8_t* cell_sigma);

/* -- (1) G3-G5: quad face geometry ----------------------------------------------------------
 * replaces hexmesh_cython.all_face_areas (hexmesh_cython.pyx:30-89), all_face_normals
 * (hexmesh_cython.pyx:91-118) and HexMesh._populate_face_centroids (hexmesh.py:68-74).
 * Any of the four outputs may be NULL. */
int mimpy_b200_geom_hex_faces(mimpy_ctx* ctx, int32_t nf, const int32_t* face_points /*nf*4*/,
                              const double* points /*np*3*/, double* face_areas /*nf*/,
                              double* face_normals /*nf*3*/, double* face_centroids /*nf*3*/,
                              double* face_geom /*nf*8*/);

/* pack separately stored face arrays (any polygonal mesh) into face_geom */
int mimpy_b200_face_pack(mimpy_ctx* ctx, int32_t nf, const double* face_areas,
                         const double* face_normals, const double* face_centroids,
                         double* face_geom);

/* -- (2) G6: Mirtich cell volumes and centroids ---------------------------------------------
 * replaces mesh_cython.all_cell_volumes_centroids (mesh_cython.pyx:5-207), caller
 * Mesh.find_volume_centroid_all (mesh.py:1758-1781).  Gather form: one lane per (cell, face),
 * no scatter, outputs need not be pre-zeroed.  face_ptr: nf+1 offsets into face_points. */
int mimpy_b200_geom_cells(mimpy_ctx* ctx, int32_t nc, int32_t max_faces, const int32_t* cell_ptr,
                          const int32_t* cell_faces, const int8_t* cell_sigma,
                          const int32_t* face_ptr, const int32_t* face_points,
                          const double* points, const double* face_normals /*nf*3*/,
                          double* cell_volume /*nc*/, double* cell_centroid /*nc*3*/);

/* -- (3) L1-L4: batched local inner-product matrices ------------------------------------------
 * replaces MFD.build_r_e / build_n_e / build_c_e / build_m_e (mfd.py:248-483), methods 0..7
 * (MFD.set_m_e_construction_method, mfd.py:131-139).  m_ptr[c] = offset of the row-major
 * n_E x n_E block of cell c in m_e (m_ptr[nc] = sum n_E^2).
 * flags: bit0 k_unity (mfd.py:300-304, 367-369).
 * face_real_centroids / cell_real_centroids: only read by methods 2 and 6 when the mesh view
 * carries SHIFTED centroids (mfd.py:398-399, 451-452 always use the real ones); may be NULL.
 * diagonality (nc, nullable): ||M_E - diag||_F / ||M_E||_F          (mfd.py:476-478)
 * consistency (nc, nullable): ||M_E N_E - R_E||_F                    (mfd.py:472-474)      */
int mimpy_b200_mfd_local_matrices(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, int32_t method,
                                  int32_t flags, const int32_t* m_ptr, double* m_e,
                                  double* diagonality, double* consistency);

/* -- (4) A1-A5: symbolic assembly ------------------------------------------------------------
 * Everything index-valued that MFD.renumber_for_neumann (mfd.py:141-159), build_m (545-599),
 * build_div (203-246), build_bottom_right (745-774) and coo_matrix(...).tocsr() (906-908)
 * decide, computed once per (mesh, Neumann set):
 *
 *   face_to_flux[nf+1] g(f) = f - #Neumann faces before f, -1 on Neumann faces; [nf] = flux_dof
 *   face_cell[nf*2]    the <=2 cells that list face f, lower cell index first, -1 = none
 *                      (derived from the cell lists; Mesh.face_to_cell is not reliable after
 *                      divide_cell_by_plane, mesh.py:2405-2410)
 *   indptr/indices     sorted CSR pattern of the saddle matrix [M -DIV^T; DIV C],
 *                      size (flux_dof+nc)^2; M keeps its full STRUCTURAL pattern (the
 *                      reference's 1e-20 value filter, mfd.py:582, only removes exact zeros)
 *   m_indptr/m_indices CSR pattern of the M block alone (= pattern of the hybridised system)
 *   m_ptr[nc+1]        offsets of the per-cell n_E^2 blocks
 *   pos_m[sum n_E^2]   uint8: offset of entry (i,j) of cell c inside row g(f_i) (0xFF = dropped:
 *                      i or j Neumann); same offset in the saddle CSR and in the M-only CSR
 *   pos_dt[sum n_E]    uint8: offset of the -DIV^T entry (g(f_i), flux_dof+c) inside its row
 *   pos_d[sum n_E]     uint8: offset of the DIV entry (flux_dof+c, g(f_i)) inside its row
 *   diag_pos[nf]       uint8: offset of the diagonal entry inside flux row g (first flux_dof used)
 *
 * Two calls because sizes are data dependent: _sizes fills face_to_flux, face_cell, m_ptr,
 * indptr, m_indptr and returns flux_dof / nnz / nnz_m (synchronises); the caller allocates
 * indices, m_indices; _fill writes them and the pos_* maps.                               */
typedef struct {
  int32_t flux_dof;
  int64_t nnz;      /* saddle matrix                */
  int64_t nnz_m;    /* M block / hybridised system  */
  int64_t n_blocks; /* sum n_E^2                    */
  int32_t* face_to_flux;
  int32_t* face_cell;
  int32_t* indptr;    /* nf+nc+1 allocated, flux_dof+nc+1 used */
  int32_t* indices;   /* nnz                                   */
  int32_t* m_indptr;  /* nf+1 allocated, flux_dof+1 used       */
  int32_t* m_indices; /* nnz_m                                 */
  int32_t* m_ptr;     /* nc+1                                  */
  uint8_t* pos_m;     /* n_blocks                              */
  uint8_t* pos_dt;    /* cell_ptr[nc]                          */
  uint8_t* pos_d;     /* cell_ptr[nc]                          */
  uint8_t* diag_pos;  /* nf                                    */
  uint8_t* role;      /* cell_ptr[nc]: 0 = the face has this cell only, 1 = this is the lower-index
                         cell of a shared face, 2 = the higher-index one (finishes the diagonal)   */
  int32_t* cell_rowstart; /* cell_ptr[nc]: indptr[g(f_i)] per (cell, face), -1 on Neumann faces --
                         saves the numeric kernel the dependent gathers f -> g(f) -> indptr[g]    */
  const uint8_t* face_flag; /* nf or NULL (input, set by the caller on a slab-partitioned mesh):
                         1 = interface face whose other cell lives on the upper neighbour rank
                         (this rank holds the lower-index cell), 2 = ... on the lower neighbour
                         rank.  Interface faces are interior faces of the global problem, not
                         boundary faces, for the hybridised solve.                                */
  int64_t serial;     /* output of _fill: process-unique id of this plan (> 0).  The pipelined brick
                         kernel keeps the plan's row offsets inside the face records ("bind" word below)
                         and re-binds when a different plan is used with the same records.            */
} mimpy_symbolic;

int mimpy_b200_mfd_symbolic_sizes(mimpy_ctx* ctx, const mimpy_mesh_view* mesh,
                                  const uint8_t* neumann_mask /*nf, 1 = Neumann*/,
                                  mimpy_symbolic* sym);
int mimpy_b200_mfd_symbolic_fill(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, mimpy_symbolic* sym);

/* -- (5) A2-A5 + A8: numeric assembly (fused with the M_E construction) ------------------------
 * vals[nnz] of the saddle CSR in one pass over the cells: per cell the n_E^2 block is built
 * in registers/shared memory and scattered straight into its CSR slots
 *   M[g(i),g(j)] += s_i s_j mult[c] M_E[i,j]      (mfd.py:575-588; mult: update_m,
 *                                                  mfd.py:729-743 / mfd_cython.pyx:6-22)
 *   A[F+c, g(i)] = s_i a_i ; A[g(i), F+c] = -s_i a_i            (mfd.py:232-243)
 *   A[F+c, F+c]  = c_diag ? c_diag[c] : alpha * vol(c)          (mfd.py:766-772, twophase.py:783)
 * Every slot has <= 2 contributors (the two cells of a face), so the sum is order-free and
 * bit-reproducible.  multipliers / c_diag / m_e_out may be NULL.  m_e_out receives the
 * UNSCALED blocks (what the reference keeps in m_data_for_update, mfd.py:593-595).       */
int mimpy_b200_mfd_numeric(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym,
                           int32_t method, int32_t flags, const double* multipliers, double alpha,
                           const double* c_diag, double* vals, double* m_e_out);

/* same scatter from already built blocks (m_e, layout of (3)); used when M_E comes from the
 * caller (e.g. mesh-dependent user construction) */
/* mimpy_b200_mfd_numeric restricted to the cell layers [layer_begin, layer_end) of a structured hex
 * mesh (both multiples of 4, or layer_end == hex_nz), called for increasing ranges that tile the mesh:
 * after the call for a range, the CSR rows of the faces k*L .. (in HexMesh numbering, L faces per
 * layer) of those layers and of their cells are final, so a caller can overlap the download of one
 * slab with the upload of cell_k for the next (bench.py e2e).  MIMPY_E_UNSUPPORTED when the staged
 * brick kernel does not apply. */
int mimpy_b200_mfd_numeric_layers(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym,
                                  int32_t method, int32_t flags, const double* multipliers, double alpha,
                                  const double* c_diag, double* vals, int32_t layer_begin, int32_t layer_end);
int mimpy_b200_mfd_numeric_from_blocks(mimpy_ctx* ctx, const mimpy_mesh_view* mesh,
                                       const mimpy_symbolic* sym, const double* m_e,
                                       const double* multipliers, double alpha,
                                       const double* c_diag, double* vals);

/* COO view of the M block in the reference's own order (cell-major, row-major inside the cell,
 * Neumann rows/cols skipped; mfd.py:575-588): coo_ptr[nc+1] = m_e_locations (mfd.py:590-595),
 * and optionally rows/cols/vals of each triple.  Any output may be NULL. */
int mimpy_b200_mfd_coo(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym,
                       const double* m_e, int64_t* coo_ptr, int32_t* rows, int32_t* cols,
                       double* vals);

/* -- (6) A6: right-hand side -------------------------------------------------------------------
 * replaces MFD.build_rhs (mfd.py:991-1041, 1209-1215): rhs = 0; Neumann rhs[g(f)] = value with
 * g = -1, i.e. rhs[last] = value of the LAST Neumann face in insertion order (reference quirk,
 * mfd.py:1041 with mfd.py:153); Dirichlet rhs[g(f)] = -value; forcing rhs[F+c] += F_c.     */
int mimpy_b200_mfd_rhs(mimpy_ctx* ctx, int32_t nc, int32_t flux_dof, const int32_t* face_to_flux,
                       int32_t n_dirichlet, const int32_t* dirichlet_faces,
                       const double* dirichlet_values, int32_t has_neumann,
                       double last_neumann_value, const double* forcing /*nc, nullable*/,
                       double* rhs /*flux_dof+nc*/);

/* P1: per-step terms of SinglePhase.start_solving (singlephase.py:263-281):
 * multipliers[c] = mu / (((p - p_ref) c_f + 1) rho_ref)   (update_m factor, inverse of rho/mu)
 * c_diag[c]      = c_f phi / dt |E|                        (accumulation diagonal)
 * rhs[flux_dof+c] += c_diag[c] * p[c] */
int mimpy_b200_singlephase_terms(mimpy_ctx* ctx, int32_t nc, int32_t flux_dof, double ref_pressure,
                                 double compressibility, double ref_density, double viscosity,
                                 double dt, const double* pressure, const double* porosity,
                                 const double* cell_volume, double* rhs, double* multipliers,
                                 double* c_diag);

/* Transport.update_concentration + find_upwinding_direction (transport.py:303-361): explicit upwind
 * step of a concentration carried by the flux field `velocity` (flux-DOF indexed, as returned by the
 * solve).  boundary_conc: NULL or one value per FACE, NaN = none (transport.py:308-317 applies it on
 * inflow faces of the markers given to set_concentration_boundaries).  conc_new must not alias
 * conc_old.  The result is not divided by the cell volume (transport.py:344 is commented out). */
int mimpy_b200_transport_step(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym,
                              const double* velocity, const double* conc_old,
                              const double* boundary_conc, const double* p_prev, const double* p_cur,
                              const double* porosity, double ref_pressure, double compressibility,
                              double ref_density, double dt, double* conc_new);

/* -- (7) S1: linear solve -----------------------------------------------------------------------
 * replaces MFD.solve('spsolve') (mfd.py:1349-1358) for the saddle system
 *     M v - DIV^T p = b_v ,  DIV v + C p = b_p
 * by exact algebraic hybridisation (face multipliers lambda, one per flux DOF) followed by
 * Jacobi-preconditioned CG on the SPD face system  A lambda = h  (pattern = m_indptr/m_indices)
 * and cell-local recovery of [v ; p] in the reference's DOF order.
 *
 * hybrid_setup : per cell W_E = (masked, scaled M_E)^-1, A_E = a W a - r c^T / S scattered into
 *                a_vals[nnz_m]; boundary (single-cell) flux DOFs become identity rows/cols;
 *                w_e (n_blocks) keeps W_E for rhs/recovery; a_diag_inv[flux_dof] = 1/diag(A).
 * hybrid_rhs   : h[flux_dof] for given saddle rhs (b = [b_v ; b_p]); boundary rows carry the
 *                prescribed multiplier.
 * pcg          : solves A x = h; x holds the initial guess on entry.
 * hybrid_recover: solution[flux_dof+nc] = [v ; p] from lambda.                              */
int mimpy_b200_hybrid_setup(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym,
                            const double* m_e, const double* multipliers, double alpha,
                            const double* c_diag, double* w_e, double* a_vals, double* a_diag_inv,
                            const int32_t* sell_ptr /* NULL: a_vals in CSR order (m_indptr);
                            else a_vals in the SELL-32 layout of mimpy_b200_sell_sizes/_fill */);
/* Same as mimpy_b200_mfd_local_matrices + mimpy_b200_hybrid_setup in one pass for hexahedral cells
 * (uniform_faces == 6, methods 0/1/6): M_E is built and inverted per cell without ever being stored
 * (mfd.py:357-483 feeding the elimination above).  Returns MIMPY_E_UNSUPPORTED otherwise -- callers
 * then use the two-step path.  flags: MIMPY_FLAG_K_UNITY.
 * face_rowbase_work (nf ints, nullable): scratch that enables the pipelined brick kernel on structured hex
 * meshes with even nx (bulk-copy staging, closed-form row positions; csrc/assemble_brick3.cu).  That kernel
 * also writes tau_out[nc*6] (nullable) = a_f^2 W_E[f,f], the half transmissibilities mimpy_b200_auxmg_setup
 * needs, and sets *tau_written_host (nullable) to 1 when it did. */
int mimpy_b200_hybrid_setup_fused(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym,
                                  int32_t method, int32_t flags, const double* multipliers, double alpha,
                                  const double* c_diag, double* w_e, double* a_vals, double* a_diag_inv,
                                  const int32_t* sell_ptr, int32_t* face_rowbase_work, double* tau_out,
                                  int32_t* tau_written_host);
int mimpy_b200_hybrid_rhs(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym,
                          const double* w_e, double alpha, const double* c_diag,
                          const double* rhs /*flux_dof+nc*/, double* cell_tmp /*cell_ptr[nc]*/,
                          double* h /*flux_dof*/);
int mimpy_b200_hybrid_recover(mimpy_ctx* ctx, const mimpy_mesh_view* mesh,
                              const mimpy_symbolic* sym, const double* w_e, double alpha,
                              const double* c_diag, const double* rhs, const double* lambda,
                              double* solution /*flux_dof+nc*/);

typedef struct {
  double rtol;          /* stop when ||r||_2 <= rtol * ||h||_2                        */
  double atol;          /* ... or ||r||_2 <= atol                                      */
  int32_t maxit;
  int32_t check_every;  /* host looks at the residual every this many iterations       */
  int32_t use_graph;    /* 1: replay each block of check_every iterations as a graph   */
  /* outputs */
  int32_t iterations;
  double residual;      /* final ||r||_2 (recurrence)                                  */
  double rhs_norm;
  float ms_total;       /* device time of the whole solve (CUDA events)               */
  /* inputs (zero-initialised = the behaviour above) */
  int32_t norm_kind;    /* 0: ||r||_2 against rtol * (ref_norm > 0 ? ref_norm : ||h||_2);
                           1: preconditioned residual, sqrt(r.Br) <= rtol * sqrt(r0.Br0) (r0 = the
                              residual of the initial guess) -- an estimate of the energy-norm error
                              that does not depend on how the rows of h are scaled              */
  double ref_norm;      /* norm_kind 0: reference norm replacing ||h||_2, e.g. the norm of the saddle
                           right-hand side the face system was condensed from                   */
  /* outputs */
  double rz0, rz;       /* r.Br of the initial guess and at the end                            */
  /* inputs, SELL layout only (NULL / 0 = columns are read from sell_cols): the column patterns of
     mimpy_b200_sell_patterns -- the SpMV then reads one byte per ROW instead of four per entry      */
  const uint8_t* sell_pat_id;
  const int32_t* sell_pat_table;
  int32_t sell_npat;
} mimpy_pcg_info;

/* work: 4*n doubles (r, p, Ap, spare) */
int mimpy_b200_pcg(mimpy_ctx* ctx, int32_t n, const int32_t* indptr, const int32_t* indices,
                   const double* vals, const double* diag_inv, const double* b, double* x,
                   double* work, mimpy_pcg_info* info);

/* y = A x for a CSR matrix with int32 indices (the solver's SpMV, exposed for tests/bench) */
/* SELL-32 layout of a CSR pattern (the solver's preferred layout for the face system): 32-row
 * slices stored column-major, padded to the longest row of the slice; entry k of row g lives at
 * sell_ptr[g>>5] + 32*k + (g&31); padding = value 0, column g.  sell_ptr: (n+31)/32+1 ints.
 * sell_sizes returns the padded entry count (synchronises); sell_fill writes the columns.
 * pcg_sell / spmv_sell are pcg / spmv on that layout. */
int mimpy_b200_sell_sizes(mimpy_ctx* ctx, int32_t n, const int32_t* indptr, int32_t* sell_ptr,
                          int64_t* total_host);
/* Column patterns of a SELL-32 matrix: pat_id[32 * slices] (one byte per row), pat_table[255 *
 * MIMPY_SELL_PAT_W] offset vectors (column = row + offset) of the 254 most frequent patterns; id 255 =
 * escape, the row reads sell_cols.  *count_host = number of table patterns, 0 when more than a quarter of
 * the rows would escape.  Every row is verified against its pattern on the device.  Synchronises. */
#define MIMPY_SELL_PAT_W 12
int mimpy_b200_sell_patterns(mimpy_ctx* ctx, int32_t n, const int32_t* sell_ptr, const int32_t* sell_cols,
                             uint8_t* pat_id, int32_t* pat_table, int32_t* count_host);
int mimpy_b200_spmv_sell_pat(mimpy_ctx* ctx, int32_t n, const int32_t* sell_ptr, const int32_t* sell_cols,
                             const uint8_t* pat_id, const int32_t* pat_table, int32_t npat, const double* sell_vals,
                             const double* x, double* y);
int mimpy_b200_sell_fill(mimpy_ctx* ctx, int32_t n, const int32_t* indptr, const int32_t* indices,
                         const int32_t* sell_ptr, int32_t* sell_cols);
int mimpy_b200_pcg_sell(mimpy_ctx* ctx, int32_t n, const int32_t* sell_ptr, const int32_t* sell_cols,
                        const double* sell_vals, const double* diag_inv, const double* b, double* x,
                        double* work, mimpy_pcg_info* info);

/* Auxiliary-space multigrid preconditioner for the face system on structured hex meshes
 * (csrc/auxmg.cu): B = D^-1 + Pi V(A_c) Pi^T with A_c the cell-centred two-point-flux operator built
 * from the inverse local matrices w_e of hybrid_setup.  No reference counterpart: MFD.solve
 * (mfd.py:1349-1358) factorises; the reference's unused Schur CG preconditions with the same kind of
 * operator (twophase.py:665-686).  The handle owns its device memory; the mesh / symbolic arrays it
 * was given must outlive it.  auxmg_setup must follow every hybrid_setup (new w_e). */
int mimpy_b200_auxmg_create(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym,
                            void** handle);
int mimpy_b200_auxmg_setup(mimpy_ctx* ctx, void* handle, const mimpy_mesh_view* mesh,
                           const mimpy_symbolic* sym, const double* w_e, double alpha,
                           const double* c_diag, const double* tau /* nc*6 from hybrid_setup_fused, or NULL:
                           read from the diagonals of w_e */);
int mimpy_b200_auxmg_destroy(mimpy_ctx* ctx, void* handle);
/* Cell-block smoother of that preconditioner: after auxmg_setup, hand the assembled face system (CSR: ptr = row
 * offsets, sell = 0; SELL-32: ptr = slice offsets, sell = 1) and its inverted (complete) diagonal; the face-space
 * term D^-1 of B becomes the weighted additive Schwarz sum over the cells of the inverses of the 6 x 6 blocks of A
 * on the faces of each cell (stored as 21 floats per cell).  No reference counterpart.  MIMPY_B200_AUX_BLOCKS=0 in
 * the environment of auxmg_create keeps the Jacobi term (the call is then a no-op). */
/* Number of cells (summed over the ranks of a partition) with tr(K)^3 / (27 det K) > threshold: how anisotropic the
 * permeability field is; decides between the two smoothers (host side: device.HybridSystem(smoother="auto")). */
int mimpy_b200_k_anisotropy(mimpy_ctx* ctx, int32_t nc, const double* cell_k /*nc*9*/, double threshold,
                            double* count_out /*host*/);
int mimpy_b200_auxmg_cellblocks(mimpy_ctx* ctx, void* auxmg, const mimpy_mesh_view* mesh,
                                const mimpy_symbolic* sym, const int32_t* ptr, const int32_t* idx,
                                const double* vals, int32_t sell, const double* diag_inv);
/* PCG with that preconditioner.  (ptr, idx, vals): the face system in CSR (sell = 0) or SELL-32
 * (sell = 1) layout; work: 4 n doubles.  Same info / error behaviour as mimpy_b200_pcg. */
int mimpy_b200_pcg_aux(mimpy_ctx* ctx, int32_t n, const int32_t* ptr, const int32_t* idx,
                       const double* vals, int32_t sell, const double* diag_inv, void* auxmg,
                       const double* b, double* x, double* work, mimpy_pcg_info* info);
int mimpy_b200_spmv_sell(mimpy_ctx* ctx, int32_t n, const int32_t* sell_ptr, const int32_t* sell_cols,
                         const double* sell_vals, const double* x, double* y);
int mimpy_b200_spmv(mimpy_ctx* ctx, int32_t n, const int32_t* indptr, const int32_t* indices,
                    const double* vals, const double* x, double* y);

/* -- (8)(9) T2, T4, T5: two-phase saturation transport -----------------------------------------
 * upwind: upwind_cell[g] = the cell c with sigma_{c,f} * u_t[g(f)] > 0, -1 if none, INCLUDING
 * the reference quirk that Neumann faces test u_t[-1] and may overwrite the entry of the last
 * flux DOF (twophase.py:928-958).
 * saturation_step: f_w (persistent, twophase.py:957, 968-971) refreshed on the upwinded DOFs,
 * boundary inflow DOFs set from bnd_fw where sigma_b*u_t < 0 (twophase.py:974-994), then
 *   s_w <- (1+c_w p_prev)/(1+c_w p) s_w - dt div(f_w u_t)/(phi rho_w (1+c_w p) vol) + wells
 * (twophase.py:996-1092).  Corey relperm with clipping (twophase.py:268-344).              */
typedef struct {
  double mu_w, mu_o, rho_w, rho_o, c_w, c_o, s_wr, s_or, n_w, n_o, k_rw_o;
} mimpy_fluid;

int mimpy_b200_twophase_mobility(mimpy_ctx* ctx, int32_t nc, const mimpy_fluid* fluid_host,
                                 const double* s_w, const double* p, double* inv_total_mobility,
                                 double* water_mob, double* oil_mob);
/* Newton set-up of TwoPhase.update_pressure (twophase.py:764-816): c_diag[c] = accumulation
 * diagonal (rho_w s c_w + rho_o c_o (1-s)) phi |E| / dt, and rhs[flux_dof+c] += f1 + f2 - f3 (the
 * phase masses at time n minus their linearisation at pressure p_n).  rhs / c_diag may be NULL. */
int mimpy_b200_twophase_newton_terms(mimpy_ctx* ctx, int32_t nc, int32_t flux_dof,
                                     const mimpy_fluid* fluid_host, double dt, const double* s_w,
                                     const double* p_n, const double* porosity,
                                     const double* cell_volume, double* rhs, double* c_diag);
int mimpy_b200_twophase_upwind(mimpy_ctx* ctx, const mimpy_mesh_view* mesh,
                               const mimpy_symbolic* sym, const double* u_t,
                               int32_t* upwind_cell /*flux_dof*/);
int mimpy_b200_twophase_saturation_step(mimpy_ctx* ctx, const mimpy_mesh_view* mesh,
                                        const mimpy_symbolic* sym, const mimpy_fluid* fluid_host,
                                        double dt, const int32_t* upwind_cell, const double* u_t,
                                        const double* p, const double* p_prev,
                                        const double* porosity, int32_t n_bnd,
                                        const int32_t* bnd_faces, const int8_t* bnd_sigma,
                                        const double* bnd_fw, int32_t n_rate_wells,
                                        const int32_t* well_cells, const double* well_rate_water,
                                        double* f_w /*flux_dof, in/out*/, double* s_w /*nc, in/out*/,
                                        const double* water_mob /*nc or NULL*/,
                                        const double* oil_mob /*nc or NULL*/);
/* water_mob / oil_mob: per-cell phase mobilities evaluated by the caller from the saturation BEFORE this
 * step -- the path of user relative-permeability callables (TwoPhase.set_krw / set_kro,
 * twophase.py:292-300); NULL = Corey curves of `fluid_host` evaluated in the kernel. */
/* The same step with flux_area[g] = area of the face of flux DOF g (8 B per face instead of the 32-B sector of
 * the packed record; filled once per (mesh, plan) by mimpy_b200_flux_areas; NULL = read the records). */
int mimpy_b200_twophase_saturation_step2(mimpy_ctx* ctx, const mimpy_mesh_view* mesh,
                                         const mimpy_symbolic* sym, const mimpy_fluid* fluid_host,
                                         double dt, const int32_t* upwind_cell, const double* u_t,
                                         const double* p, const double* p_prev,
                                         const double* porosity, int32_t n_bnd,
                                         const int32_t* bnd_faces, const int8_t* bnd_sigma,
                                         const double* bnd_fw, int32_t n_rate_wells,
                                         const int32_t* well_cells, const double* well_rate_water,
                                         double* f_w, double* s_w, const double* water_mob,
                                         const double* oil_mob, const double* flux_area /*flux_dof or NULL*/);
int mimpy_b200_flux_areas(mimpy_ctx* ctx, const mimpy_mesh_view* mesh, const mimpy_symbolic* sym,
                          double* flux_area /*flux_dof*/);

/* -- (10) multi-GPU: z-slab partition of the cell set (host side in
 * mimpy_b200/partition.py) ----------------------------------------------------------------------
 * Each rank assembles and solves on its own slab (a standalone local mesh); the face system is
 * SUB-ASSEMBLED: interface rows hold this rank's half until hybrid_setup / hybrid_rhs / every SpMV
 * of pcg exchange-add the two halves with the neighbour (ncclSend/ncclRecv in one group), and the
 * dot products are all-reduced (ncclAllReduce, FP64 sum).  No other collective is used.
 *
 * NCCL is resolved at run time from the library already resident in the process (PyTorch's), see
 * comm.cu.  nccl_unique_id: rank 0 creates the 128-byte id, the caller broadcasts it (e.g. with
 * torch.distributed) and every rank calls nccl_init.
 * set_halo: n_owned = number of leading face-system rows counted in dot products (ghost copies of
 * the upper interface layer come last in the local numbering); idx_lo / idx_hi = positions of the
 * rows shared with the lower / upper neighbour; buf = 2*(n_lo+n_hi) doubles. */
int mimpy_b200_set_comm(mimpy_ctx* ctx, void* nccl_comm, int32_t rank, int32_t world);
int mimpy_b200_nccl_load(const char* path_host);
int mimpy_b200_nccl_unique_id(void* out128_host);
int mimpy_b200_nccl_init(mimpy_ctx* ctx, const void* id128_host, int32_t rank, int32_t world);
int mimpy_b200_set_halo(mimpy_ctx* ctx, int32_t n_owned, int32_t n_lo, const int32_t* idx_lo,
                        int32_t rank_lo, int32_t n_hi, const int32_t* idx_hi, int32_t rank_hi,
                        double* buf);
int mimpy_b200_halo_exchange_add(mimpy_ctx* ctx, double* v);
/* NVLink peer-memory path for the same two collectives (latency, not bandwidth, is what they
 * cost): peer_alloc creates this rank's symmetric buffer (halo_cap_doubles >= nx*ny) and returns
 * its 64-byte cudaIpc handle; the caller all-gathers the handles of the `world` (<= 8) ranks of
 * the node and calls peer_attach.  From then on halo_exchange_add / allreduce_sum / pcg push their
 * data straight into the peers' buffers (st.global over NVLink + system fence + epoch flag)
 * instead of calling NCCL.  peer_disable switches back to NCCL. */
int mimpy_b200_peer_alloc(mimpy_ctx* ctx, int64_t halo_cap_doubles, void* handle64_host);
int mimpy_b200_peer_attach(mimpy_ctx* ctx, int32_t rank, int32_t world, const void* handles_host);
int mimpy_b200_peer_disable(mimpy_ctx* ctx);
int mimpy_b200_allreduce_sum(mimpy_ctx* ctx, double* buf, int32_t count);

#ifdef __cplusplus
}
#endif
#endif /* MIMPY_B200_H */
