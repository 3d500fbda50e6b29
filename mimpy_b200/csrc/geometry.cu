// Mesh geometry kernels (compiled with -fmad=false so that the arithmetic is the reference's
// own sequence of IEEE operations: the outputs are bit-identical to the Cython natives).
//
//   k_hex_faces : G3-G5  hexmesh_cython.pyx:30-118, hexmesh.py:26-42
//   k_face_pack : gathers three face arrays into the packed 64-byte record
//   k_cells     : G6     mesh_cython.pyx:5-207 in gather form (one lane per (cell, face))
#include "common.cuh"

__device__ __forceinline__ double dist3(const double* a, const double* b) {
  double s = (a[0] - b[0]) * (a[0] - b[0]);
  s += (a[1] - b[1]) * (a[1] - b[1]);
  s += (a[2] - b[2]) * (a[2] - b[2]);
  return sqrt(s);
}

__device__ __forceinline__ double heron(const double* p, const double* q, const double* c) {
  const double a = dist3(p, q);
  const double b = dist3(q, c);
  const double cc = dist3(c, p);
  const double s = (a + b + cc) / 2.;
  return sqrt(s * (s - a) * (s - b) * (s - cc));
}

__global__ void __launch_bounds__(256)
k_hex_faces(int nf, const int4* __restrict__ face_points, const double* __restrict__ points,
            double* __restrict__ areas, double* __restrict__ normals,
            double* __restrict__ centroids, double* __restrict__ geom) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  const int4 id = face_points[f];
  double p[4][3];
  const int ids[4] = {id.x, id.y, id.z, id.w};
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const double* src = points + 3 * (size_t)ids[v];
    p[v][0] = __ldg(src);
    p[v][1] = __ldg(src + 1);
    p[v][2] = __ldg(src + 2);
  }
  double c[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) c[d] = .25 * (p[0][d] + p[1][d] + p[2][d] + p[3][d]);
  // hexmesh_cython.pyx:56-88 -- four triangles (p_i, p_{i+1}, centre)
  double area = 0;
  area += heron(p[0], p[1], c);
  area += heron(p[1], p[2], c);
  area += heron(p[2], p[3], c);
  area += heron(p[3], p[0], c);
  // hexmesh_cython.pyx:103-117 -- (p1-p2) x (p2-p3), normalised
  double v1[3], v2[3], n[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    v2[d] = p[1][d] - p[2][d];
    v1[d] = p[0][d] - p[1][d];
  }
  n[0] = v1[1] * v2[2] - v1[2] * v2[1];
  n[1] = v1[2] * v2[0] - v1[0] * v2[2];
  n[2] = v1[0] * v2[1] - v1[1] * v2[0];
  double nn = n[0] * n[0];
  nn += n[1] * n[1];
  nn += n[2] * n[2];
  nn = sqrt(nn);
  n[0] /= nn;
  n[1] /= nn;
  n[2] /= nn;
  if (areas) areas[f] = area;
  if (normals) {
    normals[3 * (size_t)f + 0] = n[0];
    normals[3 * (size_t)f + 1] = n[1];
    normals[3 * (size_t)f + 2] = n[2];
  }
  if (centroids) {
    centroids[3 * (size_t)f + 0] = c[0];
    centroids[3 * (size_t)f + 1] = c[1];
    centroids[3 * (size_t)f + 2] = c[2];
  }
  if (geom) rec_store(geom, (size_t)f, area, n, c);
}

__global__ void __launch_bounds__(256)
k_face_pack(int nf, const double* __restrict__ areas, const double* __restrict__ normals,
            const double* __restrict__ centroids, double* __restrict__ geom) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  rec_store(geom, (size_t)f, areas[f], normals + 3 * (size_t)f, centroids + 3 * (size_t)f);
}

// One lane per (cell, face): the face's projection/face integrals (mesh_cython.pyx:62-150),
// then the cell sums them in ascending global face order -- the order in which the reference's
// face loop scatters into a cell (mesh_cython.pyx:152-202) -- and divides (204-207).
template <int G>
__global__ void __launch_bounds__(128)
k_cells(int nc, const int* __restrict__ cell_ptr, const int* __restrict__ cell_faces,
        const int8_t* __restrict__ cell_sigma, const int* __restrict__ face_ptr,
        const int* __restrict__ face_points, const double* __restrict__ points,
        const double* __restrict__ normals, double* __restrict__ cell_volume,
        double* __restrict__ cell_centroid) {
  constexpr int CPB = 128 / G;
  __shared__ double s_val[CPB][G][4];
  __shared__ int s_id[CPB][G];
  const int slot = threadIdx.x / G;
  const int j = threadIdx.x % G;
  const int cell = blockIdx.x * CPB + slot;
  const bool cell_ok = cell < nc;
  const int base = cell_ok ? cell_ptr[cell] : 0;
  const int n = cell_ok ? cell_ptr[cell + 1] - base : 0;
  const bool valid = j < n;
  double cv = 0, cc[3] = {0, 0, 0};
  int fid = 0x7fffffff;
  if (valid) {
    fid = cell_faces[base + j];
    const double orient = (double)cell_sigma[base + j];
    const double nrm[3] = {normals[3 * (size_t)fid], normals[3 * (size_t)fid + 1],
                           normals[3 * (size_t)fid + 2]};
    int C;
    if (fabs(nrm[0]) > fabs(nrm[1]) && fabs(nrm[0]) > fabs(nrm[2]))
      C = 0;
    else if (fabs(nrm[1]) > fabs(nrm[2]))
      C = 1;
    else
      C = 2;
    const int A = (C + 1) % 3;
    const int B = (A + 1) % 3;
    double P1 = 0., Pa = 0., Pb = 0., Paa = 0., Pab = 0., Pbb = 0.;
    const int p0 = face_ptr[fid];
    const int len = face_ptr[fid + 1] - p0;
    for (int l = 0; l < len; ++l) {
      const double* cur = points + 3 * (size_t)face_points[p0 + l];
      const double* nxt = points + 3 * (size_t)face_points[p0 + ((l + 1) % len)];
      const double a0 = cur[A], b0 = cur[B], a1 = nxt[A], b1 = nxt[B];
      const double da = a1 - a0, db = b1 - b0;
      const double a0_2 = a0 * a0, a0_3 = a0_2 * a0;
      const double b0_2 = b0 * b0, b0_3 = b0_2 * b0;
      const double a1_2 = a1 * a1;
      const double C1 = a1 + a0;
      const double Ca = a1 * C1 + a0_2;
      const double Caa = a1 * Ca + a0_3;
      const double Cb = b1 * (b1 + b0) + b0_2;
      const double Cbb = b1 * Cb + b0_3;
      const double Cab = 3. * a1_2 + 2. * a1 * a0 + a0_2;
      const double Kab = a1_2 + 2 * a1 * a0 + 3 * a0_2;
      P1 += db * C1;
      Pa += db * Ca;
      Paa += db * Caa;
      Pb += da * Cb;
      Pbb += da * Cbb;
      Pab += db * (b1 * Cab + b0 * Kab);
    }
    P1 /= 2.0;
    Pa /= 6.0;
    Paa /= 12.0;
    Pb /= -6.0;
    Pbb /= -12.0;
    Pab /= 24.0;
    const double* fp = points + 3 * (size_t)face_points[p0];
    double w = 0.;
    w -= nrm[0] * fp[0];
    w -= nrm[1] * fp[1];
    w -= nrm[2] * fp[2];
    const double k1 = 1. / nrm[C];
    const double k2 = k1 * k1;
    const double k3 = k2 * k1;
    const double Fa = k1 * Pa;
    const double Fb = k1 * Pb;
    const double Fc = -k2 * (nrm[A] * Pa + nrm[B] * Pb + w * P1);
    const double Faa = k1 * Paa;
    const double Fbb = k1 * Pbb;
    const double Fcc = k3 * ((nrm[A] * nrm[A]) * Paa + 2 * nrm[A] * nrm[B] * Pab +
                             (nrm[B] * nrm[B]) * Pbb +
                             w * (2. * (nrm[A] * Pa + nrm[B] * Pb) + w * P1));
    if (A == 0)
      cv = nrm[0] * Fa * orient;
    else if (B == 0)
      cv = nrm[0] * Fb * orient;
    else
      cv = nrm[0] * Fc * orient;
    cc[A] = nrm[A] * Faa * orient;
    cc[B] = nrm[B] * Fbb * orient;
    cc[C] = nrm[C] * Fcc * orient;
  }
  s_id[slot][j] = fid;
  __syncwarp();
  // rank of this face among the cell's faces by global id (ascending)
  int rank = 0;
  for (int i = 0; i < n; ++i) rank += (s_id[slot][i] < fid) || (s_id[slot][i] == fid && i < j);
  if (valid) {
    s_val[slot][rank][0] = cv;
    s_val[slot][rank][1] = cc[0];
    s_val[slot][rank][2] = cc[1];
    s_val[slot][rank][3] = cc[2];
  }
  __syncwarp();
  if (j == 0 && cell_ok) {
    double v = 0., x = 0., y = 0., z = 0.;
    for (int i = 0; i < n; ++i) {
      v += s_val[slot][i][0];
      x += s_val[slot][i][1];
      y += s_val[slot][i][2];
      z += s_val[slot][i][3];
    }
    cell_volume[cell] = v;
    cell_centroid[3 * (size_t)cell + 0] = x / (v * 2.);
    cell_centroid[3 * (size_t)cell + 1] = y / (v * 2.);
    cell_centroid[3 * (size_t)cell + 2] = z / (v * 2.);
  }
}

extern "C" {

int mimpy_b200_geom_hex_faces(mimpy_ctx* ctx, int32_t nf, const int32_t* face_points,
                              const double* points, double* face_areas, double* face_normals,
                              double* face_centroids, double* face_geom) {
  REQUIRE(ctx && face_points && points, "NULL argument");
  if (nf <= 0) return MIMPY_OK;
  const int threads = 256;
  k_hex_faces<<<(nf + threads - 1) / threads, threads, 0, ctx->stream>>>(
      nf, reinterpret_cast<const int4*>(face_points), points, face_areas, face_normals,
      face_centroids, face_geom);
  KERNEL_CHECK();
  if (face_geom && ctx->bound_geom == face_geom) ctx->bound_geom = nullptr;  // the bind words were overwritten
  return MIMPY_OK;
}

int mimpy_b200_face_pack(mimpy_ctx* ctx, int32_t nf, const double* face_areas,
                         const double* face_normals, const double* face_centroids,
                         double* face_geom) {
  REQUIRE(ctx && face_areas && face_normals && face_centroids && face_geom, "NULL argument");
  if (nf <= 0) return MIMPY_OK;
  const int threads = 256;
  k_face_pack<<<(nf + threads - 1) / threads, threads, 0, ctx->stream>>>(
      nf, face_areas, face_normals, face_centroids, face_geom);
  KERNEL_CHECK();
  if (ctx->bound_geom == face_geom) ctx->bound_geom = nullptr;
  return MIMPY_OK;
}

int mimpy_b200_geom_cells(mimpy_ctx* ctx, int32_t nc, int32_t max_faces, const int32_t* cell_ptr,
                          const int32_t* cell_faces, const int8_t* cell_sigma,
                          const int32_t* face_ptr, const int32_t* face_points,
                          const double* points, const double* face_normals, double* cell_volume,
                          double* cell_centroid) {
  REQUIRE(ctx && cell_ptr && cell_faces && cell_sigma && face_ptr && face_points && points &&
              face_normals && cell_volume && cell_centroid,
          "NULL argument");
  REQUIRE(max_faces >= 1 && max_faces <= MIMPY_MAX_CELL_FACES, "max_faces out of range");
  if (nc <= 0) return MIMPY_OK;
  const int G = mimpy_group_size(max_faces);
  const int cpb = 128 / G;
  const int blocks = (nc + cpb - 1) / cpb;
#define LAUNCH(GG)                                                                          \
  k_cells<GG><<<blocks, 128, 0, ctx->stream>>>(nc, cell_ptr, cell_faces, cell_sigma,        \
                                               face_ptr, face_points, points, face_normals, \
                                               cell_volume, cell_centroid)
  if (G == 8)
    LAUNCH(8);
  else if (G == 16)
    LAUNCH(16);
  else
    LAUNCH(32);
#undef LAUNCH
  KERNEL_CHECK();
  return MIMPY_OK;
}

}  // extern "C"

// ---- G2: structured hex topology in closed form (hexmesh.py:161-234, 269-350) ----------------
// HexMesh._build_faces emits, in one k->j->i sweep over the nodes, a z-face (i<cx, j<cy), a y-face
// (k<cz, i<cx) and an x-face (j<cy, k<cz); the resulting interleaved numbering is
//   L = cx*cy + cx*(cy+1) + (cx+1)*cy faces per cell layer,  row stride 3*cx+1.
struct HexIdx {
  int cx, cy, cz;
  long long L;
  __host__ __device__ long long zf(int i, int j, int k) const {
    return k < cz ? k * L + (long long)j * (3 * cx + 1) + 3 * i : (long long)cz * L + (long long)j * cx + i;
  }
  __host__ __device__ long long yf(int i, int j, int k) const {
    return j < cy ? k * L + (long long)j * (3 * cx + 1) + 3 * i + 1 : k * L + (long long)cy * (3 * cx + 1) + i;
  }
  __host__ __device__ long long xf(int i, int j, int k) const {
    return i < cx ? k * L + (long long)j * (3 * cx + 1) + 3 * i + 2 : k * L + (long long)j * (3 * cx + 1) + 3 * cx;
  }
};

__global__ void __launch_bounds__(256)
k_hex_topology(HexIdx h, double dx, double dy, double dz, double* __restrict__ points,
               int4* __restrict__ face_points, int* __restrict__ cell_faces,
               int8_t* __restrict__ cell_sigma) {
  const int ni = h.cx + 1, nj = h.cy + 1, nk = h.cz + 1;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)ni * nj * nk) return;
  const int i = (int)(t % ni);
  const int j = (int)((t / ni) % nj);
  const int k = (int)(t / ((long long)ni * nj));
  if (points) {  // hexmesh.py:310-313
    points[3 * t + 0] = (double)i * dx;
    points[3 * t + 1] = (double)j * dy;
    points[3 * t + 2] = (double)k * dz;
  }
  auto pid = [&](int a, int b, int c) { return (int)(a + (long long)ni * b + (long long)ni * nj * c); };
  if (face_points) {
    if (i < h.cx && j < h.cy)  // hexmesh.py:178-182
      face_points[h.zf(i, j, k)] = make_int4(pid(i, j, k), pid(i + 1, j, k), pid(i + 1, j + 1, k), pid(i, j + 1, k));
    if (k < h.cz && i < h.cx)  // hexmesh.py:195-199
      face_points[h.yf(i, j, k)] = make_int4(pid(i, j, k), pid(i, j, k + 1), pid(i + 1, j, k + 1), pid(i + 1, j, k));
    if (j < h.cy && k < h.cz)  // hexmesh.py:211-215
      face_points[h.xf(i, j, k)] = make_int4(pid(i, j, k), pid(i, j + 1, k), pid(i, j + 1, k + 1), pid(i, j, k + 1));
  }
  if (cell_faces && i < h.cx && j < h.cy && k < h.cz) {  // hexmesh.py:329-342
    const long long c = i + (long long)h.cx * j + (long long)h.cx * h.cy * k;
    int* cf = cell_faces + 6 * c;
    cf[0] = (int)h.zf(i, j, k);
    cf[1] = (int)h.yf(i, j, k);
    cf[2] = (int)h.xf(i, j, k);
    cf[3] = (int)h.xf(i + 1, j, k);
    cf[4] = (int)h.yf(i, j + 1, k);
    cf[5] = (int)h.zf(i, j, k + 1);
    if (cell_sigma) {
      int8_t* cs = cell_sigma + 6 * c;
      cs[0] = -1; cs[1] = -1; cs[2] = -1; cs[3] = 1; cs[4] = 1; cs[5] = 1;
    }
  }
}

extern "C" int mimpy_b200_hex_topology(mimpy_ctx* ctx, int32_t cx, int32_t cy, int32_t cz, double dx,
                                       double dy, double dz, double* points, int32_t* face_points,
                                       int32_t* cell_faces, int8_t* cell_sigma) {
  REQUIRE(ctx, "ctx is NULL");
  REQUIRE(cx > 0 && cy > 0 && cz > 0, "cell counts must be positive");
  HexIdx h;
  h.cx = cx; h.cy = cy; h.cz = cz;
  h.L = (long long)cx * cy + (long long)cx * (cy + 1) + (long long)(cx + 1) * cy;
  const long long nf = (long long)cz * h.L + (long long)cx * cy;
  REQUIRE(nf < 0x7fffffffLL, "mesh too large for int32 face ids");
  const long long nodes = (long long)(cx + 1) * (cy + 1) * (cz + 1);
  const int T = 256;
  k_hex_topology<<<(unsigned)((nodes + T - 1) / T), T, 0, ctx->stream>>>(
      h, dx, dy, dz, points, reinterpret_cast<int4*>(face_points), cell_faces, cell_sigma);
  KERNEL_CHECK();
  return MIMPY_OK;
}
