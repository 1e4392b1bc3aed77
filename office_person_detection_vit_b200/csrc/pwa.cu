// Piecewise-affine camera -> floormap transform (the reference's shipped default, config.yaml:91): per point, find the
// Delaunay triangle that contains it, apply that triangle's 2x3 affine map; points outside the triangulation use the triangle
// with the nearest centroid.  Replaces (reference, NumPy / SciPy, one Python call per box):
//   src/transform/piecewise_affine.py:155-205  transform_pixel  (find_simplex, nearest-centroid fallback, A @ [x, y, 1],
//                                              bounds check, mm scale)
//   src/transform/piecewise_affine.py:207-236  transform_detection / transform_batch  (foot point x + w / 2, y + h)
// The triangulation itself (scipy.spatial.Delaunay) and the affine matrices (numpy.linalg.lstsq) are built on the host by
// the Python class exactly as the reference builds them; this file only evaluates them.  Arithmetic is float64 with explicit
// roundings in the order of the NumPy restatement (oracle/pwa_oracle.py):
//   inside test: c_i = T_i0 (x - r_x) + T_i1 (y - r_y), c_2 = 1 - c_0 - c_1, all in [-eps, 1 + eps]   (scipy qhull.pyx
//   _barycentric_inside with Delaunay.transform, eps = 100 * DBL_EPSILON); first triangle in index order;
//   fallback: argmin_i sqrt((cx_i - x)^2 + (cy_i - y)^2), first minimum   (piecewise_affine.py:138-153).
// Bandwidth-bound streaming kernel: tables staged in shared memory, one point per thread and iteration, coalesced loads.
#include <algorithm>
#include <vector>

#include "opd_common.h"

struct opd_pwa_table {
  int device = 0;
  int T = 0;
  double eps = 0.0;
  double* d_tri = nullptr;   // [T][14]: T00 T01 T10 T11 rx ry | a00 a01 a02 a10 a11 a12 | cx cy
};

namespace {

constexpr int kTriDoubles = 14;
constexpr int kMaxSmemTris = 1024;   // 112 KB of tables; larger triangulations are read through L2

struct PwaK {
  const double* tri;
  int T;
  double eps;
  const double* in;
  int input_is_bbox;
  long long N;
  double sx, sy, mw, mh;
  double* floor_px;
  double* floor_mm;
  uint8_t* in_bounds;
  int32_t* tri_idx;
  uint8_t* extrapolated;
};

__global__ void __launch_bounds__(256) pwa_transform_kernel(const PwaK p, int stage) {
  extern __shared__ __align__(16) double s_tri[];
  if (stage)
    for (int i = threadIdx.x; i < p.T * kTriDoubles; i += blockDim.x) s_tri[i] = p.tri[i];
  __syncthreads();
  const double* tri = stage ? s_tri : p.tri;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.N; i += stride) {
    double x, y;
    if (p.input_is_bbox) {
      const double bx = p.in[4 * i + 0], by = p.in[4 * i + 1], bw = p.in[4 * i + 2], bh = p.in[4 * i + 3];
      x = __dadd_rn(bx, __ddiv_rn(bw, 2.0));   // piecewise_affine.py:219-220  (x + w / 2, y + h)
      y = __dadd_rn(by, bh);
    } else {
      x = p.in[2 * i + 0];
      y = p.in[2 * i + 1];
    }
    int found = -1;
    for (int t = 0; t < p.T; ++t) {
      const double* q = tri + t * kTriDoubles;
      const double dx = __dsub_rn(x, q[4]), dy = __dsub_rn(y, q[5]);
      const double c0 = __dadd_rn(__dmul_rn(q[0], dx), __dmul_rn(q[1], dy));
      const double c1 = __dadd_rn(__dmul_rn(q[2], dx), __dmul_rn(q[3], dy));
      const double c2 = __dsub_rn(__dsub_rn(1.0, c0), c1);
      const double lo = -p.eps, hi = __dadd_rn(1.0, p.eps);
      if (c0 >= lo && c0 <= hi && c1 >= lo && c1 <= hi && c2 >= lo && c2 <= hi) {   // NaN -> outside
        found = t;
        break;
      }
    }
    const bool extra = found < 0;
    if (extra) {
      double best = INFINITY;
      found = 0;
      for (int t = 0; t < p.T; ++t) {
        const double* q = tri + t * kTriDoubles;
        const double ex = __dsub_rn(q[12], x), ey = __dsub_rn(q[13], y);
        const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)));
        if (d < best) {   // first minimum, like numpy.argmin (NaN distances never win; argmin of all-NaN is 0 as well)
          best = d;
          found = t;
        }
      }
    }
    const double* q = tri + found * kTriDoubles;
    const double fx = fma(q[6], x, fma(q[7], y, q[8])), fy = fma(q[9], x, fma(q[10], y, q[11]));
    if (p.floor_px) {
      p.floor_px[2 * i + 0] = fx;
      p.floor_px[2 * i + 1] = fy;
    }
    if (p.floor_mm) {
      p.floor_mm[2 * i + 0] = __dmul_rn(fx, p.sx);
      p.floor_mm[2 * i + 1] = __dmul_rn(fy, p.sy);
    }
    if (p.in_bounds) p.in_bounds[i] = (0.0 <= fx && fx < p.mw && 0.0 <= fy && fy < p.mh) ? 1 : 0;
    if (p.tri_idx) p.tri_idx[i] = found;
    if (p.extrapolated) p.extrapolated[i] = extra ? 1 : 0;
  }
}

}  // namespace

extern "C" int opd_pwa_table_create(const double* bary /*[T,3,2]*/, const double* affine /*[T,2,3]*/, const double* centroids /*[T,2]*/,
                                    int32_t T, double eps, int32_t device, opd_pwa_table** out) {
  OPD_REQUIRE(bary && affine && centroids && out && T > 0, "opd_pwa_table_create: bad argument (T=%d)", T);
  OPD_CUDA_OK(cudaSetDevice(device));
  std::vector<double> h((size_t)T * kTriDoubles);
  for (int t = 0; t < T; ++t) {
    double* q = h.data() + (size_t)t * kTriDoubles;
    const double* b = bary + (size_t)t * 6;
    q[0] = b[0]; q[1] = b[1]; q[2] = b[2]; q[3] = b[3]; q[4] = b[4]; q[5] = b[5];   // Delaunay.transform[t]: rows 0-1 = T^-1, row 2 = r
    for (int j = 0; j < 6; ++j) q[6 + j] = affine[(size_t)t * 6 + j];
    q[12] = centroids[2 * t];
    q[13] = centroids[2 * t + 1];
  }
  opd_pwa_table* tb = new opd_pwa_table();
  tb->device = device;
  tb->T = T;
  tb->eps = eps;
  if (cudaMalloc(&tb->d_tri, h.size() * sizeof(double)) != cudaSuccess) {
    delete tb;
    return opd::fail(OPD_ERR_CUDA, "opd_pwa_table_create: cudaMalloc of %zu bytes failed", h.size() * sizeof(double));
  }
  OPD_CUDA_OK(cudaMemcpy(tb->d_tri, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
  *out = tb;
  return OPD_OK;
}

extern "C" void opd_pwa_table_destroy(opd_pwa_table* t) {
  if (!t) return;
  cudaFree(t->d_tri);
  delete t;
}

extern "C" int opd_pwa_transform_f64(const opd_pwa_table* t, const double* in_dev, int32_t input_is_bbox, int64_t N, double scale_x_mm,
                                     double scale_y_mm, double map_w_px, double map_h_px, double* floor_px_dev, double* floor_mm_dev,
                                     uint8_t* in_bounds_dev, int32_t* tri_idx_dev, uint8_t* extrapolated_dev, void* stream) {
  OPD_REQUIRE(t && N >= 0 && (N == 0 || in_dev), "opd_pwa_transform_f64: bad argument");
  if (N == 0) return OPD_OK;
  PwaK k{t->d_tri, t->T, t->eps, in_dev, input_is_bbox, (long long)N, scale_x_mm, scale_y_mm, map_w_px, map_h_px,
         floor_px_dev, floor_mm_dev, in_bounds_dev, tri_idx_dev, extrapolated_dev};
  const int stage = t->T <= kMaxSmemTris;
  const size_t smem = stage ? (size_t)t->T * kTriDoubles * sizeof(double) : 0;
  static bool configured = false;
  if (!configured) {
    OPD_CUDA_OK(cudaFuncSetAttribute(pwa_transform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemTris * kTriDoubles * 8));
    configured = true;
  }
  const long long blocks = std::min<long long>((N + 255) / 256, 148LL * 8);
  pwa_transform_kernel<<<(unsigned)blocks, 256, smem, static_cast<cudaStream_t>(stream)>>>(k, stage);
  opd::count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}
