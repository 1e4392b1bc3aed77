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
#include <cmath>
#include <vector>

#include "opd_common.h"

struct opd_pwa_table {
  int device = 0;
  int T = 0;
  double eps = 0.0;
  double* d_tri = nullptr;   // [T][14]: T00 T01 T10 T11 rx ry | a00 a01 a02 a10 a11 a12 | cx cy
  // lookup grid over the hull's bounding box grown by half its size: cell code t = "every point of the cell is strictly inside
  // triangle t", T + t = "the cell misses every triangle and its nearest centroid is t's", -1 = decide by the full search
  int16_t* d_grid = nullptr;
  int gw = 0, gh = 0;
  double gx0 = 0, gy0 = 0, inv_cw = 0, inv_ch = 0;
  int uniform_cells = 0;
};

namespace {

constexpr int kTriDoubles = 14;
constexpr int kMaxSmemTris = 1024;   // 112 KB of tables; larger triangulations are read through L2

struct PwaK {
  const double* tri;
  int T;
  double eps;
  const int16_t* grid;
  int gw, gh;
  double gx0, gy0, inv_cw, inv_ch;
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

__device__ __forceinline__ void load_point(const PwaK& p, long long i, double& x, double& y) {
  if (p.input_is_bbox) {
    const double bx = p.in[4 * i + 0], by = p.in[4 * i + 1], bw = p.in[4 * i + 2], bh = p.in[4 * i + 3];
    x = __dadd_rn(bx, __ddiv_rn(bw, 2.0));   // piecewise_affine.py:219-220  (x + w / 2, y + h)
    y = __dadd_rn(by, bh);
  } else {
    x = p.in[2 * i + 0];
    y = p.in[2 * i + 1];
  }
}

// the reference's search: first triangle (index order) whose barycentric test passes, else the nearest centroid
__device__ __forceinline__ int full_search(const PwaK& p, const double* tri, double x, double y, bool& extra) {
  const double lo = -p.eps, hi = __dadd_rn(1.0, p.eps);
  for (int t = 0; t < p.T; ++t) {
    const double* q = tri + t * kTriDoubles;
    const double dx = __dsub_rn(x, q[4]), dy = __dsub_rn(y, q[5]);
    const double c0 = __dadd_rn(__dmul_rn(q[0], dx), __dmul_rn(q[1], dy));
    const double c1 = __dadd_rn(__dmul_rn(q[2], dx), __dmul_rn(q[3], dy));
    const double c2 = __dsub_rn(__dsub_rn(1.0, c0), c1);
    if (c0 >= lo && c0 <= hi && c1 >= lo && c1 <= hi && c2 >= lo && c2 <= hi) {   // NaN -> outside
      extra = false;
      return t;
    }
  }
  extra = true;
  double best = INFINITY;
  int found = 0;
  for (int t = 0; t < p.T; ++t) {
    const double* q = tri + t * kTriDoubles;
    const double ex = __dsub_rn(q[12], x), ey = __dsub_rn(q[13], y);
    const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)));
    if (d < best) {   // first minimum, like numpy.argmin (NaN distances never win; argmin of all-NaN is 0 as well)
      best = d;
      found = t;
    }
  }
  return found;
}

__device__ __forceinline__ void emit(const PwaK& p, const double* tri, long long i, double x, double y, int found, bool extra) {
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

// One point per lane and iteration.  Points whose grid cell has a single answer are finished at once; the others (cells cut by
// a triangle edge or a Voronoi border, points outside the grid: a few per cent) are queued per warp in shared memory and
// searched 32 at a time, so the long search never runs with most lanes idle.
__global__ void __launch_bounds__(256) pwa_transform_kernel(const PwaK p, int stage) {
  extern __shared__ __align__(16) double s_tri[];
  __shared__ unsigned s_queue[8][64];
  if (stage)
    for (int i = threadIdx.x; i < p.T * kTriDoubles; i += blockDim.x) s_tri[i] = p.tri[i];
  __syncthreads();
  const double* tri = stage ? s_tri : p.tri;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1;
  unsigned* q = s_queue[warp];
  unsigned qn = 0;   // warp-uniform
  auto drain = [&](unsigned take) {   // the last `take` queued points, one per lane
    qn -= take;
    if ((unsigned)lane < take) {
      const long long i = q[qn + lane];
      double x, y;
      load_point(p, i, x, y);
      bool extra;
      const int found = full_search(p, tri, x, y, extra);
      emit(p, tri, i, x, y, found, extra);
    }
    __syncwarp();
  };
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n_rounded = (p.N + 31) / 32 * 32;   // whole warps iterate together (ballots)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_rounded; i += stride) {
    const bool live = i < p.N;
    bool slow = false;
    if (live) {
      double x, y;
      load_point(p, i, x, y);
      int code = -1;
      if (p.grid) {
        const double gx = (x - p.gx0) * p.inv_cw, gy = (y - p.gy0) * p.inv_ch;
        if (gx >= 0.0 && gx < (double)p.gw && gy >= 0.0 && gy < (double)p.gh) code = p.grid[(int)gy * p.gw + (int)gx];
      }
      if (code >= 0) {
        const bool extra = code >= p.T;
        emit(p, tri, i, x, y, extra ? code - p.T : code, extra);
      } else {
        slow = true;
      }
    }
    const unsigned sb = __ballot_sync(0xffffffffu, slow);
    if (sb) {
      if (slow) q[qn + __popc(sb & lt_mask)] = (unsigned)i;
      qn += __popc(sb);
      __syncwarp();
      if (qn >= 32) drain(32);
    }
  }
  if (qn) drain(qn);
}

}  // namespace

extern "C" int opd_pwa_table_create(const double* bary /*[T,3,2]*/, const double* affine /*[T,2,3]*/, const double* centroids /*[T,2]*/,
                                    int32_t T, double eps, int32_t device, opd_pwa_table** out) {
  OPD_REQUIRE(bary && affine && centroids && out && T > 0, "opd_pwa_table_create: bad argument (T=%d)", T);
  opd::DeviceGuard guard(device);
  OPD_CUDA_OK(guard.err);
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
  // ---- lookup grid (float64 on the host).  Triangles and Voronoi regions are convex, so the four corners decide a cell:
  //   all barycentrics of all corners >= margin for ONE triangle            -> the cell lies strictly inside it;
  //   for EVERY triangle some barycentric is <= -margin at all four corners -> the cell misses every triangle, and if one
  //   centroid is the nearest (by a margin) at all four corners it is the nearest everywhere in the cell.
  // The margin (1e-7 in barycentric units, 1e-6 px between centroid distances) dwarfs the rounding of the cell index.
  std::vector<int16_t> grid;
  if (T < 16000) {
    double minx = INFINITY, miny = INFINITY, maxx = -INFINITY, maxy = -INFINITY;
    for (int t = 0; t < T; ++t) {   // vertices of triangle t: r (the third vertex) and r + columns of T^-1's inverse; use the centroid spread
      const double* q = h.data() + (size_t)t * kTriDoubles;
      // invert the 2x2 barycentric transform to recover the edge vectors: [v0 - r, v1 - r] = inv(Tinv)
      const double det = q[0] * q[3] - q[1] * q[2];
      const double e00 = q[3] / det, e01 = -q[1] / det, e10 = -q[2] / det, e11 = q[0] / det;   // columns: v0 - r = (e00, e10), v1 - r = (e01, e11)
      const double vx[3] = {q[4] + e00, q[4] + e01, q[4]}, vy[3] = {q[5] + e10, q[5] + e11, q[5]};
      for (int k = 0; k < 3; ++k) {
        minx = std::min(minx, vx[k]); maxx = std::max(maxx, vx[k]);
        miny = std::min(miny, vy[k]); maxy = std::max(maxy, vy[k]);
      }
    }
    const double gxs = 0.5 * (maxx - minx), gys = 0.5 * (maxy - miny);
    minx -= gxs; maxx += gxs; miny -= gys; maxy += gys;
    const int gw = 256, gh = 256;
    const double cw = (maxx - minx) / gw, ch = (maxy - miny) / gh;
    if (cw > 0 && ch > 0 && std::isfinite(cw) && std::isfinite(ch)) {
      grid.assign((size_t)gw * gh, (int16_t)-1);
      const double margin = 1e-7;
      auto bary = [&](int t, double x, double y, double* c) {
        const double* q = h.data() + (size_t)t * kTriDoubles;
        const double dx = x - q[4], dy = y - q[5];
        c[0] = q[0] * dx + q[1] * dy;
        c[1] = q[2] * dx + q[3] * dy;
        c[2] = 1.0 - c[0] - c[1];
      };
      for (int cy = 0; cy < gh; ++cy)
        for (int cx = 0; cx < gw; ++cx) {
          // corners pushed outwards by 1e-9 of a cell: the float64 cell index of a point can be off by one ulp at a cell border
          const double x0 = minx + (cx - 1e-9) * cw, x1 = minx + (cx + 1 + 1e-9) * cw, y0 = miny + (cy - 1e-9) * ch, y1 = miny + (cy + 1 + 1e-9) * ch;
          const double px[4] = {x0, x1, x0, x1}, py[4] = {y0, y0, y1, y1};
          int inside_t = -1;
          bool misses_all = true;
          for (int t = 0; t < T; ++t) {
            double c[4][3];
            for (int k = 0; k < 4; ++k) bary(t, px[k], py[k], c[k]);
            bool all_in = true;
            for (int k = 0; k < 4; ++k)
              for (int i = 0; i < 3; ++i) all_in = all_in && c[k][i] >= margin && c[k][i] <= 1.0 - margin;
            if (all_in) {
              inside_t = t;
              break;
            }
            bool separated = false;
            for (int i = 0; i < 3 && !separated; ++i) separated = c[0][i] <= -margin && c[1][i] <= -margin && c[2][i] <= -margin && c[3][i] <= -margin;
            misses_all = misses_all && separated;
          }
          if (inside_t >= 0) {
            grid[(size_t)cy * gw + cx] = (int16_t)inside_t;
          } else if (misses_all) {
            int best_t = -1;
            bool same = true;
            for (int k = 0; k < 4 && same; ++k) {
              double b0 = INFINITY, b1 = INFINITY;
              int bt = -1;
              for (int t = 0; t < T; ++t) {
                const double* q = h.data() + (size_t)t * kTriDoubles;
                const double d = std::sqrt((q[12] - px[k]) * (q[12] - px[k]) + (q[13] - py[k]) * (q[13] - py[k]));
                if (d < b0) { b1 = b0; b0 = d; bt = t; } else if (d < b1) { b1 = d; }
              }
              same = (b1 - b0 > 1e-6) && (k == 0 || bt == best_t);
              best_t = bt;
            }
            if (same && best_t >= 0) grid[(size_t)cy * gw + cx] = (int16_t)(T + best_t);
          }
        }
      tb->gw = gw; tb->gh = gh; tb->gx0 = minx; tb->gy0 = miny; tb->inv_cw = 1.0 / cw; tb->inv_ch = 1.0 / ch;
      for (int16_t v : grid) tb->uniform_cells += v >= 0;
    }
  }
  if (cudaMalloc(&tb->d_tri, h.size() * sizeof(double)) != cudaSuccess) {
    delete tb;
    return opd::fail(OPD_ERR_CUDA, "opd_pwa_table_create: cudaMalloc of %zu bytes failed", h.size() * sizeof(double));
  }
  OPD_CUDA_OK(cudaMemcpy(tb->d_tri, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
  if (!grid.empty()) {
    if (cudaMalloc(&tb->d_grid, grid.size() * sizeof(int16_t)) != cudaSuccess) {
      cudaFree(tb->d_tri);
      delete tb;
      return opd::fail(OPD_ERR_CUDA, "opd_pwa_table_create: cudaMalloc of the lookup grid failed");
    }
    OPD_CUDA_OK(cudaMemcpy(tb->d_grid, grid.data(), grid.size() * sizeof(int16_t), cudaMemcpyHostToDevice));
  }
  *out = tb;
  return OPD_OK;
}

extern "C" void opd_pwa_table_destroy(opd_pwa_table* t) {
  if (!t) return;
  cudaFree(t->d_tri);
  cudaFree(t->d_grid);
  delete t;
}

extern "C" int opd_pwa_transform_f64(const opd_pwa_table* t, const double* in_dev, int32_t input_is_bbox, int64_t N, double scale_x_mm,
                                     double scale_y_mm, double map_w_px, double map_h_px, double* floor_px_dev, double* floor_mm_dev,
                                     uint8_t* in_bounds_dev, int32_t* tri_idx_dev, uint8_t* extrapolated_dev, void* stream) {
  OPD_REQUIRE(t && N >= 0 && (N == 0 || in_dev), "opd_pwa_transform_f64: bad argument");
  if (N == 0) return OPD_OK;
  OPD_REQUIRE(N < (1ll << 32), "opd_pwa_transform_f64: at most 2^32 - 1 points per call");
  const bool use_grid = t->d_grid != nullptr && opd::g_option_probe.load() != 64;   // probe 64: full search for every point (tests)
  PwaK k{t->d_tri, t->T, t->eps, use_grid ? t->d_grid : nullptr, t->gw, t->gh, t->gx0, t->gy0, t->inv_cw, t->inv_ch, in_dev, input_is_bbox, (long long)N, scale_x_mm, scale_y_mm, map_w_px, map_h_px,
         floor_px_dev, floor_mm_dev, in_bounds_dev, tri_idx_dev, extrapolated_dev};
  const int stage = t->T <= kMaxSmemTris;
  const size_t smem = stage ? (size_t)t->T * kTriDoubles * sizeof(double) : 0;
  static opd::PerDeviceOnce configured;
  if (int rc = opd::once_per_device(configured, []() -> int {
        OPD_CUDA_OK(cudaFuncSetAttribute(pwa_transform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemTris * kTriDoubles * 8));
        return OPD_OK;
      }))
    return rc;
  const long long blocks = std::min<long long>((N + 255) / 256, 148LL * 8);
  pwa_transform_kernel<<<(unsigned)blocks, 256, smem, static_cast<cudaStream_t>(stream)>>>(k, stage);
  opd::count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Thin-plate-spline transform (src/transform/piecewise_affine.py:398-545): f(p) = a0 + a1 x + a2 y + sum_i w_i U(|p - c_i|),
// U(r) = r^2 log r (0 at r = 0), one such function per output coordinate; the coefficients are solved on the host exactly as
// the reference solves them (:445-485).  The sum runs in the reference's order (i = 0 .. n-1, product rounded, then added:
// Python floats) with explicit float64 roundings; only log() may differ from NumPy's by an ulp.
// ------------------------------------------------------------------------------------------------------------
struct opd_tps_table {
  int device = 0;
  int n = 0;
  double* d_ctrl = nullptr;   // [n][4]: cx cy wx wy
  double ax[3] = {0, 0, 0}, ay[3] = {0, 0, 0};
};

namespace {

struct TpsK {
  const double* ctrl;
  int n;
  double ax0, ax1, ax2, ay0, ay1, ay2;
  const double* in;
  int input_is_bbox;
  long long N;
  double sx, sy, mw, mh;
  double* floor_px;
  double* floor_mm;
  uint8_t* in_bounds;
};

__global__ void __launch_bounds__(256) tps_transform_kernel(const TpsK p) {
  extern __shared__ __align__(16) double s_ctrl[];
  for (int i = threadIdx.x; i < p.n * 4; i += blockDim.x) s_ctrl[i] = p.ctrl[i];
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.N; i += stride) {
    double x, y;
    if (p.input_is_bbox) {
      x = __dadd_rn(p.in[4 * i + 0], __ddiv_rn(p.in[4 * i + 2], 2.0));   // :531-533 foot point
      y = __dadd_rn(p.in[4 * i + 1], p.in[4 * i + 3]);
    } else {
      x = p.in[2 * i + 0];
      y = p.in[2 * i + 1];
    }
    double rx = 0.0, ry = 0.0;
    for (int k = 0; k < p.n; ++k) {
      const double* c = s_ctrl + 4 * k;
      const double dx = __dsub_rn(x, c[0]), dy = __dsub_rn(y, c[1]);
      const double r = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));   // np.linalg.norm of a 2-vector
      const double u = r > 0.0 ? __dmul_rn(__dmul_rn(r, r), log(r)) : 0.0;            // _radial_basis (:437-443)
      rx = __dadd_rn(rx, __dmul_rn(c[2], u));
      ry = __dadd_rn(ry, __dmul_rn(c[3], u));
    }
    const double fx = __dadd_rn(__dadd_rn(__dadd_rn(p.ax0, __dmul_rn(p.ax1, x)), __dmul_rn(p.ax2, y)), rx);   // :504-505, left to right
    const double fy = __dadd_rn(__dadd_rn(__dadd_rn(p.ay0, __dmul_rn(p.ay1, x)), __dmul_rn(p.ay2, y)), ry);
    if (p.floor_px) {
      p.floor_px[2 * i + 0] = fx;
      p.floor_px[2 * i + 1] = fy;
    }
    if (p.floor_mm) {
      p.floor_mm[2 * i + 0] = __dmul_rn(fx, p.sx);
      p.floor_mm[2 * i + 1] = __dmul_rn(fy, p.sy);
    }
    if (p.in_bounds) p.in_bounds[i] = (0.0 <= fx && fx < p.mw && 0.0 <= fy && fy < p.mh) ? 1 : 0;
  }
}

}  // namespace

extern "C" int opd_tps_table_create(const double* src_points /*[n,2]*/, const double* weights_x, const double* weights_y,
                                    const double* affine_x /*[3]*/, const double* affine_y /*[3]*/, int32_t n, int32_t device,
                                    opd_tps_table** out) {
  OPD_REQUIRE(src_points && weights_x && weights_y && affine_x && affine_y && out && n > 0 && n <= 4096,
              "opd_tps_table_create: bad argument (n=%d, at most 4096 control points)", n);
  opd::DeviceGuard guard(device);
  OPD_CUDA_OK(guard.err);
  std::vector<double> h((size_t)n * 4);
  for (int i = 0; i < n; ++i) {
    h[4 * i + 0] = src_points[2 * i];
    h[4 * i + 1] = src_points[2 * i + 1];
    h[4 * i + 2] = weights_x[i];
    h[4 * i + 3] = weights_y[i];
  }
  opd_tps_table* tb = new opd_tps_table();
  tb->device = device;
  tb->n = n;
  for (int j = 0; j < 3; ++j) {
    tb->ax[j] = affine_x[j];
    tb->ay[j] = affine_y[j];
  }
  if (cudaMalloc(&tb->d_ctrl, h.size() * sizeof(double)) != cudaSuccess) {
    delete tb;
    return opd::fail(OPD_ERR_CUDA, "opd_tps_table_create: cudaMalloc failed");
  }
  OPD_CUDA_OK(cudaMemcpy(tb->d_ctrl, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
  *out = tb;
  return OPD_OK;
}

extern "C" void opd_tps_table_destroy(opd_tps_table* t) {
  if (!t) return;
  cudaFree(t->d_ctrl);
  delete t;
}

extern "C" int opd_tps_transform_f64(const opd_tps_table* t, const double* in_dev, int32_t input_is_bbox, int64_t N, double scale_x_mm,
                                     double scale_y_mm, double map_w_px, double map_h_px, double* floor_px_dev, double* floor_mm_dev,
                                     uint8_t* in_bounds_dev, void* stream) {
  OPD_REQUIRE(t && N >= 0 && (N == 0 || in_dev), "opd_tps_transform_f64: bad argument");
  if (N == 0) return OPD_OK;
  TpsK k{t->d_ctrl, t->n, t->ax[0], t->ax[1], t->ax[2], t->ay[0], t->ay[1], t->ay[2], in_dev, input_is_bbox, (long long)N,
         scale_x_mm, scale_y_mm, map_w_px, map_h_px, floor_px_dev, floor_mm_dev, in_bounds_dev};
  const size_t smem = (size_t)t->n * 4 * sizeof(double);
  static opd::PerDeviceOnce configured;
  if (int rc = opd::once_per_device(configured, []() -> int {
        OPD_CUDA_OK(cudaFuncSetAttribute(tps_transform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 4 * 8));
        return OPD_OK;
      }))
    return rc;
  const long long blocks = std::min<long long>((N + 255) / 256, 148LL * 8);
  tps_transform_kernel<<<(unsigned)blocks, 256, smem, static_cast<cudaStream_t>(stream)>>>(k);
  opd::count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Lens-distortion correction of points (src/calibration/lens_distortion.py:156-203 -> cv2.undistortPoints(pts, K, dist, P=K)):
// OpenCV's fixed-point iteration with its default criteria (exactly 5 iterations), distortion model (k1, k2, p1, p2, k3),
// output re-projected with the same camera matrix.  Restated from OpenCV's cvUndistortPointsInternal; float64.
// ------------------------------------------------------------------------------------------------------------
namespace {

struct UndistortK {
  double fx, fy, cx, cy, k1, k2, p1, p2, k3;
  const double* in;
  int input_is_bbox;
  long long N;
  double* out;
};

__global__ void __launch_bounds__(256) undistort_points_kernel(const UndistortK p) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.N; i += stride) {
    double u, v;
    if (p.input_is_bbox) {
      u = __dadd_rn(p.in[4 * i + 0], __ddiv_rn(p.in[4 * i + 2], 2.0));   // foot point of an (x, y, w, h) box
      v = __dadd_rn(p.in[4 * i + 1], p.in[4 * i + 3]);
    } else {
      u = p.in[2 * i + 0];
      v = p.in[2 * i + 1];
    }
    const double ifx = 1.0 / p.fx, ify = 1.0 / p.fy;
    double x = (u - p.cx) * ifx, y = (v - p.cy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; ++j) {
      const double r2 = x * x + y * y;
      const double icdist = 1.0 / (1.0 + ((p.k3 * r2 + p.k2) * r2 + p.k1) * r2);
      if (icdist < 0.0) {
        x = (u - p.cx) * ifx;
        y = (v - p.cy) * ify;
        break;
      }
      const double dx = 2.0 * p.p1 * x * y + p.p2 * (r2 + 2.0 * x * x);
      const double dy = p.p1 * (r2 + 2.0 * y * y) + 2.0 * p.p2 * x * y;
      x = (x0 - dx) * icdist;
      y = (y0 - dy) * icdist;
    }
    p.out[2 * i + 0] = p.fx * x + p.cx;
    p.out[2 * i + 1] = p.fy * y + p.cy;
  }
}

}  // namespace

extern "C" int opd_undistort_points_f64(double fx, double fy, double cx, double cy, double k1, double k2, double p1, double p2, double k3,
                                        const double* in_dev, int32_t input_is_bbox, int64_t N, double* out_dev, void* stream) {
  OPD_REQUIRE(N >= 0 && (N == 0 || (in_dev && out_dev)) && fx != 0.0 && fy != 0.0, "opd_undistort_points_f64: bad argument");
  if (N == 0) return OPD_OK;
  UndistortK k{fx, fy, cx, cy, k1, k2, p1, p2, k3, in_dev, input_is_bbox, (long long)N, out_dev};
  const long long blocks = std::min<long long>((N + 255) / 256, 148LL * 8);
  undistort_points_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(k);
  opd::count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}
