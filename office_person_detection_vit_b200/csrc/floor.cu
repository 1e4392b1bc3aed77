// K9 + K10 + K11: foot-point homography, polygon zone classification and per-timestamp zone counting,
// fused into one pass over the points.
//
// Replaces (reference, pure Python / NumPy float64):
//   src/transform/homography.py:150-197   transform_batch  (foot point, H·p, perspective divide, mm, bounds)
//   src/zone/zone_classifier.py:114-197   classify + _point_in_polygon (ray casting, priority pick)
//   src/aggregation/aggregator.py:52-75   get_zone_counts  (histogram with an "unclassified" bin)
//
// Design: the polygons are rasterised ON THE HOST, in float64, into a uniform grid of
// one byte per cell.  A cell whose Δ-dilated box is crossed by no polygon edge has a single answer for every
// point in it ("uniform" cell: the byte is the class id); the other cells are "boundary" cells (byte 255) and
// carry the list of polygons that must be tested exactly.  The exact test is the reference's ray casting,
// operation by operation, in float64.  Two kernels use the table:
//   floor_exact_kernel  — float64 arithmetic throughout, every output, any histogram shape;
//   floor_fast_kernel   — the HBM-roofline path: float32 projection with a rigorous per-point error bound
//                         as a *filter*; a point is answered from the grid only when bound <= Δ/2 and its
//                         cell is uniform, otherwise its index goes to a shared-memory queue that the CTA
//                         drains densely through the float64 exact path.  Counting uses warp ballots
//                         (no per-point atomics).  Results are identical to floor_exact_kernel.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

#include "opd_common.h"

namespace {

constexpr int kBoundary = 255;          // grid byte of a boundary cell
constexpr int kMaxClasses = 255;        // class ids 0..254
constexpr int kGridCellBudget = 128 * 1024;
// Winner grid of the float32 path, one unsigned byte per cell: 0 .. Z-1 = the cell lies inside that zone (after the priority
// pick), kByteNone = outside every zone, kByteSlow = the float64 path decides, Z + t (t < 254 - Z) = the cell is crossed by exactly
// one line, type t of the line table (see opd_zone_table_create).  Per-point codes in the kernel: zone, kCodeNone or kCodeSlow.
constexpr int kByteNone = 255, kByteSlow = 254;
constexpr int kMaxLineRecords = 256;    // one table for the whole grid: up to 254 - Z line types
constexpr int kTypeTile = 32;           // tiled type tables (more triples than a byte can name): tiles of 32 x 32 cells
constexpr int kMaxTiledRecords = 2048;  // 32 KB of shared memory
constexpr int kMaxTiles = 1024;
constexpr int kCodeNone = -1, kCodeSlow = -2;
constexpr int kMaxSmemVerts = 1024;     // polygons vertices staged in shared memory (16 KB)
constexpr int kFastThreads = 512;
constexpr int kFastPointsPerThread = 8;
constexpr int kFastChunk = kFastThreads * kFastPointsPerThread;

// ---------------------------------------------------------------------------------------------------------
// The reference's ray casting (zone_classifier.py:162-197), float64, same operations in the same order.
// Rounding of every operation is explicit (__d*_rn on the device) so no FMA contraction can change it.
// ---------------------------------------------------------------------------------------------------------
__host__ __device__ inline bool point_in_polygon_ref(double x, double y, const double2* v, int n) {
  bool inside = false;
  double p1x = v[0].x, p1y = v[0].y;
  for (int i = 1; i <= n; ++i) {
    const double2 q = v[i == n ? 0 : i];
    const double p2x = q.x, p2y = q.y;
    const double ymin = p1y < p2y ? p1y : p2y;   // Python min(p1y, p2y) on finite floats
    const double ymax = p1y < p2y ? p2y : p1y;
    const double xmax = p1x < p2x ? p2x : p1x;
    if (y > ymin && y <= ymax && x <= xmax) {
      // here p1y != p2y (ymin < ymax), so the reference always (re)computes xinters (:187-189)
#ifdef __CUDA_ARCH__
      // vertical edge: `p1x == p2x or ...` is True whatever xinters is; skip the float64 division (same result)
      if (p1x == p2x) {
        inside = !inside;
        p1x = p2x;
        p1y = p2y;
        continue;
      }
      const double xinters =
          __dadd_rn(__ddiv_rn(__dmul_rn(__dsub_rn(y, p1y), __dsub_rn(p2x, p1x)), __dsub_rn(p2y, p1y)), p1x);
#else
      volatile double t0 = y - p1y, t1 = p2x - p1x;
      volatile double t2 = t0 * t1;
      volatile double t3 = p2y - p1y;
      volatile double t4 = t2 / t3;
      const double xinters = t4 + p1x;
#endif
      if (p1x == p2x || x <= xinters) inside = !inside;
    }
    p1x = p2x;
    p1y = p2y;
  }
  return inside;
}

struct FloorK {
  // projection
  double H[9];
  double sx, sy, mw, mh;
  float Hf[9];        // float32 copy of H for the filter path
  float Sxy[3];       // max(|H0j|, |H1j|) — magnitude bound of the X and Y rows
  float Sw[3];        // |H2j|
  float err_max;      // Δ/2
  float fk[3];        // fast filter: K(x, y) = fk0 |x| + fk1 |y| + fk2 >= Sxy + PMAX Sw
  float fT1, fT2;     // thresholds on |r̂| K for in-grid / out-of-grid points (floor_fast_kernel)
  float fgx, fgy;     // -gx0 * inv_cw, -gy0 * inv_ch
  // grid
  double gx0, gy0, inv_cw, inv_ch;
  float gx0_f, gy0_f, inv_cw_f, inv_ch_f;
  int gw, gh;
  int Z, allow_overlap, input_is_bbox, skip_projection;
  int n_verts, stage_grid, stage_verts;
  const uint8_t* grid;
  const uint64_t* class_mask;    // [256]
  const int32_t* class_winner;   // [256]
  const int32_t* cell_entry;     // [gw*gh]
  const uint64_t* entry_inside;  // [n_boundary]
  const uint64_t* entry_cand;    // [n_boundary]
  const double2* verts;          // [n_verts]
  const int32_t* poly_off;       // [Z+1]
  const int32_t* zone_rank;      // [64]
  const uint8_t* wgrid;          // [gw*gh] winner grid of the float32 path (see kByteNone)
  const float4* line_types;      // [n_line_types]
  int n_line_types;
  const uint16_t* tile_base;     // tiled type tables: [n_tiles] first record of every 32 x 32-cell tile
  int tiles_x, n_tiles;
  float line_band;               // a point closer than this to its cell's line goes to the float64 path
  int l2_ahead;                  // floor_fast_kernel: units (per warp) pulled into L2 ahead of the register loads
  // io
  const void* in;
  const int32_t* slot;
  long long N;
  int T;
  void* floor_px;
  void* floor_mm;
  uint8_t* in_bounds;
  int32_t* zone_idx;
  uint64_t* zone_mask;
  int32_t* hist;
};

struct SmemTables {
  const uint8_t* grid;
  const uint64_t* class_mask;
  const int32_t* class_winner;
  const double2* verts;
  const int32_t* poly_off;
  const int32_t* zone_rank;
};

__device__ __forceinline__ int winner_of(uint64_t m, const int32_t* zone_rank) {
  int best = -1, best_rank = 0x7fffffff;
  while (m) {
    const int z = __ffsll((long long)m) - 1;
    m &= m - 1;
    const int r = zone_rank[z];
    if (r < best_rank) {
      best_rank = r;
      best = z;
    }
  }
  return best;
}

// float64 classification of one floor point through the grid (+ exact ray casting in boundary cells).
// Returns the containing-zones mask (declaration-order bits); *win = selected zone or -1.
__device__ __forceinline__ uint64_t classify_exact(const FloorK& p, const SmemTables& t, double px, double py,
                                                   int* win) {
  uint64_t mask = 0;
  int w = -1;
  const double fx = (px - p.gx0) * p.inv_cw;
  const double fy = (py - p.gy0) * p.inv_ch;
  if (fx >= 0.0 && fx < (double)p.gw && fy >= 0.0 && fy < (double)p.gh) {  // NaN -> outside, like the reference
    const int cell = (int)fy * p.gw + (int)fx;
    const int code = t.grid[cell];
    if (code != kBoundary) {
      mask = t.class_mask[code];
      w = t.class_winner[code];
    } else {
      const int e = p.cell_entry[cell];
      mask = p.entry_inside[e];
      uint64_t cand = p.entry_cand[e];
      while (cand) {
        const int z = __ffsll((long long)cand) - 1;
        cand &= cand - 1;
        const int o = t.poly_off[z];
        if (point_in_polygon_ref(px, py, t.verts + o, t.poly_off[z + 1] - o)) mask |= 1ull << z;
      }
      w = winner_of(mask, t.zone_rank);
      if (!p.allow_overlap) mask = w >= 0 ? 1ull << w : 0ull;
    }
  }
  *win = w;
  return mask;
}

// H·[x, y, 1] then perspective divide, float64 (homography.py:172-175).
__device__ __forceinline__ void project_exact(const FloorK& p, double x, double y, double* px, double* py) {
  const double X = fma(p.H[2], 1.0, fma(p.H[1], y, p.H[0] * x));
  const double Y = fma(p.H[5], 1.0, fma(p.H[4], y, p.H[3] * x));
  const double W = fma(p.H[8], 1.0, fma(p.H[7], y, p.H[6] * x));
  *px = __ddiv_rn(X, W);
  *py = __ddiv_rn(Y, W);
}

__device__ __forceinline__ SmemTables stage_tables(const FloorK& p, unsigned char* smem, int cells_rounded) {
  // layout: [grid bytes][class_mask 256*8][class_winner 256*4][zone_rank 64*4][poly_off 68*4][verts]
  SmemTables t;
  unsigned char* cur = smem;
  uint8_t* s_grid = cur;
  cur += p.stage_grid ? cells_rounded : 0;
  uint64_t* s_cm = reinterpret_cast<uint64_t*>(cur);
  cur += 256 * 8;
  int32_t* s_cw = reinterpret_cast<int32_t*>(cur);
  cur += 256 * 4;
  int32_t* s_rank = reinterpret_cast<int32_t*>(cur);
  cur += 64 * 4;
  int32_t* s_off = reinterpret_cast<int32_t*>(cur);
  cur += 68 * 4;
  double2* s_verts = reinterpret_cast<double2*>(cur);

  if (p.stage_grid) {
    const uint4* src = reinterpret_cast<const uint4*>(p.grid);
    uint4* dst = reinterpret_cast<uint4*>(s_grid);
    for (int i = threadIdx.x; i < cells_rounded / 16; i += blockDim.x) dst[i] = src[i];
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    s_cm[i] = p.class_mask[i];
    s_cw[i] = p.class_winner[i];
  }
  for (int i = threadIdx.x; i < 64; i += blockDim.x) s_rank[i] = p.zone_rank[i];
  for (int i = threadIdx.x; i <= p.Z; i += blockDim.x) s_off[i] = p.poly_off[i];
  if (p.stage_verts)
    for (int i = threadIdx.x; i < p.n_verts; i += blockDim.x) s_verts[i] = p.verts[i];
  t.grid = p.stage_grid ? s_grid : p.grid;
  t.class_mask = s_cm;
  t.class_winner = s_cw;
  t.zone_rank = s_rank;
  t.poly_off = s_off;
  t.verts = p.stage_verts ? s_verts : p.verts;
  return t;
}

__host__ __device__ inline size_t tables_smem_bytes(int stage_grid, int cells_rounded, int stage_verts, int n_verts) {
  return (size_t)(stage_grid ? cells_rounded : 0) + 256 * 8 + 256 * 4 + 64 * 4 + 68 * 4 +
         (stage_verts ? (size_t)n_verts * 16 : 0);
}

// ---------------------------------------------------------------------------------------------------------
// floor_exact_kernel: one point per thread and iteration, float64 arithmetic, every output optional.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) floor_exact_kernel(const FloorK p, int cells_rounded) {
  extern __shared__ __align__(16) unsigned char smem[];
  const SmemTables t = stage_tables(p, smem, cells_rounded);
  __syncthreads();

  const T* in = static_cast<const T*>(p.in);
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n_rounded = (p.N + 31) / 32 * 32;  // keep warps converged for match_any
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_rounded; i += stride) {
    const bool live = i < p.N;
    int key = -1;
    if (live) {
      double x, y;
      if (p.input_is_bbox) {
        const double bx = (double)in[4 * i + 0], by = (double)in[4 * i + 1];
        const double bw = (double)in[4 * i + 2], bh = (double)in[4 * i + 3];
        x = __dadd_rn(bx, __ddiv_rn(bw, 2.0));  // homography.py:167  [x + w / 2, y + h]
        y = __dadd_rn(by, bh);
      } else {
        x = (double)in[2 * i + 0];
        y = (double)in[2 * i + 1];
      }
      double px = x, py = y;
      if (!p.skip_projection) project_exact(p, x, y, &px, &py);
      if (p.floor_px) {
        static_cast<T*>(p.floor_px)[2 * i + 0] = (T)px;
        static_cast<T*>(p.floor_px)[2 * i + 1] = (T)py;
      }
      if (p.floor_mm) {
        static_cast<T*>(p.floor_mm)[2 * i + 0] = (T)__dmul_rn(px, p.sx);  // homography.py:183-186
        static_cast<T*>(p.floor_mm)[2 * i + 1] = (T)__dmul_rn(py, p.sy);
      }
      if (p.in_bounds) p.in_bounds[i] = (0.0 <= px && px < p.mw && 0.0 <= py && py < p.mh) ? 1 : 0;
      // a row whose slot lies outside [0, T) is an UNUSED row (the padding rows of the detector's [B, 100] tables carry
      // slot -1): it is neither classified nor counted, its zone index is -1 and its mask 0
      const int s = p.slot ? p.slot[i] : 0;
      const bool used = p.slot == nullptr || (unsigned)s < (unsigned)p.T;
      if (!used) {
        if (p.zone_idx) p.zone_idx[i] = -1;
        if (p.zone_mask) p.zone_mask[i] = 0;
      } else if (p.zone_idx || p.zone_mask || p.hist) {
        int win;
        const uint64_t mask = classify_exact(p, t, px, py, &win);
        if (p.zone_idx) p.zone_idx[i] = win;
        if (p.zone_mask) p.zone_mask[i] = mask;
        if (p.hist) {
          {
            const int row = s * (p.Z + 1);
            if (mask == 0) {
              key = row + p.Z;  // aggregator.py:71-73 "unclassified"
            } else if ((mask & (mask - 1)) == 0) {
              key = row + (__ffsll((long long)mask) - 1);
            } else {  // several zones: one count each (aggregator.py:66-69); rare, plain atomics
              uint64_t m = mask;
              while (m) {
                atomicAdd(p.hist + row + (__ffsll((long long)m) - 1), 1);
                m &= m - 1;
              }
            }
          }
        }
      }
    }
    if (p.hist) {  // warp-aggregated atomics: one per distinct (slot, bin) in the warp
      const unsigned peers = __match_any_sync(0xffffffffu, key);
      if (key >= 0 && lane == __ffs(peers) - 1) atomicAdd(p.hist + key, __popc(peers));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// floor_fast_kernel: float32 points in, int32 zone index and/or row-0 histogram out.
//
// Work unit = 256 consecutive points per warp and iteration: four fully coalesced 16-byte loads per lane (load j, lane l: points
// 2 (32 j + l) and + 1 of the unit) and four 8-byte index stores.  Everything loop-invariant lives in registers (512 threads x
// up to 128 registers), the winner grid is one signed byte per cell in shared memory, the per-point work is branch-free.
// Warps never synchronise with each other inside the loop: every warp owns a private slow-path queue in shared memory and
// drains it, 32 points at a time, through the float64 exact path.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_stream(const float4* ptr) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(ptr));
  return r;
}

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// A value the compiler must keep in a register: without this it re-reads the loop-invariant kernel parameters from the constant
// bank (and re-derives shared-memory window addresses) in every iteration, a sixth of the main loop's instructions.
__device__ __forceinline__ float pin_reg(float x) {
  float r;
  asm volatile("mov.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ uint32_t pin_reg(uint32_t x) {
  uint32_t r;
  asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(x));
  return r;
}

__device__ __forceinline__ void atoms_inc(uint32_t addr) {   // shared-memory counter += 1 (no return value)
  asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
}

__device__ __forceinline__ int lds_u8(uint32_t addr) {   // byte load from a 32-bit shared-memory address
  int v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// The float32 filter.  p̂ = (X̂ r̂, Ŷ r̂) with r̂ = rcp(Ŵ) satisfies, per coordinate (u = 2^-24; X̂, Ŷ, Ŵ carry <= 4u·S from
// two fma roundings and the float32 rounding of H, the approximate reciprocal 2u, the products u; constants generous):
//   |p̂ - p| <= 16u·(|r̂|·(Sxy + max|p̂|·Sw) + max|p̂|),   Sxy >= |H0·|(|x|,|y|,1), |H1·|(|x|,|y|,1),  Sw = |H2·|(|x|,|y|,1).
// With PMAX = the largest |coordinate| of the grid box and K(x, y) = Sxy + PMAX·Sw (three multiply-adds with constants
// from the host) the per-point decision needs ONE product q = |r̂|·K:
//   p̂ inside the grid box (max|p̂| <= PMAX):  |p̂ - p| <= 16u·(q + PMAX) <= Δ/2  <=>  q <= T1 = Δ/(32u) - PMAX
//       -> the cell of p̂ answers for p unless it is a boundary cell (cells within Δ of an edge are boundary cells);
//   p̂ outside the grid box by d >= 0 (max|p̂| <= PMAX + d), m = the box's margin around every polygon (>= 1 px):
//       |p̂ - p| <= 16u·(q + PMAX) + 16u·d·(|r̂|·Sw + 1) <= m/2 + d/2 < m + d   when  q <= T2 = min(m/(32u), PMAX/(32u)) - PMAX
//       (|r̂|·Sw <= q/PMAX), so p is outside every polygon's bounding box -> no zone.
// Everything else (q too large, NaN / inf, boundary cells) is decided by the float64 path, so the result never depends
// on float32 rounding.  The constants are computed in float64 by fast_consts() and rounded toward the safe side.
//
// Per-point code (int): 0 .. Z-1 = zone, kCodeNone = outside every zone, kCodeSlow = the float64 path decides.  The winner grid
// holds exactly these codes as signed bytes, so the loaded byte IS the stored zone index; a slow point's slot temporarily holds
// kCodeSlow and is overwritten by the drain.  Counters: bin code + 2 of the warp's private histogram, unconditionally (bins 0 and
// 1 absorb the slow placeholders and the zone-less points).
constexpr int kLaneSlots = 16;    // per-lane queue depth (a unit adds at most 8 entries per lane)
constexpr int kExactQueue = 128;  // dense float64 queue entries per warp
constexpr int kHistBins = 66;     // bin of code c: c + 2 (final codes -2 .. 63)
constexpr int kUnitPoints = 256;

// Per-point code (int): 0 .. Z-1 = zone, kCodeNone = outside every zone, kCodeSlow = the float64 path decides.  For most points
// the winner grid byte u of the point's cell (see kByteNone) IS the code (sign-extended).  A byte in [Z, Z + n_types) names the
// line type of a cell crossed by exactly one line: that lane reads the 16-byte record {a, b, c, codes} and the SIGN of
// a x + b y + c picks one of the record's two codes unless the point lies inside the error band of the line (-> kCodeSlow).
// The test is branch-free (one predicated shared-memory load).
// Points whose final code is kCodeSlow are "queued": their index goes to the owning lane's private queue in shared memory (three
// predicated instructions, no warp-wide coordination), their slot temporarily holds kCodeSlow and is overwritten by the drain.
// Counters: bin code + 2 of the warp's private histogram, unconditionally (bins 0 and 1 absorb the placeholders and the zone-less
// points and are never read).
//
// Drain, when a lane's queue is half full and at the end: the lane queues are compacted row by row (ballot) into the warp's dense
// float64 queue, which is processed 32 points at a time through the reference's exact arithmetic.
// kTiled: the line-type byte is relative to the cell's 32 x 32-cell tile (record = tile_base[tile] + t), see opd_zone_table_create.
template <bool kTiled>
__global__ void __launch_bounds__(kFastThreads, 1) floor_fast_kernel(const FloorK p, int cells_rounded) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int kWarps = kFastThreads / 32;
  FloorK pg = p;
  pg.stage_grid = 0;   // the float64 path reads the CLASS grid from global memory (L2); shared memory holds the WINNER grid
  const SmemTables t = stage_tables(pg, smem + cells_rounded, cells_rounded);
  uint8_t* s_wgrid = smem;                                                              // [cells_rounded] winner grid
  unsigned char* cur = smem + cells_rounded + (tables_smem_bytes(0, cells_rounded, p.stage_verts, p.n_verts) + 15) / 16 * 16;
  float4* s_types = reinterpret_cast<float4*>(cur);      // [kMaxLineRecords] or, tiled, [kMaxTiledRecords] + tile_base [kMaxTiles] u16
  cur += (kTiled ? kMaxTiledRecords : kMaxLineRecords) * 16;
  uint16_t* s_tile_base = reinterpret_cast<uint16_t*>(cur);
  if (kTiled) cur += kMaxTiles * 2;
  unsigned* s_lq = reinterpret_cast<unsigned*>(cur);     // [warps][kLaneSlots][32]: lane queues
  cur += kWarps * kLaneSlots * 32 * 4;
  unsigned* s_xq = reinterpret_cast<unsigned*>(cur);     // [warps][kExactQueue]: dense float64 queues
  cur += kWarps * kExactQueue * 4;
  unsigned* s_whist = reinterpret_cast<unsigned*>(cur);  // [warps][kHistBins]: private counters, one shared-memory atomic per point
  unsigned* s_extra = s_whist + kWarps * kHistBins;      // [warps]: counts beyond the first of points in several zones
  for (int i = threadIdx.x; i < kWarps * (kHistBins + 1); i += kFastThreads) s_whist[i] = 0;
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.wgrid);
    uint4* dst = reinterpret_cast<uint4*>(s_wgrid);
    for (int i = threadIdx.x; i < cells_rounded / 16; i += kFastThreads) dst[i] = src[i];
    for (int i = threadIdx.x; i < p.n_line_types; i += kFastThreads) s_types[i] = p.line_types[i];
    if (kTiled)
      for (int i = threadIdx.x; i < p.n_tiles; i += kFastThreads) s_tile_base[i] = p.tile_base[i];
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned* xq = s_xq + warp * kExactQueue;
  unsigned* wh = s_whist + warp * kHistBins;
  const bool do_hist = p.hist != nullptr;
  unsigned xn = 0;   // fill of the dense float64 queue (warp-uniform)

  const float* in = static_cast<const float*>(p.in);
  const float2* in2 = reinterpret_cast<const float2*>(in);

  // ---- constants of the float32 filter, pinned in registers ----
  const float h0 = pin_reg(p.Hf[0]), h1 = pin_reg(p.Hf[1]), h2 = pin_reg(p.Hf[2]), h3 = pin_reg(p.Hf[3]), h4 = pin_reg(p.Hf[4]),
              h5 = pin_reg(p.Hf[5]), h6 = pin_reg(p.Hf[6]), h7 = pin_reg(p.Hf[7]), h8 = pin_reg(p.Hf[8]);
  const float k0 = pin_reg(p.fk[0]), k1 = pin_reg(p.fk[1]), k2 = pin_reg(p.fk[2]), T1 = pin_reg(p.fT1), T2 = pin_reg(p.fT2);
  const float icw = pin_reg(p.inv_cw_f), ich = pin_reg(p.inv_ch_f), ox = pin_reg(p.fgx), oy = pin_reg(p.fgy);   // fx = px * icw - gx0 * icw
  const unsigned gw = pin_reg((uint32_t)p.gw), gh = pin_reg((uint32_t)p.gh);
  const uint32_t grid_base = pin_reg((uint32_t)__cvta_generic_to_shared(s_wgrid));
  const uint32_t hist_base = pin_reg((uint32_t)__cvta_generic_to_shared(wh + 2));   // bin of code c: hist_base + 4 c
  const uint32_t type_base = pin_reg((uint32_t)__cvta_generic_to_shared(s_types));
  const float band = pin_reg(p.line_band);
  // a type byte names a line when it is below n_types: the table's size, or (tiled) every byte short of kByteSlow / kByteNone
  const unsigned zu = pin_reg((uint32_t)p.Z), n_types = pin_reg(kTiled ? (uint32_t)(kByteSlow - p.Z) : (uint32_t)p.n_line_types);
  const uint32_t tile_addr = pin_reg((uint32_t)__cvta_generic_to_shared(s_tile_base));
  const unsigned tiles_x = pin_reg((uint32_t)p.tiles_x);
  float la = 0.f, lb = 0.f, lc = 1.f;   // the last line record read by this lane
  uint32_t lcodes = 0;
  const uint32_t lq_base = pin_reg((uint32_t)__cvta_generic_to_shared(s_lq + warp * kLaneSlots * 32 + lane));
  uint32_t lq_top = lq_base;   // shared address of this lane's next free slot (slots are 128 bytes apart)
  int32_t* const out_idx = p.zone_idx;
  const unsigned n_points = (unsigned)p.N;

  // tier 2: float64 pass over the dense queue, 32 points at a time, until at most `keep` remain
  auto drain_exact = [&](unsigned keep) {
    __syncwarp();
    while (xn > keep) {
      const unsigned take = xn < 32u ? xn : 32u;
      xn -= take;
      if ((unsigned)lane < take) {
        const unsigned i = xq[xn + lane];
        const float2 xy = in2[i];
        double px, py;
        int win = -1;
        project_exact(p, (double)xy.x, (double)xy.y, &px, &py);
        uint64_t m = classify_exact(pg, t, px, py, &win);
        if (p.zone_idx) p.zone_idx[i] = win;
        if (do_hist && win >= 0) {
          if (!p.allow_overlap) {
            atomicAdd(&wh[win + 2], 1u);
          } else {   // one count per containing zone
            atomicAdd(&s_extra[warp], (unsigned)__popcll(m) - 1u);
            while (m) {
              atomicAdd(&wh[__ffsll((long long)m) + 1], 1u);
              m &= m - 1;
            }
          }
        }
      }
      __syncwarp();
    }
  };

  // one point in float32 -> its final code
  auto filter = [&](float x, float y) -> int {
    const float X = fmaf(h0, x, fmaf(h1, y, h2));
    const float Y = fmaf(h3, x, fmaf(h4, y, h5));
    const float W = fmaf(h6, x, fmaf(h7, y, h8));
    const float r = rcp_approx(W);
    const float px = X * r, py = Y * r;
    const float K = fmaf(fabsf(x), k0, fmaf(fabsf(y), k1, k2));
    const float qq = fabsf(r) * K;
    const int ix = __float2int_rd(fmaf(px, icw, ox)), iy = __float2int_rd(fmaf(py, ich, oy));
    const bool in_grid = ((unsigned)ix < gw) & ((unsigned)iy < gh);      // NaN -> cell 0, guarded by the q tests (NaN: false)
    int u = (!in_grid & (qq <= T2)) ? kByteNone : kByteSlow;
    if (in_grid & (qq <= T1)) u = lds_u8(grid_base + (unsigned)iy * gw + (unsigned)ix);
    // single-line cells only (a few percent of the lanes) read their 16-byte record {a, b, c, codes}: a predicated load, so the
    // other lanes cost the shared-memory pipe nothing; their registers keep the previous record (finite, unused)
    unsigned t = (unsigned)u - zu;
    const bool is_line = t < n_types;
    if (kTiled) {   // + the first record of the cell's tile (the same few lanes; ix, iy are in the grid whenever is_line)
      unsigned tb = 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p ld.shared.u16 %0, [%1];\n\t}"
                   : "+r"(tb)
                   : "r"(tile_addr + 2u * (((unsigned)iy / kTypeTile) * tiles_x + (unsigned)ix / kTypeTile)), "r"((unsigned)is_line));
      t += tb;
    }
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n\t}"
                 : "+f"(la), "+f"(lb), "+f"(lc), "+r"(lcodes)
                 : "r"(type_base + 16u * t), "r"((unsigned)is_line));
    const float d = is_line ? fmaf(la, px, fmaf(lb, py, lc)) : 1.f;
    // code A (byte 0) or code B (byte 1) of the record / the grid byte itself (254, 255 -> kCodeSlow, kCodeNone), sign-extended
    // by the permute (selector bit 3 replicates the sign)
    int z, zu8;
    asm("prmt.b32 %0, %1, 0, %2;" : "=r"(z) : "r"(lcodes), "r"(d > 0.f ? 0x8880u : 0x9991u));
    asm("prmt.b32 %0, %1, 0, 0x8880;" : "=r"(zu8) : "r"(u));
    const int code = is_line ? z : zu8;
    return fabsf(d) >= band ? code : kCodeSlow;   // inside the band of the cell's line (or NaN): float64 path
  };

  // compaction of the lane queues (row s = every lane's s-th entry) into the dense float64 queue, processed down to `keep` entries
  auto drain = [&](unsigned keep) {
    __syncwarp();
    const unsigned mine = (lq_top - lq_base) >> 7;
    unsigned rows = mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rows = max(rows, __shfl_xor_sync(0xffffffffu, rows, o));
    for (unsigned sl = 0; sl < rows; ++sl) {
      const bool have = sl < mine;
      const unsigned eb = __ballot_sync(0xffffffffu, have);
      if (have) xq[xn + __popc(eb & ((1u << lane) - 1u))] = s_lq[(warp * kLaneSlots + sl) * 32 + lane];
      xn += __popc(eb);
      if (xn > (unsigned)(kExactQueue - 32)) drain_exact(kExactQueue - 64);
    }
    lq_top = lq_base;
    drain_exact(keep);
  };

  // Full units.  The next unit's loads are in flight while this one is processed, and the units kL2Ahead strides ahead are pulled
  // into L2 (one 128-byte line per lane 0..15): one unit per warp in registers is only 32 KB in flight per SM.
  const unsigned n_full = n_points / (unsigned)kUnitPoints;
  const unsigned w_stride = gridDim.x * kWarps;
  const unsigned kL2Ahead = (unsigned)p.l2_ahead;   // 0: no L2 prefetch
  unsigned unit = blockIdx.x * kWarps + warp;
  const float4* src = reinterpret_cast<const float4*>(in) + ((size_t)unit * 128u + lane);
  const size_t src_stride = (size_t)w_stride * 128u;
  const char* l2p = reinterpret_cast<const char*>(in) + ((size_t)unit * 2048u + (lane & 15) * 128u);
  const bool l2_lane = lane < 16 && kL2Ahead > 0;
  for (unsigned a = 1; a < kL2Ahead; ++a)
    if (l2_lane && unit + a * w_stride < n_full) asm volatile("prefetch.global.L2 [%0];" ::"l"(l2p + a * (src_stride * 16u)));
  // queue a point whose code is kCodeSlow: three predicated instructions
  auto push = [&](int code, unsigned index) {
    if (code == kCodeSlow) {
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(lq_top), "r"(index) : "memory");
      lq_top += 128u;
    }
  };
  // one unit: filter its 8 points per lane, store the codes, count, queue the undecided ones
  auto process = [&](const float4 (&v)[4], unsigned u) {
    int c[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      c[2 * j] = filter(v[j].x, v[j].y);
      c[2 * j + 1] = filter(v[j].z, v[j].w);
    }
    const unsigned first = u * (unsigned)kUnitPoints + 2u * lane;   // point index of c[j]: first + 64 (j >> 1) + (j & 1)
    if (out_idx) {
      int2* dst = reinterpret_cast<int2*>(out_idx + first);
#pragma unroll
      for (int j = 0; j < 4; ++j) dst[32 * j] = make_int2(c[2 * j], c[2 * j + 1]);
    }
    if (do_hist) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atoms_inc(hist_base + 4u * (unsigned)c[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) push(c[j], first + 64u * (j >> 1) + (j & 1));
    // a lane queue more than half full could overflow in the next unit
    if (__any_sync(0xffffffffu, lq_top - lq_base > (unsigned)(kLaneSlots - 8) * 128u)) drain(kExactQueue - 64);
  };
  auto load = [&](float4 (&v)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = ldg_stream(src + 32 * j);
  };
  auto advance = [&]() {   // to this warp's next unit: issue its L2 prefetch kL2Ahead units ahead
    unit += w_stride;
    src += src_stride;
    l2p += src_stride * 16u;
    if (l2_lane && unit + (kL2Ahead - 1) * w_stride < n_full)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(l2p + (kL2Ahead - 1) * (src_stride * 16u)));
  };
  // two register sets in ping-pong (no register-to-register copies): while one unit is processed the next one's loads are in flight
  float4 va[4], vb[4];
  if (unit < n_full) load(va);
  while (unit < n_full) {
    const unsigned ua = unit;
    advance();
    if (unit < n_full) load(vb);
    process(va, ua);
    if (unit >= n_full) break;
    const unsigned ub = unit;
    advance();
    if (unit < n_full) load(va);
    process(vb, ub);
  }
  // ragged tail (< 256 points): the warp that would own unit n_full, same point-to-lane map, out-of-range points skipped
  if (unit == n_full && (n_points % (unsigned)kUnitPoints)) {
    const unsigned first = n_full * (unsigned)kUnitPoints + 2u * lane;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const unsigned i = first + 64u * (j >> 1) + (j & 1);
      if (i < n_points) {
        const int code = filter(in[2ull * i], in[2ull * i + 1]);
        if (out_idx) out_idx[i] = code;
        if (do_hist) atoms_inc(hist_base + 4u * (unsigned)code);
        push(code, i);
      }
    }
  }
  drain(0);

  if (do_hist) {
    __syncthreads();
    // classified counts go to their bins; bin Z receives (points handled) - (classified), summed over CTAs
    if (threadIdx.x < 64 && threadIdx.x < p.Z) {
      unsigned c = 0;
      for (int w = 0; w < kWarps; ++w) c += s_whist[w * kHistBins + threadIdx.x + 2];
      if (c) {
        atomicAdd(p.hist + threadIdx.x, (int)c);
        atomicSub(p.hist + p.Z, (int)c);
      }
    }
    if (threadIdx.x == 64) {   // points inside k zones were subtracted k times above
      unsigned e = 0;
      for (int w = 0; w < kWarps; ++w) e += s_extra[w];
      if (e) atomicAdd(p.hist + p.Z, (int)e);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.hist + p.Z, (int)p.N);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Histogram-only kernel (Aggregator.get_zone_counts on classified points).
// ---------------------------------------------------------------------------------------------------------
__global__ void zone_histogram_kernel(const int32_t* zone_idx, const uint64_t* zone_mask, const int32_t* slot,
                                      long long N, int Z, int T, int32_t* hist) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    const int s = slot ? slot[i] : 0;
    if ((unsigned)s >= (unsigned)T) continue;
    int32_t* row = hist + (long long)s * (Z + 1);
    if (zone_idx) {
      const int z = zone_idx[i];
      atomicAdd(row + ((unsigned)z < (unsigned)Z ? z : Z), 1);
    } else {
      uint64_t m = zone_mask[i];
      if (Z < 64) m &= (1ull << Z) - 1;
      if (m == 0) atomicAdd(row + Z, 1);
      while (m) {
        atomicAdd(row + (__ffsll((long long)m) - 1), 1);
        m &= m - 1;
      }
    }
  }
}

// More than 64 zones (zone_classifier.py:44-112 sets no limit): the zones are split into groups of <= 64, one table per group; every
// group answers with its own winner, this kernel keeps the best of them by the global rank (priority or +inf, declaration order):
// idx_groups [G, N] group-local zone index or -1, rank [Z] -> out_idx [N] global zone index or -1, out_count [N] (optional) number of
// GROUPS with a hit (0 = unclassified).
__global__ void zone_combine_kernel(const int32_t* __restrict__ idx_groups, const int32_t* __restrict__ rank, int G, long long N,
                                    int32_t* __restrict__ out_idx, int32_t* __restrict__ out_count) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
    int best = -1, best_rank = 0x7fffffff, hits = 0;
    for (int g = 0; g < G; ++g) {
      const int z = idx_groups[(long long)g * N + i];
      if (z >= 0) {
        ++hits;
        const int gz = g * OPD_MAX_ZONES + z;
        const int r = rank[gz];
        if (r < best_rank) {
          best_rank = r;
          best = gz;
        }
      }
    }
    out_idx[i] = best;
    if (out_count) out_count[i] = hits;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Host: zone table construction (float64).
// ---------------------------------------------------------------------------------------------------------
bool segment_hits_box(double ax, double ay, double bx, double by, double x0, double y0, double x1, double y1) {
  // Liang–Barsky clip of segment a->b against [x0,x1]x[y0,y1]; callers pass an already dilated box.
  double t0 = 0.0, t1 = 1.0;
  const double dx = bx - ax, dy = by - ay;
  const double pq[4][2] = {{-dx, ax - x0}, {dx, x1 - ax}, {-dy, ay - y0}, {dy, y1 - ay}};
  for (int k = 0; k < 4; ++k) {
    const double pp = pq[k][0], qq = pq[k][1];
    if (pp == 0.0) {
      if (qq < 0.0) return false;
    } else {
      const double r = qq / pp;
      if (pp < 0.0) {
        if (r > t1) return false;
        if (r > t0) t0 = r;
      } else {
        if (r < t0) return false;
        if (r < t1) t1 = r;
      }
    }
  }
  return true;
}

}  // namespace

struct opd_zone_table {
  int device = 0, Z = 0, allow_overlap = 0;
  int gw = 1, gh = 1, cells_rounded = 16;
  double gx0 = 0, gy0 = 0, inv_cw = 1, inv_ch = 1, delta = 0.125;
  double margin = 1.0;   // the grid box extends at least this far beyond every polygon's bounding box
  int n_classes = 1, n_boundary = 0, n_verts = 0;
  bool fast_ok = false;
  unsigned char* d_blob = nullptr;  // one allocation, carved below
  uint8_t* d_grid = nullptr;
  uint64_t* d_class_mask = nullptr;
  int32_t* d_class_winner = nullptr;
  int32_t* d_cell_entry = nullptr;
  uint64_t* d_entry_inside = nullptr;
  uint64_t* d_entry_cand = nullptr;
  double2* d_verts = nullptr;
  int32_t* d_poly_off = nullptr;
  int32_t* d_zone_rank = nullptr;
  uint8_t* d_wgrid = nullptr;     // winner grid (see kByteNone)
  float4* d_line_types = nullptr; // [n_line_types] {a, b, c, codes}: a x + b y + c > 0 -> code A (low byte) else code B (second byte)
  int n_line_types = 0, n_line_cells = 0;
  int tiles_x = 0, n_tiles = 0;   // tiles_x > 0: tiled type tables
  uint16_t* d_tile_base = nullptr;
};

extern "C" int opd_zone_table_create(const double* verts_xy, const int32_t* poly_offsets, const double* priority,
                                     int32_t Z, int32_t allow_overlap, int32_t device, opd_zone_table** out) {
  OPD_REQUIRE(out != nullptr, "opd_zone_table_create: out is NULL");
  OPD_REQUIRE(Z >= 0 && Z <= OPD_MAX_ZONES, "opd_zone_table_create: Z=%d outside [0,%d]", Z, OPD_MAX_ZONES);
  OPD_REQUIRE(Z == 0 || (verts_xy && poly_offsets && priority), "opd_zone_table_create: NULL table input");
  const int n_verts = Z ? poly_offsets[Z] : 0;
  for (int z = 0; z < Z; ++z)
    OPD_REQUIRE(poly_offsets[z + 1] - poly_offsets[z] >= 3, "opd_zone_table_create: polygon %d has < 3 vertices", z);
  for (int i = 0; i < 2 * n_verts; ++i)
    OPD_REQUIRE(std::isfinite(verts_xy[i]), "opd_zone_table_create: non-finite vertex coordinate");

  auto* zt = new opd_zone_table();
  zt->device = device;
  zt->Z = Z;
  zt->allow_overlap = allow_overlap ? 1 : 0;
  zt->n_verts = n_verts;

  // rank = position in the order sorted by (priority or +inf, declaration order)  (zone_classifier.py:138-146)
  std::vector<int32_t> rank(64, 0x7fffffff);
  {
    std::vector<int> order(Z);
    for (int z = 0; z < Z; ++z) order[z] = z;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      const double pa = std::isnan(priority[a]) ? INFINITY : priority[a];
      const double pb = std::isnan(priority[b]) ? INFINITY : priority[b];
      return pa < pb;
    });
    for (int r = 0; r < Z; ++r) rank[order[r]] = r;
  }
  auto winner = [&](uint64_t m) {
    int best = -1, br = 0x7fffffff;
    for (int z = 0; z < Z; ++z)
      if ((m >> z) & 1ull)
        if (rank[z] < br) {
          br = rank[z];
          best = z;
        }
    return best;
  };

  const double2* V = reinterpret_cast<const double2*>(verts_xy);
  std::vector<uint8_t> grid;
  std::vector<uint64_t> class_mask(256, 0);
  std::vector<int32_t> class_winner(256, -1);
  std::vector<int32_t> cell_entry;
  std::vector<uint64_t> entry_inside, entry_cand;
  std::vector<uint8_t> wgrid;
  std::vector<float4> line_types;
  std::vector<uint16_t> tile_base;   // tiled type tables: first record of every tile (empty: one table for the whole grid)

  if (Z == 0) {
    zt->gw = zt->gh = 1;
    grid.assign(16, 0);
    wgrid.assign(16, (uint8_t)kByteNone);
    cell_entry.assign(1, -1);
    zt->n_classes = 1;  // class 0 = no zone
    zt->gx0 = zt->gy0 = 0.0;
    zt->inv_cw = zt->inv_ch = 1.0;
    zt->fast_ok = true;
  } else {
    double minx = INFINITY, miny = INFINITY, maxx = -INFINITY, maxy = -INFINITY;
    for (int i = 0; i < n_verts; ++i) {
      minx = std::min(minx, V[i].x); maxx = std::max(maxx, V[i].x);
      miny = std::min(miny, V[i].y); maxy = std::max(maxy, V[i].y);
    }
    const double mx = std::max(1.0, 0.01 * (maxx - minx)), my = std::max(1.0, 0.01 * (maxy - miny));
    minx -= mx; maxx += mx; miny -= my; maxy += my;
    zt->margin = std::min(mx, my) * (1.0 - 1e-9);
    const double ex = maxx - minx, ey = maxy - miny;
    int gw = (int)std::floor(std::sqrt((double)kGridCellBudget * ex / ey));
    gw = std::max(1, std::min(gw, kGridCellBudget));
    int gh = std::max(1, kGridCellBudget / gw);
    const double cw = ex / gw, ch = ey / gh;
    zt->gw = gw; zt->gh = gh;
    zt->gx0 = minx; zt->gy0 = miny;
    zt->inv_cw = 1.0 / cw; zt->inv_ch = 1.0 / ch;
    zt->delta = std::min(0.125, 0.125 * std::min(cw, ch));
    // float32 cell-index slop must stay far below delta for the filter path to be sound
    const double maxabs = std::max(std::max(std::fabs(minx), std::fabs(maxx)), std::max(std::fabs(miny), std::fabs(maxy)));
    zt->fast_ok = (8.0 * 5.97e-8 * maxabs) < zt->delta * 0.25 && zt->delta * 0.5 < std::min(mx, my);

    const size_t cells = (size_t)gw * gh;
    std::vector<uint64_t> cand(cells, 0), inside(cells, 0);
    const double dil = zt->delta * (1.0 + 1e-9) + 1e-9 * (1.0 + maxabs);
    // Single-line cells: a boundary cell whose dilated box is crossed by segments of ONE line only, with no polygon vertex inside,
    // is split by that line into two uniform half cells - the float32 path can answer from the sign of the line function unless
    // the point is within the error band of the line (floor_fast_kernel, tier 1).  cell_line: -1 no edge yet, -2 several lines.
    struct Line { double a, b, c; };
    std::vector<Line> lines;
    std::vector<int32_t> cell_line(cells, -1);
    std::vector<uint8_t> cell_vertex(cells, 0);
    auto line_of = [&](const double2& u, const double2& v) -> int {
      const double dx = v.x - u.x, dy = v.y - u.y, len = std::sqrt(dx * dx + dy * dy);
      if (!(len > 0.0)) return -2;
      double a = dy / len, b = -dx / len;
      if (a < 0.0 || (a == 0.0 && b < 0.0)) { a = -a; b = -b; }
      const double c = -(a * u.x + b * u.y);
      for (size_t i = 0; i < lines.size(); ++i)   // coincident edges of neighbouring polygons are one line
        if (std::fabs(lines[i].a - a) < 1e-9 && std::fabs(lines[i].b - b) < 1e-9 && std::fabs(lines[i].c - c) < 1e-6) return (int)i;
      lines.push_back({a, b, c});
      return (int)lines.size() - 1;
    };
    for (int z = 0; z < Z; ++z) {
      const int o = poly_offsets[z], n = poly_offsets[z + 1] - o;
      double bx0 = INFINITY, by0 = INFINITY, bx1 = -INFINITY, by1 = -INFINITY;
      for (int i = 0; i < n; ++i) {
        const double2 a = V[o + i], b = V[o + (i + 1) % n];
        bx0 = std::min(bx0, a.x); bx1 = std::max(bx1, a.x);
        by0 = std::min(by0, a.y); by1 = std::max(by1, a.y);
        const int lid = line_of(a, b);
        {   // cells whose dilated box contains the vertex a
          const int vx0 = std::max(0, (int)std::floor((a.x - dil - minx) / cw) - 1), vx1 = std::min(gw - 1, (int)std::floor((a.x + dil - minx) / cw) + 1);
          const int vy0 = std::max(0, (int)std::floor((a.y - dil - miny) / ch) - 1), vy1 = std::min(gh - 1, (int)std::floor((a.y + dil - miny) / ch) + 1);
          for (int cy = vy0; cy <= vy1; ++cy)
            for (int cx = vx0; cx <= vx1; ++cx)
              if (a.x >= minx + cx * cw - dil && a.x <= minx + (cx + 1) * cw + dil && a.y >= miny + cy * ch - dil && a.y <= miny + (cy + 1) * ch + dil)
                cell_vertex[(size_t)cy * gw + cx] = 1;
        }
        // cells whose dilated box the edge crosses
        const double ex0 = std::min(a.x, b.x) - dil, ex1 = std::max(a.x, b.x) + dil;
        const double ey0 = std::min(a.y, b.y) - dil, ey1 = std::max(a.y, b.y) + dil;
        const int cx0 = std::max(0, (int)std::floor((ex0 - minx) / cw) - 1), cx1 = std::min(gw - 1, (int)std::floor((ex1 - minx) / cw) + 1);
        const int cy0 = std::max(0, (int)std::floor((ey0 - miny) / ch) - 1), cy1 = std::min(gh - 1, (int)std::floor((ey1 - miny) / ch) + 1);
        for (int cy = cy0; cy <= cy1; ++cy)
          for (int cx = cx0; cx <= cx1; ++cx) {
            const double x0 = minx + cx * cw - dil, x1 = minx + (cx + 1) * cw + dil;
            const double y0 = miny + cy * ch - dil, y1 = miny + (cy + 1) * ch + dil;
            if (segment_hits_box(a.x, a.y, b.x, b.y, x0, y0, x1, y1)) {
              const size_t c = (size_t)cy * gw + cx;
              cand[c] |= 1ull << z;
              cell_line[c] = (cell_line[c] == -1 || cell_line[c] == lid) ? lid : -2;
            }
          }
      }
      // uniform status of the remaining cells inside the polygon's bounding box: test the cell centre
      const int cx0 = std::max(0, (int)std::floor((bx0 - minx) / cw) - 1), cx1 = std::min(gw - 1, (int)std::floor((bx1 - minx) / cw) + 1);
      const int cy0 = std::max(0, (int)std::floor((by0 - miny) / ch) - 1), cy1 = std::min(gh - 1, (int)std::floor((by1 - miny) / ch) + 1);
      for (int cy = cy0; cy <= cy1; ++cy)
        for (int cx = cx0; cx <= cx1; ++cx) {
          const size_t c = (size_t)cy * gw + cx;
          if ((cand[c] >> z) & 1ull) continue;
          if (point_in_polygon_ref(minx + (cx + 0.5) * cw, miny + (cy + 0.5) * ch, V + o, n)) inside[c] |= 1ull << z;
        }
    }
    // classes
    zt->cells_rounded = (int)((cells + 15) / 16 * 16);
    grid.assign(zt->cells_rounded, 0);
    cell_entry.assign(cells, -1);
    std::map<uint64_t, int> class_of;
    int n_classes = 0;
    auto get_class = [&](uint64_t m) -> int {
      const int w = winner(m);
      const uint64_t key = zt->allow_overlap ? m : (w >= 0 ? 1ull << w : 0ull);
      auto it = class_of.find(key);
      if (it != class_of.end()) return it->second;
      if (n_classes >= kMaxClasses) return -1;
      class_mask[n_classes] = key;
      class_winner[n_classes] = w;
      class_of[key] = n_classes;
      return n_classes++;
    };
    get_class(0);  // class 0 = no zone
    for (size_t c = 0; c < cells; ++c) {
      int code = cand[c] ? -1 : get_class(inside[c]);
      if (code < 0) {
        cell_entry[c] = (int32_t)entry_inside.size();
        entry_inside.push_back(inside[c]);
        entry_cand.push_back(cand[c]);
        code = kBoundary;
      }
      grid[c] = (uint8_t)code;
    }
    zt->n_classes = n_classes;
    zt->n_boundary = (int)entry_inside.size();

    // winner grid of the float32 path
    wgrid.assign(zt->cells_rounded, (uint8_t)kByteNone);
    auto code_of_mask = [&](uint64_t m) -> int {   // what the float32 path may answer for a region inside exactly these zones
      if (m == 0) return kCodeNone;
      if (zt->allow_overlap && (m & (m - 1))) return kCodeSlow;   // counted once per zone: float64 path
      return winner(m);
    };
    const double band = 2.0 * zt->delta;   // a half cell must reach at least this far from the line to be answered by sign
    struct TypeKey { int line, a, b; bool operator<(const TypeKey& o) const { return line != o.line ? line < o.line : (a != o.a ? a < o.a : b < o.b); } };
    std::map<TypeKey, std::vector<size_t>> type_cells;
    for (size_t c = 0; c < cells; ++c) {
      if (!cand[c]) {
        wgrid[c] = (uint8_t)code_of_mask(inside[c]);   // kCodeNone / kCodeSlow wrap to kByteNone / kByteSlow
        continue;
      }
      wgrid[c] = (uint8_t)kByteSlow;
      if (cell_vertex[c] || cell_line[c] < 0) continue;
      const Line& L = lines[cell_line[c]];
      const int cx = (int)(c % gw), cy = (int)(c / gw);
      double best[2] = {0.0, 0.0};
      double2 probe[2] = {{0, 0}, {0, 0}};
      for (int k = 0; k < 4; ++k) {   // the cell corner farthest from the line on either side
        const double x = minx + (cx + (k & 1)) * cw, y = miny + (cy + (k >> 1)) * ch, d = L.a * x + L.b * y + L.c;
        if (d > best[0]) { best[0] = d; probe[0] = {x, y}; }
        if (-d > best[1]) { best[1] = -d; probe[1] = {x, y}; }
      }
      int side[2];
      for (int sd = 0; sd < 2; ++sd) {
        if (best[sd] < band) { side[sd] = kCodeSlow; continue; }   // a sliver inside the band: the float64 path answers it anyway
        uint64_t m = inside[c];
        for (int z = 0; z < Z; ++z)
          if ((cand[c] >> z) & 1ull)
            if (point_in_polygon_ref(probe[sd].x, probe[sd].y, V + poly_offsets[z], poly_offsets[z + 1] - poly_offsets[z])) m |= 1ull << z;
        side[sd] = code_of_mask(m);
      }
      if (side[0] == kCodeSlow && side[1] == kCodeSlow) continue;
      type_cells[{cell_line[c], side[0], side[1]}].push_back(c);
    }
    // the kMaxLineTypes most frequent (line, code A, code B) triples get a code; the rest stay slow cells
    std::vector<std::pair<size_t, TypeKey>> order;
    for (auto& kv : type_cells) order.push_back({kv.second.size(), kv.first});
    std::sort(order.begin(), order.end(), [](const auto& x, const auto& y) { return x.first > y.first; });
    const bool lines_ok = zt->delta >= 0.01;   // the band argument needs delta well above the float32 error of the line function
    auto record_of = [&](const TypeKey& k) {
      const Line& L = lines[k.line];
      const uint32_t codes = ((uint32_t)k.a & 255u) | (((uint32_t)k.b & 255u) << 8);
      float cf;
      memcpy(&cf, &codes, 4);
      return make_float4((float)L.a, (float)L.b, (float)L.c, cf);
    };
    const size_t cap = (size_t)(254 - Z);
    if (lines_ok && order.size() <= cap) {
      // one table for the whole grid: byte Z + t = type t
      for (size_t t = 0; t < order.size(); ++t) {
        const TypeKey& k = order[t].second;
        line_types.push_back(record_of(k));
        for (size_t c : type_cells[k]) wgrid[c] = (uint8_t)(Z + (int)t);
        zt->n_line_cells += (int)type_cells[k].size();
      }
    } else if (lines_ok) {
      // more (line, code A, code B) triples than a byte can name (e.g. 64 many-sided polygons): one table per TILE of
      // kTypeTile x kTypeTile cells; byte Z + t = type t of the cell's tile, record tile_base[tile] + t.  The most frequent triples
      // of every tile first; a tile keeps at most 254 - Z types and the grid at most kMaxTiledRecords records (the rest: slow cells).
      zt->tiles_x = (gw + kTypeTile - 1) / kTypeTile;
      const int tiles_y = (gh + kTypeTile - 1) / kTypeTile;
      std::vector<std::map<TypeKey, std::vector<size_t>>> per_tile((size_t)zt->tiles_x * tiles_y);
      for (auto& kv : type_cells)
        for (size_t c : kv.second) per_tile[(c / gw / kTypeTile) * zt->tiles_x + (c % gw) / kTypeTile][kv.first].push_back(c);
      tile_base.assign(per_tile.size() + 1, 0);
      for (size_t ti = 0; ti < per_tile.size(); ++ti) {
        tile_base[ti] = (uint16_t)line_types.size();
        std::vector<std::pair<size_t, TypeKey>> ord;
        for (auto& kv : per_tile[ti]) ord.push_back({kv.second.size(), kv.first});
        std::sort(ord.begin(), ord.end(), [](const auto& x, const auto& y) { return x.first > y.first; });
        for (size_t t = 0; t < ord.size() && t < cap && line_types.size() < (size_t)kMaxTiledRecords; ++t) {
          line_types.push_back(record_of(ord[t].second));
          for (size_t c : per_tile[ti][ord[t].second]) wgrid[c] = (uint8_t)(Z + (int)t);
          zt->n_line_cells += (int)per_tile[ti][ord[t].second].size();
        }
      }
    }
    zt->n_line_types = (int)line_types.size();
  }
  if (entry_inside.empty()) {
    entry_inside.push_back(0);
    entry_cand.push_back(0);
  }
  std::vector<int32_t> poly_off(68, 0);
  for (int z = 0; z <= Z && Z > 0; ++z) poly_off[z] = poly_offsets[z];

  // one device blob
  auto up16 = [](size_t v) { return (v + 15) / 16 * 16; };
  const size_t o_grid = 0;
  const size_t o_cm = o_grid + up16(grid.size());
  const size_t o_cw = o_cm + 256 * 8;
  const size_t o_ce = o_cw + 256 * 4;
  const size_t o_ei = up16(o_ce + cell_entry.size() * 4);
  const size_t o_ec = o_ei + up16(entry_inside.size() * 8);
  const size_t o_v = o_ec + up16(entry_cand.size() * 8);
  const size_t o_po = o_v + up16((size_t)std::max(1, n_verts) * 16);
  const size_t o_rk = o_po + up16(68 * 4);
  const size_t o_wg = o_rk + up16(64 * 4);
  const size_t o_lt = o_wg + up16(wgrid.size());
  const size_t o_tb = o_lt + up16(std::max<size_t>(kMaxLineRecords, line_types.size()) * 16);
  const size_t total = o_tb + up16((tile_base.size() + 8) * 2) + 16;
  std::vector<unsigned char> blob(total, 0);
  memcpy(&blob[o_grid], grid.data(), grid.size());
  memcpy(&blob[o_cm], class_mask.data(), 256 * 8);
  memcpy(&blob[o_cw], class_winner.data(), 256 * 4);
  memcpy(&blob[o_ce], cell_entry.data(), cell_entry.size() * 4);
  memcpy(&blob[o_ei], entry_inside.data(), entry_inside.size() * 8);
  memcpy(&blob[o_ec], entry_cand.data(), entry_cand.size() * 8);
  if (n_verts) memcpy(&blob[o_v], verts_xy, (size_t)n_verts * 16);
  memcpy(&blob[o_po], poly_off.data(), 68 * 4);
  memcpy(&blob[o_rk], rank.data(), 64 * 4);
  memcpy(&blob[o_wg], wgrid.data(), wgrid.size());
  if (!line_types.empty()) memcpy(&blob[o_lt], line_types.data(), line_types.size() * 16);
  if (!tile_base.empty()) memcpy(&blob[o_tb], tile_base.data(), tile_base.size() * 2);

  opd::DeviceGuard guard(device);
  cudaError_t e = guard.err;
  if (e == cudaSuccess) e = cudaMalloc(&zt->d_blob, total);
  if (e == cudaSuccess) e = cudaMemcpy(zt->d_blob, blob.data(), total, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (zt->d_blob) cudaFree(zt->d_blob);
    delete zt;
    return opd::fail(OPD_ERR_CUDA, "opd_zone_table_create: %s", cudaGetErrorString(e));
  }
  zt->d_grid = zt->d_blob + o_grid;
  zt->d_class_mask = reinterpret_cast<uint64_t*>(zt->d_blob + o_cm);
  zt->d_class_winner = reinterpret_cast<int32_t*>(zt->d_blob + o_cw);
  zt->d_cell_entry = reinterpret_cast<int32_t*>(zt->d_blob + o_ce);
  zt->d_entry_inside = reinterpret_cast<uint64_t*>(zt->d_blob + o_ei);
  zt->d_entry_cand = reinterpret_cast<uint64_t*>(zt->d_blob + o_ec);
  zt->d_verts = reinterpret_cast<double2*>(zt->d_blob + o_v);
  zt->d_poly_off = reinterpret_cast<int32_t*>(zt->d_blob + o_po);
  zt->d_zone_rank = reinterpret_cast<int32_t*>(zt->d_blob + o_rk);
  zt->d_wgrid = zt->d_blob + o_wg;
  zt->d_line_types = reinterpret_cast<float4*>(zt->d_blob + o_lt);
  zt->d_tile_base = reinterpret_cast<uint16_t*>(zt->d_blob + o_tb);
  if (zt->tiles_x > 0 && (int)tile_base.size() > kMaxTiles) {   // absurdly large grids: no line types at all
    zt->tiles_x = 0;
    zt->n_line_types = 0;
  }
  zt->n_tiles = zt->tiles_x > 0 ? (int)tile_base.size() : 0;
  *out = zt;
  return OPD_OK;
}

extern "C" void opd_zone_table_destroy(opd_zone_table* zt) {
  if (!zt) return;
  if (zt->d_blob) cudaFree(zt->d_blob);
  delete zt;
}

extern "C" int opd_zone_table_info(const opd_zone_table* zt, int32_t* grid_w, int32_t* grid_h,
                                   int32_t* n_boundary_cells, int32_t* n_classes) {
  OPD_REQUIRE(zt != nullptr, "opd_zone_table_info: NULL table");
  if (grid_w) *grid_w = zt->gw;
  if (grid_h) *grid_h = zt->gh;
  if (n_boundary_cells) *n_boundary_cells = zt->n_boundary;
  if (n_classes) *n_classes = zt->n_classes;
  return OPD_OK;
}

namespace {

int fill_params(FloorK& k, const opd_floor_params* p, const opd_zone_table* zt, const void* in, const int32_t* slot,
                int64_t N, int32_t T, void* floor_px, void* floor_mm, uint8_t* in_bounds, int32_t* zone_idx,
                uint64_t* zone_mask, int32_t* hist) {
  OPD_REQUIRE(p && zt, "floor: NULL params / zone table");
  OPD_REQUIRE(N >= 0, "floor: N=%lld < 0", (long long)N);
  OPD_REQUIRE(N == 0 || in != nullptr, "floor: NULL input");
  OPD_REQUIRE(hist == nullptr || T >= 1, "floor: T=%d must be >= 1 when a histogram is requested", T);
  memset(&k, 0, sizeof(k));
  for (int i = 0; i < 9; ++i) {
    k.H[i] = p->H[i];
    k.Hf[i] = (float)p->H[i];
  }
  for (int j = 0; j < 3; ++j) {
    // magnitudes rounded up so the float32 bound stays an upper bound
    k.Sxy[j] = nextafterf((float)std::max(std::fabs(p->H[j]), std::fabs(p->H[3 + j])), INFINITY);
    k.Sw[j] = nextafterf((float)std::fabs(p->H[6 + j]), INFINITY);
  }
  k.sx = p->scale_x_mm; k.sy = p->scale_y_mm; k.mw = p->map_w_px; k.mh = p->map_h_px;
  k.err_max = (float)(zt->delta * 0.5);
  k.gx0 = zt->gx0; k.gy0 = zt->gy0; k.inv_cw = zt->inv_cw; k.inv_ch = zt->inv_ch;
  k.gx0_f = (float)zt->gx0; k.gy0_f = (float)zt->gy0; k.inv_cw_f = (float)zt->inv_cw; k.inv_ch_f = (float)zt->inv_ch;
  k.gw = zt->gw; k.gh = zt->gh;
  k.Z = zt->Z; k.allow_overlap = zt->allow_overlap;
  k.input_is_bbox = p->input_is_bbox; k.skip_projection = p->skip_projection;
  k.n_verts = zt->n_verts;
  k.grid = zt->d_grid; k.class_mask = zt->d_class_mask; k.class_winner = zt->d_class_winner;
  k.cell_entry = zt->d_cell_entry; k.entry_inside = zt->d_entry_inside; k.entry_cand = zt->d_entry_cand;
  k.verts = zt->d_verts; k.poly_off = zt->d_poly_off; k.zone_rank = zt->d_zone_rank;
  k.wgrid = zt->d_wgrid; k.line_types = zt->d_line_types; k.n_line_types = zt->n_line_types;
  k.tile_base = zt->d_tile_base; k.tiles_x = zt->tiles_x; k.n_tiles = zt->n_tiles;
  // |p̂ - p| <= Δ/2 per coordinate (q <= T1) -> <= Δ/sqrt(2) along the unit normal; + float32 rounding of a, b, c and of the two
  // multiply-adds (<= 1e-3 px for |p| <= 4000) + the 1e-6 px tolerance under which coincident edges were merged into one line
  k.line_band = (float)(0.75 * zt->delta + 1.5e-3);
  k.l2_ahead = 4;
  if (const int pr = opd::g_option_probe.load(); pr >= 100 && pr < 200) k.l2_ahead = pr - 100;   // measurement: opd_set_option("probe", 100 + n)
  k.in = in; k.slot = slot; k.N = N; k.T = T;
  k.floor_px = floor_px; k.floor_mm = floor_mm; k.in_bounds = in_bounds;
  k.zone_idx = zone_idx; k.zone_mask = zone_mask; k.hist = hist;
  k.stage_verts = zt->n_verts <= kMaxSmemVerts;
  return OPD_OK;
}

int device_sm_count(int device, int* sms) {
  static std::mutex mu;
  static int cached[opd::kMaxDevices] = {0};
  std::lock_guard<std::mutex> lk(mu);
  if (device < 0 || device >= opd::kMaxDevices) return opd::fail(OPD_ERR_INVALID, "floor: device %d out of range", device);
  if (cached[device] <= 0) OPD_CUDA_OK(cudaDeviceGetAttribute(&cached[device], cudaDevAttrMultiProcessorCount, device));
  *sms = cached[device];
  return OPD_OK;
}

template <typename T>
int launch_exact(FloorK& k, const opd_zone_table* zt, cudaStream_t s) {
  if (k.N == 0) return OPD_OK;
  int sms = 148;
  if (int rc = device_sm_count(zt->device, &sms)) return rc;
  // small inputs read the grid from L2; large ones stage it in shared memory once per CTA
  k.stage_grid = k.N >= (1 << 18);
  const size_t smem = tables_smem_bytes(k.stage_grid, zt->cells_rounded, k.stage_verts, k.n_verts);
  auto kern = floor_exact_kernel<T>;
  OPD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = 256;
  long long blocks = (k.N + threads - 1) / threads;
  const long long cap = k.stage_grid ? sms : (long long)sms * 8;
  if (blocks > cap) blocks = cap;
  kern<<<(unsigned)blocks, threads, smem, s>>>(k, zt->cells_rounded);
  opd::count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

}  // namespace

extern "C" int opd_floor_project_classify_count_f64(const opd_floor_params* p, const opd_zone_table* zt,
                                                    const double* in_dev, const int32_t* slot_dev, int64_t N,
                                                    int32_t T, double* floor_px_dev, double* floor_mm_dev,
                                                    uint8_t* in_bounds_dev, int32_t* zone_idx_dev,
                                                    uint64_t* zone_mask_dev, int32_t* hist_dev, void* stream) {
  FloorK k;
  if (int rc = fill_params(k, p, zt, in_dev, slot_dev, N, T, floor_px_dev, floor_mm_dev, in_bounds_dev, zone_idx_dev,
                           zone_mask_dev, hist_dev))
    return rc;
  return launch_exact<double>(k, zt, static_cast<cudaStream_t>(stream));
}

extern "C" int opd_floor_project_classify_count_f32(const opd_floor_params* p, const opd_zone_table* zt,
                                                    const float* in_dev, const int32_t* slot_dev, int64_t N,
                                                    int32_t T, float* floor_px_dev, float* floor_mm_dev,
                                                    uint8_t* in_bounds_dev, int32_t* zone_idx_dev,
                                                    uint64_t* zone_mask_dev, int32_t* hist_dev, void* stream) {
  FloorK k;
  if (int rc = fill_params(k, p, zt, in_dev, slot_dev, N, T, floor_px_dev, floor_mm_dev, in_bounds_dev, zone_idx_dev,
                           zone_mask_dev, hist_dev))
    return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // The filtered kernel covers the bandwidth-critical shape: many points, zone index and/or a single
  // histogram row.  Everything else (coordinates out, masks, per-timestamp slots, boxes) takes the exact kernel.
  const bool fast = zt->fast_ok && N >= (1 << 16) && N < (1ll << 32) && !p->input_is_bbox && !p->skip_projection &&
                    !floor_px_dev && !floor_mm_dev && !in_bounds_dev && !zone_mask_dev && !slot_dev &&
                    (zone_idx_dev || hist_dev) && (reinterpret_cast<uintptr_t>(in_dev) % 16 == 0) &&
                    (zone_idx_dev == nullptr || reinterpret_cast<uintptr_t>(zone_idx_dev) % 16 == 0);
  if (!fast) return launch_exact<float>(k, zt, s);
  // constants of the float32 filter (derivation above floor_fast_kernel), float64 on the host, rounded to the safe side
  const double u32 = 32.0 * 5.9604645e-8;
  const double gx1 = zt->gx0 + zt->gw / zt->inv_cw, gy1 = zt->gy0 + zt->gh / zt->inv_ch;
  const double pmax = 1.0001 * std::max(std::max(std::fabs(zt->gx0), std::fabs(gx1)), std::max(std::fabs(zt->gy0), std::fabs(gy1))) + 1.0;
  const double t1 = zt->delta / u32 - pmax;
  const double t2 = std::min(zt->margin, pmax * 0.99) / u32 - pmax;
  if (!(t1 > 0.0 && t2 > 0.0)) return launch_exact<float>(k, zt, s);
  for (int j = 0; j < 3; ++j) k.fk[j] = nextafterf((float)(((double)k.Sxy[j] + pmax * (double)k.Sw[j]) * (1.0 + 1e-6)), INFINITY);
  k.fT1 = nextafterf((float)(t1 * (1.0 - 1e-6)), 0.0f);
  k.fT2 = nextafterf((float)(t2 * (1.0 - 1e-6)), 0.0f);
  k.fgx = (float)(-zt->gx0 * zt->inv_cw);
  k.fgy = (float)(-zt->gy0 * zt->inv_ch);
  int sms = 148;
  if (int rc = device_sm_count(zt->device, &sms)) return rc;
  k.stage_grid = 1;
  const size_t smem = (size_t)zt->cells_rounded + (tables_smem_bytes(0, zt->cells_rounded, k.stage_verts, k.n_verts) + 15) / 16 * 16 +
                      (zt->tiles_x > 0 ? kMaxTiledRecords * 16 + kMaxTiles * 2 : kMaxLineRecords * 16) +
                      (kFastThreads / 32) * (kLaneSlots * 32 + kExactQueue) * 4 +
                      (kFastThreads / 32) * (kHistBins + 1) * 4 + 16;
  if (smem > 227 * 1024) return launch_exact<float>(k, zt, s);   // cannot happen with the budgets above; never launch past the limit
  auto kern = zt->tiles_x > 0 ? floor_fast_kernel<true> : floor_fast_kernel<false>;
  OPD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long chunks = (N + kFastChunk - 1) / kFastChunk;
  const unsigned blocks = (unsigned)std::min<long long>(chunks, sms);
  kern<<<blocks, kFastThreads, smem, s>>>(k, zt->cells_rounded);
  opd::count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

extern "C" int opd_zone_histogram(const int32_t* zone_idx_dev, const uint64_t* zone_mask_dev, const int32_t* slot_dev,
                                  int64_t N, int32_t Z, int32_t T, int32_t* hist_dev, void* stream) {
  OPD_REQUIRE((zone_idx_dev != nullptr) != (zone_mask_dev != nullptr), "opd_zone_histogram: pass exactly one of zone_idx / zone_mask");
  OPD_REQUIRE(hist_dev && Z >= 0 && T >= 1 && N >= 0, "opd_zone_histogram: bad argument");
  OPD_REQUIRE(zone_idx_dev || Z <= OPD_MAX_ZONES, "opd_zone_histogram: a mask covers at most %d zones", OPD_MAX_ZONES);
  if (N == 0) return OPD_OK;
  const int threads = 256;
  const unsigned blocks = (unsigned)std::min<long long>((N + threads - 1) / threads, 148 * 8);
  zone_histogram_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(zone_idx_dev, zone_mask_dev, slot_dev,
                                                                                  N, Z, T, hist_dev);
  opd::count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

extern "C" int opd_zone_combine_groups(const int32_t* idx_groups_dev, const int32_t* rank_dev, int32_t G, int64_t N,
                                       int32_t* out_idx_dev, int32_t* out_count_dev, void* stream) {
  OPD_REQUIRE(idx_groups_dev && rank_dev && out_idx_dev && G >= 1 && N >= 0, "opd_zone_combine_groups: bad argument");
  if (N == 0) return OPD_OK;
  const int threads = 256;
  const unsigned blocks = (unsigned)std::min<long long>((N + threads - 1) / threads, 148 * 8);
  zone_combine_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(idx_groups_dev, rank_dev, G, N, out_idx_dev, out_count_dev);
  opd::count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}
