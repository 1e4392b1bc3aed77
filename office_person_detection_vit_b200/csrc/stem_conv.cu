// K1 + K2: preprocessing and the ResNet stem (7x7 / stride 2 / pad 3 convolution, 3 -> 64, + frozen BN + ReLU) on the
// tensor cores without an im2col buffer.
//
//   K1  s2d_preprocess_kernel: uint8 frame -> normalised bf16 "space-to-depth" tensor S[B, H/2, W/2, 16]:
//       channel (dy*2+dx)*3 + c = pixel (2y+dy, 2x+dx), colour c (12 real channels, 4 zero lanes; 32 B per pixel).
//       (BGR->RGB and (u8 - 255*mean) / (255*std): transformers image_processing_backends.py:308-331.)
//   K2  stem_kernel: in S space the 7x7/s2 convolution is a 4x4 convolution with pad (2 before, 1 after) and
//       K = 16 taps x 16 channels.  One output tile = 16 rows x 8 columns; its 19 x 11 pixel halo (6.7 KB) is loaded ONCE
//       by a tiled TMA (32-byte swizzle, zero fill = padding) and each of the 16 taps is ONE tcgen05.mma (M 128, N 64,
//       K 16) whose A descriptor points into the patch at (r * 11 + s) * 32 B with 8-pixel groups one patch row (352 B)
//       apart.  The 32 KB of weights stay resident in shared memory.  Two epilogue warpgroups alternate tiles:
//       TMEM -> +shift, ReLU -> bf16 -> swizzled staging -> 4D TMA store into the NHWC stem output.
// Arithmetic replaced: transformers models/resnet/modeling_resnet.py:57-88 (ResNetEmbeddings) with DetrFrozenBatchNorm2d
// (models/detr/modeling_detr.py:185-222) folded into the weights.
#include <algorithm>
#include <cstdio>

#include "detr_kernels.h"
#include "opd_common.h"
#include "sm100_ptx.cuh"
#include "tc_gemm.h"

namespace opd {
namespace {

constexpr int TILE_W = 8, TILE_H = 16, PATCH_W = TILE_W + 3, PATCH_H = TILE_H + 3;
constexpr int PATCH_BYTES = PATCH_W * PATCH_H * 32;   // 6688
constexpr int PATCH_SLOT = 7168;
constexpr int kPatchStages = 12;
constexpr int W_TAP_BYTES = 64 * 32;                   // one tap: 64 output channels x 16 input lanes
constexpr int OUT_BYTES = 128 * 128;                   // staging box: 128 pixels x 64 channels bf16
constexpr int kAccStages = 8;
constexpr int kThreads = 384;
#ifdef OPD_STEM_PROBE
constexpr bool kCounters = true;    // per-role clock64 counters printed by CTA 0 when opd_set_option("probe", 4)
#else
constexpr bool kCounters = false;   // clock64 reads cost ~30 cycles each on the single-thread roles: compiled out
#endif
__device__ __forceinline__ long long probe_clock() { return kCounters ? clock64() : 0; }
constexpr int kSmemBytes = 16 * W_TAP_BYTES + kPatchStages * PATCH_SLOT + 4 * OUT_BYTES + 1024;
// fused pooling: 8 more warps (12-19) reduce the 16 x 16 blocks the epilogue warpgroups leave in a ring of kBlkSlots
constexpr int kPoolThreads = 640, kBlkSlots = 3, BLK_BYTES = 2 * OUT_BYTES;
constexpr int kPoolSmemBytes = 16 * W_TAP_BYTES + kPatchStages * PATCH_SLOT + kBlkSlots * BLK_BYTES + 1024;

struct StemParams {
  CUtensorMap tmS, tmW, tmD;
  int tiles_x, tiles_y, num_tiles;
  const float* bias;
  __nv_bfloat16* pooled;   // kPool: max-pool output [B, P, Q, 64]
  int H2, W2, P, Q;        // stem output size, pooled size
  int probe;   // measurement probes (opd_set_option("probe")): bit 0 no patch reloads, bit 1 no output stores
};

// The input is uint8: (u - 255 * mean[c]) / (255 * std[c]) rounded to bf16 takes 3 x 256 values, each reproduced bit for bit by
// one FMA (see `norm` below).
// valid_hw (optional, [B, 2] int32): frame b holds a picture of valid_hw[b] = (h, w) pixels in the top-left corner of its
// Hs x Ws canvas; the rest is the batch padding and becomes 0 AFTER normalisation, like DetrImageProcessor.pad (pixel_mask = 0).
__global__ void __launch_bounds__(256) s2d_preprocess_kernel(const uint8_t* __restrict__ src, int B, int Hs, int Ws, int bgr,
                                                             __nv_bfloat16* __restrict__ S, int H2, int W2,
                                                             const int32_t* __restrict__ valid_hw) {
  // (u - 255 mean[c]) / (255 std[c]) -> bf16, as fma(u, 1 / s, -m / s): for every one of the 3 x 256 possible inputs the bf16 result
  // has the bits of the reference's subtract-then-divide (enumerated offline and in tests/test_detr_gpu.py; the two float32
  // values differ by at most one ulp, never across a bf16 rounding boundary).  A 3 x 256-entry shared-memory table of those
  // values (round 1) cost ~3.5 bank-conflicted wavefronts per look-up: 80 us of the kernel's 180.
  constexpr float kMean[3] = {0.485f * 255.0f, 0.456f * 255.0f, 0.406f * 255.0f};
  constexpr float kStd[3] = {0.229f * 255.0f, 0.224f * 255.0f, 0.225f * 255.0f};
  constexpr float kInv[3] = {1.0f / kStd[0], 1.0f / kStd[1], 1.0f / kStd[2]};
  constexpr float kOff[3] = {-kMean[0] / kStd[0], -kMean[1] / kStd[1], -kMean[2] / kStd[2]};
  auto norm = [&](int c, uint32_t u) {   // float(u) through the 2^23 trick (LOP3 + FADD instead of an I2F on the XU pipe)
    return fmaf(__uint_as_float(0x4B000000u | u) - 8388608.0f, kInv[c], kOff[c]);
  };
  // Work unit = one row of S (frame b, row y): blocks stride over the B * H2 rows, threads over the row's pixels.  (A flat index
  // cost two 64-bit divisions per pixel: ~300 of the thread's ~400 instructions.)
  const int rows = B * H2;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const int b = r / H2, y = r - b * H2;
    const int vh = valid_hw ? valid_hw[2 * b] : Hs, vw = valid_hw ? valid_hw[2 * b + 1] : Ws;
    const uint8_t* row0 = src + ((long long)b * Hs + 2 * y) * Ws * 3;
    const bool ok0 = 2 * y < vh, ok1 = 2 * y + 1 < vh;
    uint4* dst_row = reinterpret_cast<uint4*>(S + (long long)r * W2 * 16);
    // kPx pixels per thread and trip, every byte load issued before the first table look-up: the kernel is paced by the bytes a
    // thread keeps in flight (12 per pixel), not by instructions
    constexpr int kPx = 3;
    for (int x0 = threadIdx.x; x0 < W2; x0 += kPx * blockDim.x) {
      uint8_t raw[kPx][12];
#pragma unroll
      for (int u = 0; u < kPx; ++u) {
        const int x = x0 + u * blockDim.x;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const int ix = 2 * x + dx;
            const bool ok = x < W2 && (dy ? ok1 : ok0) && ix < vw;
            const uint8_t* px = row0 + (ok ? (long long)dy * Ws * 3 + ix * 3 : 0);
#pragma unroll
            for (int c = 0; c < 3; ++c) raw[u][(dy * 2 + dx) * 3 + c] = px[bgr ? 2 - c : c];
          }
      }
#pragma unroll
      for (int u = 0; u < kPx; ++u) {
        const int x = x0 + u * blockDim.x;
        if (x >= W2) break;
        float v[12];
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const bool ok = (dy ? ok1 : ok0) && 2 * x + dx < vw;
#pragma unroll
            for (int c = 0; c < 3; ++c) v[(dy * 2 + dx) * 3 + c] = ok ? norm(c, raw[u][(dy * 2 + dx) * 3 + c]) : 0.f;
          }
        uint4 o0, o1;
        o0.x = ptx::pack_bf16(v[0], v[1]); o0.y = ptx::pack_bf16(v[2], v[3]); o0.z = ptx::pack_bf16(v[4], v[5]); o0.w = ptx::pack_bf16(v[6], v[7]);
        o1.x = ptx::pack_bf16(v[8], v[9]); o1.y = ptx::pack_bf16(v[10], v[11]); o1.z = 0u; o1.w = 0u;
        dst_row[2 * x] = o0;
        dst_row[2 * x + 1] = o1;
      }
    }
  }
}

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(ptx::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// K-major operand with 32-byte rows and the 32-byte swizzle; sbo = byte distance between 8-row groups
__device__ __forceinline__ uint64_t desc_sw32(uint32_t addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;   // SWIZZLE_32B
  return d;
}

// kPool: the 3x3 / stride 2 / pad 1 max pooling (modeling_resnet.py:69, nn.MaxPool2d) runs in the epilogue.  A work tile is a
// 16 x 16 block of stem outputs = two MMA tiles side by side (columns 0-7 / 8-15, one per epilogue warpgroup) whose origin
// is (14 ty - 1, 14 tx - 1): it holds every input of pooled outputs (7 ty .. 7 ty + 6, 7 tx .. 7 tx + 6).  Both warpgroups
// write bias + ReLU'd bf16 pixels (zeros outside the image, which is what the pooling's -inf padding amounts to after a
// ReLU) into a shared-memory block, and the 256 epilogue threads reduce it to 7 x 7 x 64 pooled values stored straight to
// global memory.  The stem output (2.2 GB per 64 frames) is never written or read back; the price is (16/14)^2 = 1.31x
// the MMA work, on a kernel whose tensor pipe was two thirds idle.
template <bool kPool>
__global__ void __launch_bounds__(kPool ? kPoolThreads : kThreads, 1) stem_kernel(const __grid_constant__ StemParams p) {
  constexpr uint32_t kIdesc = ptx::umma_idesc_bf16(128, 64);
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_w = smem;                                   // 16 taps x [64 x 32 B]
  uint8_t* smem_patch = smem_w + 16 * W_TAP_BYTES;          // [kPatchStages]
  uint8_t* smem_out = smem_patch + kPatchStages * PATCH_SLOT;   // 2 staging boxes per epilogue warpgroup
  float* s_bias = reinterpret_cast<float*>(smem_out + (kPool ? kBlkSlots * BLK_BYTES : 4 * OUT_BYTES));   // [64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 64);
  uint64_t* patch_full = bars;           // [kPatchStages <= 12]
  uint64_t* patch_empty = bars + 12;     // [kPatchStages]
  uint64_t* acc_full = bars + 24;        // [8]
  uint64_t* acc_empty = bars + 32;       // [8]
  uint64_t* w_full = bars + 40;
  uint64_t* blk_full = bars + 41;        // [kBlkSlots]  kPool: block written by both epilogue warpgroups
  uint64_t* blk_empty = bars + 44;       // [kBlkSlots]  kPool: block reduced by the pooling warps
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 48);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  if (warp == 8 && lane == 0) {
    ptx::prefetch_tmap(&p.tmS);
    ptx::prefetch_tmap(&p.tmW);
    ptx::prefetch_tmap(&p.tmD);
    for (int i = 0; i < kPatchStages; ++i) {
      ptx::mbar_init(&patch_full[i], 1);
      ptx::mbar_init(&patch_empty[i], 1);
    }
    for (int i = 0; i < kAccStages; ++i) {
      ptx::mbar_init(&acc_full[i], 1);
      ptx::mbar_init(&acc_empty[i], 4);    // one arrival per epilogue warp (32 per-thread arrivals serialise on the barrier word)
    }
    ptx::mbar_init(w_full, 1);
    for (int i = 0; i < kBlkSlots; ++i) {
      ptx::mbar_init(&blk_full[i], 8);     // one arrival per warp, after __syncwarp()
      ptx::mbar_init(&blk_empty[i], 8);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 9) ptx::tmem_alloc<kAccStages * 64>(tmem_ptr);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  const int first = blockIdx.x, step = gridDim.x, n_tiles = p.num_tiles;
  // Work units (kPool: 16 x 16 blocks = two MMA tiles; else single MMA tiles) are dealt round-robin: unit first + i * step.
  // Every role walks them with a cursor (tile column, tile row, image) advanced by the precomputed decomposition of `step`:
  // the three integer divisions per tile of the obvious formula were a third of the epilogue's instruction stream.
  const int n_units = kPool ? n_tiles / 2 : n_tiles;
  const uint32_t n_my = first < n_units ? (uint32_t)((n_units - first + step - 1) / step) * (kPool ? 2u : 1u) : 0u;
  const int per_img = p.tiles_x * p.tiles_y;
  const int step_b = step / per_img, step_y = (step - step_b * per_img) / p.tiles_x, step_x = step - step_b * per_img - step_y * p.tiles_x;
  struct Cursor {
    int tx, ty, b;
  };
  auto cursor_at = [&](int unit) {
    Cursor c;
    c.b = unit / per_img;
    const int r = unit - c.b * per_img;
    c.ty = r / p.tiles_x;
    c.tx = r - c.ty * p.tiles_x;
    return c;
  };
  auto advance = [&](Cursor& c) {
    c.tx += step_x;
    c.ty += step_y;
    c.b += step_b;
    if (c.tx >= p.tiles_x) {
      c.tx -= p.tiles_x;
      ++c.ty;
    }
    if (c.ty >= p.tiles_y) {
      c.ty -= p.tiles_y;
      ++c.b;
    }
  };
  // stem-output coordinates of the first pixel of the unit's MMA tile `half` (kPool: 0 / 1; else 0)
  auto origin = [&](const Cursor& c, int half, int& y0, int& x0) {
    if (kPool) {
      x0 = c.tx * 14 - 1 + half * TILE_W;
      y0 = c.ty * 14 - 1;
    } else {
      x0 = c.tx * TILE_W;
      y0 = c.ty * TILE_H;
    }
  };

  if (warp == 8) {
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(w_full, 16 * W_TAP_BYTES);
      for (int tap = 0; tap < 16; ++tap) ptx::tma_load_2d(&p.tmW, w_full, smem_w + tap * W_TAP_BYTES, tap * 16, 0);
      int ps = 0;
      uint32_t pphase = 0;
      Cursor cur = cursor_at(first);
      for (uint32_t n = 0; n < n_my; ++n) {
        if ((p.probe & 1) && n >= kPatchStages) break;
        int y0, x0;
        origin(cur, kPool ? (int)(n & 1) : 0, y0, x0);
        const int b = cur.b;
        if (!kPool || (n & 1)) advance(cur);
        ptx::mbar_wait(&patch_empty[ps], pphase ^ 1);
        ptx::mbar_expect_tx(&patch_full[ps], PATCH_BYTES);
        tma_load_4d(&p.tmS, &patch_full[ps], smem_patch + ps * PATCH_SLOT, 0, x0 - 2, y0 - 2, b);
        if (++ps == kPatchStages) {
          ps = 0;
          pphase ^= 1;
        }
      }
    }
  } else if (warp == 9) {
    if (ptx::elect_one()) {
      ptx::mbar_wait(w_full, 0);
      const uint64_t w0 = desc_sw32(ptx::smem_u32(smem_w), 256);
      // One tile = 16 MMAs (128 x 64 x 16, ~45 cycles each: operand-fetch bound, benchmarks/mma_probe.py).  A successful
      // mbarrier test costs ~100 cycles of latency, so the two barriers of tile n + 1 are tested BEFORE the MMAs of tile n
      // are issued and the results consumed afterwards; the blocking wait only runs when that early test failed.
      long long w_acc = 0, w_patch = 0, t_begin = probe_clock();   // probe bit 2: where the issuing thread waits
      int nr_acc = 0, nr_patch = 0;                            // early tests that found the barrier not ready
      auto acc_par = [&](uint32_t m) { return ((m / kAccStages) & 1) ^ 1; };
      auto patch_par = [&](uint32_t m) { return (m / kPatchStages) & 1; };
      const bool no_patch_wait = (p.probe & 1) != 0;
      bool acc_ok = n_my > 0 && ptx::mbar_try_wait(&acc_empty[0], acc_par(0));
      bool patch_ok = n_my > 0 && ptx::mbar_try_wait(&patch_full[0], patch_par(0));
      uint32_t n = 0;
      for (; n < n_my; ++n) {
        const uint32_t as = n % kAccStages, ps = n % kPatchStages;
        const long long c0 = probe_clock();
        if (!acc_ok) {
          ++nr_acc;
          ptx::mbar_wait(&acc_empty[as], acc_par(n));
        }
        const long long c1 = probe_clock();
        if (!patch_ok && !(no_patch_wait && n >= kPatchStages)) {
          ++nr_patch;
          ptx::mbar_wait(&patch_full[ps], patch_par(n));
        }
        w_acc += c1 - c0;
        w_patch += probe_clock() - c1;
        ptx::tc_fence_after_sync();
        if (n + 1 < n_my) {
          acc_ok = ptx::mbar_try_wait(&acc_empty[(n + 1) % kAccStages], acc_par(n + 1));
          patch_ok = ptx::mbar_try_wait(&patch_full[(n + 1) % kPatchStages], patch_par(n + 1));
        }
        const uint32_t d = tmem_base + as * 64;
        // descriptors differ only in the 16-byte-granular start address field: base descriptor + constant per tap
        const uint64_t a0 = desc_sw32(ptx::smem_u32(smem_patch + ps * PATCH_SLOT), PATCH_W * 32);
#pragma unroll
        for (int tap = 0; tap < 16; ++tap) {
          const int r = tap >> 2, sx = tap & 3;
          ptx::umma_bf16_ss(d, a0 + (uint64_t)(((r * PATCH_W + sx) * 32) >> 4), w0 + (uint64_t)((tap * W_TAP_BYTES) >> 4), kIdesc, tap != 0);
        }
        ptx::umma_commit(&patch_empty[ps]);
        ptx::umma_commit(&acc_full[as]);
      }
      if (kCounters && (p.probe & 4) && blockIdx.x == 0)
        printf("stem CTA 0 MMA thread: %u tiles, %lld cycles total, %lld waiting for accumulators (%d not ready), %lld waiting for patches (%d not ready)\n",
               n, probe_clock() - t_begin, w_acc, nr_acc, w_patch, nr_patch);
    }
  } else if (warp < 8) {
    // two epilogue warpgroups; warpgroup g handles this CTA's tiles number g, g + 2, ... (accumulator stages g, g + 2)
    const int wg = warp >> 2;
    const int et = threadIdx.x - wg * 128;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    if (threadIdx.x < 64) s_bias[threadIdx.x] = p.bias[threadIdx.x];
    ptx::named_bar_sync(3, 256);
    uint8_t* my_out = smem_out + wg * 2 * OUT_BYTES;
    uint32_t k = 0;   // tiles processed by this warpgroup; its k-th tile is the CTA's tile n = 2k + wg: stage n % kAccStages
    long long w_full = 0, t_begin = probe_clock();
    Cursor cur = cursor_at(kPool ? first : first + wg * step);   // !kPool: warpgroup g owns units g, g + 2, ...: two steps at a time
    for (uint32_t n = wg; n < n_my; n += 2) {
      int y0, x0;
      origin(cur, kPool ? wg : 0, y0, x0);
      const int b = cur.b;
      advance(cur);
      if (!kPool) advance(cur);
      const int as = (2 * k + wg) % kAccStages;
      const long long c0 = probe_clock();
      ptx::mbar_wait(&acc_full[as], ((2 * k + wg) / kAccStages) & 1);
      w_full += probe_clock() - c0;
      ptx::tc_fence_after_sync();
      uint32_t packed[32];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(tmem_base + lane_addr + as * 64 + h * 32, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          packed[h * 16 + j] = ptx::epi_bias_relu2(v[2 * j], v[2 * j + 1], *reinterpret_cast<const float2*>(s_bias + h * 32 + 2 * j));
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&acc_empty[as]);
      if constexpr (kPool) {
        // this thread's stem pixel: local (ly, lx) of the 16 x 16 block; zero outside the image
        const int ly = row >> 3, lx = wg * 8 + (row & 7);
        const bool inside = (unsigned)(y0 + ly) < (unsigned)p.H2 && (unsigned)(x0 + (row & 7)) < (unsigned)p.W2;
        const uint32_t slot = k % kBlkSlots;
        ptx::mbar_wait(&blk_empty[slot], ((k / kBlkSlots) & 1) ^ 1);     // the pooling warps are done with this slot
        uint8_t* rowp = smem_out + slot * BLK_BYTES + (ly * 16 + lx) * 128;   // [16][16] pixels x 128 B, 16-byte chunks XOR (lx & 7)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(rowp + ((j ^ (lx & 7)) << 4)) =
              inside ? make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]) : make_uint4(0, 0, 0, 0);
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&blk_full[slot]);
        ++k;
        continue;
      }
      uint8_t* buf = my_out + (k & 1) * OUT_BYTES;
      if (et == 0) ptx::tma_store_wait_read<1>();   // the store issued two tiles ago (same buffer) has read its data
      ptx::named_bar_sync(1 + wg, 128);
      uint8_t* rowp = buf + row * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(rowp + ((j ^ (row & 7)) << 4)) =
            make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
      ptx::fence_proxy_async_smem();
      ptx::named_bar_sync(1 + wg, 128);
      if (et == 0 && !(p.probe & 2)) {
        tma_store_4d(&p.tmD, buf, 0, x0, y0, b);
        ptx::tma_store_commit();
      }
      ++k;
    }
    if (et == 0) ptx::tma_store_wait_all<0>();
    if (kCounters && (p.probe & 4) && blockIdx.x == 0 && et == 0)
      printf("stem CTA 0 epilogue warpgroup %d: %u tiles, %lld cycles total, %lld waiting for accumulators\n", wg, k, probe_clock() - t_begin,
             w_full);
  }
  if (kPool && warp >= 12) {
    // ===== pooling warps: 224 of 256 threads = (pooled column j, 16-byte channel chunk, row pair rg); pooled rows 2 rg and
    // 2 rg + 1 share stem row 4 rg + 2, so five stem rows x three columns serve two outputs =====
    const int tid = threadIdx.x - 384;
    const int ch = tid & 7, jr = tid >> 3;      // jr = rg * 7 + j
    const int rg = (jr * 37) >> 8, j = jr - rg * 7;
    long long w_blk = 0, t_begin = probe_clock();
    Cursor cur = cursor_at(first);
    for (uint32_t kk = 0; 2 * kk < n_my; ++kk) {
      int y0, x0;
      origin(cur, 0, y0, x0);
      const int b = cur.b;
      advance(cur);
      const int pi0 = (y0 + 1) >> 1, pj0 = (x0 + 1) >> 1;   // first pooled row / column of the block
      const uint32_t slot = kk % kBlkSlots;
      const long long c0 = probe_clock();
      ptx::mbar_wait(&blk_full[slot], (kk / kBlkSlots) & 1);
      w_blk += probe_clock() - c0;
      const uint8_t* blk = smem_out + slot * BLK_BYTES;
      if (tid < 224 && !(p.probe & 8)) {
        __nv_bfloat162 hm[5][4];                     // horizontal 3-max of stem rows 4 rg .. 4 rg + 4
#pragma unroll
        for (int r = 0; r < 5; ++r) {
          const int sy = 4 * rg + r;
          if (sy < 15) {
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              const int cx = 2 * j + dx;
              const uint4 v = *reinterpret_cast<const uint4*>(blk + (sy * 16 + cx) * 128 + ((ch ^ (cx & 7)) << 4));
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
              for (int e = 0; e < 4; ++e) hm[r][e] = dx == 0 ? h[e] : __hmax2(hm[r][e], h[e]);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) hm[r][e] = __floats2bfloat162_rn(0.f, 0.f);
          }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&blk_empty[slot]);          // every read of the slot has returned
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const int i = 2 * rg + o;
          if (i < 7 && pi0 + i < p.P && pj0 + j < p.Q && !(p.probe & 16)) {
            __nv_bfloat162 m[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) m[e] = __hmax2(__hmax2(hm[2 * o][e], hm[2 * o + 1][e]), hm[2 * o + 2][e]);
            *reinterpret_cast<uint4*>(p.pooled + ((((long long)b * p.P + pi0 + i) * p.Q + pj0 + j) * 64 + ch * 8)) =
                *reinterpret_cast<const uint4*>(m);
          }
        }
      } else {
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&blk_empty[slot]);
      }
    }
    if (kCounters && (p.probe & 4) && blockIdx.x == 0 && tid == 0)
      printf("stem CTA 0 pooling warps: %lld cycles total, %lld waiting for blocks\n", probe_clock() - t_begin, w_blk);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 9) ptx::tmem_dealloc<kAccStages * 64>(tmem_base);
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode(CUtensorMap* tm, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
           CUtensorMapSwizzle swz) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  OPD_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  OPD_REQUIRE(fn && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled unavailable");
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(fn)(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), dims, strides,
                                                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(OPD_ERR_CUDA, "cuTensorMapEncodeTiled (stem) failed (%d)", (int)r);
  return OPD_OK;
}

}  // namespace

int stem_plan(StemPlan* plan, const __nv_bfloat16* s2d, int B, int H2, int W2, const __nv_bfloat16* w_taps, const float* bias,
              __nv_bfloat16* y, __nv_bfloat16* pooled) {
  *plan = StemPlan{};
  plan->B = B; plan->H2 = H2; plan->W2 = W2;
  plan->bias = bias;
  plan->pooled = pooled;
  plan->P = (H2 - 1) / 2 + 1;
  plan->Q = (W2 - 1) / 2 + 1;
  {
    cuuint64_t dims[4] = {16, (cuuint64_t)W2, (cuuint64_t)H2, (cuuint64_t)B};
    cuuint64_t strides[3] = {32, (cuuint64_t)W2 * 32, (cuuint64_t)H2 * W2 * 32};
    cuuint32_t box[4] = {16, PATCH_W, PATCH_H, 1};
    if (int rc = encode(&plan->tmS, s2d, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_32B)) return rc;
  }
  {
    cuuint64_t dims[2] = {256, 64};
    cuuint64_t strides[1] = {512};
    cuuint32_t box[2] = {16, 64};
    if (int rc = encode(&plan->tmW, w_taps, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_32B)) return rc;
  }
  if (pooled) {
    plan->tmD = plan->tmS;   // unused
    const int tiles = B * ((plan->P + 6) / 7) * ((plan->Q + 6) / 7);
    plan->grid = std::min(tiles, sm_count());
    return OPD_OK;
  }
  if (int rc = make_tmap_nhwc_patch(&plan->tmD, y, B, H2, W2, 64, TILE_W, TILE_H)) return rc;
  const int tiles = B * ((H2 + TILE_H - 1) / TILE_H) * ((W2 + TILE_W - 1) / TILE_W);
  plan->grid = std::min(tiles, sm_count());
  return OPD_OK;
}

int stem_launch(const StemPlan& plan, cudaStream_t stream) {
  StemParams p;
  p.tmS = plan.tmS; p.tmW = plan.tmW; p.tmD = plan.tmD;
  const bool pool = plan.pooled != nullptr;
  p.tiles_x = pool ? (plan.Q + 6) / 7 : (plan.W2 + TILE_W - 1) / TILE_W;
  p.tiles_y = pool ? (plan.P + 6) / 7 : (plan.H2 + TILE_H - 1) / TILE_H;
  p.num_tiles = plan.B * p.tiles_x * p.tiles_y * (pool ? 2 : 1);
  p.bias = plan.bias;
  p.pooled = plan.pooled;
  p.H2 = plan.H2; p.W2 = plan.W2; p.P = plan.P; p.Q = plan.Q;
  p.probe = g_option_probe.load();
  static PerDeviceOnce configured;
  if (int rc = once_per_device(configured, []() -> int {
        OPD_CUDA_OK(cudaFuncSetAttribute(stem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        OPD_CUDA_OK(cudaFuncSetAttribute(stem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPoolSmemBytes));
        return OPD_OK;
      }))
    return rc;
  if (pool)
    stem_kernel<true><<<plan.grid, kPoolThreads, kPoolSmemBytes, stream>>>(p);
  else
    stem_kernel<false><<<plan.grid, kThreads, kSmemBytes, stream>>>(p);
  count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

int launch_preprocess_s2d(const uint8_t* src, int B, int Hs, int Ws, int src_is_bgr, __nv_bfloat16* s2d, cudaStream_t s,
                          const int32_t* valid_hw) {
  const int H2 = (Hs + 1) / 2, W2 = (Ws + 1) / 2;
  s2d_preprocess_kernel<<<std::min(B * H2, sm_count() * 8), 224, 0, s>>>(src, B, Hs, Ws, src_is_bgr, s2d, H2, W2, valid_hw);
  count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

}  // namespace opd
