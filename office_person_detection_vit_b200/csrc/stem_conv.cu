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
constexpr int kPatchStages = 8;
constexpr int W_TAP_BYTES = 64 * 32;                   // one tap: 64 output channels x 16 input lanes
constexpr int OUT_BYTES = 128 * 128;                   // staging box: 128 pixels x 64 channels bf16
constexpr int kAccStages = 8;
constexpr int kMmaGroup = 4;   // tiles whose MMAs are interleaved: consecutive MMAs never accumulate into the same TMEM tile
constexpr int kThreads = 384;
constexpr int kSmemBytes = 16 * W_TAP_BYTES + kPatchStages * PATCH_SLOT + 4 * OUT_BYTES + 1024;

struct StemParams {
  CUtensorMap tmS, tmW, tmD;
  int tiles_x, tiles_y, num_tiles;
  const float* bias;
  __nv_bfloat16* pooled;   // kPool: max-pool output [B, P, Q, 64]
  int H2, W2, P, Q;        // stem output size, pooled size
  int probe;   // measurement probes (opd_set_option("probe")): bit 0 no patch reloads, bit 1 no output stores
};

__global__ void s2d_preprocess_kernel(const uint8_t* __restrict__ src, int B, int Hs, int Ws, int bgr,
                                      __nv_bfloat16* __restrict__ S, int H2, int W2) {
  const float mean[3] = {0.485f * 255.0f, 0.456f * 255.0f, 0.406f * 255.0f};
  const float stdv[3] = {0.229f * 255.0f, 0.224f * 255.0f, 0.225f * 255.0f};
  const long long total = (long long)B * H2 * W2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W2);
    long long r = i / W2;
    const int y = (int)(r % H2);
    const int b = (int)(r / H2);
    float v[12];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int iy = 2 * y + dy, ix = 2 * x + dx;
        const bool ok = iy < Hs && ix < Ws;
        const uint8_t* px = src + (((long long)b * Hs + (ok ? iy : 0)) * Ws + (ok ? ix : 0)) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float u = (float)px[bgr ? 2 - c : c];
          v[(dy * 2 + dx) * 3 + c] = ok ? (u - mean[c]) / stdv[c] : 0.f;
        }
      }
    uint4 o0, o1;
    o0.x = ptx::pack_bf16(v[0], v[1]); o0.y = ptx::pack_bf16(v[2], v[3]); o0.z = ptx::pack_bf16(v[4], v[5]); o0.w = ptx::pack_bf16(v[6], v[7]);
    o1.x = ptx::pack_bf16(v[8], v[9]); o1.y = ptx::pack_bf16(v[10], v[11]); o1.z = 0u; o1.w = 0u;
    uint4* dst = reinterpret_cast<uint4*>(S + i * 16);
    dst[0] = o0;
    dst[1] = o1;
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
__global__ void __launch_bounds__(kThreads, 1) stem_kernel(const __grid_constant__ StemParams p) {
  constexpr uint32_t kIdesc = ptx::umma_idesc_bf16(128, 64);
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_w = smem;                                   // 16 taps x [64 x 32 B]
  uint8_t* smem_patch = smem_w + 16 * W_TAP_BYTES;          // [kPatchStages]
  uint8_t* smem_out = smem_patch + kPatchStages * PATCH_SLOT;   // 2 staging boxes per epilogue warpgroup
  float* s_bias = reinterpret_cast<float*>(smem_out + 4 * OUT_BYTES);   // [64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 64);
  uint64_t* patch_full = bars;           // [8]
  uint64_t* patch_empty = bars + 8;      // [8]
  uint64_t* acc_full = bars + 16;        // [8]
  uint64_t* acc_empty = bars + 24;       // [8]
  uint64_t* w_full = bars + 32;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 33);

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
      ptx::mbar_init(&acc_empty[i], 128);
    }
    ptx::mbar_init(w_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 9) ptx::tmem_alloc<kAccStages * 64>(tmem_ptr);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  const int first = blockIdx.x, step = gridDim.x, n_tiles = p.num_tiles;
  // MMA tile t -> image b and the stem-output coordinates of its first pixel.  kPool: t = 2 * work tile + half.
  auto tile_origin = [&](int t, int& b, int& y0, int& x0) {
    const int w = kPool ? t >> 1 : t;
    const int tx = w % p.tiles_x;
    const int r = w / p.tiles_x;
    const int ty = r % p.tiles_y;
    b = r / p.tiles_y;
    if (kPool) {
      x0 = tx * 14 - 1 + (t & 1) * TILE_W;
      y0 = ty * 14 - 1;
    } else {
      x0 = tx * TILE_W;
      y0 = ty * TILE_H;
    }
  };

  // this CTA's n-th MMA tile.  kPool: both halves of a work tile belong to the same CTA (n = 2 * local work tile + half)
  const int n_units = kPool ? n_tiles / 2 : n_tiles;
  const uint32_t n_my = first < n_units ? (uint32_t)((n_units - first + step - 1) / step) * (kPool ? 2u : 1u) : 0u;
  auto tile_of = [&](uint32_t n) -> int { return kPool ? 2 * (first + (int)(n >> 1) * step) + (int)(n & 1) : first + (int)n * step; };

  if (warp == 8) {
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(w_full, 16 * W_TAP_BYTES);
      for (int tap = 0; tap < 16; ++tap) ptx::tma_load_2d(&p.tmW, w_full, smem_w + tap * W_TAP_BYTES, tap * 16, 0);
      int ps = 0;
      uint32_t pphase = 0;
      for (uint32_t n = 0; n < n_my; ++n) {
        if ((p.probe & 1) && n >= kPatchStages) break;
        int b, y0, x0;
        tile_origin(tile_of(n), b, y0, x0);
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
      // Tiles are issued in groups of kMmaGroup with their MMAs interleaved tap by tap: a 128x64x16 MMA occupies the
      // tensor pipe for 32 cycles but a dependent accumulate into the SAME TMEM tile waits ~100 cycles for the previous
      // one, so back-to-back taps of one tile ran at a third of the pipe rate (measured: 16 MMAs = 1700 cycles).
      uint32_t n = 0;   // tiles issued so far by this CTA: patch slot n % kPatchStages, accumulator stage n % kAccStages
      long long w_acc = 0, w_patch = 0, t_begin = clock64();   // probe bit 2: where the issuing thread waits
      while (n < n_my) {
        uint32_t d[kMmaGroup];
        uint64_t a[kMmaGroup];
        int g = 0;
#pragma unroll
        for (int j = 0; j < kMmaGroup; ++j) {
          if (n + j < n_my) {
            const uint32_t m = n + j, as = m % kAccStages, ps = m % kPatchStages;
            const long long c0 = clock64();
            ptx::mbar_wait(&acc_empty[as], ((m / kAccStages) & 1) ^ 1);
            const long long c1 = clock64();
            if (!(p.probe & 1) || m < kPatchStages) ptx::mbar_wait(&patch_full[ps], (m / kPatchStages) & 1);
            w_acc += c1 - c0;
            w_patch += clock64() - c1;
            d[j] = tmem_base + as * 64;
            // descriptors differ only in the 16-byte-granular start address field: base descriptor + constant per tap
            a[j] = desc_sw32(ptx::smem_u32(smem_patch + ps * PATCH_SLOT), PATCH_W * 32);
            g = j + 1;
          }
        }
        ptx::tc_fence_after_sync();
#pragma unroll
        for (int tap = 0; tap < 16; ++tap) {
          const int r = tap >> 2, sx = tap & 3;
#pragma unroll
          for (int j = 0; j < kMmaGroup; ++j)
            if (j < g)
              ptx::umma_bf16_ss(d[j], a[j] + (uint64_t)(((r * PATCH_W + sx) * 32) >> 4), w0 + (uint64_t)((tap * W_TAP_BYTES) >> 4), kIdesc,
                                tap != 0);
        }
#pragma unroll
        for (int j = 0; j < kMmaGroup; ++j)
          if (j < g) {
            ptx::umma_commit(&patch_empty[(n + j) % kPatchStages]);
            ptx::umma_commit(&acc_full[(n + j) % kAccStages]);
          }
        n += g;
      }
      if ((p.probe & 4) && blockIdx.x == 0)
        printf("stem CTA 0 MMA thread: %u tiles, %lld cycles total, %lld waiting for accumulators, %lld waiting for patches\n", n,
               clock64() - t_begin, w_acc, w_patch);
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
    long long w_full = 0, t_begin = clock64();
    for (uint32_t n = 0; n < n_my; ++n) {
      if ((int)(n & 1) != wg) continue;
      const int t = tile_of(n);
      int b, y0, x0;
      tile_origin(t, b, y0, x0);
      const int as = (2 * k + wg) % kAccStages;
      const long long c0 = clock64();
      ptx::mbar_wait(&acc_full[as], ((2 * k + wg) / kAccStages) & 1);
      w_full += clock64() - c0;
      ptx::tc_fence_after_sync();
      uint32_t packed[32];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(tmem_base + lane_addr + as * 64 + h * 32, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          packed[h * 16 + j] = ptx::pack_bf16(fmaxf(__uint_as_float(v[2 * j]) + s_bias[h * 32 + 2 * j], 0.f),
                                              fmaxf(__uint_as_float(v[2 * j + 1]) + s_bias[h * 32 + 2 * j + 1], 0.f));
      }
      ptx::tc_fence_before_sync();
      ptx::mbar_arrive(&acc_empty[as]);
      if constexpr (kPool) {
        // this thread's stem pixel: local (ly, lx) of the 16 x 16 block; zero outside the image
        const int ly = row >> 3, lx = wg * 8 + (row & 7);
        const bool inside = (unsigned)(y0 + ly) < (unsigned)p.H2 && (unsigned)(x0 + (row & 7)) < (unsigned)p.W2;
        uint8_t* blk = smem_out + (k & 1) * (2 * OUT_BYTES);          // [16][16] pixels x 128 B, 16-byte chunks XOR (lx & 7)
        uint8_t* rowp = blk + (ly * 16 + lx) * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(rowp + ((j ^ (lx & 7)) << 4)) =
              inside ? make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]) : make_uint4(0, 0, 0, 0);
        ptx::named_bar_sync(4, 256);   // both halves of the block are in shared memory (the other warpgroup's tile n ^ 1)
        const int pi0 = ((t >> 1) / p.tiles_x % p.tiles_y) * 7, pj0 = ((t >> 1) % p.tiles_x) * 7;
        for (int item = threadIdx.x; item < 49 * 8; item += 256) {
          const int px = item >> 3, ch = item & 7;
          const int i = px / 7, j = px - i * 7;
          if (pi0 + i < p.P && pj0 + j < p.Q) {
            __nv_bfloat162 m[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) m[e] = __floats2bfloat162_rn(0.f, 0.f);   // inputs are >= 0
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                const int cx = 2 * j + dx;
                const uint4 v = *reinterpret_cast<const uint4*>(blk + ((2 * i + dy) * 16 + cx) * 128 + ((ch ^ (cx & 7)) << 4));
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
                for (int e = 0; e < 4; ++e) m[e] = __hmax2(m[e], h[e]);
              }
            *reinterpret_cast<uint4*>(p.pooled + ((((long long)b * p.P + pi0 + i) * p.Q + pj0 + j) * 64 + ch * 8)) =
                *reinterpret_cast<const uint4*>(m);
          }
        }
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
    if ((p.probe & 4) && blockIdx.x == 0 && et == 0)
      printf("stem CTA 0 epilogue warpgroup %d: %u tiles, %lld cycles total, %lld waiting for accumulators\n", wg, k, clock64() - t_begin,
             w_full);
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
  static bool configured = false;
  if (!configured) {
    OPD_CUDA_OK(cudaFuncSetAttribute(stem_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    OPD_CUDA_OK(cudaFuncSetAttribute(stem_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  if (pool)
    stem_kernel<true><<<plan.grid, kThreads, kSmemBytes, stream>>>(p);
  else
    stem_kernel<false><<<plan.grid, kThreads, kSmemBytes, stream>>>(p);
  count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

int launch_preprocess_s2d(const uint8_t* src, int B, int Hs, int Ws, int src_is_bgr, __nv_bfloat16* s2d, cudaStream_t s) {
  const int H2 = (Hs + 1) / 2, W2 = (Ws + 1) / 2;
  const long long total = (long long)B * H2 * W2;
  const long long blocks = (total + 255) / 256;
  s2d_preprocess_kernel<<<(int)std::min<long long>(blocks, 148LL * 16), 256, 0, s>>>(src, B, Hs, Ws, src_is_bgr, s2d, H2, W2);
  count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

}  // namespace opd
