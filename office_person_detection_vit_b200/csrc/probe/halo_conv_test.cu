// EXPERIMENT (unit-test surface only): 3x3 / stride 1 / pad 1 convolution, C = N = 64, where the A operand of all nine
// taps is ONE halo patch in shared memory (18 x 10 pixels x 128 B, loaded once by a tiled TMA with zero fill) and
// each tap's tcgen05.mma reads it through a shifted shared-memory descriptor (start + (r * 10 + s) * 128 B, stride
// between 8-pixel groups = one patch row = 1280 B) instead of nine im2col copies.  Output tile = 16 rows x 8 columns.
#include "opd_common.h"
#include "opd_probe.h"
#include "sm100_ptx.cuh"
#include "tc_gemm.h"

namespace opd {
namespace {

constexpr int PATCH_W = 10, PATCH_H = 18, TILE_W = 8, TILE_H = 16;
constexpr int PATCH_BYTES = PATCH_W * PATCH_H * 128;   // 23040
constexpr int PATCH_SLOT = 24576;

struct HaloParams {
  CUtensorMap tmX, tmW;
  int B, H, W;
  int tiles_x, tiles_y;
  const float* bias;
  __nv_bfloat16* y;
  int mode;
};

__device__ __forceinline__ void tma_load_4d_tiled(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(ptx::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t sbo_bytes, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) halo_conv_kernel(const __grid_constant__ HaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* patch = smem;
  uint8_t* wts = smem + PATCH_SLOT;                 // 9 x [64 x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(wts + 9 * 8192);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bars[0], 1);
    ptx::mbar_init(&bars[1], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) ptx::tmem_alloc<64>(tmem_ptr);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = *tmem_ptr;

  const int tile = blockIdx.x;
  const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, b = tile / (p.tiles_x * p.tiles_y);
  const int x0 = tx * TILE_W, y0 = ty * TILE_H;

  if (threadIdx.x == 0) {
    ptx::mbar_expect_tx(&bars[0], PATCH_BYTES + 9 * 8192);
    tma_load_4d_tiled(&p.tmX, &bars[0], patch, 0, x0 - 1, y0 - 1, b);
    for (int tap = 0; tap < 9; ++tap) ptx::tma_load_2d(&p.tmW, &bars[0], wts + tap * 8192, tap * 64, 0);
    ptx::mbar_wait(&bars[0], 0);
    ptx::tc_fence_after_sync();
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, 64);
    for (int tap = 0; tap < 9; ++tap) {
      const int r = tap / 3, s = tap % 3;
      const uint32_t a_addr = ptx::smem_u32(patch) + (r * PATCH_W + s) * 128;
      const uint32_t b_addr = ptx::smem_u32(wts + tap * 8192);
      const uint32_t bo = p.mode == 1 ? ((a_addr >> 7) & 7) : 0;
      for (int k = 0; k < 4; ++k)
        ptx::umma_bf16_ss(tmem, desc_sw128(a_addr + k * 32, PATCH_W * 128, bo), desc_sw128(b_addr + k * 32, 1024, 0), idesc,
                          (tap | k) != 0);
    }
    ptx::umma_commit(&bars[1]);
  }
  ptx::mbar_wait(&bars[1], 0);
  ptx::tc_fence_after_sync();
  const int row = threadIdx.x;                     // = TMEM lane = (y, x) = (row / 8, row % 8)
  const int oy = y0 + row / TILE_W, ox = x0 + row % TILE_W;
  const uint32_t t_acc = tmem + ((uint32_t)(warp * 32) << 16);
  for (int h = 0; h < 2; ++h) {
    uint32_t v[32];
    ptx::tmem_ld_32x32(t_acc + h * 32, v);
    ptx::tmem_ld_wait();
    if (oy < p.H && ox < p.W) {
      __nv_bfloat16* dst = p.y + (((long long)b * p.H + oy) * p.W + ox) * 64 + h * 32;
      for (int j = 0; j < 32; ++j) dst[j] = __float2bfloat16(fmaxf(__uint_as_float(v[j]) + p.bias[h * 32 + j], 0.f));
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<64>(tmem);
}

}  // namespace

}  // namespace opd

extern "C" int opd_halo_conv3x3_test(const void* x_dev, int32_t B, int32_t H, int32_t W, const void* w_dev, const float* bias_dev,
                                     void* y_dev, int32_t mode, void* stream) {
  using namespace opd;
  HaloParams p;
  if (int rc = make_tmap_nhwc_patch(&p.tmX, x_dev, B, H, W, 64, PATCH_W, PATCH_H)) return rc;
  if (int rc = make_tmap_2d(&p.tmW, w_dev, 64, 576, 576, 64)) return rc;
  p.B = B; p.H = H; p.W = W;
  p.tiles_x = (W + TILE_W - 1) / TILE_W;
  p.tiles_y = (H + TILE_H - 1) / TILE_H;
  p.bias = bias_dev;
  p.y = static_cast<__nv_bfloat16*>(y_dev);
  p.mode = mode;
  const int smem = PATCH_SLOT + 9 * 8192 + 64;
  static bool configured = false;
  if (!configured) {
    OPD_CUDA_OK(cudaFuncSetAttribute(halo_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  halo_conv_kernel<<<B * p.tiles_x * p.tiles_y, 128, smem, static_cast<cudaStream_t>(stream)>>>(p);
  count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}
