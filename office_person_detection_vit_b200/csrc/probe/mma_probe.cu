// MEASUREMENT PROBE (benchmarks/mma_probe.py; not on the product path): cycles per tcgen05.mma as a function of the
// instruction shape (M = 128, N in {32 .. 256}, K = 16), the operand swizzle (32-byte rows as in the stem, 128-byte rows
// as everywhere else), how many TMEM accumulators the instruction stream rotates over (1 = every MMA accumulates into the
// tile the previous one wrote), and whether consecutive MMAs read the same or different shared-memory operand tiles.
// One CTA per SM, one issuing thread, operands are whatever the shared memory holds (values do not matter for timing).
#include "opd_common.h"
#include "opd_probe.h"
#include "sm100_ptx.cuh"

namespace opd {
namespace {

__device__ __forceinline__ uint64_t probe_desc(uint32_t addr, uint32_t sbo_bytes, int swizzle32) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(swizzle32 ? 6 : 2) << 61;
  return d;
}

template <int kAcc>
__global__ void __launch_bounds__(256, 1) mma_probe_kernel(int N, int swizzle32, int iters, int walk, int a_sbo, int a_step,
                                                           int ld_iters, int commit_every, unsigned long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t dummy_bar[2];   // commit_every > 0: tcgen05.commit targets nobody waits on
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::mbar_init(&dummy_bar[0], 1);
    ptx::mbar_init(&dummy_bar[1], 1);
    ptx::fence_barrier_init();
  }
  if (threadIdx.x < 32) ptx::tmem_alloc<512>(&tmem_slot);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = ptx::umma_idesc_bf16(128, N);
    const uint32_t a_base = ptx::smem_u32(smem), b_base = ptx::smem_u32(smem + 64 * 1024);
    const uint32_t row = swizzle32 ? 32 : 128;          // bytes per operand row
    const uint32_t sbo = 8 * row;
    uint64_t da[4], db[4];
    uint32_t d[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // walk: every MMA reads a different operand tile (16 KB apart, inside 64 KB); else the K step inside a 128-byte row
      const uint32_t off = walk ? (uint32_t)j * 16384u : (swizzle32 ? 0u : (uint32_t)j * 32u);
      // a_sbo / a_step != 0: the shifted-view operand of the halo kernels (8-row groups a_sbo bytes apart, MMA j starts a_step
      // bytes after MMA j - 1), e.g. stem 352 / 32, stage-1 tail 2304 / 128
      da[j] = a_sbo ? probe_desc(a_base + (uint32_t)j * (uint32_t)a_step, (uint32_t)a_sbo, swizzle32) : probe_desc(a_base + off, sbo, swizzle32);
      db[j] = probe_desc(b_base + off, sbo, swizzle32);
      d[j] = tmem + (uint32_t)(j % kAcc) * (uint32_t)N;
    }
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) ptx::umma_bf16_ss(d[j & 3], da[j & 3], db[j & 3], idesc, 1u);
      if (commit_every > 0 && ((i + 8) % commit_every) == 0) {   // a tile boundary: two commits (operand slot + accumulator)
        ptx::umma_commit(&dummy_bar[0]);
        ptx::umma_commit(&dummy_bar[1]);
      }
    }
    ptx::umma_commit(&bar);
    ptx::mbar_wait(&bar, 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  } else if (threadIdx.x >= 128 && ld_iters > 0) {
    // warps 4-7: the epilogue's TMEM reads (tcgen05.ld 32x32b.x32 of the warp's lane quarter) running against the MMAs
    const uint32_t lane_addr = (uint32_t)(((threadIdx.x >> 5) & 3) * 32) << 16;
    uint32_t sink = 0;
    const long long t0 = clock64();
    for (int i = 0; i < ld_iters; ++i) {
      uint32_t v[32];
      ptx::tmem_ld_32x32(tmem + lane_addr + (uint32_t)(i & 7) * 32, v);
      ptx::tmem_ld_wait();
      sink ^= v[i & 31];
    }
    const long long t1 = clock64();
    if (threadIdx.x == 128) cycles[gridDim.x + blockIdx.x] = (unsigned long long)(t1 - t0) + (sink == 0x12345u ? 1 : 0);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tmem);
}

}  // namespace
}  // namespace opd

extern "C" int opd_debug_mma_probe(int32_t N, int32_t swizzle32, int32_t n_acc, int32_t iters, int32_t walk, int32_t grid,
                                   int32_t a_sbo, int32_t a_step, int32_t ld_iters, int32_t commit_every, uint64_t* cycles_dev, void* stream) {
  OPD_REQUIRE(N >= 8 && N <= 256 && N % 8 == 0 && n_acc >= 1 && n_acc * N <= 512 && iters > 0 && grid > 0 && cycles_dev,
              "opd_debug_mma_probe: bad argument");
  OPD_REQUIRE(n_acc == 1 || n_acc == 2 || n_acc == 4, "opd_debug_mma_probe: n_acc must be 1, 2 or 4");
  static bool configured = false;
  if (!configured) {
    OPD_CUDA_OK(cudaFuncSetAttribute(opd::mma_probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    OPD_CUDA_OK(cudaFuncSetAttribute(opd::mma_probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    OPD_CUDA_OK(cudaFuncSetAttribute(opd::mma_probe_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured = true;
  }
  auto* out = reinterpret_cast<unsigned long long*>(cycles_dev);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n_acc == 1) opd::mma_probe_kernel<1><<<grid, 256, 160 * 1024, s>>>(N, swizzle32, iters, walk, a_sbo, a_step, ld_iters, commit_every, out);
  else if (n_acc == 2) opd::mma_probe_kernel<2><<<grid, 256, 160 * 1024, s>>>(N, swizzle32, iters, walk, a_sbo, a_step, ld_iters, commit_every, out);
  else opd::mma_probe_kernel<4><<<grid, 256, 160 * 1024, s>>>(N, swizzle32, iters, walk, a_sbo, a_step, ld_iters, commit_every, out);
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}
