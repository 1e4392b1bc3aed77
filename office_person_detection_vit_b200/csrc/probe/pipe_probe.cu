// Instruction-pipe throughput probes (benchmarks/pipe_probe.py): how many MUFU.EX2 / F2FP / FFMA2 per clock an SM sustains, alone
// and mixed in the proportions of the attention kernel's softmax loop.  Measurement only (libopd_probe.so).
#include <cuda_runtime.h>

#include <cstdint>

#include "opd_common.h"
#include "opd_probe.h"

namespace {

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// mode 0: 8 independent ex2 chains; 1: 8 cvt.bf16x2 chains; 2: 8 fma.f32x2 chains; 3: softmax mix per element pair
// (fma2, 2 x ex2, add2 via fma2, cvt pack)
__global__ void __launch_bounds__(256) pipe_probe_kernel(int mode, int iters, float seed, float* out, unsigned long long* cycles) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + 0.001f * (threadIdx.x + i);
  uint64_t p[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = ((uint64_t)__float_as_uint(a[2 * i]) << 32) | __float_as_uint(a[2 * i + 1]);
  uint32_t acc = 0;
  const unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = ex2(a[i]);
    } else if (mode == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = __uint_as_float(pack_bf16(a[i], a[(i + 1) & 7]));
    } else if (mode == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = fma2(p[i], p[(i + 1) & 3], p[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint64_t x = fma2(p[i], p[(i + 1) & 3], p[(i + 2) & 3]);
        const float e0 = ex2(__uint_as_float((uint32_t)x)), e1 = ex2(__uint_as_float((uint32_t)(x >> 32)));
        const uint64_t e = ((uint64_t)__float_as_uint(e1) << 32) | __float_as_uint(e0);
        p[i] = fma2(e, p[i], p[i]);
        acc ^= pack_bf16(e0, e1);
      }
    }
  }
  const unsigned long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) s += __uint_as_float((uint32_t)p[i]) + __uint_as_float((uint32_t)(p[i] >> 32));
  if (s == 12345.678f || acc == 0x12345u) out[0] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

}  // namespace

extern "C" int opd_debug_pipe_probe(int32_t mode, int32_t iters, int32_t grid, int32_t threads, float* out_dev, uint64_t* cycles_dev,
                                    void* stream) {
  pipe_probe_kernel<<<grid, threads, 0, static_cast<cudaStream_t>(stream)>>>(mode, iters, 0.5f, out_dev,
                                                                              reinterpret_cast<unsigned long long*>(cycles_dev));
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}
