// Internal C++ interface of the tcgen05 GEMM / implicit-GEMM convolution kernel (tc_gemm.cu).
//   D[M,N] = epilogue( A[M,K] * W[N,K]^T + bias[N] )        bf16 operands, fp32 accumulation in TMEM, bf16 out
// A is either a row-major matrix (linear layers, 1x1 stride-1 convolutions on NHWC activations) or an NHWC
// activation tensor read through TMA im2col mode (3x3 convolutions, any stride; strided 1x1 shortcuts).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace opd {

enum EpiMode : int {
  EPI_BIAS = 0,           // D = acc + bias
  EPI_BIAS_RELU = 1,      // D = relu(acc + bias)
  EPI_BIAS_RES_RELU = 2,  // D = relu(acc + bias + R)                      (bottleneck output)
  EPI_BIAS_RES_LN = 3,    // D = LayerNorm(acc + bias + R) * gamma + beta   (N == 256 == one tile row per thread)
};
// Any epilogue may also write D2 = bf16(D + pos[row % pos_rows]) (the query/key input of the next attention).

struct ConvGeom {
  int B, H, W, C;      // NHWC input
  int KH, KW, stride;
  int pad_h, pad_w;    // zero padding before the first row / column
  int P, Q;            // output height / width (fixes the padding after the last row / column)
};

struct GemmPlan {
  CUtensorMap tmA, tmB, tmD, tmD2;
  int M, N, K;
  int block_n;
  int im2col;          // 0: A is [M,K] rows; 1: A is NHWC through im2col TMA
  ConvGeom g;
  int epi;
  const float* bias;
  const __nv_bfloat16* residual;
  int ldr;
  const float* gamma;
  const float* beta;
  const float* pos;
  int pos_rows;
  int has_d2;
  int grid;
};

// Row-major A [M,K] (row stride lda elements), W [N,K] (row stride K), D [M,N] (row stride ldd).
int gemm_plan_linear(GemmPlan* plan, const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* W, __nv_bfloat16* D,
                     int64_t ldd, int M, int N, int K, int epi, const float* bias, const __nv_bfloat16* residual,
                     int64_t ldr, const float* gamma, const float* beta, __nv_bfloat16* D2, const float* pos,
                     int pos_rows);
// NHWC x [B,H,W,C]; W [N, KH*KW*C] with K ordered (kh, kw, c); D [B*P*Q, N].
int gemm_plan_conv(GemmPlan* plan, const __nv_bfloat16* x, const ConvGeom& g, const __nv_bfloat16* W, __nv_bfloat16* D,
                   int N, int epi, const float* bias, const __nv_bfloat16* residual);
int gemm_launch(const GemmPlan& plan, cudaStream_t stream);

int sm_count();

}  // namespace opd
