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
  int reverse = 0;   // walk the m-blocks from the last one down (set by the engine for layers that follow an ascending writer)
  CUtensorMap tmA, tmB, tmD, tmD2, tmR;   // tmR: residual, read by TMA in 64-column chunks
  CUtensorMap tmA2;    // second A source (k-blocks >= k_split): NHWC tensor through a 1x1 / stride-s im2col map
  int k_split;         // k-blocks served by tmA; == K / 64 without a second source
  ConvGeom g2;         // geometry of the second source (1x1, stride s, pad 0)
  int M, N, K;
  int block_n;
  int m_tiles;         // 2: the CTA works on pairs of m-blocks that share every weight tile (long-K layers)
  int out_bufs;        // staging boxes per epilogue warpgroup (2 for short-K layers)
  int cluster;         // 1: 2-CTA cluster variant (weight tiles multicast to both CTAs)
  int b_resident;      // 1: weight-stationary kernel variant (one n-block per CTA, its weights resident in shared memory)
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
  int pos_row0;        // pos row of this GEMM's row 0 (row-range GEMMs); 0 by default
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
// D = epi( A[M,K1] * W[:, :K1]^T + im2col_1x1(x2; stride)[M,K2] * W[:, K1:]^T + bias ): a bottleneck's 1x1 expansion and its
// strided projection shortcut accumulated into the same TMEM tile (no shortcut tensor).  W [N, K1 + K2]; g2: 1x1 geometry.
int gemm_plan_linear_plus_shortcut(GemmPlan* plan, const __nv_bfloat16* A, int K1, const __nv_bfloat16* x2, const ConvGeom& g2,
                                   const __nv_bfloat16* W, __nv_bfloat16* D, int N, int epi, const float* bias);
int gemm_launch(const GemmPlan& plan, cudaStream_t stream);

int sm_count();

// Tensor maps (128-byte swizzle): 2D bf16 matrix [rows, cols] with row pitch ld elements, box = [box_rows, 64 columns];
// NHWC activation in im2col mode, box = 128 output pixels x 64 channels of one filter tap.
int make_tmap_2d(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
int make_tmap_im2col(CUtensorMap* tm, const void* ptr, const ConvGeom& g);

// Fused bottleneck tail (tc_bottleneck.cu): y = relu(conv1x1(relu(conv3x3(x, stride) + b2)) + b3 + residual).
// x [B,H,W,MID] NHWC, w2 [MID,3,3,MID], w3 [WIDTH,MID], residual / y [B,P,Q,WIDTH]; MID in {64, 128}, WIDTH % 128 == 0.
// 4D tiled map over an NHWC bf16 tensor: box = [64 channels, box_w, box_h, 1 image], 128-byte swizzle, zero fill
int make_tmap_nhwc_patch(CUtensorMap* tm, const void* ptr, int B, int H, int W, int C, int box_w, int box_h);

// Fused attention on tcgen05 (attention_tc.cu): o = softmax(q k^T / sqrt(32)) v, head h = columns [32h, 32h + 32).
struct AttnPlan {
  CUtensorMap tmQ, tmK, tmV;
  __nv_bfloat16* o;
  int64_t ldo;
  int B, heads, Lq, Lk;
  int kv_tile;   // keys per tile: 64 (4 CTAs per SM) or 128 (2 CTAs per SM)
  // key-padding mask of padded batches (modeling_detr.py:386-411 attention_mask): bit k % 32 of word [b, k / 32] set <=> key k of
  // frame b takes part; nullptr: every key does.  key_mask_stride = words per frame.
  const uint32_t* key_mask = nullptr;
  int key_mask_stride = 0;
};
int attn_plan(AttnPlan* plan, const __nv_bfloat16* q, int64_t ldq, const __nv_bfloat16* k, int64_t ldk, const __nv_bfloat16* v,
              int64_t ldv, __nv_bfloat16* o, int64_t ldo, int B, int heads, int Lq, int Lk);
int attn_launch(const AttnPlan& plan, cudaStream_t stream);

// Fused feed-forward block (tc_mlp.cu): y = LayerNorm(relu(x W1^T + b1) W2^T + b2 + x) * gamma + beta, y2 = bf16(y + pos[row % pos_rows]);
// x, y, y2 [M, 256] bf16 (contiguous rows), W1 [2048, 256], W2 [256, 2048].  Bit-identical to the two GEMM launches it replaces.
struct MlpPlan {
  CUtensorMap tmX, tmW1, tmW2, tmD, tmD2;
  int M;
  const float* b1;
  const float* b2;
  const float* gamma;
  const float* beta;
  const float* pos;
  int pos_rows;
  int pos_row0 = 0;
  int has_d2;
  int pair;            // 1: cta_group::2 variant (two row tiles = one 256-row MMA, half of every weight tile per SM)
  int grid;
};
int mlp_plan(MlpPlan* plan, const __nv_bfloat16* x, const __nv_bfloat16* w1, const float* b1, const __nv_bfloat16* w2, const float* b2,
             const float* gamma, const float* beta, __nv_bfloat16* d, __nv_bfloat16* d2, const float* pos, int pos_rows, int M);
int mlp_launch(const MlpPlan& plan, cudaStream_t stream);

// Stem (stem_conv.cu): 4x4-tap convolution over the space-to-depth tensor S[B,H2,W2,16] -> y[B,H2,W2,64], bias + ReLU.
struct StemPlan {
  CUtensorMap tmS, tmW, tmD;
  int B, H2, W2, P, Q;
  const float* bias;
  __nv_bfloat16* pooled;   // != nullptr: the 3x3 / stride 2 max pooling is fused; y is not written
  int grid;
};
// pooled != nullptr: y[B,P,Q,64] = maxpool3x3s2(relu(stem)) in one kernel (y_stem unused)
int stem_plan(StemPlan* plan, const __nv_bfloat16* s2d, int B, int H2, int W2, const __nv_bfloat16* w_taps, const float* bias,
              __nv_bfloat16* y, __nv_bfloat16* pooled = nullptr);
int stem_launch(const StemPlan& plan, cudaStream_t stream);

struct BneckPlan {
  CUtensorMap tmA, tmB1, tmB2, tmR, tmD;
  int fused_shortcut;  // halo variant only: the projection shortcut is a second k-block of the 1x1 expansion
  int pair;            // 1: cta_group::2 variant of tc_bneck_kernel<128> (two CTAs = one 256-pixel MMA, half a weight tile per SM)
  int halo;            // 1: stride-1, MID = 64 variant (tc_bottleneck_halo.cu): 16 x 8 pixel tiles, A = one halo patch per tile
  int M, mid, width;
  ConvGeom g;
  const float* bias2;
  const float* bias3;
  int grid;
};
int bneck_plan(BneckPlan* plan, const __nv_bfloat16* x, const ConvGeom& g, const __nv_bfloat16* w2, const float* bias2,
               const __nv_bfloat16* w3, const float* bias3, int width, const __nv_bfloat16* residual, __nv_bfloat16* y);
int bneck_launch(const BneckPlan& plan, cudaStream_t stream);
void g_bneck_trace_set(unsigned long long* p);   // debug timeline of tc_bneck_kernel (see opd_debug_set_bneck_trace)
int bneck_halo_plan(BneckPlan* plan, const __nv_bfloat16* x, const ConvGeom& g, const __nv_bfloat16* w2, const float* bias2,
                    const __nv_bfloat16* w3, const float* bias3, int width, const __nv_bfloat16* residual, __nv_bfloat16* y,
                    const __nv_bfloat16* shortcut_in = nullptr);
int bneck_halo_launch(const BneckPlan& plan, cudaStream_t stream);

}  // namespace opd
