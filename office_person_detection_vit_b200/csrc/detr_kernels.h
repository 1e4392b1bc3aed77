// Internal C++ interface of the non-GEMM DETR kernels (detr_kernels.cu).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace opd {

// K1: uint8 frames -> normalised bf16 stem input in the "space-to-depth + column-unrolled" layout
//   X2[b, y, x, kw*16 + (dy*2+dx)*3 + c] = norm(src[b, 2y+dy, 2(x+kw-2)+dx, c])   (0 outside the image, 0 for the 4 pad lanes)
// so that the 7x7/s2 stem convolution is a 4x1 convolution over 64 channels (K = 256) on the tensor cores.
int launch_preprocess(const uint8_t* src, int B, int Hs, int Ws, int src_is_bgr, __nv_bfloat16* x2, cudaStream_t s);
// K1 (stem_conv.cu): uint8 frames -> normalised bf16 space-to-depth tensor S[B, (Hs+1)/2, (Ws+1)/2, 16]
// valid_hw (optional, device [B, 2]): per-frame picture size inside the Hs x Ws canvas; the padding becomes 0 after normalisation
int launch_preprocess_s2d(const uint8_t* src, int B, int Hs, int Ws, int src_is_bgr, __nv_bfloat16* s2d, cudaStream_t s,
                          const int32_t* valid_hw = nullptr);
// uint8 bilinear resize with antialias (ATen's separable uint8 kernel: horizontal pass, then vertical pass,
// fixed-point weights); src [B,H0,W0,3] (BGR or RGB) -> dst [B,H1,W1,3] RGB
// dst_frame_stride / dst_row_pitch (bytes, 0 = contiguous [B,H1,W1,3]): write into the top-left corner of a larger canvas
int launch_resize_u8(const uint8_t* src, int B, int H0, int W0, int src_is_bgr, uint8_t* tmp, uint8_t* dst, int H1,
                     int W1, const int16_t* wx, const int32_t* x0, int kx, int px, const int16_t* wy, const int32_t* y0,
                     int ky, int py, cudaStream_t s, long long dst_frame_stride = 0, long long dst_row_pitch = 0);
// frames that need no resize, copied (BGR -> RGB) into the top-left corner of their canvas
int launch_copy_into_canvas(const uint8_t* src, int B, int H, int W, int src_is_bgr, uint8_t* dst, long long dst_frame_stride,
                            long long dst_row_pitch, cudaStream_t s);
// synthetic uint8 frames [B,H,W,3] generated on the device (bench.py config 4), frame b = global frame frame0 + b
int launch_synthetic_frames(unsigned long long seed_base, long long frame0, int B, int H, int W, uint8_t* out, cudaStream_t s);
// K3: 3x3 / stride 2 / pad 1 max pooling, NHWC bf16
int launch_maxpool(const __nv_bfloat16* x, int B, int H, int W, int C, __nv_bfloat16* y, int P, int Q, cudaStream_t s);
// K7: sine positional embedding table [h*w, 256] fp32 (all-ones mask)
int launch_pos_embed(float* pos, int h, int w, cudaStream_t s);
// K7 for padded batches: per-frame table [B, h*w, 256] from the per-frame valid feature rectangle fvalid [B, 2] = (fh, fw)
int launch_pos_embed_masked(float* pos, int B, int h, int w, const int32_t* fvalid, cudaStream_t s);
// K6: multi-head attention, head_dim 32.  q/k/v/o are [B, L, *] with row strides ld* (elements); head h uses
// columns [32h, 32h+32).  o = softmax(q k^T / sqrt(32)) v
int launch_attention(const __nv_bfloat16* q, int64_t ldq, const __nv_bfloat16* k, int64_t ldk, const __nv_bfloat16* v,
                     int64_t ldv, __nv_bfloat16* o, int64_t ldo, int B, int heads, int Lq, int Lk, cudaStream_t s);
// y = 0, yp = bf16(qpos) broadcast over the batch
int launch_decoder_init(__nv_bfloat16* y, __nv_bfloat16* yp, const float* qpos, int B, int Q, cudaStream_t s);
// final decoder LayerNorm (rows of 256)
int launch_layernorm(const __nv_bfloat16* x, const float* gamma, const float* beta, __nv_bfloat16* y, int rows,
                     cudaStream_t s);

struct HeadWeights {
  const float* wc_t;   // [256, 92]   class_labels_classifier.weight^T
  const float* bc;     // [92]
  const float* w0_t;   // [256, 256]  bbox_predictor.layers.0.weight^T
  const float* b0;
  const float* w1_t;   // [256, 256]
  const float* b1;
  const float* w2;     // [4, 256]
  const float* b2;
};
// K8a: class logits [rows, 92] and sigmoid boxes [rows, 4] from the decoder output, fp32
int launch_heads(const __nv_bfloat16* y, const HeadWeights& w, float* logits, float* boxes, int rows, cudaStream_t s);
// ROI mean-pool + L2 norm of the encoder map over each compacted detection box (feature_extractor.py:39-88)
int launch_roi_features(const __nv_bfloat16* feat, int B, int fh, int fw, int D, const double* xywh, const int32_t* n_keep,
                        int Q, int img_h, int img_w, float* out, cudaStream_t s);

// K8b: softmax + best class over the first C-1 logits + threshold + person filter + cxcywh -> xyxy -> pixel xywh +
// foot point + per-frame stable compaction
int launch_postprocess(const float* logits, const float* boxes, int B, int Q, int C, int H0, int W0, float threshold,
                       int person_label, float* scores, int32_t* labels, float* xyxy, double* det_xywh, float* det_score,
                       double* det_foot, int32_t* det_query, int32_t* n_keep, int32_t* det_slot, int slot_base,
                       cudaStream_t s);

}  // namespace opd
