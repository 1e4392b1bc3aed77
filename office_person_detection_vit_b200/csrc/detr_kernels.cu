// The non-GEMM kernels of the DETR-ResNet-50 detector: preprocessing (K1), uint8 resize, max pooling (K3), sine
// positional embedding (K7), multi-head attention (K6), prediction heads and post-processing (K8).
// Reference arithmetic: third-party `transformers` (file:line map in oracle/detr_oracle.py) and the reference's
// src/detection/yolov8_detector.py:210-225, 229-241 (xyxy -> xywh, foot point).
#include <algorithm>

#include "detr_kernels.h"

#include <cmath>

#include "opd_common.h"
#include "tc_gemm.h"

namespace opd {

namespace {

constexpr int kD = 256;        // d_model
constexpr int kHeadDim = 32;

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float lo_f(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float hi_f(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// ------------------------------------------------------------------------------------------------------------
// K1 preprocess.  One thread per (b, y, x, kw): 12 normalised values + 4 zeros = 32 bytes, fully coalesced.
// normalise: (u8 - 255*mean) / (255*std) in fp32 (image_processing_backends.py:308-331), then bf16.
// ------------------------------------------------------------------------------------------------------------
__global__ void preprocess_kernel(const uint8_t* __restrict__ src, int B, int Hs, int Ws, int bgr,
                                  __nv_bfloat16* __restrict__ x2, int H2, int W2) {
  const float mean[3] = {0.485f * 255.0f, 0.456f * 255.0f, 0.406f * 255.0f};
  const float stdv[3] = {0.229f * 255.0f, 0.224f * 255.0f, 0.225f * 255.0f};
  const long long total = (long long)B * H2 * W2 * 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int kw = (int)(i & 3);
    long long r = i >> 2;
    const int x = (int)(r % W2);
    r /= W2;
    const int y = (int)(r % H2);
    const int b = (int)(r / H2);
    const int sx = x + kw - 2;   // space-to-depth column this lane group holds
    float v[12];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int iy = 2 * y + dy, ix = 2 * sx + dx;
        const bool ok = sx >= 0 && iy < Hs && ix < Ws && ix >= 0;
        const uint8_t* px = src + (((long long)b * Hs + (ok ? iy : 0)) * Ws + (ok ? ix : 0)) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float u = (float)px[bgr ? 2 - c : c];
          v[(dy * 2 + dx) * 3 + c] = ok ? (u - mean[c]) / stdv[c] : 0.f;
        }
      }
    uint4 o0, o1;
    o0.x = pack2(v[0], v[1]); o0.y = pack2(v[2], v[3]); o0.z = pack2(v[4], v[5]); o0.w = pack2(v[6], v[7]);
    o1.x = pack2(v[8], v[9]); o1.y = pack2(v[10], v[11]); o1.z = 0u; o1.w = 0u;
    uint4* dst = reinterpret_cast<uint4*>(x2 + i * 16);
    dst[0] = o0;
    dst[1] = o1;
  }
}

// ------------------------------------------------------------------------------------------------------------
// uint8 antialias bilinear resize (separable, fixed point), ATen UpSampleKernel.cpp restated:
//   out = clamp((2^(p-1) + sum_j src[x0 + j] * w[j]) >> p, 0, 255), horizontal pass first, then vertical.
// ------------------------------------------------------------------------------------------------------------
__global__ void resize_h_kernel(const uint8_t* __restrict__ src, int B, int H0, int W0, int bgr,
                                uint8_t* __restrict__ tmp, int W1, const int16_t* __restrict__ wx,
                                const int32_t* __restrict__ x0, int kx, int px) {
  const long long total = (long long)B * H0 * W1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % W1);
    const long long row = i / W1;   // b * H0 + y
    const uint8_t* s = src + (row * W0 + x0[ox]) * 3;
    int acc[3] = {1 << (px - 1), 1 << (px - 1), 1 << (px - 1)};
    for (int j = 0; j < kx; ++j) {
      const int w = wx[ox * kx + j];
      if (w == 0) continue;       // padding taps may point past the row end
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[c] += (int)s[j * 3 + c] * w;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int v = min(max(acc[bgr ? 2 - c : c] >> px, 0), 255);
      tmp[i * 3 + c] = (uint8_t)v;   // RGB from here on
    }
  }
}

__global__ void resize_v_kernel(const uint8_t* __restrict__ tmp, int B, int H0, int W1, uint8_t* __restrict__ dst,
                                int H1, const int16_t* __restrict__ wy, const int32_t* __restrict__ y0, int ky, int py,
                                long long dst_frame_stride, long long dst_row_pitch) {
  const long long total = (long long)B * H1 * W1 * 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int xc = (int)(i % (W1 * 3));
    long long r = i / (W1 * 3);
    const int oy = (int)(r % H1);
    const int b = (int)(r / H1);
    const uint8_t* s = tmp + ((long long)b * H0 + y0[oy]) * W1 * 3 + xc;
    int acc = 1 << (py - 1);
    for (int j = 0; j < ky; ++j) {
      const int w = wy[oy * ky + j];
      if (w != 0) acc += (int)s[(long long)j * W1 * 3] * w;
    }
    dst[b * dst_frame_stride + oy * dst_row_pitch + xc] = (uint8_t)min(max(acc >> py, 0), 255);
  }
}

// frames that need no resize, into the top-left corner of their canvas (BGR -> RGB on the way)
__global__ void copy_into_canvas_kernel(const uint8_t* __restrict__ src, int B, int H, int W, int bgr, uint8_t* __restrict__ dst,
                                        long long dst_frame_stride, long long dst_row_pitch) {
  const long long total = (long long)B * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const long long r = i / W;
    const int y = (int)(r % H);
    const long long b = r / H;
    const uint8_t* s = src + i * 3;
    uint8_t* d = dst + b * dst_frame_stride + y * dst_row_pitch + x * 3;
    d[0] = s[bgr ? 2 : 0];
    d[1] = s[1];
    d[2] = s[bgr ? 0 : 2];
  }
}

// ------------------------------------------------------------------------------------------------------------
// K3 max pooling 3x3 / s2 / p1 on NHWC bf16; one thread per (pixel, 8 channels)
// ------------------------------------------------------------------------------------------------------------
__global__ void maxpool_kernel(const __nv_bfloat16* __restrict__ x, int B, int H, int W, int C,
                               __nv_bfloat16* __restrict__ y, int P, int Q) {
  const int c8 = C / 8;
  const long long total = (long long)B * P * Q * c8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % c8);
    long long r = i / c8;
    const int q = (int)(r % Q);
    r /= Q;
    const int p = (int)(r % P);
    const int b = (int)(r / P);
    __nv_bfloat162 m[4];
    const __nv_bfloat162 ninf = __floats2bfloat162_rn(-INFINITY, -INFINITY);
#pragma unroll
    for (int k = 0; k < 4; ++k) m[k] = ninf;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int iy = 2 * p - 1 + dy;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int ix = 2 * q - 1 + dx;
        if (ix < 0 || ix >= W) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + (((long long)b * H + iy) * W + ix) * C) + cg);
        const __nv_bfloat162* vv = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int k = 0; k < 4; ++k) m[k] = __hmax2(m[k], vv[k]);
      }
    }
    uint4 o;
    o.x = *reinterpret_cast<uint32_t*>(&m[0]);
    o.y = *reinterpret_cast<uint32_t*>(&m[1]);
    o.z = *reinterpret_cast<uint32_t*>(&m[2]);
    o.w = *reinterpret_cast<uint32_t*>(&m[3]);
    reinterpret_cast<uint4*>(y + (((long long)b * P + p) * Q + q) * C)[cg] = o;
  }
}

// ------------------------------------------------------------------------------------------------------------
// K7 sine positional embedding (modeling_detr.py:322-349, all-ones mask): pos[y*w + x, 0:128] from y, [128:256] from x
// ------------------------------------------------------------------------------------------------------------
__global__ void pos_embed_kernel(float* __restrict__ pos, int h, int w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= h * w * kD) return;
  const int c = i % kD, t = i / kD;
  const int x = t % w, y = t / w;
  const bool from_y = c < 128;
  const int k = from_y ? c : c - 128;
  const float scale = 6.283185307179586f;   // 2*pi rounded to float32, like torch's python-float * tensor
  const float embed = from_y ? ((float)(y + 1) / ((float)h + 1e-6f)) * scale : ((float)(x + 1) / ((float)w + 1e-6f)) * scale;
  // dim_t = 10000 ** (2 * (k // 2) / 128)   (float32 pow of a float32 exponent)
  const float dim_t = powf(10000.0f, (float)(2 * (k / 2)) / 128.0f);
  const float v = embed / dim_t;
  pos[i] = (k & 1) ? cosf(v) : sinf(v);
}

// Padded batches (modeling_detr.py:322-349 with a real mask): frame b's valid feature cells are the top-left fvalid[b] = (fh, fw)
// rectangle of the h x w map.  y_embed = cumsum of the mask along y = min(y + 1, fh) in a valid column, 0 elsewhere, normalised by
// its last row (fh, or 0 -> 0 / eps = 0); x_embed likewise.  pos [B, h * w, 256].
__global__ void pos_embed_masked_kernel(float* __restrict__ pos, int B, int h, int w, const int32_t* __restrict__ fvalid) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * h * w * kD) return;
  const int c = (int)(i % kD);
  const long long t = i / kD;
  const int x = (int)(t % w), y = (int)((t / w) % h), b = (int)(t / ((long long)w * h));
  const int fh = fvalid[2 * b], fw = fvalid[2 * b + 1];
  const bool from_y = c < 128;
  const int k = from_y ? c : c - 128;
  const float scale = 6.283185307179586f;
  float num, den;
  if (from_y) {
    num = x < fw ? (float)min(y + 1, fh) : 0.f;
    den = x < fw ? (float)fh : 0.f;
  } else {
    num = y < fh ? (float)min(x + 1, fw) : 0.f;
    den = y < fh ? (float)fw : 0.f;
  }
  const float embed = (num / (den + 1e-6f)) * scale;
  const float dim_t = powf(10000.0f, (float)(2 * (k / 2)) / 128.0f);
  const float v = embed / dim_t;
  pos[i] = (k & 1) ? cosf(v) : sinf(v);
}

// ------------------------------------------------------------------------------------------------------------
// K6 attention.  CTA = 4 warps x 16 query rows; keys/values stream through shared memory in tiles of 64 via
// cp.async double buffering; S = Q K^T and O += P V on mma.sync m16n8k16 (bf16 in, fp32 accumulate); online
// softmax in fp32 with exp2.  (A tcgen05 version is the next step for this kernel; see DESIGN.md.)
// ------------------------------------------------------------------------------------------------------------
constexpr int kKvTile = 64;
constexpr int kKvPitch = 40;   // bf16 elements per smem row (80 B): conflict-free fragment loads

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem, bool valid) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 16 : 0;   // src-size 0 -> 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(s));
}

__global__ void __launch_bounds__(128) attention_kernel(const __nv_bfloat16* __restrict__ q, long long ldq,
                                                        const __nv_bfloat16* __restrict__ k, long long ldk,
                                                        const __nv_bfloat16* __restrict__ v, long long ldv,
                                                        __nv_bfloat16* __restrict__ o, long long ldo, int Lq, int Lk) {
  __shared__ __align__(16) __nv_bfloat16 s_k[2][kKvTile][kKvPitch];
  __shared__ __align__(16) __nv_bfloat16 s_v[2][kKvTile][kKvPitch];

  const int head = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * 64 + warp * 16;

  const __nv_bfloat16* qb = q + (long long)b * Lq * ldq + head * kHeadDim;
  const __nv_bfloat16* kb = k + (long long)b * Lk * ldk + head * kHeadDim;
  const __nv_bfloat16* vb = v + (long long)b * Lk * ldv + head * kHeadDim;

  // Q fragments: rows q0+g and q0+g+8 (clamped), two k-steps of 16 dims
  uint32_t qa[2][4];
  {
    const int r0 = min(q0 + g, Lq - 1), r1 = min(q0 + g + 8, Lq - 1);
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      qa[kk][0] = *reinterpret_cast<const uint32_t*>(qb + (long long)r0 * ldq + kk * 16 + 2 * t);
      qa[kk][1] = *reinterpret_cast<const uint32_t*>(qb + (long long)r1 * ldq + kk * 16 + 2 * t);
      qa[kk][2] = *reinterpret_cast<const uint32_t*>(qb + (long long)r0 * ldq + kk * 16 + 8 + 2 * t);
      qa[kk][3] = *reinterpret_cast<const uint32_t*>(qb + (long long)r1 * ldq + kk * 16 + 8 + 2 * t);
    }
  }

  auto load_tile = [&](int buf, int key0) {
    // 64 rows x 4 chunks of 16 B for K and for V: 512 chunks, 4 per thread
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int c = threadIdx.x + it * 128;
      const int row = c >> 2, ch = c & 3;
      const bool ok = key0 + row < Lk;
      const long long kr = ok ? key0 + row : 0;
      cp_async_16(&s_k[buf][row][ch * 8], kb + kr * ldk + ch * 8, ok);
      cp_async_16(&s_v[buf][row][ch * 8], vb + kr * ldv + ch * 8, ok);
    }
    cp_async_commit();
  };

  float oacc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) oacc[i][j] = 0.f;
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
  const float sl2 = 0.17677669529663687f * 1.4426950408889634f;   // 1/sqrt(32) * log2(e)

  const int n_tiles = (Lk + kKvTile - 1) / kKvTile;
  load_tile(0, 0);
  for (int tile = 0; tile < n_tiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < n_tiles) {
      load_tile(buf ^ 1, (tile + 1) * kKvTile);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    // S = Q K^T : 8 key tiles of 8
    float sacc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sacc[j][0] = sacc[j][1] = sacc[j][2] = sacc[j][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&s_k[buf][8 * j + g][kk * 16 + 2 * t]);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&s_k[buf][8 * j + g][kk * 16 + 8 + 2 * t]);
        mma_bf16_16816(sacc[j], qa[kk], b0, b1);
      }
    }
    // mask keys past Lk (only the last tile)
    const int key0 = tile * kKvTile;
    if (key0 + kKvTile > Lk) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kc = key0 + 8 * j + 2 * t;
        if (kc >= Lk) sacc[j][0] = sacc[j][2] = -INFINITY;
        if (kc + 1 >= Lk) sacc[j][1] = sacc[j][3] = -INFINITY;
      }
    }
    // online softmax (rows g and g+8)
    float mx[2] = {mrow[0], mrow[1]};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx[0] = fmaxf(mx[0], fmaxf(sacc[j][0], sacc[j][1]));
      mx[1] = fmaxf(mx[1], fmaxf(sacc[j][2], sacc[j][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float alpha[2], psum[2] = {0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      alpha[r] = exp2f((mrow[r] - mx[r]) * sl2);   // first tile: exp2(-inf) = 0
      mrow[r] = mx[r];
    }
    uint32_t pa[4][4];   // P as A fragments: k-step ks covers keys 16ks..16ks+15
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = exp2f((sacc[j][0] - mx[0]) * sl2), p1 = exp2f((sacc[j][1] - mx[0]) * sl2);
      const float p2 = exp2f((sacc[j][2] - mx[1]) * sl2), p3 = exp2f((sacc[j][3] - mx[1]) * sl2);
      psum[0] += p0 + p1;
      psum[1] += p2 + p3;
      pa[j >> 1][(j & 1) * 2 + 0] = pack2(p0, p1);
      pa[j >> 1][(j & 1) * 2 + 1] = pack2(p2, p3);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) lrow[r] = lrow[r] * alpha[r] + psum[r];
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      oacc[n][0] *= alpha[0];
      oacc[n][1] *= alpha[0];
      oacc[n][2] *= alpha[1];
      oacc[n][3] *= alpha[1];
    }
    // O += P V : 4 k-steps (16 keys) x 4 dim tiles (8); V fragments by ldmatrix.trans, two dim tiles at a time
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t vb4[4];
        const int mat = lane >> 3, rin = lane & 7;
        ldmatrix_x4_trans(vb4, &s_v[buf][16 * ks + (mat & 1) * 8 + rin][8 * (2 * np + (mat >> 1))]);
        mma_bf16_16816(oacc[2 * np], pa[ks], vb4[0], vb4[1]);
        mma_bf16_16816(oacc[2 * np + 1], pa[ks], vb4[2], vb4[3]);
      }
    }
    __syncthreads();   // everyone is done with `buf` before the next iteration's prefetch overwrites it
  }

  // finalise: row sums across the 4 lanes of a row, normalise, store
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 1);
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 2);
  }
  const float inv0 = 1.f / lrow[0], inv1 = 1.f / lrow[1];
  __nv_bfloat16* ob = o + (long long)b * Lq * ldo + head * kHeadDim;
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    if (q0 + g < Lq)
      *reinterpret_cast<uint32_t*>(ob + (long long)(q0 + g) * ldo + 8 * n + 2 * t) =
          pack2(oacc[n][0] * inv0, oacc[n][1] * inv0);
    if (q0 + g + 8 < Lq)
      *reinterpret_cast<uint32_t*>(ob + (long long)(q0 + g + 8) * ldo + 8 * n + 2 * t) =
          pack2(oacc[n][2] * inv1, oacc[n][3] * inv1);
  }
}

// ------------------------------------------------------------------------------------------------------------
// decoder helpers
// ------------------------------------------------------------------------------------------------------------
__global__ void decoder_init_kernel(__nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ yp,
                                    const float* __restrict__ qpos, int B, int Q) {
  const long long total = (long long)B * Q * kD;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    y[i] = __float2bfloat16(0.f);
    yp[i] = __float2bfloat16(qpos[i % ((long long)Q * kD)]);
  }
}

// one warp per row of 256: fp32 statistics (two-pass in registers)
__global__ void layernorm_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, int rows) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const uint4 raw = *reinterpret_cast<const uint4*>(x + (long long)row * kD + lane * 8);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&raw);
  float v[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[2 * j] = lo_f(w[j]);
    v[2 * j + 1] = hi_f(w[j]);
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += v[j];
#pragma unroll
  for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  const float mean = s * (1.f / kD);
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) sq += (v[j] - mean) * (v[j] - mean);
#pragma unroll
  for (int d = 16; d; d >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, d);
  const float rstd = rsqrtf(sq * (1.f / kD) + 1e-5f);
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = lane * 8 + 2 * j;
    o[j] = pack2((v[2 * j] - mean) * rstd * gamma[c] + beta[c], (v[2 * j + 1] - mean) * rstd * gamma[c + 1] + beta[c + 1]);
  }
  *reinterpret_cast<uint4*>(y + (long long)row * kD + lane * 8) = make_uint4(o[0], o[1], o[2], o[3]);
}

// ------------------------------------------------------------------------------------------------------------
// K8a heads: one CTA (256 threads) per kHeadRows decoder rows; fp32 weights, fp32 math.  Every weight element is loaded
// once per CTA and used for all of its rows (the one-row version re-read 0.6 MB of weights from L2 per row).
//   logits = y Wc^T + bc ; boxes = sigmoid(W2 relu(W1 relu(W0 y + b0) + b1) + b2)   (modeling_detr.py:1275-1297, 1401-1402)
// ------------------------------------------------------------------------------------------------------------
constexpr int kHeadRows = 16;   // 16 rows per CTA: every weight element fetched from L2 serves 16 FMAs (8 rows: 0.15 ms at batch 64, weight-fetch bound)
// acc[r] += x[k][r] * wv for the CTA's rows: the row values of one k sit in 64 consecutive bytes (four LDS.128, broadcast)
__device__ __forceinline__ void head_fma8(float (&acc)[kHeadRows], const float* xk, float wv) {
#pragma unroll
  for (int q = 0; q < kHeadRows / 4; ++q) {
    const float4 a = *reinterpret_cast<const float4*>(xk + 4 * q);
    acc[4 * q + 0] = fmaf(a.x, wv, acc[4 * q + 0]);
    acc[4 * q + 1] = fmaf(a.y, wv, acc[4 * q + 1]);
    acc[4 * q + 2] = fmaf(a.z, wv, acc[4 * q + 2]);
    acc[4 * q + 3] = fmaf(a.w, wv, acc[4 * q + 3]);
  }
}
__global__ void __launch_bounds__(256) heads_kernel(const __nv_bfloat16* __restrict__ y, HeadWeights w,
                                                    float* __restrict__ logits, float* __restrict__ boxes, int n_cls, int rows) {
  __shared__ __align__(16) float s_y[kD][kHeadRows], s_h0[kD][kHeadRows];   // [k][row]
  float (*s_h1)[kHeadRows] = s_y;   // the second hidden layer overwrites the input rows (last read before the barrier above it)
  const int row0 = blockIdx.x * kHeadRows, j = threadIdx.x;
  const int nr = min(kHeadRows, rows - row0);
#pragma unroll
  for (int r = 0; r < kHeadRows; ++r) s_y[j][r] = r < nr ? __bfloat162float(y[(long long)(row0 + r) * kD + j]) : 0.f;
  __syncthreads();
  if (j < n_cls) {
    float acc[kHeadRows] = {};
#pragma unroll 8
    for (int k = 0; k < kD; ++k) head_fma8(acc, s_y[k], w.wc_t[k * n_cls + j]);
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r)
      if (r < nr) logits[(long long)(row0 + r) * n_cls + j] = acc[r] + w.bc[j];
  }
  {
    float acc[kHeadRows] = {};
#pragma unroll 8
    for (int k = 0; k < kD; ++k) head_fma8(acc, s_y[k], w.w0_t[k * kD + j]);
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r) s_h0[j][r] = fmaxf(acc[r] + w.b0[j], 0.f);
  }
  __syncthreads();
  {
    float acc[kHeadRows] = {};
#pragma unroll 8
    for (int k = 0; k < kD; ++k) head_fma8(acc, s_h0[k], w.w1_t[k * kD + j]);
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r) s_h1[j][r] = fmaxf(acc[r] + w.b1[j], 0.f);
  }
  __syncthreads();
  // 8 warps x 4 outputs: a warp per row (two rows each), each lane strides the 256 hidden values
  const int warp = j >> 5, lane = j & 31;
  for (int r = warp; r < nr; r += 8) {
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      float acc = 0.f;
      for (int k = lane; k < kD; k += 32) acc = fmaf(s_h1[k][r], w.w2[o * kD + k], acc);
#pragma unroll
      for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
      if (lane == 0) boxes[(long long)(row0 + r) * 4 + o] = 1.f / (1.f + expf(-(acc + w.b2[o])));
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// ROI features of the detections (detect_with_features / extract_features): mean of the encoder map over the box,
// L2-normalised.  Restates FeatureExtractor.extract_roi_features + normalize_features
// (src/tracking/feature_extractor.py:39-88, :21-37): box -> feature-map cells with Python-float arithmetic and int()
// truncation, clamps, mean over [y_min:y_max, x_min:x_max], v / (|v| + 1e-8).
// One CTA per (detection row, frame), thread = channel (coalesced 2-byte loads of one cell's D channels).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) roi_features_kernel(const __nv_bfloat16* __restrict__ feat, int fh, int fw, int D,
                                                            const double* __restrict__ xywh, const int32_t* __restrict__ n_keep,
                                                            int Q, double img_h, double img_w, float* __restrict__ out) {
  __shared__ float s_part[32];
  const int b = blockIdx.y, r = blockIdx.x, c = threadIdx.x;
  float* o = out + ((long long)b * Q + r) * D;
  if (r >= n_keep[b]) {
    if (c < D) o[c] = 0.f;
    return;
  }
  const double* bb = xywh + ((long long)b * Q + r) * 4;
  const double x = bb[0], y = bb[1], w = bb[2], h = bb[3];
  int x_min = (int)((x / img_w) * fw), y_min = (int)((y / img_h) * fh);
  int x_max = (int)(((x + w) / img_w) * fw), y_max = (int)(((y + h) / img_h) * fh);
  x_min = max(0, min(x_min, fw - 1));
  y_min = max(0, min(y_min, fh - 1));
  x_max = max(x_min + 1, min(x_max, fw));
  y_max = max(y_min + 1, min(y_max, fh));
  float acc = 0.f;
  if (c < D) {
    const __nv_bfloat16* base = feat + (long long)b * fh * fw * D + c;
    for (int yy = y_min; yy < y_max; ++yy)
      for (int xx = x_min; xx < x_max; ++xx) acc += __bfloat162float(base[((long long)yy * fw + xx) * D]);
    acc /= (float)((y_max - y_min) * (x_max - x_min));
  }
  float sq = acc * acc;
#pragma unroll
  for (int d = 16; d; d >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, d);
  if ((c & 31) == 0) s_part[c >> 5] = sq;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += s_part[i];
  if (c < D) o[c] = acc / (sqrtf(tot) + 1e-8f);
}

// ------------------------------------------------------------------------------------------------------------
// K8b post-processing: one CTA of 128 threads per frame, thread = query
//   image_processing_detr.py:826-843 (softmax, best of the first C-1 classes, cxcywh -> xyxy, scale to pixels, threshold)
//   + person filter, xyxy -> xywh and foot point (x + w/2, y + h) as in yolov8_detector.py:210-225, 229-241
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) postprocess_kernel(const float* __restrict__ logits, const float* __restrict__ boxes,
                                                          int Q, int C, int H0, int W0, float threshold, int person_label,
                                                          float* __restrict__ scores, int32_t* __restrict__ labels,
                                                          float* __restrict__ xyxy, double* __restrict__ det_xywh,
                                                          float* __restrict__ det_score, double* __restrict__ det_foot,
                                                          int32_t* __restrict__ det_query, int32_t* __restrict__ n_keep,
                                                          int32_t* __restrict__ det_slot, int slot_base) {
  __shared__ int s_warp_count[4];
  const int b = blockIdx.x, qi = threadIdx.x;
  const int lane = qi & 31, warp = qi >> 5;
  bool keep = false;
  float sc = 0.f, x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f;
  if (qi < Q) {
    const float* lg = logits + ((long long)b * Q + qi) * C;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, lg[c]);
    float sum = 0.f;
    for (int c = 0; c < C; ++c) sum += expf(lg[c] - mx);
    int best = 0;
    float bestv = -INFINITY;
    for (int c = 0; c < C - 1; ++c)
      if (lg[c] > bestv) {   // first maximum, like torch.max
        bestv = lg[c];
        best = c;
      }
    sc = expf(bestv - mx) / sum;
    const float* bx = boxes + ((long long)b * Q + qi) * 4;
    const float cx = bx[0], cy = bx[1], bw = bx[2], bh = bx[3];
    x1 = (cx - 0.5f * bw) * (float)W0;
    y1 = (cy - 0.5f * bh) * (float)H0;
    x2 = (cx + 0.5f * bw) * (float)W0;
    y2 = (cy + 0.5f * bh) * (float)H0;
    const long long r = (long long)b * Q + qi;
    scores[r] = sc;
    labels[r] = best;
    xyxy[r * 4 + 0] = x1; xyxy[r * 4 + 1] = y1; xyxy[r * 4 + 2] = x2; xyxy[r * 4 + 3] = y2;
    keep = sc > threshold && best == person_label;
  }
  const unsigned bal = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) s_warp_count[warp] = __popc(bal);
  __syncthreads();
  int base = 0;
  for (int wq = 0; wq < warp; ++wq) base += s_warp_count[wq];
  if (keep) {
    const int slot = base + __popc(bal & ((1u << lane) - 1));
    const long long r = (long long)b * Q + slot;
    const double dx1 = x1, dy1 = y1, dw = (double)x2 - (double)x1, dh = (double)y2 - (double)y1;
    det_xywh[r * 4 + 0] = dx1;   // Python floats in the reference (yolov8_detector.py:214-218): float64, x2 - x1 exact
    det_xywh[r * 4 + 1] = dy1;
    det_xywh[r * 4 + 2] = dw;
    det_xywh[r * 4 + 3] = dh;
    det_score[r] = sc;
    det_foot[r * 2 + 0] = dx1 + dw / 2.0;   // python floats in the reference: float64
    det_foot[r * 2 + 1] = dy1 + dh;
    det_query[r] = qi;
  }
  const int total = s_warp_count[0] + s_warp_count[1] + s_warp_count[2] + s_warp_count[3];
  if (qi == 0) n_keep[b] = total;
  // histogram row of every compacted detection row (-1 = unused row, skipped by the floor kernels)
  if (det_slot && qi < Q) det_slot[(long long)b * Q + qi] = qi < total ? slot_base + b : -1;
}

int grid_for(long long total, int threads) {
  long long blocks = (total + threads - 1) / threads;
  const long long cap = 148LL * 16;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

// Synthetic timelapse frames generated on the device (bench.py config 4, SURVEY.md §8d: "frames generated on device per rank
// from seed = 1000 + global_frame_idx // 64", no host I/O in the timed region).  Same picture family as
// detection/synthetic.py synthetic_frames (192-pixel flat-colour blocks + 32-pixel blocks + pixel noise around 128), from a
// counter-based hash so that any frame can be regenerated anywhere: value(frame g, y, x, c) =
//   clamp(128 + (h(0, y/192, x/192, c) % 221 - 110) + (h(1, y/32, x/32, c) % 81 - 40) + (h(2, y, x, c) % 25 - 12), 0, 255)
// with h(level, yy, xx, c) = high 32 bits of mix64((seed_base + g / 64) * C1 + (g % 64) * C2 + key * C3), key =
// ((level * 4096 + yy) * 4096 + xx) * 4 + c.  detection/synthetic.py device_frames_reference() restates it in NumPy.
__device__ __forceinline__ uint32_t synth_hash(unsigned long long base, uint32_t level, uint32_t yy, uint32_t xx, uint32_t c) {
  const unsigned long long key = ((((unsigned long long)level * 4096ull + yy) * 4096ull + xx) * 4ull + c);
  unsigned long long z = base + key * 0x94D049BB133111EBull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (uint32_t)(z >> 32);
}

// one thread = 16 consecutive output bytes (one aligned 16-byte store); the pixel cursor (frame, y, x, channel) advances byte by
// byte and the two block-level hashes are recomputed only when the cursor enters another 32-pixel block, row or frame
__global__ void synthetic_frames_kernel(unsigned long long seed_base, long long frame0, int H, int W, uint8_t* __restrict__ out,
                                        long long n_bytes) {
  const long long n_chunks = (n_bytes + 15) / 16;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_chunks; i += (long long)gridDim.x * blockDim.x) {
    const long long byte0 = i * 16;
    long long pix = byte0 / 3;
    int c = (int)(byte0 - pix * 3);
    long long row = pix / W;
    int x = (int)(pix - row * W);
    long long b = row / H;
    int y = (int)(row - b * H);
    unsigned long long base = 0;
    int coarse[3] = {0, 0, 0};
    bool stale = true;
    union { uint8_t u8[16]; uint4 v; } buf;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (stale) {
        const long long g = frame0 + b;
        base = (seed_base + (unsigned long long)(g / 64)) * 0x9E3779B97F4A7C15ull + (unsigned long long)(g % 64) * 0xD1B54A32D192ED03ull;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
          coarse[cc] = 128 + (int)(synth_hash(base, 0, y / 192, x / 192, cc) % 221u) - 110 + (int)(synth_hash(base, 1, y / 32, x / 32, cc) % 81u) - 40;
        stale = false;
      }
      const int v = coarse[c] + (int)(synth_hash(base, 2, y, x, c) % 25u) - 12;
      buf.u8[k] = (uint8_t)min(max(v, 0), 255);
      if (++c == 3) {
        c = 0;
        if (++x == W) {
          x = 0;
          stale = true;
          if (++y == H) {
            y = 0;
            ++b;
          }
        } else if ((x & 31) == 0) {
          stale = true;
        }
      }
    }
    if (byte0 + 16 <= n_bytes) {
      *reinterpret_cast<uint4*>(out + byte0) = buf.v;
    } else {
      for (int k = 0; byte0 + k < n_bytes; ++k) out[byte0 + k] = buf.u8[k];
    }
  }
}

}  // namespace

int launch_synthetic_frames(unsigned long long seed_base, long long frame0, int B, int H, int W, uint8_t* out, cudaStream_t s) {
  OPD_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "synthetic frames: the output must be 16-byte aligned");
  const long long n = (long long)B * H * W * 3;
  const unsigned blocks = (unsigned)std::min<long long>(((n + 15) / 16 + 255) / 256, 148LL * 16);
  synthetic_frames_kernel<<<blocks, 256, 0, s>>>(seed_base, frame0, H, W, out, n);
  count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

int launch_preprocess(const uint8_t* src, int B, int Hs, int Ws, int src_is_bgr, __nv_bfloat16* x2, cudaStream_t s) {
  const int H2 = (Hs + 1) / 2, W2 = (Ws + 1) / 2;
  const long long total = (long long)B * H2 * W2 * 4;
  preprocess_kernel<<<grid_for(total, 256), 256, 0, s>>>(src, B, Hs, Ws, src_is_bgr, x2, H2, W2);
  count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

int launch_copy_into_canvas(const uint8_t* src, int B, int H, int W, int src_is_bgr, uint8_t* dst, long long dst_frame_stride,
                            long long dst_row_pitch, cudaStream_t s) {
  copy_into_canvas_kernel<<<grid_for((long long)B * H * W, 256), 256, 0, s>>>(src, B, H, W, src_is_bgr, dst, dst_frame_stride, dst_row_pitch);
  count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

int launch_resize_u8(const uint8_t* src, int B, int H0, int W0, int src_is_bgr, uint8_t* tmp, uint8_t* dst, int H1,
                     int W1, const int16_t* wx, const int32_t* x0, int kx, int px, const int16_t* wy, const int32_t* y0,
                     int ky, int py, cudaStream_t s, long long dst_frame_stride, long long dst_row_pitch) {
  if (dst_frame_stride == 0) dst_frame_stride = (long long)H1 * W1 * 3;
  if (dst_row_pitch == 0) dst_row_pitch = (long long)W1 * 3;
  resize_h_kernel<<<grid_for((long long)B * H0 * W1, 256), 256, 0, s>>>(src, B, H0, W0, src_is_bgr, tmp, W1, wx, x0, kx, px);
  count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  resize_v_kernel<<<grid_for((long long)B * H1 * W1 * 3, 256), 256, 0, s>>>(tmp, B, H0, W1, dst, H1, wy, y0, ky, py, dst_frame_stride,
                                                                           dst_row_pitch);
  count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

int launch_maxpool(const __nv_bfloat16* x, int B, int H, int W, int C, __nv_bfloat16* y, int P, int Q, cudaStream_t s) {
  OPD_REQUIRE(C % 8 == 0, "maxpool: C=%d must be a multiple of 8", C);
  const long long total = (long long)B * P * Q * (C / 8);
  maxpool_kernel<<<grid_for(total, 256), 256, 0, s>>>(x, B, H, W, C, y, P, Q);
  count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

int launch_pos_embed(float* pos, int h, int w, cudaStream_t s) {
  const int total = h * w * kD;
  pos_embed_kernel<<<(total + 255) / 256, 256, 0, s>>>(pos, h, w);
  count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

int launch_pos_embed_masked(float* pos, int B, int h, int w, const int32_t* fvalid, cudaStream_t s) {
  const long long total = (long long)B * h * w * kD;
  pos_embed_masked_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(pos, B, h, w, fvalid);
  count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

int launch_attention(const __nv_bfloat16* q, int64_t ldq, const __nv_bfloat16* k, int64_t ldk, const __nv_bfloat16* v,
                     int64_t ldv, __nv_bfloat16* o, int64_t ldo, int B, int heads, int Lq, int Lk, cudaStream_t s) {
  OPD_REQUIRE(Lq > 0 && Lk > 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldq % 2 == 0 && ldo % 2 == 0, "attention: bad strides");
  dim3 grid((Lq + 63) / 64, heads, B);
  attention_kernel<<<grid, 128, 0, s>>>(q, ldq, k, ldk, v, ldv, o, ldo, Lq, Lk);
  count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

int launch_decoder_init(__nv_bfloat16* y, __nv_bfloat16* yp, const float* qpos, int B, int Q, cudaStream_t s) {
  decoder_init_kernel<<<grid_for((long long)B * Q * kD, 256), 256, 0, s>>>(y, yp, qpos, B, Q);
  count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

int launch_roi_features(const __nv_bfloat16* feat, int B, int fh, int fw, int D, const double* xywh, const int32_t* n_keep,
                        int Q, int img_h, int img_w, float* out, cudaStream_t s) {
  OPD_REQUIRE(D > 0 && D <= 1024 && D % 32 == 0, "roi features: D=%d must be a multiple of 32, at most 1024", D);
  roi_features_kernel<<<dim3(Q, B), D, 0, s>>>(feat, fh, fw, D, xywh, n_keep, Q, (double)img_h, (double)img_w, out);
  count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

int launch_layernorm(const __nv_bfloat16* x, const float* gamma, const float* beta, __nv_bfloat16* y, int rows,
                     cudaStream_t s) {
  layernorm_kernel<<<(rows + 7) / 8, 256, 0, s>>>(x, gamma, beta, y, rows);
  count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

int launch_heads(const __nv_bfloat16* y, const HeadWeights& w, float* logits, float* boxes, int rows, cudaStream_t s) {
  heads_kernel<<<(rows + kHeadRows - 1) / kHeadRows, 256, 0, s>>>(y, w, logits, boxes, 92, rows);
  count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

int launch_postprocess(const float* logits, const float* boxes, int B, int Q, int C, int H0, int W0, float threshold,
                       int person_label, float* scores, int32_t* labels, float* xyxy, double* det_xywh, float* det_score,
                       double* det_foot, int32_t* det_query, int32_t* n_keep, int32_t* det_slot, int slot_base,
                       cudaStream_t s) {
  OPD_REQUIRE(Q <= 128, "postprocess: at most 128 queries per frame (got %d)", Q);
  postprocess_kernel<<<B, 128, 0, s>>>(logits, boxes, Q, C, H0, W0, threshold, person_label, scores, labels, xyxy,
                                       det_xywh, det_score, det_foot, det_query, n_keep, det_slot, slot_base);
  count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

}  // namespace opd

extern "C" int opd_attention_bf16(const void* q_dev, int64_t ldq, const void* k_dev, int64_t ldk, const void* v_dev,
                                  int64_t ldv, void* o_dev, int64_t ldo, int32_t B, int32_t heads, int32_t Lq,
                                  int32_t Lk, void* stream) {
  if (opd::g_option_attention_tc.load()) {
    opd::AttnPlan plan;
    if (int rc = opd::attn_plan(&plan, static_cast<const __nv_bfloat16*>(q_dev), ldq, static_cast<const __nv_bfloat16*>(k_dev), ldk,
                                static_cast<const __nv_bfloat16*>(v_dev), ldv, static_cast<__nv_bfloat16*>(o_dev), ldo, B, heads, Lq, Lk))
      return rc;
    return opd::attn_launch(plan, static_cast<cudaStream_t>(stream));
  }
  return opd::launch_attention(static_cast<const __nv_bfloat16*>(q_dev), ldq, static_cast<const __nv_bfloat16*>(k_dev),
                               ldk, static_cast<const __nv_bfloat16*>(v_dev), ldv, static_cast<__nv_bfloat16*>(o_dev),
                               ldo, B, heads, Lq, Lk, static_cast<cudaStream_t>(stream));
}

extern "C" int opd_synthetic_frames_u8(uint64_t seed_base, int64_t global_frame0, int32_t B, int32_t H, int32_t W, uint8_t* frames_dev,
                                       void* stream) {
  OPD_REQUIRE(frames_dev && B > 0 && H > 0 && W > 0 && H <= 4096 * 32 && W <= 4096 * 32 && global_frame0 >= 0,
              "opd_synthetic_frames_u8: bad argument");
  return opd::launch_synthetic_frames(seed_base, global_frame0, B, H, W, frames_dev, static_cast<cudaStream_t>(stream));
}

extern "C" int opd_roi_features_bf16(const void* feat_dev, int32_t B, int32_t fh, int32_t fw, int32_t D, const double* det_xywh_dev,
                                     const int32_t* n_keep_dev, int32_t Q, int32_t img_h, int32_t img_w, float* out_dev,
                                     void* stream) {
  OPD_REQUIRE(feat_dev && det_xywh_dev && n_keep_dev && out_dev, "opd_roi_features_bf16: NULL argument");
  OPD_REQUIRE(B > 0 && fh > 0 && fw > 0 && Q > 0 && img_h > 0 && img_w > 0, "opd_roi_features_bf16: bad shape");
  return opd::launch_roi_features(static_cast<const __nv_bfloat16*>(feat_dev), B, fh, fw, D, det_xywh_dev, n_keep_dev, Q, img_h,
                                  img_w, out_dev, static_cast<cudaStream_t>(stream));
}
