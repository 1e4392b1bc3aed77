// Fused transformer feed-forward block on the 5th-generation tensor cores:
//     y  = LayerNorm( relu(x W1^T + b1) W2^T + b2 + x ) * gamma + beta          x, y [M, 256], hidden width 2048
//     y2 = bf16(y + pos[row % pos_rows])                                         (query / key input of the next attention)
// i.e. DetrEncoderLayer / DetrDecoderLayer: mlp.fc1 -> activation -> mlp.fc2 -> residual -> final_layer_norm (transformers
// models/detr/modeling_detr.py:560-640, 643-760; ReLU, post-norm), in ONE persistent kernel: the [M, 2048] hidden activations
// (275 MB per encoder layer at batch 64) are neither written nor read back.  As two GEMM launches, fc1 ran at the WRITE roofline of
// HBM (3.9 TB/s) and fc2 re-read what it had written: 79 + 98 us per encoder layer, 19 + 36 us per decoder layer.
//
// Per 128-row tile (persistent CTAs, static schedule), for the eight 256-wide chunks c of the hidden layer:
//   G1(c): acc1[128 x 256] = X[128 x 256] * W1[c]^T           X resident in shared memory (also the LayerNorm residual)
//   E1(c): H = bf16(relu(acc1 + b1[c]))  -> shared memory      four K-chunks [128 x 64]; G2 starts on a chunk as soon as it is written
//   G2(c): acc2[128 x 256] += H * W2[:, c]^T
//   E2   : LayerNorm(acc2 + b2 + X) -> y (staged over H), y + pos -> y2 (staged over X), TMA stores
// Weight tiles [128 rows x 64 k] stream through one ring in the order the MMA thread consumes them (W1[0]; W2[0], W1[1]; ...).
// Every output element is accumulated in the order of the two-launch path (k ascending, fp32 in TMEM) and both epilogues apply
// the same operations in the same order, so the results are bit-identical to gemm(EPI_BIAS_RELU) + gemm(EPI_BIAS_RES_LN)
// (tests/test_tc_ops_gpu.py).
// Warps: 0-7 epilogue (two warpgroups: warpgroup g owns the K-chunks / column chunks g and g + 2 of every step), 8 TMA producer,
// 9 MMA issuer + TMEM owner.  TMEM: acc1 = columns [0, 256), acc2 = [256, 512).
#include <algorithm>

#include "opd_common.h"
#include "sm100_ptx.cuh"
#include "tc_gemm.h"

namespace opd {
namespace {

constexpr int BLOCK_M = 128, kDm = 256, kHidden = 2048, kChunk = 256, kNC = kHidden / kChunk;
constexpr int BLOCK_K = 64, UMMA_K = 16;
constexpr int CHUNK_BYTES = BLOCK_M * 64 * 2;        // [128 x 64] bf16, 128-byte swizzle
constexpr int B_STAGE_BYTES = 128 * 64 * 2;          // half a weight tile: 128 output rows x 64 k
constexpr int kStages = 5;
constexpr int kThreads = 320;
constexpr int kSmemBytes = 8 * CHUNK_BYTES + kStages * B_STAGE_BYTES + (kHidden + 3 * kDm) * 4 + 4096 + 256;

struct MlpParams {
  CUtensorMap tmX, tmW1, tmW2, tmD, tmD2;
  int M, num_tiles;
  const float* b1;
  const float* b2;
  const float* gamma;
  const float* beta;
  const float* pos;
  int pos_rows, pos_row0, has_d2;
};

__global__ void __launch_bounds__(kThreads, 1) tc_mlp_kernel(const __grid_constant__ MlpParams p) {
  constexpr uint32_t kIdesc = ptx::umma_idesc_bf16(BLOCK_M, 128);
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_a1 = smem;                              // X tile: 4 K-chunks; E2: residual, then y2 staging
  uint8_t* smem_a2 = smem_a1 + 4 * CHUNK_BYTES;         // H chunk: 4 K-chunks; E2: y staging
  uint8_t* smem_b = smem_a2 + 4 * CHUNK_BYTES;          // weight ring
  float* s_b1 = reinterpret_cast<float*>(smem_b + kStages * B_STAGE_BYTES);   // [2048]
  float* s_b2 = s_b1 + kHidden;                                               // [256]
  float* s_gamma = s_b2 + kDm;
  float* s_beta = s_gamma + kDm;
  float* s_stat = s_beta + kDm;                                               // LayerNorm partials [2 tiles][2 warpgroups][128 rows][2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stat + 1024);
  uint64_t* full_bar = bars;            // [kStages]
  uint64_t* empty_bar = bars + 5;       // [kStages]
  uint64_t* a1_full = bars + 10;
  uint64_t* a1_free = bars + 11;        // both warpgroups' stores out of the X / H buffers have been read
  uint64_t* acc1_full = bars + 12;
  uint64_t* acc1_empty = bars + 13;     // 8 warps
  uint64_t* a2_ready = bars + 14;       // [4] K-chunk of H written (4 warps each)
  uint64_t* a2_free = bars + 18;        // G2 has read H
  uint64_t* acc2_full = bars + 19;
  uint64_t* acc2_empty = bars + 20;     // 8 warps
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 21);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  if (warp == 8 && lane == 0) {
    ptx::prefetch_tmap(&p.tmX);
    ptx::prefetch_tmap(&p.tmW1);
    ptx::prefetch_tmap(&p.tmW2);
    ptx::prefetch_tmap(&p.tmD);
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    ptx::mbar_init(a1_full, 1);
    ptx::mbar_init(a1_free, 2);
    ptx::mbar_init(acc1_full, 1);
    ptx::mbar_init(acc1_empty, 8);
    for (int i = 0; i < 4; ++i) ptx::mbar_init(&a2_ready[i], 4);
    ptx::mbar_init(a2_free, 1);
    ptx::mbar_init(acc2_full, 1);
    ptx::mbar_init(acc2_empty, 8);
    ptx::fence_barrier_init();
  }
  if (warp == 9) ptx::tmem_alloc<512>(tmem_ptr);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::grid_dependency_wait();      // programmatic dependent launch: everything above overlaps the previous kernel's tail
  ptx::grid_launch_dependents();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_acc1 = tmem_base, tmem_acc2 = tmem_base + 256;
  const int n_my = (int)blockIdx.x < p.num_tiles ? (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 8) {
    // ===================================== TMA producer =====================================
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      auto ring_load = [&](const CUtensorMap* tm, int c0, int c1) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        ptx::mbar_expect_tx(&full_bar[stage], B_STAGE_BYTES);
        ptx::tma_load_2d(tm, &full_bar[stage], smem_b + stage * B_STAGE_BYTES, c0, c1);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      };
      auto load_w1 = [&](int c) {   // rows = hidden units [256 c, 256 c + 256), k = model columns
        for (int kb = 0; kb < 4; ++kb)
          for (int nh = 0; nh < 2; ++nh) ring_load(&p.tmW1, kb * BLOCK_K, c * kChunk + nh * 128);
      };
      auto load_w2 = [&](int c) {   // rows = the 256 outputs, k = hidden units [256 c, 256 c + 256)
        for (int kb = 0; kb < 4; ++kb)
          for (int nh = 0; nh < 2; ++nh) ring_load(&p.tmW2, c * kChunk + kb * BLOCK_K, nh * 128);
      };
      for (int it = 0; it < n_my; ++it) {
        const int m0 = ((int)blockIdx.x + it * (int)gridDim.x) * BLOCK_M;
        ptx::mbar_wait(a1_free, (it & 1) ^ 1);
        ptx::mbar_expect_tx(a1_full, 4 * CHUNK_BYTES);
        for (int kb = 0; kb < 4; ++kb) ptx::tma_load_2d(&p.tmX, a1_full, smem_a1 + kb * CHUNK_BYTES, kb * BLOCK_K, m0);
        load_w1(0);
        for (int c = 0; c < kNC; ++c) {
          load_w2(c);
          if (c + 1 < kNC) load_w1(c + 1);
        }
      }
    }
  } else if (warp == 9) {
    // ===================================== MMA issuer =====================================
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      // one k-block of a [128 x 256] x [256 x 64]^T product: two ring stages (output halves), four MMAs each
      auto kblock = [&](uint32_t acc, const uint8_t* a_chunk, bool first) {
        const uint64_t da = ptx::umma_desc_kmajor_sw128(ptx::smem_u32(a_chunk));
        for (int nh = 0; nh < 2; ++nh) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after_sync();
          const uint64_t db = ptx::umma_desc_kmajor_sw128(ptx::smem_u32(smem_b + stage * B_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            ptx::umma_bf16_ss(acc + nh * 128, da + 2 * k, db + 2 * k, kIdesc, !(first && k == 0));
          ptx::umma_commit(&empty_bar[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      };
      uint32_t q = 0;   // hidden chunks issued by this CTA (G1 count)
      auto issue_g1 = [&]() {
        ptx::mbar_wait(acc1_empty, (q & 1) ^ 1);
        ptx::tc_fence_after_sync();
        for (int kb = 0; kb < 4; ++kb) kblock(tmem_acc1, smem_a1 + kb * CHUNK_BYTES, kb == 0);
        ptx::umma_commit(acc1_full);
        ++q;
      };
      for (int it = 0; it < n_my; ++it) {
        ptx::mbar_wait(a1_full, it & 1);
        ptx::tc_fence_after_sync();
        issue_g1();
        for (int c = 0; c < kNC; ++c) {
          const uint32_t qc = (uint32_t)it * kNC + c;
          if (c == 0) {
            ptx::mbar_wait(acc2_empty, (it & 1) ^ 1);
            ptx::tc_fence_after_sync();
          }
          for (int kb = 0; kb < 4; ++kb) {
            ptx::mbar_wait(&a2_ready[kb], qc & 1);
            ptx::tc_fence_after_sync();
            kblock(tmem_acc2, smem_a2 + kb * CHUNK_BYTES, c == 0 && kb == 0);
          }
          ptx::umma_commit(a2_free);
          if (c + 1 == kNC) ptx::umma_commit(acc2_full);
          else issue_g1();
        }
      }
    }
  } else if (warp < 8) {
    // ===================================== epilogue warps =====================================
    const int wg = warp >> 2;
    const int et = threadIdx.x - wg * 128;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int bar_id = 1 + wg;
    for (int i = threadIdx.x; i < kHidden; i += 256) s_b1[i] = p.b1[i];
    for (int i = threadIdx.x; i < kDm; i += 256) {
      s_b2[i] = p.b2[i];
      s_gamma[i] = p.gamma[i];
      s_beta[i] = p.beta[i];
    }
    ptx::named_bar_sync(3, 256);
    auto warp_arrive = [&](uint64_t* bar) {
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar);
    };
    for (int it = 0; it < n_my; ++it) {
      const int m0 = ((int)blockIdx.x + it * (int)gridDim.x) * BLOCK_M;
      const long long m = (long long)m0 + row;
      const bool row_ok = m < p.M;
      // ---------------- E1: eight hidden chunks ----------------
      for (int c = 0; c < kNC; ++c) {
        const uint32_t qc = (uint32_t)it * kNC + c;
        ptx::mbar_wait(acc1_full, qc & 1);
        ptx::mbar_wait(a2_free, (qc & 1) ^ 1);     // G2 of the previous chunk has read H (E2's stores out of H: waited for below)
        ptx::tc_fence_after_sync();
        for (int i = wg; i < 4; i += 2) {          // K-chunk i of H = hidden units [64 i, 64 i + 64) of this chunk
          uint32_t packed[32];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(tmem_acc1 + lane_addr + i * 64 + h * 32, v);
            ptx::tmem_ld_wait();
            const float4* bias4 = reinterpret_cast<const float4*>(s_b1 + c * kChunk + i * 64 + h * 32);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 bq = bias4[j4];
              const uint64_t s0 = ptx::add_f32x2(ptx::f32x2(v[4 * j4], v[4 * j4 + 1]), ptx::f32x2(__float_as_uint(bq.x), __float_as_uint(bq.y)));
              const uint64_t s1 = ptx::add_f32x2(ptx::f32x2(v[4 * j4 + 2], v[4 * j4 + 3]), ptx::f32x2(__float_as_uint(bq.z), __float_as_uint(bq.w)));
              packed[h * 16 + 2 * j4] = ptx::relu_bf16x2(ptx::cvt_bf16x2(s0));
              packed[h * 16 + 2 * j4 + 1] = ptx::relu_bf16x2(ptx::cvt_bf16x2(s1));
            }
          }
          uint8_t* rowp = smem_a2 + i * CHUNK_BYTES + row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(rowp + ((j ^ (row & 7)) << 4)) = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          ptx::fence_proxy_async_smem();
          warp_arrive(&a2_ready[i]);
        }
        ptx::tc_fence_before_sync();
        warp_arrive(acc1_empty);
      }
      // ---------------- E2: bias + residual + LayerNorm (+ pos) ----------------
      ptx::mbar_wait(acc2_full, it & 1);
      ptx::tc_fence_after_sync();
      const uint32_t t_acc = tmem_acc2 + lane_addr;
      float sum = 0.f, sq = 0.f;
      for (int c = wg; c < 4; c += 2) {
        const uint8_t* resp = smem_a1 + c * CHUNK_BYTES + row * 128;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int g = c * 2 + h;
          uint32_t v[32];
          ptx::tmem_ld_32x32(t_acc + g * 32, v);
          uint4 rr[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) rr[j] = *reinterpret_cast<const uint4*>(resp + (((h * 4 + j) ^ (row & 7)) << 4));
          ptx::tmem_ld_wait();
          const uint32_t* rw = reinterpret_cast<const uint32_t*>(rr);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float a = __uint_as_float(v[2 * j]) + s_b2[g * 32 + 2 * j] + ptx::bf16_lo(rw[j]);
            const float b = __uint_as_float(v[2 * j + 1]) + s_b2[g * 32 + 2 * j + 1] + ptx::bf16_hi(rw[j]);
            sum += a + b;
            sq += a * a + b * b;
            v[2 * j] = __float_as_uint(a);
            v[2 * j + 1] = __float_as_uint(b);
          }
          ptx::tmem_st_32x32(t_acc + g * 32, v);
        }
      }
      ptx::tmem_st_wait();
      float* st = s_stat + (it & 1) * 512;
      st[(wg * 128 + row) * 2 + 0] = sum;
      st[(wg * 128 + row) * 2 + 1] = sq;
      ptx::named_bar_sync(3, 256);
      sum += st[((wg ^ 1) * 128 + row) * 2 + 0];
      sq += st[((wg ^ 1) * 128 + row) * 2 + 1];
      const float mean = sum * (1.f / kDm);
      const float var = fmaxf(sq * (1.f / kDm) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + 1e-5f);
      for (int c = wg; c < 4; c += 2) {
        uint32_t packed[32];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int g = c * 2 + h;
          uint32_t v[32];
          ptx::tmem_ld_32x32(t_acc + g * 32, v);
          ptx::tmem_ld_wait();
          const uint64_t nmean2 = ptx::f32x2(__float_as_uint(-mean), __float_as_uint(-mean));
          const uint64_t rstd2 = ptx::f32x2(__float_as_uint(rstd), __float_as_uint(rstd));
          const float4* g4 = reinterpret_cast<const float4*>(s_gamma + g * 32);
          const float4* b4 = reinterpret_cast<const float4*>(s_beta + g * 32);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 gq = g4[j4], bq = b4[j4];
            uint64_t x0 = ptx::mul_f32x2(ptx::add_f32x2(ptx::f32x2(v[4 * j4], v[4 * j4 + 1]), nmean2), rstd2);
            uint64_t x1 = ptx::mul_f32x2(ptx::add_f32x2(ptx::f32x2(v[4 * j4 + 2], v[4 * j4 + 3]), nmean2), rstd2);
            x0 = ptx::fma_f32x2(x0, ptx::f32x2(__float_as_uint(gq.x), __float_as_uint(gq.y)), ptx::f32x2(__float_as_uint(bq.x), __float_as_uint(bq.y)));
            x1 = ptx::fma_f32x2(x1, ptx::f32x2(__float_as_uint(gq.z), __float_as_uint(gq.w)), ptx::f32x2(__float_as_uint(bq.z), __float_as_uint(bq.w)));
            packed[h * 16 + 2 * j4] = ptx::cvt_bf16x2(x0);
            packed[h * 16 + 2 * j4 + 1] = ptx::cvt_bf16x2(x1);
          }
        }
        const int n_out = p.has_d2 ? 2 : 1;
        for (int o = 0; o < n_out; ++o) {
          if (o == 1) {
            const float* pp = p.pos + (long long)(row_ok ? ((m + p.pos_row0) % p.pos_rows) : 0) * kDm + c * 64;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float4 q4 = __ldg(reinterpret_cast<const float4*>(pp) + j);
              packed[2 * j] = ptx::pack_bf16(ptx::bf16_lo(packed[2 * j]) + q4.x, ptx::bf16_hi(packed[2 * j]) + q4.y);
              packed[2 * j + 1] = ptx::pack_bf16(ptx::bf16_lo(packed[2 * j + 1]) + q4.z, ptx::bf16_hi(packed[2 * j + 1]) + q4.w);
            }
          }
          // y is staged over H chunk c (G2 of the last hidden chunk has read it: acc2_full), y2 over X chunk c (this warpgroup read
          // its residual columns in the first pass; the barrier below orders the other rows' reads before the box is overwritten)
          uint8_t* box = (o == 0 ? smem_a2 : smem_a1) + c * CHUNK_BYTES;
          ptx::named_bar_sync(bar_id, 128);
          uint8_t* rowp = box + row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(rowp + ((j ^ (row & 7)) << 4)) = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          ptx::fence_proxy_async_smem();
          ptx::named_bar_sync(bar_id, 128);
          if (et == 0) {
            ptx::tma_store_2d(o == 0 ? &p.tmD : &p.tmD2, box, c * 64, m0);
            ptx::tma_store_commit();
          }
        }
      }
      ptx::tc_fence_before_sync();
      warp_arrive(acc2_empty);
      // this warpgroup's stores out of the X / H buffers have been read: the next X tile may land, E1 may write H again
      if (et == 0) {
        ptx::tma_store_wait_read<0>();
        ptx::mbar_arrive(a1_free);
      }
      ptx::named_bar_sync(bar_id, 128);
    }
    if (et == 0) ptx::tma_store_wait_all<0>();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 9) ptx::tmem_dealloc<512>(tmem_base);
}

}  // namespace

int mlp_plan(MlpPlan* plan, const __nv_bfloat16* x, const __nv_bfloat16* w1, const float* b1, const __nv_bfloat16* w2, const float* b2,
             const float* gamma, const float* beta, __nv_bfloat16* d, __nv_bfloat16* d2, const float* pos, int pos_rows, int M) {
  *plan = MlpPlan{};
  OPD_REQUIRE(M > 0 && x && w1 && b1 && w2 && b2 && gamma && beta && d, "mlp: NULL argument");
  if (d2) OPD_REQUIRE(pos && pos_rows > 0, "mlp: D2 needs pos");
  plan->M = M;
  plan->b1 = b1; plan->b2 = b2; plan->gamma = gamma; plan->beta = beta; plan->pos = pos; plan->pos_rows = pos_rows;
  plan->has_d2 = d2 != nullptr;
  if (int rc = make_tmap_2d(&plan->tmX, x, M, kDm, kDm, BLOCK_M)) return rc;
  if (int rc = make_tmap_2d(&plan->tmW1, w1, kHidden, kDm, kDm, 128)) return rc;
  if (int rc = make_tmap_2d(&plan->tmW2, w2, kDm, kHidden, kHidden, 128)) return rc;
  if (int rc = make_tmap_2d(&plan->tmD, d, M, kDm, kDm, BLOCK_M)) return rc;
  if (d2) {
    if (int rc = make_tmap_2d(&plan->tmD2, d2, M, kDm, kDm, BLOCK_M)) return rc;
  } else {
    plan->tmD2 = plan->tmD;
  }
  plan->grid = std::min((M + BLOCK_M - 1) / BLOCK_M, sm_count());
  return OPD_OK;
}

int mlp_launch(const MlpPlan& plan, cudaStream_t stream) {
  MlpParams p;
  p.tmX = plan.tmX; p.tmW1 = plan.tmW1; p.tmW2 = plan.tmW2; p.tmD = plan.tmD; p.tmD2 = plan.tmD2;
  p.M = plan.M;
  p.num_tiles = (plan.M + BLOCK_M - 1) / BLOCK_M;
  p.b1 = plan.b1; p.b2 = plan.b2; p.gamma = plan.gamma; p.beta = plan.beta; p.pos = plan.pos;
  p.pos_rows = plan.pos_rows > 0 ? plan.pos_rows : 1; p.pos_row0 = plan.pos_row0; p.has_d2 = plan.has_d2;
  static PerDeviceOnce configured;
  if (int rc = once_per_device(configured, []() -> int {
        OPD_CUDA_OK(cudaFuncSetAttribute(tc_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        return OPD_OK;
      }))
    return rc;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.gridDim = dim3(plan.grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = g_option_pdl.load() ? 1 : 0;
  OPD_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_mlp_kernel, p));
  count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

}  // namespace opd

// C ABI: the building block on its own (unit-test surface)
extern "C" int opd_mlp_ln_bf16(const void* x_dev, const void* w1_dev, const float* b1_dev, const void* w2_dev, const float* b2_dev,
                               const float* gamma_dev, const float* beta_dev, void* d_dev, void* d2_dev, const float* pos_dev,
                               int32_t pos_rows, int32_t M, void* stream) {
  opd::MlpPlan plan;
  if (int rc = opd::mlp_plan(&plan, static_cast<const __nv_bfloat16*>(x_dev), static_cast<const __nv_bfloat16*>(w1_dev), b1_dev,
                             static_cast<const __nv_bfloat16*>(w2_dev), b2_dev, gamma_dev, beta_dev, static_cast<__nv_bfloat16*>(d_dev),
                             static_cast<__nv_bfloat16*>(d2_dev), pos_dev, pos_rows, M))
    return rc;
  return opd::mlp_launch(plan, static_cast<cudaStream_t>(stream));
}
