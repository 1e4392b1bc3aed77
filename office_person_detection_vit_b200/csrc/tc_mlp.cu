// Fused transformer feed-forward block on the 5th-generation tensor cores:
//     y  = LayerNorm( relu(x W1^T + b1) W2^T + b2 + x ) * gamma + beta          x, y [M, 256], hidden width 2048
//     y2 = bf16(y + pos[row % pos_rows])                                         (query / key input of the next attention)
// i.e. DetrEncoderLayer / DetrDecoderLayer: mlp.fc1 -> activation -> mlp.fc2 -> residual -> final_layer_norm (transformers
// models/detr/modeling_detr.py:560-640, 643-760; ReLU, post-norm), in ONE persistent kernel: the [M, 2048] hidden activations
// (275 MB per encoder layer at batch 64) are neither written nor read back.  As two GEMM launches, fc1 ran at the WRITE roofline of
// HBM (3.9 TB/s) and fc2 re-read what it had written: 79 + 98 us per encoder layer, 19 + 36 us per decoder layer.
//
// Per 128-row tile (persistent CTAs, static schedule), for the sixteen 128-wide chunks c of the hidden layer:
//   G1(c): acc1[c & 1][128 x 128] = X[128 x 256] * W1[c]^T    X resident in shared memory (also the LayerNorm residual)
//   E1(c): H[c & 1] = bf16(relu(acc1 + b1[c])) -> shared memory   two K-chunks [128 x 64], one per epilogue warpgroup
//   G2(c): acc2[128 x 256] += H[c & 1] * W2[:, c]^T
//   E2   : LayerNorm(acc2 + b2 + X) -> y, y + pos -> y2 (both staged over H), TMA stores; X is released after the residual pass, so
//          the next tile's X load and first GEMMs overlap the rest of E2
// acc1 and H are double-buffered and the MMA thread issues G1(c + 2) right after G2(c): the tensor pipe works on G2(c) / G1(c + 2)
// while the epilogue warps convert chunk c + 1 (with 256-wide chunks and single buffers the pipe idled through every E1: 86 k
// cycles per tile against 33 k of MMA work).
// Weight tiles [128 rows x 64 k] stream through one ring in the order the MMA thread consumes them (W1[0]; W2[0], W1[1]; ...).
// Every output element is accumulated in the order of the two-launch path (k ascending, fp32 in TMEM) and both epilogues apply
// the same operations in the same order, so the results are bit-identical to gemm(EPI_BIAS_RELU) + gemm(EPI_BIAS_RES_LN)
// (tests/test_tc_ops_gpu.py).
// Warps: 0-7 epilogue (two warpgroups: warpgroup g owns K-chunk g of H and the column chunks g and g + 2 of the output), 8 TMA
// producer, 9 MMA issuer + TMEM owner.  TMEM: acc1 buffers = columns [0, 128) and [128, 256), acc2 = [256, 512).
#include <algorithm>
#include <cstdio>

#include "opd_common.h"
#include "sm100_ptx.cuh"
#include "tc_gemm.h"

namespace opd {
namespace {

constexpr int BLOCK_M = 128, kDm = 256, kHidden = 2048, kChunk = 128, kNC = kHidden / kChunk;
constexpr int BLOCK_K = 64, UMMA_K = 16;
constexpr int CHUNK_BYTES = BLOCK_M * 64 * 2;        // [128 x 64] bf16, 128-byte swizzle
constexpr int B_STAGE_BYTES = 128 * 64 * 2;          // ring slot: up to 128 weight rows x 64 k
constexpr int kStages = 5;
constexpr int kThreads = 320;
constexpr int kSmemBytes = 8 * CHUNK_BYTES + kStages * B_STAGE_BYTES + (kHidden + 3 * kDm) * 4 + 4096 + 256;

#ifdef OPD_MLP_PROBE
constexpr bool kMlpProbe = true;    // clock64 counters of the MMA thread's waits, printed by CTA 0
#else
constexpr bool kMlpProbe = false;
#endif
__device__ __forceinline__ long long mclk() { return kMlpProbe ? clock64() : 0; }

struct MlpParams {
  CUtensorMap tmX, tmW1, tmW2, tmD, tmD2;   // tmW1: box of 128 rows (kPair: 64 = this CTA's half of a 128-unit hidden chunk)
  int M, num_tiles;
  const float* b1;
  const float* b2;
  const float* gamma;
  const float* beta;
  const float* pos;
  int pos_rows, pos_row0, has_d2;
};

// kPair (cta_group::2, clusters of two CTAs): the two CTAs work on two consecutive row tiles as ONE 256-row MMA.  Each CTA keeps
// its own X / H tiles and loads HALF of every weight tile, so a ring slot feeds twice the MMA cycles and the L2 -> SM weight
// traffic halves.  The leader's MMA thread issues both GEMMs for both CTAs; its commits are multicast to both CTAs' barriers; both
// CTAs' TMA loads signal the leader's full barriers and both CTAs' epilogue warps arrive on the leader's accumulator / H barriers
// (same protocol as tc_gemm.cu / tc_bottleneck.cu).
template <bool kPair>
__global__ void __launch_bounds__(kThreads, 1) tc_mlp_kernel(const __grid_constant__ MlpParams p) {
  constexpr uint32_t kIdesc1 = kPair ? ptx::umma_idesc_bf16(2 * BLOCK_M, kChunk) : ptx::umma_idesc_bf16(BLOCK_M, kChunk);
  constexpr uint32_t kIdesc2 = kPair ? ptx::umma_idesc_bf16(2 * BLOCK_M, 256) : ptx::umma_idesc_bf16(BLOCK_M, 128);
  constexpr int kEpiWarps = kPair ? 16 : 8;      // arrivals on the MMA thread's barriers: one per epilogue warp (of both CTAs)
  constexpr uint32_t kW1Bytes = (kPair ? 64 : 128) * BLOCK_K * 2;   // one k-block of a hidden chunk's W1 rows (this CTA's share)
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_a1 = smem;                              // X tile: 4 K-chunks; E2: the LayerNorm residual
  uint8_t* smem_a2 = smem_a1 + 4 * CHUNK_BYTES;         // H: 2 buffers x 2 K-chunks; E2: y / y2 staging (4 boxes)
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
  uint64_t* a1_free = bars + 11;        // the epilogue warps have read the residual out of X
  uint64_t* acc1_full = bars + 12;      // [2]
  uint64_t* acc1_empty = bars + 14;     // [2] all epilogue warps have read the buffer
  uint64_t* a2_ready = bars + 16;       // [2 buffers][2 K-chunks] written (one warpgroup each)
  uint64_t* a2_free = bars + 20;        // [2] G2 has read the H buffer
  uint64_t* acc2_full = bars + 22;
  uint64_t* acc2_empty = bars + 23;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 24);

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
    ptx::mbar_init(a1_free, 8);     // this CTA's eight epilogue warps, after the residual pass
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&acc1_full[i], 1);
      ptx::mbar_init(&acc1_empty[i], kEpiWarps);
      ptx::mbar_init(&a2_free[i], 1);
    }
    for (int i = 0; i < 4; ++i) ptx::mbar_init(&a2_ready[i], kEpiWarps / 2);
    ptx::mbar_init(acc2_full, 1);
    ptx::mbar_init(acc2_empty, kEpiWarps);
    ptx::fence_barrier_init();
  }
  if (warp == 9) {
    if (kPair) ptx::tmem_alloc_2sm<512>(tmem_ptr);
    else ptx::tmem_alloc<512>(tmem_ptr);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (kPair) ptx::cluster_sync();   // the peer's barriers are initialised before anything signals them
  ptx::grid_dependency_wait();      // programmatic dependent launch: everything above overlaps the previous kernel's tail
  ptx::grid_launch_dependents();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_acc2 = tmem_base + 256;   // acc1 buffers: columns [0, 128) and [128, 256)
  // work units: row tiles, or (kPair) pairs of consecutive row tiles, dealt round-robin to the CTAs / clusters
  const int cta_rank = kPair ? (int)ptx::cluster_ctarank() : 0;
  const int n_units = kPair ? (p.num_tiles + 1) / 2 : p.num_tiles;
  const int my_id = kPair ? (int)blockIdx.x / 2 : (int)blockIdx.x, n_ids = kPair ? (int)gridDim.x / 2 : (int)gridDim.x;
  const int n_my = my_id < n_units ? (n_units - my_id + n_ids - 1) / n_ids : 0;
  // odd tail: rank 1 repeats the last tile (identical stores)
  auto tile_m0 = [&](int it) { return (kPair ? min(2 * (my_id + it * n_ids) + cta_rank, p.num_tiles - 1) : my_id + it * n_ids) * BLOCK_M; };

  if (warp == 8) {
    // ===================================== TMA producer =====================================
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      // kPair: both CTAs' bytes are counted by the LEADER's barrier; each CTA waits for its own copy of the empty barrier
      auto ring_load = [&](const CUtensorMap* tm, uint32_t bytes, int c0, int c1) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        if (!kPair) ptx::mbar_expect_tx(&full_bar[stage], bytes);
        else if (cta_rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * bytes);
        if (kPair) ptx::tma_load_2d_2sm(tm, &full_bar[stage], smem_b + stage * B_STAGE_BYTES, c0, c1);
        else ptx::tma_load_2d(tm, &full_bar[stage], smem_b + stage * B_STAGE_BYTES, c0, c1);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      };
      auto load_w1 = [&](int c) {   // rows = hidden units [128 c, 128 c + 128), k = model columns
        if constexpr (kPair) {      // this CTA's 64 rows: two k-blocks (8 KB each) share a ring slot, so that every slot feeds 512 MMA cycles
          for (int kp = 0; kp < 2; ++kp) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            if (cta_rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 4 * kW1Bytes);
            for (int hk = 0; hk < 2; ++hk)
              ptx::tma_load_2d_2sm(&p.tmW1, &full_bar[stage], smem_b + stage * B_STAGE_BYTES + hk * kW1Bytes, (2 * kp + hk) * BLOCK_K,
                                   c * kChunk + cta_rank * 64);
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        } else {
          for (int kb = 0; kb < 4; ++kb) ring_load(&p.tmW1, kW1Bytes, kb * BLOCK_K, c * kChunk);
        }
      };
      auto load_w2 = [&](int c) {   // rows = the 256 outputs (two halves; kPair: this CTA's), k = hidden units [128 c, 128 c + 128)
        for (int kb = 0; kb < 2; ++kb)
          for (int nh = (kPair ? cta_rank : 0); nh < (kPair ? cta_rank + 1 : 2); ++nh)
            ring_load(&p.tmW2, B_STAGE_BYTES, c * kChunk + kb * BLOCK_K, nh * 128);
      };
      for (int it = 0; it < n_my; ++it) {
        const int m0 = tile_m0(it);
        ptx::mbar_wait(a1_free, (it & 1) ^ 1);
        if (!kPair) ptx::mbar_expect_tx(a1_full, 4 * CHUNK_BYTES);
        else if (cta_rank == 0) ptx::mbar_expect_tx(a1_full, 8 * CHUNK_BYTES);
        for (int kb = 0; kb < 4; ++kb) {
          if (kPair) ptx::tma_load_2d_2sm(&p.tmX, a1_full, smem_a1 + kb * CHUNK_BYTES, kb * BLOCK_K, m0);
          else ptx::tma_load_2d(&p.tmX, a1_full, smem_a1 + kb * CHUNK_BYTES, kb * BLOCK_K, m0);
        }
        load_w1(0);   // the order in which the MMA thread consumes them: G1(0) G1(1) | G2(c) G1(c + 2) ...
        load_w1(1);
        for (int c = 0; c < kNC; ++c) {
          load_w2(c);
          if (c + 2 < kNC) load_w1(c + 2);
        }
      }
    }
  } else if (warp == 9) {
    // ===================================== MMA issuer =====================================
    if ((!kPair || cta_rank == 0) && ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      long long w_ring = 0, w_a2 = 0, w_acc1 = 0, w_a1 = 0, w_acc2 = 0, t_begin = mclk();
      auto commit = [&](uint64_t* bar) {   // kPair: to both CTAs' copies of the barrier
        if (kPair) ptx::umma_commit_2sm(bar, (uint16_t)0x3);
        else ptx::umma_commit(bar);
      };
      // four MMAs (k = 0 .. 3) of one ring slot
      auto slot_mmas = [&](uint32_t acc, uint64_t da, uint32_t idesc, bool first) {
        const long long t0 = mclk();
        ptx::mbar_wait(&full_bar[stage], phase);
        w_ring += mclk() - t0;
        ptx::tc_fence_after_sync();
        const uint64_t db = ptx::umma_desc_kmajor_sw128(ptx::smem_u32(smem_b + stage * B_STAGE_BYTES));
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          if (kPair) ptx::umma_bf16_ss_2sm(acc, da + 2 * k, db + 2 * k, idesc, !(first && k == 0));
          else ptx::umma_bf16_ss(acc, da + 2 * k, db + 2 * k, idesc, !(first && k == 0));
        }
        commit(&empty_bar[stage]);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      };
      uint32_t q1 = 0, q2 = 0;   // hidden chunks issued: first / second GEMM
      auto issue_g1 = [&]() {    // acc1[q1 & 1] = X * W1[chunk]^T
        const uint32_t b = q1 & 1;
        const long long t0 = mclk();
        ptx::mbar_wait(&acc1_empty[b], ((q1 >> 1) & 1) ^ 1);
        w_acc1 += mclk() - t0;
        ptx::tc_fence_after_sync();
        if constexpr (kPair) {   // two k-blocks per ring slot
          for (int kp = 0; kp < 2; ++kp) {
            const long long t4 = mclk();
            ptx::mbar_wait(&full_bar[stage], phase);
            w_ring += mclk() - t4;
            ptx::tc_fence_after_sync();
#pragma unroll
            for (int hk = 0; hk < 2; ++hk) {
              const uint64_t da = ptx::umma_desc_kmajor_sw128(ptx::smem_u32(smem_a1 + (2 * kp + hk) * CHUNK_BYTES));
              const uint64_t db = ptx::umma_desc_kmajor_sw128(ptx::smem_u32(smem_b + stage * B_STAGE_BYTES + hk * kW1Bytes));
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                ptx::umma_bf16_ss_2sm(tmem_base + b * kChunk, da + 2 * k, db + 2 * k, kIdesc1, (kp | hk | k) != 0);
            }
            commit(&empty_bar[stage]);
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        } else {
          for (int kb = 0; kb < 4; ++kb)
            slot_mmas(tmem_base + b * kChunk, ptx::umma_desc_kmajor_sw128(ptx::smem_u32(smem_a1 + kb * CHUNK_BYTES)), kIdesc1, kb == 0);
        }
        commit(&acc1_full[b]);
        ++q1;
      };
      auto issue_g2 = [&](int c, int it) {   // acc2 += H[q2 & 1] * W2[:, chunk]^T
        const uint32_t b = q2 & 1;
        if (c == 0) {
          const long long t2 = mclk();
          ptx::mbar_wait(acc2_empty, (it & 1) ^ 1);
          w_acc2 += mclk() - t2;
          ptx::tc_fence_after_sync();
        }
        for (int kb = 0; kb < 2; ++kb) {
          const long long t3 = mclk();
          ptx::mbar_wait(&a2_ready[b * 2 + kb], (q2 >> 1) & 1);
          w_a2 += mclk() - t3;
          ptx::tc_fence_after_sync();
          const uint64_t da = ptx::umma_desc_kmajor_sw128(ptx::smem_u32(smem_a2 + (b * 2 + kb) * CHUNK_BYTES));
          for (int nh = 0; nh < (kPair ? 1 : 2); ++nh) slot_mmas(tmem_acc2 + nh * 128, da, kIdesc2, c == 0 && kb == 0);
        }
        commit(&a2_free[b]);
        if (c + 1 == kNC) commit(acc2_full);
        ++q2;
      };
      for (int it = 0; it < n_my; ++it) {
        const long long t1 = mclk();
        ptx::mbar_wait(a1_full, it & 1);
        w_a1 += mclk() - t1;
        ptx::tc_fence_after_sync();
        issue_g1();
        issue_g1();
        for (int c = 0; c < kNC; ++c) {
          issue_g2(c, it);
          if (c + 2 < kNC) issue_g1();
        }
      }
      if (kMlpProbe && blockIdx.x == 0)
        printf("mlp CTA 0 MMA thread: %d tile(s), %lld cycles: waiting for X %lld, weight ring %lld, H chunks %lld, acc1 free %lld, acc2 free %lld\n",
               n_my, mclk() - t_begin, w_a1, w_ring, w_a2, w_acc1, w_acc2);
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
    auto warp_arrive = [&](uint64_t* bar) {   // the MMA thread's barriers live in the leader CTA
      __syncwarp();
      if (lane == 0) {
        if (kPair) ptx::mbar_arrive_cluster(bar, 0);
        else ptx::mbar_arrive(bar);
      }
    };
    for (int it = 0; it < n_my; ++it) {
      const int m0 = tile_m0(it);
      const long long m = (long long)m0 + row;
      const bool row_ok = m < p.M;
      // ---------------- E1: sixteen hidden chunks of 128; warpgroup g converts hidden units [64 g, 64 g + 64) = K-chunk g of H ----------------
      for (int c = 0; c < kNC; ++c) {
        const uint32_t qc = (uint32_t)it * kNC + c, b = qc & 1;
        ptx::mbar_wait(&acc1_full[b], (qc >> 1) & 1);
        ptx::mbar_wait(&a2_free[b], ((qc >> 1) & 1) ^ 1);     // G2 of the chunk two back has read this H buffer
        ptx::tc_fence_after_sync();
        uint32_t packed[32];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(tmem_base + lane_addr + b * kChunk + wg * 64 + h * 32, v);
          ptx::tmem_ld_wait();
          const float4* bias4 = reinterpret_cast<const float4*>(s_b1 + c * kChunk + wg * 64 + h * 32);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 bq = bias4[j4];
            const uint64_t s0 = ptx::add_f32x2(ptx::f32x2(v[4 * j4], v[4 * j4 + 1]), ptx::f32x2(__float_as_uint(bq.x), __float_as_uint(bq.y)));
            const uint64_t s1 = ptx::add_f32x2(ptx::f32x2(v[4 * j4 + 2], v[4 * j4 + 3]), ptx::f32x2(__float_as_uint(bq.z), __float_as_uint(bq.w)));
            packed[h * 16 + 2 * j4] = ptx::relu_bf16x2(ptx::cvt_bf16x2(s0));
            packed[h * 16 + 2 * j4 + 1] = ptx::relu_bf16x2(ptx::cvt_bf16x2(s1));
          }
        }
        ptx::tc_fence_before_sync();
        warp_arrive(&acc1_empty[b]);              // this warp's TMEM reads are complete: the accumulator buffer is free first
        uint8_t* rowp = smem_a2 + (b * 2 + wg) * CHUNK_BYTES + row * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(rowp + ((j ^ (row & 7)) << 4)) = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        ptx::fence_proxy_async_smem();
        warp_arrive(&a2_ready[b * 2 + wg]);
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
      // X has been read for the last time (the MMAs of this tile completed before acc2_full): the next tile's X may land and its
      // first GEMMs run while the second pass and the stores below are still going on
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(a1_free);
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
          // both outputs are staged over this warpgroup's two H boxes (G2 of the last hidden chunk has read them: acc2_full):
          // y in box wg, y2 in box wg + 2; the second column chunk waits until the first one's stores have read them
          uint8_t* box = smem_a2 + (wg + 2 * o) * CHUNK_BYTES;
          if (c >= 2 && o == 0 && et == 0) ptx::tma_store_wait_read<0>();
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
      // this warpgroup's stores out of its H boxes have been read: E1 of the next tile may write them again
      if (et == 0) ptx::tma_store_wait_read<0>();
      ptx::named_bar_sync(bar_id, 128);
    }
    if (et == 0) ptx::tma_store_wait_all<0>();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (kPair) ptx::cluster_sync();   // no CTA leaves while its peer may still signal its barriers or read its shared memory
  if (warp == 9) {
    if (kPair) ptx::tmem_dealloc_2sm<512>(tmem_base);
    else ptx::tmem_dealloc<512>(tmem_base);
  }
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
  const int tiles = (M + BLOCK_M - 1) / BLOCK_M;
  plan->grid = std::min(tiles, sm_count());
  plan->pair = g_option_mlp_pair.load() != 0 && tiles >= 2;   // cta_group::2 pairs (default) unless there is a single tile
  if (int rc = make_tmap_2d(&plan->tmX, x, M, kDm, kDm, BLOCK_M)) return rc;
  if (int rc = make_tmap_2d(&plan->tmW1, w1, kHidden, kDm, kDm, plan->pair ? 64 : 128)) return rc;
  if (int rc = make_tmap_2d(&plan->tmW2, w2, kDm, kHidden, kHidden, 128)) return rc;
  if (int rc = make_tmap_2d(&plan->tmD, d, M, kDm, kDm, BLOCK_M)) return rc;
  if (d2) {
    if (int rc = make_tmap_2d(&plan->tmD2, d2, M, kDm, kDm, BLOCK_M)) return rc;
  } else {
    plan->tmD2 = plan->tmD;
  }
  return OPD_OK;
}

int mlp_launch(const MlpPlan& plan, cudaStream_t stream) {
  MlpParams p;
  p.tmX = plan.tmX; p.tmW1 = plan.tmW1; p.tmW2 = plan.tmW2; p.tmD = plan.tmD; p.tmD2 = plan.tmD2;
  p.M = plan.M;
  p.num_tiles = (plan.M + BLOCK_M - 1) / BLOCK_M;
  p.b1 = plan.b1; p.b2 = plan.b2; p.gamma = plan.gamma; p.beta = plan.beta; p.pos = plan.pos;
  p.pos_rows = plan.pos_rows > 0 ? plan.pos_rows : 1; p.pos_row0 = plan.pos_row0; p.has_d2 = plan.has_d2;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[2];
  int n_attr = 0;
  if (plan.pair) {
    attr[n_attr].id = cudaLaunchAttributeClusterDimension;
    attr[n_attr].val.clusterDim.x = 2;
    attr[n_attr].val.clusterDim.y = 1;
    attr[n_attr].val.clusterDim.z = 1;
    ++n_attr;
  }
  if (g_option_pdl.load()) {
    attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
    ++n_attr;
  }
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = n_attr;
  if (plan.pair) {
    static PerDeviceInt cluster_limit;
    int max_clusters = -1;
    if (int rc = cached_per_device(cluster_limit, &max_clusters, [&](int* n) -> int {
          OPD_CUDA_OK(cudaFuncSetAttribute(tc_mlp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
          cfg.gridDim = dim3(sm_count() / 2 * 2);
          OPD_CUDA_OK(cudaOccupancyMaxActiveClusters(n, tc_mlp_kernel<true>, &cfg));
          return OPD_OK;
        }))
      return rc;
    OPD_REQUIRE(max_clusters > 0, "mlp: no 2-CTA cluster of the kernel fits on this device");
    cfg.gridDim = dim3(2 * std::min((p.num_tiles + 1) / 2, max_clusters));
    OPD_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_mlp_kernel<true>, p));
  } else {
    static PerDeviceOnce configured;
    if (int rc = once_per_device(configured, []() -> int {
          OPD_CUDA_OK(cudaFuncSetAttribute(tc_mlp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
          return OPD_OK;
        }))
      return rc;
    cfg.gridDim = dim3(plan.grid);
    OPD_CUDA_OK(cudaLaunchKernelEx(&cfg, tc_mlp_kernel<false>, p));
  }
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
