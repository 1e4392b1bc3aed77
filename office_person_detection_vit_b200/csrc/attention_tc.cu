// K6: fused multi-head attention on the 5th-generation tensor cores (head_dim 32, bf16 in, fp32 softmax):
//     o = softmax(q k^T / sqrt(32)) v          per (batch, head)
// replaces transformers models/detr/modeling_detr.py:386-411 (eager_attention_forward) inside DetrSelfAttention /
// DetrCrossAttention (:414-557), including the key-padding mask of padded batches (masked keys get probability 0).
//
// One CTA = 128 query rows of one (batch, head); key/value tiles of 128 stream through a 2-stage TMA ring.
//   warp 0   TMA producer: Q tile once, then K_j / V_j tiles ([128 x 32] bf16, 64-byte rows, 64-byte swizzle)
//   warp 1   MMA issuer:   S = Q K_j^T   (2 x tcgen05.mma 128x128x16, K-major operands)          -> TMEM columns [0,128)
//                          O += P V_j    (8 x tcgen05.mma 128x32x16, A = P from shared memory, B = V_j MN-major) -> [128,160)
//   warps 2-5 softmax:     thread = query row: TMEM -> registers, online max / exp2 / sum in fp32, P -> bf16 -> 128B-swizzled
//                          shared memory (the A operand of the second MMA), O rescaled in TMEM when the row maximum moves
// S_{j+1} is issued ahead of P V_j so the next tile's softmax overlaps the value MMA; TMEM use is 256 columns, shared
// memory 73 KB, so two CTAs share an SM and cover each other's barrier latencies.
#include <algorithm>

#include "detr_kernels.h"
#include "opd_common.h"
#include "sm100_ptx.cuh"
#include "tc_gemm.h"

namespace opd {
namespace {

constexpr int kQ = 128, kD = 32;
constexpr int Q_BYTES = 128 * 64;             // [128 rows x 32 bf16]
constexpr int P_SUB = 128 * 64;               // P sub-chunk: [128 rows x 32 keys] bf16, 64-byte rows, 64-byte swizzle (like Q)
constexpr int kThreads = 192;
#ifndef OPD_ATTN_POLY_EXP
#define OPD_ATTN_POLY_EXP 0   // measured with 96-key tiles: 219.7 us against 225.8 us per encoder layer - not worth 7.5e-5 of relative error in P
#endif
#ifndef OPD_ATTN_PDOUBLE
#define OPD_ATTN_PDOUBLE 0
#endif
constexpr bool kPDoubleOpt = OPD_ATTN_PDOUBLE != 0;   // kRegS: P double-buffered by tile parity, 3 CTAs per SM (measured: slower)
constexpr bool kPolyExp = OPD_ATTN_POLY_EXP != 0;   // every other pair of exponentials on the FMA pipe (exp2_poly_pair)
// kKV = keys per tile.  The kernel is paced by the per-tile handshake chain (softmax -> barrier -> MMA issue -> tensor pipe ->
// commit -> barrier -> softmax: ~1.7 us per tile with ALL arithmetic removed, benchmarks/attention_microbench.py --probe), not by
// MUFU, TMEM reads or issue slots, so what counts is keys per handshake x CTAs in flight:
//   96 (default): S 96 + O 32 = the 128 TMEM columns of a quarter SM, 4 CTAs per SM, 11 tiles for 1050 keys (1056: no ragged waste);
//                 the V ring is one stage deep so that four CTAs fit in shared memory (51 KB each)
//   64: 4 CTAs per SM, 17 tiles;   128: TMEM 128 + 32 -> 256 columns, 2 CTAs per SM
template <int kKV>
struct AttnCfg {
  static constexpr int kTileBytes = kKV * 64;                  // K / V tile [kKV rows x 32 bf16]
  static constexpr int kPSubs = kKV / 32;
  static constexpr int kVStages = kKV == 96 ? 1 : 2;
  static constexpr int kSmemBytes = Q_BYTES + (2 + kVStages) * kTileBytes + kPSubs * P_SUB + 128;
  static constexpr int kTmemCols = kKV == 128 ? 256 : 128;
  static constexpr int kCtasPerSm = kKV == 128 ? 2 : 4;
};

struct AttnParams {
  CUtensorMap tmQ, tmK, tmV;
  __nv_bfloat16* o;
  long long ldo;
  int Lq, Lk;
  const uint32_t* key_mask;   // [B, key_mask_stride] 32-key words, or nullptr (see AttnPlan)
  int key_mask_stride;
};

// shared-memory descriptors: 64-byte swizzle (rows of 32 bf16), 8-row groups 512 B apart
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
constexpr uint32_t kSw64 = 4, kSw128 = 2;

__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// packed fp32 pairs (Blackwell FFMA2 / FADD2): two lanes per instruction
__device__ __forceinline__ uint64_t pack2f(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2f(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for a pair of x <= 0 on the FMA / integer pipes (Cody-Waite split + cubic, relative error 7.5e-5 = 1/27 of a bf16 half-ulp;
// P is rounded to bf16 for the value MMA and the row sum is taken over the same values), for every other pair of a row
// (-DOPD_ATTN_POLY_EXP=1).  The second pass's samples sit on the MUFU.EX2 instructions (ncu), but the kernel is paced by the
// per-tile handshake chain, not by the MUFU pipe: halving the MUFU count buys 3 %, so this is off by default.
// x is clamped at -125 (2^-125 instead of a flushed 0: invisible in the sums).
__device__ __forceinline__ void exp2_poly_pair(uint64_t x2, float& e0, float& e1) {
  float x0, x1;
  unpack2f(x2, x0, x1);
  const uint64_t xc = pack2f(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
  const uint64_t t = add2(xc, pack2f(12582912.f, 12582912.f));           // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const uint64_t r = add2(t, pack2f(-12582912.f, -12582912.f));          // round(x)
  const uint64_t f = fma2(r, pack2f(-1.f, -1.f), xc);                    // x - round(x) in [-0.5, 0.5]
  uint64_t q = fma2(f, pack2f(0.05517587438225746f, 0.05517587438225746f), pack2f(0.24261151254177094f, 0.24261151254177094f));
  q = fma2(q, f, pack2f(0.6932601928710938f, 0.6932601928710938f));
  q = fma2(q, f, pack2f(0.9999279975891113f, 0.9999279975891113f));
  float t0, t1, q0, q1;
  unpack2f(t, t0, t1);
  unpack2f(q, q0, q1);
  e0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));  // * 2^round(x): straight into the exponent field
  e1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}


// kRegS (64-key tiles): the softmax warps copy the whole S tile into registers (64 per thread) as soon as it is complete and hand
// the TMEM columns straight back, so S_{j+1} is computed WHILE tile j's maximum / exponentials run instead of after them: the
// MMA round trip (softmax done -> barrier -> issue -> commit -> barrier -> tcgen05.ld, ~600 cycles per tile) leaves the per-CTA
// chain.  80 registers per thread (a few spilled words) keep 4 CTAs per SM.  MEASURED (benchmarks/attention_microbench.py, batch
// 64, encoder shape): 267 us, the same as 64-key tiles without it, and 96-key tiles (which do not fit in registers) run in 220 us,
// so it stays an option (opd_set_option("attention_kv", 65)); -DOPD_ATTN_PDOUBLE=1 adds a second P buffer (3 CTAs per SM): slower.
// Also measured: nanosleep back-off in the two single-thread roles' barrier spins (a third of all issued instructions, on two of
// the four schedulers): no change.  ncu of the 64-key kernel (profiles/r02_ncu_attention_raw.csv): MUFU pipe 60 % busy, issue 52 %.
constexpr int kRegSSmemBytes = AttnCfg<64>::kSmemBytes + (kPDoubleOpt ? AttnCfg<64>::kPSubs * P_SUB : 0);   // second P buffer
template <int kKV, bool kRegS = false>
__global__ void __launch_bounds__(kThreads, (kRegS && kPDoubleOpt) ? 3 : AttnCfg<kKV>::kCtasPerSm) attention_tc_kernel(const __grid_constant__ AttnParams p) {
  static_assert(!kRegS || kKV == 64, "S in registers: 64-key tiles only");
  constexpr int kVS = AttnCfg<kKV>::kVStages;
  using C = AttnCfg<kKV>;
  constexpr bool kPD = kRegS && kPDoubleOpt;
  constexpr int TILE_BYTES = C::kTileBytes;
  constexpr int kTmemCols = C::kTmemCols;
  constexpr uint32_t kIdescS = ptx::umma_idesc_bf16(128, kKV);
  constexpr uint32_t kIdescPV = ptx::umma_idesc_bf16(128, 32) | (1u << 16);   // B (= V tile) is MN-major

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_q = smem;
  uint8_t* s_k = s_q + Q_BYTES;             // [2]
  uint8_t* s_v = s_k + 2 * TILE_BYTES;      // [kVS]
  uint8_t* s_p = s_v + kVS * TILE_BYTES;    // sub-chunks of 32 keys
  constexpr int P_BUF = C::kPSubs * P_SUB;   // kPD: two P buffers (tile parity)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_p + (kPD ? 2 : 1) * P_BUF);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;     // [2]
  uint64_t* k_empty = bars + 3;    // [2]
  uint64_t* v_full = bars + 5;     // [2]
  uint64_t* v_empty = bars + 7;    // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* p_ready = bars + 10;
  uint64_t* pv_done = bars + 11;   // [2]  kRegS: value MMA of tile j signals pv_done[j & 1]; else [0] only
  uint64_t* s_free = bars + 13;    // kRegS: every softmax thread holds its S row in registers
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kQ, head = blockIdx.y, b = blockIdx.z;
  const int n_tiles = (p.Lk + kKV - 1) / kKV;
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.tmQ);
    ptx::prefetch_tmap(&p.tmK);
    ptx::prefetch_tmap(&p.tmV);
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&k_full[i], 1);
      ptx::mbar_init(&k_empty[i], 1);
      ptx::mbar_init(&v_full[i], 1);
      ptx::mbar_init(&v_empty[i], 1);
    }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(p_ready, 4);      // one arrival per softmax warp (after __syncwarp): 128 per-thread arrivals serialise on the barrier word
    ptx::mbar_init(&pv_done[0], 1);
    ptx::mbar_init(&pv_done[1], 1);
    ptx::mbar_init(s_free, 4);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<kTmemCols>(tmem_ptr);
  ptx::tc_fence_before_sync();
  __syncthreads();
  // programmatic dependent launch (opd_set_option("pdl", 1)): everything above overlaps the previous kernel's tail
  ptx::grid_dependency_wait();
  ptx::grid_launch_dependents();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + kKV;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(q_full, Q_BYTES);
      ptx::tma_load_2d(&p.tmQ, q_full, s_q, head * kD, b * p.Lq + q0);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        ptx::mbar_wait(&k_empty[st], ph ^ 1);
        ptx::mbar_expect_tx(&k_full[st], TILE_BYTES);
        ptx::tma_load_2d(&p.tmK, &k_full[st], s_k + st * TILE_BYTES, head * kD, b * p.Lk + j * kKV);
        const int vst = kVS == 1 ? 0 : st;
        const uint32_t vph = kVS == 1 ? (uint32_t)(j & 1) : ph;
        ptx::mbar_wait(&v_empty[vst], vph ^ 1);
        ptx::mbar_expect_tx(&v_full[vst], TILE_BYTES);
        ptx::tma_load_2d(&p.tmV, &v_full[vst], s_v + vst * TILE_BYTES, head * kD, b * p.Lk + j * kKV);
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    if (ptx::elect_one()) {
      const uint32_t q_addr = ptx::smem_u32(s_q), p_addr = ptx::smem_u32(s_p);
      auto issue_s = [&](int j) {
        const int st = j & 1;
        ptx::mbar_wait(&k_full[st], (j >> 1) & 1);
        ptx::tc_fence_after_sync();
        const uint32_t k_addr = ptx::smem_u32(s_k + st * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < 2; ++k)
          ptx::umma_bf16_ss(tmem_s, desc(q_addr + k * 32, 512, kSw64), desc(k_addr + k * 32, 512, kSw64), kIdescS, k != 0);
        ptx::umma_commit(&k_empty[st]);
        ptx::umma_commit(s_full);
      };
      ptx::mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        if (j + 1 < n_tiles) {
          // the softmax warps have read S_j out of TMEM for the last time (kRegS: into registers, right away; else the last
          // tcgen05.ld of the second pass): the next scores go out while the rest of tile j's exponentials, the P stores and the
          // O rescale still run - that much of the MMA round trip leaves the per-tile chain
          ptx::mbar_wait(s_free, j & 1);
          ptx::tc_fence_after_sync();
          issue_s(j + 1);
        }
        ptx::mbar_wait(p_ready, j & 1);          // softmax_j: P written, O rescaled
        ptx::tc_fence_after_sync();
        const int vst = kVS == 1 ? 0 : st;
        ptx::mbar_wait(&v_full[vst], kVS == 1 ? (j & 1) : ((j >> 1) & 1));
        ptx::tc_fence_after_sync();
        const uint32_t v_addr = ptx::smem_u32(s_v + vst * TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < kKV / 16; ++kk)   // 16 keys: half of a P sub-chunk's 64-byte rows, 16 rows of the V tile
          ptx::umma_bf16_ss(tmem_o, desc(p_addr + (kPD ? st * P_BUF : 0) + (kk >> 1) * P_SUB + (kk & 1) * 32, 512, kSw64),
                            desc(v_addr + kk * 1024, 512, kSw64), kIdescPV, (j | kk) != 0);
        ptx::umma_commit(&v_empty[vst]);
        ptx::umma_commit(&pv_done[kPD ? st : 0]);
      }
    }
  } else {
    // ===================================== softmax warps (thread = query row) =====================================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float sl2 = 0.17677669529663687f * 1.4426950408889634f;   // 1/sqrt(32) * log2(e)
    float m = -INFINITY, l = 0.f;
    uint8_t* prow = s_p + row * 64;
    auto warp_arrive = [&](uint64_t* bar) {   // every lane's fences precede the one arrival
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar);
    };
    // A warp whose 32 query rows all lie past Lq (the ragged last tile: 1050 = 8 x 128 + 26 leaves three of its four warps
    // without a row, 8 % of all softmax work) only keeps the barrier protocol going.  Its rows of P stay whatever shared memory
    // held: every row of P V depends on its own row of P alone, and these rows are never stored.
    const bool idle_warp = q0 + quarter * 32 >= p.Lq;
    for (int j = 0; j < n_tiles; ++j) {
      ptx::mbar_wait(s_full, j & 1);
      ptx::tc_fence_after_sync();
      if (idle_warp) {
        ptx::tc_fence_before_sync();
        warp_arrive(s_free);
        if (!kPD && j > 0) ptx::mbar_wait(&pv_done[0], (j - 1) & 1);
        ptx::tc_fence_before_sync();
        warp_arrive(p_ready);
        // kRegS: S_{j+1} may already be complete; this warp must not arrive for tile j + 1 before p_ready's phase j has closed
        // (the value MMA of tile j is issued after it)
        if (kPD) ptx::mbar_wait(&pv_done[j & 1], (j >> 1) & 1);
        else if (kRegS) ptx::mbar_wait(&pv_done[0], j & 1);
        continue;
      }
      const int valid = p.Lk - j * kKV;           // keys of this tile that exist (>= kKV: all)
      // keys of each 32-key chunk that take part: those that exist (the last tile of a row of tiles is ragged) and, in a padded
      // batch, are not padding; all-ones for almost every chunk -> the fast loops
      uint32_t kbits[kKV / 32];
#pragma unroll
      for (int c = 0; c < kKV / 32; ++c) {
        const int left = valid - c * 32;
        kbits[c] = left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : (1u << left) - 1u);
        if (p.key_mask) kbits[c] &= p.key_mask[(long long)b * p.key_mask_stride + (j * kKV) / 32 + c];
      }
      uint32_t sv[kRegS ? kKV / 32 : 1][32];
      if constexpr (kRegS) {
#pragma unroll
        for (int c = 0; c < kKV / 32; ++c) ptx::tmem_ld_32x32(tmem_s + lane_addr + c * 32, sv[c]);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before_sync();
        warp_arrive(s_free);
      }
      // pass 1: row maximum (3-input max)
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < kKV / 32; ++c) {
        uint32_t vv[32];
        uint32_t(&v)[32] = kRegS ? sv[kRegS ? c : 0] : vv;
        if constexpr (!kRegS) {
          ptx::tmem_ld_32x32(tmem_s + lane_addr + c * 32, v);
          ptx::tmem_ld_wait();
        }
        if (kbits[c] == 0xFFFFFFFFu) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) mx = max3(mx, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if ((kbits[c] >> i) & 1u) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
      }
      const float m_new = fmaxf(m, mx * sl2);
      const float alpha = ex2(m - m_new);         // first tile: ex2(-inf) = 0
      // P and O belong to the previous tile's value MMA until it has completed.  kRegS: P is double-buffered by tile parity, so
      // pass 2 only waits for the value MMA of tile j - 2 (long done); the previous tile's is awaited further down, and only by a
      // warp that has to rescale O.  Each of the two barriers is waited for at every one of its phases, in order (at tile j for
      // tile j - 2; the rescale wait for tile j - 1 merely comes one tile early), which is what a parity wait needs.
      bool prev_pv_seen = false;
      if constexpr (kPD) {
        if (j >= 2) ptx::mbar_wait(&pv_done[j & 1], ((j - 2) >> 1) & 1);
      } else if (j > 0) {
        ptx::mbar_wait(&pv_done[0], (j - 1) & 1);
        ptx::tc_fence_after_sync();
        prev_pv_seen = true;
      }
      (void)prev_pv_seen;
      // pass 2: p = 2^(s * sl2 - m_new) -> bf16 -> shared memory (K-major A operand, 128-byte swizzle)
      float lsum = 0.f;
      const uint64_t sl2_2 = pack2f(sl2, sl2), neg_m2 = pack2f(-m_new, -m_new);
      uint64_t sum2 = pack2f(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < kKV / 32; ++c) {
        uint32_t vv[32];
        uint32_t(&v)[32] = kRegS ? sv[kRegS ? c : 0] : vv;
        if constexpr (!kRegS) {
          ptx::tmem_ld_32x32(tmem_s + lane_addr + c * 32, v);
          ptx::tmem_ld_wait();
          if (c == kKV / 32 - 1) {   // S_j is not read again
            ptx::tc_fence_before_sync();
            warp_arrive(s_free);
          }
        }
        uint32_t packed[16];
        if (kbits[c] == 0xFFFFFFFFu) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint64_t x2 = fma2(pack2f(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sl2_2, neg_m2);
            float a, bb;
            if (kPolyExp && (i & 1)) {
              exp2_poly_pair(x2, a, bb);
            } else {
              float x0, x1;
              unpack2f(x2, x0, x1);
              a = ex2(x0);
              bb = ex2(x1);
            }
            sum2 = add2(sum2, pack2f(a, bb));
            packed[i] = ptx::pack_bf16(a, bb);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float a = ex2(fmaf(__uint_as_float(v[2 * i]), sl2, -m_new));
            float bb = ex2(fmaf(__uint_as_float(v[2 * i + 1]), sl2, -m_new));
            if (!((kbits[c] >> (2 * i)) & 1u)) a = 0.f;
            if (!((kbits[c] >> (2 * i + 1)) & 1u)) bb = 0.f;
            lsum += a + bb;
            packed[i] = ptx::pack_bf16(a, bb);
          }
        }
        // 32 keys = one sub-chunk row of 64 bytes; 64-byte swizzle: 16-byte piece i sits at i ^ ((row >> 1) & 3)
        uint8_t* chunk = prow + (kPD ? (j & 1) * P_BUF : 0) + c * P_SUB;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<uint4*>(chunk + ((i ^ ((row >> 1) & 3)) << 4)) =
              make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
      }
      {
        float s0, s1;
        unpack2f(sum2, s0, s1);
        lsum += s0 + s1;
      }
      l = l * alpha + lsum;
      m = m_new;
      // rescale the running output when any row of this warp moved its maximum
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
        if constexpr (kPD) {
          ptx::mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
          ptx::tc_fence_after_sync();
        }
        uint32_t o[32];
        ptx::tmem_ld_32x32(tmem_o + lane_addr, o);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        ptx::tmem_st_32x32(tmem_o + lane_addr, o);
        ptx::tmem_st_wait();
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before_sync();
      warp_arrive(p_ready);
    }
    // epilogue: O / l -> bf16 -> global (the warp of rows 0-31 always holds valid rows, so the last value MMA is awaited before TMEM is freed)
    uint32_t o[32];
    if (!idle_warp) {
      if constexpr (kPD) ptx::mbar_wait(&pv_done[(n_tiles - 1) & 1], ((n_tiles - 1) >> 1) & 1);
      else ptx::mbar_wait(&pv_done[0], (n_tiles - 1) & 1);
      ptx::tc_fence_after_sync();
      ptx::tmem_ld_32x32(tmem_o + lane_addr, o);
      ptx::tmem_ld_wait();
    }
    if (!idle_warp && q0 + row < p.Lq) {
      const float inv = 1.f / l;
      uint32_t w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = ptx::pack_bf16(__uint_as_float(o[2 * i]) * inv, __uint_as_float(o[2 * i + 1]) * inv);
      uint4* dst = reinterpret_cast<uint4*>(p.o + ((long long)b * p.Lq + q0 + row) * p.ldo + head * kD);
#pragma unroll
      for (int i = 0; i < 4; ++i) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<kTmemCols>(tmem_base);
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// [rows, cols] bf16 with row pitch ld; box = [32 columns, 128 rows], 64-byte swizzle
int make_tmap_head(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  OPD_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  OPD_REQUIRE(fn && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled unavailable");
  OPD_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld * 2) % 16 == 0, "attention: operand not 16-byte aligned");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(fn)(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box,
                                                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(OPD_ERR_CUDA, "cuTensorMapEncodeTiled (attention) failed (%d)", (int)r);
  return OPD_OK;
}

}  // namespace

int attn_plan(AttnPlan* plan, const __nv_bfloat16* q, int64_t ldq, const __nv_bfloat16* k, int64_t ldk, const __nv_bfloat16* v,
              int64_t ldv, __nv_bfloat16* o, int64_t ldo, int B, int heads, int Lq, int Lk) {
  *plan = AttnPlan{};
  OPD_REQUIRE(B > 0 && heads > 0 && Lq > 0 && Lk > 0, "attention: bad shape");
  OPD_REQUIRE(ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0, "attention: output not 16-byte aligned");
  const uint64_t cols = (uint64_t)heads * kD;
  const int kv_opt = g_option_attention_kv.load();   // 96 (default) / 64 / 128 keys per tile; 65 = 64 keys with S held in registers
  plan->kv_tile = kv_opt == 128 ? 128 : (kv_opt == 65 ? 65 : (kv_opt == 64 ? 64 : 96));
  if (int rc = make_tmap_head(&plan->tmQ, q, (uint64_t)B * Lq, cols, ldq, 128)) return rc;
  if (int rc = make_tmap_head(&plan->tmK, k, (uint64_t)B * Lk, cols, ldk, plan->kv_tile & ~1)) return rc;
  if (int rc = make_tmap_head(&plan->tmV, v, (uint64_t)B * Lk, cols, ldv, plan->kv_tile & ~1)) return rc;
  plan->o = o; plan->ldo = ldo; plan->B = B; plan->heads = heads; plan->Lq = Lq; plan->Lk = Lk;
  return OPD_OK;
}

int attn_launch(const AttnPlan& plan, cudaStream_t stream) {
  AttnParams p;
  p.tmQ = plan.tmQ; p.tmK = plan.tmK; p.tmV = plan.tmV;
  p.o = plan.o; p.ldo = plan.ldo; p.Lq = plan.Lq; p.Lk = plan.Lk;
  p.key_mask = plan.key_mask; p.key_mask_stride = plan.key_mask_stride;
  static PerDeviceOnce configured;
  if (int rc = once_per_device(configured, []() -> int {
        OPD_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<128>::kSmemBytes));
        OPD_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<64>::kSmemBytes));
        OPD_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<96>::kSmemBytes));
        OPD_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRegSSmemBytes));
        return OPD_OK;
      }))
    return rc;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.gridDim = dim3((plan.Lq + kQ - 1) / kQ, plan.heads, plan.B);
  cfg.blockDim = dim3(kThreads);
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = g_option_pdl.load() ? 1 : 0;
  if (plan.kv_tile == 128) {
    cfg.dynamicSmemBytes = AttnCfg<128>::kSmemBytes;
    OPD_CUDA_OK(cudaLaunchKernelEx(&cfg, attention_tc_kernel<128>, p));
  } else if (plan.kv_tile == 96) {
    cfg.dynamicSmemBytes = AttnCfg<96>::kSmemBytes;
    OPD_CUDA_OK(cudaLaunchKernelEx(&cfg, attention_tc_kernel<96>, p));
  } else if (plan.kv_tile == 65) {
    cfg.dynamicSmemBytes = kRegSSmemBytes;
    OPD_CUDA_OK(cudaLaunchKernelEx(&cfg, attention_tc_kernel<64, true>, p));
  } else {
    cfg.dynamicSmemBytes = AttnCfg<64>::kSmemBytes;
    OPD_CUDA_OK(cudaLaunchKernelEx(&cfg, attention_tc_kernel<64>, p));
  }
  count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

}  // namespace opd
