// Fused bottleneck tail on the 5th-generation tensor cores:
//     y = relu( conv1x1( relu(conv3x3(x, stride) + b2) ) + b3 + residual )
// i.e. layer.1 + layer.2 + shortcut add of a ResNet bottleneck (transformers models/resnet/modeling_resnet.py:134-200,
// frozen BN folded) in ONE persistent kernel.  The 3x3 convolution is L2-bound (im2col re-reads its input 9 times) and
// the 1x1 expansion is HBM-bound (residual in, 4x wider tensor out): run back to back they leave HBM, then L2, idle.
// Fused, the mid activation never leaves the SM (TMEM -> registers -> shared memory as the A operand of the second
// GEMM) and the two phases of different tiles overlap inside every SM.
//
// Per 128-pixel tile t (persistent CTAs, static tile schedule):
//   G1(t): acc1[128 x MID]  = im2col(x)[128 x 9*MID] * W2^T        tcgen05.mma, TMA im2col A + TMA B1 through the ring
//   E1(t): A2 = bf16(relu(acc1 + b2))                               TMEM -> regs -> 128B-swizzled smem (K-major operand)
//   G2(t): acc2[128 x 128] = A2[128 x MID] * W3[n2]^T  per n2 tile  B2 tiles through the same ring
//   E2(t): y = bf16(relu(acc2 + b3 + residual))                     residual by TMA ring, output by TMA store
// Warp roles: 0-7 epilogue (two warpgroups), 8 TMA producer, 9 MMA issuer + TMEM owner, 10 residual producer, 11 W3 tile
// producer; the single-thread roles are elected with elect.sync (the epilogue warpgroups
// split the columns: the epilogue, not the tensor pipe, paces these HBM-bound layers).
// MMA issue order G1(t0) G1(t1) G2(t0) G1(t2) G2(t1) ... so that E1 / E2 of one tile overlap the MMAs of the next.
//
// kPair (cta_group::2, clusters of two CTAs; MID = 128 layers with an even number of 128-pixel tiles): the two CTAs work on
// two consecutive pixel tiles as ONE 256-row MMA.  Each CTA loads its own im2col A slot and HALF of every W2 / W3 tile
// (64 of the 128 output channels), so the L2 -> SM weight traffic that bounds this kernel (profiles/README.md: 10.8 TB/s =
// the LTS cap) halves; the leader's MMA thread issues both GEMMs for both CTAs, its commits are multicast to both CTAs'
// barriers, and both CTAs' epilogue warps arrive (one arrival per warp) on the leader's accumulator / A2 barriers.
// Each CTA's epilogue, residual ring and output stores stay local: its 128 accumulator rows live in its own TMEM.
#include <algorithm>

#include "opd_common.h"
#include "sm100_ptx.cuh"
#include "tc_gemm.h"

namespace opd {
namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int BLOCK_N2 = 128;
constexpr int UMMA_K = 16;
constexpr int CHUNK_BYTES = BLOCK_M * 64 * 2;        // one [128 x 64] bf16 box = 16 KB
constexpr int kThreads = 384;   // warps: 0-7 epilogue (two warpgroups), 8 TMA, 9 MMA, 10 residual TMA, 11 W3 TMA
constexpr int kSmemBudget = 232448;

template <int MID, bool kPair = false>
struct Cfg {
  static constexpr int kA2Chunks = MID / 64;
  static constexpr int kResStages = 4;                               // two residual chunk slots per epilogue warpgroup
  static constexpr int kBSlot = CHUNK_BYTES / (kPair ? 2 : 1);       // B1: MID x 64, B2: 128 x 64; kPair: this CTA's half of the rows
  static constexpr int kStageBytes = CHUNK_BYTES + kBSlot;          // A slot 16 KB + B slot
  static constexpr int kB2Stages = kPair ? 4 : 2;                    // ring of W3 tiles (second GEMM), fed by its own producer
  static constexpr int kFixed = (kA2Chunks + kResStages) * CHUNK_BYTES + kB2Stages * kBSlot + 3072;   // residual slots double as output staging
  static constexpr int kStages = (kSmemBudget - kFixed) / kStageBytes > 6 ? 6 : (kSmemBudget - kFixed) / kStageBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + kFixed;
  static constexpr int kTmemCols = 512;                             // acc1 2 x MID + acc2 2 x 128
  static_assert(MID == 64 || MID == 128, "fused bottleneck tail: MID must be 64 or 128");
  static_assert(kStages >= 3, "ring too shallow");
};

struct BneckParams {
  CUtensorMap tmA, tmB1, tmB2, tmR, tmD;
  int M, width;
  int num_m_blocks, num_n2;
  int KW, stride, pad_h, pad_w, P, Q;
  const float* bias2;
  const float* bias3;
  int early_release;           // residual slots are released early in the next epilogue step (default) instead of at its end
  unsigned long long* trace;   // debug timeline (CTA 0): [0] = count, then (event id << 48 | globaltimer ns) records
};

#ifdef OPD_BNECK_PROBE
constexpr bool kBneckProbe = true;
#else
constexpr bool kBneckProbe = false;   // -DOPD_BNECK_PROBE: clock64 counters of the epilogue's waits, printed by CTA 0
#endif
__device__ __forceinline__ long long pclk() { return kBneckProbe ? clock64() : 0; }

__device__ __forceinline__ void trace_ev(unsigned long long* tr, int id) {
  if (kBneckProbe && tr && blockIdx.x == 0) {   // timeline trace: measurement builds only (-DOPD_BNECK_PROBE)
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    const unsigned long long i = atomicAdd(tr, 1ull);
    if (i < 4000) tr[1 + i] = ((unsigned long long)id << 48) | (t & 0xFFFFFFFFFFFFull);
  }
}

template <int MID, bool kPair>
__global__ void __launch_bounds__(kThreads, 1) tc_bneck_kernel(const __grid_constant__ BneckParams p) {
  using C = Cfg<MID, kPair>;
  constexpr int kMmaM = kPair ? 2 * BLOCK_M : BLOCK_M;
  constexpr uint32_t kB1Bytes = MID * BLOCK_K * 2 / (kPair ? 2 : 1);   // kPair: this CTA's half of the weight tile
  constexpr uint32_t kB2Bytes = CHUNK_BYTES / (kPair ? 2 : 1);
  constexpr int kEpiArrivals = kPair ? 16 : 256;                       // kPair: one arrival per epilogue warp of both CTAs
  constexpr int kStages = C::kStages;
  constexpr int kResStages = C::kResStages;
  constexpr int kK1Blocks = 9 * MID / BLOCK_K;      // 3x3 taps x channel blocks
  constexpr int kCBlocks = MID / BLOCK_K;
  constexpr uint32_t kIdesc1 = ptx::umma_idesc_bf16(kMmaM, MID);
  constexpr uint32_t kIdesc2 = ptx::umma_idesc_bf16(kMmaM, BLOCK_N2);

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_a = smem;                                    // ring: A slots
  uint8_t* smem_b = smem_a + kStages * CHUNK_BYTES;          // ring: B slots
  uint8_t* smem_b2 = smem_b + kStages * C::kBSlot;           // ring: W3 tiles of the second GEMM
  uint8_t* smem_a2 = smem_b2 + C::kB2Stages * C::kBSlot;     // A operand of the second GEMM
  uint8_t* smem_res = smem_a2 + C::kA2Chunks * CHUNK_BYTES;  // residual slots; the output chunk is written in place and stored from there
  float* s_bias2 = reinterpret_cast<float*>(smem_res + kResStages * CHUNK_BYTES);   // [MID]
  float* s_bias3 = s_bias2 + 128;                                                   // [width <= 512]: whole layer, loaded once
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias3 + 512);
  uint64_t* full_bar = bars;                  // [kStages]
  uint64_t* empty_bar = bars + 8;             // [kStages]
  uint64_t* acc1_full = bars + 16;            // [2]
  uint64_t* acc1_empty = bars + 18;           // [2]
  uint64_t* acc2_full = bars + 20;            // [2]
  uint64_t* acc2_empty = bars + 22;           // [2]
  uint64_t* a2_ready = bars + 24;             // [1]
  uint64_t* a2_free = bars + 25;              // [1]
  uint64_t* res_full = bars + 26;             // [4]
  uint64_t* res_empty = bars + 30;            // [4]
  uint64_t* b2_full = bars + 34;              // [kB2Stages <= 4]
  uint64_t* b2_empty = bars + 38;             // [kB2Stages <= 4]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 42);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();

  if (warp == 8 && lane == 0) {
    ptx::prefetch_tmap(&p.tmA);
    ptx::prefetch_tmap(&p.tmB1);
    ptx::prefetch_tmap(&p.tmB2);
    ptx::prefetch_tmap(&p.tmD);
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&acc1_full[i], 1);
      ptx::mbar_init(&acc1_empty[i], kEpiArrivals);
      ptx::mbar_init(&acc2_full[i], 1);
      ptx::mbar_init(&acc2_empty[i], kEpiArrivals);
    }
    ptx::mbar_init(a2_ready, kEpiArrivals);
    ptx::mbar_init(a2_free, 1);
    for (int i = 0; i < C::kB2Stages; ++i) {
      ptx::mbar_init(&b2_full[i], 1);
      ptx::mbar_init(&b2_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      ptx::mbar_init(&res_full[i], 1);
      ptx::mbar_init(&res_empty[i], 1);   // released by the warpgroup's store thread once the TMA store has read the slot
    }
    ptx::fence_barrier_init();
  }
  if (warp == 9) {
    if (kPair) ptx::tmem_alloc_2sm<C::kTmemCols>(tmem_ptr);
    else ptx::tmem_alloc<C::kTmemCols>(tmem_ptr);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (kPair) ptx::cluster_sync();   // the peer's barriers are initialised before anything arrives on them
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const int cta_rank = kPair ? (int)ptx::cluster_ctarank() : 0;
  const uint32_t tmem_acc1 = tmem_base;                 // 2 x MID columns
  const uint32_t tmem_acc2 = tmem_base + 2 * MID;       // 2 x 128 columns

  // kPair: blockIdx.x = 2 * cluster + rank, an even grid and an even number of m-blocks: the pair walks tiles (2j, 2j + 1) together
  const int first = blockIdx.x, step = gridDim.x, n_mblk = p.num_m_blocks;

  if (warp == 8) {
    // ===================================== TMA producer =====================================
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      };
      auto load_g1 = [&](int m_blk) {
        const int m0 = m_blk * BLOCK_M;
        const int pq = p.P * p.Q;
        const int img = m0 / pq;
        const int rem = m0 - img * pq;
        const int op = rem / p.Q, oq = rem - op * p.Q;
        const int base_w = oq * p.stride - p.pad_w, base_h = op * p.stride - p.pad_h;
        for (int kb = 0; kb < kK1Blocks; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          const int tap = kb / kCBlocks, cb = kb - tap * kCBlocks;
          const int r = tap / p.KW, s = tap - r * p.KW;
          if (kPair) {   // both CTAs' bytes are counted by the LEADER's barrier; my half of the weight tile goes into MY slot
            if (cta_rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * (CHUNK_BYTES + kB1Bytes));
            ptx::tma_load_im2col_4d_2sm(&p.tmA, &full_bar[stage], smem_a + stage * CHUNK_BYTES, cb * BLOCK_K, base_w, base_h, img,
                                        (uint16_t)s, (uint16_t)r);
            ptx::tma_load_2d_2sm(&p.tmB1, &full_bar[stage], smem_b + stage * C::kBSlot, kb * BLOCK_K, cta_rank * (MID / 2));
          } else {
            ptx::mbar_expect_tx(&full_bar[stage], CHUNK_BYTES + kB1Bytes);
            ptx::tma_load_im2col_4d(&p.tmA, &full_bar[stage], smem_a + stage * CHUNK_BYTES, cb * BLOCK_K, base_w, base_h, img,
                                    (uint16_t)s, (uint16_t)r);
            ptx::tma_load_2d(&p.tmB1, &full_bar[stage], smem_b + stage * C::kBSlot, kb * BLOCK_K, 0);
          }
          advance();
        }
      };
      for (int t = first; t < n_mblk; t += step) load_g1(t);
    }
  } else if (warp == 11) {
    // ===================================== W3 tile producer (second GEMM) =====================================
    if (ptx::elect_one()) {
      int st = 0;
      uint32_t ph = 0;
      for (int t = first; t < n_mblk; t += step)
        for (int n2 = 0; n2 < p.num_n2; ++n2)
          for (int kb = 0; kb < kCBlocks; ++kb) {
            ptx::mbar_wait(&b2_empty[st], ph ^ 1);
            if (kPair) {
              if (cta_rank == 0) ptx::mbar_expect_tx(&b2_full[st], 2 * kB2Bytes);
              ptx::tma_load_2d_2sm(&p.tmB2, &b2_full[st], smem_b2 + st * C::kBSlot, kb * BLOCK_K, n2 * BLOCK_N2 + cta_rank * (BLOCK_N2 / 2));
            } else {
              ptx::mbar_expect_tx(&b2_full[st], kB2Bytes);
              ptx::tma_load_2d(&p.tmB2, &b2_full[st], smem_b2 + st * C::kBSlot, kb * BLOCK_K, n2 * BLOCK_N2);
            }
            if (++st == C::kB2Stages) {
              st = 0;
              ph ^= 1;
            }
          }
    }
  } else if (warp == 9) {
    // ===================================== MMA issuer =====================================
    // One thread interleaves two streams of work so that neither blocks the other:
    //   G1(i1): the 3x3 convolution of this CTA's tile number i1 (L2 -> SM bound: 18 x 32 KB through the main ring),
    //   G2(i2): the 1x1 expansion of tile i2 <= i1 (paced by the epilogue: it needs A2 from E1 and free acc2 buffers).
    // G2 steps have priority (they unblock the epilogue); whenever G2 cannot advance, G1 of the next tile keeps the ring
    // draining, so the fabric-bound phase of tile i+1 overlaps the HBM / epilogue-bound phase of tile i.
    // kPair: the leader's thread issues for both CTAs.
    auto mma = [](uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
      if (kPair) ptx::umma_bf16_ss_2sm(d, da, db, idesc, acc);
      else ptx::umma_bf16_ss(d, da, db, idesc, acc);
    };
    auto commit = [](uint64_t* bar) {
      if (kPair) ptx::umma_commit_2sm(bar, (uint16_t)0x3);   // the same barrier in both CTAs
      else ptx::umma_commit(bar);
    };
    if ((!kPair || cta_rank == 0) && ptx::elect_one()) {
      const int n_my = first < n_mblk ? (n_mblk - first + step - 1) / step : 0;
      int i1 = 0, kb1 = 0, stage = 0, a1 = 0;          // G1 cursor, main ring, acc1 buffer
      uint32_t phase = 0, a1_phase = 0;
      bool acc1_ok = false;
      int i2 = 0, n2 = 0, kb2 = 0, st2 = 0, a2 = 0;      // G2 cursor, W3 ring, acc2 buffer
      uint32_t ph2 = 0, a2_phase = 0, ready_phase = 0;
      bool ready_seen = false, acc2_ok = false;
      uint32_t idle = 0;
      while (i2 < n_my) {
        bool progressed = false;
        // ---- one k-block of G2(i2) ----
        if (i2 < i1) {   // G1(i2) fully issued (acc1_full committed), E1 can have run
          if (!ready_seen && ptx::mbar_try_wait(a2_ready, ready_phase)) {
            ready_seen = true;
            trace_ev(p.trace, 13);
          }
          if (ready_seen) {
            if (!acc2_ok && ptx::mbar_try_wait(&acc2_empty[a2], a2_phase ^ 1)) acc2_ok = true;
            if (acc2_ok && ptx::mbar_try_wait(&b2_full[st2], ph2)) {
              ptx::tc_fence_after_sync();
              const uint32_t d = tmem_acc2 + a2 * BLOCK_N2;
              const uint64_t da = ptx::umma_desc_kmajor_sw128(ptx::smem_u32(smem_a2 + kb2 * CHUNK_BYTES));
              const uint64_t db = ptx::umma_desc_kmajor_sw128(ptx::smem_u32(smem_b2 + st2 * C::kBSlot));
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) mma(d, da + 2 * k, db + 2 * k, kIdesc2, (kb2 | k) != 0);
              commit(&b2_empty[st2]);
              if (++st2 == C::kB2Stages) {
                st2 = 0;
                ph2 ^= 1;
              }
              if (++kb2 == kCBlocks) {
                kb2 = 0;
                commit(&acc2_full[a2]);
                acc2_ok = false;
                if (++a2 == 2) {
                  a2 = 0;
                  a2_phase ^= 1;
                }
                if (++n2 == p.num_n2) {
                  n2 = 0;
                  commit(a2_free);   // every MMA that reads A2 has completed
                  trace_ev(p.trace, 12);
                  ready_seen = false;
                  ready_phase ^= 1;
                  ++i2;
                }
              }
              progressed = true;
            }
          }
        }
        // ---- one k-block of G1(i1) ----
        if (i1 < n_my) {
          if (!acc1_ok && ptx::mbar_try_wait(&acc1_empty[a1], a1_phase ^ 1)) acc1_ok = true;
          if (acc1_ok && ptx::mbar_try_wait(&full_bar[stage], phase)) {
            ptx::tc_fence_after_sync();
            const uint32_t d = tmem_acc1 + a1 * MID;
            const uint64_t da = ptx::umma_desc_kmajor_sw128(ptx::smem_u32(smem_a + stage * CHUNK_BYTES));
            const uint64_t db = ptx::umma_desc_kmajor_sw128(ptx::smem_u32(smem_b + stage * C::kBSlot));
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) mma(d, da + 2 * k, db + 2 * k, kIdesc1, (kb1 | k) != 0);
            commit(&empty_bar[stage]);
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
            if (++kb1 == kK1Blocks) {
              kb1 = 0;
              commit(&acc1_full[a1]);
              trace_ev(p.trace, 11);
              acc1_ok = false;
              if (++a1 == 2) {
                a1 = 0;
                a1_phase ^= 1;
              }
              ++i1;
            }
            progressed = true;
          }
        }
        if (progressed) idle = 0;
        else if (++idle > (1u << 26)) __trap();   // protocol bug: fail the launch instead of hanging the GPU
      }
    }
  } else if (warp == 10) {
    // ===================================== residual TMA producer =====================================
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&p.tmR);
      uint32_t k = 0;   // chunk counter per warpgroup: slot = wg * 2 + (k & 1), parity = (k >> 1) & 1
      for (int t = first; t < n_mblk; t += step) {
        for (int n2 = 0; n2 < p.num_n2; ++n2, ++k)
          for (int c = 0; c < 2; ++c) {   // chunk c of the n2 tile belongs to epilogue warpgroup c
            const int slot = c * 2 + (k & 1);
            ptx::mbar_wait(&res_empty[slot], ((k >> 1) & 1) ^ 1);
            ptx::mbar_expect_tx(&res_full[slot], CHUNK_BYTES);
            ptx::tma_load_2d(&p.tmR, &res_full[slot], smem_res + slot * CHUNK_BYTES, n2 * BLOCK_N2 + c * 64, t * BLOCK_M);
          }
      }
    }
  } else if (warp < 8) {
    // ============================ epilogue: two warpgroups (warps 4-7, 8-11) ============================
    // Both warpgroups cover all 128 rows (TMEM lane quarter = warp % 4) and split the COLUMNS: in E1 each converts
    // half of acc1, in E2 warpgroup g owns the 64-column chunk g of every 128-column n2 tile (its own residual
    // ring slots, bias slice, staging buffers and TMA stores).
    const int wg = warp >> 2;             // 0 or 1
    const int et = threadIdx.x - wg * 128;   // 0..127 inside the warpgroup
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int bar_id = 1 + wg;
    for (int i = threadIdx.x; i < MID; i += 256) s_bias2[i] = p.bias2[i];
    for (int i = threadIdx.x; i < p.width; i += 256) s_bias3[i] = p.bias3[i];   // (was reloaded per n2 tile behind a 128-thread barrier)
    ptx::named_bar_sync(3, 256);

    int a1 = 0, a2 = 0;
    uint32_t a1_phase = 0, a2_phase = 0, rk = 0, free_phase = 0;
    int prev_slot = -1;   // slot whose TMA store may still be reading it
    const bool early_release = p.early_release != 0;
    long long c_acc1 = 0, c_free = 0, c_e1 = 0, c_bar = 0, c_acc2 = 0, c_res = 0, c_math = 0, c_store = 0;
    const long long c_begin = pclk();

    // The slot an output chunk was stored from goes back to the residual producer as soon as that store has READ it.  Waiting
    // for that at the end of the NEXT step (tma_store_wait_read<1>) left the producer no lead: every residual load was issued
    // just before its data was needed and the epilogue waited a full load latency per step (43 % of the kernel, measured with
    // -DOPD_BNECK_PROBE).  Now the store thread releases the slot early in the next step (the store has had the barrier and the
    // accumulator wait to drain), a whole step before the slot's next chunk is consumed.
    auto release_prev = [&]() {
      if (early_release && et == 0 && prev_slot >= 0) {
        ptx::tma_store_wait_read<0>();
        ptx::mbar_arrive(&res_empty[prev_slot]);
      }
      if (early_release) prev_slot = -1;
    };

    auto e1 = [&]() {
      if (wg == 0 && et == 0) trace_ev(p.trace, 20);
      const long long q0 = pclk();
      ptx::mbar_wait(&acc1_full[a1], a1_phase);
      if (wg == 0 && et == 0) trace_ev(p.trace, 21);
      release_prev();
      const long long q1 = pclk();
      ptx::mbar_wait(a2_free, free_phase ^ 1);   // the previous tile's second GEMM no longer reads A2
      if (wg == 0 && et == 0) trace_ev(p.trace, 22);
      const long long q2 = pclk();
      c_acc1 += q1 - q0;
      c_free += q2 - q1;
      free_phase ^= 1;
      ptx::tc_fence_after_sync();
      const uint32_t t_acc = tmem_acc1 + lane_addr + a1 * MID;
      constexpr int kUnits = MID / 64;           // 32-column units per warpgroup
#pragma unroll
      for (int uu = 0; uu < kUnits; ++uu) {
        const int u = wg * kUnits + uu;          // global 32-column unit: chunk u / 2, half u % 2
        uint32_t v[32];
        ptx::tmem_ld_32x32(t_acc + u * 32, v);
        ptx::tmem_ld_wait();
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          packed[j] = ptx::epi_bias_relu2(v[2 * j], v[2 * j + 1], *reinterpret_cast<const float2*>(s_bias2 + u * 32 + 2 * j));
        uint8_t* rowp = smem_a2 + (u >> 1) * CHUNK_BYTES + row * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(rowp + ((((u & 1) * 4 + j) ^ (row & 7)) << 4)) =
              make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
      }
      ptx::tc_fence_before_sync();
      if (kPair) {
        // Every thread's A2 rows are complete and visible to the async proxy of ITS OWN SM (whose tensor core is the one that
        // reads them) before the warp's arrival leaves for the leader.  Plain remote arrivals, as CUTLASS's 2-SM kernels signal
        // the leader after fence.proxy.async; .release.cluster on the arrival cost ~1500 cycles per tile (measured).
        ptx::fence_proxy_async_smem();           // A2 is read by the tensor core through the async proxy
        __syncwarp();
        if (lane == 0) {
          ptx::mbar_arrive_cluster(&acc1_empty[a1], 0);
          ptx::mbar_arrive_cluster(a2_ready, 0);
        }
      } else {
        ptx::mbar_arrive(&acc1_empty[a1]);
        ptx::fence_proxy_async_smem();             // A2 is read by the tensor core through the async proxy
        ptx::mbar_arrive(a2_ready);
      }
      if (wg == 0 && et == 0) trace_ev(p.trace, 23);
      c_e1 += pclk() - q2;
      if (++a1 == 2) {
        a1 = 0;
        a1_phase ^= 1;
      }
    };

    auto e2 = [&](int m_blk, int n2) {
      const int m0 = m_blk * BLOCK_M, n0 = n2 * BLOCK_N2 + wg * 64;
      const long long q0 = pclk();
      const float* my_bias3 = s_bias3 + n0;
      if (wg == 0 && et == 0) trace_ev(p.trace, 30);
      const long long q1 = pclk();
      ptx::mbar_wait(&acc2_full[a2], a2_phase);
      if (wg == 0 && et == 0) trace_ev(p.trace, 31);
      release_prev();
      const long long q2 = pclk();
      ptx::tc_fence_after_sync();
      const uint32_t t_acc = tmem_acc2 + lane_addr + a2 * BLOCK_N2 + wg * 64;
      uint32_t packed[32];
      const int rs = wg * 2 + (rk & 1);
      ptx::mbar_wait(&res_full[rs], (rk >> 1) & 1);
      if (wg == 0 && et == 0) trace_ev(p.trace, 32);
      const long long q3 = pclk();
      ++rk;
      const uint8_t* rrow = smem_res + rs * CHUNK_BYTES + row * 128;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(t_acc + h * 32, v);
        uint4 rr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) rr[j] = *reinterpret_cast<const uint4*>(rrow + (((h * 4 + j) ^ (row & 7)) << 4));
        ptx::tmem_ld_wait();
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(rr);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          packed[h * 16 + j] = ptx::epi_bias_res_relu2(v[2 * j], v[2 * j + 1], *reinterpret_cast<const float2*>(my_bias3 + h * 32 + 2 * j), rw[j]);
      }
      // accumulator and residual chunk are in registers: hand both back before the store path
      ptx::tc_fence_before_sync();
      if (kPair) {
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(&acc2_empty[a2], 0);
      } else {
        ptx::mbar_arrive(&acc2_empty[a2]);
      }
      if (++a2 == 2) {
        a2 = 0;
        a2_phase ^= 1;
      }
      const long long q4 = pclk();
      // The output chunk overwrites the residual chunk in place (every thread rewrites exactly the 128 bytes it has just
      // read) and is stored from there; the slot goes back to the residual producer once the store has read it.
      uint8_t* rowp = smem_res + rs * CHUNK_BYTES + row * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(rowp + ((j ^ (row & 7)) << 4)) =
            make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
      ptx::fence_proxy_async_smem();
      ptx::named_bar_sync(bar_id, 128);
      if (et == 0) {
        ptx::tma_store_2d(&p.tmD, smem_res + rs * CHUNK_BYTES, n0, m0);
        ptx::tma_store_commit();
        if (!early_release && prev_slot >= 0) {
          ptx::tma_store_wait_read<1>();   // every store but the one just issued has finished reading shared memory
          ptx::mbar_arrive(&res_empty[prev_slot]);
        }
      }
      prev_slot = rs;
      if (wg == 0 && et == 0) trace_ev(p.trace, 33);
      c_bar += q1 - q0;
      c_acc2 += q2 - q1;
      c_res += q3 - q2;
      c_math += q4 - q3;
      c_store += pclk() - q4;
    };

    // E1 of the NEXT tile runs before E2 of this one, so that the next tile's second GEMM overlaps this tile's output phase
    // (G2 of a tile needs both acc2 buffers back for its n2 >= 2 tiles, so E1(next) goes before the LAST TWO E2's.)
    if (first < n_mblk) e1();
    const int pre = p.num_n2 > 2 ? p.num_n2 - 2 : 0;
    for (int t = first; t < n_mblk; t += step) {
      for (int n2 = 0; n2 < pre; ++n2) e2(t, n2);
      if (t + step < n_mblk) e1();
      for (int n2 = pre; n2 < p.num_n2; ++n2) e2(t, n2);
    }
    if (kBneckProbe && (blockIdx.x == 0 || blockIdx.x == 77) && (et == 0 || et == 127))
      printf("bneck CTA %d wg %d thread %d: %lld cycles; E1: acc1_full wait %lld, a2_free wait %lld, convert %lld; E2: bias + barrier %lld, acc2_full wait %lld, "
             "res_full wait %lld, math %lld, st.shared + barrier + store (+ read wait) %lld\n",
             (int)blockIdx.x, wg, et, pclk() - c_begin, c_acc1, c_free, c_e1, c_bar, c_acc2, c_res, c_math, c_store);
    if (et == 0) ptx::tma_store_wait_all<0>();
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (kPair) ptx::cluster_sync();   // no CTA leaves while its peer may still arrive on its barriers or read its shared memory
  if (warp == 9) {
    if (kPair) ptx::tmem_dealloc_2sm<C::kTmemCols>(tmem_base);
    else ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

unsigned long long* g_bneck_trace = nullptr;

template <int MID>
int launch_t(const BneckParams& p, int grid, cudaStream_t s) {
  static PerDeviceOnce configured;
  auto kern = tc_bneck_kernel<MID, false>;
  if (int rc = once_per_device(configured, [&]() -> int {
        OPD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<MID>::kSmemBytes));
        return OPD_OK;
      }))
    return rc;
  kern<<<grid, kThreads, Cfg<MID>::kSmemBytes, s>>>(p);
  count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

// cta_group::2 variant: clusters of two CTAs, as many pairs as fit (GPCs with an odd SM count leave one SM without a partner)
int launch_pair(const BneckParams& p, int grid, cudaStream_t s) {
  auto kern = tc_bneck_kernel<128, true>;
  static PerDeviceInt cluster_limit;
  int max_clusters = -1;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg<128, true>::kSmemBytes;
  cfg.stream = s;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (int rc = cached_per_device(cluster_limit, &max_clusters, [&](int* n) -> int {
        OPD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128, true>::kSmemBytes));
        cfg.gridDim = dim3(sm_count() / 2 * 2);
        OPD_CUDA_OK(cudaOccupancyMaxActiveClusters(n, kern, &cfg));
        return OPD_OK;
      }))
    return rc;
  OPD_REQUIRE(max_clusters > 0, "bottleneck tail: no 2-CTA cluster of the kernel fits on this device");
  cfg.gridDim = dim3(2 * std::min(grid / 2, max_clusters));
  OPD_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p));
  count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

}  // namespace

void g_bneck_trace_set(unsigned long long* p) { g_bneck_trace = p; }

int bneck_plan(BneckPlan* plan, const __nv_bfloat16* x, const ConvGeom& g, const __nv_bfloat16* w2, const float* bias2,
               const __nv_bfloat16* w3, const float* bias3, int width, const __nv_bfloat16* residual, __nv_bfloat16* y) {
  if (g_option_bneck_halo.load() && (g.C == 64 || (g.C == 128 && g_option_bneck_halo.load() > 1)) && g.stride == 1 && g.KH == 3 && g.KW == 3 && g.pad_h == 1 && g.pad_w == 1)
    return bneck_halo_plan(plan, x, g, w2, bias2, w3, bias3, width, residual, y);
  *plan = BneckPlan{};
  OPD_REQUIRE(g.KH == 3 && g.KW == 3 && (g.C == 64 || g.C == 128), "bottleneck tail: 3x3 convolution over 64 or 128 channels (C=%d)", g.C);
  OPD_REQUIRE(width % BLOCK_N2 == 0 && width > 0 && width <= 512, "bottleneck tail: width=%d must be a multiple of 128, <= 512", width);
  OPD_REQUIRE(bias2 && bias3 && residual && y && x && w2 && w3, "bottleneck tail: NULL argument");
  plan->M = g.B * g.P * g.Q;
  plan->mid = g.C;
  plan->width = width;
  plan->g = g;
  plan->bias2 = bias2;
  plan->bias3 = bias3;
  if (int rc = make_tmap_im2col(&plan->tmA, x, g)) return rc;
  const int m_blocks = (plan->M + BLOCK_M - 1) / BLOCK_M;
  const int pr = g_option_bneck_pair.load();
  plan->pair = pr && g.C == 128 && m_blocks % 2 == 0 && (pr == 3 || m_blocks >= sm_count());
  const int halves = plan->pair ? 2 : 1;   // cta_group::2: each CTA of the pair loads half of the output channels of a weight tile
  if (int rc = make_tmap_2d(&plan->tmB1, w2, g.C, 9 * g.C, 9 * g.C, g.C / halves)) return rc;
  if (int rc = make_tmap_2d(&plan->tmB2, w3, width, g.C, g.C, BLOCK_N2 / halves)) return rc;
  if (int rc = make_tmap_2d(&plan->tmR, residual, plan->M, width, width, BLOCK_M)) return rc;
  if (int rc = make_tmap_2d(&plan->tmD, y, plan->M, width, width, BLOCK_M)) return rc;
  plan->grid = std::min(m_blocks, sm_count());
  return OPD_OK;
}

int bneck_launch(const BneckPlan& plan, cudaStream_t stream) {
  if (plan.halo) return bneck_halo_launch(plan, stream);
  BneckParams p;
  p.tmA = plan.tmA; p.tmB1 = plan.tmB1; p.tmB2 = plan.tmB2; p.tmR = plan.tmR; p.tmD = plan.tmD;
  p.M = plan.M; p.width = plan.width;
  p.num_m_blocks = (plan.M + BLOCK_M - 1) / BLOCK_M;
  p.num_n2 = plan.width / BLOCK_N2;
  p.KW = plan.g.KW; p.stride = plan.g.stride; p.pad_h = plan.g.pad_h; p.pad_w = plan.g.pad_w; p.P = plan.g.P; p.Q = plan.g.Q;
  p.bias2 = plan.bias2; p.bias3 = plan.bias3;
  p.trace = g_bneck_trace;
  p.early_release = (g_option_bneck_release.load() & 1) != 0;
  if (plan.pair) return launch_pair(p, plan.grid, stream);
  return plan.mid == 64 ? launch_t<64>(p, plan.grid, stream) : launch_t<128>(p, plan.grid, stream);
}

}  // namespace opd

extern "C" int opd_bottleneck_tail_bf16(const void* x_dev, int32_t B, int32_t H, int32_t W, int32_t mid, const void* w2_dev,
                                        const float* bias2_dev, int32_t stride, const void* w3_dev, const float* bias3_dev,
                                        int32_t width, const void* residual_dev, void* y_dev, void* stream) {
  opd::ConvGeom g{B, H, W, mid, 3, 3, stride, 1, 1, (H + 2 - 3) / stride + 1, (W + 2 - 3) / stride + 1};
  opd::BneckPlan plan;
  if (int rc = opd::bneck_plan(&plan, static_cast<const __nv_bfloat16*>(x_dev), g, static_cast<const __nv_bfloat16*>(w2_dev),
                               bias2_dev, static_cast<const __nv_bfloat16*>(w3_dev), bias3_dev, width,
                               static_cast<const __nv_bfloat16*>(residual_dev), static_cast<__nv_bfloat16*>(y_dev)))
    return rc;
  return opd::bneck_launch(plan, static_cast<cudaStream_t>(stream));
}

#ifdef OPD_BNECK_PROBE
// measurement builds only (benchmarks/bneck_trace.py): timeline trace of CTA 0 of the next tc_bneck_kernel launches
// (device buffer of >= 4001 uint64, [0] zeroed)
extern "C" int opd_debug_set_bneck_trace(unsigned long long* buf_dev) {
  opd::g_bneck_trace_set(buf_dev);
  return 0;
}
#endif
