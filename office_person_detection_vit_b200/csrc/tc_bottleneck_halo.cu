// Fused bottleneck tail, halo variant (3x3 stride 1, MID = 64 channels: ResNet stage 1):
//     y = relu( conv1x1( relu(conv3x3(x) + b2) ) + b3 + residual )            (or + Wsc * block_input: fused projection shortcut)
// Same pipeline as tc_bottleneck.cu, but the A operand of the 3x3 convolution is NOT nine im2col copies.  One work tile is
// a 16-row x 16-column pixel patch = two 128-pixel MMA tiles side by side (h = 0: columns 0-7, h = 1: columns 8-15).  Its
// 18 x 18 x 64-channel input halo is loaded ONCE by a tiled TMA (zero fill = the convolution's padding) and every filter
// tap reads it through a shifted shared-memory descriptor (start + (r * 18 + s + 8h) * 128 B, 8-pixel groups one patch row
// = 2304 B apart; the 128-byte swizzle is a function of the absolute shared-memory address, so shifted views stay
// consistent with what TMA wrote).  Both half tiles share every weight tile that streams through the B ring, which halves
// the L2 -> SM weight traffic that paced the one-tile version (profiles/README.md).
// Output, residual and shortcut-input tiles are 16 x 8 pixel patches moved by 4D TMA boxes [64 ch, 8, 16, 1].
//
// Warps: 0-7 epilogue (two warpgroups; E1: warpgroup g converts half tile g; E2: warpgroup g owns 64-column chunk g),
//        8 TMA producer (patch + W2 / W3 tiles), 9 MMA issuer + TMEM owner, 10 residual (or shortcut-input) producer;
//        the single-thread roles are elected with elect.sync.
#include <algorithm>

#include "opd_common.h"
#include "sm100_ptx.cuh"
#include "tc_gemm.h"

namespace opd {
namespace {

constexpr int BLOCK_M = 128, BLOCK_N2 = 128, UMMA_K = 16;
constexpr int SUB_W = 8, TILE_W = 16, TILE_H = 16, PATCH_W = TILE_W + 2, PATCH_H = TILE_H + 2;
constexpr int PATCH_BYTES = PATCH_W * PATCH_H * 128;   // 41472
constexpr int PATCH_SLOT = 41984;                      // 41 KB
constexpr int CHUNK_BYTES = BLOCK_M * 64 * 2;          // 16 KB: one [128 x 64] bf16 box
constexpr int kBStages = 3;
constexpr int kThreads = 384;
constexpr int kTmemCols = 512;   // acc1[h] 2 x MID + acc2[h] 2 x 128
// MID = 64: A2[h] one chunk each, residual 2 slots per warpgroup.  MID = 128 (stage 2): two 64-channel patch slabs per
// tile through the same patch buffer, A2[h] two chunks each, residual 1 slot per warpgroup.  Same shared-memory total.
template <int MID>
struct HaloCfg {
  static constexpr int kCB = MID / 64;                    // 64-channel blocks
  static constexpr int kA2Chunks = 2 * kCB;               // both half tiles
  // pool of [128 x 64] chunk slots after A2: residual chunks land here and the output chunk is written IN PLACE over the
  // residual it has just consumed, then stored from there (kSC, no residual: 2 staging boxes + the 2 block-input tiles)
  static constexpr int kPool = MID == 64 ? 6 : 4;
  static constexpr int kResStages = kPool;
  static constexpr int kResPerWg = kPool / 2;
  static constexpr int kTapBytes = MID * 128;             // one W2 tap tile [MID x 64] bf16
  static constexpr int kTapsPerSlot = CHUNK_BYTES / kTapBytes;
  static constexpr int kSmemBytes = PATCH_SLOT + kBStages * CHUNK_BYTES + kA2Chunks * CHUNK_BYTES + kPool * CHUNK_BYTES + 4096;
  static_assert(MID == 64 || MID == 128, "halo bottleneck tail: MID must be 64 or 128");
  static_assert(kSmemBytes <= 232448, "shared memory budget");
};

struct HaloParams {
  CUtensorMap tmA, tmB1, tmB2, tmR, tmD;
  int tiles_x, tiles_y, num_tiles, num_n2;
  const float* bias2;
  const float* bias3;
  int early_release;   // residual / output slots go back to the producer early in the next epilogue step (see tc_bottleneck.cu)
};

#ifdef OPD_BNECK_PROBE
constexpr bool kHaloProbe = true;
#else
constexpr bool kHaloProbe = false;   // -DOPD_BNECK_PROBE: clock64 counters of the epilogue's waits, printed by two CTAs
#endif
__device__ __forceinline__ long long hclk() { return kHaloProbe ? clock64() : 0; }

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(ptx::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// K-major, 128-byte swizzle, explicit stride between 8-row groups
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// kSC: the block's 1x1 projection shortcut is computed here too (second k-block of the second GEMM: the block input
// tile times Wsc), so the residual tensor is neither written by a separate kernel nor read back.
template <int MID, bool kSC>
__global__ void __launch_bounds__(kThreads, 1) tc_bneck_halo_kernel(const __grid_constant__ HaloParams p) {
  using C = HaloCfg<MID>;
  constexpr int kCB = C::kCB, kResStages = C::kResStages, kResPerWg = C::kResPerWg;
  constexpr uint32_t kIdesc1 = ptx::umma_idesc_bf16(BLOCK_M, MID);
  constexpr uint32_t kIdesc2 = ptx::umma_idesc_bf16(BLOCK_M, BLOCK_N2);
  constexpr int kB2Blocks = kCB + (kSC ? 1 : 0);   // k-blocks of the second GEMM: W3 (kCB) [+ Wsc]
  static_assert(!kSC || MID == 64, "fused shortcut: 64-channel blocks only");

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_patch = smem;                                      // one halo patch (the next one loads during E2)
  uint8_t* smem_b = smem_patch + PATCH_SLOT;                       // [kBStages] W2 tap pairs / W3 tiles
  uint8_t* smem_a2 = smem_b + kBStages * CHUNK_BYTES;              // [2 halves][kCB chunks] A operand of the second GEMM
  uint8_t* smem_pool = smem_a2 + C::kA2Chunks * CHUNK_BYTES;       // chunk slots (see HaloCfg::kPool)
  uint8_t* smem_out = smem_pool;                                   // kSC: one staging box per epilogue warpgroup (slots 0, 1)
  uint8_t* smem_res = kSC ? smem_pool + 2 * CHUNK_BYTES : smem_pool;   // kSC: X[h] in pool slots 2, 3; else all slots = residual ring
  float* s_bias2 = reinterpret_cast<float*>(smem_pool + C::kPool * CHUNK_BYTES);    // [MID]
  float* s_bias3 = s_bias2 + 128;                                                   // [width <= 512]: whole layer, loaded once
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias3 + 512);
  uint64_t* patch_full = bars;          // [1]
  uint64_t* patch_empty = bars + 1;     // [1]
  uint64_t* b_full = bars + 2;          // [kBStages]
  uint64_t* b_empty = bars + 6;         // [kBStages]
  uint64_t* acc1_full = bars + 10;      // [1]  (both halves)
  uint64_t* acc1_empty = bars + 11;     // [1]
  uint64_t* acc2_full = bars + 12;      // [2]  per half tile
  uint64_t* acc2_empty = bars + 14;     // [2]
  uint64_t* a2_ready = bars + 16;       // [1]
  uint64_t* a2_free = bars + 17;        // [1]
  uint64_t* res_full = bars + 18;       // [kPool]
  uint64_t* res_empty = bars + 24;      // [kPool]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 30);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();

  if (warp == 8 && lane == 0) {
    ptx::prefetch_tmap(&p.tmA);
    ptx::prefetch_tmap(&p.tmB1);
    ptx::prefetch_tmap(&p.tmB2);
    ptx::prefetch_tmap(&p.tmD);
    ptx::mbar_init(patch_full, 1);
    ptx::mbar_init(patch_empty, 1);
    for (int i = 0; i < kBStages; ++i) {
      ptx::mbar_init(&b_full[i], 1);
      ptx::mbar_init(&b_empty[i], 1);
    }
    ptx::mbar_init(acc1_full, 1);
    ptx::mbar_init(acc1_empty, 256);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&acc2_full[i], 1);
      ptx::mbar_init(&acc2_empty[i], 256);
    }
    ptx::mbar_init(a2_ready, 256);
    ptx::mbar_init(a2_free, 1);
    for (int i = 0; i < kResStages; ++i) {
      ptx::mbar_init(&res_full[i], 1);
      ptx::mbar_init(&res_empty[i], 1);   // kSC: MMA commit; else: the warpgroup's store thread, once the TMA store has read the slot
    }
    ptx::fence_barrier_init();
  }
  if (warp == 9) ptx::tmem_alloc<kTmemCols>(tmem_ptr);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_acc1 = tmem_base;             // [h] 64 columns each
  const uint32_t tmem_acc2 = tmem_base + 2 * MID;   // [h] 128 columns each

  const int first = blockIdx.x, step = gridDim.x, n_tiles = p.num_tiles;
  auto tile_origin = [&](int t, int& b, int& y0, int& x0) {
    const int tx = t % p.tiles_x;
    const int r = t / p.tiles_x;
    x0 = tx * TILE_W;
    y0 = (r % p.tiles_y) * TILE_H;
    b = r / p.tiles_y;
  };

  if (warp == 8) {
    // ===================================== TMA producer =====================================
    if (ptx::elect_one()) {
      int bs = 0;
      uint32_t bphase = 0, pc = 0;
      auto next_b = [&]() {
        if (++bs == kBStages) {
          bs = 0;
          bphase ^= 1;
        }
      };
      auto load_g1 = [&](int t) {   // the tile's halo patch (per 64-channel slab) and its W2 tap tiles
        int b, y0, x0;
        tile_origin(t, b, y0, x0);
        for (int cb = 0; cb < kCB; ++cb, ++pc) {
          ptx::mbar_wait(patch_empty, (pc & 1) ^ 1);
          ptx::mbar_expect_tx(patch_full, PATCH_BYTES);
          tma_load_4d(&p.tmA, patch_full, smem_patch, cb * 64, x0 - 1, y0 - 1, b);
          for (int tap = 0; tap < 9; tap += C::kTapsPerSlot) {   // W2 tap tiles [MID x 64]: two per 16 KB slot when MID = 64
            const int n_taps = tap + C::kTapsPerSlot <= 9 ? C::kTapsPerSlot : 9 - tap;
            ptx::mbar_wait(&b_empty[bs], bphase ^ 1);
            ptx::mbar_expect_tx(&b_full[bs], n_taps * C::kTapBytes);
            for (int j = 0; j < n_taps; ++j)
              ptx::tma_load_2d(&p.tmB1, &b_full[bs], smem_b + bs * CHUNK_BYTES + j * C::kTapBytes, (tap + j) * MID + cb * 64, 0);
            next_b();
          }
        }
      };
      auto load_g2 = [&](int n2) {
        for (int kb = 0; kb < kB2Blocks; ++kb) {   // kb >= kCB: the shortcut weights (last 64 columns of [W3 | Wsc])
          ptx::mbar_wait(&b_empty[bs], bphase ^ 1);
          ptx::mbar_expect_tx(&b_full[bs], CHUNK_BYTES);
          ptx::tma_load_2d(&p.tmB2, &b_full[bs], smem_b + bs * CHUNK_BYTES, kb * 64, n2 * BLOCK_N2);
          next_b();
        }
      };
      // Same order as the MMA issuer consumes the ring (see there): G1 of the next tile goes before the last n2 tile of this one.
      if (first < n_tiles) load_g1(first);
      for (int t = first; t < n_tiles; t += step) {
        for (int n2 = 0; n2 < p.num_n2 - 1; ++n2) load_g2(n2);
        if (t + step < n_tiles) load_g1(t + step);
        load_g2(p.num_n2 - 1);
      }
    }
  } else if (warp == 9) {
    // ===================================== MMA issuer =====================================
    if (ptx::elect_one()) {
      int bs = 0;
      uint32_t bphase = 0, n = 0, pc = 0, acc2_phase[2] = {0, 0};
      auto next_b = [&]() {
        if (++bs == kBStages) {
          bs = 0;
          bphase ^= 1;
        }
      };
      const uint64_t patch_desc = desc_sw128(ptx::smem_u32(smem_patch), PATCH_W * 128);
      // Issue order per tile i:  G2(i, 0 .. last-1)  G1(i+1)  G2(i, last)   (G1(0) first).
      // The epilogue converts tile i+1's mid activation (E1) between the two half tiles of tile i's last n2 tile, so G2(i+1, 0)
      // can start while the epilogue is still on tile i: with the strict G1(i) G2(i) order every tile ended in a bubble of
      // about 1600 cycles (E1, then the first second-GEMM tile, with the epilogue idle; -DOPD_BNECK_PROBE: acc2_full wait 13 %).
      uint32_t n1 = 0;   // tiles whose G1 has been issued
      auto g1 = [&]() {
        // ---- G1: both half tiles, every W2 tap tile used twice; one 64-channel patch slab at a time ----
        ptx::mbar_wait(acc1_empty, (n1 & 1) ^ 1);
        for (int cb = 0; cb < kCB; ++cb, ++pc) {
          ptx::mbar_wait(patch_full, pc & 1);
          ptx::tc_fence_after_sync();
          for (int tap0 = 0; tap0 < 9; tap0 += C::kTapsPerSlot) {
            ptx::mbar_wait(&b_full[bs], bphase);
            ptx::tc_fence_after_sync();
            for (int tap = tap0; tap < tap0 + C::kTapsPerSlot && tap < 9; ++tap) {
              const int r = tap / 3, s = tap - r * 3;
              // descriptors differ only in the 16-byte-granular start address field: base descriptor + small offsets
              const uint64_t db = desc_sw128(ptx::smem_u32(smem_b + bs * CHUNK_BYTES + (tap - tap0) * C::kTapBytes), 1024);
              const uint64_t da = patch_desc + (uint64_t)(((r * PATCH_W + s) * 128) >> 4);
              // the two half tiles alternate MMA by MMA: a dependent accumulate into the same TMEM tile waits for the
              // previous MMA's full latency (~100 cycles), three times the 32 cycles a 128x64x16 MMA occupies the pipe
#pragma unroll
              for (int k = 0; k < 64 / UMMA_K; ++k)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                  ptx::umma_bf16_ss(tmem_acc1 + h * MID, da + (uint64_t)((h * SUB_W * 128 + k * 32) >> 4), db + (uint64_t)(2 * k), kIdesc1,
                                    (cb | tap | k) != 0);
            }
            ptx::umma_commit(&b_empty[bs]);
            next_b();
          }
          ptx::umma_commit(patch_empty);   // the slab may be overwritten once these MMAs have read it
        }
        ptx::umma_commit(acc1_full);
        ++n1;
      };
      // ---- G2: per n2 tile both half tiles share the W3 (and Wsc) tile ----
      auto g2 = [&](int n2, uint32_t par) {
        if (n2 == 0) {
          ptx::mbar_wait(a2_ready, par);
          if (kSC) {
            ptx::mbar_wait(&res_full[0], par);
            ptx::mbar_wait(&res_full[1], par);
          }
          ptx::tc_fence_after_sync();
        }
        if (kB2Blocks == 1 && n2 == 0) {
          // first n2 tile of a tile: half tile 0's accumulator is free one epilogue step before half tile 1's - issue (and
          // commit) the halves separately so that the epilogue's first step of the tile does not wait for both
          ptx::mbar_wait(&b_full[bs], bphase);
          const uint64_t db = desc_sw128(ptx::smem_u32(smem_b + bs * CHUNK_BYTES), 1024);
          for (int h = 0; h < 2; ++h) {
            ptx::mbar_wait(&acc2_empty[h], acc2_phase[h] ^ 1);
            acc2_phase[h] ^= 1;
            ptx::tc_fence_after_sync();
            const uint64_t da = desc_sw128(ptx::smem_u32(smem_a2 + (h * kCB) * CHUNK_BYTES), 1024);
#pragma unroll
            for (int k = 0; k < 64 / UMMA_K; ++k)
              ptx::umma_bf16_ss(tmem_acc2 + h * BLOCK_N2, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), kIdesc2, k != 0);
            ptx::umma_commit(&acc2_full[h]);
          }
          ptx::umma_commit(&b_empty[bs]);
          next_b();
          return;
        }
        for (int kb = 0; kb < kB2Blocks; ++kb) {
          ptx::mbar_wait(&b_full[bs], bphase);
          const uint32_t b_addr = ptx::smem_u32(smem_b + bs * CHUNK_BYTES);
          uint64_t da[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (kb == 0) {
              ptx::mbar_wait(&acc2_empty[h], acc2_phase[h] ^ 1);
              acc2_phase[h] ^= 1;
            }
            da[h] = desc_sw128(ptx::smem_u32(kb < kCB ? smem_a2 + (h * kCB + kb) * CHUNK_BYTES : smem_res + h * CHUNK_BYTES), 1024);
          }
          ptx::tc_fence_after_sync();
          const uint64_t db = desc_sw128(b_addr, 1024);
#pragma unroll
          for (int k = 0; k < 64 / UMMA_K; ++k)
#pragma unroll
            for (int h = 0; h < 2; ++h)   // alternate the two accumulators (see G1)
              ptx::umma_bf16_ss(tmem_acc2 + h * BLOCK_N2, da[h] + (uint64_t)(2 * k), db + (uint64_t)(2 * k), kIdesc2, (kb | k) != 0);
          if (kb == kB2Blocks - 1) {
            ptx::umma_commit(&acc2_full[0]);
            ptx::umma_commit(&acc2_full[1]);
          }
          ptx::umma_commit(&b_empty[bs]);
          next_b();
        }
      };
      if (first < n_tiles) g1();
      for (int t = first; t < n_tiles; t += step, ++n) {
        const uint32_t par = n & 1;
        for (int n2 = 0; n2 < p.num_n2 - 1; ++n2) g2(n2, par);
        if (t + step < n_tiles) g1();
        g2(p.num_n2 - 1, par);
        ptx::umma_commit(a2_free);
        if (kSC) {
          ptx::umma_commit(&res_empty[0]);
          ptx::umma_commit(&res_empty[1]);
        }
      }
    }
  } else if (warp == 10) {
    // ===================================== residual / shortcut-input TMA producer =====================================
    if (ptx::elect_one()) {
      ptx::prefetch_tmap(&p.tmR);
      uint32_t k = 0, n = 0;   // k: chunk counter per warpgroup: slot = wg * 2 + (k & 1), parity = (k >> 1) & 1
      for (int t = first; t < n_tiles; t += step, ++n) {
        int b, y0, x0;
        tile_origin(t, b, y0, x0);
        if (kSC) {   // block-input tile of each half tile (slots 0 / 1, released by the MMA warp's commit)
          for (int h = 0; h < 2; ++h) {
            ptx::mbar_wait(&res_empty[h], (n & 1) ^ 1);
            ptx::mbar_expect_tx(&res_full[h], CHUNK_BYTES);
            tma_load_4d(&p.tmR, &res_full[h], smem_res + h * CHUNK_BYTES, 0, x0 + h * SUB_W, y0, b);
          }
          continue;
        }
        for (int n2 = 0; n2 < p.num_n2; ++n2)
          for (int h = 0; h < 2; ++h, ++k)
            for (int c = 0; c < 2; ++c) {   // chunk c of the n2 tile belongs to epilogue warpgroup c
              const int slot = c * kResPerWg + (k % kResPerWg);
              ptx::mbar_wait(&res_empty[slot], ((k / kResPerWg) & 1) ^ 1);
              ptx::mbar_expect_tx(&res_full[slot], CHUNK_BYTES);
              tma_load_4d(&p.tmR, &res_full[slot], smem_res + slot * CHUNK_BYTES, n2 * BLOCK_N2 + c * 64, x0 + h * SUB_W, y0, b);
            }
      }
    }
  } else if (warp < 8) {
    // ============================ epilogue: two warpgroups (warps 0-3, 4-7) ============================
    const int wg = warp >> 2;
    const int et = threadIdx.x - wg * 128;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;           // half-tile row = pixel (row / 8, row % 8) = TMEM lane
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int bar_id = 1 + wg;
    if (threadIdx.x < MID) s_bias2[threadIdx.x] = p.bias2[threadIdx.x];   // 256 epilogue threads >= MID
    for (int i = threadIdx.x; i < p.num_n2 * BLOCK_N2; i += 256) s_bias3[i] = p.bias3[i];
    ptx::named_bar_sync(3, 256);

    uint32_t rk = 0, n = 0, acc2_phase[2] = {0, 0};
    uint8_t* my_out = smem_out + wg * CHUNK_BYTES;
    int prev_slot = -1;   // residual slot whose TMA store may still be reading it
    const bool early_release = !kSC && p.early_release != 0;
    // The slot an output chunk was stored from is released as soon as the store has read it: early in the NEXT step, not at its
    // end, so that the slot's next residual chunk is requested a whole step earlier (tc_bottleneck.cu: the residual wait was
    // 43 % of that kernel).
    auto release_prev = [&]() {
      if (early_release && et == 0 && prev_slot >= 0) {
        ptx::tma_store_wait_read<0>();
        ptx::mbar_arrive(&res_empty[prev_slot]);
      }
      if (early_release) prev_slot = -1;
    };

    long long c_acc1 = 0, c_free = 0, c_e1 = 0, c_acc2 = 0, c_res = 0, c_math = 0, c_store = 0;
    const long long c_begin = hclk();
    // ---- E1: warpgroup g converts half tile g: acc1[g] -> +b2, ReLU -> bf16 -> A2[g] ----
    auto e1 = [&](uint32_t par) {
      const long long q0 = hclk();
      ptx::mbar_wait(acc1_full, par);
      release_prev();
      const long long q1 = hclk();
      ptx::mbar_wait(a2_free, par ^ 1);           // the previous tile's second GEMM no longer reads A2
      const long long q2 = hclk();
      c_acc1 += q1 - q0;
      c_free += q2 - q1;
      ptx::tc_fence_after_sync();
      {
#pragma unroll
        for (int u = 0; u < MID / 32; ++u) {   // 32-column units: chunk u / 2, half u % 2
          uint32_t v[32];
          ptx::tmem_ld_32x32(tmem_acc1 + lane_addr + wg * MID + u * 32, v);
          ptx::tmem_ld_wait();
          uint32_t packed[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            packed[j] = ptx::epi_bias_relu2(v[2 * j], v[2 * j + 1], *reinterpret_cast<const float2*>(s_bias2 + u * 32 + 2 * j));
          uint8_t* rowp = smem_a2 + (wg * kCB + (u >> 1)) * CHUNK_BYTES + row * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(rowp + ((((u & 1) * 4 + j) ^ (row & 7)) << 4)) =
                make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
        }
      }
      ptx::tc_fence_before_sync();
      ptx::mbar_arrive(acc1_empty);
      ptx::fence_proxy_async_smem();             // A2 is read by the tensor core through the async proxy
      ptx::mbar_arrive(a2_ready);
      c_e1 += hclk() - q2;
    };

    // ---- E2 step (n2, h), in MMA order; warpgroup g owns columns [n2 * 128 + 64 g, + 64) ----
    auto e2 = [&](int n2, int h, int b, int y0, int x0) {
        const int n0 = n2 * BLOCK_N2 + wg * 64;
        const float* my_bias3 = s_bias3 + n0;
        {
          const long long r0 = hclk();
          ptx::mbar_wait(&acc2_full[h], acc2_phase[h]);
          acc2_phase[h] ^= 1;
          const long long r1 = hclk();
          ptx::tc_fence_after_sync();
          const uint32_t t_acc = tmem_acc2 + lane_addr + h * BLOCK_N2 + wg * 64;
          uint32_t packed[32];
          const int rslot = wg * kResPerWg + (rk % kResPerWg);
          const uint8_t* rrow = smem_res + rslot * CHUNK_BYTES + row * 128;
          if (!kSC) {
            ptx::mbar_wait(&res_full[rslot], (rk / kResPerWg) & 1);
            ++rk;
          }
          const long long r2 = hclk();
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(t_acc + u * 32, v);
            uint4 rr[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              rr[j] = kSC ? make_uint4(0, 0, 0, 0) : *reinterpret_cast<const uint4*>(rrow + (((u * 4 + j) ^ (row & 7)) << 4));
            ptx::tmem_ld_wait();
            const uint32_t* rw = reinterpret_cast<const uint32_t*>(rr);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 b2 = *reinterpret_cast<const float2*>(my_bias3 + u * 32 + 2 * j);
              packed[u * 16 + j] = kSC ? ptx::epi_bias_relu2(v[2 * j], v[2 * j + 1], b2) : ptx::epi_bias_res_relu2(v[2 * j], v[2 * j + 1], b2, rw[j]);
            }
          }
          ptx::tc_fence_before_sync();
          ptx::mbar_arrive(&acc2_empty[h]);
          release_prev();   // after this step's arithmetic: the previous store has had that long to read its slot
          const long long r3 = hclk();
          if (kSC) {
            // staging box: its previous TMA store must have finished READING it; waited for here, after the arithmetic
            if (et == 0) ptx::tma_store_wait_read<0>();
            ptx::named_bar_sync(bar_id, 128);
          }
          // non-kSC: the output chunk overwrites the residual chunk in place (every thread rewrites exactly the 128 bytes it
          // has just read) and is stored from there; the slot returns to the residual producer once the store has read it
          uint8_t* obuf = kSC ? my_out : smem_res + rslot * CHUNK_BYTES;
          uint8_t* rowp = obuf + row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(rowp + ((j ^ (row & 7)) << 4)) =
                make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          ptx::fence_proxy_async_smem();
          ptx::named_bar_sync(bar_id, 128);
          if (et == 0) {
            tma_store_4d(&p.tmD, obuf, n0, x0 + h * SUB_W, y0, b);
            ptx::tma_store_commit();
            if (!kSC && !early_release && prev_slot >= 0) {
              ptx::tma_store_wait_read<1>();   // every store but the one just issued has finished reading shared memory
              ptx::mbar_arrive(&res_empty[prev_slot]);
            }
          }
          prev_slot = rslot;
          c_acc2 += r1 - r0;
          c_res += r2 - r1;
          c_math += r3 - r2;
          c_store += hclk() - r3;
        }
    };

    // Tile i+1's conversion (E1) runs between the two half tiles of tile i's last n2 tile: its A2 hand-off lets the second GEMM of
    // tile i+1 start while the epilogue still has a step of tile i to do (the MMA issuer's order matches: see there).
    if (first < n_tiles) e1(0);
    for (int t = first; t < n_tiles; t += step, ++n) {
      int b, y0, x0;
      tile_origin(t, b, y0, x0);
      const int last = p.num_n2 - 1;
      for (int n2 = 0; n2 < last; ++n2) {
        e2(n2, 0, b, y0, x0);
        e2(n2, 1, b, y0, x0);
      }
      e2(last, 0, b, y0, x0);
      if (t + step < n_tiles) e1((n + 1) & 1);
      e2(last, 1, b, y0, x0);
    }
    if (kHaloProbe && (blockIdx.x == 0 || blockIdx.x == 77) && et == 0)
      printf("halo<%d,%d> CTA %d wg %d: %lld cycles, %u tiles; E1: acc1_full wait %lld, a2_free wait %lld, convert %lld; E2: acc2_full wait %lld, res_full wait %lld, "
             "math %lld, st.shared + barrier + store (+ read wait) %lld\n",
             MID, (int)kSC, (int)blockIdx.x, wg, hclk() - c_begin, n, c_acc1, c_free, c_e1, c_acc2, c_res, c_math, c_store);
    if (et == 0) ptx::tma_store_wait_all<0>();
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 9) ptx::tmem_dealloc<kTmemCols>(tmem_base);
}

}  // namespace

// shortcut_in != nullptr: fused projection shortcut; then w3 is [width, 128] = [W3 | Wsc], bias3 = b3 + b_sc, `residual` is
// ignored and shortcut_in is the block input [B, P, Q, 64].
int bneck_halo_plan(BneckPlan* plan, const __nv_bfloat16* x, const ConvGeom& g, const __nv_bfloat16* w2, const float* bias2,
                    const __nv_bfloat16* w3, const float* bias3, int width, const __nv_bfloat16* residual, __nv_bfloat16* y,
                    const __nv_bfloat16* shortcut_in) {
  *plan = BneckPlan{};
  if (shortcut_in) residual = shortcut_in;
  const int MID = g.C;
  OPD_REQUIRE(g.KH == 3 && g.KW == 3 && (MID == 64 || MID == 128) && g.stride == 1 && g.pad_h == 1 && g.pad_w == 1 && g.P == g.H &&
                  g.Q == g.W,
              "bottleneck tail (halo): 3x3 / stride 1 / pad 1 over 64 or 128 channels only");
  OPD_REQUIRE(!shortcut_in || MID == 64, "bottleneck tail (halo): fused shortcut needs 64 channels");
  OPD_REQUIRE(width % BLOCK_N2 == 0 && width > 0 && width <= 512, "bottleneck tail (halo): width=%d must be a multiple of 128, <= 512", width);
  OPD_REQUIRE(bias2 && bias3 && residual && y && x && w2 && w3, "bottleneck tail (halo): NULL argument");
  plan->halo = 1;
  plan->M = g.B * g.P * g.Q;
  plan->mid = g.C;
  plan->width = width;
  plan->g = g;
  plan->bias2 = bias2;
  plan->bias3 = bias3;
  if (int rc = make_tmap_nhwc_patch(&plan->tmA, x, g.B, g.H, g.W, MID, PATCH_W, PATCH_H)) return rc;
  if (int rc = make_tmap_2d(&plan->tmB1, w2, MID, 9 * MID, 9 * MID, MID)) return rc;
  plan->fused_shortcut = shortcut_in != nullptr;
  const int k2 = shortcut_in ? 2 * MID : MID;
  if (int rc = make_tmap_2d(&plan->tmB2, w3, width, k2, k2, BLOCK_N2)) return rc;
  if (int rc = make_tmap_nhwc_patch(&plan->tmR, residual, g.B, g.P, g.Q, shortcut_in ? MID : width, SUB_W, TILE_H)) return rc;
  if (int rc = make_tmap_nhwc_patch(&plan->tmD, y, g.B, g.P, g.Q, width, SUB_W, TILE_H)) return rc;
  const int tiles = g.B * ((g.P + TILE_H - 1) / TILE_H) * ((g.Q + TILE_W - 1) / TILE_W);
  plan->grid = std::min(tiles, sm_count());
  return OPD_OK;
}

int bneck_halo_launch(const BneckPlan& plan, cudaStream_t stream) {
  HaloParams p;
  p.tmA = plan.tmA; p.tmB1 = plan.tmB1; p.tmB2 = plan.tmB2; p.tmR = plan.tmR; p.tmD = plan.tmD;
  p.tiles_x = (plan.g.Q + TILE_W - 1) / TILE_W;
  p.tiles_y = (plan.g.P + TILE_H - 1) / TILE_H;
  p.num_tiles = plan.g.B * p.tiles_x * p.tiles_y;
  p.num_n2 = plan.width / BLOCK_N2;
  p.bias2 = plan.bias2;
  p.bias3 = plan.bias3;
  p.early_release = (g_option_bneck_release.load() & 2) != 0;
  static PerDeviceOnce configured;
  if (int rc = once_per_device(configured, []() -> int {
        OPD_CUDA_OK(cudaFuncSetAttribute(tc_bneck_halo_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, HaloCfg<64>::kSmemBytes));
        OPD_CUDA_OK(cudaFuncSetAttribute(tc_bneck_halo_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HaloCfg<64>::kSmemBytes));
        OPD_CUDA_OK(cudaFuncSetAttribute(tc_bneck_halo_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, HaloCfg<128>::kSmemBytes));
        return OPD_OK;
      }))
    return rc;
  if (plan.mid == 128)
    tc_bneck_halo_kernel<128, false><<<plan.grid, kThreads, HaloCfg<128>::kSmemBytes, stream>>>(p);
  else if (plan.fused_shortcut)
    tc_bneck_halo_kernel<64, true><<<plan.grid, kThreads, HaloCfg<64>::kSmemBytes, stream>>>(p);
  else
    tc_bneck_halo_kernel<64, false><<<plan.grid, kThreads, HaloCfg<64>::kSmemBytes, stream>>>(p);
  count_launch();
  OPD_CUDA_OK(cudaGetLastError());
  return OPD_OK;
}

}  // namespace opd
