// Shared host-side helpers of libopd_b200.so: thread-local error string, CUDA error mapping,
// kernel-launch counter.  Nothing here is part of the C ABI (see include/opd_b200.h).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>

#include "opd_b200.h"

namespace opd {

std::string& last_error_ref();
extern std::atomic<int64_t> g_launches;
extern std::atomic<int> g_option_attention_kv;  // keys per attention tile: 96 (default), 64 or 128; 65 = 64 with S in registers
extern std::atomic<int> g_option_attention_tc;  // 1 (default): tcgen05 attention kernel; 0: the mma.sync kernel
extern std::atomic<int> g_option_probe;        // measurement probes, 0 in production: bit 0 stem without patch reloads, bit 1 stem without stores
extern std::atomic<int> g_option_gemm_cluster; // 0 (default) / 1: BLOCK_N = 256 layers as 2-CTA clusters with multicast weight tiles (measured: no gain)
extern std::atomic<int> g_option_gemm_outbufs; // 1 (default): short-K GEMMs double-buffer the epilogue's staging boxes
extern std::atomic<int> g_option_pdl;          // 1 (default): GEMM / attention launches allow programmatic dependent launch (see opd_set_option)
extern std::atomic<int> g_option_gemm_mpairs;  // 0 (default) / 1: long-K BLOCK_N = 256 layers on pairs of m-blocks sharing every weight tile (measured: slower)
extern std::atomic<int> g_option_bneck_release;   // see opd_set_option
extern std::atomic<int> g_option_bneck_pair;   // cta_group::2 fused bottleneck tail (see opd_set_option)
extern std::atomic<int> g_option_dec0_const;   // see opd_set_option
extern std::atomic<int> g_option_gemm_reverse;   // see opd_set_option
extern std::atomic<int> g_option_gemm_pair;    // cta_group::2 GEMM variant (see opd_set_option)
extern std::atomic<int> g_option_gemm_res_wide;   // see opd_set_option
extern std::atomic<int> g_option_mlp_fused;   // see opd_set_option
extern std::atomic<int> g_option_mlp_pair;    // see opd_set_option
extern std::atomic<int> g_option_gemm_bres;    // 1 (default): short-K bottleneck outputs use the weight-stationary GEMM variant
extern std::atomic<int> g_option_stem_pool;    // 1 (default): max pooling fused into the stem kernel's epilogue
extern std::atomic<int> g_option_bneck_halo;   // 1 (default): stride-1 / 64-channel bottleneck tails use the halo-patch kernel

inline int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// cudaFuncSetAttribute and the occupancy / SM-count queries are per DEVICE: one-time setup and cached limits are keyed by
// the current device (handles of several devices may live in one process) and guarded by a mutex (handles are used from
// several threads).
constexpr int kMaxDevices = 64;
struct PerDeviceOnce {
  std::mutex mu;
  uint64_t done = 0;
};
template <typename F>
inline int once_per_device(PerDeviceOnce& o, F&& setup) {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) return setup();
  std::lock_guard<std::mutex> lk(o.mu);
  if ((o.done >> d) & 1ull) return 0;
  const int rc = setup();
  if (rc == 0) o.done |= 1ull << d;
  return rc;
}
// Makes `device` current for the lifetime of the guard and restores the caller's device afterwards (handles are bound to one
// device; the caller's current device is not ours to change).
struct DeviceGuard {
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
    else if (err == cudaSuccess) prev = -1;   // nothing to restore
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
struct PerDeviceInt {   // cached per-device integer, -1 = not yet computed
  std::mutex mu;
  int v[kMaxDevices];
  PerDeviceInt() { for (int& x : v) x = -1; }
};
template <typename F>
inline int cached_per_device(PerDeviceInt& c, int* out, F&& compute) {   // compute(int* value) -> status
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) return compute(out);
  std::lock_guard<std::mutex> lk(c.mu);
  if (c.v[d] < 0) {
    int val = -1;
    if (int rc = compute(&val)) return rc;
    c.v[d] = val;
  }
  *out = c.v[d];
  return 0;
}

}  // namespace opd

#define OPD_CUDA_OK(expr)                                                                          \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return opd::fail(OPD_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,                  \
                       cudaGetErrorString(_e));                                                    \
  } while (0)

#define OPD_REQUIRE(cond, ...)                                                                     \
  do {                                                                                             \
    if (!(cond)) return opd::fail(OPD_ERR_INVALID, __VA_ARGS__);                                   \
  } while (0)
