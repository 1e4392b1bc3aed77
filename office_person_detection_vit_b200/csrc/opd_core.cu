// Library-wide state of libopd_b200.so (error string, launch counter, version).
#include "opd_common.h"

namespace opd {
std::string& last_error_ref() {
  static thread_local std::string s;
  return s;
}
std::atomic<int64_t> g_launches{0};
std::atomic<int> g_option_bneck_halo{1};
std::atomic<int> g_option_attention_tc{1};
std::atomic<int> g_option_attention_kv{96};
std::atomic<int> g_option_probe{0};
std::atomic<int> g_option_stem_pool{1};
std::atomic<int> g_option_gemm_bres{1};
std::atomic<int> g_option_gemm_cluster{0};
std::atomic<int> g_option_gemm_outbufs{1};
std::atomic<int> g_option_pdl{1};
std::atomic<int> g_option_gemm_pair{1};
std::atomic<int> g_option_gemm_reverse{1};
std::atomic<int> g_option_dec0_const{1};
std::atomic<int> g_option_bneck_pair{1};
std::atomic<int> g_option_bneck_release{3};
std::atomic<int> g_option_gemm_mpairs{0};
std::atomic<int> g_option_gemm_res_wide{1};
std::atomic<int> g_option_mlp_fused{1};
std::atomic<int> g_option_mlp_pair{1};
}  // namespace opd

extern "C" {
int opd_version(void) { return OPD_ABI_VERSION; }
const char* opd_last_error(void) { return opd::last_error_ref().c_str(); }
int64_t opd_launch_count(void) { return opd::g_launches.load(); }
int opd_set_option(const char* name, int32_t value) {
  if (name && std::string(name) == "bneck_halo") {
    opd::g_option_bneck_halo.store(value);
    return OPD_OK;
  }
  if (name && std::string(name) == "attention_kv") {
    opd::g_option_attention_kv.store(value);
    return OPD_OK;
  }
  if (name && std::string(name) == "probe") {   // measurement probes (benchmarks/step_times.py): results are WRONG when set
    opd::g_option_probe.store(value);
    return OPD_OK;
  }
  if (name && std::string(name) == "gemm_bres") {   // 0: never use the weight-stationary GEMM variant (A/B runs; new plans only)
    opd::g_option_gemm_bres.store(value);
    return OPD_OK;
  }
  if (name && std::string(name) == "gemm_cluster") {   // 0 (default): no clusters; 1: 2-CTA clusters for the big BLOCK_N = 256 layers; 2: whenever BLOCK_N = 256 (tests); new plans only
    opd::g_option_gemm_cluster.store(value);
    return OPD_OK;
  }
  if (name && std::string(name) == "gemm_outbufs") {   // 0: one staging box per epilogue warpgroup everywhere (A/B runs; new plans only)
    opd::g_option_gemm_outbufs.store(value);
    return OPD_OK;
  }
  if (name && std::string(name) == "pdl") {   // 1 (default): GEMM and attention launches allow programmatic dependent launch (their prologues overlap the previous kernel's tail: 18.71 -> 18.63 ms at batch 64, 1.72 -> 1.65 ms at batch 1); 0: plain stream order
    opd::g_option_pdl.store(value);
    return OPD_OK;
  }
  if (name && std::string(name) == "gemm_mpairs") {   // 0 (default): no m-block pairs; 1: long-K BLOCK_N = 256 layers; 2: whenever BLOCK_N = 256 (tests); new plans only
    opd::g_option_gemm_mpairs.store(value);
    return OPD_OK;
  }
  if (name && std::string(name) == "bneck_release") {   // fused bottleneck tails: residual / output slots are released early in the next epilogue step (bit 0: im2col kernel, right after the accumulator wait: 23 % faster; bit 1: halo kernel, after the step's arithmetic: 3 % faster in the pipeline; default 3) or at its end
    opd::g_option_bneck_release.store(value);
    return OPD_OK;
  }
  if (name && std::string(name) == "bneck_pair") {   // cta_group::2 fused bottleneck tail (MID = 128): 0 off, 1 (default) layers of at least one tile per SM, 3 whenever the tile count is even (tests); new plans only
    opd::g_option_bneck_pair.store(value);
    return OPD_OK;
  }
  if (name && std::string(name) == "dec0_const") {   // 1 (default): decoder layer 0's frame-independent self-attention block runs once per plan; 0: in every step; new plans only
    opd::g_option_dec0_const.store(value);
    return OPD_OK;
  }
  if (name && std::string(name) == "gemm_reverse") {   // engine: 1 = a GEMM layer walks its m-blocks in the direction opposite to the launch before it (L2 reuse); 0 = always ascending; new plans only
    opd::g_option_gemm_reverse.store(value);
    return OPD_OK;
  }
  if (name && std::string(name) == "gemm_pair") {   // cta_group::2 GEMM: 0 off, 1 (default) BLOCK_N = 256 layers with a tile pair per cluster, 3 whenever BLOCK_N = 256 (tests); new plans only
    opd::g_option_gemm_pair.store(value);
    return OPD_OK;
  }
  if (name && std::string(name) == "gemm_res_wide") {   // bias + residual + ReLU GEMMs: 0 = 128-column tiles; 1 (default) = layers with K >= 256 keep 256-column cta_group::2 tiles (deeper operand ring, three residual slots); 2 = whenever N allows it (tests); new plans only
    opd::g_option_gemm_res_wide.store(value);
    return OPD_OK;
  }
  if (name && std::string(name) == "mlp_fused") {   // 1 (default): fc1 + ReLU + fc2 + residual + LayerNorm of a layer in one kernel (tc_mlp.cu) when it has at least 16 row tiles; 2: always; 0: two GEMM launches; new plans only
    opd::g_option_mlp_fused.store(value);
    return OPD_OK;
  }
  if (name && std::string(name) == "mlp_pair") {   // 1 (default): the fused feed-forward kernel runs as cta_group::2 pairs; 0: one CTA per row tile; new plans only
    opd::g_option_mlp_pair.store(value);
    return OPD_OK;
  }
  if (name && std::string(name) == "stem_pool") {   // 0: stem and max pooling as two kernels; 1: fused (default); 2: fused in debug plans too
    opd::g_option_stem_pool.store(value);
    return OPD_OK;
  }
  if (name && std::string(name) == "attention_tc") {
    opd::g_option_attention_tc.store(value);
    return OPD_OK;
  }
  return opd::fail(OPD_ERR_INVALID, "opd_set_option: unknown option '%s'", name ? name : "(null)");
}
}
