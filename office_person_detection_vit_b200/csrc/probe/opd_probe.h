/* C entry points of libopd_probe.so: MEASUREMENT PROBES used by benchmarks/*.py while the kernels were designed.
 * They are not part of the product library (libopd_b200.so, include/opd_b200.h) and nothing on the product path loads them. */
#ifndef OPD_PROBE_H
#define OPD_PROBE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* benchmarks/mma_probe.py: `iters` tcgen05.mma 128 x N x 16 issued by one thread per CTA, rotating over n_acc TMEM accumulators,
 * operands with 32-byte (swizzle32 = 1) or 128-byte swizzled rows; a_sbo / a_step != 0: A is a shifted view (8-row groups a_sbo
 * bytes apart, consecutive MMAs a_step bytes apart); commit_every > 0: two tcgen05.commit after every commit_every MMAs;
 * ld_iters > 0: four more warps run that many tcgen05.ld 32x32b.x32 (+ wait) against the MMAs, their ticks go to
 * cycles_dev[grid .. 2 grid); cycles_dev[0 .. grid) receives clock64 ticks from the first issue to the completion of the last. */
int opd_debug_mma_probe(int32_t N, int32_t swizzle32, int32_t n_acc, int32_t iters, int32_t walk, int32_t grid,
                        int32_t a_sbo, int32_t a_step, int32_t ld_iters, int32_t commit_every, uint64_t* cycles_dev,
                        void* stream);
/* benchmarks/halo_experiment.py: 3x3 / stride 1 / pad 1 convolution (C = N = 64) whose nine taps read ONE halo patch in shared
 * memory through shifted UMMA descriptors (mode 0) or nine im2col copies (mode 1). */
int opd_halo_conv3x3_test(const void* x_dev, int32_t B, int32_t H, int32_t W, const void* w_dev, const float* bias_dev,
                          void* y_dev, int32_t mode, void* stream);
/* benchmarks/pipe_probe.py: instruction-pipe throughput (mode 0 MUFU.EX2, 1 cvt.rn.bf16x2.f32, 2 fma.rn.f32x2, 3 the softmax mix);
 * cycles_dev[grid] receives each CTA's clock64 ticks for `iters` iterations of 8 (modes 0-2) or 4 pairs (mode 3) per thread. */
int opd_debug_pipe_probe(int32_t mode, int32_t iters, int32_t grid, int32_t threads, float* out_dev, uint64_t* cycles_dev, void* stream);
#ifdef __cplusplus
}
#endif
#endif
