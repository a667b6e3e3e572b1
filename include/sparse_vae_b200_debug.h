/*
 * Micro-benchmarks of sm_100a primitives (tcgen05.mma issue cost, per-pipe instruction rates) used to size the
 * attention kernels (profiles/README.md).  NOT part of the product library: they live in libsvae_b200_dbg.so, built
 * by `python sparse_vae_b200/csrc/build.py --debug` and loaded only by tests/mma_bench.py / tests/pipe_bench.py.
 */
#ifndef SPARSE_VAE_B200_DEBUG_H_
#define SPARSE_VAE_B200_DEBUG_H_

#include "sparse_vae_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Debug micro-benchmark (one CTA): clock64 cycles to issue `count` back-to-back tcgen05.mma (M=128, K=16, N=n).
 * variant bit 0: A from TMEM, bit 1: B MN-major, bit 2: two issuing warps.  out: int64[4] =
 * {issue, issue+drain} per issuing warp. */
SVAE_API int svae_debug_mma_bench(int variant, int n, int count, long long* out, void* stream);
/* Debug micro-benchmark (one CTA, `warps` warps): cycles per warp for `iters` x 8 back-to-back instructions of
 * mode 0 MUFU.EX2, 1 F2FP bf16x2 pack, 2 FFMA, 3 FMNMX3, 4 tcgen05.ld 32x32b.x32, 5 tcgen05.st 32x32b.x16,
 * 6 the softmax step (FFMA, EX2, FADD, pack).  out: int64[64]. */
SVAE_API int svae_debug_pipe_bench(int mode, int warps, int iters, long long* out, void* stream);

/* libsvae_b200_dbg.so also holds the product kernels compiled with -DSVAE_DEBUG_BUILD (same entry points as
 * sparse_vae_b200.h).  When non-NULL, the next one-pass svae_attn_bwd launches write, per CTA and warp, the cycles spent
 * in every kind of wait: int64 [num_ctas][16 warps][12] = {full, stat, s_ready, p_ready, u_free, group, ds_free,
 * acc_ready, acc_free, free, total cycles of the CTA, tiles of the CTA}.  Process-global; debug library only. */
SVAE_API void svae_debug_set_b1_timeline(long long* timeline);

#ifdef __cplusplus
}
#endif
#endif
