// Debug micro-benchmark of the per-SM pipes the softmax / epilogue phases of the attention kernels lean on:
// MUFU.EX2, F2FP (fp32 pair -> bf16x2), FFMA, FMNMX3, tcgen05.ld (32x32b.x32) and tcgen05.st (32x32b.x16).
// One CTA; `warps` warps run the same instruction stream (warp w sits on SMSP w % 4 and on TMEM lane quarter
// w % 4).  out[w] = clock64 cycles of warp w for `iters` iterations of an 8-instruction unrolled body.
// Not part of the product path: it sizes the kernels (profiles/README.md).
#include "attn_sm100.cuh"

namespace svae {
namespace sm100 {

using namespace ptx;

__global__ void __launch_bounds__(512, 1) pipe_bench_kernel(int mode, int iters, long long* out, float seed) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t trow = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + 0.01f * (float)(i + lane);
  uint32_t acc = 0;
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = (uint32_t)(lane + i);
  __syncthreads();
  const long long t0 = clock64();
  if (mode == 0) {                       // MUFU.EX2, 8 independent chains
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fast_exp2(a[i]);
  } else if (mode == 1) {                // F2FP.BF16.F32.PACK_AB
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t pk = Elem<__nv_bfloat16>::pack(a[i], a[(i + 1) & 7]);
        acc ^= pk;
        a[i] = __uint_as_float(pk | 0x3f000000u);
      }
  } else if (mode == 2) {                // FFMA
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], 1.0001f, 0.5f);
  } else if (mode == 3) {                // FMNMX3
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fmax3(a[i], a[(i + 1) & 7], a[(i + 2) & 7]);
  } else if (mode == 4) {                // tcgen05.ld 32x32b.x32 (4 KB per warp instruction), 4 in flight per wait
    uint32_t v1[32], v2[32], v3[32];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int rep = 0; rep < 2; ++rep) {
        const uint32_t base = trow + (uint32_t)(((it * 2 + rep) & 3) * 128);
        tmem_ld32(base, v);
        tmem_ld32(base + 32, v1);
        tmem_ld32(base + 64, v2);
        tmem_ld32(base + 96, v3);
        tmem_wait_ld(v, v1);
        tmem_dep(v2); tmem_dep(v3);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += (v[i] ^ v1[i]) + (v2[i] ^ v3[i]);
      }
    }
  } else if (mode == 5) {                // tcgen05.st 32x32b.x16 (2 KB per warp instruction)
    uint32_t s[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = v[i];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) tmem_st16(trow + 16 * i, s);
      tmem_wait_st();
    }
  } else if (mode == 6) {                // the softmax inner step: FFMA -> MUFU.EX2 -> FADD, pack every pair
    float l = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const float p0 = fast_exp2(fmaf(a[i], 0.125f, -1.0f)), p1 = fast_exp2(fmaf(a[i + 1], 0.125f, -1.0f));
        l += p0 + p1;
        acc ^= Elem<__nv_bfloat16>::pack(p0, p1);
        a[i] = p0; a[i + 1] = p1;
      }
    }
    a[0] += l;
  } else if (mode == 7) {                // the same step, 96 elements held in registers, fully unrolled (straight-line)
    float b[96];
#pragma unroll
    for (int i = 0; i < 96; ++i) b[i] = seed * (float)(i + 1) + 0.001f * (float)lane;
    float l0 = 0.f, l1 = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 96; i += 2) {
        const float p0 = fast_exp2(fmaf(b[i], 0.125f, -1.0f)), p1 = fast_exp2(fmaf(b[i + 1], 0.125f, -1.0f));
        l0 += p0; l1 += p1;
        acc ^= Elem<__nv_bfloat16>::pack(p0, p1);
        b[i] = p0; b[i + 1] = p1;
      }
    }
    a[0] += l0 + l1;
  }
  const long long t1 = clock64();
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += a[i];
  if (lane == 0) out[warp] = t1 - t0;
  if (sum == 1234.5678f || acc == 0x12345u || v[3] == 0xdeadbeefu) out[63] = 1;     // keep the work alive
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_slot);
}

}  // namespace sm100
}  // namespace svae

extern "C" __attribute__((visibility("default"))) int svae_debug_pipe_bench(int mode, int warps, int iters, long long* out,
                                                                            void* stream) {
  using namespace svae;
  SVAE_REQUIRE(warps >= 1 && warps <= 16 && mode >= 0 && mode <= 7, SVAE_ERR_INVALID, "svae_debug_pipe_bench: bad arguments");
  sm100::pipe_bench_kernel<<<1, warps * 32, 0, static_cast<cudaStream_t>(stream)>>>(mode, iters, out, 0.25f);
  SVAE_CUDA_CHECK(cudaGetLastError());
  return SVAE_OK;
}
