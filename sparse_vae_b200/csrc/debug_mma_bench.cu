// Debug micro-benchmark: issue cost of back-to-back tcgen05.mma instructions from one elected thread
// (M = 128, K = 16, variable N; A from SMEM or TMEM; one or two issuing warps).  Not part of the product path;
// it exists to size the MMA instruction counts of the attention kernels (profiles/README.md).
#include "attn_sm100.cuh"
#include "../../include/sparse_vae_b200_debug.h"

namespace svae {
namespace sm100 {

using namespace ptx;

// Warp-convergent issue: every lane runs the loop with warp-uniform operands, one lane is elected inside the asm.
__device__ __forceinline__ void mma_ss_e(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts_e(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// variant bit 0: A from TMEM (TS) instead of SMEM (SS) ; bit 1: B MN-major ; bit 2: two issuing warps ;
// bit 3: warp-convergent issue (elect inside the asm) instead of an `if (lane == 0)` branch
__global__ void __launch_bounds__(96, 1) mma_issue_bench_kernel(int variant, int n, int count, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(&tmem_slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const bool ts = variant & 1, mn = variant & 2, two = variant & 4, conv = variant & 8, unrolled = variant & 16;
  if (unrolled && warp < (two ? 2 : 1)) {
    // fully unrolled groups of 16 MMAs whose descriptors differ by compile-time offsets (how a kernel with a fixed
    // schedule issues them), warp-convergent
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 32 * 1024);
    const uint32_t idesc = make_idesc(128, n, 1, 0, mn ? 1 : 0);
    const uint32_t d = tmem_base + warp * 256;
    const uint32_t ta = tmem_base + 128 + warp * 256;
    const long long t0 = clock64();
    for (int g = 0; g < count / 16; ++g) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint64_t bd = mn ? make_smem_desc(b_addr + (i & 7) * 2048, 4096, 1024, 128)
                               : make_smem_desc(b_addr + (i & 3) * 32, 16, 1024, 128);
        if (ts) mma_ts_e(d, ta + (i & 7) * 8, bd, idesc, (g | i) > 0);
        else mma_ss_e(d, make_smem_desc(a_addr + (i & 3) * 32, 16, 1024, 128), bd, idesc, (g | i) > 0);
      }
    }
    const long long t1 = clock64();
    if (elect_one()) tc_commit(&bars[warp]);
    __syncwarp();
    mbar_wait(&bars[warp], 0);
    const long long t2 = clock64();
    if (lane == 0) {
      out[warp * 2 + 0] = t1 - t0;
      out[warp * 2 + 1] = t2 - t0;
    }
  } else if (conv && warp < (two ? 2 : 1)) {
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 32 * 1024);
    const uint32_t idesc = make_idesc(128, n, 1, 0, mn ? 1 : 0);
    const uint32_t d = tmem_base + warp * 256;
    const long long t0 = clock64();
    for (int i = 0; i < count; ++i) {
      const uint32_t off = (i & 3) * 32;
      const uint64_t bd = mn ? make_smem_desc(b_addr + (i & 7) * 2048, 4096, 1024, 128)
                             : make_smem_desc(b_addr + off, 16, 1024, 128);
      if (ts) mma_ts_e(d, tmem_base + 128 + warp * 256 + (i & 7) * 8, bd, idesc, i > 0);
      else mma_ss_e(d, make_smem_desc(a_addr + off, 16, 1024, 128), bd, idesc, i > 0);
    }
    const long long t1 = clock64();
    if (elect_one()) tc_commit(&bars[warp]);
    __syncwarp();
    mbar_wait(&bars[warp], 0);
    const long long t2 = clock64();
    if (lane == 0) {
      out[warp * 2 + 0] = t1 - t0;
      out[warp * 2 + 1] = t2 - t0;
    }
  } else if (!conv && warp < (two ? 2 : 1) && lane == 0) {
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 32 * 1024);
    const uint32_t idesc = make_idesc(128, n, 1, 0, mn ? 1 : 0);
    const uint32_t d = tmem_base + warp * 256;
    const long long t0 = clock64();
    for (int i = 0; i < count; ++i) {
      const uint32_t off = (i & 3) * 32;
      const uint64_t bd = mn ? make_smem_desc(b_addr + (i & 7) * 2048, 4096, 1024, 128)
                             : make_smem_desc(b_addr + off, 16, 1024, 128);
      if (ts) mma_ts(d, tmem_base + 128 + warp * 256 + (i & 7) * 8, bd, idesc, i > 0);
      else mma_ss(d, make_smem_desc(a_addr + off, 16, 1024, 128), bd, idesc, i > 0);
    }
    const long long t1 = clock64();
    tc_commit(&bars[warp]);
    mbar_wait(&bars[warp], 0);
    const long long t2 = clock64();
    out[warp * 2 + 0] = t1 - t0;     // issue time
    out[warp * 2 + 1] = t2 - t0;     // issue + drain
  }
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

}  // namespace sm100
}  // namespace svae

extern "C" __attribute__((visibility("default"))) int svae_debug_mma_bench(int variant, int n, int count, long long* out,
                                                                           void* stream) {
  using namespace svae;
  auto kern = sm100::mma_issue_bench_kernel;
  SVAE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  kern<<<1, 96, 64 * 1024, static_cast<cudaStream_t>(stream)>>>(variant, n, count, out);
  SVAE_CUDA_CHECK(cudaGetLastError());
  return SVAE_OK;
}
