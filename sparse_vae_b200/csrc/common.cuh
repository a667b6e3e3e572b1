// Shared host/device helpers for libsvae_b200: error plumbing, dtype conversion, block-band geometry.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/sparse_vae_b200.h"

namespace svae {

// ---- error plumbing --------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define SVAE_CUDA_CHECK(expr)                                         \
  do {                                                                \
    cudaError_t _e = (expr);                                          \
    if (_e != cudaSuccess) return ::svae::cuda_fail(_e, #expr);       \
  } while (0)

#define SVAE_REQUIRE(cond, code, ...)                                 \
  do {                                                                \
    if (!(cond)) {                                                    \
      ::svae::set_error(__VA_ARGS__);                                 \
      return (code);                                                  \
    }                                                                 \
  } while (0)

// ---- opt-in to > 48 KB of dynamic shared memory: a per-DEVICE function attribute, set once per device and kernel.
// Expands to a block with its own static mask, so every call site (= kernel instantiation) keeps its own record.
#define SVAE_CONFIGURE_SMEM(kernel, bytes)                                                            \
  do {                                                                                                \
    static std::atomic<unsigned long long> _done[2] = {{0ull}, {0ull}};                              \
    int _dev = 0;                                                                                     \
    SVAE_CUDA_CHECK(cudaGetDevice(&_dev));                                                            \
    const unsigned long long _bit = 1ull << (_dev & 63);                                              \
    if (!(_done[(_dev >> 6) & 1].load(std::memory_order_acquire) & _bit)) {                           \
      SVAE_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)); \
      _done[(_dev >> 6) & 1].fetch_or(_bit, std::memory_order_release);                               \
    }                                                                                                 \
  } while (0)

// number of SMs of the current device (cached per device)
int sm_count_of_current_device();

// ---- optional per-kernel device timing (svae_profile_begin / svae_profile_end) ---------------------
// When enabled, every kernel launch of the library is bracketed by two CUDA events on its own stream.
struct ScopedKernelTimer {
  ScopedKernelTimer(const char* name, cudaStream_t st);
  ~ScopedKernelTimer();
  int slot;
  cudaStream_t st;
};

// ---- geometry of the banded + global-column layout -------------------------------------------
// SparseAttention.get_master_layout (reference core/sparse_attention.py:38-59) only ever produces
// "band + optional column 0": block-row r attends key blocks [r - (left-1), r + nsup] and block 0.
struct Band {
  int left;   // sub-diagonals including the main one  (= left_context)
  int nsup;   // super-diagonals                        (= max(right_context - 1, 0))
  int cls;    // column 0 always attended
};

__host__ __device__ inline Band make_band(int window, int causal, int include_cls) {
  int sides = causal ? 1 : 2;
  int left = window / sides + window % sides;
  int right = window - left;
  Band b;
  b.left = left;
  b.nsup = right > 1 ? right - 1 : 0;
  b.cls = include_cls ? 1 : 0;
  return b;
}

__host__ __device__ inline bool block_live(const Band& g, int r, int c) {
  if (g.cls && c == 0) return true;
  return c >= r - (g.left - 1) && c <= r + g.nsup;
}

// ---- dtype helpers -----------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T x);
template <> __device__ __forceinline__ float to_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }
template <> __device__ __forceinline__ float to_f32<__half>(__half x) { return __half2float(x); }

template <typename T> __device__ __forceinline__ T from_f32(float x);
template <> __device__ __forceinline__ float from_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }
template <> __device__ __forceinline__ __half from_f32<__half>(float x) { return __float2half_rn(x); }

// ---- packed fp32 arithmetic of sm_100 (FFMA2 / FMUL2 / FADD2: two lanes per issued instruction) -----------------
#if defined(__CUDACC__)
__device__ __forceinline__ uint64_t pk2(float2 v) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y));
  return r;
}
__device__ __forceinline__ float2 up2(uint64_t r) {
  float2 v;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(r));
  return v;
}
// packed fp32 FMA / MUL of sm_100 (two lanes per instruction)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(b)), "l"(pk2(c)));
  return up2(d);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(b)));
  return up2(d);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk2(a)), "l"(pk2(b)));
  return up2(d);
}
__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }
#endif

inline size_t dtype_size(int dtype) { return dtype == SVAE_DTYPE_F32 ? 4 : 2; }

}  // namespace svae
