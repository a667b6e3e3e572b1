// Fused multi-tensor optimizer step for the data-parallel training path: global gradient-norm clipping
// (reference core/language_model.py:120-122, torch.nn.utils.clip_grad_norm_) and the RAdam update
// (reference core/rectified_adam.py:15-88) over ALL parameter tensors in a handful of launches instead of a
// Python loop with ~10 element-wise launches per parameter.
//
// HBM-bound element-wise work: sum of squares reads 4 B / element, the in-place scale 8 B, the RAdam update
// 28 B (read p, g, m, v; write p, m, v).  Tensors are described by HOST arrays of device pointers; the library
// packs them into kernel-parameter tables (no device-side table, no allocation, no synchronisation).
#include <math.h>

#include "common.cuh"

namespace svae {

constexpr int kMtChunk = 65536;        // elements per CTA
constexpr int kMtThreads = 512;
constexpr int kMtMaxTensors = 120;      // (tables of ~13 KB: kernel parameters may be up to 32 KB on sm_70+ with CUDA >= 12.1)
constexpr int kMtMaxBlocks = 1600;

template <int NPTR>
struct MtTable {
  void* ptr[NPTR][kMtMaxTensors];
  int64_t numel[kMtMaxTensors];
  int block_chunk[kMtMaxBlocks];
  unsigned char block_tensor[kMtMaxBlocks];
};

struct RadamArgs {
  float beta1, beta2, omb1, omb2, eps, decay, step_size, bias_v;   // omb = 1 - beta (rounded from double); decay = 1 - lr*wd ; step_size = lr / bias_m
  int rectified;
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ void radam_one(float& p, float g, float& m, float& v, const RadamArgs& a) {
  // exp_avg.mul_(beta1).add_(grad, alpha=1-beta1) ; exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2)
  m = fmaf(g, a.omb1, m * a.beta1);
  v = fmaf(g * g, a.omb2, v * a.beta2);
  p *= a.decay;                                            // param.mul_(1 - lr * weight_decay)
  if (a.rectified) {
    const float denom = sqrtf(v) / a.bias_v + a.eps;       // (exp_avg_sq.sqrt() / bias_correction_v).add_(eps)
    p -= a.step_size * (m / denom);                        // addcdiv_(exp_avg, denom, value=-step_size)
  } else {
    p -= a.step_size * m;                                  // SGD with momentum while the variance is intractable
  }
}

__global__ void __launch_bounds__(kMtThreads) radam_kernel(const __grid_constant__ MtTable<4> t, const RadamArgs a_host,
                                                             const RadamArgs* __restrict__ a_dev) {
  const RadamArgs a = a_dev ? *a_dev : a_host;        // graph replay: the step's scalars are read from device memory
  const int ti = t.block_tensor[blockIdx.x];
  const int64_t base = (int64_t)t.block_chunk[blockIdx.x] * kMtChunk;
  const int64_t n = t.numel[ti];
  float* p = reinterpret_cast<float*>(t.ptr[0][ti]) + base;
  const float* g = reinterpret_cast<const float*>(t.ptr[1][ti]) + base;
  float* m = reinterpret_cast<float*>(t.ptr[2][ti]) + base;
  float* v = reinterpret_cast<float*>(t.ptr[3][ti]) + base;
  const int cnt = (int)((n - base) < kMtChunk ? (n - base) : kMtChunk);
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (vec) {
    const int n4 = cnt >> 2;
    for (int i = threadIdx.x; i < n4; i += kMtThreads) {
      float4 pp = ld4(p + 4 * i), gg = ld4(g + 4 * i), mm = ld4(m + 4 * i), vv = ld4(v + 4 * i);
      radam_one(pp.x, gg.x, mm.x, vv.x, a);
      radam_one(pp.y, gg.y, mm.y, vv.y, a);
      radam_one(pp.z, gg.z, mm.z, vv.z, a);
      radam_one(pp.w, gg.w, mm.w, vv.w, a);
      st4(p + 4 * i, pp); st4(m + 4 * i, mm); st4(v + 4 * i, vv);
    }
    for (int i = (n4 << 2) + threadIdx.x; i < cnt; i += kMtThreads) radam_one(p[i], g[i], m[i], v[i], a);
  } else {
    for (int i = threadIdx.x; i < cnt; i += kMtThreads) radam_one(p[i], g[i], m[i], v[i], a);
  }
}

// partial[first_partial + blockIdx.x] = sum of squares of the block's chunk (fixed summation order -> deterministic)
__global__ void __launch_bounds__(kMtThreads) sumsq_kernel(const __grid_constant__ MtTable<1> t, float* partial, int first_partial) {
  const int ti = t.block_tensor[blockIdx.x];
  const int64_t base = (int64_t)t.block_chunk[blockIdx.x] * kMtChunk;
  const int64_t n = t.numel[ti];
  const float* g = reinterpret_cast<const float*>(t.ptr[0][ti]) + base;
  const int cnt = (int)((n - base) < kMtChunk ? (n - base) : kMtChunk);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const int n4 = cnt >> 2;
    for (int i = threadIdx.x; i < n4; i += kMtThreads) {
      const float4 x = ld4(g + 4 * i);
      s0 = fmaf(x.x, x.x, s0); s1 = fmaf(x.y, x.y, s1); s2 = fmaf(x.z, x.z, s2); s3 = fmaf(x.w, x.w, s3);
    }
    for (int i = (n4 << 2) + threadIdx.x; i < cnt; i += kMtThreads) s0 = fmaf(g[i], g[i], s0);
  } else {
    for (int i = threadIdx.x; i < cnt; i += kMtThreads) s0 = fmaf(g[i], g[i], s0);
  }
  float s = (s0 + s1) + (s2 + s3);
  __shared__ float red[kMtThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float r = threadIdx.x < kMtThreads / 32 ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (threadIdx.x == 0) partial[first_partial + blockIdx.x] = r;
  }
}

// norm_coef[0] = sqrt(sum partial) ; norm_coef[1] = min(1, max_norm / (norm + 1e-6))   (clip_grad_norm_)
__global__ void __launch_bounds__(1024) norm_finish_kernel(const float* partial, int n, float max_norm, float* norm_coef) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) s += (double)partial[i];
  __shared__ double red[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double r = red[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (threadIdx.x == 0) {
      const float norm = (float)sqrt(r);
      norm_coef[0] = norm;
      norm_coef[1] = fminf(max_norm / (norm + 1e-6f), 1.0f);
    }
  }
}

__global__ void __launch_bounds__(kMtThreads) scale_kernel(const __grid_constant__ MtTable<1> t, const float* coef_ptr) {
  const float coef = *coef_ptr;
  const int ti = t.block_tensor[blockIdx.x];
  const int64_t base = (int64_t)t.block_chunk[blockIdx.x] * kMtChunk;
  const int64_t n = t.numel[ti];
  float* g = reinterpret_cast<float*>(t.ptr[0][ti]) + base;
  const int cnt = (int)((n - base) < kMtChunk ? (n - base) : kMtChunk);
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const int n4 = cnt >> 2;
    for (int i = threadIdx.x; i < n4; i += kMtThreads) {
      float4 x = ld4(g + 4 * i);
      x.x *= coef; x.y *= coef; x.z *= coef; x.w *= coef;
      st4(g + 4 * i, x);
    }
    for (int i = (n4 << 2) + threadIdx.x; i < cnt; i += kMtThreads) g[i] *= coef;
  } else {
    for (int i = threadIdx.x; i < cnt; i += kMtThreads) g[i] *= coef;
  }
}

// dst[i] = src[i] * scale over every tensor of the list (gathers per-parameter gradients into a flat all-reduce bucket)
__global__ void __launch_bounds__(kMtThreads) scale_copy_kernel(const __grid_constant__ MtTable<2> t, float scale) {
  const int ti = t.block_tensor[blockIdx.x];
  const int64_t base = (int64_t)t.block_chunk[blockIdx.x] * kMtChunk;
  const int64_t n = t.numel[ti];
  float* d = reinterpret_cast<float*>(t.ptr[0][ti]) + base;
  const float* s = reinterpret_cast<const float*>(t.ptr[1][ti]) + base;
  const int cnt = (int)((n - base) < kMtChunk ? (n - base) : kMtChunk);
  if (((reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(s)) & 15) == 0) {
    const int n4 = cnt >> 2;
    for (int i = threadIdx.x; i < n4; i += kMtThreads) {
      float4 x = ld4(s + 4 * i);
      x.x *= scale; x.y *= scale; x.z *= scale; x.w *= scale;
      st4(d + 4 * i, x);
    }
    for (int i = (n4 << 2) + threadIdx.x; i < cnt; i += kMtThreads) d[i] = s[i] * scale;
  } else {
    for (int i = threadIdx.x; i < cnt; i += kMtThreads) d[i] = s[i] * scale;
  }
}

// dst[i] = T(src[i]) over every tensor of the list: the autocast-dtype copies of all projection weights in a few launches
// (core/weight_shadows.py) instead of one ATen cast launch per weight and per step.  Same rounding as `.to(dtype)`.
template <typename T>
__global__ void __launch_bounds__(kMtThreads) cast_copy_kernel(const __grid_constant__ MtTable<2> t) {
  const int ti = t.block_tensor[blockIdx.x];
  const int64_t base = (int64_t)t.block_chunk[blockIdx.x] * kMtChunk;
  const int64_t n = t.numel[ti];
  T* d = reinterpret_cast<T*>(t.ptr[0][ti]) + base;
  const float* s = reinterpret_cast<const float*>(t.ptr[1][ti]) + base;
  const int cnt = (int)((n - base) < kMtChunk ? (n - base) : kMtChunk);
  if ((reinterpret_cast<uintptr_t>(s) & 15) == 0 && (reinterpret_cast<uintptr_t>(d) & 7) == 0) {
    const int n4 = cnt >> 2;
    for (int i = threadIdx.x; i < n4; i += kMtThreads) {
      const float4 x = ld4(s + 4 * i);
      struct alignas(8) Out4 { T v[4]; } o;
      o.v[0] = from_f32<T>(x.x); o.v[1] = from_f32<T>(x.y); o.v[2] = from_f32<T>(x.z); o.v[3] = from_f32<T>(x.w);
      *reinterpret_cast<Out4*>(d + 4 * i) = o;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < cnt; i += kMtThreads) d[i] = from_f32<T>(s[i]);
  } else {
    for (int i = threadIdx.x; i < cnt; i += kMtThreads) d[i] = from_f32<T>(s[i]);
  }
}

// Walks the tensor list, filling launch tables; calls launch(table, blocks, first_block_index) whenever one is full.
template <int NPTR, typename Launch>
static int for_each_table(int n, void* const* const* lists, const int64_t* numel, Launch&& launch) {
  MtTable<NPTR> t;
  int nt = 0, nb = 0, done_blocks = 0;
  auto flush = [&]() -> int {
    if (nb == 0) { nt = 0; return SVAE_OK; }
    int rc = launch(t, nb, done_blocks);
    done_blocks += nb;
    nt = 0; nb = 0;
    return rc;
  };
  for (int i = 0; i < n; ++i) {
    if (numel[i] <= 0) continue;
    const int chunks = (int)((numel[i] + kMtChunk - 1) / kMtChunk);
    int c = 0;
    while (c < chunks) {
      if (nt == kMtMaxTensors || nb == kMtMaxBlocks) { int rc = flush(); if (rc) return rc; }
      for (int k = 0; k < NPTR; ++k) t.ptr[k][nt] = lists[k][i];
      t.numel[nt] = numel[i];
      while (c < chunks && nb < kMtMaxBlocks) {
        t.block_tensor[nb] = (unsigned char)nt;
        t.block_chunk[nb] = c;
        ++nb; ++c;
      }
      ++nt;
    }
  }
  return flush();
}

}  // namespace svae

using namespace svae;

extern "C" int64_t svae_multi_tensor_chunks(int32_t n, const int64_t* numel) {
  int64_t total = 0;
  for (int i = 0; i < n; ++i)
    if (numel[i] > 0) total += (numel[i] + kMtChunk - 1) / kMtChunk;
  return total;
}

extern "C" int svae_multi_tensor_scale_copy(int32_t n, void* const* dst, void* const* src, const int64_t* numel, float scale,
                                           void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SVAE_REQUIRE(n >= 0 && dst && src && numel, SVAE_ERR_INVALID, "svae_multi_tensor_scale_copy: null argument");
  void* const* lists[2] = {dst, src};
  ScopedKernelTimer timer("grad_gather", st);
  return for_each_table<2>(n, lists, numel, [&](const MtTable<2>& t, int nb, int) -> int {
    scale_copy_kernel<<<nb, kMtThreads, 0, st>>>(t, scale);
    SVAE_CUDA_CHECK(cudaGetLastError());
    return SVAE_OK;
  });
}

extern "C" int svae_multi_tensor_cast(int32_t n, void* const* dst, void* const* src, const int64_t* numel, int32_t dst_dtype,
                                     void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SVAE_REQUIRE(n >= 0 && dst && src && numel, SVAE_ERR_INVALID, "svae_multi_tensor_cast: null argument");
  SVAE_REQUIRE(dst_dtype == SVAE_DTYPE_BF16 || dst_dtype == SVAE_DTYPE_F16, SVAE_ERR_UNSUPPORTED,
               "svae_multi_tensor_cast: destination dtype %d (bf16 / f16 only)", dst_dtype);
  void* const* lists[2] = {dst, src};
  ScopedKernelTimer timer("weight_cast", st);
  return for_each_table<2>(n, lists, numel, [&](const MtTable<2>& t, int nb, int) -> int {
    if (dst_dtype == SVAE_DTYPE_BF16) cast_copy_kernel<__nv_bfloat16><<<nb, kMtThreads, 0, st>>>(t);
    else cast_copy_kernel<__half><<<nb, kMtThreads, 0, st>>>(t);
    SVAE_CUDA_CHECK(cudaGetLastError());
    return SVAE_OK;
  });
}

extern "C" int svae_clip_grad_norm(int32_t n, void* const* grads, const int64_t* numel, float max_norm, float* partials,
                                   int64_t partials_len, float* norm_coef, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SVAE_REQUIRE(n >= 0 && grads && numel && partials && norm_coef, SVAE_ERR_INVALID, "svae_clip_grad_norm: null argument");
  const int64_t chunks = svae_multi_tensor_chunks(n, numel);
  SVAE_REQUIRE(chunks <= partials_len, SVAE_ERR_INVALID, "svae_clip_grad_norm: workspace of %lld floats < %lld chunks",
               (long long)partials_len, (long long)chunks);
  void* const* lists[1] = {grads};
  int rc;
  {
    ScopedKernelTimer timer("grad_sumsq", st);
    rc = for_each_table<1>(n, lists, numel, [&](const MtTable<1>& t, int nb, int first) -> int {
      sumsq_kernel<<<nb, kMtThreads, 0, st>>>(t, partials, first);
      SVAE_CUDA_CHECK(cudaGetLastError());
      return SVAE_OK;
    });
    if (rc) return rc;
    norm_finish_kernel<<<1, 1024, 0, st>>>(partials, (int)chunks, max_norm, norm_coef);
    SVAE_CUDA_CHECK(cudaGetLastError());
  }
  {
    ScopedKernelTimer timer("grad_scale", st);
    rc = for_each_table<1>(n, lists, numel, [&](const MtTable<1>& t, int nb, int) -> int {
      scale_kernel<<<nb, kMtThreads, 0, st>>>(t, norm_coef + 1);
      SVAE_CUDA_CHECK(cudaGetLastError());
      return SVAE_OK;
    });
  }
  return rc;
}

// scalar schedule in double precision, exactly the reference's Python arithmetic (core/rectified_adam.py:24-36,72)
static RadamArgs radam_args(double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step) {
  const double b1 = beta1, b2 = beta2;
  const double beta2_t = pow(b2, (double)step);
  const double bias_v = sqrt(1.0 - beta2_t);
  const double bias_m = 1.0 - pow(b1, (double)step);
  const double rho_inf = 2.0 / (1.0 - b2) - 1.0;
  const double rho_t = rho_inf - 2.0 * (double)step * beta2_t / (1.0 - beta2_t);
  double lr_eff = lr;
  RadamArgs a;
  a.rectified = rho_t > 4.0 ? 1 : 0;
  if (a.rectified) {
    const double r_t = sqrt(((rho_t - 4.0) * (rho_t - 2.0) * rho_inf) / ((rho_inf - 4.0) * (rho_inf - 2.0) * rho_t));
    lr_eff *= r_t * bias_v;
  }
  a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.eps = (float)eps;
  a.omb1 = (float)(1.0 - b1); a.omb2 = (float)(1.0 - b2);
  a.decay = (float)(1.0 - lr_eff * weight_decay);
  a.step_size = (float)(lr_eff / bias_m);
  a.bias_v = (float)bias_v;
  return a;
}

extern "C" int32_t svae_radam_args_bytes(void) { return (int32_t)sizeof(RadamArgs); }

extern "C" int svae_radam_args(double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step,
                               void* args_out) {
  SVAE_REQUIRE(args_out && step >= 1, SVAE_ERR_INVALID, "svae_radam_args: null output or step < 1");
  *reinterpret_cast<RadamArgs*>(args_out) = radam_args(lr, beta1, beta2, eps, weight_decay, step);
  return SVAE_OK;
}

static int radam_launch(int32_t n, void* const* params, void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                        const int64_t* numel, const RadamArgs& a, const RadamArgs* a_dev, cudaStream_t st) {
  void* const* lists[4] = {params, grads, exp_avg, exp_avg_sq};
  ScopedKernelTimer timer("radam_step", st);
  return for_each_table<4>(n, lists, numel, [&](const MtTable<4>& t, int nb, int) -> int {
    radam_kernel<<<nb, kMtThreads, 0, st>>>(t, a, a_dev);
    SVAE_CUDA_CHECK(cudaGetLastError());
    return SVAE_OK;
  });
}

extern "C" int svae_radam_step(int32_t n, void* const* params, void* const* grads, void* const* exp_avg,
                               void* const* exp_avg_sq, const int64_t* numel, double lr, double beta1, double beta2, double eps,
                               double weight_decay, int64_t step, void* stream) {
  SVAE_REQUIRE(n >= 0 && params && grads && exp_avg && exp_avg_sq && numel, SVAE_ERR_INVALID, "svae_radam_step: null argument");
  SVAE_REQUIRE(step >= 1, SVAE_ERR_INVALID, "svae_radam_step: step is 1-indexed");
  return radam_launch(n, params, grads, exp_avg, exp_avg_sq, numel, radam_args(lr, beta1, beta2, eps, weight_decay, step), nullptr,
                      reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int svae_radam_step_g(int32_t n, void* const* params, void* const* grads, void* const* exp_avg,
                                 void* const* exp_avg_sq, const int64_t* numel, const void* args_dev, void* stream) {
  SVAE_REQUIRE(n >= 0 && params && grads && exp_avg && exp_avg_sq && numel && args_dev, SVAE_ERR_INVALID,
               "svae_radam_step_g: null argument");
  RadamArgs unused{};
  return radam_launch(n, params, grads, exp_avg, exp_avg_sq, numel, unused, reinterpret_cast<const RadamArgs*>(args_dev),
                      reinterpret_cast<cudaStream_t>(stream));
}
