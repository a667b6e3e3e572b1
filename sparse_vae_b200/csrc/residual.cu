// Residual-stream update of the transformer blocks under autocast (reference core/transformer_layer.py:41,49,61:
// `x = x + h`): x is the fp32 residual stream, h the 16-bit output of the attention / feed-forward branch.  ATen
// runs this mixed-dtype add through its non-vectorised `unrolled_elementwise_kernel` (112 us at [65536, 512] on a
// B200, 3.0 TB/s); this is the same fp32 addition (promote h, one rounding) with 16-byte accesses.
// HBM-bound: numel * (4 + sizeof(T) + 4) bytes.
#include <curand_kernel.h>
#include <curand_philox4x32_x.h>

#include "common.cuh"

namespace svae {

template <typename T> __device__ __forceinline__ void unpack8(const uint4& t, float (&f)[8]);
template <> __device__ __forceinline__ void unpack8<__nv_bfloat16>(const uint4& t, float (&f)[8]) {
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 p = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
    f[2 * i] = p.x; f[2 * i + 1] = p.y;
  }
}
template <> __device__ __forceinline__ void unpack8<__half>(const uint4& t, float (&f)[8]) {
  const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 p = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    f[2 * i] = p.x; f[2 * i + 1] = p.y;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) residual_add_kernel(const float* __restrict__ x, const T* __restrict__ h,
                                                            float* __restrict__ out, int64_t vecs) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < vecs; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = *reinterpret_cast<const float4*>(x + i * 8), b = *reinterpret_cast<const float4*>(x + i * 8 + 4);
    float f[8];
    unpack8<T>(*reinterpret_cast<const uint4*>(h + i * 8), f);
    *reinterpret_cast<float4*>(out + i * 8) = make_float4(a.x + f[0], a.y + f[1], a.z + f[2], a.w + f[3]);
    *reinterpret_cast<float4*>(out + i * 8 + 4) = make_float4(b.x + f[4], b.y + f[5], b.z + f[6], b.w + f[7]);
  }
}

// ---- out = x + dropout(h) (reference core/transformer_layer.py:61 `x + self.dropout(self.ffn(...))`, training mode) ----
// The keep mask is never stored: element 8 i + e keeps its value iff word (e & 3) of Philox4x32-10(counter = (i, call
// e >> 2, offset), key = seed) is >= p * 2^32, and the backward kernel regenerates it from the same (seed, offset).
// Arithmetic like ATen's fused dropout followed by the promoted add: bf16(h * 1/(1-p)) for kept elements, then fp32
// x + that; backward: bf16(bf16(g) * 1/(1-p)) for kept elements, 0 otherwise.  (Same distribution as nn.Dropout, its
// own use of the generator's Philox stream: the host advances the generator offset by 4 per launch.)
__device__ __forceinline__ void keep_mask8(int64_t i, uint64_t seed, uint64_t offset, unsigned threshold, bool (&keep)[8]) {
  const uint2 key = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
#pragma unroll
  for (int call = 0; call < 2; ++call) {
    const uint4 ctr = make_uint4((unsigned)i, (unsigned)((uint64_t)i >> 32), (unsigned)offset + call, (unsigned)(offset >> 32));
    const uint4 r = curand_Philox4x32_10(ctr, key);
    keep[4 * call + 0] = r.x >= threshold;
    keep[4 * call + 1] = r.y >= threshold;
    keep[4 * call + 2] = r.z >= threshold;
    keep[4 * call + 3] = r.w >= threshold;
  }
}

template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  const __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}

template <typename T>
__global__ void __launch_bounds__(256) residual_dropout_add_kernel(const float* __restrict__ x, const T* __restrict__ h,
                                                                    float* __restrict__ out, int64_t vecs, unsigned threshold,
                                                                    float scale, uint64_t seed, uint64_t offset,
                                                                    const uint64_t* __restrict__ philox_dev) {
  if (philox_dev) { seed = philox_dev[0]; offset += philox_dev[1]; }      // graph replay: (seed, base offset) live on the device
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < vecs; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = *reinterpret_cast<const float4*>(x + i * 8), b = *reinterpret_cast<const float4*>(x + i * 8 + 4);
    float f[8];
    unpack8<T>(*reinterpret_cast<const uint4*>(h + i * 8), f);
    bool keep[8];
    keep_mask8(i, seed, offset, threshold, keep);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = keep[e] ? to_f32<T>(from_f32<T>(f[e] * scale)) : 0.f;
    *reinterpret_cast<float4*>(out + i * 8) = make_float4(a.x + f[0], a.y + f[1], a.z + f[2], a.w + f[3]);
    *reinterpret_cast<float4*>(out + i * 8 + 4) = make_float4(b.x + f[4], b.y + f[5], b.z + f[6], b.w + f[7]);
  }
}

// dh = dropout-backward of the branch gradient: bf16(bf16(g) * scale) where kept
template <typename T>
__global__ void __launch_bounds__(256) dropout_branch_grad_kernel(const float* __restrict__ g, T* __restrict__ dh, int64_t vecs,
                                                                   unsigned threshold, float scale, uint64_t seed,
                                                                   uint64_t offset, const uint64_t* __restrict__ philox_dev) {
  if (philox_dev) { seed = philox_dev[0]; offset += philox_dev[1]; }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < vecs; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = *reinterpret_cast<const float4*>(g + i * 8), b = *reinterpret_cast<const float4*>(g + i * 8 + 4);
    float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    bool keep[8];
    keep_mask8(i, seed, offset, threshold, keep);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = keep[e] ? to_f32<T>(from_f32<T>(f[e])) * scale : 0.f;
    uint4 o;
    o.x = pack2<T>(f[0], f[1]); o.y = pack2<T>(f[2], f[3]); o.z = pack2<T>(f[4], f[5]); o.w = pack2<T>(f[6], f[7]);
    *reinterpret_cast<uint4*>(dh + i * 8) = o;
  }
}

}  // namespace svae

using namespace svae;

static int residual_dropout_common(const void* a, const void* b, const void* c, int32_t dtype, int64_t numel, float p,
                                   const char* who) {
  SVAE_REQUIRE(a && b && c && numel >= 0, SVAE_ERR_INVALID, "%s: null argument", who);
  SVAE_REQUIRE(numel % 8 == 0, SVAE_ERR_UNSUPPORTED, "%s: numel must be a multiple of 8", who);
  SVAE_REQUIRE(dtype == SVAE_DTYPE_BF16 || dtype == SVAE_DTYPE_F16, SVAE_ERR_UNSUPPORTED, "%s: 16-bit branch only (dtype %d)", who, dtype);
  SVAE_REQUIRE(p >= 0.f && p < 1.f, SVAE_ERR_INVALID, "%s: dropout probability %f outside [0, 1)", who, (double)p);
  const uintptr_t al = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c);
  SVAE_REQUIRE((al & 15) == 0, SVAE_ERR_INVALID, "%s: tensors must be 16-byte aligned", who);
  return SVAE_OK;
}

static unsigned dropout_threshold(float p) {
  const double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 4294967295u : (unsigned)t;
}

extern "C" int svae_residual_dropout_add_g(const float* x, const void* h, int32_t h_dtype, float* out, int64_t numel, float p,
                                           uint64_t seed, uint64_t offset, const uint64_t* philox_dev, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (int rc = residual_dropout_common(x, h, out, h_dtype, numel, p, "svae_residual_dropout_add")) return rc;
  if (numel == 0) return SVAE_OK;
  ScopedKernelTimer timer("residual_dropout_add", st);
  const int64_t vecs = numel / 8;
  int64_t blocks = (vecs + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  const float scale = 1.0f / (1.0f - p);
  if (h_dtype == SVAE_DTYPE_BF16)
    residual_dropout_add_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(x, (const __nv_bfloat16*)h, out, vecs,
                                                                               dropout_threshold(p), scale, seed, offset, philox_dev);
  else
    residual_dropout_add_kernel<__half><<<(unsigned)blocks, 256, 0, st>>>(x, (const __half*)h, out, vecs, dropout_threshold(p),
                                                                        scale, seed, offset, philox_dev);
  SVAE_CUDA_CHECK(cudaGetLastError());
  return SVAE_OK;
}

extern "C" int svae_residual_dropout_add(const float* x, const void* h, int32_t h_dtype, float* out, int64_t numel, float p,
                                         uint64_t seed, uint64_t offset, void* stream) {
  return svae_residual_dropout_add_g(x, h, h_dtype, out, numel, p, seed, offset, nullptr, stream);
}

extern "C" int svae_dropout_branch_grad_g(const float* g, void* dh, int32_t h_dtype, int64_t numel, float p, uint64_t seed,
                                          uint64_t offset, const uint64_t* philox_dev, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (int rc = residual_dropout_common(g, dh, dh, h_dtype, numel, p, "svae_dropout_branch_grad")) return rc;
  if (numel == 0) return SVAE_OK;
  ScopedKernelTimer timer("dropout_branch_grad", st);
  const int64_t vecs = numel / 8;
  int64_t blocks = (vecs + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  const float scale = 1.0f / (1.0f - p);
  if (h_dtype == SVAE_DTYPE_BF16)
    dropout_branch_grad_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(g, (__nv_bfloat16*)dh, vecs, dropout_threshold(p),
                                                                              scale, seed, offset, philox_dev);
  else
    dropout_branch_grad_kernel<__half><<<(unsigned)blocks, 256, 0, st>>>(g, (__half*)dh, vecs, dropout_threshold(p), scale, seed,
                                                                       offset, philox_dev);
  SVAE_CUDA_CHECK(cudaGetLastError());
  return SVAE_OK;
}


extern "C" int svae_residual_add(const float* x, const void* h, int32_t h_dtype, float* out, int64_t numel, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SVAE_REQUIRE(x && h && out && numel >= 0, SVAE_ERR_INVALID, "svae_residual_add: null argument");
  SVAE_REQUIRE(numel % 8 == 0, SVAE_ERR_UNSUPPORTED, "svae_residual_add: numel must be a multiple of 8");
  SVAE_REQUIRE(h_dtype == SVAE_DTYPE_BF16 || h_dtype == SVAE_DTYPE_F16, SVAE_ERR_UNSUPPORTED,
               "svae_residual_add: the branch must be a 16-bit tensor (dtype %d)", h_dtype);
  const uintptr_t a = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(out);
  SVAE_REQUIRE((a & 15) == 0, SVAE_ERR_INVALID, "svae_residual_add: tensors must be 16-byte aligned");
  if (numel == 0) return SVAE_OK;
  ScopedKernelTimer timer("residual_add", st);
  const int64_t vecs = numel / 8;
  int64_t blocks = (vecs + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (h_dtype == SVAE_DTYPE_BF16)
    residual_add_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(x, (const __nv_bfloat16*)h, out, vecs);
  else
    residual_add_kernel<__half><<<(unsigned)blocks, 256, 0, st>>>(x, (const __half*)h, out, vecs);
  SVAE_CUDA_CHECK(cudaGetLastError());
  return SVAE_OK;
}

extern "C" int svae_dropout_branch_grad(const float* g, void* dh, int32_t h_dtype, int64_t numel, float p, uint64_t seed,
                                        uint64_t offset, void* stream) {
  return svae_dropout_branch_grad_g(g, dh, h_dtype, numel, p, seed, offset, nullptr, stream);
}
