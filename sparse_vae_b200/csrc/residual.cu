// Residual-stream update of the transformer blocks under autocast (reference core/transformer_layer.py:41,49,61:
// `x = x + h`): x is the fp32 residual stream, h the 16-bit output of the attention / feed-forward branch.  ATen
// runs this mixed-dtype add through its non-vectorised `unrolled_elementwise_kernel` (112 us at [65536, 512] on a
// B200, 3.0 TB/s); this is the same fp32 addition (promote h, one rounding) with 16-byte accesses.
// HBM-bound: numel * (4 + sizeof(T) + 4) bytes.
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

}  // namespace svae

using namespace svae;

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
