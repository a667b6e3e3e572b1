// Weight gradient of the token embedding (reference core/transformer_language_model.py:47-53 `nn.Embedding` at the head
// of `input_layer`): dW[v, :] = sum over the positions p with ids[p] == v of g[p, :].
//
// ATen's deterministic embedding_dense_backward costs 372 us at 65536 tokens x 512 features (plus its own radix sort).
// Here the caller sorts the token ids once (stable, so equal ids keep their position order), locates every token's segment
// of the sorted order (one searchsorted over the vocabulary) and ONE launch writes the whole gradient: block v sums the
// gradient rows of its positions in position order -- fixed order, no atomics: bit-deterministic -- and writes row v (zeros when the token does not occur),
// so the [vocab, d] gradient needs no zero-fill either.  HBM-bound: n * d * s bytes read + vocab * d * 4 written.
// (A token that fills most of the batch is summed by one block at that block's load rate; the synthetic and any natural
// token distribution spread the rows over thousands of blocks.)
#include "common.cuh"

namespace svae {

template <typename T> __device__ __forceinline__ float4 load4(const T* p);       // 4 consecutive elements, one request
template <> __device__ __forceinline__ float4 load4<float>(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
template <> __device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 u = __ldcs(reinterpret_cast<const uint2*>(p));
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                     __uint_as_float(u.y & 0xffff0000u));
}
template <> __device__ __forceinline__ float4 load4<__half>(const __half* p) {
  const uint2 u = __ldcs(reinterpret_cast<const uint2*>(p));
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

template <typename T>
__global__ void __launch_bounds__(128) embedding_bwd_kernel(const T* __restrict__ g, const int64_t* __restrict__ bounds,
                                                            const int64_t* __restrict__ perm, int d, float* __restrict__ dw) {
  const int64_t v = blockIdx.x;
  const int64_t first = bounds[v], last = bounds[v + 1];      // positions of token v in the sorted order
  for (int c = threadIdx.x * 4; c < d; c += 128 * 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int64_t j = first;
    for (; j + 1 < last; j += 2) {               // two rows in flight, added in position order
      const float4 a = load4<T>(g + perm[j] * d + c), b = load4<T>(g + perm[j + 1] * d + c);
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
      acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
    }
    if (j < last) {
      const float4 a = load4<T>(g + perm[j] * d + c);
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
    *reinterpret_cast<float4*>(dw + v * d + c) = acc;
  }
}

}  // namespace svae

using namespace svae;

extern "C" int svae_embedding_bwd(const void* grad, int32_t dtype, const int64_t* bounds, const int64_t* perm, int64_t n,
                                  int32_t vocab, int32_t d, float* dweight, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SVAE_REQUIRE(grad && bounds && perm && dweight && n >= 0 && vocab > 0, SVAE_ERR_INVALID, "svae_embedding_bwd: null argument");
  SVAE_REQUIRE(d > 0 && d % 4 == 0 && (reinterpret_cast<uintptr_t>(dweight) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(grad) & (dtype == SVAE_DTYPE_F32 ? 15 : 7)) == 0,
               SVAE_ERR_INVALID, "svae_embedding_bwd: d must be a multiple of 4 and the tensors 16-byte aligned");
  ScopedKernelTimer timer("embedding_bwd", st);
  if (dtype == SVAE_DTYPE_F32)
    embedding_bwd_kernel<float><<<vocab, 128, 0, st>>>((const float*)grad, bounds, perm, d, dweight);
  else if (dtype == SVAE_DTYPE_BF16)
    embedding_bwd_kernel<__nv_bfloat16><<<vocab, 128, 0, st>>>((const __nv_bfloat16*)grad, bounds, perm, d, dweight);
  else if (dtype == SVAE_DTYPE_F16)
    embedding_bwd_kernel<__half><<<vocab, 128, 0, st>>>((const __half*)grad, bounds, perm, d, dweight);
  else
    SVAE_REQUIRE(false, SVAE_ERR_INVALID, "svae_embedding_bwd: dtype %d", dtype);
  SVAE_CUDA_CHECK(cudaGetLastError());
  return SVAE_OK;
}
