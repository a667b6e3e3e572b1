// erf-GELU of the feed-forward blocks (reference core/transformer_layer.py:20-24 `nn.GELU()` between the two ffn
// projections; F.gelu = x * Phi(x), Phi the normal CDF), forward and backward, HBM-bound.
//
// ATen's kernels evaluate erff per element (~25 FP32 instructions) and are ALU-bound on a B200: 148 us forward /
// ~200 us backward at [65536, 2048] bf16 where the traffic (537 / 805 MB) allows 84 / 125 us.  Here
//   Phi(-a) = 2^-G(a),  G(a) = 1 + a * Q(2a/6 - 1),  a = min(|x|, 6),  Q a degree-6 polynomial
// (minimax fit of -log2(Phi(-a)); relative error of Phi(-a) <= 7e-6 over the whole range, absolute error of Phi <= 2e-6;
// beyond |x| = 6 the tail is held at Phi(-6) = 1e-9, below the 6e-8 granularity of torch's own 1 + erff(x / sqrt 2)),
// evaluated two elements at a time with the packed FP32 FMA of sm_100 (FFMA2) and ONE MUFU.EX2 per element;
//   gelu(x)  = x * Phi(x),            Phi(x) = x < 0 ? Phi(-a) : 1 - Phi(-a)
//   gelu'(x) = Phi(x) + x * phi(x),   phi(x) = 2^(-x^2 * log2(e)/2) / sqrt(2 pi)   (a second MUFU.EX2)
// The 16-bit result differs from ATen's rounding of its fp32 value by one 16-bit ulp on ~0.4 % of the elements (both are
// within half an ulp + 2e-6 of the exact value; tests/test_gpu_gelu.py states the bound).
//
// The backward also produces the COLUMN SUMS of its (unrounded fp32) output -- the bias gradient of the up-projection that precedes the
// GELU -- in the same pass (per-slab fp32 partials, summed in slab order by the last block of a column group to finish:
// deterministic, one launch), which removes the separate read of the [rows, 4 d_model] gradient by `svae_colsum`.
// Bytes: forward 2 * rows * n * s, backward 3 * rows * n * s.
#include "common.cuh"

namespace svae {

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr float kGeluA = 6.0f;
// Q(s), s = 2a/6 - 1 in [-1, 1] (monomial coefficients c0..c6)
__device__ __forceinline__ float2 gelu_tail2(float2 x) {      // Phi(-|x|) for two elements
  const float2 a = make_float2(fminf(fabsf(x.x), kGeluA), fminf(fabsf(x.y), kGeluA));
  const float2 s = ffma2(a, splat(2.0f / kGeluA), splat(-1.0f));
  float2 q = ffma2(splat(1.307678630e-03f), s, splat(-6.910470003e-03f));
  q = ffma2(q, s, splat(2.061073352e-02f));
  q = ffma2(q, s, splat(-5.117394290e-02f));
  q = ffma2(q, s, splat(1.191125043e-01f));
  q = ffma2(q, s, splat(1.892212900e+00f));
  q = ffma2(q, s, splat(2.844311945e+00f));
  const float2 g = ffma2(a, q, splat(1.0f));                  // -log2 Phi(-a)
  return make_float2(ex2_approx(-g.x), ex2_approx(-g.y));
}
__device__ __forceinline__ float2 gelu_cdf2(float2 x) {
  const float2 t = gelu_tail2(x);
  return make_float2(x.x < 0.f ? t.x : 1.0f - t.x, x.y < 0.f ? t.y : 1.0f - t.y);
}
__device__ __forceinline__ float2 gelu2(float2 x) { return fmul2(x, gelu_cdf2(x)); }
__device__ __forceinline__ float2 gelu_grad2(float2 x) {      // Phi(x) + x phi(x)
  const float2 cdf = gelu_cdf2(x);
  const float2 e = fmul2(fmul2(x, x), splat(-0.7213475204444817f));
  const float2 pdf = make_float2(ex2_approx(e.x), ex2_approx(e.y));
  return ffma2(fmul2(x, splat(0.3989422804014327f)), pdf, cdf);
}

template <typename T> struct Pair;      // two consecutive 16-bit elements in one 32-bit word
template <> struct Pair<__nv_bfloat16> {
  static __device__ __forceinline__ float2 unpack(uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }
  static __device__ __forceinline__ uint32_t pack(float2 v) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
};
template <> struct Pair<__half> {
  static __device__ __forceinline__ float2 unpack(uint32_t w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }
  static __device__ __forceinline__ uint32_t pack(float2 v) {
    const __half2 h = __floats2half2_rn(v.x, v.y);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
};

// ------------------------------------------------------------------------------------------------ forward
// one 16-byte vector (8 elements) per thread and iteration, two in flight
template <typename T>
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int64_t nvec) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  auto one = [](uint4 v) {
    uint4 o;
    o.x = Pair<T>::pack(gelu2(Pair<T>::unpack(v.x)));
    o.y = Pair<T>::pack(gelu2(Pair<T>::unpack(v.y)));
    o.z = Pair<T>::pack(gelu2(Pair<T>::unpack(v.z)));
    o.w = Pair<T>::pack(gelu2(Pair<T>::unpack(v.w)));
    return o;
  };
  for (; i + stride < nvec; i += 2 * stride) {
    const uint4 a = __ldcs(x + i), b = __ldcs(x + i + stride);
    y[i] = one(a);
    y[i + stride] = one(b);
  }
  if (i < nvec) y[i] = one(__ldcs(x + i));
}

// ------------------------------------------------------------------------------------------------ backward (+ column sums)
// block = 32 column vectors x 8 row lanes walking a slab of rows (the layout of colsum_partial_kernel); dx may alias dy
template <typename T>
__global__ void __launch_bounds__(256, 4) gelu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, T* __restrict__ dx,
                                                       int64_t rows, int n, float* __restrict__ partial,
                                                       unsigned* __restrict__ counters, float* __restrict__ colsum) {
  __shared__ float red[8][32][9];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int vec = blockIdx.x * 32 + cx;
  const bool ok = vec * 8 < n;
  float acc[8];
  float2 acc2[4] = {splat(0.f), splat(0.f), splat(0.f), splat(0.f)};
  auto one = [&](uint4 g, uint4 v, int64_t off) {
    const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, vw[4] = {v.x, v.y, v.z, v.w};
    uint32_t ow[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 d = fmul2(Pair<T>::unpack(gw[k]), gelu_grad2(Pair<T>::unpack(vw[k])));
      ow[k] = Pair<T>::pack(d);
      acc2[k] = fadd2(acc2[k], d);                    // the bias gradient sums the fp32 products (before the 16-bit rounding)
    }
    *reinterpret_cast<uint4*>(dx + off) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  };
  if (ok) {
    // software pipeline: the loads of the NEXT pair of rows are in flight while this pair is evaluated (~270 instructions
    // per thread and pair; without it the warps spent half their time waiting: ncu issue-active 44 %, DRAM 55 %)
    const int64_t step = (int64_t)gridDim.y * 8;
    int64_t r = (int64_t)blockIdx.y * 8 + ry;
    auto ld = [&](const T* p, int64_t row) {
      return row < rows ? __ldcs(reinterpret_cast<const uint4*>(p + row * n + vec * 8)) : make_uint4(0, 0, 0, 0);
    };
    uint4 g0 = ld(dy, r), v0 = ld(x, r), g1 = ld(dy, r + step), v1 = ld(x, r + step);
    for (; r < rows; r += 2 * step) {
      const int64_t rn = r + 2 * step;
      const uint4 ng0 = ld(dy, rn), nv0 = ld(x, rn), ng1 = ld(dy, rn + step), nv1 = ld(x, rn + step);
      one(g0, v0, r * n + vec * 8);
      if (r + step < rows) one(g1, v1, (r + step) * n + vec * 8);
      g0 = ng0; v0 = nv0; g1 = ng1; v1 = nv1;
    }
  }
  if (colsum == nullptr) return;
#pragma unroll
  for (int k = 0; k < 4; ++k) { acc[2 * k] = acc2[k].x; acc[2 * k + 1] = acc2[k].y; }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[ry][cx][e] = acc[e];
  __syncthreads();
  if (ry == 0 && ok) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += red[k][cx][e];
      partial[(int64_t)blockIdx.y * n + vec * 8 + e] = t;
    }
  }
  // the LAST block of this column group to finish sums the slabs in slab order and leaves the ticket at zero
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(&counters[blockIdx.x], 1u) == gridDim.y - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (ok)
    for (int b = ry; b < (int)gridDim.y; b += 8) {
      const float4 u = __ldcg(reinterpret_cast<const float4*>(partial + (int64_t)b * n + vec * 8));
      const float4 w = __ldcg(reinterpret_cast<const float4*>(partial + (int64_t)b * n + vec * 8 + 4));
      acc[0] += u.x; acc[1] += u.y; acc[2] += u.z; acc[3] += u.w; acc[4] += w.x; acc[5] += w.y; acc[6] += w.z; acc[7] += w.w;
    }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[ry][cx][e] = acc[e];
  __syncthreads();
  if (ry == 0 && ok) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += red[k][cx][e];
      colsum[vec * 8 + e] = t;
    }
  }
  if (threadIdx.x == 0) counters[blockIdx.x] = 0u;
}

static int gelu_slabs(int64_t rows, int n) {
  const int col_blocks = (n / 8 + 31) / 32;
  const int sms = sm_count_of_current_device();
  int slabs = (sms * 4) / col_blocks;                            // ONE wave: 4 resident blocks of 256 threads per SM (<= 64 registers)
  if (slabs < 1) slabs = 1;
  const int64_t max_slabs = (rows + 15) / 16;                   // at least two rows per row lane
  if (slabs > max_slabs) slabs = (int)(max_slabs > 0 ? max_slabs : 1);
  return slabs;
}

}  // namespace svae

using namespace svae;

extern "C" int32_t svae_gelu_supported(int32_t dtype, int64_t rows, int32_t n) {
  return (dtype == SVAE_DTYPE_BF16 || dtype == SVAE_DTYPE_F16) && rows > 0 && n > 0 && n % 8 == 0;
}

extern "C" int svae_gelu_fwd(const void* x, void* y, int32_t dtype, int64_t rows, int32_t n, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SVAE_REQUIRE(svae_gelu_supported(dtype, rows, n), SVAE_ERR_UNSUPPORTED, "svae_gelu_fwd: 16-bit dtype and n %% 8 == 0 required");
  SVAE_REQUIRE(x && y && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, SVAE_ERR_INVALID,
               "svae_gelu_fwd: null or misaligned pointer");
  const int64_t nvec = rows * (n / 8);
  const int sms = sm_count_of_current_device();
  int64_t blocks = (nvec + 2 * 256 - 1) / (2 * 256);
  if (blocks > (int64_t)sms * 8 * 4) blocks = (int64_t)sms * 8 * 4;
  ScopedKernelTimer timer("gelu_fwd", st);
  if (dtype == SVAE_DTYPE_BF16)
    gelu_fwd_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((const uint4*)x, (uint4*)y, nvec);
  else
    gelu_fwd_kernel<__half><<<(unsigned)blocks, 256, 0, st>>>((const uint4*)x, (uint4*)y, nvec);
  SVAE_CUDA_CHECK(cudaGetLastError());
  return SVAE_OK;
}

extern "C" int64_t svae_gelu_bwd_workspace_floats(int64_t rows, int32_t n) { return (int64_t)gelu_slabs(rows, n) * n; }

extern "C" int32_t svae_gelu_bwd_counters(int32_t n) { return (n / 8 + 31) / 32; }

extern "C" int svae_gelu_bwd(const void* dy, const void* x, void* dx, int32_t dtype, int64_t rows, int32_t n, float* colsum,
                             float* workspace, int64_t workspace_floats, uint32_t* counters, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SVAE_REQUIRE(svae_gelu_supported(dtype, rows, n), SVAE_ERR_UNSUPPORTED, "svae_gelu_bwd: 16-bit dtype and n %% 8 == 0 required");
  SVAE_REQUIRE(dy && x && dx &&
                   ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0,
               SVAE_ERR_INVALID, "svae_gelu_bwd: null or misaligned pointer");
  const int slabs = gelu_slabs(rows, n);
  if (colsum)
    SVAE_REQUIRE(workspace && counters && workspace_floats >= (int64_t)slabs * n, SVAE_ERR_INVALID,
                 "svae_gelu_bwd: column sums need a workspace of svae_gelu_bwd_workspace_floats() and zeroed counters");
  dim3 grid((n / 8 + 31) / 32, slabs);
  ScopedKernelTimer timer("gelu_bwd", st);
  if (dtype == SVAE_DTYPE_BF16)
    gelu_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (__nv_bfloat16*)dx, rows, n,
                                                        workspace, counters, colsum);
  else
    gelu_bwd_kernel<__half><<<grid, 256, 0, st>>>((const __half*)dy, (const __half*)x, (__half*)dx, rows, n, workspace, counters, colsum);
  SVAE_CUDA_CHECK(cudaGetLastError());
  return SVAE_OK;
}
