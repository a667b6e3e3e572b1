// Row-wise log-softmax cross-entropy over the vocabulary with the gradient written IN PLACE over the logits
// (fused vocabulary head + chunked cross-entropy; reference
// core/language_model.py:98-113,161-170 `get_nll` / `robust_cross_entropy` -> F.cross_entropy(ignore_index=0)).
//
// The host side (core/fused_ce.py) computes the logits of a few thousand rows at a time with a library GEMM, calls
// this kernel, and immediately feeds the in-place gradient to the two backward GEMMs, so the [B*L, 32768] logits
// (4.3 GB at the C2 shape, 8.6 GB more for ATen's fp32 copy) never exist as a whole and are touched exactly twice
// (one read, one write) while L2-resident.  One CTA per row, the row lives in registers:
//   nll[r]      = logsumexp(logits[r, :]) - logits[r, label[r]]          (fp32 arithmetic on the stored logits)
//   logits[r,j] <- weight[r] * (softmax(logits[r, :])[j] - [j == label[r]])   rounded to the logits' dtype
// weight[r] = 0 marks an ignored row (label == ignore_index): nll 0, gradient 0.
// HBM/L2-bound: V * sizeof(T) bytes read + written per row.
#include "common.cuh"

namespace svae {

constexpr int kCeThreads = 1024;

__device__ __forceinline__ float ce_exp2(float x) {      // x <= 0 here; flush-to-zero of tiny results is exact enough
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <typename T> struct Row8;      // 8 consecutive elements <-> 8 floats
template <> struct Row8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Row8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <> struct Row8<__half> {
  static __device__ __forceinline__ void load(const __half* p, float (&v)[8]) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(__half* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

template <bool IS_MAX>
__device__ __forceinline__ float block_reduce(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float other = __shfl_xor_sync(0xffffffffu, v, o);
    v = IS_MAX ? fmaxf(v, other) : v + other;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();                       // `red` may still be read from the previous reduction
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[lane];                   // kCeThreads / 32 == 32 partials
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float other = __shfl_xor_sync(0xffffffffu, r, o);
    r = IS_MAX ? fmaxf(r, other) : r + other;
  }
  return r;
}

// V = 8192 * K ; thread t owns elements (i*1024 + t)*8 .. +7 for i < K
template <typename T, int K>
__global__ void __launch_bounds__(kCeThreads, 1)
vocab_ce_kernel(T* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels, const float* __restrict__ weight,
                float* __restrict__ nll, int write_grad) {
  __shared__ float red[32];
  const int64_t r = blockIdx.x;
  T* row = logits + r * ld;
  const float w = weight[r];
  const int64_t label = labels[r];
  if (w == 0.f) {                        // ignored row (block-uniform branch)
    if (threadIdx.x == 0) nll[r] = 0.f;
    if (write_grad) {
      float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < K; ++i) Row8<T>::store(row + (i * kCeThreads + threadIdx.x) * 8, z);
    }
    return;
  }
  float v[K][8];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    Row8<T>::load(row + (i * kCeThreads + threadIdx.x) * 8, v[i]);
#pragma unroll
    for (int e = 0; e < 8; ++e) m = fmaxf(m, v[i][e]);
  }
  m = block_reduce<true>(m, red);
  constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
  const float neg_m2 = -m * kLog2e;
  float s = 0.f;
  float at_label = 0.f;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    const int64_t col0 = (int64_t)(i * kCeThreads + threadIdx.x) * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (col0 + e == label) at_label = v[i][e];
      v[i][e] = ce_exp2(fmaf(v[i][e], kLog2e, neg_m2));
      s += v[i][e];
    }
  }
  s = block_reduce<false>(s, red);
  // the label's logit lives in exactly one thread: broadcast it through the sum reduction
  at_label = block_reduce<false>(at_label, red);
  if (threadIdx.x == 0) nll[r] = (m + log2f(s) * kLn2) - at_label;
  if (write_grad) {
    const float scale = w / s;
#pragma unroll
    for (int i = 0; i < K; ++i) {
      const int64_t col0 = (int64_t)(i * kCeThreads + threadIdx.x) * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[i][e] = fmaf(v[i][e], scale, (col0 + e == label) ? -w : 0.f);
      Row8<T>::store(row + col0, v[i]);
    }
  }
}

// 16-bit logits: 512 threads per row, the row stays PACKED in registers (8 x 16 bytes per thread at V = 32768) and the
// exponentials are evaluated twice (sum pass, gradient pass) -- MUFU has headroom, registers do not: at <= 64
// registers two CTAs share an SM, so one row's loads / stores overlap the other row's arithmetic.
constexpr int kCe16Threads = 512;

template <typename T> __device__ __forceinline__ float2 unpack2(uint32_t u);
template <> __device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
}
template <> __device__ __forceinline__ float2 unpack2<__half>(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }
template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// reduces (a: max or sum, b: sum) over the CTA in one round trip through shared memory
template <bool A_IS_MAX>
__device__ __forceinline__ void block_reduce2_512(float& a, float& b, float (*red)[2]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float oa = __shfl_xor_sync(0xffffffffu, a, o), ob = __shfl_xor_sync(0xffffffffu, b, o);
    a = A_IS_MAX ? fmaxf(a, oa) : a + oa;
    b += ob;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[warp][0] = a; red[warp][1] = b; }
  __syncthreads();
  a = lane < kCe16Threads / 32 ? red[lane][0] : (A_IS_MAX ? -INFINITY : 0.f);
  b = lane < kCe16Threads / 32 ? red[lane][1] : 0.f;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    const float oa = __shfl_xor_sync(0xffffffffu, a, o), ob = __shfl_xor_sync(0xffffffffu, b, o);
    a = A_IS_MAX ? fmaxf(a, oa) : a + oa;
    b += ob;
  }
  a = __shfl_sync(0xffffffffu, a, 0);
  b = __shfl_sync(0xffffffffu, b, 0);
}

// V = 4096 * G ; thread t owns the 8 elements at (i*512 + t)*8 for i < G
template <typename T, int G>
__global__ void __launch_bounds__(kCe16Threads, 2)
vocab_ce16_kernel(T* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels, const float* __restrict__ weight,
                  float* __restrict__ nll, int write_grad) {
  __shared__ float red[2][kCe16Threads / 32][2];
  const int64_t r = blockIdx.x;
  T* row = logits + r * ld;
  const float w = weight[r];
  const int label = (int)labels[r];
  if (w == 0.f) {                        // ignored row (block-uniform branch)
    if (threadIdx.x == 0) nll[r] = 0.f;
    if (write_grad) {
#pragma unroll
      for (int i = 0; i < G; ++i) *reinterpret_cast<uint4*>(row + (i * kCe16Threads + threadIdx.x) * 8) = make_uint4(0, 0, 0, 0);
    }
    return;
  }
  uint4 raw[G];
#pragma unroll
  for (int i = 0; i < G; ++i) raw[i] = *reinterpret_cast<const uint4*>(row + (i * kCe16Threads + threadIdx.x) * 8);
  const float x_label = to_f32<T>(row[label]);            // same value in every thread (L1 broadcast)
  constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
  float m = -INFINITY, dummy = 0.f;
#pragma unroll
  for (int i = 0; i < G; ++i) {
    const uint32_t u[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack2<T>(u[e]);
      m = fmaxf(m, fmaxf(f.x, f.y));
    }
  }
  block_reduce2_512<true>(m, dummy, red[0]);
  const float neg_m2 = -m * kLog2e;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < G; ++i) {
    const uint32_t u[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack2<T>(u[e]);
      s += ce_exp2(fmaf(f.x, kLog2e, neg_m2)) + ce_exp2(fmaf(f.y, kLog2e, neg_m2));
    }
  }
  block_reduce2_512<false>(s, dummy, red[1]);
  if (threadIdx.x == 0) nll[r] = (m + log2f(s) * kLn2) - x_label;
  if (write_grad) {
    const float scale = w / s;
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const uint32_t u[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack2<T>(u[e]);
        o[e] = pack2<T>(ce_exp2(fmaf(f.x, kLog2e, neg_m2)) * scale, ce_exp2(fmaf(f.y, kLog2e, neg_m2)) * scale);
      }
      *reinterpret_cast<uint4*>(row + (i * kCe16Threads + threadIdx.x) * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    __syncthreads();                                       // the label's element has been written by its owner
    if (threadIdx.x == 0)                                  // softmax - 1 at the label, rounded once from fp32
      row[label] = from_f32<T>(fmaf(ce_exp2(fmaf(x_label, kLog2e, neg_m2)), scale, -w));
  }
}

template <typename T>
static int launch_ce(int k, void* logits, int64_t ld, const int64_t* labels, const float* weight, float* nll, int write_grad,
                     int64_t rows, cudaStream_t st) {
#define SVAE_CE(KK) \
  case KK: vocab_ce_kernel<T, KK><<<(unsigned)rows, kCeThreads, 0, st>>>((T*)logits, ld, labels, weight, nll, write_grad); break
  switch (k) { SVAE_CE(1); SVAE_CE(2); SVAE_CE(3); SVAE_CE(4); default: return SVAE_ERR_UNSUPPORTED; }
#undef SVAE_CE
  SVAE_CUDA_CHECK(cudaGetLastError());
  return SVAE_OK;
}

template <typename T>
static int launch_ce16(int k, void* logits, int64_t ld, const int64_t* labels, const float* weight, float* nll, int write_grad,
                       int64_t rows, cudaStream_t st) {
#define SVAE_CE16(KK) \
  case KK: vocab_ce16_kernel<T, 2 * KK><<<(unsigned)rows, kCe16Threads, 0, st>>>((T*)logits, ld, labels, weight, nll, write_grad); break
  switch (k) { SVAE_CE16(1); SVAE_CE16(2); SVAE_CE16(3); SVAE_CE16(4); default: return SVAE_ERR_UNSUPPORTED; }
#undef SVAE_CE16
  SVAE_CUDA_CHECK(cudaGetLastError());
  return SVAE_OK;
}

}  // namespace svae

using namespace svae;

extern "C" int svae_vocab_ce_supported(int32_t vocab) { return (vocab % 8192 == 0 && vocab >= 8192 && vocab <= 32768) ? 1 : 0; }

extern "C" int svae_vocab_ce(void* logits, int32_t dtype, int64_t rows, int32_t vocab, int64_t ld, const int64_t* labels,
                             const float* weight, float* nll, int32_t write_grad, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SVAE_REQUIRE(logits && labels && weight && nll && rows >= 0, SVAE_ERR_INVALID, "svae_vocab_ce: null argument");
  SVAE_REQUIRE(svae_vocab_ce_supported(vocab), SVAE_ERR_UNSUPPORTED, "svae_vocab_ce: vocabulary %d is not 8192*k, k <= 4", vocab);
  SVAE_REQUIRE(ld >= vocab && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0, SVAE_ERR_INVALID,
               "svae_vocab_ce: rows must be 16-byte aligned (ld %% 8 == 0)");
  SVAE_REQUIRE(rows < (int64_t)1 << 31, SVAE_ERR_INVALID, "svae_vocab_ce: too many rows for one launch");
  if (rows == 0) return SVAE_OK;
  ScopedKernelTimer timer("vocab_ce", st);
  const int k = vocab / 8192;
  if (dtype == SVAE_DTYPE_F32) return launch_ce<float>(k, logits, ld, labels, weight, nll, write_grad, rows, st);
  if (dtype == SVAE_DTYPE_BF16) return launch_ce16<__nv_bfloat16>(k, logits, ld, labels, weight, nll, write_grad, rows, st);
  if (dtype == SVAE_DTYPE_F16) return launch_ce16<__half>(k, logits, ld, labels, weight, nll, write_grad, rows, st);
  SVAE_REQUIRE(false, SVAE_ERR_INVALID, "svae_vocab_ce: dtype %d", dtype);
}
