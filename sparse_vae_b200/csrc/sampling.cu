// One-launch token sampler for graphed decoding (reference core/generation.py:40-72
// `GenerationState.process_logits` with its defaults: repetition penalty over the last 512 tokens, temperature,
// nucleus (top-p) filtering, one multinomial draw, end-token bookkeeping).
//
// The reference does this with ~25 launches, the expensive ones being a full descending sort of every row
// (2 segmented radix-sort passes over [B, 32768]) and a cumsum -- 480 us per token at B = 256.  A nucleus does not
// need the order, only the threshold: one CTA per row keeps the row's exp() values in registers and
//   1. marks the tokens generated in the last `window` steps in a shared-memory bitmap and applies the penalty
//      (x < 0 ? x * r : x / r, evaluated like ATen: fp32 product rounded to the logits dtype, division by a host
//      scalar = multiplication by its fp32 reciprocal), then the temperature the same way;
//   2. softmax statistics in fp32 (block max, block sum, deterministic tree reductions);
//   3. finds the smallest 16-bit logit value whose upper tail has mass <= top_p by bisection over the 65,536
//      possible keys -- the set "sorted cumsum <= top_p" of the reference (its first token is always kept).  Tokens
//      tying at the boundary are admitted in index order while they fit (the reference admits them in its
//      unstable sort's order);
//   4. draws the token by inverse CDF over the kept tokens from ONE uniform per row supplied by the caller (torch.rand,
//      so the generator state stays torch's, also under graph capture).  Same distribution as the reference's
//      exponential-race multinomial, not the same random stream;
//   5. writes the token into ids[b, column] if the sample is alive, clears `alive` on the end token and counts the
//      samples that finished in this step.
// HBM traffic: the logits once (B * V * 2 bytes); the bisection runs from registers.
#include "common.cuh"

namespace svae {

constexpr int kSampThreads = 1024;
constexpr int kSampPerThread = 32;                 // V <= 32768
constexpr int kSampVec = 8;                        // 16-byte loads of 16-bit logits

template <typename T> struct alignas(16) SPack { T v[kSampVec]; };

__device__ __forceinline__ float block_reduce_sum(float v, float* s_red) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();                                  // s_red may still be read from the previous call
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = s_red[threadIdx.x & 31];
#pragma unroll
  for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;                                         // identical in every thread
}

__device__ __forceinline__ float block_reduce_max(float v, float* s_red) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = s_red[threadIdx.x & 31];
#pragma unroll
  for (int o = 16; o; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
  return t;
}

// exclusive prefix of v over the thread index (thread-major order); total returned through `total`
__device__ __forceinline__ float block_scan_exclusive(float v, float* s_red, float& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  __syncthreads();
  if (lane == 31) s_red[warp] = inc;
  __syncthreads();
  float w = s_red[lane];
  float winc = w;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_up_sync(0xffffffffu, winc, o);
    if (lane >= o) winc += n;
  }
  total = __shfl_sync(0xffffffffu, winc, 31);
  const float warp_offset = __shfl_sync(0xffffffffu, winc - w, warp);
  return warp_offset + inc - v;
}

// Keys order the 16-bit patterns like their values.  kKeyLo = key of -inf, kKeyHi = one past the key of +inf; the keys
// outside [kKeyLo, kKeyHi) are NaN patterns and never bisected.
template <typename T> struct KeyRange;
template <> struct KeyRange<__half> { static constexpr int lo = 0x03ff, hi = 0xfc01; };
template <> struct KeyRange<__nv_bfloat16> { static constexpr int lo = 0x007f, hi = 0xff81; };
template <typename T> __device__ __forceinline__ float key_to_value(int key);
// ascending 16-bit key -> the 16-bit float with that rank (sign-magnitude order)
template <> __device__ __forceinline__ float key_to_value<__half>(int key) {
  const unsigned short bits = (key & 0x8000) ? (unsigned short)(key & 0x7fff) : (unsigned short)(~key & 0xffff);
  return __half2float(__ushort_as_half(bits));
}
template <> __device__ __forceinline__ float key_to_value<__nv_bfloat16>(int key) {
  const unsigned short bits = (key & 0x8000) ? (unsigned short)(key & 0x7fff) : (unsigned short)(~key & 0xffff);
  return __bfloat162float(__ushort_as_bfloat16(bits));
}

template <typename T> __device__ __forceinline__ int value_to_key(float x);   // x exactly representable in T
template <> __device__ __forceinline__ int value_to_key<__half>(float x) {
  const int bits = __half_as_ushort(__float2half_rn(x));
  return (bits & 0x8000) ? (~bits & 0xffff) : (bits | 0x8000);
}
template <> __device__ __forceinline__ int value_to_key<__nv_bfloat16>(float x) {
  const int bits = __bfloat16_as_ushort(__float2bfloat16_rn(x));
  return (bits & 0x8000) ? (~bits & 0xffff) : (bits | 0x8000);
}

template <typename T>
__global__ void __launch_bounds__(kSampThreads) sample_top_p_kernel(
    const T* __restrict__ logits, int V, int64_t* __restrict__ ids, int64_t ids_stride, const int64_t* __restrict__ column_ptr,
    const float* __restrict__ uniforms, unsigned char* __restrict__ alive, int* __restrict__ finished, int window,
    float penalty, float temperature, float top_p, int64_t end_token) {
  __shared__ unsigned s_seen[1024];                 // bitmap over V <= 32768 token ids
  __shared__ float s_red[32];
  __shared__ int s_pick[2];
  const int b = blockIdx.x, t = threadIdx.x;
  const int64_t column = *column_ptr;
  int64_t* row_ids = ids + (int64_t)b * ids_stride;

  s_seen[t] = 0u;
  if (t == 0) { s_pick[0] = -1; s_pick[1] = -1; }
  __syncthreads();
  if (penalty > 1.0f && t < window && column - 1 - t >= 0) {
    const int64_t tok = row_ids[column - 1 - t];
    if (tok >= 0 && tok < V) atomicOr(&s_seen[tok >> 5], 1u << (tok & 31));
  }
  __syncthreads();

  // ---- load, penalty, temperature (ATen's arithmetic), exp --------------------------------------------------------
  const float inv_penalty = 1.0f / penalty, inv_temperature = 1.0f / temperature;
  float e[kSampPerThread];
  float row_max = -INFINITY, row_min = INFINITY;
#pragma unroll
  for (int c = 0; c < kSampPerThread / kSampVec; ++c) {
    const int base = c * (kSampThreads * kSampVec) + t * kSampVec;
    if (base < V) {
      const SPack<T> x = *reinterpret_cast<const SPack<T>*>(logits + (size_t)b * V + base);
      const unsigned seen = (s_seen[base >> 5] >> (base & 31)) & 0xffu;
#pragma unroll
      for (int i = 0; i < kSampVec; ++i) {
        float v = to_f32<T>(x.v[i]);
        if ((seen >> i) & 1u) v = to_f32<T>(from_f32<T>(v < 0.f ? v * penalty : v * inv_penalty));
        if (temperature != 1.0f) v = to_f32<T>(from_f32<T>(v * inv_temperature));
        e[c * kSampVec + i] = v;
        row_max = fmaxf(row_max, v);
        row_min = fminf(row_min, v);
      }
    } else {
#pragma unroll
      for (int i = 0; i < kSampVec; ++i) e[c * kSampVec + i] = -INFINITY;
    }
  }
  row_max = block_reduce_max(row_max, s_red);
  row_min = -block_reduce_max(-row_min, s_red);
  float part = 0.f;
#pragma unroll
  for (int i = 0; i < kSampPerThread; ++i) {
    e[i] = expf(e[i] - row_max);                    // exp(-inf) = 0 for the padding beyond V
    part += e[i];
  }
  const float Z = block_reduce_sum(part, s_red);

  // ---- nucleus threshold: smallest key whose upper tail {x >= value(key)} has mass <= top_p * Z --------------------
  float thr_e = 0.f;                                // keep e >= thr_e ...
  float tie_e = -1.f;                               // ... plus the first `n_tie` tokens with e == tie_e
  int n_tie = 0;
  if (top_p < 1.0f) {
    const float budget = top_p * Z;
    // invariant: tail(hi) <= budget < tail(lo - 1).  Only keys between the row's extremes can be the answer:
    // tail(key(max) + 1) = 0 and tail(key(min)) = Z > budget, which saves a third of the rounds on typical rows
    int lo = max(value_to_key<T>(row_min) + 1, KeyRange<T>::lo), hi = min(value_to_key<T>(row_max) + 1, KeyRange<T>::hi);
    if (!(row_min == row_min && row_max == row_max)) { lo = KeyRange<T>::lo; hi = KeyRange<T>::hi; }
    float tail_hi = 0.f;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const float te = expf(key_to_value<T>(mid) - row_max);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < kSampPerThread; ++i) s += e[i] >= te ? e[i] : 0.f;
      const float g = block_reduce_sum(s, s_red);
      if (g <= budget) { hi = mid; tail_hi = g; } else { lo = mid + 1; }
    }
    thr_e = hi >= KeyRange<T>::hi ? INFINITY : expf(key_to_value<T>(hi) - row_max);
    if (hi > KeyRange<T>::lo) {
      tie_e = expf(key_to_value<T>(hi - 1) - row_max);
      if (tie_e >= thr_e) tie_e = -1.f;             // exp() collided with the threshold: already inside the tail
    }
    if (tie_e > 0.f) n_tie = (int)floorf((budget - tail_hi) / tie_e);
    if (tail_hi == 0.f && n_tie < 1) n_tie = 1;     // the most likely token is always kept
  }

  // ---- weights of the kept tokens, in thread-major order -----------------------------------------------------------
  int my_ties = 0;
#pragma unroll
  for (int i = 0; i < kSampPerThread; ++i) my_ties += (e[i] == tie_e);
  float tie_count;
  int tie_rank = (int)block_scan_exclusive((float)my_ties, s_red, tie_count);     // counts < 2^24: exact in fp32
  float w_sum = 0.f;
#pragma unroll
  for (int i = 0; i < kSampPerThread; ++i) {
    bool keep = e[i] >= thr_e;
    if (e[i] == tie_e) keep = tie_rank++ < n_tie;
    if (!keep) e[i] = 0.f;
    w_sum += e[i];
  }
  float total;
  const float before = block_scan_exclusive(w_sum, s_red, total);
  const float target = uniforms[b] * total;
  // the last thread holding mass at or before the target owns the draw
  if (w_sum > 0.f && before <= target) atomicMax(&s_pick[0], t);
  __syncthreads();
  if (s_pick[0] == t) {
    float run = before;
    int chosen = -1;
#pragma unroll
    for (int i = 0; i < kSampPerThread; ++i) {
      if (e[i] > 0.f && (chosen < 0 || run <= target)) chosen = i;     // first kept element, then advance while run <= target
      run += e[i];
    }
    // `chosen` is the last kept element whose preceding mass is <= target
    const int c = chosen / kSampVec, i = chosen % kSampVec;
    s_pick[1] = c * (kSampThreads * kSampVec) + t * kSampVec + i;
  }
  __syncthreads();
  if (t == 0 && alive[b]) {
    const int64_t tok = s_pick[1] < 0 ? 0 : s_pick[1];           // < 0 only for a row of NaNs
    row_ids[column] = tok;
    if (tok == end_token) {
      alive[b] = 0;
      atomicAdd(finished, 1);
    }
  }
}

}  // namespace svae

using namespace svae;

extern "C" int svae_sample_top_p_supported(int32_t vocab, int32_t dtype) {
  return vocab > 0 && vocab <= kSampThreads * kSampPerThread && vocab % kSampVec == 0 &&
         (dtype == SVAE_DTYPE_F16 || dtype == SVAE_DTYPE_BF16);
}

extern "C" int svae_sample_top_p(const void* logits, int32_t dtype, int32_t B, int32_t vocab, int64_t* ids,
                                 int64_t ids_stride, const int64_t* column, const float* uniforms, uint8_t* alive,
                                 int32_t* finished, int32_t penalty_window, float repetition_penalty, float temperature,
                                 float top_p, int64_t end_token, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SVAE_REQUIRE(logits && ids && column && uniforms && alive && finished, SVAE_ERR_INVALID, "svae_sample_top_p: null argument");
  SVAE_REQUIRE(svae_sample_top_p_supported(vocab, dtype), SVAE_ERR_UNSUPPORTED,
               "svae_sample_top_p: vocab %d / dtype %d not supported (16-bit logits, vocab <= 32768 and a multiple of 8)",
               vocab, dtype);
  SVAE_REQUIRE(B >= 0 && penalty_window >= 0 && penalty_window <= kSampThreads && temperature > 0.f && top_p > 0.f &&
                   repetition_penalty > 0.f,
               SVAE_ERR_INVALID, "svae_sample_top_p: bad sampling parameters");
  SVAE_REQUIRE((reinterpret_cast<uintptr_t>(logits) & 15) == 0, SVAE_ERR_INVALID, "svae_sample_top_p: misaligned logits");
  if (B == 0) return SVAE_OK;
  ScopedKernelTimer timer("sample_top_p", st);
  if (dtype == SVAE_DTYPE_F16)
    sample_top_p_kernel<__half><<<B, kSampThreads, 0, st>>>((const __half*)logits, vocab, ids, ids_stride, column, uniforms,
                                                            alive, finished, penalty_window, repetition_penalty,
                                                            temperature, top_p, end_token);
  else
    sample_top_p_kernel<__nv_bfloat16><<<B, kSampThreads, 0, st>>>((const __nv_bfloat16*)logits, vocab, ids, ids_stride,
                                                                   column, uniforms, alive, finished, penalty_window,
                                                                   repetition_penalty, temperature, top_p, end_token);
  SVAE_CUDA_CHECK(cudaGetLastError());
  return SVAE_OK;
}
