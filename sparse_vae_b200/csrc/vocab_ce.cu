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
#include "sm100_ptx.cuh"

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

// ---- 16-bit logits, STREAMED: one persistent CTA per SM, rows through a 3-deep shared-memory ring ---------------
// The register-resident kernel above alternates per CTA between a load phase, two reduction round trips and a store
// phase; with two rows per SM in flight HBM idles a third of the time (4.3 TB/s).  Here the row of CTA-iteration i is
// computed from shared memory while the bulk copies (cp.async.bulk, mbarrier completion) of rows i+1 and i+2 are in
// flight and the gradient of row i-1 is being written back by a bulk store from the same buffer it arrived in: the
// memory system always holds two row loads and one row store per SM, independent of the arithmetic.
constexpr int kCeSThreads = 1024;
constexpr int kCeSRing = 3;

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ptx::smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ptx::smem_u32(smem_src)), "r"(bytes) : "memory");
}

template <bool A_IS_MAX>
__device__ __forceinline__ float block_reduce_1024(float a, float* red) {      // red: 32 floats, one use per call site
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float oa = __shfl_xor_sync(0xffffffffu, a, o);
    a = A_IS_MAX ? fmaxf(a, oa) : a + oa;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = a;
  __syncthreads();
  a = red[lane];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float oa = __shfl_xor_sync(0xffffffffu, a, o);
    a = A_IS_MAX ? fmaxf(a, oa) : a + oa;
  }
  return a;
}

// V = 8192 * G; thread t owns the 8 elements at (i * 1024 + t) * 8, i < G
template <typename T, int G>
__global__ void __launch_bounds__(kCeSThreads, 1)
vocab_ce16s_kernel(T* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels, const float* __restrict__ weight,
                   float* __restrict__ nll, int write_grad, int64_t rows) {
  constexpr int V = 8192 * G;
  constexpr uint32_t ROW_BYTES = V * sizeof(T);
  extern __shared__ uint8_t ce_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ce_smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kCeSRing * ROW_BYTES);
  float* red = reinterpret_cast<float*>(full + kCeSRing);              // [2][2][32]: max / sum, double-buffered over rows
  constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;

  if (threadIdx.x == 0) {
    for (int b = 0; b < kCeSRing; ++b) ptx::mbar_init(full + b, 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  const int64_t first = blockIdx.x, stride = gridDim.x;
  const int64_t mine = first < rows ? (rows - first + stride - 1) / stride : 0;      // rows of this CTA
  auto row_of = [&](int64_t i) { return first + i * stride; };
  auto request = [&](int64_t i) {                 // thread 0: bring row i into its ring buffer (ignored rows: nothing to read)
    const int b = (int)(i % kCeSRing);
    const int64_t r = row_of(i);
    if (weight[r] == 0.f) { ptx::mbar_arrive(full + b); return; }
    ptx::mbar_arrive_expect_tx(full + b, ROW_BYTES);
#pragma unroll
    for (int c = 0; c < 4; ++c)
      bulk_load_1d(smem + (size_t)b * ROW_BYTES + c * (ROW_BYTES / 4), reinterpret_cast<const uint8_t*>(logits + r * ld) + c * (ROW_BYTES / 4),
                   ROW_BYTES / 4, full + b);
  };
  if (threadIdx.x == 0)
    for (int64_t i = 0; i < mine && i < kCeSRing - 1; ++i) request(i);

  for (int64_t i = 0; i < mine; ++i) {
    const int b = (int)(i % kCeSRing);
    const int64_t r = row_of(i);
    T* buf = reinterpret_cast<T*>(smem + (size_t)b * ROW_BYTES);
    T* grow = logits + r * ld;
    const float w = weight[r];
    float* red_m = red + (i & 1) * 64, *red_s = red_m + 32;
    ptx::mbar_wait(full + b, (uint32_t)((i / kCeSRing) & 1));
    if (w == 0.f) {                         // ignored row (block-uniform): nll 0, zero gradient straight to global memory
      if (threadIdx.x == 0) nll[r] = 0.f;
      if (write_grad) {
#pragma unroll
        for (int k = 0; k < G; ++k) *reinterpret_cast<uint4*>(grow + (k * kCeSThreads + threadIdx.x) * 8) = make_uint4(0, 0, 0, 0);
      }
    } else {
      const int label = (int)labels[r];
      uint4 raw[G];
#pragma unroll
      for (int k = 0; k < G; ++k) raw[k] = *reinterpret_cast<const uint4*>(buf + (k * kCeSThreads + threadIdx.x) * 8);
      const float x_label = to_f32<T>(buf[label]);
      float m = -INFINITY;
#pragma unroll
      for (int k = 0; k < G; ++k) {
        const uint32_t u[4] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack2<T>(u[e]);
          m = fmaxf(m, fmaxf(f.x, f.y));
        }
      }
      m = block_reduce_1024<true>(m, red_m);
      const float neg_m2 = -m * kLog2e;
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < G; ++k) {
        const uint32_t u[4] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack2<T>(u[e]);
          s += ce_exp2(fmaf(f.x, kLog2e, neg_m2)) + ce_exp2(fmaf(f.y, kLog2e, neg_m2));
        }
      }
      s = block_reduce_1024<false>(s, red_s);
      if (threadIdx.x == 0) nll[r] = (m + log2f(s) * kLn2) - x_label;
      if (write_grad) {
        const float scale = w / s;
#pragma unroll
        for (int k = 0; k < G; ++k) {
          const uint32_t u[4] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w};
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = unpack2<T>(u[e]);
            o[e] = pack2<T>(ce_exp2(fmaf(f.x, kLog2e, neg_m2)) * scale, ce_exp2(fmaf(f.y, kLog2e, neg_m2)) * scale);
          }
          *reinterpret_cast<uint4*>(buf + (k * kCeSThreads + threadIdx.x) * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        __syncthreads();                                     // the label's element has been written by its owner
        if (threadIdx.x == 0)                                // softmax - 1 at the label, rounded once from fp32
          buf[label] = from_f32<T>(fmaf(ce_exp2(fmaf(x_label, kLog2e, neg_m2)), scale, -w));
        ptx::fence_proxy_async();                            // generic-proxy writes -> visible to the bulk store
      }
    }
    __syncthreads();                                         // every thread is done with buffer b (and has fenced its writes)
    if (threadIdx.x == 0) {
      if (write_grad && w != 0.f) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          bulk_store_1d(reinterpret_cast<uint8_t*>(grow) + c * (ROW_BYTES / 4), smem + (size_t)b * ROW_BYTES + c * (ROW_BYTES / 4), ROW_BYTES / 4);
      }
      ptx::tma_store_commit();
      // row i + 2 goes into the buffer row i - 1 was stored from: that store (the group before the one just committed)
      // must have finished READING shared memory
      if (i + kCeSRing - 1 < mine) {
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        request(i + kCeSRing - 1);
      }
    }
  }
  if (threadIdx.x == 0) ptx::tma_store_wait_all();           // the gradient rows are in global memory when the kernel ends
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
  if (k <= 0 || k > 4) return SVAE_ERR_UNSUPPORTED;
  const int sms = sm_count_of_current_device();
  const unsigned grid = (unsigned)(rows < sms ? rows : sms);
#define SVAE_CE16S(KK)                                                                                              \
  case KK: {                                                                                                        \
    const size_t smem = (size_t)kCeSRing * 8192 * KK * sizeof(T) + kCeSRing * 8 + 2 * 64 * sizeof(float) + 128;     \
    auto kern = vocab_ce16s_kernel<T, KK>;                                                                          \
    SVAE_CONFIGURE_SMEM(kern, (int)smem);                                                                           \
    kern<<<grid, kCeSThreads, smem, st>>>((T*)logits, ld, labels, weight, nll, write_grad, rows);                   \
  } break
  switch (k) { SVAE_CE16S(1); SVAE_CE16S(2); SVAE_CE16S(3); SVAE_CE16S(4); }
#undef SVAE_CE16S
  SVAE_CUDA_CHECK(cudaGetLastError());
  return SVAE_OK;
}

// the register-resident variant (two CTAs of 512 threads per SM, one row each); kept as the cross-check of the
// streamed kernel (tests/test_gpu_fused_ce.py)
template <typename T>
static int launch_ce16_rows(int k, void* logits, int64_t ld, const int64_t* labels, const float* weight, float* nll, int write_grad,
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
  const bool streamed = !(write_grad & 2);      // bit 1 of write_grad: the register-resident variant (cross-check)
  write_grad &= 1;
  if (dtype == SVAE_DTYPE_BF16)
    return streamed ? launch_ce16<__nv_bfloat16>(k, logits, ld, labels, weight, nll, write_grad, rows, st)
                    : launch_ce16_rows<__nv_bfloat16>(k, logits, ld, labels, weight, nll, write_grad, rows, st);
  if (dtype == SVAE_DTYPE_F16)
    return streamed ? launch_ce16<__half>(k, logits, ld, labels, weight, nll, write_grad, rows, st)
                    : launch_ce16_rows<__half>(k, logits, ld, labels, weight, nll, write_grad, rows, st);
  SVAE_REQUIRE(false, SVAE_ERR_INVALID, "svae_vocab_ce: dtype %d", dtype);
}
