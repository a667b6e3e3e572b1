// Token-by-token decoding step of one sparse-attention layer (reference
// core/attention.py:60-100 with the KV cache of :107-142, driven by TransformerVAE.sample, transformer_vae.py:112-126).
//
// The reference runs ~25 launches per layer and token: rotary on q and k (offset = cache_index), two cache row
// writes, a block shift of the cache every 32 tokens, q k^T, scale, softmax, p v.  Here it is ONE launch whose
// position comes from a device counter, so the whole per-token model step can sit in a CUDA graph:
//   * rotary (fp32 tables, one rounding to T -- the arithmetic of rotary.cu's autocast mode) on the new q and k rows,
//   * k / v appended to the cache.  Cache layout [B, (window+1)*block, d_model]: slots [0, block) hold positions
//     0..block-1 (the global block, never evicted), the other window*block slots are a RING over positions >= block
//     (position p lives in slot block + (p - block) mod (window*block)).  While p < (window+1)*block this is
//     the reference's layout (slot == p); afterwards the reference shifts the window left by one block every `block`
//     tokens where this kernel wraps around -- the set of visible keys is the same:
//     block 0 plus key blocks max(1, b - window + 1) .. b of the current block b (the rows of get_master_layout),
//   * scores, fp32 softmax and p v for every head of one sample.
// One warp per (sample, head); 4 warps per CTA.  HBM-bound: reads the live part of both caches once,
// 2 * B * min(p+1, (window+1)*block) * d_model * sizeof(T) bytes per launch.
#include "common.cuh"

namespace svae {

template <typename T, int N> struct alignas(sizeof(T) * N) DPack { T v[N]; };

// position held by ring slot `slot` when the newest position is `pos`; -1 when the slot is empty or evicted
__device__ __forceinline__ int slot_position(int slot, int pos, int block, int window) {
  if (slot < block) return slot <= pos ? slot : -1;
  if (pos < block) return -1;
  const int ring = window * block;
  const int r = slot - block, m = pos - block;
  int back = (m - r) % ring;
  if (back < 0) back += ring;
  const int q = block + m - back;
  if (q < block) return -1;
  const int first_block = max(1, pos / block - (window - 1));
  return q >= first_block * block ? q : -1;
}

template <typename T, int DH>
__global__ void __launch_bounds__(128) decode_attn_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                           const T* __restrict__ v, const float* __restrict__ cos_t,
                                                           const float* __restrict__ sin_t, T* __restrict__ key_cache,
                                                           T* __restrict__ value_cache, T* __restrict__ out,
                                                           const int* __restrict__ pos_ptr, int B, int H, int window,
                                                           int block, int table_rows, int64_t in_stride, float scale) {
  constexpr int kVec = 16 / sizeof(T);              // elements per 16-byte load
  constexpr int kMaxSlots = 15 * 32;                // window <= 14
  __shared__ float s_q[4][DH];
  __shared__ float s_p[4][kMaxSlots];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x * 4 + warp;
  if (bh >= B * H) return;
  const int b = bh / H, h = bh - b * H;
  const int d_model = H * DH;
  const int pos = *pos_ptr;
  const int C = (window + 1) * block;
  const int slot = pos < block ? pos : block + (pos - block) % (window * block);
  const int trow = min(pos, table_rows - 1);

  // ---- rotary on the new q / k rows, append k / v ------------------------------------------------------------
  const size_t row = (size_t)b * in_stride + (size_t)h * DH;          // q / k / v rows may be slices of one [B, 3D] GEMM
  const size_t out_row = (size_t)b * d_model + (size_t)h * DH;
  T* kc = key_cache + ((size_t)b * C) * d_model + (size_t)h * DH;
  T* vc = value_cache + ((size_t)b * C) * d_model + (size_t)h * DH;
  for (int pr = lane; pr < DH / 2; pr += 32) {
    const float c = cos_t[(size_t)trow * (d_model / 2) + h * (DH / 2) + pr];
    const float s = sin_t[(size_t)trow * (d_model / 2) + h * (DH / 2) + pr];
    const float qe = to_f32<T>(q[row + 2 * pr]), qo = to_f32<T>(q[row + 2 * pr + 1]);
    const float ke = to_f32<T>(k[row + 2 * pr]), ko = to_f32<T>(k[row + 2 * pr + 1]);
    const T q0 = from_f32<T>(__fsub_rn(__fmul_rn(qe, c), __fmul_rn(qo, s)));
    const T q1 = from_f32<T>(__fadd_rn(__fmul_rn(qo, c), __fmul_rn(qe, s)));
    const T k0 = from_f32<T>(__fsub_rn(__fmul_rn(ke, c), __fmul_rn(ko, s)));
    const T k1 = from_f32<T>(__fadd_rn(__fmul_rn(ko, c), __fmul_rn(ke, s)));
    s_q[warp][2 * pr] = to_f32<T>(q0);
    s_q[warp][2 * pr + 1] = to_f32<T>(q1);
    kc[(size_t)slot * d_model + 2 * pr] = k0;
    kc[(size_t)slot * d_model + 2 * pr + 1] = k1;
    vc[(size_t)slot * d_model + 2 * pr] = v[row + 2 * pr];
    vc[(size_t)slot * d_model + 2 * pr + 1] = v[row + 2 * pr + 1];
  }
  __syncwarp();

  // ---- scores: one cache slot per lane and pass ----------------------------------------------------------------
  float row_max = -INFINITY;
  for (int s = lane; s < C; s += 32) {
    float score = -INFINITY;
    if (slot_position(s, pos, block, window) >= 0) {
      const T* kr = kc + (size_t)s * d_model;
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < DH / kVec; ++i) {
        const DPack<T, kVec> kv = *reinterpret_cast<const DPack<T, kVec>*>(kr + i * kVec);
#pragma unroll
        for (int e = 0; e < kVec; ++e) acc = fmaf(s_q[warp][i * kVec + e], to_f32<T>(kv.v[e]), acc);
      }
      score = acc * scale;
    }
    s_p[warp][s] = score;
    row_max = fmaxf(row_max, score);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) row_max = fmaxf(row_max, __shfl_xor_sync(0xffffffffu, row_max, o));
  float row_sum = 0.f;
  for (int s = lane; s < C; s += 32) {
    const float sc = s_p[warp][s];
    const float p = sc == -INFINITY ? 0.f : __expf(sc - row_max);      // the newest key is always live: row_max finite
    s_p[warp][s] = p;
    row_sum += p;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) row_sum += __shfl_xor_sync(0xffffffffu, row_sum, o);
  const float inv = 1.f / row_sum;
  __syncwarp();

  // ---- out = p v: lanes split the head's features, coalesced value rows -----------------------------------------
  constexpr int kPer = DH / 32;                     // features per lane (1 or 2)
  float acc[kPer] = {};
  const int live = min(pos + 1, C);                 // slots >= live are still empty (p == 0)
#pragma unroll 4
  for (int s = 0; s < live; ++s) {
    const float p = s_p[warp][s];
    if (p != 0.f) {
      const DPack<T, kPer> vv = *reinterpret_cast<const DPack<T, kPer>*>(vc + (size_t)s * d_model + lane * kPer);
#pragma unroll
      for (int e = 0; e < kPer; ++e) acc[e] = fmaf(p, to_f32<T>(vv.v[e]), acc[e]);
    }
  }
  DPack<T, kPer> o;
#pragma unroll
  for (int e = 0; e < kPer; ++e) o.v[e] = from_f32<T>(acc[e] * inv);
  *reinterpret_cast<DPack<T, kPer>*>(out + out_row + lane * kPer) = o;
}

template <typename T>
static int launch_decode(const void* q, const void* k, const void* v, const float* c, const float* s, void* kc, void* vc,
                         void* out, const int* pos, int B, int H, int Dh, int window, int block, int table_rows,
                         int64_t in_stride, float scale, cudaStream_t st) {
  const unsigned grid = (unsigned)((B * H + 3) / 4);
#define SVAE_DEC(DH)                                                                                                  \
  decode_attn_kernel<T, DH><<<grid, 128, 0, st>>>((const T*)q, (const T*)k, (const T*)v, c, s, (T*)kc, (T*)vc, (T*)out, \
                                                   pos, B, H, window, block, table_rows, in_stride, scale)
  if (Dh == 64) SVAE_DEC(64); else SVAE_DEC(32);
#undef SVAE_DEC
  SVAE_CUDA_CHECK(cudaGetLastError());
  return SVAE_OK;
}

}  // namespace svae

using namespace svae;

extern "C" int svae_decode_attn_supported(int32_t head_dim, int32_t window, int32_t block) {
  return (head_dim == 32 || head_dim == 64) && window >= 1 && window <= 14 && block == 32;
}

extern "C" int svae_decode_attn(const void* q, const void* k, const void* v, const float* cos_table,
                                const float* sin_table, void* key_cache, void* value_cache, void* out,
                                const int32_t* position, int32_t B, int32_t H, int32_t head_dim, int32_t window,
                                int32_t block, int32_t table_rows, int64_t in_stride, int32_t dtype, float scale,
                                void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SVAE_REQUIRE(q && k && v && cos_table && sin_table && key_cache && value_cache && out && position, SVAE_ERR_INVALID,
               "svae_decode_attn: null argument");
  SVAE_REQUIRE(svae_decode_attn_supported(head_dim, window, block), SVAE_ERR_UNSUPPORTED,
               "svae_decode_attn: head_dim %d / window %d / block %d not supported (head_dim 32|64, window 1..14, block 32)",
               head_dim, window, block);
  SVAE_REQUIRE(B >= 0 && H > 0 && table_rows > 0 && in_stride >= (int64_t)H * head_dim && in_stride % 8 == 0,
               SVAE_ERR_INVALID, "svae_decode_attn: bad sizes");
  const uintptr_t a = reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                      reinterpret_cast<uintptr_t>(key_cache) | reinterpret_cast<uintptr_t>(value_cache) |
                      reinterpret_cast<uintptr_t>(out);
  SVAE_REQUIRE((a & 15) == 0, SVAE_ERR_INVALID, "svae_decode_attn: tensors must be 16-byte aligned");
  if (B == 0) return SVAE_OK;
  ScopedKernelTimer timer("decode_attn", st);
  if (dtype == SVAE_DTYPE_F32)
    return launch_decode<float>(q, k, v, cos_table, sin_table, key_cache, value_cache, out, position, B, H, head_dim,
                                window, block, table_rows, in_stride, scale, st);
  if (dtype == SVAE_DTYPE_BF16)
    return launch_decode<__nv_bfloat16>(q, k, v, cos_table, sin_table, key_cache, value_cache, out, position, B, H,
                                        head_dim, window, block, table_rows, in_stride, scale, st);
  if (dtype == SVAE_DTYPE_F16)
    return launch_decode<__half>(q, k, v, cos_table, sin_table, key_cache, value_cache, out, position, B, H, head_dim,
                                 window, block, table_rows, in_stride, scale, st);
  SVAE_REQUIRE(false, SVAE_ERR_INVALID, "svae_decode_attn: dtype %d", dtype);
}
