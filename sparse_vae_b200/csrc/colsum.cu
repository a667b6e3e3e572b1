// Column sums of a [rows, n] activation-gradient matrix: the bias gradient of every nn.Linear of the decoder blocks
// (reference core/attention.py:33-39 q/k/v/output_linear, core/transformer_layer.py:20-24 ffn) in the backward pass.
// ATen's generic reduction reaches ~1.6 TB/s on [65536, 512] bf16; this is a plain two-stage column reduction:
// stage 1, block = 32 sixteen-byte column vectors x 8 row lanes walking a slab of rows (coalesced 512-byte
// segments), fp32 accumulation, partial[slab][n]; stage 2 sums the slabs in a fixed order (deterministic) -- either a
// second small launch or, given a zeroed counter per column group, the last block of the group to finish (one launch).
// HBM-bound: rows * n * sizeof(T) bytes read.
#include "common.cuh"

namespace svae {

template <typename T> struct Vec8;      // 8 consecutive elements (16 bytes for 16-bit types, 32 for fp32)
template <> struct Vec8<float> {
  struct Raw { float4 a, b; };
  static __device__ __forceinline__ Raw load(const float* p) {
    return Raw{*reinterpret_cast<const float4*>(p), *reinterpret_cast<const float4*>(p + 4)};
  }
  static __device__ __forceinline__ void add(const Raw& r, float (&acc)[8]) {
    acc[0] += r.a.x; acc[1] += r.a.y; acc[2] += r.a.z; acc[3] += r.a.w; acc[4] += r.b.x; acc[5] += r.b.y; acc[6] += r.b.z; acc[7] += r.b.w;
  }
};
template <> struct Vec8<__nv_bfloat16> {
  using Raw = uint4;
  static __device__ __forceinline__ Raw load(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
  static __device__ __forceinline__ void add(const Raw& t, float (&acc)[8]) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
      acc[2 * i] += f.x; acc[2 * i + 1] += f.y;
    }
  }
};
template <> struct Vec8<__half> {
  using Raw = uint4;
  static __device__ __forceinline__ Raw load(const __half* p) { return *reinterpret_cast<const uint4*>(p); }
  static __device__ __forceinline__ void add(const Raw& t, float (&acc)[8]) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      acc[2 * i] += f.x; acc[2 * i + 1] += f.y;
    }
  }
};

template <typename T>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const T* __restrict__ x, int64_t ld, int64_t rows, int n,
                                                             float* __restrict__ partial, unsigned* __restrict__ counters,
                                                             float* __restrict__ out) {
  __shared__ float red[8][32][9];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int vec = blockIdx.x * 32 + cx;              // 8-column vector index
  const bool ok = vec * 8 < n;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (ok) {
    const int64_t step = (int64_t)gridDim.y * 8;
    int64_t r = (int64_t)blockIdx.y * 8 + ry;
#pragma unroll 2
    for (; r + 3 * step < rows; r += 4 * step) {     // four independent loads in flight, added in row order
      typename Vec8<T>::Raw a = Vec8<T>::load(x + r * ld + vec * 8), b = Vec8<T>::load(x + (r + step) * ld + vec * 8);
      typename Vec8<T>::Raw c = Vec8<T>::load(x + (r + 2 * step) * ld + vec * 8), d = Vec8<T>::load(x + (r + 3 * step) * ld + vec * 8);
      Vec8<T>::add(a, acc); Vec8<T>::add(b, acc); Vec8<T>::add(c, acc); Vec8<T>::add(d, acc);
    }
    for (; r < rows; r += step) Vec8<T>::add(Vec8<T>::load(x + r * ld + vec * 8), acc);
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
      partial[(int64_t)blockIdx.y * n + vec * 8 + e] = t;
    }
  }
  if (counters == nullptr) return;                   // two-launch mode: colsum_final_kernel sums the slabs
  // single-launch mode: the LAST block of this column group to finish sums the slabs, always in slab order, so the
  // result does not depend on which block that is; it leaves the counter at zero for the next launch
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
      out[vec * 8 + e] = t;
    }
  }
  if (threadIdx.x == 0) counters[blockIdx.x] = 0u;
}

__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ partial, int slabs, int n,
                                                           float* __restrict__ out) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float s = 0.f;
  if (c < n)
    for (int b = ry; b < slabs; b += 8) s += partial[(int64_t)b * n + c];
  red[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && c < n) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][cx];
    out[c] = t;
  }
}

static int colsum_slabs(int64_t rows, int n) {
  const int col_blocks = (n / 8 + 31) / 32;
  int slabs = (148 * 4 + col_blocks - 1) / col_blocks;
  const int64_t max_slabs = (rows + 63) / 64;        // at least 8 rows per row lane
  if (slabs > max_slabs) slabs = (int)(max_slabs > 0 ? max_slabs : 1);
  return slabs;
}

}  // namespace svae

using namespace svae;

extern "C" int64_t svae_colsum_workspace_floats(int64_t rows, int32_t n) { return (int64_t)colsum_slabs(rows, n) * n; }

extern "C" int32_t svae_colsum_counters(int32_t n) { return (n / 8 + 31) / 32; }

extern "C" int svae_colsum(const void* x, int32_t dtype, int64_t rows, int32_t n, int64_t ld, float* out, float* workspace,
                           int64_t workspace_floats, uint32_t* counters, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SVAE_REQUIRE(x && out && workspace && rows >= 0, SVAE_ERR_INVALID, "svae_colsum: null argument");
  SVAE_REQUIRE(n > 0 && n % 8 == 0 && ld >= n && ld % 8 == 0, SVAE_ERR_INVALID, "svae_colsum: n and ld must be multiples of 8");
  SVAE_REQUIRE((reinterpret_cast<uintptr_t>(x) & (dtype == SVAE_DTYPE_F32 ? 31 : 15)) == 0, SVAE_ERR_INVALID, "svae_colsum: misaligned input");
  const int slabs = colsum_slabs(rows, n);
  SVAE_REQUIRE(workspace_floats >= (int64_t)slabs * n, SVAE_ERR_INVALID, "svae_colsum: workspace too small");
  dim3 grid((n / 8 + 31) / 32, slabs);
  ScopedKernelTimer timer("colsum", st);
  if (dtype == SVAE_DTYPE_F32) colsum_partial_kernel<float><<<grid, 256, 0, st>>>((const float*)x, ld, rows, n, workspace, counters, out);
  else if (dtype == SVAE_DTYPE_BF16) colsum_partial_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, ld, rows, n, workspace, counters, out);
  else if (dtype == SVAE_DTYPE_F16) colsum_partial_kernel<__half><<<grid, 256, 0, st>>>((const __half*)x, ld, rows, n, workspace, counters, out);
  else SVAE_REQUIRE(false, SVAE_ERR_INVALID, "svae_colsum: dtype %d", dtype);
  SVAE_CUDA_CHECK(cudaGetLastError());
  if (counters == nullptr) {
    colsum_final_kernel<<<(n + 31) / 32, 256, 0, st>>>(workspace, slabs, n, out);
    SVAE_CUDA_CHECK(cudaGetLastError());
  }
  return SVAE_OK;
}
