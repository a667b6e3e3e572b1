// LayerNorm forward / backward for the pre-LN decoder blocks that call the sparse-attention hot path
// (reference core/transformer_layer.py:17-24,44-61: attn_layer_norm, ffn_layer_norm; core/transformer_language_model.py
// :55-63: the LayerNorm of the output head).  SURVEY.md section 2.1 #11 lists the block as the "next" fusion row.
//
// Why a kernel of our own: under autocast the reference's LayerNorm runs in fp32 and every consuming Linear casts
// its fp32 output to 16 bit again (three casts for q/k/v), and ATen's gamma/beta backward at [65536, 512] is a
// 16-CTA kernel (419 us, profiles/r01_launches_summary.md).  Here: one warp per row, the row lives in registers,
// fp32 statistics, output written directly in the consumer's dtype (bit-identical to casting the fp32 result), and
// the backward computes dx AND the per-CTA partial dgamma/dbeta in the same single pass over x and dy; a second
// tiny kernel sums the partials in a fixed order (deterministic, no atomics).
// HBM-bound: forward reads rows*n*sx + writes rows*n*sy bytes; backward reads rows*n*(sx+sy), writes rows*n*sx.
#include "common.cuh"

namespace svae {

constexpr int kLnWarps = 8;
constexpr int kLnThreads = kLnWarps * 32;

template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a); t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
};
template <> struct Vec4<__half> {
  static __device__ __forceinline__ void load(const __half* p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ void store(__half* p, const float (&v)[4]) {
    __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a); t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// lane l owns elements (i*32 + l)*4 .. +3 for i < VPT  (n = 128 * VPT)
template <typename TX, typename TY, int VPT>
__global__ void __launch_bounds__(kLnThreads)
layernorm_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     TY* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int64_t rows, float eps) {
  constexpr int N = 128 * VPT;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kLnWarps;
  float g[VPT][4], bt[VPT][4];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    Vec4<float>::load(gamma + (i * 32 + lane) * 4, g[i]);
    if (beta) Vec4<float>::load(beta + (i * 32 + lane) * 4, bt[i]);
    else bt[i][0] = bt[i][1] = bt[i][2] = bt[i][3] = 0.f;
  }
  for (int64_t r = warp0; r < rows; r += nwarps) {
    float v[VPT][4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      Vec4<TX>::load(x + r * N + (i * 32 + lane) * 4, v[i]);
      s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
    }
    const float mean = warp_sum(s) * (1.0f / N);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float d = v[i][e] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / N) + eps);
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = fmaf((v[i][e] - mean) * rstd, g[i][e], bt[i][e]);
      Vec4<TY>::store(y + r * N + (i * 32 + lane) * 4, o);
    }
    if (lane == 0) { mean_out[r] = mean; rstd_out[r] = rstd; }
  }
}

// Token-by-token decoding (core/decode.py): the residual stream update and the next LayerNorm in one pass.
// x (fp32, in place) += h; y = LayerNorm(x).  Same arithmetic as the separate `x + h` (fp32 add of the promoted
// 16-bit branch output) followed by layernorm_fwd_kernel.  x_out may be x itself (decoding) or a new buffer (training,
// where the old x is still needed by the backward pass of the norm it fed); mean / rstd are written when given.
template <typename TH, typename TY, int VPT>
__global__ void __launch_bounds__(kLnThreads)
residual_layernorm_kernel(const float* x, const TH* __restrict__ h, const float* __restrict__ gamma,
                          const float* __restrict__ beta, TY* __restrict__ y, float* x_out, float* __restrict__ mean_out,
                          float* __restrict__ rstd_out, int64_t rows, float eps) {
  constexpr int N = 128 * VPT;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kLnWarps;
  for (int64_t r = warp0; r < rows; r += nwarps) {
    float v[VPT][4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      float hv[4];
      Vec4<float>::load(x + r * N + (i * 32 + lane) * 4, v[i]);
      Vec4<TH>::load(h + r * N + (i * 32 + lane) * 4, hv);
#pragma unroll
      for (int e = 0; e < 4; ++e) v[i][e] += hv[e];
      Vec4<float>::store(x_out + r * N + (i * 32 + lane) * 4, v[i]);
      s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
    }
    const float mean = warp_sum(s) * (1.0f / N);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float d = v[i][e] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / N) + eps);
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      float g[4], bt[4], o[4];
      Vec4<float>::load(gamma + (i * 32 + lane) * 4, g);
      if (beta) Vec4<float>::load(beta + (i * 32 + lane) * 4, bt);
      else bt[0] = bt[1] = bt[2] = bt[3] = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = fmaf((v[i][e] - mean) * rstd, g[e], bt[e]);
      Vec4<TY>::store(y + r * N + (i * 32 + lane) * 4, o);
    }
    if (mean_out && lane == 0) { mean_out[r] = mean; rstd_out[r] = rstd; }
  }
}

// dx = rstd * (dy*gamma - mean(dy*gamma) - xhat * mean(dy*gamma*xhat)) ; partial[block] = {sum dy*xhat, sum dy}
template <typename TX, typename TY, int VPT>
__global__ void __launch_bounds__(kLnThreads)
layernorm_bwd_kernel(const TY* __restrict__ dy, const TX* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in, TX* __restrict__ dx,
                     const TX* __restrict__ dres, TY* __restrict__ dx_low, float* __restrict__ partial, int64_t rows) {
  constexpr int N = 128 * VPT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warp0 = (int64_t)blockIdx.x * kLnWarps + warp;
  const int64_t nwarps = (int64_t)gridDim.x * kLnWarps;
  float g[VPT][4], dg[VPT][4], db[VPT][4];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    Vec4<float>::load(gamma + (i * 32 + lane) * 4, g[i]);
#pragma unroll
    for (int e = 0; e < 4; ++e) dg[i][e] = db[i][e] = 0.f;
  }
  for (int64_t r = warp0; r < rows; r += nwarps) {
    const float mean = mean_in[r], rstd = rstd_in[r];
    float xh[VPT][4], d[VPT][4];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      Vec4<TX>::load(x + r * N + (i * 32 + lane) * 4, xh[i]);
      Vec4<TY>::load(dy + r * N + (i * 32 + lane) * 4, d[i]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        xh[i][e] = (xh[i][e] - mean) * rstd;
        dg[i][e] = fmaf(d[i][e], xh[i][e], dg[i][e]);
        db[i][e] += d[i][e];
        d[i][e] *= g[i][e];
        c1 += d[i][e];
        c2 = fmaf(d[i][e], xh[i][e], c2);
      }
    }
    c1 = warp_sum(c1) * (1.0f / N);
    c2 = warp_sum(c2) * (1.0f / N);
    if (dx) {
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = rstd * (d[i][e] - c1 - xh[i][e] * c2);
        if (dres) {                                  // gradient arriving through the residual path around this norm
          float a[4];
          Vec4<TX>::load(dres + r * N + (i * 32 + lane) * 4, a);
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] += a[e];
        }
        Vec4<TX>::store(dx + r * N + (i * 32 + lane) * 4, o);
        if (dx_low) Vec4<TY>::store(dx_low + r * N + (i * 32 + lane) * 4, o);     // the same gradient for a 16-bit branch
      }
    }
  }
  // cross-warp reduction of the column sums in a fixed order, one [2, N] slab per CTA (dgamma, then dbeta)
  __shared__ float red[kLnWarps][N];
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    if (which) __syncthreads();
#pragma unroll
    for (int i = 0; i < VPT; ++i) Vec4<float>::store(red[warp] + (i * 32 + lane) * 4, which ? db[i] : dg[i]);
    __syncthreads();
    for (int c = threadIdx.x; c < N; c += kLnThreads) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kLnWarps; ++w) s += red[w][c];
      partial[(int64_t)blockIdx.x * (2 * N) + which * N + c] = s;
    }
  }
}

// dgamma / dbeta = column sums of the per-CTA partials [nblocks, 2n]; block = 32 columns x 8 row groups, fixed order
__global__ void __launch_bounds__(256) layernorm_param_grad_kernel(const float* __restrict__ partial, int nblocks, int n2,
                                                                   float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float s = 0.f;
  if (c < n2)
    for (int b = ry; b < nblocks; b += 8) s += partial[(int64_t)b * n2 + c];
  red[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && c < n2) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][cx];
    const int n = n2 >> 1;
    if (c < n) { if (dgamma) dgamma[c] = t; }
    else if (dbeta) dbeta[c - n] = t;
  }
}

static int ln_grid(int64_t rows) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
  }
  const int64_t want = (rows + kLnWarps - 1) / kLnWarps;
  const int64_t cap = (int64_t)sms * 4;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

template <typename TX, typename TY>
static int launch_fwd(int vpt, const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                      int64_t rows, float eps, cudaStream_t st) {
  const int grid = ln_grid(rows);
#define SVAE_LN_FWD(V)                                                                                         \
  case V: layernorm_fwd_kernel<TX, TY, V><<<grid, kLnThreads, 0, st>>>((const TX*)x, gamma, beta, (TY*)y, mean, rstd, rows, eps); break
  switch (vpt) { SVAE_LN_FWD(1); SVAE_LN_FWD(2); SVAE_LN_FWD(4); SVAE_LN_FWD(8); default: return SVAE_ERR_UNSUPPORTED; }
#undef SVAE_LN_FWD
  SVAE_CUDA_CHECK(cudaGetLastError());
  return SVAE_OK;
}

template <typename TX, typename TY>
static int launch_bwd(int vpt, const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                      void* dx, const void* dres, void* dx_low, float* partial, int grid, int64_t rows, cudaStream_t st) {
#define SVAE_LN_BWD(V)                                                                                         \
  case V: layernorm_bwd_kernel<TX, TY, V><<<grid, kLnThreads, 0, st>>>((const TY*)dy, (const TX*)x, gamma, mean, rstd, (TX*)dx, (const TX*)dres, (TY*)dx_low, partial, rows); break
  switch (vpt) { SVAE_LN_BWD(1); SVAE_LN_BWD(2); SVAE_LN_BWD(4); SVAE_LN_BWD(8); default: return SVAE_ERR_UNSUPPORTED; }
#undef SVAE_LN_BWD
  SVAE_CUDA_CHECK(cudaGetLastError());
  return SVAE_OK;
}

template <typename TH, typename TY>
static int launch_residual(int vpt, const float* x, const void* h, const float* gamma, const float* beta, void* y,
                           float* x_out, float* mean, float* rstd, int64_t rows, float eps, cudaStream_t st) {
  const int grid = ln_grid(rows);
#define SVAE_LN_RES(V)                                                                                         \
  case V: residual_layernorm_kernel<TH, TY, V><<<grid, kLnThreads, 0, st>>>(x, (const TH*)h, gamma, beta, (TY*)y, x_out, mean, rstd, rows, eps); break
  switch (vpt) { SVAE_LN_RES(1); SVAE_LN_RES(2); SVAE_LN_RES(4); SVAE_LN_RES(8); default: return SVAE_ERR_UNSUPPORTED; }
#undef SVAE_LN_RES
  SVAE_CUDA_CHECK(cudaGetLastError());
  return SVAE_OK;
}

static bool ln_shape_ok(int32_t n) { return n == 128 || n == 256 || n == 512 || n == 1024; }

}  // namespace svae

using namespace svae;

extern "C" int svae_layernorm_supported(int32_t n) { return ln_shape_ok(n) ? 1 : 0; }

extern "C" int64_t svae_layernorm_bwd_workspace_floats(int64_t rows, int32_t n) { return (int64_t)ln_grid(rows) * 2 * n; }

#define SVAE_LN_DISPATCH(FN, ...)                                                                     \
  do {                                                                                                \
    if (x_dtype == SVAE_DTYPE_F32 && y_dtype == SVAE_DTYPE_F32) return FN<float, float>(__VA_ARGS__); \
    if (x_dtype == SVAE_DTYPE_F32 && y_dtype == SVAE_DTYPE_BF16) return FN<float, __nv_bfloat16>(__VA_ARGS__); \
    if (x_dtype == SVAE_DTYPE_F32 && y_dtype == SVAE_DTYPE_F16) return FN<float, __half>(__VA_ARGS__); \
    if (x_dtype == SVAE_DTYPE_BF16 && y_dtype == SVAE_DTYPE_BF16) return FN<__nv_bfloat16, __nv_bfloat16>(__VA_ARGS__); \
    if (x_dtype == SVAE_DTYPE_BF16 && y_dtype == SVAE_DTYPE_F32) return FN<__nv_bfloat16, float>(__VA_ARGS__); \
    if (x_dtype == SVAE_DTYPE_F16 && y_dtype == SVAE_DTYPE_F16) return FN<__half, __half>(__VA_ARGS__); \
    if (x_dtype == SVAE_DTYPE_F16 && y_dtype == SVAE_DTYPE_F32) return FN<__half, float>(__VA_ARGS__); \
  } while (0)

extern "C" int svae_layernorm_fwd(const void* x, int32_t x_dtype, const float* gamma, const float* beta, int64_t rows,
                                  int32_t n, float eps, void* y, int32_t y_dtype, float* mean, float* rstd, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SVAE_REQUIRE(x && gamma && y && mean && rstd && rows >= 0, SVAE_ERR_INVALID, "svae_layernorm_fwd: null argument");
  SVAE_REQUIRE(ln_shape_ok(n), SVAE_ERR_UNSUPPORTED, "svae_layernorm_fwd: width %d not in {128, 256, 512, 1024}", n);
  if (rows == 0) return SVAE_OK;
  ScopedKernelTimer timer("layernorm_fwd", st);
  SVAE_LN_DISPATCH(launch_fwd, n / 128, x, gamma, beta, y, mean, rstd, rows, eps, st);
  SVAE_REQUIRE(false, SVAE_ERR_UNSUPPORTED, "svae_layernorm_fwd: dtype pair (%d -> %d) not supported", x_dtype, y_dtype);
}

extern "C" int svae_layernorm_bwd(const void* dy, int32_t y_dtype, const void* x, int32_t x_dtype, const float* gamma,
                                  const float* mean, const float* rstd, int64_t rows, int32_t n, void* dx,
                                  const void* dx_residual, void* dx_low, float* dgamma, float* dbeta, float* workspace,
                                  int64_t workspace_floats, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SVAE_REQUIRE(dy && x && gamma && mean && rstd && workspace && rows >= 0, SVAE_ERR_INVALID, "svae_layernorm_bwd: null argument");
  SVAE_REQUIRE(ln_shape_ok(n), SVAE_ERR_UNSUPPORTED, "svae_layernorm_bwd: width %d not in {128, 256, 512, 1024}", n);
  const int grid = ln_grid(rows);
  SVAE_REQUIRE(workspace_floats >= (int64_t)grid * 2 * n, SVAE_ERR_INVALID, "svae_layernorm_bwd: workspace too small");
  ScopedKernelTimer timer("layernorm_bwd", st);
  auto run = [&]() -> int {
    SVAE_LN_DISPATCH(launch_bwd, n / 128, dy, x, gamma, mean, rstd, dx, dx_residual, dx_low, workspace, grid, rows, st);
    SVAE_REQUIRE(false, SVAE_ERR_UNSUPPORTED, "svae_layernorm_bwd: dtype pair (%d -> %d) not supported", x_dtype, y_dtype);
  };
  int rc = run();
  if (rc) return rc;
  if (dgamma || dbeta) {
    layernorm_param_grad_kernel<<<(2 * n + 31) / 32, 256, 0, st>>>(workspace, rows > 0 ? grid : 0, 2 * n, dgamma, dbeta);
    SVAE_CUDA_CHECK(cudaGetLastError());
  }
  return SVAE_OK;
}

extern "C" int svae_residual_layernorm(const float* x, const void* h, int32_t h_dtype, const float* gamma, const float* beta,
                                       int64_t rows, int32_t n, float eps, void* y, int32_t y_dtype, float* x_out, float* mean,
                                       float* rstd, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SVAE_REQUIRE(x && h && gamma && y && x_out && rows >= 0, SVAE_ERR_INVALID, "svae_residual_layernorm: null argument");
  SVAE_REQUIRE((mean == nullptr) == (rstd == nullptr), SVAE_ERR_INVALID, "svae_residual_layernorm: mean and rstd go together");
  SVAE_REQUIRE(ln_shape_ok(n), SVAE_ERR_UNSUPPORTED, "svae_residual_layernorm: width %d not in {128, 256, 512, 1024}", n);
  SVAE_REQUIRE(h_dtype == y_dtype && (h_dtype == SVAE_DTYPE_F16 || h_dtype == SVAE_DTYPE_BF16 || h_dtype == SVAE_DTYPE_F32),
               SVAE_ERR_UNSUPPORTED, "svae_residual_layernorm: branch and output must share one dtype (%d, %d)", h_dtype, y_dtype);
  if (rows == 0) return SVAE_OK;
  ScopedKernelTimer timer("residual_layernorm", st);
  if (h_dtype == SVAE_DTYPE_F16)
    return launch_residual<__half, __half>(n / 128, x, h, gamma, beta, y, x_out, mean, rstd, rows, eps, st);
  if (h_dtype == SVAE_DTYPE_BF16)
    return launch_residual<__nv_bfloat16, __nv_bfloat16>(n / 128, x, h, gamma, beta, y, x_out, mean, rstd, rows, eps, st);
  return launch_residual<float, float>(n / 128, x, h, gamma, beta, y, x_out, mean, rstd, rows, eps, st);
}
