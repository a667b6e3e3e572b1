// libsvae_b200: error plumbing, version/device queries and the host-side block-layout / LUT builder.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace svae {

static thread_local char g_last_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return SVAE_ERR_CUDA;
}

}  // namespace svae

using namespace svae;

extern "C" int svae_abi_version(void) { return SVAE_ABI_VERSION; }

extern "C" const char* svae_last_error(void) { return g_last_error; }

extern "C" int svae_device_check(void) {
  int dev = 0;
  SVAE_CUDA_CHECK(cudaGetDevice(&dev));
  int major = 0;
  SVAE_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  SVAE_REQUIRE(major == 10, SVAE_ERR_DEVICE, "device %d has compute capability %d.x; libsvae_b200 is built for sm_100a only",
               dev, major);
  return SVAE_OK;
}

// ---- layout -------------------------------------------------------------------------------------
// Restates SparseAttention.get_master_layout()[..., :nb, :nb] (reference core/sparse_attention.py:38-59)
// as a predicate per (row, col) block instead of diagonal fills on a 3600x3600 master matrix.
static inline bool layout_bit(int r, int c, int left, int right_ctx, int cls) {
  if (cls && c == 0) return true;
  int d = r - c;                       // >0: sub-diagonal, <0: super-diagonal
  if (d >= 0) return d < left;         // range(left_context)
  return -d < right_ctx;               // range(1, right_context)
}

static inline void contexts(int window, int causal, int* left, int* right_ctx) {
  int sides = causal ? 1 : 2;
  *left = window / sides + window % sides;
  *right_ctx = window - *left;
}

extern "C" int64_t svae_layout_nnz(int32_t nb, int32_t window, int32_t causal, int32_t include_cls) {
  if (nb < 0 || window < 0) return -1;
  int left, right_ctx;
  contexts(window, causal, &left, &right_ctx);
  int64_t n = 0;
  for (int r = 0; r < nb; ++r)
    for (int c = 0; c < nb; ++c) n += layout_bit(r, c, left, right_ctx, include_cls) ? 1 : 0;
  return n;
}

extern "C" int svae_layout_build(int32_t nb, int32_t window, int32_t causal, int32_t include_cls, int32_t num_heads,
                                 int64_t* layout, int32_t* row_ptr, int32_t* col_idx, int32_t* colT_ptr,
                                 int32_t* rowT_idx) {
  SVAE_REQUIRE(nb >= 0 && window >= 0 && num_heads >= 0, SVAE_ERR_INVALID,
               "svae_layout_build: negative size (nb=%d window=%d heads=%d)", nb, window, num_heads);
  int left, right_ctx;
  contexts(window, causal, &left, &right_ctx);
  if (layout) {
    for (int r = 0; r < nb; ++r)
      for (int c = 0; c < nb; ++c) layout[(int64_t)r * nb + c] = layout_bit(r, c, left, right_ctx, include_cls) ? 1 : 0;
    for (int h = 1; h < num_heads; ++h) memcpy(layout + (int64_t)h * nb * nb, layout, sizeof(int64_t) * nb * nb);
  }
  if (row_ptr || col_idx) {
    int32_t n = 0;
    for (int r = 0; r < nb; ++r) {
      if (row_ptr) row_ptr[r] = n;
      for (int c = 0; c < nb; ++c)
        if (layout_bit(r, c, left, right_ctx, include_cls)) {
          if (col_idx) col_idx[n] = c;
          ++n;
        }
    }
    if (row_ptr) row_ptr[nb] = n;
  }
  if (colT_ptr || rowT_idx) {
    int32_t n = 0;
    for (int c = 0; c < nb; ++c) {
      if (colT_ptr) colT_ptr[c] = n;
      for (int r = 0; r < nb; ++r)
        if (layout_bit(r, c, left, right_ctx, include_cls)) {
          if (rowT_idx) rowT_idx[n] = r;
          ++n;
        }
    }
    if (colT_ptr) colT_ptr[nb] = n;
  }
  return SVAE_OK;
}
