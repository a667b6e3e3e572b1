// libsvae_b200: error plumbing, version/device queries and the host-side block-layout / LUT builder.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

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

int sm_count_of_current_device() {
  static std::atomic<int> cache[128];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 128) return 148;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

// ---- per-kernel timing ----------------------------------------------------------------------------
struct ProfRecord {
  const char* name;
  cudaEvent_t a, b;
};
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRecord> g_prof;

ScopedKernelTimer::ScopedKernelTimer(const char* name, cudaStream_t s) : slot(-1), st(s) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRecord r;
  r.name = name;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, st);
  g_prof.push_back(r);
  slot = (int)g_prof.size() - 1;
}

ScopedKernelTimer::~ScopedKernelTimer() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (slot < (int)g_prof.size()) cudaEventRecord(g_prof[slot].b, st);
}

}  // namespace svae

using namespace svae;

extern "C" void svae_profile_begin(void) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.clear();
  g_prof_on = true;
}

// Stops recording, waits for the recorded events and writes {"kernel": {"launches": n, "ms": total}, ...}.
extern "C" int svae_profile_end(char* buf, size_t n) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = false;
  std::map<std::string, std::pair<long, double>> agg;
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      auto& e = agg[r.name];
      e.first += 1;
      e.second += ms;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.clear();
  std::string out = "{";
  bool first = true;
  for (auto& kv : agg) {
    char tmp[256];
    snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"launches\": %ld, \"ms\": %.6f}", first ? "" : ", ", kv.first.c_str(),
             kv.second.first, kv.second.second);
    out += tmp;
    first = false;
  }
  out += "}";
  SVAE_REQUIRE(buf && out.size() + 1 <= n, SVAE_ERR_INVALID, "svae_profile_end: buffer of %zu bytes too small (%zu needed)",
               n, out.size() + 1);
  memcpy(buf, out.c_str(), out.size() + 1);
  return SVAE_OK;
}

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
