// Error plumbing, launch accounting and host-side TMA descriptor encoding.
#include <stdarg.h>
#include <stdlib.h>
#include <atomic>
#include <mutex>
#include <unordered_map>
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"

namespace vlg {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return VLG_ERR_CUDA;
}

void count_launches(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VLG_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
  return dev;
}

int num_sms() {          // of the CURRENT device (function attributes and SM counts are per device)
  static std::atomic<int> n[kMaxDevices];
  const int dev = current_device();
  int v = n[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    n[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// Encoded descriptors are pure functions of their arguments: the Python-driven step API re-launches the same kernels on
// the same buffers every position (7 maps per layer-tail launch), so they are cached instead of re-encoded
// (cuTensorMapEncodeTiled is ~1.5 us of host time each).
namespace {
struct TmapKey {
  const void* base; uint64_t rows, cols, stride; uint32_t box_rows, box_cols; int sw;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && stride == o.stride && box_rows == o.box_rows &&
           box_cols == o.box_cols && sw == o.sw;
  }
};
struct TmapHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = reinterpret_cast<uintptr_t>(k.base) * 0x9E3779B97F4A7C15ull;
    for (uint64_t v : {k.rows, k.cols, k.stride, (uint64_t)k.box_rows << 32 | k.box_cols, (uint64_t)k.sw})
      h = (h ^ v) * 0x9E3779B97F4A7C15ull + (h >> 29);
    return (size_t)h;
  }
};
std::mutex g_tmap_mu;
std::unordered_map<TmapKey, TensorMap, TmapHash> g_tmap_cache;
}  // namespace

int make_tmap_bf16_2d(TensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                      uint32_t box_rows, uint32_t box_cols, bool swizzle128) {
  const TmapKey key{base, rows, cols, row_stride_bytes, box_rows, box_cols, swizzle128 ? 1 : 0};
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) { *out = it->second; return VLG_OK; }
  }
  auto fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return VLG_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base),
                  gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rows=%llu cols=%llu stride=%llu box=%ux%u", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)row_stride_bytes, box_rows,
              box_cols);
    return VLG_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmap_cache.size() > 8192) g_tmap_cache.clear();
    g_tmap_cache.emplace(key, *out);
  }
  return VLG_OK;
}

}  // namespace vlg

extern "C" {
int vlg_abi_version(void) { return VLG_ABI_VERSION; }
const char* vlg_last_error(void) { return vlg::g_err; }
uint64_t vlg_launch_count(void) { return vlg::g_launches.load(); }
}
