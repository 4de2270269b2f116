#include "tma_host.h"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <unordered_map>

namespace modcr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

static std::atomic<unsigned long long> g_launches{0};
unsigned long long launch_count() { return g_launches.load(); }
void count_launches(int n) { g_launches.fetch_add((unsigned long long)n); }

// SM budget of the persistent kernels (0 = every SM).  While gradient all-reduces run on a side stream their CTAs occupy
// SMs; a persistent kernel that launches one CTA per SM with a static tile partition then has CTAs queued behind a
// collective for its whole duration (measured: + 3.3 ms per step at 8 GPUs).  The data-parallel reducer lowers the budget
// for the part of backward that overlaps communication, so collective and compute CTAs are co-resident.
static std::atomic<int> g_sm_limit{0};
int sm_limit() { return g_sm_limit.load(); }
int set_sm_limit(int sms) { return g_sm_limit.exchange(sms < 0 ? 0 : sms); }

int check_launch(const char* what) {
  g_launches.fetch_add(1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return MODCR_ERR_CUDA;
  }
  return MODCR_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
  });
  return fn;
}

struct Key {
  uint64_t v[9];
  bool operator==(const Key& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct KeyHash {
  size_t operator()(const Key& k) const {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 9; ++i) {
      h ^= k.v[i];
      h *= 1099511628211ull;
    }
    return (size_t)h;
  }
};
static std::mutex g_mu;
static std::unordered_map<Key, CUtensorMap, KeyHash> g_cache;

// The swizzle mode follows the inner box width: 128-byte rows -> SWIZZLE_128B, 64-byte rows -> SWIZZLE_64B.
static int encode(CUtensorMap* out, const void* ptr, int elem_bytes, int rank, const cuuint64_t* dims,
                  const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
    return MODCR_ERR_DRIVER;
  }
  if (((uintptr_t)ptr & 15) != 0) {
    set_error("tensor base %p is not 16-byte aligned", ptr);
    return MODCR_ERR_ALIGN;
  }
  for (int i = 0; i < rank - 1; ++i) {
    if (strides_bytes[i] % 16 != 0) {
      set_error("tensor stride %llu bytes is not a multiple of 16", (unsigned long long)strides_bytes[i]);
      return MODCR_ERR_ALIGN;
    }
  }
  Key key;
  memset(&key, 0, sizeof(key));
  key.v[0] = (uint64_t)(uintptr_t)ptr;
  key.v[1] = ((uint64_t)elem_bytes << 32) | (uint64_t)rank;
  for (int i = 0; i < rank; ++i) key.v[2 + i] = dims[i] | ((uint64_t)box[i] << 48);
  for (int i = 0; i < rank - 1; ++i) key.v[5 + i] = strides_bytes[i];
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_cache.find(key);
    if (it != g_cache.end()) {
      *out = it->second;
      return MODCR_OK;
    }
  }
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const CUtensorMapSwizzle sw = box[0] * (cuuint32_t)elem_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu box %u,%u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return MODCR_ERR_DRIVER;
  }
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_cache.size() > 65536) g_cache.clear();
    g_cache.emplace(key, *out);
  }
  return MODCR_OK;
}

int make_tmap_2d(CUtensorMap* out, const void* ptr, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows, uint32_t box_cols) {
  if ((box_cols * elem_bytes != 128 && box_cols * elem_bytes != 64) || box_rows > 256 || box_rows == 0) {
    set_error("bad TMA box %u x %u (elem %d B)", box_rows, box_cols, elem_bytes);
    return MODCR_ERR_SHAPE;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * (uint64_t)elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  return encode(out, ptr, elem_bytes, 2, dims, strides, box);
}

int make_tmap_3d(CUtensorMap* out, const void* ptr, int elem_bytes, uint64_t batch, uint64_t rows, uint64_t cols,
                 uint64_t row_stride, uint64_t batch_stride, uint32_t box_rows, uint32_t box_cols) {
  if (box_cols * elem_bytes != 128 || box_rows > 256 || box_rows == 0) {
    set_error("bad TMA box %u x %u (elem %d B)", box_rows, box_cols, elem_bytes);
    return MODCR_ERR_SHAPE;
  }
  cuuint64_t dims[3] = {cols, rows, batch};
  cuuint64_t strides[2] = {row_stride * (uint64_t)elem_bytes, batch_stride * (uint64_t)elem_bytes};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  return encode(out, ptr, elem_bytes, 3, dims, strides, box);
}

}  // namespace modcr
