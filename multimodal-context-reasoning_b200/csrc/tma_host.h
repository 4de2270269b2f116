// Host-side TMA tensor-map construction (driver entry point fetched at run time,
// so the library links against cudart only) + a small descriptor cache.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace modcr {

// Row-major 2-D view [rows, cols] with leading dimension `ld` (elements).
// The box is [box_rows, box_cols]; box_cols * elem_bytes must be 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B).
// Returns 0 on success, negative error code otherwise (error text via set_error()).
int make_tmap_2d(CUtensorMap* out, const void* ptr, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows, uint32_t box_cols);
// 3-D view [batch, rows, cols], strides in elements.
int make_tmap_3d(CUtensorMap* out, const void* ptr, int elem_bytes, uint64_t batch, uint64_t rows, uint64_t cols,
                 uint64_t row_stride, uint64_t batch_stride, uint32_t box_rows, uint32_t box_cols);

void set_error(const char* fmt, ...);
const char* last_error();

enum : int {
  MODCR_OK = 0,
  MODCR_ERR_SHAPE = -1,
  MODCR_ERR_ALIGN = -2,
  MODCR_ERR_ARCH = -3,
  MODCR_ERR_CUDA = -4,
  MODCR_ERR_DRIVER = -5,
  MODCR_ERR_WORKSPACE = -6,
  MODCR_ERR_UNSUPPORTED = -7,
};

int check_launch(const char* what);   // also counts one kernel launch
unsigned long long launch_count();
void count_launches(int n);           // extra launches behind a single check_launch
int sm_limit();                       // SM budget of the persistent kernels (0 = all)
int set_sm_limit(int sms);            // returns the previous budget

}  // namespace modcr
