// dta_common.cuh — host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#define DTA_OK 0
#define DTA_ERR_INVALID 1
#define DTA_ERR_CUDA 2
#define DTA_ERR_UNSUPPORTED 3

namespace dta {

// Last error text, per host thread; read through dta_last_error().
char* last_error_buf();
int set_error(int code, const char* fmt, ...);

#define DTA_CHECK_CUDA(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return dta::set_error(DTA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                               \
  } while (0)

#define DTA_REQUIRE(cond, ...)                                  \
  do {                                                          \
    if (!(cond)) return dta::set_error(DTA_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// bf16 tensor viewed as (D innermost, heads, rows); box = (64 cols, 1 head, box_rows rows), 128B swizzle.
// row_stride / head_stride in ELEMENTS.  Returns DTA_OK or an error code.
int make_tmap_bf16_rows(CUtensorMap* out, const void* base, int64_t rows, int heads, int head_dim,
                        int64_t row_stride, int64_t head_stride, int box_rows);

// fp32 row-major [rows][cols] tensor, box = (box_cols, box_rows), no swizzle (target of the dQ tensor reduce)
int make_tmap_f32_2d(CUtensorMap* out, const void* base, int64_t cols, int64_t rows, int box_cols, int box_rows);

int sm_count();

// one-shot profiling events armed by dta_profile_next_attn (thread-local)
struct ProfileHook { cudaEvent_t start = nullptr, stop = nullptr; };
ProfileHook take_profile_hook();

}  // namespace dta
