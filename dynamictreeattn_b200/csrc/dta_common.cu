// dta_common.cu — error plumbing, TMA descriptor encoding, misc C-ABI entry points.
#include "dta_common.cuh"
#include "../../include/dta_b200.h"
#include <cstring>
#include <mutex>

namespace dta {

static thread_local char g_err[512] = "";

char* last_error_buf() { return g_err; }

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void arm_profile_hook(void* a, void* b);

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

int make_tmap_bf16_rows(CUtensorMap* out, const void* base, int64_t rows, int heads, int head_dim,
                        int64_t row_stride, int64_t head_stride, int box_rows) {
  auto enc = encode_fn();
  if (!enc) return set_error(DTA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0)
    return set_error(DTA_ERR_INVALID, "tensor base must be 16-byte aligned");
  if ((row_stride * 2) % 16 != 0 || (head_stride * 2) % 16 != 0)
    return set_error(DTA_ERR_INVALID, "row/head strides must be multiples of 8 elements");
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(head_dim), static_cast<cuuint64_t>(heads),
                        static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(head_stride) * 2, static_cast<cuuint64_t>(row_stride) * 2};
  cuuint32_t box[3] = {64, 1, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(DTA_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return DTA_OK;
}

int make_tmap_f32_2d(CUtensorMap* out, const void* base, int64_t cols, int64_t rows, int box_cols, int box_rows) {
  auto enc = encode_fn();
  if (!enc) return set_error(DTA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (cols * 4) % 16 != 0)
    return set_error(DTA_ERR_INVALID, "fp32 tensor base / row pitch must be 16-byte aligned");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 4};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(DTA_ERR_CUDA, "cuTensorMapEncodeTiled (fp32 2-D) failed with CUresult %d", (int)r);
  return DTA_OK;
}

static thread_local ProfileHook g_hook;
ProfileHook take_profile_hook() {
  ProfileHook h = g_hook;
  g_hook = ProfileHook{};
  return h;
}
void arm_profile_hook(void* a, void* b) {
  g_hook.start = static_cast<cudaEvent_t>(a);
  g_hook.stop = static_cast<cudaEvent_t>(b);
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

}  // namespace dta

extern "C" {

const char* dta_last_error(void) { return dta::last_error_buf(); }

int dta_abi_version(void) { return DTA_ABI_VERSION; }

int dta_profile_next_attn(void* start_event, void* stop_event) {
  dta::arm_profile_hook(start_event, stop_event);
  return DTA_OK;
}

}  // extern "C"
