// logprob_kernels.cu — fused per-token log-prob / entropy over a vocab shard, fwd + bwd.
//
// Replaces the fp32 log_softmax + gather + entropy of the reference
// (vocab_parallel.py:13-27) and its vocab-parallel autograd functions
// (vocab_parallel.py:117-166, 249-321, 324-370).  HBM-bound: the forward reads each logit
// once (online softmax, 16-byte loads), the backward reads it once and writes the gradient
// once, optionally in place.  Per row the forward emits 16 bytes {m, sum e^(x-m),
// sum x e^(x-m), 0}; that is all a vocab-parallel rank has to exchange.
#include "dta_common.cuh"
#include "../../include/dta_b200.h"
#include <cuda_bf16.h>
#include <cstdlib>

namespace dta {

static constexpr float kLog2e = 1.4426950408889634f;
static constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct RowStat {  // running (max, sum e^(x-m), sum raw*e^(x-m)); x = raw * inv_temp
  float m, s, t;
};
__device__ __forceinline__ RowStat merge(RowStat a, RowStat b) {
  const float m = fmaxf(a.m, b.m);
  if (m == -INFINITY) return RowStat{m, 0.f, 0.f};
  const float fa = ex2((a.m - m) * kLog2e), fb = ex2((b.m - m) * kLog2e);
  return RowStat{m, a.s * fa + b.s * fb, a.t * fa + b.t * fb};
}

template <typename T> struct Vec;
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void load(const __nv_bfloat16* p, float* out) {
    const uint4 v = __ldcs(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      out[2 * i] = __uint_as_float(w[i] << 16);
      out[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static void store(__nv_bfloat16* p, const float* in) {
    uint4 v;
    uint32_t* w = &v.x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(in[2 * i], in[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    __stcs(reinterpret_cast<uint4*>(p), v);
  }
  __device__ static float get(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  __device__ static void put(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};
template <> struct Vec<float> {
  static constexpr int N = 4;
  __device__ static void load(const float* p, float* out) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(p));
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
  }
  __device__ static void store(float* p, const float* in) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(in[0], in[1], in[2], in[3]));
  }
  __device__ static float get(const float* p) { return *p; }
  __device__ static void put(float* p, float v) { *p = v; }
};

// ------------------------------------------------------------------------------------
// forward: per-row online softmax statistics
// ------------------------------------------------------------------------------------
template <typename T, int kRowThreads>
__global__ void __launch_bounds__(kRowThreads) row_stats_kernel(const T* __restrict__ logits, int64_t n_rows,
                                                               int64_t n_cols, int64_t row_stride, float inv_temp,
                                                               const int64_t* __restrict__ labels, int64_t vocab_start,
                                                               float* __restrict__ stats) {
  constexpr int V = Vec<T>::N;
  constexpr int U = 4;  // vectors in flight per thread
  __shared__ RowStat s_part[kRowThreads / 32];
  __shared__ float s_xl;
  const float c = inv_temp * kLog2e;
  for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const T* p = logits + row * row_stride;
    // the row's label logit (one label per row): requested first by one thread of the LAST warp, so the dependent
    // load is in flight under the whole row pass and thread 0 finds it in shared memory after the block reduction
    float xl = 0.f;
    if (labels != nullptr && threadIdx.x == kRowThreads - 1) {
      const int64_t lc = labels[row] - vocab_start;
      if (lc >= 0 && lc < n_cols) xl = Vec<T>::get(p + lc) * inv_temp;
    }
    RowStat st{-INFINITY, 0.f, 0.f};
    const bool vec_ok = (reinterpret_cast<uintptr_t>(p) & 15) == 0;
    const int64_t n_vec = vec_ok ? n_cols / V : 0;
    // vector u of a thread is kRowThreads vectors further along the row: every load instruction of a warp covers 512
    // contiguous bytes (a thread-contiguous mapping fetched each 32-byte sector twice)
    for (int64_t v0 = threadIdx.x; v0 < n_vec; v0 += static_cast<int64_t>(kRowThreads) * U) {
      float x[U * V];
      int cnt = 0;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (v0 + u * kRowThreads < n_vec) { Vec<T>::load(p + (v0 + u * kRowThreads) * V, x + u * V); cnt = (u + 1) * V; }
      }
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < U * V; ++i) if (i < cnt) mx = fmaxf(mx, x[i]);
      mx *= inv_temp;  // inv_temp > 0
      if (mx > st.m) {
        const float f = ex2((st.m - mx) * kLog2e);  // ex2(-inf) = 0 on the first chunk
        st.s *= f; st.t *= f; st.m = mx;
      }
      const float mo = st.m * kLog2e;
#pragma unroll
      for (int i = 0; i < U * V; ++i) {
        if (i < cnt) {
          const float e = ex2(fmaf(x[i], c, -mo));
          st.s += e;
          st.t = fmaf(x[i], e, st.t);
        }
      }
    }
    // scalar tail (and whole row when the row is not 16-byte aligned)
    for (int64_t i = n_vec * V + threadIdx.x; i < n_cols; i += kRowThreads) {
      const float raw = Vec<T>::get(p + i);
      st = merge(st, RowStat{raw * inv_temp, 1.f, raw});
    }
    // block reduction
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      RowStat other{__shfl_xor_sync(0xffffffffu, st.m, o), __shfl_xor_sync(0xffffffffu, st.s, o),
                    __shfl_xor_sync(0xffffffffu, st.t, o)};
      st = merge(st, other);
    }
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = st;
    if (threadIdx.x == kRowThreads - 1) s_xl = xl;
    __syncthreads();
    if (threadIdx.x < 32) {
      RowStat r = threadIdx.x < kRowThreads / 32 ? s_part[threadIdx.x] : RowStat{-INFINITY, 0.f, 0.f};
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        RowStat other{__shfl_xor_sync(0xffffffffu, r.m, o), __shfl_xor_sync(0xffffffffu, r.s, o),
                      __shfl_xor_sync(0xffffffffu, r.t, o)};
        r = merge(r, other);
      }
      if (threadIdx.x == 0)
        reinterpret_cast<float4*>(stats)[row] = make_float4(r.m, r.s, r.t * inv_temp, s_xl);
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void gather_labels_kernel(const T* __restrict__ logits, int64_t n_cols, int64_t row_stride, float inv_temp,
                                     const int64_t* __restrict__ labels, const int32_t* __restrict__ row_of,
                                     int64_t n_labels, int64_t vocab_start, float* __restrict__ x_label,
                                     int64_t x_stride) {
  for (int64_t m = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; m < n_labels;
       m += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = row_of ? row_of[m] : m;
    const int64_t c = labels[m] - vocab_start;
    float v = 0.f;
    if (r >= 0 && c >= 0 && c < n_cols) v = Vec<T>::get(logits + r * row_stride + c) * inv_temp;
    x_label[m * x_stride] = v;
  }
}

__global__ void finalize_kernel(const float4* __restrict__ stats_all, int world, int64_t world_stride4,
                                int64_t n_rows, float* __restrict__ logz, float* __restrict__ entropy,
                                float* __restrict__ logprob) {
  for (int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; r < n_rows;
       r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float m = -INFINITY;
    for (int w = 0; w < world; ++w) m = fmaxf(m, stats_all[w * world_stride4 + r].x);
    float s = 0.f, t = 0.f, xl = 0.f;
    for (int w = 0; w < world; ++w) {
      const float4 st = stats_all[w * world_stride4 + r];
      xl += st.w;
      const float f = st.x == -INFINITY ? 0.f : expf(st.x - m);
      s = fmaf(st.y, f, s);
      t = fmaf(st.z, f, t);
    }
    const float lz = m + logf(s);
    logz[r] = lz;
    entropy[r] = lz - t / s;
    if (logprob) logprob[r] = xl - lz;
  }
}

// ------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------
template <typename T, int kRowThreads>
__global__ void __launch_bounds__(kRowThreads) logprob_bwd_kernel(const T* __restrict__ logits, int64_t n_rows,
                                                                 int64_t n_cols, int64_t row_stride, float inv_temp,
                                                                 const float* __restrict__ logz,
                                                                 const float* __restrict__ entropy,
                                                                 const float* __restrict__ g_entropy,
                                                                 const float* __restrict__ g_lp_rowsum,
                                                                 const int64_t* __restrict__ labels, int64_t vocab_start,
                                                                 T* __restrict__ dlogits, int64_t d_row_stride) {
  constexpr int V = Vec<T>::N;
  constexpr int U = 4;
  for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const T* p = logits + row * row_stride;
    T* q = dlogits + row * d_row_stride;
    const float lz = logz[row];
    const float a = g_entropy ? g_entropy[row] : 0.f;
    const float G = g_lp_rowsum ? g_lp_rowsum[row] : 0.f;
    // d = inv_temp * p * (a * (-H - lp) - G),  lp = x - lz,  p = e^lp
    const float k0 = -(a * entropy[row] + G);
    const float c = inv_temp * kLog2e, lz2 = lz * kLog2e;
    // one label per row: the one-hot term  + inv_temp * g_lp[row]  lands in the same pass (G is that row's g_lp)
    const int64_t lc = labels != nullptr ? labels[row] - vocab_start : -1;
    const float hot = inv_temp * G;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(q)) & 15) == 0;
    const int64_t n_vec = vec_ok ? n_cols / V : 0;
    for (int64_t v0 = threadIdx.x; v0 < n_vec; v0 += static_cast<int64_t>(kRowThreads) * U) {
      float x[U * V];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (v0 + u * kRowThreads < n_vec) Vec<T>::load(p + (v0 + u * kRowThreads) * V, x + u * V);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (v0 + u * kRowThreads < n_vec) {
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const float lp2 = fmaf(x[u * V + i], c, -lz2);  // log2 p
            const float pr = ex2(lp2);
            x[u * V + i] = inv_temp * pr * fmaf(-a * kLn2, lp2, k0);
          }
          const int64_t e0 = (v0 + u * kRowThreads) * V;           // first element of this vector
          if (static_cast<uint64_t>(lc - e0) < static_cast<uint64_t>(V)) {
#pragma unroll
            for (int i = 0; i < V; ++i)
              if (e0 + i == lc) x[u * V + i] += hot;
          }
          Vec<T>::store(q + e0, x + u * V);
        }
      }
    }
    for (int64_t i = n_vec * V + threadIdx.x; i < n_cols; i += kRowThreads) {
      const float lp2 = fmaf(Vec<T>::get(p + i), c, -lz2);
      const float pr = ex2(lp2);
      Vec<T>::put(q + i, inv_temp * pr * fmaf(-a * kLn2, lp2, k0) + (i == lc ? hot : 0.f));
    }
  }
}

template <typename T>
__global__ void bwd_scatter_kernel(T* __restrict__ dlogits, int64_t n_cols, int64_t d_row_stride, float inv_temp,
                                   const int64_t* __restrict__ labels, const int32_t* __restrict__ row_of,
                                   const float* __restrict__ g_lp, int64_t n_labels, int64_t vocab_start) {
  for (int64_t m = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; m < n_labels;
       m += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = row_of ? row_of[m] : m;
    const int64_t c = labels[m] - vocab_start;
    if (r >= 0 && c >= 0 && c < n_cols) {
      T* p = dlogits + r * d_row_stride + c;
      Vec<T>::put(p, Vec<T>::get(p) + inv_temp * g_lp[m]);
    }
  }
}

__global__ void rowsum_grad_kernel(const float* __restrict__ g_lp, const int32_t* __restrict__ row_of,
                                   int64_t n_labels, float* __restrict__ out) {
  for (int64_t m = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; m < n_labels;
       m += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = row_of ? row_of[m] : m;
    if (r >= 0) atomicAdd(out + r, g_lp[m]);
  }
}

static int small_grid(int64_t n, int threads) {
  int64_t g = (n + threads - 1) / threads;
  const int64_t cap = 16 * static_cast<int64_t>(sm_count());
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}
static int row_grid(int64_t n_rows, int threads) {
  const int64_t cap = (8 * 512 / threads) * static_cast<int64_t>(sm_count());  // grid-stride over rows; multiple of SM count
  return static_cast<int>(n_rows < cap ? n_rows : cap);
}
// Threads per row (= per block).  Measured on the B200 with the coalesced vector mapping (profiles/r02/logprob_threads_ab.log):
// the forward is latency-bound on its online-softmax chain and likes MANY small blocks (64-128 threads: 4.6 TB/s at
// 151,936 columns, 4.2-4.5 TB/s on vocab-parallel shards; 512 threads: 3.9 / 2.2-2.9); the backward is a pure stream
// and holds 5.8-6.0 TB/s (88-92 % of the measured copy bandwidth) from 64 to 256 threads.  DTA_LOGPROB_THREADS forces one
// size for both (A/B knob).
static int row_threads(bool backward, int64_t n_cols) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("DTA_LOGPROB_THREADS");
    const int v = e ? atoi(e) : 0;
    forced = (v == 32 || v == 64 || v == 128 || v == 256 || v == 512 || v == 1024) ? v : 0;
  }
  if (forced) return forced;
  if (backward) return 256;
  return n_cols >= 65536 ? 128 : 64;
}

}  // namespace dta

using namespace dta;

extern "C" {

int dta_logprob_row_stats(const void* logits, int dtype, int64_t n_rows, int64_t n_cols, int64_t row_stride,
                          float inv_temperature, const int64_t* labels, int64_t vocab_start, float* stats,
                          dta_stream_t stream) {
  DTA_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (bf16) or 1 (fp32)");
  DTA_REQUIRE(n_rows >= 0 && n_cols >= 1 && row_stride >= n_cols, "bad logits shape");
  DTA_REQUIRE(inv_temperature > 0.f, "temperature must be positive");
  if (n_rows == 0) return DTA_OK;
  DTA_REQUIRE(logits && stats, "NULL pointer argument");
  DTA_REQUIRE((reinterpret_cast<uintptr_t>(stats) & 15) == 0, "stats must be 16-byte aligned");
  const int th = row_threads(false, n_cols);
#define DTA_ROW_STATS(T, TH)                                                                                       \
  row_stats_kernel<T, TH><<<row_grid(n_rows, TH), TH, 0, stream>>>(static_cast<const T*>(logits), n_rows, n_cols, \
                                                                   row_stride, inv_temperature, labels, vocab_start, stats)
  if (dtype == 0) {
    switch (th) {
      case 32: DTA_ROW_STATS(__nv_bfloat16, 32); break;
      case 64: DTA_ROW_STATS(__nv_bfloat16, 64); break;
      case 128: DTA_ROW_STATS(__nv_bfloat16, 128); break;
      case 256: DTA_ROW_STATS(__nv_bfloat16, 256); break;
      case 1024: DTA_ROW_STATS(__nv_bfloat16, 1024); break;
      default: DTA_ROW_STATS(__nv_bfloat16, 512);
    }
  } else {
    switch (th) {
      case 32: DTA_ROW_STATS(float, 32); break;
      case 64: DTA_ROW_STATS(float, 64); break;
      case 128: DTA_ROW_STATS(float, 128); break;
      case 256: DTA_ROW_STATS(float, 256); break;
      case 1024: DTA_ROW_STATS(float, 1024); break;
      default: DTA_ROW_STATS(float, 512);
    }
  }
#undef DTA_ROW_STATS
  DTA_CHECK_CUDA(cudaGetLastError());
  return DTA_OK;
}

int dta_logprob_gather_labels(const void* logits, int dtype, int64_t n_cols, int64_t row_stride,
                              float inv_temperature, const int64_t* labels, const int32_t* row_of,
                              int64_t n_labels, int64_t vocab_start, float* x_label, int64_t x_stride,
                              dta_stream_t stream) {
  DTA_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (bf16) or 1 (fp32)");
  if (n_labels == 0) return DTA_OK;
  DTA_REQUIRE(logits && labels && x_label && x_stride >= 1, "NULL pointer argument / bad x_stride");
  if (dtype == 0)
    gather_labels_kernel<__nv_bfloat16><<<small_grid(n_labels, 256), 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(logits), n_cols, row_stride, inv_temperature, labels, row_of, n_labels,
        vocab_start, x_label, x_stride);
  else
    gather_labels_kernel<float><<<small_grid(n_labels, 256), 256, 0, stream>>>(
        static_cast<const float*>(logits), n_cols, row_stride, inv_temperature, labels, row_of, n_labels,
        vocab_start, x_label, x_stride);
  DTA_CHECK_CUDA(cudaGetLastError());
  return DTA_OK;
}

int dta_logprob_finalize(const float* stats_all, int world, int64_t world_stride, int64_t n_rows, float* logz,
                         float* entropy, float* logprob, dta_stream_t stream) {
  DTA_REQUIRE(world >= 1, "world must be >= 1");
  if (n_rows == 0) return DTA_OK;
  DTA_REQUIRE(stats_all && logz && entropy, "NULL pointer argument");
  DTA_REQUIRE(world_stride >= 4 * n_rows && world_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(stats_all) & 15) == 0,
              "stats_all: 16-byte aligned blocks of at least 4 * n_rows floats, world_stride a multiple of 4");
  finalize_kernel<<<small_grid(n_rows, 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(stats_all), world,
                                                               world_stride / 4, n_rows, logz, entropy, logprob);
  DTA_CHECK_CUDA(cudaGetLastError());
  return DTA_OK;
}

int dta_logprob_bwd(const void* logits, int dtype, int64_t n_rows, int64_t n_cols, int64_t row_stride,
                    float inv_temperature, const float* logz, const float* entropy, const float* g_entropy,
                    const float* g_lp_rowsum, const int64_t* labels, int64_t vocab_start, void* dlogits,
                    int64_t d_row_stride, dta_stream_t stream) {
  DTA_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (bf16) or 1 (fp32)");
  DTA_REQUIRE(n_rows >= 0 && n_cols >= 1 && row_stride >= n_cols && d_row_stride >= n_cols, "bad logits shape");
  if (n_rows == 0) return DTA_OK;
  DTA_REQUIRE(logits && logz && entropy && dlogits, "NULL pointer argument");
  DTA_REQUIRE(labels == nullptr || g_lp_rowsum != nullptr, "labels (one per row) need the per-row g_lp in g_lp_rowsum");
  const int th = row_threads(true, n_cols);
#define DTA_ROW_BWD(T, TH)                                                                                              \
  logprob_bwd_kernel<T, TH><<<row_grid(n_rows, TH), TH, 0, stream>>>(                                                  \
      static_cast<const T*>(logits), n_rows, n_cols, row_stride, inv_temperature, logz, entropy, g_entropy, g_lp_rowsum, \
      labels, vocab_start, static_cast<T*>(dlogits), d_row_stride)
  if (dtype == 0) {
    switch (th) {
      case 32: DTA_ROW_BWD(__nv_bfloat16, 32); break;
      case 64: DTA_ROW_BWD(__nv_bfloat16, 64); break;
      case 128: DTA_ROW_BWD(__nv_bfloat16, 128); break;
      case 256: DTA_ROW_BWD(__nv_bfloat16, 256); break;
      case 1024: DTA_ROW_BWD(__nv_bfloat16, 1024); break;
      default: DTA_ROW_BWD(__nv_bfloat16, 512);
    }
  } else {
    switch (th) {
      case 32: DTA_ROW_BWD(float, 32); break;
      case 64: DTA_ROW_BWD(float, 64); break;
      case 128: DTA_ROW_BWD(float, 128); break;
      case 256: DTA_ROW_BWD(float, 256); break;
      case 1024: DTA_ROW_BWD(float, 1024); break;
      default: DTA_ROW_BWD(float, 512);
    }
  }
#undef DTA_ROW_BWD
  DTA_CHECK_CUDA(cudaGetLastError());
  return DTA_OK;
}

int dta_logprob_bwd_scatter(void* dlogits, int dtype, int64_t n_cols, int64_t d_row_stride, float inv_temperature,
                            const int64_t* labels, const int32_t* row_of, const float* g_lp, int64_t n_labels,
                            int64_t vocab_start, dta_stream_t stream) {
  DTA_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (bf16) or 1 (fp32)");
  if (n_labels == 0) return DTA_OK;
  DTA_REQUIRE(dlogits && labels && g_lp, "NULL pointer argument");
  if (dtype == 0)
    bwd_scatter_kernel<__nv_bfloat16><<<small_grid(n_labels, 256), 256, 0, stream>>>(
        static_cast<__nv_bfloat16*>(dlogits), n_cols, d_row_stride, inv_temperature, labels, row_of, g_lp, n_labels,
        vocab_start);
  else
    bwd_scatter_kernel<float><<<small_grid(n_labels, 256), 256, 0, stream>>>(
        static_cast<float*>(dlogits), n_cols, d_row_stride, inv_temperature, labels, row_of, g_lp, n_labels,
        vocab_start);
  DTA_CHECK_CUDA(cudaGetLastError());
  return DTA_OK;
}

int dta_logprob_rowsum_grad(const float* g_lp, const int32_t* row_of, int64_t n_labels, float* g_lp_rowsum,
                            dta_stream_t stream) {
  if (n_labels == 0) return DTA_OK;
  DTA_REQUIRE(g_lp && g_lp_rowsum, "NULL pointer argument");
  rowsum_grad_kernel<<<small_grid(n_labels, 256), 256, 0, stream>>>(g_lp, row_of, n_labels, g_lp_rowsum);
  DTA_CHECK_CUDA(cudaGetLastError());
  return DTA_OK;
}

}  // extern "C"
