// tma_bench.cu — latency of the backward kernel's Q / dO tile loads.  Every CTA (one per SM) loads 32 KB tiles
// ([128 rows][128 bf16], two 128B-swizzled [128][64] boxes) one after the other and waits for each:
//   layout 0: token-major [N][16 heads][128]  (row stride 4 KB: what the HF projections hand us)
//   layout 1: head-major  [16 heads][N][128]  (a tile is 32 contiguous KB)
// with `inflight` = 1 or 2 tiles requested back to back (Q and dO).  Prints cycles from issue to mbarrier completion.
#include "dta_common.cuh"
#include "dta_ptx.cuh"
#include <vector>
using namespace dta;
static constexpr int T = 128, HALF = T * 128, TILE = 2 * HALF;

__global__ void __launch_bounds__(128, 1) tma_kernel(const __grid_constant__ CUtensorMap tm, int n_tiles_rows, int iters, int inflight,
                                                     long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * TILE);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long total = 0;
    for (int it = 0; it < iters; ++it) {
      const int tile = (it * gridDim.x + blockIdx.x) % n_tiles_rows;
      const int head = (it * 7 + blockIdx.x) % 16;
      const long long t0 = clock64();
      mbar_arrive_expect_tx(bar, inflight * TILE);
      for (int f = 0; f < inflight; ++f)
        for (int h = 0; h < 2; ++h) tma_load_3d(smem + f * TILE + h * HALF, &tm, bar, h * 64, (head + f) % 16, tile * T);
      mbar_wait(bar, it & 1);
      total += clock64() - t0;
    }
    cyc[blockIdx.x] = total / iters;
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d_cyc;
  cudaMalloc(&d_cyc, sms * sizeof(long long));
  cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE + 64);
  for (int n_rows : {4096, 65536}) {              // 16 MB (L2 resident) and 268 MB
    __nv_bfloat16* buf;
    const size_t bytes = static_cast<size_t>(n_rows) * 16 * 128 * 2;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 0, bytes);
    for (int layout = 0; layout < 2; ++layout) {
      CUtensorMap tm;
      // make_tmap_bf16_rows(out, base, rows, heads, head_dim, row_stride, head_stride, box_rows) — strides in elements
      const int rc = layout == 0 ? make_tmap_bf16_rows(&tm, buf, n_rows, 16, 128, 16 * 128, 128, T)
                                 : make_tmap_bf16_rows(&tm, buf, n_rows, 16, 128, 128, static_cast<int64_t>(n_rows) * 128, T);
      if (rc) { printf("tensor map failed: %s\n", last_error_buf()); return 2; }
      for (int grid : {1, sms}) {
        for (int inflight = 1; inflight <= 2; ++inflight) {
          tma_kernel<<<grid, 128, 2 * TILE + 64>>>(tm, n_rows / T, 50, inflight, d_cyc);
          tma_kernel<<<grid, 128, 2 * TILE + 64>>>(tm, n_rows / T, 400, inflight, d_cyc);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
          std::vector<long long> h(grid);
          cudaMemcpy(h.data(), d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
          long long mx = 0, sum = 0;
          for (auto c : h) { mx = c > mx ? c : mx; sum += c; }
          printf("tma_bench %4zu MB %s grid=%3d tiles in flight=%d: mean %5lld  max %5lld cycles per wait\n", bytes >> 20,
                 layout == 0 ? "token-major" : "head-major ", grid, inflight, sum / grid, mx);
        }
      }
    }
    cudaFree(buf);
  }
  return 0;
}
