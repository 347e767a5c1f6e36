// umma_bench.cu — issue-rate / throughput floor of each tcgen05.mma operand path used by the attention
// kernels (128x128x16 bf16, cta_group::1), one CTA per SM, back-to-back MMAs from one thread.
//   mode 0  SS  A K-major,  B K-major     (S = Q K^T, dP = V dO^T)
//   mode 1  SS  A K-major,  B MN-major    (dK = dS^T Q)
//   mode 2  TS  A in TMEM,  B MN-major    (O = P V, dV = P^T dO)
//   mode 3  SS  A MN-major, B MN-major    (dQ = dS K)
//   mode 4  mix of the backward step: 0,0,2,3,1 per 40 MMAs
// Prints cycles per MMA (ideal 64) and the implied dense TFLOP/s at the measured SM clock.
#include "dta_common.cuh"
#include "dta_ptx.cuh"
#include <vector>

using namespace dta;
static constexpr int T = 128;
static constexpr int HALF = T * 128;

__global__ void __launch_bounds__(128, 1) bench_kernel(int mode, int iters, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * HALF);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 4 * HALF / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(s_tmem, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  if (threadIdx.x == 0) {
    const uint32_t kA = sdesc_lo_kmajor(smem_u32(smem)), kB = sdesc_lo_kmajor(smem_u32(smem + 2 * HALF));
    const uint32_t nA = sdesc_lo_mnmajor(smem_u32(smem), HALF), nB = sdesc_lo_mnmajor(smem_u32(smem + 2 * HALF), HALF);
    constexpr uint32_t i_kk = make_idesc_bf16(T, T, false, false), i_kn = make_idesc_bf16(T, T, false, true),
                       i_nn = make_idesc_bf16(T, T, true, true);
    auto opaque = [](uint32_t x) { asm volatile("" : "+r"(x)); return x; };
    auto gemm = [&](int m, uint32_t d) {
      const uint32_t ka = opaque(kA), kb = opaque(kB), na = opaque(nA), nb = opaque(nB), ta = opaque(tmem + 384);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t ko = (kk >> 2) * (HALF >> 4) + (kk & 3) * 2, no = kk * 128;
        if (m == 0) umma_ss_lo(d, ka + ko, kb + ko, i_kk, kk > 0);
        if (m == 1) umma_ss_lo(d, ka + ko, nb + no, i_kn, kk > 0);
        if (m == 2) umma_ts_lo(d, ta + kk * 8, nb + no, i_kn, kk > 0);
        if (m == 3) umma_ss_lo(d, na + no, nb + no, i_nn, kk > 0);
        if (m == 5) umma_ss_lo(d, ka + ko, kb + ko, make_idesc_bf16(T, 64, false, false), kk > 0);   // N = 64
        if (m == 6) umma_ss_lo(d, ka + ko, kb + ko, make_idesc_bf16(T, 256, false, false), kk > 0);  // N = 256
      }
    };
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (mode < 4) {
        gemm(mode, tmem + (it & 1) * 128);
      } else if (mode == 4) {
        gemm(0, tmem); gemm(0, tmem + 128); gemm(2, tmem + 256); gemm(3, tmem + 128); gemm(1, tmem + 256);
      } else if (mode == 5) {
        gemm(5, tmem); gemm(5, tmem + 64); gemm(5, tmem + 128); gemm(5, tmem + 192); gemm(5, tmem + 256);
      } else {
        gemm(6, tmem); gemm(6, tmem + 256); gemm(6, tmem); gemm(6, tmem + 256); gemm(6, tmem);
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    cycles_out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  long long* d_cycles;
  cudaMalloc(&d_cycles, sms * sizeof(long long));
  const int smem = 4 * HALF + 64;
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4000;
  for (int grid : {1, sms}) {
    for (int mode = 0; mode < 7; ++mode) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      bench_kernel<<<grid, 128, smem>>>(mode, 200, d_cycles);
      cudaEventRecord(e0);
      bench_kernel<<<grid, 128, smem>>>(mode, iters, d_cycles);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      std::vector<long long> h(grid);
      cudaMemcpy(h.data(), d_cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (auto c : h) mx = c > mx ? c : mx;
      const double n_mma = static_cast<double>(iters) * (mode < 4 ? 8 : 40);
      const double flops = n_mma * 2.0 * T * (mode == 5 ? 64 : mode == 6 ? 256 : T) * 16 * grid;
      printf("umma_bench grid=%d mode=%d: %.1f cycles/MMA (max over CTAs), %.3f ms, %.0f TFLOP/s, eff. clock %.0f MHz\n",
             grid, mode, mx / n_mma, ms, flops / (ms * 1e-3) / 1e12, mx / (ms * 1e-3) / 1e6);
    }
  }
  return 0;
}
