// red_bench.cu — how fast can one SM push fp32 reductions into an L2-resident / HBM-resident accumulator?
// The backward attention kernel adds one 128x128 fp32 dQ tile (64 KB) per step; this measures the paths:
//   mode 0  red.global.add.f32        lanes -> 32 consecutive floats (128 B per warp instruction)   [current]
//   mode 1  red.global.add.v2.f32     lanes -> 32 x 8 B   (256 B contiguous per warp instruction)
//   mode 2  red.global.add.v4.f32     lanes -> 32 x 16 B  (512 B contiguous per warp instruction)
//   mode 3  red.global.add.v4.f32     lanes -> 16 B each in 32 different rows (row stride 8 KB: half-used sectors)
//   mode 4  cp.reduce.async.bulk (TMA) shared -> global add.f32, 512 B rows (128 ops per 64 KB tile)
//   mode 5  cp.reduce.async.bulk (TMA) shared -> global add.f32, 16 KB contiguous pieces (4 ops per tile)
// Every CTA (one per SM, 128 threads like the drain warpgroup) adds `tiles` tiles of 64 KB; tiles of
// different CTAs interleave over a buffer of `mb` MB so the footprint can be L2-resident or not.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void __launch_bounds__(128, 1) red_kernel(float* acc, size_t n_tiles_buf, int tiles, int mode, long long* cyc) {
  extern __shared__ __align__(128) float stage[];   // 64 KB for the TMA modes
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 1.0f + i;
  if (mode >= 4) {
    for (int i = threadIdx.x; i < 16384; i += 128) stage[i] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int t = 0; t < tiles; ++t) {
    const size_t tile = (static_cast<size_t>(t) * gridDim.x + blockIdx.x) % n_tiles_buf;
    float* base = acc + tile * 16384;                     // 128 rows x 128 floats, contiguous here
    if (mode == 0) {
      // warp w handles rows 32w..32w+31; per row 4 instructions of 32 lanes
#pragma unroll 8
      for (int r = 0; r < 32; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c)
          asm volatile("red.global.add.f32 [%0], %1;" ::"l"(base + (warp * 32 + r) * 128 + c * 32 + lane), "f"(v[c]) : "memory");
    } else if (mode == 1) {
#pragma unroll 8
      for (int r = 0; r < 32; ++r)
#pragma unroll
        for (int c = 0; c < 2; ++c)
          asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(base + (warp * 32 + r) * 128 + c * 64 + lane * 2), "f"(v[c]), "f"(v[c + 1]) : "memory");
    } else if (mode == 2) {
#pragma unroll 8
      for (int r = 0; r < 32; ++r)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(base + (warp * 32 + r) * 128 + lane * 4), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    } else if (mode == 3) {
      // lane = row, 32 instructions walk along the row; rows 8 KB apart (16 heads x 512 B)
      float* b3 = acc + (tile % (n_tiles_buf / 16)) * 16384 * 16;
#pragma unroll 8
      for (int j = 0; j < 32; ++j)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(b3 + static_cast<size_t>(warp * 32 + lane) * 2048 + j * 4), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    } else if (mode == 4) {
      if (threadIdx.x == 0) {
        for (int r = 0; r < 128; ++r)
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 512;" ::"l"(base + r * 128), "r"(smem_u32(stage + r * 128)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      __syncthreads();
    } else {
      if (threadIdx.x == 0) {
        for (int r = 0; r < 4; ++r)
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 16384;" ::"l"(base + r * 4096), "r"(smem_u32(stage + r * 4096)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      __syncthreads();
    }
  }
  if (mode >= 4 && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}

int main(int argc, char** argv) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int tiles = 400;
  long long* d_cyc;
  cudaMalloc(&d_cyc, sms * sizeof(long long));
  cudaFuncSetAttribute(red_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int mb : {32, 1024}) {
    const size_t n_tiles_buf = static_cast<size_t>(mb) * 1024 * 1024 / 65536;
    float* acc;
    cudaMalloc(&acc, n_tiles_buf * 65536);
    cudaMemset(acc, 0, n_tiles_buf * 65536);
    for (int grid : {1, sms}) {
      for (int mode = 0; mode < 6; ++mode) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        red_kernel<<<grid, 128, 65536>>>(acc, n_tiles_buf, 20, mode, d_cyc);
        cudaEventRecord(e0);
        red_kernel<<<grid, 128, 65536>>>(acc, n_tiles_buf, tiles, mode, d_cyc);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        std::vector<long long> h(grid);
        cudaMemcpy(h.data(), d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (auto c : h) mx = c > mx ? c : mx;
        printf("red_bench buf=%4d MB grid=%3d mode=%d: %7.0f cycles per 64 KB tile, %7.1f GB/s aggregate (%.3f ms)\n", mb, grid, mode,
               static_cast<double>(mx) / tiles, static_cast<double>(tiles) * grid * 65536 / (ms * 1e-3) / 1e9, ms);
      }
    }
    cudaFree(acc);
  }
  return 0;
}
