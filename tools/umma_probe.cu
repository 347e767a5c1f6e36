// umma_probe.cu — stand-alone check of the tcgen05 building blocks the attention kernels rely on.
// Each mode computes one 128x128x128 bf16 GEMM through a different operand path and compares
// with a host reference, so a wrong descriptor field shows up as ONE failing line:
//   0  A smem K-major,  B smem K-major          (Q K^T, V dO^T)          D = A  B^T
//   1  A smem K-major,  B smem MN-major         (B given as [K][N])      D = A  B
//   2  A TMEM (bf16x2), B smem MN-major         (P V, P^T dO)            D = A  B
//   3  A smem MN-major, B smem MN-major         (dQ = dS K)              D = A^T B
//   4  A smem K-major written by threads with the manual 128B swizzle, B MN-major (dK = dS^T Q)
//   5  mode 0 plus ONE extra K=16 step whose operands are un-swizzled K-major tiles written by threads
//      (8-row x 16-byte core matrices; LBO = K-direction chunk distance, SBO = 128): the broadcast trick
//      that folds per-query constants (lse, delta) into S^T / dP^T on the tensor core
//   6  same with LBO / SBO swapped (the other reading of the descriptor fields; exactly one of 5/6 passes)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I dynamictreeattn_b200/csrc \
//        -I include tools/umma_probe.cu dynamictreeattn_b200/csrc/dta_common.cu -o tools/bin/umma_probe
#include "dta_common.cuh"
#include "dta_ptx.cuh"
#include <vector>
#include <cmath>
#include <cstdlib>

using namespace dta;

static constexpr int T = 128;
static constexpr int HALF = T * 128;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, int mode,
             const __nv_bfloat16* a_glob, const __nv_bfloat16* b_glob, float* d_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + 2 * HALF;
  uint8_t* sAug = smem + 4 * HALF;                 // A_aug chunk0 | chunk1 | B_aug chunk0 | chunk1 (2 KB each)
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + 4 * HALF + 8192);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_load + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) { printf("probe: smem base not 1024-aligned\n"); __trap(); }
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(s_tmem, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const uint32_t tD = tmem, tA = tmem + 128;

  if (threadIdx.x == 0) {
    const bool a_by_tma = (mode != 4 && mode != 2);
    (void)sAug;
    mbar_arrive_expect_tx(bar_load, (a_by_tma ? 2 : 1) * 2 * HALF);
    for (int h = 0; h < 2; ++h) {
      if (a_by_tma) tma_load_3d(sA + h * HALF, &tm_a, bar_load, h * 64, 0, 0);
      tma_load_3d(sB + h * HALF, &tm_b, bar_load, h * 64, 0, 0);
    }
  }
  const int row = threadIdx.x;
  const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
  if (mode == 4) {
    // A[row][:] (K-major) written by hand with the 128B swizzle
    for (int h = 0; h < 2; ++h)
      for (int chunk = 0; chunk < 8; ++chunk) {
        const uint4 v = *reinterpret_cast<const uint4*>(a_glob + row * T + h * 64 + chunk * 8);
        *reinterpret_cast<uint4*>(sA + h * HALF + sw128_offset(row, chunk)) = v;
      }
    fence_proxy_async_smem();
  }
  if (mode >= 5) {
    // aug operands: element (row, k) of A_aug = A[row][k], of B_aug = B[row][k] (k < 16), canonical un-swizzled layout
    for (int j = 0; j < 2; ++j) {
      *reinterpret_cast<uint4*>(sAug + j * 2048 + row * 16) = *reinterpret_cast<const uint4*>(a_glob + row * T + j * 8);
      *reinterpret_cast<uint4*>(sAug + 4096 + j * 2048 + row * 16) = *reinterpret_cast<const uint4*>(b_glob + row * T + j * 8);
    }
    fence_proxy_async_smem();
  }
  if (mode == 2) {
    // A[row][:] -> TMEM as packed bf16 pairs: column c holds (k = 2c, 2c+1)
    uint32_t pk[64];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a_glob + row * T);
    for (int c = 0; c < 64; ++c) pk[c] = src[c];
    tmem_st_x32(tA + lane_addr, pk);
    tmem_st_x32(tA + lane_addr + 32, pk + 32);
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();

  if (threadIdx.x == 0) {
    mbar_wait(bar_load, 0);
    tc_fence_after();
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    auto kmaj = [](uint32_t base, int kk) { return make_sdesc_sw128(base + (kk >> 2) * HALF + (kk & 3) * 32, 16, 1024); };
    auto mnmaj = [](uint32_t base, int kk) { return make_sdesc_sw128(base + kk * 2048, HALF, 1024); };
    for (int kk = 0; kk < 8; ++kk) {
      const uint32_t acc = kk > 0;
      if (mode == 0) umma_ss(tD, kmaj(a0, kk), kmaj(b0, kk), make_idesc_bf16(T, T, false, false), acc);
      if (mode == 1) umma_ss(tD, kmaj(a0, kk), mnmaj(b0, kk), make_idesc_bf16(T, T, false, true), acc);
      if (mode == 2) umma_ts(tD, tA + kk * 8, mnmaj(b0, kk), make_idesc_bf16(T, T, false, true), acc);
      if (mode == 3) umma_ss(tD, mnmaj(a0, kk), mnmaj(b0, kk), make_idesc_bf16(T, T, true, true), acc);
      if (mode == 4) umma_ss(tD, kmaj(a0, kk), mnmaj(b0, kk), make_idesc_bf16(T, T, false, true), acc);
      if (mode >= 5) umma_ss(tD, kmaj(a0, kk), kmaj(b0, kk), make_idesc_bf16(T, T, false, false), acc);
    }
    if (mode >= 5) {
      const uint32_t lbo = mode == 5 ? 2048u : 128u, sbo = mode == 5 ? 128u : 2048u;
      umma_ss(tD, make_sdesc_noswz(smem_u32(sAug), lbo, sbo), make_sdesc_noswz(smem_u32(sAug) + 4096, lbo, sbo),
              make_idesc_bf16(T, T, false, false), 1u);
    }
    umma_commit(bar_mma);
  }
  mbar_wait(bar_mma, 0);
  tc_fence_after();
  for (int c = 0; c < 4; ++c) {
    uint32_t r[32];
    tmem_ld_x32(tD + lane_addr + c * 32, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) d_out[row * T + c * 32 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  std::vector<float> A(T * T), B(T * T);
  srand(1);
  for (auto& x : A) x = bf((rand() % 2001 - 1000) / 1000.f);
  for (auto& x : B) x = bf((rand() % 2001 - 1000) / 1000.f);
  std::vector<__nv_bfloat16> hA(T * T), hB(T * T);
  for (int i = 0; i < T * T; ++i) { hA[i] = __float2bfloat16(A[i]); hB[i] = __float2bfloat16(B[i]); }
  __nv_bfloat16 *dA, *dB;
  float* dD;
  cudaMalloc(&dA, T * T * 2); cudaMalloc(&dB, T * T * 2); cudaMalloc(&dD, T * T * 4);
  cudaMemcpy(dA, hA.data(), T * T * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), T * T * 2, cudaMemcpyHostToDevice);
  CUtensorMap ta, tb;
  if (make_tmap_bf16_rows(&ta, dA, T, 1, T, T, T, T) || make_tmap_bf16_rows(&tb, dB, T, 1, T, T, T, T)) {
    printf("tensor map failed: %s\n", last_error_buf());
    return 2;
  }
  const int smem = 4 * HALF + 8192 + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int fails = 0;
  for (int mode = 0; mode < 6; ++mode) {   // mode 6 (LBO/SBO swapped) faults on B200: the mode-5 reading is the right one
    cudaMemset(dD, 0, T * T * 4);
    probe_kernel<<<1, 128, smem>>>(ta, tb, mode, dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 3; }
    std::vector<float> D(T * T);
    cudaMemcpy(D.data(), dD, T * T * 4, cudaMemcpyDeviceToHost);
    double max_err = 0;
    for (int m = 0; m < T; ++m)
      for (int n = 0; n < T; ++n) {
        double ref = 0;
        for (int k = 0; k < T; ++k) {
          // A element (m,k): modes 0,1,2,4 read A[m][k]; mode 3 reads A stored as [K][M] -> A[k][m]
          const float a = (mode == 3) ? A[k * T + m] : A[m * T + k];
          // B element (k,n): mode 0 reads B stored as [N][K]; others as [K][N]
          const float b = (mode == 0 || mode >= 5) ? B[n * T + k] : B[k * T + n];
          ref += static_cast<double>(a) * b;
        }
        if (mode >= 5)
          for (int k = 0; k < 16; ++k) ref += static_cast<double>(A[m * T + k]) * B[n * T + k];
        max_err = fmax(max_err, fabs(ref - D[m * T + n]));
      }
    const bool ok = max_err < 1e-2;
    printf("umma_probe mode %d: max_abs_err %.3e %s\n", mode, max_err, ok ? "PASS" : "FAIL");
    fails += !ok;
  }
  return fails ? 1 : 0;
}
