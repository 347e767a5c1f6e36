// umma_interf.cu — what slows tcgen05.mma down inside the attention kernels?  One warp issues the backward
// step's MMA mix (S^T, dP^T, dV(TS), dQ^T, dK: 40 x 128x128x16 bf16, nominal 64 cycles each) while the other
// warps of the CTA generate ONE kind of interference; prints cycles per MMA for each kind.
//   0 none                      1 ALU/MUFU loop (8 warps)          2 STS.128 stream (8 warps)
//   3 global reds (4 warps)     4 tcgen05.ld of another TMEM region (8 warps)
//   5 mbarrier try_wait polling on a barrier that never completes (12 warps)
//   6 LDS.128 broadcast stream (8 warps)                           7 tcgen05.ld + tcgen05.st (8 warps)
#include "dta_common.cuh"
#include "dta_ptx.cuh"
#include <vector>

using namespace dta;
static constexpr int T = 128;
static constexpr int HALF = T * 128;

__global__ void __launch_bounds__(512, 1) interf_kernel(int kind, int iters, float* gacc, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* scratch = smem + 4 * HALF;                         // 32 KB for STS / LDS interference
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * HALF + 32768);
  uint64_t* bar_never = bar + 1;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar + 3);
  volatile int* s_stop = reinterpret_cast<volatile int*>(bar + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 4 * HALF / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar_never, 1 << 20); mbar_init(bar_never + 1, 1); *s_stop = 0; fence_mbar_init(); }
  if (warp == 15) tmem_alloc(s_tmem, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  if (warp == 15) {
    const uint32_t kA = sdesc_lo_kmajor(smem_u32(smem)), kB = sdesc_lo_kmajor(smem_u32(smem + 2 * HALF));
    const uint32_t nA = sdesc_lo_mnmajor(smem_u32(smem), HALF), nB = sdesc_lo_mnmajor(smem_u32(smem + 2 * HALF), HALF);
    constexpr uint32_t i_kk = make_idesc_bf16(T, T, false, false), i_kn = make_idesc_bf16(T, T, false, true),
                       i_nn = make_idesc_bf16(T, T, true, true);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (kind == 10) {
        // the kernel's order: dV (A = P^T in TMEM columns [0,32) u [64,96)) then S^T overwriting [0,128) right behind it
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_ts_lo(tmem + 256, tmem + (kk >> 2) * 64 + (kk & 3) * 8, nB + kk * 128, i_kn, kk > 0);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_ss_lo(tmem, kA + (kk >> 2) * (HALF >> 4) + (kk & 3) * 2, kB + (kk >> 2) * (HALF >> 4) + (kk & 3) * 2, i_kk, kk > 0);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_ss_lo(tmem + 128, nA + kk * 128, nB + kk * 128, i_nn, kk > 0);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_ss_lo(tmem + 384, kA + (kk >> 2) * (HALF >> 4) + (kk & 3) * 2, nB + kk * 128, i_kn, kk > 0);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_ss_lo(tmem + 128, kA + (kk >> 2) * (HALF >> 4) + (kk & 3) * 2, kB + (kk >> 2) * (HALF >> 4) + (kk & 3) * 2, i_kk, kk > 0);
        }
        __syncwarp();
        continue;
      }
      if (kind == 8) {
        // same mix, one tcgen05.commit after every GEMM (6 per step in the kernel)
        if (elect_one()) {
#pragma unroll
          for (int g = 0; g < 5; ++g) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) umma_ss_lo(tmem + (g & 3) * 128, kA + (kk >> 2) * (HALF >> 4) + (kk & 3) * 2, kB + (kk >> 2) * (HALF >> 4) + (kk & 3) * 2, i_kk, kk > 0);
            umma_commit(bar_never);
          }
        }
        __syncwarp();
        continue;
      }
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) umma_ss_lo(tmem, kA + (kk >> 2) * (HALF >> 4) + (kk & 3) * 2, kB + (kk >> 2) * (HALF >> 4) + (kk & 3) * 2, i_kk, kk > 0);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) umma_ss_lo(tmem + 128, kA + (kk >> 2) * (HALF >> 4) + (kk & 3) * 2, kB + (kk >> 2) * (HALF >> 4) + (kk & 3) * 2, i_kk, kk > 0);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) umma_ts_lo(tmem + 256, tmem + (kk >> 2) * 64 + (kk & 3) * 8, nB + kk * 128, i_kn, kk > 0);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) umma_ss_lo(tmem + 128, nA + kk * 128, nB + kk * 128, i_nn, kk > 0);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) umma_ss_lo(tmem + 384, kA + (kk >> 2) * (HALF >> 4) + (kk & 3) * 2, nB + kk * 128, i_kn, kk > 0);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(bar);
    __syncwarp();
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (lane == 0) { cycles_out[blockIdx.x] = t1 - t0; *s_stop = 1; }
  } else if (kind != 0) {
    float x = 1.0f + lane * 1e-3f;
    uint32_t regs[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) regs[i] = i + lane;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) << 5) << 16;
    long long n = 0;
    while (*s_stop == 0 && n < (1ll << 24)) {
      ++n;
      if (kind >= 11 && kind <= 14) {
        // combinations: 11 = compute-like (LDTM, MUFU, STTM, LDTM, STS) on warps 0-7 + drain-like (LDTM x4, 128 reds) on 8-11
        //               12 = compute-like only      13 = drain-like only       14 = 11 without the reds (LDTM only on 8-11)
        if (warp < 8 && (kind == 11 || kind == 12 || kind == 14)) {
          tmem_ld_x32(tmem + lane_addr + 448, regs);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) { x = fast_exp2(x * 0.999f + __uint_as_float(regs[i] & 0xff)); }
          tmem_st_x32(tmem + lane_addr + 480, regs);
          tmem_st_wait();
          tmem_ld_x32(tmem + lane_addr + 448, regs);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<uint4*>(scratch + ((warp * 32 + lane) * 128 + ((i ^ (lane & 7)) << 4)) % 32768) = make_uint4(regs[0], regs[1], regs[2], regs[3]);
          fence_proxy_async_smem();
        } else if (warp >= 8 && warp < 12 && kind != 12) {
          uint32_t acc[128];
#pragma unroll
          for (int c = 0; c < 4; ++c) tmem_ld_x32(tmem + lane_addr + 384 + c * 32, acc + c * 32);
          tmem_ld_wait();
          if (kind != 14) {
            float* dst = gacc + ((static_cast<size_t>(blockIdx.x) * 64 + (n & 63)) * 16384) + (warp - 8) * 32 + lane;
#pragma unroll
            for (int i = 0; i < 128; ++i) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + i * 128), "f"(__uint_as_float(acc[i])) : "memory");
          } else if (acc[5] == 0xdeadbeefu) x += 1.f;
        } else {
          __nanosleep(200);
        }
      } else if (kind == 1 && warp < 8) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { x = fast_exp2(x * 0.999f); x = fmaf(x, 1.0001f, 0.5f); }
      } else if (kind == 2 && warp < 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<uint4*>(scratch + ((warp * 32 + lane) * 128 + ((i ^ (lane & 7)) << 4)) % 32768) = make_uint4(regs[0], regs[1], regs[2], regs[3]);
      } else if (kind == 3 && warp >= 8 && warp < 12) {
        float* dst = gacc + ((static_cast<size_t>(blockIdx.x) * 64 + (n & 63)) * 16384) + (warp - 8) * 4096 + lane;
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + i * 32), "f"(x) : "memory");
      } else if (kind == 4 && warp < 8) {
        tmem_ld_x32(tmem + lane_addr + 448 + 0 * (warp >> 2), regs);     // columns [448,480): outside every MMA's D / A range? (dK acc 384..511 is live: read-only)
        tmem_ld_wait();
      } else if (kind == 5 && warp < 12) {
        (void)mbar_try_wait(bar_never, 0);
      } else if (kind == 6 && warp < 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint32_t a, b, c, d;
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(smem_u32(scratch + i * 16)));
          regs[i] ^= a ^ b ^ c ^ d;
        }
      } else if (kind == 9 && warp == 14) {
        // 64 KB of bulk-async copies global -> shared per ~step, like the Q / dO loads
        if (lane == 0) {
          mbar_arrive_expect_tx(bar_never + 1, 65536);
          for (int c = 0; c < 4; ++c) bulk_load_1d(scratch, reinterpret_cast<const uint8_t*>(gacc) + (((n - 1) * 4 + c) % 4096) * 16384, 16384, bar_never + 1);
          mbar_wait(bar_never + 1, (n - 1) & 1);
        }
        __syncwarp();
      } else if (kind == 7 && warp < 8) {
        tmem_ld_x32(tmem + lane_addr + 448, regs);
        tmem_ld_wait();
        tmem_st_x32(tmem + lane_addr + 480, regs);
        tmem_st_wait();
      } else {
        __nanosleep(200);
      }
    }
    if (x == 123.456f || regs[3] == 0xdeadbeefu) cycles_out[0] = 0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 15) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---- mbarrier hand-off latency: warps 13 and 14 ping-pong over two mbarriers while the other warps interfere ----
__device__ __forceinline__ bool test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool try_wait_nohint(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void wait_mode(uint64_t* bar, uint32_t parity, int mode) {
  if (mode == 0) { while (!mbar_try_wait(bar, parity)) {} }
  else if (mode == 1) { while (!try_wait_nohint(bar, parity)) {} }
  else { while (!test_wait(bar, parity)) {} }
}
__global__ void __launch_bounds__(512, 1) pingpong_kernel(int kind, int mode, int rounds, float* gacc, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* scratch = smem;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
  volatile int* s_stop = reinterpret_cast<volatile int*>(bar + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); mbar_init(bar + 2, 1); *s_stop = 0; fence_mbar_init(); }
  __syncthreads();
  if (warp == 13) {
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      if (lane == 0) mbar_arrive(bar);
      wait_mode(bar + 1, r & 1, mode);
    }
    const long long t1 = clock64();
    if (lane == 0) { cycles_out[blockIdx.x] = t1 - t0; *s_stop = 1; }
  } else if (warp == 14) {
    for (int r = 0; r < rounds; ++r) {
      wait_mode(bar, r & 1, mode);
      if (lane == 0) mbar_arrive(bar + 1);
    }
  } else if (kind != 0) {
    float x = 1.0f + lane * 1e-3f;
    uint32_t regs[4] = {1u, 2u, 3u, 4u};
    long long n = 0;
    while (*s_stop == 0 && n < (1ll << 24)) {
      ++n;
      if (kind == 1 && warp < 8) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { x = fast_exp2(x * 0.999f); x = fmaf(x, 1.0001f, 0.5f); }
      } else if (kind == 2 && warp < 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<uint4*>(scratch + ((warp * 32 + lane) * 128 + ((i ^ (lane & 7)) << 4)) % 32768) = make_uint4(regs[0], regs[1], regs[2], regs[3]);
      } else if (kind == 3 && warp >= 8 && warp < 12) {
        float* dst = gacc + ((static_cast<size_t>(blockIdx.x) * 64 + (n & 63)) * 16384) + (warp - 8) * 4096 + lane;
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + i * 32), "f"(x) : "memory");
      } else if (kind == 13 && warp >= 8 && warp < 12) {
        float* dst = gacc + ((static_cast<size_t>(blockIdx.x) * 64 + (n & 63)) * 16384) + (warp - 8) * 32 + lane;
#pragma unroll
        for (int i = 0; i < 128; ++i) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + i * 128), "f"(x) : "memory");
      } else if (kind == 5 && warp < 12) {
        (void)mbar_try_wait(bar + 2, 0);
      } else {
        __nanosleep(200);
      }
    }
    if (x == 123.456f) cycles_out[0] = 0;
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d_cycles;
  cudaMalloc(&d_cycles, sms * sizeof(long long));
  float* gacc;
  cudaMalloc(&gacc, static_cast<size_t>(sms) * 64 * 65536);
  cudaMemset(gacc, 0, static_cast<size_t>(sms) * 64 * 65536);
  const int smem = 4 * HALF + 32768 + 64;
  cudaFuncSetAttribute(interf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 400;
  const char* names[] = {"none", "ALU+MUFU x8 warps", "STS.128 x8 warps", "global red x4 warps", "tcgen05.ld x8 warps",
                         "mbarrier try_wait polling x12 warps", "LDS.128 broadcast x8 warps", "tcgen05.ld+st x8 warps",
                         "commit after every GEMM", "64 KB bulk copies into smem per step", "kernel order: dV(TS) then S^T over its A operand",
                         "compute-like x8 + drain-like x4 (LDTM + 128 reds)", "compute-like x8 only", "drain-like x4 only (LDTM + 128 reds)",
                         "compute-like x8 + drain LDTM without reds"};
  for (int kind = 0; kind < 15; ++kind) {
    interf_kernel<<<sms, 512, smem>>>(kind, 20, gacc, d_cycles);
    interf_kernel<<<sms, 512, smem>>>(kind, iters, gacc, d_cycles);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kind %d: %s\n", kind, cudaGetErrorString(e)); return 1; }
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), d_cycles, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (auto c : h) mx = c > mx ? c : mx;
    printf("umma_interf kind %d (%s): %.1f cycles per MMA\n", kind, names[kind], static_cast<double>(mx) / (iters * 40.0));
  }
  cudaFuncSetAttribute(pingpong_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 64);
  const char* modes[] = {"try_wait + 10 ms suspend hint (kernels today)", "try_wait, default hint", "test_wait spin"};
  for (int kind : {0, 3, 13}) {
    for (int mode = 0; mode < 3; ++mode) {
      pingpong_kernel<<<sms, 512, 32768 + 64>>>(kind, mode, 2000, gacc, d_cycles);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("pingpong %d/%d: %s\n", kind, mode, cudaGetErrorString(e)); return 1; }
      std::vector<long long> h(sms);
      cudaMemcpy(h.data(), d_cycles, sms * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (auto c : h) mx = c > mx ? c : mx;
      printf("mbarrier hand-off, interference '%s', wait = %s: %.0f cycles per hand-off\n", names[kind], modes[mode],
             static_cast<double>(mx) / (2000 * 2.0));
    }
  }
  return 0;
}
