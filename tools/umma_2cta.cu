// umma_2cta.cu — probe + rate of tcgen05.mma.cta_group::2 (CTA pair, M = 256) for the next step of the
// attention backward kernel (the single-CTA kernel is bound by the shared-memory data pipe; a
// CTA pair fetches each B operand once for 256 rows of A).
//
//   probe : D[256 x N] = A[256 x 64] B[N x 64]^T with small-integer bf16 operands, A rows 0-127 in CTA 0's
//           shared memory and 128-255 in CTA 1's, B rows (N index) 0..N/2-1 in CTA 0 and N/2..N-1 in CTA 1
//           (each CTA holds HALF of B), every tile K-major / 128B-swizzled like the attention kernels' tiles.
//           Prints the max error against that hypothesis (0 = the operand split is understood).
//   rate  : cycles per MMA for back-to-back M=256 MMAs (N = 128, 256) from the leader CTA, with and without
//           eight warps per CTA streaming STS.128 (the interference that throttles the single-CTA kernel's
//           dependency chain), next to the cta_group::1 M=128 N=128 rate in the same harness.
//   mix   : the backward step's five GEMMs per iteration (S^T, dP^T SS K-major; dV TS; dQ^T SS MN/MN; dK SS K/MN; 40
//           MMA-equivalents per SM), once as five cta_group::1 GEMMs per CTA and once as the CTA-pair version (S^T,
//           dP^T, dV, dK as M = 256 cta_group::2 MMAs issued by the leader, dQ^T as one M = 128 MMA per CTA), with
//           eight warps per CTA streaming STS.128: cycles per MMA-equivalent AND the shared-memory store bandwidth the
//           stream still gets -- the quantity that bounds the real kernel (its LSU/TMA traffic is served in the gaps).
// Every wait is bounded (mbar_wait traps); run under `timeout`.
#include "dta_common.cuh"
#include "dta_ptx.cuh"
#include <cmath>
#include <vector>

using namespace dta;
static constexpr int kRows = 128;            // rows of A per CTA
static constexpr int kK = 64;                // K extent of the probe tiles (one 128-byte swizzled row)
static constexpr int kABytes = kRows * 128;  // 16 KB
static constexpr int kBBytes = 128 * 128;    // room for up to 128 rows of B per CTA (N = 256)

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_ss_2cta(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(kSdescHi)
      : "memory");
}
__device__ __forceinline__ void umma_ts_2cta(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(acc), "r"(kSdescHi)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {   // arrives on `bar` of BOTH CTAs of the pair
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// element (row, k) of a K-major 128B-swizzled [rows][64 bf16] tile
__device__ __forceinline__ uint32_t sw_off(uint32_t row, uint32_t k) { return sw128_offset(row, k >> 3) + (k & 7) * 2; }
__host__ __device__ inline float a_val(int m, int k) { return static_cast<float>(((m * 7 + k * 3) % 9) - 4); }
__host__ __device__ inline float b_val(int n, int k) { return static_cast<float>(((n * 5 + k * 11) % 7) - 3); }

// mode 0: probe (writes D to out[256][n_dim]);  mode 1: rate, cta_group::2;  mode 2: rate, cta_group::1 (each CTA alone)
// mode 3: backward mix, cta_group::1 per CTA;  mode 4: backward mix as a CTA pair
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
pair_kernel(int mode, int n_dim, int iters, int interfere, float* out, long long* cycles_out, long long* sts_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + kABytes;
  uint8_t* scratch = smem + kABytes + kBBytes;                        // 32 KB for the STS stream
  uint64_t* bar = reinterpret_cast<uint64_t*>(scratch + 32768);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar + 2);
  volatile int* s_stop = reinterpret_cast<volatile int*>(bar + 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int n_half = n_dim / 2;

  // operands: this CTA's 128 rows of A, and its half of B
  for (int i = threadIdx.x; i < kRows * kK; i += blockDim.x) {
    const int r = i / kK, k = i % kK;
    *reinterpret_cast<__nv_bfloat16*>(sA + sw_off(r, k)) = __float2bfloat16(a_val(rank * kRows + r, k));
  }
  const int b_rows = (mode == 2 || mode >= 3) ? 128 : n_half;         // single-CTA MMAs need all of B locally
  for (int i = threadIdx.x; i < b_rows * kK; i += blockDim.x) {
    const int r = i / kK, k = i % kK;
    *reinterpret_cast<__nv_bfloat16*>(sB + sw_off(r, k)) = __float2bfloat16(b_val((mode == 2 ? 0 : rank * n_half) + r, k));
  }
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    *s_stop = 0;
    fence_mbar_init();
  }
  if (warp == 0) {
    if (mode == 2 || mode == 3) tmem_alloc(s_tmem, 512); else tmem_alloc_2cta(s_tmem, 512);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (mode >= 3 && warp == 0) {
    // ---- backward-step mix.  Tiles: sA = [128][64] K-major (also read MN-major), sB likewise; K = 128 is emulated by
    // two passes over the 64-wide tile (the operand traffic per MMA is what matters here, not the values).
    const uint32_t kA = sdesc_lo_kmajor(smem_u32(sA)), kB = sdesc_lo_kmajor(smem_u32(sB));
    const uint32_t nA = sdesc_lo_mnmajor(smem_u32(sA), kABytes), nB = sdesc_lo_mnmajor(smem_u32(sB), kABytes);
    constexpr uint32_t i1_kk = make_idesc_bf16(128, 128, false, false), i1_kn = make_idesc_bf16(128, 128, false, true),
                       i1_nn = make_idesc_bf16(128, 128, true, true);
    constexpr uint32_t i2_kk = make_idesc_bf16(256, 128, false, false), i2_kn = make_idesc_bf16(256, 128, false, true);
    const bool pair = mode == 4;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
        if (!pair || rank == 0) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {      // S^T
            if (pair) umma_ss_2cta(tmem, kA + (kk & 3) * 2, kB + (kk & 3) * 2, i2_kk, kk > 0);
            else umma_ss_lo(tmem, kA + (kk & 3) * 2, kB + (kk & 3) * 2, i1_kk, kk > 0);
          }
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {      // dP^T
            if (pair) umma_ss_2cta(tmem + 128, kA + (kk & 3) * 2, kB + (kk & 3) * 2, i2_kk, kk > 0);
            else umma_ss_lo(tmem + 128, kA + (kk & 3) * 2, kB + (kk & 3) * 2, i1_kk, kk > 0);
          }
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {      // dV += P^T dO   (A in TMEM, B MN-major)
            if (pair) umma_ts_2cta(tmem + 256, tmem + (kk >> 2) * 64 + (kk & 3) * 8, nB + (kk & 3) * 128, i2_kn, kk > 0);
            else umma_ts_lo(tmem + 256, tmem + (kk >> 2) * 64 + (kk & 3) * 8, nB + (kk & 3) * 128, i1_kn, kk > 0);
          }
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)          // dQ^T: always one M = 128 MMA stream per CTA
          umma_ss_lo(tmem + 128, nA + (kk & 3) * 128, nB + (kk & 3) * 128, i1_nn, kk > 0);
        if (!pair || rank == 0) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {      // dK += dS^T Q   (A K-major, B MN-major)
            if (pair) umma_ss_2cta(tmem + 384, kA + (kk & 3) * 2, nB + (kk & 3) * 128, i2_kn, kk > 0);
            else umma_ss_lo(tmem + 384, kA + (kk & 3) * 2, nB + (kk & 3) * 128, i1_kn, kk > 0);
          }
        }
      }
      __syncwarp();
    }
    if (elect_one()) {
      umma_commit(bar + 1);                                   // this CTA's own cta_group::1 MMAs
      if (pair && rank == 0) umma_commit_2cta(bar);           // the pair's MMAs, both CTAs' barriers
      if (!pair) umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar + 1, 0);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (lane == 0) {
      cycles_out[blockIdx.x] = t1 - t0;
      *s_stop = 1;
    }
  } else if (warp == 0) {
    const bool issuer = mode == 2 || rank == 0;
    const uint32_t kA = sdesc_lo_kmajor(smem_u32(sA)), kB = sdesc_lo_kmajor(smem_u32(sB));
    const uint32_t idesc = mode == 2 ? make_idesc_bf16(128, n_dim, false, false) : make_idesc_bf16(256, n_dim, false, false);
    const long long t0 = clock64();
    if (issuer) {
      // 8 accumulation chains (4 k-steps each) per elected block: the issuing lane must not be the bottleneck
      for (int it = 0; it < iters; it += (mode == 0 ? 1 : 8)) {
        if (elect_one()) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (mode == 0 && u > 0) break;
            const uint32_t d = tmem + (mode == 0 ? 0 : (u & 1) * (n_dim > 128 ? 256 : 128));
#pragma unroll
            for (int kk = 0; kk < kK / 16; ++kk) {
              if (mode == 2) umma_ss_lo(d, kA + kk * 2, kB + kk * 2, idesc, kk > 0);
              else umma_ss_2cta(d, kA + kk * 2, kB + kk * 2, idesc, kk > 0);
            }
          }
        }
        __syncwarp();
      }
      if (elect_one()) {
        if (mode == 2) umma_commit(bar); else umma_commit_2cta(bar);
      }
      __syncwarp();
    }
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (lane == 0) {
      cycles_out[blockIdx.x] = t1 - t0;
      *s_stop = 1;
    }
  } else if (mode != 0 && interfere && warp >= 4) {
    uint32_t v = lane;
    long long n = 0;
    while (*s_stop == 0 && n < (1ll << 22)) {
      ++n;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<uint4*>(scratch + (((warp - 4) * 32 + lane) * 128 + ((i ^ (lane & 7)) << 4)) % 32768) = make_uint4(v, v, v, v);
    }
    if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(sts_out + blockIdx.x), static_cast<unsigned long long>(n) * 8 * 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (mode == 0 && warp < 4) {
    // each CTA reads its own 128 accumulator rows
    const uint32_t lane_addr = static_cast<uint32_t>(warp << 5) << 16;
    for (int c = 0; c < n_dim; c += 32) {
      uint32_t v[32];
      tmem_ld_x32(tmem + lane_addr + c, v);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) out[(rank * kRows + warp * 32 + lane) * n_dim + c + i] = __uint_as_float(v[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if (mode == 2 || mode == 3) tmem_dealloc(tmem, 512); else tmem_dealloc_2cta(tmem, 512);
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem = kABytes + kBBytes + 32768 + 64;
  cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  float* d_out;
  long long* d_cyc;
  cudaMalloc(&d_out, 256 * 256 * sizeof(float));
  cudaMalloc(&d_cyc, 2 * sms * sizeof(long long));
  long long* d_sts;
  cudaMalloc(&d_sts, 2 * sms * sizeof(long long));
  // ---- probe
  for (int n_dim : {128, 256}) {
    cudaMemset(d_out, 0, 256 * 256 * sizeof(float));
    pair_kernel<<<2, 384, smem>>>(0, n_dim, 1, 0, d_out, d_cyc, d_sts);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("probe N=%d: %s\n", n_dim, cudaGetErrorString(e)); return 1; }
    std::vector<float> h(256 * n_dim);
    cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
    double err = 0, err_swapped = 0;
    for (int m = 0; m < 256; ++m)
      for (int n = 0; n < n_dim; ++n) {
        float ref = 0, ref_sw = 0;
        for (int k = 0; k < kK; ++k) {
          ref += a_val(m, k) * b_val(n, k);
          ref_sw += a_val(m, k) * b_val((n + n_dim / 2) % n_dim, k);     // hypothesis: the B halves are swapped
        }
        err = fmax(err, fabs(h[m * n_dim + n] - ref));
        err_swapped = fmax(err_swapped, fabs(h[m * n_dim + n] - ref_sw));
      }
    printf("umma_2cta probe M=256 N=%d K=64: max_abs_err %.3g (B halves swapped: %.3g) %s\n", n_dim, err, err_swapped,
           err == 0 ? "PASS" : "FAIL");
  }
  // ---- rate
  const int iters = 2048;
  struct Case { int mode, n, interfere; const char* name; };
  const Case cases[] = {{2, 128, 0, "cta_group::1 M=128 N=128"},          {2, 128, 1, "cta_group::1 M=128 N=128 + STS.128 x8 warps"},
                        {1, 128, 0, "cta_group::2 M=256 N=128"},          {1, 128, 1, "cta_group::2 M=256 N=128 + STS.128 x8 warps"},
                        {1, 256, 0, "cta_group::2 M=256 N=256"},          {1, 256, 1, "cta_group::2 M=256 N=256 + STS.128 x8 warps"}};
  for (const Case& c : cases) {
    const int grid = (sms / 2) * 2;
    pair_kernel<<<grid, 384, smem>>>(c.mode, c.n, 50, c.interfere, d_out, d_cyc, d_sts);
    pair_kernel<<<grid, 384, smem>>>(c.mode, c.n, iters, c.interfere, d_out, d_cyc, d_sts);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("rate %s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (auto v : h) mx = v > mx ? v : mx;
    const double per_mma = static_cast<double>(mx) / (iters * 4.0);
    const double flops_per_mma_per_sm = 2.0 * 128 * c.n * 16;             // per SM: 128 rows of D either way
    printf("umma_2cta rate %-46s: %.1f cycles per MMA, %.0f flop/clk/SM (8192 = peak), smem operand bytes/clk/SM %.0f\n", c.name,
           per_mma, flops_per_mma_per_sm / per_mma, (c.mode == 2 ? (128 + c.n) : (128 + c.n / 2)) * 32.0 / per_mma);
  }
  // ---- backward mix: five GEMMs per step, single-CTA vs CTA pair, with and without the STS stream
  for (int mode : {3, 4}) {
    for (int interfere : {0, 1}) {
      const int grid = (sms / 2) * 2, mix_iters = 400;
      pair_kernel<<<grid, 384, smem>>>(mode, 128, 20, interfere, d_out, d_cyc, d_sts);
      cudaMemset(d_sts, 0, 2 * sms * sizeof(long long));
      pair_kernel<<<grid, 384, smem>>>(mode, 128, mix_iters, interfere, d_out, d_cyc, d_sts);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mix mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      std::vector<long long> h(grid), st(grid);
      cudaMemcpy(h.data(), d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      cudaMemcpy(st.data(), d_sts, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0, sts_min = 1ll << 62;
      for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; sts_min = st[i] < sts_min ? st[i] : sts_min; }
      printf("umma_2cta mix %-14s %s: %.1f cycles per MMA-equivalent (64 = peak), STS stream gets %.1f B/clk/SM\n",
             mode == 3 ? "cta_group::1" : "CTA pair", interfere ? "+ STS.128 x8 warps" : "alone             ",
             static_cast<double>(mx) / (mix_iters * 40.0), interfere ? static_cast<double>(sts_min) / mx : 0.0);
    }
  }
  return 0;
}
