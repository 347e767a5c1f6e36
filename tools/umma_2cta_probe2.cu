// umma_2cta_probe2.cu — second probe for the CTA-pair attention backward (docs/next_kernels.md section 1): the three
// building blocks `umma_2cta.cu` left unverified.  One M = 256 `tcgen05.mma.cta_group::2` GEMM per mode, K = 64,
// small-integer bf16 operands (exact in fp32), compared with a host reference:
//   0  SS, A K-major (rows 0-127 in CTA 0, 128-255 in CTA 1), B MN-major ([K][64 n] 128B-swizzled blocks) split
//      along N: CTA r holds columns [r N/2, (r+1) N/2) for ALL k           (dK += dS^T Q;  N = 128 and 256)
//   1  TS, A = bf16x2 in each CTA's own TMEM lanes (written with tcgen05.st), B as in mode 0   (dV += P^T dO)
//   2  SS K-major / K-major (the split `umma_2cta.cu` verified), operands brought in by
//      `cp.async.bulk.tensor ... .cta_group::2` from BOTH CTAs with complete_tx on the LEADER's mbarrier (address
//      from `mapa`), plus one plain remote `mbarrier.arrive.shared::cluster` from the peer   (the pair's "full" barriers)
// diag = 1 / 2 replace the operands by patterns whose product spells out the column (D[m][n] = n) or the k / row
// mapping (D[m][n] = m % 64); on a FAIL the first rows of D are printed so the actual split can be read off.
// Every wait is bounded (build with -DDTA_MBAR_SPIN_LIMIT=256); run each mode as its own process under `timeout`.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -DDTA_MBAR_SPIN_LIMIT=256 -I dynamictreeattn_b200/csrc \
//        -I include tools/umma_2cta_probe2.cu dynamictreeattn_b200/csrc/dta_common.cu -lcuda -o tools/bin/umma_2cta_probe2
#include "dta_common.cuh"
#include "dta_ptx.cuh"
#include <cmath>
#include <cstdlib>
#include <vector>

using namespace dta;
static constexpr int kRows = 128;             // rows of A per CTA
static constexpr int kK = 64;                 // K extent (one 128-byte swizzled row of a K-major tile)
static constexpr int kABytes = kRows * 128;   // 16 KB
static constexpr int kBBlock = kK * 128;      // one MN-major block: [64 k][64 n] bf16 = 8 KB
static constexpr int kBBytes = 2 * kBBlock;   // up to 128 columns of B per CTA
static constexpr int kATmemCol = 256;         // TMEM column of the TS-mode A operand (D occupies [0, N))

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
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// TMA tile load of a CTA pair: data lands in the EXECUTING CTA's shared memory, the transaction bytes are
// signalled on `bar_cluster_addr`, which may be an mbarrier of the peer CTA (shared::cluster address).
__device__ __forceinline__ void tma_load_3d_2cta(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0,
                                                 int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

// element (row, k) of a K-major 128B-swizzled [rows][64 bf16] tile; the same image read MN-major is [k][n]
__device__ __forceinline__ uint32_t sw_off(uint32_t row, uint32_t col) { return sw128_offset(row, col >> 3) + (col & 7) * 2; }

__host__ __device__ inline float a_val(int diag, int m, int k) {
  if (diag == 1) return k == 0 ? 1.f : 0.f;
  if (diag == 2) return k == (m % 64) ? 1.f : 0.f;
  return static_cast<float>(((m * 7 + k * 3) % 9) - 4);
}
__host__ __device__ inline float b_val(int diag, int n, int k) {
  if (diag == 1) return k == 0 ? static_cast<float>(n) : 0.f;
  if (diag == 2) return static_cast<float>(k);
  return static_cast<float>(((n * 5 + k * 11) % 7) - 3);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe2_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, int mode, int n_dim,
              int diag, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + kABytes;
  uint64_t* bar_done = reinterpret_cast<uint64_t*>(smem + kABytes + kBBytes);
  uint64_t* bar_full = bar_done + 1;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_done + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int n_half = n_dim / 2;

  if (mode != 2) {
    if (mode == 0)
      for (int i = threadIdx.x; i < kRows * kK; i += blockDim.x) {
        const int r = i / kK, k = i % kK;
        *reinterpret_cast<__nv_bfloat16*>(sA + sw_off(r, k)) = __float2bfloat16(a_val(diag, rank * kRows + r, k));
      }
    // MN-major B: block j holds this CTA's columns [64 j, 64 j + 64), rows are k
    for (int i = threadIdx.x; i < kK * n_half; i += blockDim.x) {
      const int k = i / n_half, nl = i % n_half;
      *reinterpret_cast<__nv_bfloat16*>(sB + (nl >> 6) * kBBlock + sw_off(k, nl & 63)) =
          __float2bfloat16(b_val(diag, rank * n_half + nl, k));
    }
  }
  if (threadIdx.x == 0) {
    mbar_init(bar_done, 1);
    mbar_init(bar_full, 2);          // the leader's expect_tx arrival + the peer's remote arrival
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc_2cta(s_tmem, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (mode == 1) {
    // A rows of this CTA -> its own TMEM lanes, 32 columns of bf16x2 (low half = even k)
    uint32_t pk[32];
    const int m = rank * kRows + warp * 32 + lane;
#pragma unroll
    for (int c = 0; c < 32; ++c) pk[c] = pack_bf16x2(a_val(diag, m, 2 * c), a_val(diag, m, 2 * c + 1));
    tmem_st_x32(tmem + (static_cast<uint32_t>(warp << 5) << 16) + kATmemCol, pk);
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
  }
  if (mode == 2 && threadIdx.x == 0) {
    const uint32_t full_leader = mapa_u32(smem_u32(bar_full), 0);
    if (rank == 0) mbar_arrive_expect_tx(bar_full, 2 * (kABytes + 64 * 128));
    tma_load_3d_2cta(sA, &tm_a, full_leader, 0, 0, rank * kRows);       // A rows [128 r, 128 r + 128)
    tma_load_3d_2cta(sB, &tm_b, full_leader, 0, 0, rank * 64);          // B rows (n) [64 r, 64 r + 64)
    if (rank == 1) mbar_arrive_remote(full_leader);
  }
  if (warp == 0) {
    if (rank == 0) {
      if (mode == 2) {
        mbar_wait(bar_full, 0);
        tc_fence_after();
      }
      if (elect_one()) {
        const uint32_t kA = sdesc_lo_kmajor(smem_u32(sA)), kB = sdesc_lo_kmajor(smem_u32(sB));
        const uint32_t nB = sdesc_lo_mnmajor(smem_u32(sB), kBBlock);
        const uint32_t i_kk = make_idesc_bf16(256, n_dim, false, false), i_kn = make_idesc_bf16(256, n_dim, false, true);
#pragma unroll
        for (int kk = 0; kk < kK / 16; ++kk) {
          if (mode == 0) umma_ss_2cta(tmem, kA + kk * 2, nB + kk * 128, i_kn, kk > 0);
          else if (mode == 1) umma_ts_2cta(tmem, tmem + kATmemCol + kk * 8, nB + kk * 128, i_kn, kk > 0);
          else umma_ss_2cta(tmem, kA + kk * 2, kB + kk * 2, i_kk, kk > 0);
        }
        umma_commit_2cta(bar_done);
      }
      __syncwarp();
    }
    mbar_wait(bar_done, 0);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  {
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
    tmem_dealloc_2cta(tmem, 512);
  }
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const char* names[] = {"SS  A K-major, B MN-major split along N", "TS  A in TMEM,  B MN-major split along N",
                         "SS  K/K, TMA .cta_group::2 -> leader mbarrier + remote arrive"};
  if (mode < 0 || mode > 2) { printf("mode 0..2\n"); return 2; }
  const int smem = kABytes + kBBytes + 64;
  cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  float* d_out;
  cudaMalloc(&d_out, 256 * 256 * sizeof(float));
  __nv_bfloat16 *d_a, *d_b;
  cudaMalloc(&d_a, 256 * kK * sizeof(__nv_bfloat16));
  cudaMalloc(&d_b, 256 * kK * sizeof(__nv_bfloat16));
  int fails = 0;
  for (int n_dim : {128, 256}) {
    if (mode == 2 && n_dim == 256) continue;
    for (int diag = 0; diag < 3; ++diag) {
      std::vector<__nv_bfloat16> ha(256 * kK), hb(256 * kK);
      for (int m = 0; m < 256; ++m)
        for (int k = 0; k < kK; ++k) {
          ha[m * kK + k] = __float2bfloat16(a_val(diag, m, k));
          hb[m * kK + k] = __float2bfloat16(b_val(diag, m, k));
        }
      cudaMemcpy(d_a, ha.data(), ha.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice);
      cudaMemcpy(d_b, hb.data(), hb.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice);
      CUtensorMap tm_a, tm_b;
      // (out, base, rows, heads, head_dim, row_stride, head_stride, box_rows): box = 64 columns x box_rows rows
      if (make_tmap_bf16_rows(&tm_a, d_a, 256, 1, kK, kK, kK, kRows) || make_tmap_bf16_rows(&tm_b, d_b, 256, 1, kK, kK, kK, 64)) {
        printf("tensor map failed: %s\n", last_error_buf());
        return 2;
      }
      cudaMemset(d_out, 0, 256 * 256 * sizeof(float));
      probe2_kernel<<<2, 128, smem>>>(tm_a, tm_b, mode, n_dim, diag, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("probe2 mode %d N=%d diag %d: %s\n", mode, n_dim, diag, cudaGetErrorString(e)); return 1; }
      std::vector<float> h(256 * n_dim);
      cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
      double err = 0, err_sw = 0;
      for (int m = 0; m < 256; ++m)
        for (int n = 0; n < n_dim; ++n) {
          float ref = 0, ref_sw = 0;
          for (int k = 0; k < kK; ++k) {
            ref += a_val(diag, m, k) * b_val(diag, n, k);
            ref_sw += a_val(diag, m, k) * b_val(diag, (n + n_dim / 2) % n_dim, k);
          }
          err = fmax(err, fabs(h[m * n_dim + n] - ref));
          err_sw = fmax(err_sw, fabs(h[m * n_dim + n] - ref_sw));
        }
      printf("umma_2cta_probe2 mode %d (%s) N=%d diag %d: max_abs_err %.3g (B halves swapped: %.3g) %s\n", mode, names[mode],
             n_dim, diag, err, err_sw, err == 0 ? "PASS" : "FAIL");
      if (err != 0) {
        ++fails;
        for (int m : {0, 1, 128, 129}) {
          printf("  D[%3d][0..%d):", m, n_dim);
          for (int n = 0; n < n_dim; ++n) printf(" %g", h[m * n_dim + n]);
          printf("\n");
        }
      }
    }
  }
  return fails ? 1 : 0;
}
