// pair_handoff_bench.cu — latency of the CTA-pair hand-offs (docs/next_kernels.md 1.2).  NOT YET RUN (written after the
// round-1 GPU budget was spent); first visit of round 2.
//
// The pair backward kernel's leader waits ~370-930 cycles for the peer's `mbarrier.arrive.shared::cluster` after the
// peer's warps are done, while the leader's `tcgen05.commit ... multicast` reaches the peer within ~20 cycles
// (profiles/r01/bwd_pair_trace.log).  This measures the hand-off alone, as a ping-pong between the two CTAs of a
// cluster timed on the LEADER's clock (per-SM clocks are not comparable), `iters` round trips:
//   mode 0  arrive / arrive   : leader -> peer and peer -> leader by mbarrier.arrive.shared::cluster (default .release.cta)
//   mode 1  same with .release.cluster                                      (the form that cost 1000+ cycles in the kernel)
//   mode 2  commit / arrive   : leader -> peer by tcgen05.commit.cta_group::2 multicast (mask 2, no MMA outstanding),
//                               peer -> leader by mbarrier.arrive
//   mode 3  commit / commit   : both directions by tcgen05.commit.cta_group::2 multicast (peer -> leader: mask 1).
//                               Answers: may a NON-leader thread issue a cta_group::2 commit, and is it faster?
// `load` = 1 adds the traffic the real kernel has in flight: warps 4-7 of both CTAs stream 16-byte global reds
// (stand-in for the dQ TMA reduce-adds) while the ping-pong runs.
// Every wait is bounded by time; run each mode as its own process under `timeout` (mode 3 may be illegal and trap).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I dynamictreeattn_b200/csrc -I include \
//        tools/pair_handoff_bench.cu dynamictreeattn_b200/csrc/dta_common.cu -lcuda -o tools/bin/pair_handoff_bench
#include "dta_common.cuh"
#include "dta_ptx.cuh"
#include <cstdlib>
#include <vector>

using namespace dta;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void arrive_remote(uint32_t addr, bool cluster_release) {
  if (cluster_release)
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
  else
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ bool wait_bounded(uint64_t* bar, uint32_t parity) {   // false = timed out (~1 s)
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > (1ll << 31)) return false;
  return true;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
handoff_kernel(int mode, int iters, int load, float* sink, long long* cycles_out, int* status) {
  __shared__ __align__(8) uint64_t bar[2];            // bar[0]: "ping" (waited by the peer), bar[1]: "pong" (waited by the leader)
  __shared__ uint32_t s_tmem;
  __shared__ volatile int s_stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    s_stop = 0;
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc_2cta(&s_tmem, 32);         // the commit forms need the pair's tensor-core context
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();

  if (warp == 1 && lane == 0) {
    const bool cl = mode == 1;
    const uint32_t ping_peer = mapa_u32(smem_u32(bar), 1), pong_leader = mapa_u32(smem_u32(bar + 1), 0);
    bool ok = true;
    if (rank == 0) {
      const long long t0 = clock64();
      for (int it = 0; it < iters && ok; ++it) {
        if (mode >= 2) commit_multicast(bar, 2);        // leader -> peer's bar[0]
        else arrive_remote(ping_peer, cl);
        ok = wait_bounded(bar + 1, it & 1);
      }
      cycles_out[blockIdx.x >> 1] = clock64() - t0;
      if (!ok) status[0] = 1;
    } else {
      for (int it = 0; it < iters && ok; ++it) {
        ok = wait_bounded(bar, it & 1);
        if (mode == 3) commit_multicast(bar + 1, 1);    // peer -> leader's bar[1]
        else arrive_remote(pong_leader, cl);
      }
      if (!ok) status[0] = 2;
    }
    s_stop = 1;
  } else if (load && warp >= 4) {
    // background: 16-byte reds to a 1 MB window per CTA (L2-resident), as the drain warps' reduce-adds would be
    float* base = sink + (static_cast<size_t>(blockIdx.x) << 18);
    uint32_t i = threadIdx.x;
    long long n = 0;
    while (s_stop == 0 && n < (1ll << 20)) {
      ++n;
      float* a = base + ((i * 4u) & 0x3FFFCu);
      asm volatile("red.global.v4.f32.add [%0], {%1, %1, %1, %1};" ::"l"(a), "f"(1.0f) : "memory");
      i = i * 1664525u + 1013904223u;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc_2cta(s_tmem, 32);
  }
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const char* names[] = {"arrive(.release.cta) / arrive", "arrive(.release.cluster) / arrive", "commit multicast / arrive",
                         "commit multicast / commit multicast (peer-issued)"};
  if (mode < 0 || mode > 3) { printf("mode 0..3\n"); return 2; }
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* d_sink;
  long long* d_cyc;
  int* d_status;
  cudaMalloc(&d_sink, static_cast<size_t>(sms) << 20);
  cudaMemset(d_sink, 0, static_cast<size_t>(sms) << 20);
  cudaMalloc(&d_cyc, sms * sizeof(long long));
  cudaMalloc(&d_status, sizeof(int));
  const int iters = 2000;
  for (int load = 0; load < 2; ++load)
    for (int grid : {2, (sms / 2) * 2}) {
      cudaMemset(d_status, 0, sizeof(int));
      handoff_kernel<<<grid, 256>>>(mode, 100, load, d_sink, d_cyc, d_status);
      handoff_kernel<<<grid, 256>>>(mode, iters, load, d_sink, d_cyc, d_status);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("pair_handoff mode %d (%s): %s\n", mode, names[mode], cudaGetErrorString(e)); return 1; }
      int st = 0;
      cudaMemcpy(&st, d_status, sizeof(int), cudaMemcpyDeviceToHost);
      std::vector<long long> h(grid / 2);
      cudaMemcpy(h.data(), d_cyc, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0, sum = 0;
      for (auto c : h) { mx = c > mx ? c : mx; sum += c; }
      printf("pair_handoff mode %d (%-50s) clusters=%3d load=%d: round trip mean %6.0f max %6.0f cycles%s\n", mode, names[mode],
             grid / 2, load, static_cast<double>(sum) / h.size() / iters, static_cast<double>(mx) / iters,
             st ? "  TIMED OUT (hand-off form not delivered)" : "");
    }
  return 0;
}
