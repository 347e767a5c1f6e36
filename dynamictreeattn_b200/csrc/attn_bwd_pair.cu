// attn_bwd_pair.cu — CTA-PAIR variant of the tree attention backward main kernel (tcgen05 cta_group::2).
//
// The DEFAULT backward for launches the caller marks balanced (dta_attn_shape_t::bwd_whole_tiles: tau2, chains, deep
// fans); the single-CTA work-list kernel (attn_bwd.cu) serves bushy tries, whose long-pole key tiles it splits, and
// DTA_BWD_PAIR=0|1 forces either.  Measured in the power-capped step on one box: 921-941 TFLOP/s against 891-905 for
// the single-CTA kernel.  Same mathematics, same pre-pass (attn_delta_kernel) and epilogue
// (attn_dq_convert_kernel), same TMEM plan and warp roles as attn_bwd.cu -- read that file's header first.  What
// changes (docs/next_kernels.md section 1; building blocks verified by tools/umma_2cta.cu and tools/umma_2cta_probe2.cu):
//
//   cluster (2,1,1) = key tiles (2j, 2j+1) x one KV head.  Both CTAs walk the UNION of the two tiles' live query
//   ranges in lock-step (a step that one of the tiles cannot see is fully masked for that CTA: P^T = 0).
//   Four of the five GEMMs of a step become ONE M = 256 MMA issued by the leader (cluster rank 0); their B operand
//   is shared by the pair, each CTA holding half of it, so every SM fetches half the B bytes per step:
//     S^T  = K Q^T    A = K  (own tile)       B = Q  K-major,  split along N = queries: rank r holds [64r, 64r+64)
//     dP^T = V dO^T   A = V                   B = dO K-major,  same split
//     dV  += P^T dO   A = P^T (own TMEM)      B = dO MN-major, split along N = head dim: rank r holds d [64r, 64r+64)
//     dK  += dS^T Q   A = dS^T (own smem)     B = Q  MN-major, same split
//     dQ^T = K^T dS^T contracts over keys, which the pair splits: stays one M = 128 cta_group::1 MMA per CTA.
//   A pair MMA reads B at the SAME shared-memory offset in both CTAs, so the two views of a Q (dO) tile cannot
//   overlap the way a single CTA's do: a stage holds [own 64 queries x 128 d] (two K-major blocks) and
//   [128 queries x own 64 d] (one MN-major block) = 32 KB, as before; the saving is in operand fetches.
//   Barriers: everything the leader's issuers wait for covers both CTAs -- TMA loads of both CTAs signal the
//   LEADER's "full" barriers (cp.async.bulk.tensor .cta_group::2), compute / drain warps of both CTAs arrive on the
//   leader's barriers (mbarrier.arrive.shared::cluster), and the leader's tcgen05.commit multicasts to both CTAs.
//   Schedule: whole pairs only (no split chunks yet), heaviest first in k_order's sequence.
#include "dta_common.cuh"
#include "dta_ptx.cuh"
#include "../../include/dta_b200.h"
#include <climits>
#include <cstdlib>
#include <cstring>

namespace dta {
namespace bwd_pair {

static constexpr int kT = 128;
static constexpr int kHalf = kT * 128;               // [128 rows][64 bf16] swizzled block, bytes
static constexpr int kTile = 2 * kHalf;              // 32 KB
static constexpr int kBlk = 64 * 128;                // [64 rows][64 bf16] swizzled block, bytes
static constexpr int kThreads = 512;
static constexpr int kProducerWarp = 14, kMmaWarp = 15, kMmaWarpB = 13;
static constexpr int kQStages = 2;
static constexpr int kDqBufs = 3;
static constexpr int kDqRows = 16;
static constexpr int kAugA = kT * 16;                // un-swizzled [128 rows][8 bf16] chunk (ones / zeros), bytes
static constexpr int kAugB = 64 * 16;                // this CTA's 64 query rows of the lse / delta operand

struct Smem {
  static constexpr int k = 0;
  static constexpr int v = k + kTile;
  static constexpr int q = v + kTile;                 // kQStages x [K-major: 2 blocks of [64 own q][64 d] | MN-major: [128 q][64 own d]]
  static constexpr int d_o = q + kQStages * kTile;    // one stage, same two views
  static constexpr int ds = d_o + kTile;
  static constexpr int aug_q = ds + kTile;
  static constexpr int aug_do = aug_q + kQStages * kAugB;
  static constexpr int aug_one = aug_do + kAugB;
  static constexpr int aug_zero = aug_one + kAugA;
  static constexpr int dq_stage = aug_zero + kAugA;
  static constexpr int bars = dq_stage + 4 * kDqBufs * kDqRows * 128;
  static constexpr int n_bars = 16;
  static constexpr int misc = bars + n_bars * 8;
  static constexpr int total = misc + 16;
};
static_assert(Smem::total <= 232448, "pair backward kernel shared memory exceeds 227 KB");
static_assert(Smem::dq_stage % 128 == 0, "dQ staging must be 128-byte aligned for the TMA reduce");

struct Params {
  float* dq_accum;
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
  const int32_t* subtree_end;
  const int32_t* kmin_end;
  const int32_t* k_qlast;
  const int32_t* pair_order;   // [n_pairs] pair index j (key tiles 2j, 2j+1), heaviest first
  int n_nodes, n_pad, n_tiles, n_q_heads, n_kv_heads, group;
  int64_t k_row_stride, k_head_stride, v_row_stride, v_head_stride;
  float scale_log2, inv_scale;
  // lane of a drain warp that sends "R1 drained" to the leader.  Lane 0 also issues the warp's TMA reduces, and its
  // arrival takes ~930 cycles to complete the leader's barrier against ~370 for the P^T arrival of a compute warp
  // (docs/next_kernels.md 1.2); DTA_PAIR_DRAIN_LANE picks another lane for the A/B (any lane may arrive after the
  // __syncwarp).  Default 0 = the measured configuration.
  int drain_lane;
  int handoff_commit;          // bit mask (P^T, dS^T, R1 free): that hand-off to the leader uses commit_signal_leader instead of a remote arrive
  long long* trace;            // debug timeline (NULL in production): [step < 48][cluster rank][16] clock64 stamps of cluster 0
};

// per-SM clocks: stamps of the two CTAs are comparable only within a CTA
#define DTA_PTRACE(ev)                                                                        \
  do {                                                                                        \
    if (p.trace != nullptr && blockIdx.x < 2 && it < 48)                                      \
      p.trace[(it * 2 + blockIdx.x) * 16 + (ev)] = clock64();                                 \
  } while (0)

// ---- cluster / cta_group::2 wrappers (forms verified on the B200 by the two probes) ----------------------------
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
__device__ __forceinline__ void umma2_ss(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                         uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(hi)
      : "memory");
}
__device__ __forceinline__ void umma2_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(acc), "r"(kSdescHi)
      : "memory");
}
// arrives on `bar` (same shared-memory offset) of BOTH CTAs of the pair when the issuing thread's MMAs retire
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// Same instruction used as a pure SIGNAL to the leader (mask 1) by a thread that has no MMA outstanding: the commit
// path delivered the leader's barriers to the peer within ~20 cycles where a remote mbarrier.arrive took 370-930
// (docs/next_kernels.md 1.2).  A/B variant (DTA_PAIR_HANDOFF=commit), measured on the B200 in round 2: a NON-leader
// thread may issue it (tools/pair_handoff_bench.cu mode 3: 427 cycles round trip, load-independent, against 715-2590
// for the arrive), gradients identical to the arrive form on every parity case -- and the kernel is 3 % SLOWER with it
// (tau2: 14.16 vs 13.70 ms, profiles/r02/bwd_pair_sweep.log): the commit orders behind the issuing thread's proxy
// traffic.  Kept for the A/B and covered by tests/test_gpu_kernels.py::test_bwd_cta_pair_commit_handoff_variant; not the default.
__device__ __forceinline__ void commit_signal_leader(uint64_t* bar) {
  const uint16_t mask = 1;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// TMA tile loads of a pair: data into the EXECUTING CTA's shared memory, bytes signalled on `bar_cluster_addr`
// (a shared::cluster address; here always the leader's barrier)
__device__ __forceinline__ void tma2_load_3d(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1,
                                             int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
// Arrival on a barrier of the pair's leader.  kClusterFences = false: the default .release.cta / .acquire.cta forms,
// as CUTLASS uses for the same hand-offs (ClusterBarrier::arrive, umma_arrive_2x1SM_sm0 in cutlass/arch/barrier.h).  What the leader's MMAs read from the peer is SM-local state
// of the peer (its TMEM after tcgen05.wait::st, its shared memory after the STS + fence.proxy.async), fetched by the
// peer's own tensor core; the first version's `.release.cluster` arrive made the arriving lane wait ~1000-1600 cycles
// on every one of the three per-step hand-offs (P^T, dS^T, R1 free) with TMA reduces and stores in flight
// (profiles/r01/bwd_pair_trace_release_cluster.log: step period 5735 cycles against the single-CTA kernel's 4356;
// 4074 with this form, profiles/r01/bwd_pair_trace.log).
static constexpr bool kClusterFences = false;
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  if (kClusterFences)
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
  else
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
// Waits of this kernel are bounded by TIME (~2 s of SM clock), not by a try count: a protocol bug must trap and
// surface as a launch error, never hang the box.  kCluster: the barrier receives arrivals from the peer CTA
// (acquire at cluster scope when kClusterFences, else the default form like every other wait).
template <bool kCluster>
__device__ __forceinline__ void pwait_impl(uint64_t* bar, uint32_t parity) {
  long long t0 = 0;
  for (;;) {
    uint32_t ok;
    if (kCluster && kClusterFences)
      asm volatile(
          "{\n\t.reg .pred P;\n\t"
          "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2, %3;\n\t"
          "selp.u32 %0, 1, 0, P;\n\t}\n"
          : "=r"(ok)
          : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
          : "memory");
    else
      ok = mbar_try_wait(bar, parity) ? 1u : 0u;
    if (ok) break;
    const long long now = clock64();
    if (t0 == 0) t0 = now;
    if (now - t0 > (1ll << 32)) {
      printf("dta: pair mbarrier timeout block=%d thread=%d bar=%d parity=%u\n", blockIdx.x, threadIdx.x,
             static_cast<int>(((smem_u32(bar) - Smem::bars) & 1023u) / 8), parity);   // barrier index (smem base is 1 KB aligned)
      __trap();
    }
  }
}
__device__ __forceinline__ void pwait(uint64_t* bar, uint32_t parity) { pwait_impl<false>(bar, parity); }
__device__ __forceinline__ void pwait_cluster(uint64_t* bar, uint32_t parity) { pwait_impl<true>(bar, parity); }

// kPolyOf8: how many of every 8 exponentials of the P^T pass run on the FMA pipes (poly_exp2), as in attn_bwd.cu
// kHeadMajor: launch order -- every key-tile pair of KV head 0, then head 1, ... (Q / dO / dQ of one head pair in L2 at a
// time) instead of the KV heads of a pair side by side; compile-time so that the step loop's code is the same
static constexpr bool kDefaultPairHeadMajor = false;
static constexpr int kDefaultHandoff = 0;            // bit mask of hand-offs sent as commit signals (see Params::handoff_commit)
// kTmaEpilogue: dV / dK leave through the (dead) Q stages and a TMA store per tensor instead of one 128-byte row per
// thread straight to global memory (32 rows per store instruction; see attn_bwd.cu's epilogue).  Default since the
// in-step A/B of profiles/r02/bench_ab10_*.json (one box, alternating): 781.2 / 782.5 ms per step against 784.9 / 785.8;
// alone the two are equal within noise.  DTA_PAIR_TMA_EPILOGUE=0 selects the row stores.
static constexpr bool kDefaultPairTmaEpilogue = true;
template <int kPolyOf8, bool kHeadMajor, bool kTmaEpilogue>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
tree_attn_bwd_pair_kernel(const __grid_constant__ CUtensorMap tmap_q64, const __grid_constant__ CUtensorMap tmap_q128,
                          const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                          const __grid_constant__ CUtensorMap tmap_do64, const __grid_constant__ CUtensorMap tmap_do128,
                          const __grid_constant__ CUtensorMap tmap_dq, const __grid_constant__ CUtensorMap tmap_aug_lse,
                          const __grid_constant__ CUtensorMap tmap_aug_delta, const __grid_constant__ CUtensorMap tmap_dk,
                          const __grid_constant__ CUtensorMap tmap_dv, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem + Smem::k;
  uint8_t* sV = smem + Smem::v;
  uint8_t* sQ = smem + Smem::q;
  uint8_t* sdO = smem + Smem::d_o;
  uint8_t* sdS = smem + Smem::ds;
  uint8_t* sAugQ = smem + Smem::aug_q;
  uint8_t* sAugDO = smem + Smem::aug_do;
  uint8_t* sAugOne = smem + Smem::aug_one;
  uint8_t* sAugZero = smem + Smem::aug_zero;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
  uint64_t* bar_kv_full = bars;                       // leader: K and V of BOTH CTAs landed
  uint64_t* bar_q_full = bars + 1;                    // [kQStages] leader: the stage of both CTAs landed
  uint64_t* bar_q_empty = bar_q_full + kQStages;      // [kQStages] both (multicast commit)
  uint64_t* bar_do_full = bar_q_empty + kQStages;     // leader
  uint64_t* bar_do_empty = bar_do_full + 1;           // both (two multicast commits)
  uint64_t* bar_st_full = bar_do_empty + 1;           // both
  uint64_t* bar_p_full = bar_st_full + 1;             // leader: 16 warp arrivals (8 per CTA)
  uint64_t* bar_dp_full = bar_p_full + 1;             // both
  uint64_t* bar_ds_full = bar_dp_full + 1;            // local: own dS^T in smem (8 warp arrivals) -> own dQ^T
  uint64_t* bar_ds_pair = bar_ds_full + 1;            // leader: dS^T of both CTAs in smem (16 warp arrivals) -> dK
  uint64_t* bar_dq_full = bar_ds_pair + 1;            // local
  uint64_t* bar_r1_free = bar_dq_full + 1;            // leader: R1 of both CTAs drained (8 warp arrivals)
  uint64_t* bar_dkv_full = bar_r1_free + 1;           // both: leader A, leader B (multicast) + own B (local)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + Smem::misc);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  const int cid = blockIdx.x >> 1;                    // cluster index
  const int n_pairs_all = (p.n_tiles + 1) >> 1;
  const int pj = p.pair_order[kHeadMajor ? cid % n_pairs_all : cid / p.n_kv_heads];
  const int kvh = kHeadMajor ? cid / n_pairs_all : cid % p.n_kv_heads;
  const int t0 = 2 * pj;
  const int jt = t0 + static_cast<int>(rank);         // this CTA's key tile (may be one past the last tile: all masked)
  int qlast = p.k_qlast[t0];
  if (t0 + 1 < p.n_tiles) qlast = max(qlast, p.k_qlast[t0 + 1]);
  const int q_first = t0;
  const int n_qt = qlast - t0 + 1;
  const int k0 = jt * kT;
  const int n_iter = n_qt * p.group;
  const int kmin_end = jt < p.n_tiles ? p.kmin_end[jt] : 0;

  if (warp == kProducerWarp && lane == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("dta: dynamic shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmap_q64);
    tma_prefetch_desc(&tmap_q128);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_do64);
    tma_prefetch_desc(&tmap_do128);
    tma_prefetch_desc(&tmap_dq);
    tma_prefetch_desc(&tmap_aug_lse);
    tma_prefetch_desc(&tmap_aug_delta);
  }
  if (warp < 4) {
    const int r = threadIdx.x;
    *reinterpret_cast<uint4*>(sAugOne + r * 16) = make_uint4(0x3f803f80u, 0x00003f80u, 0u, 0u);
    *reinterpret_cast<uint4*>(sAugZero + r * 16) = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  if (warp == kMmaWarp) {
    if (lane == 0) {
      mbar_init(bar_kv_full, 1);
      for (int s = 0; s < kQStages; ++s) {
        mbar_init(bar_q_full + s, 1);
        mbar_init(bar_q_empty + s, 1);
      }
      mbar_init(bar_do_full, 1);
      mbar_init(bar_do_empty, 2);
      mbar_init(bar_st_full, 1);
      mbar_init(bar_p_full, 16);
      mbar_init(bar_dp_full, 1);
      mbar_init(bar_ds_full, 8);
      mbar_init(bar_ds_pair, 16);
      mbar_init(bar_dq_full, 1);
      mbar_init(bar_r1_free, 8);
      mbar_init(bar_dkv_full, 3);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_2cta(s_tmem, 512);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                 // the peer's barriers and TMEM exist before anything targets them
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const uint32_t R0 = tmem, R1 = tmem + 128, RdV = tmem + 256, RdK = tmem + 384;

  if (warp == kProducerWarp) {
    // ================================ TMA producer (one per CTA) ==========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (lane == 0) {
      const uint32_t kv_l = mapa_u32(smem_u32(bar_kv_full), 0);
      const uint32_t do_l = mapa_u32(smem_u32(bar_do_full), 0);
      const int qr = static_cast<int>(rank) * 64;     // first own query row / own head-dim column
      if (leader) mbar_arrive_expect_tx(bar_kv_full, 4 * kTile);
      for (int h = 0; h < 2; ++h) {
        tma2_load_3d(sK + h * kHalf, &tmap_k, kv_l, h * 64, kvh, k0);
        tma2_load_3d(sV + h * kHalf, &tmap_v, kv_l, h * 64, kvh, k0);
      }
      for (int it = 0; it < n_iter; ++it) {
        const int head = kvh * p.group + it / n_qt;
        const int q0 = (q_first + it % n_qt) * kT;
        const int st = it % kQStages;
        const uint32_t ph = (it / kQStages) & 1;
        const int aug_row = head * p.n_pad + q0 + qr;
        pwait(bar_q_empty + st, ph ^ 1);
        const uint32_t qf_l = mapa_u32(smem_u32(bar_q_full + st), 0);
        if (leader) mbar_arrive_expect_tx(bar_q_full + st, 2 * (kTile + kAugB));
        uint8_t* sq = sQ + st * kTile;
        tma2_load_3d(sq, &tmap_q64, qf_l, 0, head, q0 + qr);               // own queries, d [0, 64)
        tma2_load_3d(sq + kBlk, &tmap_q64, qf_l, 64, head, q0 + qr);       // own queries, d [64, 128)
        tma2_load_3d(sq + kHalf, &tmap_q128, qf_l, qr, head, q0);          // all queries, own d half
        tma2_load_2d(sAugQ + st * kAugB, &tmap_aug_lse, qf_l, 0, aug_row);
        pwait(bar_do_empty, (it & 1) ^ 1);
        if (leader) mbar_arrive_expect_tx(bar_do_full, 2 * (kTile + kAugB));
        tma2_load_3d(sdO, &tmap_do64, do_l, 0, head, q0 + qr);
        tma2_load_3d(sdO + kBlk, &tmap_do64, do_l, 64, head, q0 + qr);
        tma2_load_3d(sdO + kHalf, &tmap_do128, do_l, qr, head, q0);
        tma2_load_2d(sAugDO, &tmap_aug_delta, do_l, 0, aug_row);
      }
    }
  } else if (warp == kMmaWarp || warp == kMmaWarpB) {
    // ================================ MMA issuers ========================================
    // leader warp A: S^T, dV (pair MMAs).  leader warp B: dP^T, dK (pair MMAs) and its own dQ^T.
    // peer   warp A: nothing.              peer   warp B: its own dQ^T.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    {
      constexpr uint32_t i2_kk = make_idesc_bf16(2 * kT, kT, false, false);   // S^T, dP^T  (M = 256)
      constexpr uint32_t i2_kn = make_idesc_bf16(2 * kT, kT, false, true);    // dV, dK
      constexpr uint32_t i1_nn = make_idesc_bf16(kT, kT, true, true);         // dQ^T (M = 128, per CTA)
      const uint32_t kK = sdesc_lo_kmajor(smem_u32(sK)), kV = sdesc_lo_kmajor(smem_u32(sV));
      const uint32_t kQ = sdesc_lo_kmajor(smem_u32(sQ)), kDO = sdesc_lo_kmajor(smem_u32(sdO));
      const uint32_t kDS = sdesc_lo_kmajor(smem_u32(sdS));
      const uint32_t nK = sdesc_lo_mnmajor(smem_u32(sK), kHalf), nDS = sdesc_lo_mnmajor(smem_u32(sdS), kHalf);
      const uint32_t nQ = sdesc_lo_mnmajor(smem_u32(sQ + kHalf), kHalf), nDO = sdesc_lo_mnmajor(smem_u32(sdO + kHalf), kHalf);
      constexpr uint32_t kStage = kTile >> 4;
      constexpr uint32_t kAugHi = (128u >> 4) | (1u << 14);
      const uint32_t zero_addr = smem_u32(sAugZero);
      auto aug_lo = [&](uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (((zero_addr - addr) >> 4) << 16); };
      const uint32_t aOne = aug_lo(smem_u32(sAugOne)), aDO = aug_lo(smem_u32(sAugDO)), aQ0 = aug_lo(smem_u32(sAugQ));
      constexpr uint32_t kAugStage = (kAugB >> 4) - ((kAugB >> 4) << 16);
#define DTA_KOFF(kk) ((((kk) >> 2) * (kHalf >> 4)) + (((kk) & 3) * 2))   /* K-major A: [128 rows][64] blocks          */
#define DTA_KOFF64(kk) ((((kk) >> 2) * (kBlk >> 4)) + (((kk) & 3) * 2))  /* K-major B half: [64 rows][64] blocks       */
#define DTA_NOFF(kk) ((kk) * (2048 >> 4))                                /* MN-major: 16 rows of the block            */

      if (warp == kMmaWarp) {
        if (leader) {
          auto issue_st = [&](uint32_t st) {       // S^T = K Q^T  (+ 1 (x) lse')
            if (elect_one()) {
              const uint32_t b = kQ + st * kStage;
#pragma unroll
              for (int kk = 0; kk < 8; ++kk) umma2_ss(R0, kK + DTA_KOFF(kk), b + DTA_KOFF64(kk), kSdescHi, i2_kk, kk > 0);
              umma2_ss(R0, aOne, aQ0 + st * kAugStage, kAugHi, i2_kk, 1u);
              umma2_commit(bar_st_full);
            }
            __syncwarp();
          };
          pwait(bar_kv_full, 0);
          pwait(bar_q_full + 0, 0);
          tc_fence_after();
          issue_st(0);
          uint32_t st = 0, stph = 0;
          for (int it = 0; it < n_iter; ++it) {
            const bool has_next = it + 1 < n_iter;
            const uint32_t nst = st + 1 == kQStages ? 0u : st + 1;
            const uint32_t nph = st + 1 == kQStages ? stph ^ 1u : stph;
            const uint32_t not_first = it > 0 ? 1u : 0u;
            pwait_cluster(bar_p_full, it & 1);     // P^T of both CTAs written
            if (lane == 0) DTA_PTRACE(0);
            pwait(bar_do_full, it & 1);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 8; ++kk)
                umma2_ts(RdV, R0 + (kk >> 2) * 64 + (kk & 3) * 8, nDO + DTA_NOFF(kk), i2_kn, (kk > 0) ? 1u : not_first);
              umma2_commit(bar_do_empty);
            }
            __syncwarp();
            if (lane == 0) DTA_PTRACE(13);
            if (has_next) {
              pwait(bar_q_full + nst, nph);
              if (lane == 0) DTA_PTRACE(14);
              tc_fence_after();
              issue_st(nst);
            }
            if (lane == 0) DTA_PTRACE(1);
            st = nst;
            stph = nph;
          }
          if (elect_one()) umma2_commit(bar_dkv_full);
          __syncwarp();
        }
      } else {
        auto issue_dp = [&]() {                    // dP^T = V dO^T  (- 1 (x) delta)
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) umma2_ss(R1, kV + DTA_KOFF(kk), kDO + DTA_KOFF64(kk), kSdescHi, i2_kk, kk > 0);
            umma2_ss(R1, aOne, aDO, kAugHi, i2_kk, 1u);
            umma2_commit(bar_dp_full);
            umma2_commit(bar_do_empty);
          }
          __syncwarp();
        };
        if (leader) {
          pwait(bar_kv_full, 0);
          pwait(bar_do_full, 0);
          tc_fence_after();
          issue_dp();
        }
        uint32_t st = 0;
        for (int it = 0; it < n_iter; ++it) {
          const bool has_next = it + 1 < n_iter;
          const uint32_t not_first = it > 0 ? 1u : 0u;
          // own dQ^T = K^T dS^T as soon as this CTA's dS^T is in shared memory
          pwait(bar_ds_full, it & 1);
          if (lane == 0) DTA_PTRACE(2);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) umma_ss_lo(R1, nK + DTA_NOFF(kk), nDS + DTA_NOFF(kk), i1_nn, kk > 0);
            umma_commit(bar_dq_full);
          }
          __syncwarp();
          if (leader) {
            // dK += dS^T Q for both key tiles once both CTAs' dS^T are in place
            pwait_cluster(bar_ds_pair, it & 1);
            if (lane == 0) DTA_PTRACE(15);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t b = nQ + st * kStage;
#pragma unroll
              for (int kk = 0; kk < 8; ++kk)
                umma2_ss(RdK, kDS + DTA_KOFF(kk), b + DTA_NOFF(kk), kSdescHi, i2_kn, (kk > 0) ? 1u : not_first);
              umma2_commit(bar_q_empty + st);
            }
            __syncwarp();
            if (lane == 0) DTA_PTRACE(3);
            if (has_next) {
              pwait(bar_do_full, (it + 1) & 1);
              pwait_cluster(bar_r1_free, it & 1);   // R1 of both CTAs drained
              if (lane == 0) DTA_PTRACE(4);
              tc_fence_after();
              issue_dp();
              if (lane == 0) DTA_PTRACE(5);
            }
          }
          st = st + 1 == kQStages ? 0u : st + 1;
        }
        if (elect_one()) {
          umma_commit(bar_dkv_full);                    // own cta_group::1 MMAs
          if (leader) umma2_commit(bar_dkv_full);       // the pair MMAs issued by this warp, both CTAs
        }
        __syncwarp();
      }
#undef DTA_KOFF
#undef DTA_KOFF64
#undef DTA_NOFF
    }
  } else if (warp >= 12) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  } else if (warp < 8) {
    // ================================ compute warps (as in attn_bwd.cu) ===================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 160;");
    const int r = ((warp & 3) << 5) | lane;
    const int ch = warp >> 2;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) << 5) << 16;
    const int k_idx = k0 + r;
    const int k_end = k_idx < p.n_nodes ? p.subtree_end[k_idx] : 0;
    const uint32_t p_full_l = mapa_u32(smem_u32(bar_p_full), 0);
    const uint32_t ds_pair_l = mapa_u32(smem_u32(bar_ds_pair), 0);

    for (int it = 0; it < n_iter; ++it) {
      const int qt = q_first + it % n_qt;
      const int qbase = qt * kT + ch * 64;
      const bool partial = !(qt > jt && kmin_end >= (qt + 1) * kT);
      const int vis_lo = k_idx - qbase, vis_hi = k_end - qbase;

      pwait(bar_st_full, it & 1);
      if (threadIdx.x == 0) DTA_PTRACE(6);
      tc_fence_after();
      float pr[64];
      tmem_ld_x32(R0 + lane_addr + ch * 64, reinterpret_cast<uint32_t*>(pr));
      tmem_ld_wait();
      tmem_ld_x32(R0 + lane_addr + ch * 64 + 32, reinterpret_cast<uint32_t*>(pr) + 32);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        if (hf == 1) tmem_ld_wait();
        if (!partial) {
#pragma unroll
          for (int c = hf * 32; c < hf * 32 + 32; ++c) {
            const float x = pr[c] * p.scale_log2;
            pr[c] = exp2_slot_on_fma<kPolyOf8>(c) ? poly_exp2(x) : fast_exp2(x);
          }
        } else {
#pragma unroll
          for (int c = hf * 32; c < hf * 32 + 32; ++c) {
            const float x = pr[c] * p.scale_log2;
            const float v = exp2_slot_on_fma<kPolyOf8>(c) ? poly_exp2(x) : fast_exp2(x);
            pr[c] = (c >= vis_lo && c < vis_hi) ? v : 0.f;
          }
        }
      }
      {
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) pk[c] = pack_bf16x2(pr[2 * c], pr[2 * c + 1]);
        tmem_st_x32(R0 + lane_addr + ch * 64, pk);
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (p.handoff_commit & 1) commit_signal_leader(bar_p_full); else mbar_arrive_cluster(p_full_l);
      }
      if (threadIdx.x == 0) DTA_PTRACE(7);

      pwait(bar_dp_full, it & 1);
      if (threadIdx.x == 0) DTA_PTRACE(8);
      tc_fence_after();
      uint32_t dsp[32];
      {
        float dp0[32], dp1[32];
        tmem_ld_x32(R1 + lane_addr + ch * 64, reinterpret_cast<uint32_t*>(dp0));
        tmem_ld_wait();
        tmem_ld_x32(R1 + lane_addr + ch * 64 + 32, reinterpret_cast<uint32_t*>(dp1));
#pragma unroll
        for (int i = 0; i < 16; ++i) dsp[i] = pack_bf16x2(pr[2 * i] * dp0[2 * i], pr[2 * i + 1] * dp0[2 * i + 1]);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          dsp[16 + i] = pack_bf16x2(pr[32 + 2 * i] * dp1[2 * i], pr[32 + 2 * i + 1] * dp1[2 * i + 1]);
      }
      tc_fence_before();
      // sdS is free: its readers dK(it-1) (pair) and dQ^T(it-1) (own) retired before dP^T(it) was issued -- the leader's
      // warp B issues dK(it-1) before dP^T(it) and waits for r1_free(it-1), which follows dq_full(it-1) in both CTAs
      uint8_t* dst = sdS + ch * kHalf;
#pragma unroll
      for (int chunk = 0; chunk < 8; ++chunk) {
        *reinterpret_cast<uint4*>(dst + sw128_offset(r, chunk)) =
            make_uint4(dsp[4 * chunk], dsp[4 * chunk + 1], dsp[4 * chunk + 2], dsp[4 * chunk + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_ds_full);
        if (p.handoff_commit & 2) commit_signal_leader(bar_ds_pair); else mbar_arrive_cluster(ds_pair_l);
      }
      if (threadIdx.x == 0) DTA_PTRACE(9);
    }

    // ---- epilogue: dV, dK -> bf16 ----------------------------------------------------------
    pwait(bar_dkv_full, 0);
    tc_fence_after();
    if (kTmaEpilogue) {
      const uint32_t st_v = smem_u32(sQ) + ch * kHalf, st_k = smem_u32(sQ) + kTile + ch * kHalf;
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        uint32_t a[32], b[32];
        tmem_ld_x32(RdV + lane_addr + ch * 64 + part * 32, a);
        tmem_ld_x32(RdK + lane_addr + ch * 64 + part * 32, b);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 w;
          const float is = p.inv_scale;
          w.x = pack_bf16x2(__uint_as_float(a[i + 0]) * is, __uint_as_float(a[i + 1]) * is);
          w.y = pack_bf16x2(__uint_as_float(a[i + 2]) * is, __uint_as_float(a[i + 3]) * is);
          w.z = pack_bf16x2(__uint_as_float(a[i + 4]) * is, __uint_as_float(a[i + 5]) * is);
          w.w = pack_bf16x2(__uint_as_float(a[i + 6]) * is, __uint_as_float(a[i + 7]) * is);
          const uint32_t off = sw128_offset(r, part * 4 + (i >> 3));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_v + off), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
          w.x = pack_bf16x2(__uint_as_float(b[i + 0]), __uint_as_float(b[i + 1]));
          w.y = pack_bf16x2(__uint_as_float(b[i + 2]), __uint_as_float(b[i + 3]));
          w.z = pack_bf16x2(__uint_as_float(b[i + 4]), __uint_as_float(b[i + 5]));
          w.w = pack_bf16x2(__uint_as_float(b[i + 6]), __uint_as_float(b[i + 7]));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_k + off), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 256);                  // the eight compute warps of this CTA
      if (threadIdx.x == 0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&tmap_dv),
                       "r"(smem_u32(sQ) + h * kHalf), "r"(h * 64), "r"(kvh), "r"(k0)
                       : "memory");
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&tmap_dk),
                       "r"(smem_u32(sQ) + kTile + h * kHalf), "r"(h * 64), "r"(kvh), "r"(k0)
                       : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      }
    } else {
      __nv_bfloat16* dv_row = p.dv + static_cast<int64_t>(k_idx) * p.v_row_stride + static_cast<int64_t>(kvh) * p.v_head_stride + ch * 64;
      __nv_bfloat16* dk_row = p.dk + static_cast<int64_t>(k_idx) * p.k_row_stride + static_cast<int64_t>(kvh) * p.k_head_stride + ch * 64;
  #pragma unroll
      for (int part = 0; part < 2; ++part) {
        uint32_t a[32], b[32];
        tmem_ld_x32(RdV + lane_addr + ch * 64 + part * 32, a);
        tmem_ld_x32(RdK + lane_addr + ch * 64 + part * 32, b);
        tmem_ld_wait();
        if (k_idx < p.n_nodes) {
  #pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 w;
            const float is = p.inv_scale;
            w.x = pack_bf16x2(__uint_as_float(a[i + 0]) * is, __uint_as_float(a[i + 1]) * is);
            w.y = pack_bf16x2(__uint_as_float(a[i + 2]) * is, __uint_as_float(a[i + 3]) * is);
            w.z = pack_bf16x2(__uint_as_float(a[i + 4]) * is, __uint_as_float(a[i + 5]) * is);
            w.w = pack_bf16x2(__uint_as_float(a[i + 6]) * is, __uint_as_float(a[i + 7]) * is);
            *reinterpret_cast<uint4*>(dv_row + part * 32 + i) = w;
            w.x = pack_bf16x2(__uint_as_float(b[i + 0]), __uint_as_float(b[i + 1]));
            w.y = pack_bf16x2(__uint_as_float(b[i + 2]), __uint_as_float(b[i + 3]));
            w.z = pack_bf16x2(__uint_as_float(b[i + 4]), __uint_as_float(b[i + 5]));
            w.w = pack_bf16x2(__uint_as_float(b[i + 6]), __uint_as_float(b[i + 7]));
            *reinterpret_cast<uint4*>(dk_row + part * 32 + i) = w;
          }
        }
      }
    }
  } else {
    // ================================ dQ drain warps (as in attn_bwd.cu) ===================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 136;");
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) << 5) << 16;
    const uint32_t r1_free_l = mapa_u32(smem_u32(bar_r1_free), 0);
    for (int it = 0; it < n_iter; ++it) {
      const int head = kvh * p.group + it / n_qt;
      const int q0 = (q_first + it % n_qt) * kT;
      pwait(bar_dq_full, it & 1);
      if (threadIdx.x == 256) DTA_PTRACE(10);
      tc_fence_after();
      float acc[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_x32(R1 + lane_addr + c * 32, reinterpret_cast<uint32_t*>(acc) + c * 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == p.drain_lane) {
        if (p.handoff_commit & 4) commit_signal_leader(bar_r1_free); else mbar_arrive_cluster(r1_free_l);
      }
      if (threadIdx.x == 256) DTA_PTRACE(11);
      float* stage = reinterpret_cast<float*>(smem + Smem::dq_stage) + (warp & 3) * (kDqBufs * kDqRows * 32);
      const int c0 = head * kT + ((warp & 3) << 5);
#pragma unroll
      for (int ck = 0; ck < kT / kDqRows; ++ck) {
        float* buf = stage + ((it * (kT / kDqRows) + ck) % kDqBufs) * (kDqRows * 32);
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kDqBufs - 1) : "memory");
        __syncwarp();
#pragma unroll
        for (int r = 0; r < kDqRows; ++r) buf[r * 32 + lane] = acc[ck * kDqRows + r];
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          asm volatile(
              "cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                  reinterpret_cast<uint64_t>(&tmap_dq)),
              "r"(c0), "r"(q0 + ck * kDqRows), "r"(smem_u32(buf))
              : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (threadIdx.x == 256) DTA_PTRACE(12);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                 // no arrival or MMA of the peer may target a CTA that has left
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem, 512);
  }
}

// pair_order[i] = j of the i-th pair (key tiles 2j, 2j+1), in the order their first tile appears in k_order
// (heaviest first).  One CTA; first_pos is scratch [n_pairs].
__global__ void __launch_bounds__(1024) pair_order_kernel(const int32_t* __restrict__ k_order, int n_tiles,
                                                          int32_t* __restrict__ first_pos, int32_t* __restrict__ pair_order) {
  __shared__ int s_warp[32];
  __shared__ int s_bcast;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_pairs = (n_tiles + 1) / 2;
  for (int i = tid; i < n_pairs; i += blockDim.x) first_pos[i] = INT_MAX;
  __syncthreads();
  for (int pos = tid; pos < n_tiles; pos += blockDim.x) atomicMin(first_pos + (k_order[pos] >> 1), pos);
  __syncthreads();
  int base = 0;
  for (int start = 0; start < n_tiles; start += blockDim.x) {
    const int pos = start + tid;
    int pj = 0;
    bool emit = false;
    if (pos < n_tiles) {
      pj = k_order[pos] >> 1;
      emit = __ldcg(first_pos + pj) == pos;      // written by atomics above: read it where they live (L2)
    }
    // exclusive block scan of `emit`
    const int v = emit ? 1 : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const int w = s_warp[lane];
      int winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      s_warp[lane] = winc - w;
      if (lane == 31) s_bcast = winc;
    }
    __syncthreads();
    if (emit) pair_order[base + s_warp[warp] + inc - v] = pj;
    base += s_bcast;
    __syncthreads();
  }
}

// Main-kernel launch of dta_tree_attn_bwd in pair mode.  `scratch`: >= 2 * ceil(n_tiles / 2) int32 (the work-list
// region of the workspace, unused in this mode).  aug_lse / aug_delta / dq_accum as prepared by the caller.
int launch(const void* q, const void* k, const void* v, const void* dout, const void* aug_lse, const void* aug_delta,
           float* dq_accum, void* dk, void* dv, const dta_attn_meta_t* meta, const dta_attn_shape_t* sh, int32_t* scratch,
           cudaEvent_t ev_start, cudaEvent_t ev_stop, long long* trace, cudaStream_t stream) {
  const int n_tiles = sh->n_pad / DTA_TILE;
  const int n_pairs = (n_tiles + 1) / 2;
  int32_t* first_pos = scratch;
  int32_t* pair_order = scratch + n_pairs;
  pair_order_kernel<<<1, 1024, 0, stream>>>(meta->k_order, n_tiles, first_pos, pair_order);
  DTA_CHECK_CUDA(cudaGetLastError());

  CUtensorMap tq64, tq128, tk, tv, tdo64, tdo128, tdq, taug_l, taug_d;
  int rc;
  if ((rc = make_tmap_bf16_rows(&tq64, q, sh->n_nodes, sh->n_q_heads, kT, sh->q_row_stride, sh->q_head_stride, 64))) return rc;
  if ((rc = make_tmap_bf16_rows(&tq128, q, sh->n_nodes, sh->n_q_heads, kT, sh->q_row_stride, sh->q_head_stride, kT))) return rc;
  if ((rc = make_tmap_bf16_rows(&tk, k, sh->n_nodes, sh->n_kv_heads, kT, sh->k_row_stride, sh->k_head_stride, kT))) return rc;
  if ((rc = make_tmap_bf16_rows(&tv, v, sh->n_nodes, sh->n_kv_heads, kT, sh->v_row_stride, sh->v_head_stride, kT))) return rc;
  if ((rc = make_tmap_bf16_rows(&tdo64, dout, sh->n_nodes, sh->n_q_heads, kT, sh->o_row_stride, sh->o_head_stride, 64))) return rc;
  if ((rc = make_tmap_bf16_rows(&tdo128, dout, sh->n_nodes, sh->n_q_heads, kT, sh->o_row_stride, sh->o_head_stride, kT))) return rc;
  if ((rc = make_tmap_f32_2d(&tdq, dq_accum, static_cast<int64_t>(sh->n_q_heads) * kT, sh->n_pad, 32, kDqRows))) return rc;
  // the 16-byte lse / delta operand rows [Hq * n_pad] viewed as 4 floats each: a [64 rows] box lands as the
  // un-swizzled K-major chunk (8-row x 16-byte core matrices) the broadcast k-step reads
  const int64_t aug_rows = static_cast<int64_t>(sh->n_q_heads) * sh->n_pad;
  if ((rc = make_tmap_f32_2d(&taug_l, aug_lse, 4, aug_rows, 4, 64))) return rc;
  if ((rc = make_tmap_f32_2d(&taug_d, aug_delta, 4, aug_rows, 4, 64))) return rc;

  Params p;
  p.dq_accum = dq_accum;
  p.dk = static_cast<__nv_bfloat16*>(dk);
  p.dv = static_cast<__nv_bfloat16*>(dv);
  p.subtree_end = meta->subtree_end;
  p.kmin_end = meta->kmin_end;
  p.k_qlast = meta->k_qlast;
  p.pair_order = pair_order;
  p.n_nodes = sh->n_nodes;
  p.n_pad = sh->n_pad;
  p.n_tiles = n_tiles;
  p.n_q_heads = sh->n_q_heads;
  p.n_kv_heads = sh->n_kv_heads;
  p.group = sh->n_q_heads / sh->n_kv_heads;
  p.k_row_stride = sh->k_row_stride;
  p.k_head_stride = sh->k_head_stride;
  p.v_row_stride = sh->v_row_stride;
  p.v_head_stride = sh->v_head_stride;
  p.scale_log2 = sh->scale * 1.4426950408889634f;
  p.inv_scale = 1.0f / sh->scale;
  p.trace = trace;
  const char* dl = getenv("DTA_PAIR_DRAIN_LANE");
  p.drain_lane = dl ? (atoi(dl) & 31) : 0;
  const char* ho = getenv("DTA_PAIR_HANDOFF");
  // "commit" = all three hand-offs (P^T, dS^T, R1 free); "commit_p" / "commit_ds" / "commit_r1" = one of them (A/B)
  p.handoff_commit = !ho ? kDefaultHandoff : strcmp(ho, "commit") == 0 ? 7 : strcmp(ho, "commit_p") == 0 ? 1 : strcmp(ho, "commit_ds") == 0 ? 2
                     : strcmp(ho, "commit_r1") == 0 ? 4 : strcmp(ho, "commit_p_ds") == 0 ? 3 : strcmp(ho, "arrive") == 0 ? 0 : kDefaultHandoff;

  CUtensorMap tdk, tdv;     // kTmaEpilogue stores (dk: the k strides, dv: the v strides)
  if ((rc = make_tmap_bf16_rows(&tdk, dk, sh->n_nodes, sh->n_kv_heads, kT, sh->k_row_stride, sh->k_head_stride, kT))) return rc;
  if ((rc = make_tmap_bf16_rows(&tdv, dv, sh->n_nodes, sh->n_kv_heads, kT, sh->v_row_stride, sh->v_head_stride, kT))) return rc;
  using Kernel = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                          const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                          const CUtensorMap, const Params);
  // [TMA epilogue][poly 0 / 2 / 4, then head-major with poly 0]: the rejected variants stay reachable for A/B runs
  static const Kernel kernels[8] = {tree_attn_bwd_pair_kernel<0, false, false>, tree_attn_bwd_pair_kernel<2, false, false>,
                                    tree_attn_bwd_pair_kernel<4, false, false>, tree_attn_bwd_pair_kernel<0, true, false>,
                                    tree_attn_bwd_pair_kernel<0, false, true>,  tree_attn_bwd_pair_kernel<2, false, true>,
                                    tree_attn_bwd_pair_kernel<4, false, true>,  tree_attn_bwd_pair_kernel<0, true, true>};
  static bool configured = false;
  if (!configured) {
    for (Kernel kf : kernels) DTA_CHECK_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::total));
    configured = true;
  }
  const char* pe = getenv("DTA_BWD_POLY");                 // same tuning knob as the single-CTA kernel; default 0
  const int poly = pe ? atoi(pe) : 0;
  const char* hm = getenv("DTA_BWD_HEAD_MAJOR");           // A/B knob (round 2, in-step): see kDefaultPairHeadMajor
  const bool head_major = hm ? atoi(hm) != 0 : kDefaultPairHeadMajor;
  const char* te = getenv("DTA_PAIR_TMA_EPILOGUE");        // A/B knob
  const bool tma_epilogue = te ? atoi(te) != 0 : kDefaultPairTmaEpilogue;
  const Kernel kernel = kernels[(tma_epilogue ? 4 : 0) + (head_major ? 3 : (poly <= 0 ? 0 : poly <= 2 ? 1 : 2))];
  if (ev_start) DTA_CHECK_CUDA(cudaEventRecord(ev_start, stream));
  kernel<<<2 * n_pairs * sh->n_kv_heads, kThreads, Smem::total, stream>>>(tq64, tq128, tk, tv, tdo64, tdo128, tdq, taug_l, taug_d, tdk, tdv, p);
  DTA_CHECK_CUDA(cudaGetLastError());
  if (ev_stop) DTA_CHECK_CUDA(cudaEventRecord(ev_stop, stream));
  return DTA_OK;
}

}  // namespace bwd_pair
}  // namespace dta
