// attn_bwd.cu — tree-masked flash attention backward for sm_100a (tcgen05 + TMEM + TMA).
//
// Replaces the autograd backward through HF attention that the reference runs once per
// popped block, including its O(prefix) grad-KV accumulation
// (tree_training_engine.py:404-465), by ONE pass per layer over the linearised trie.
//
// Work decomposition
//   CTA        = one 128-key tile x one KV head x one chunk of the tile's LIVE query tiles.  The live
//                query tiles of a key tile form one contiguous range [tile, k_qlast[tile]] because every
//                key's visible queries are the interval [k, subtree_end[k]).  The CTA loops over the
//                query heads of the GQA group and over its chunk; dK/dV accumulate in TMEM across the
//                whole loop.  Normally the chunk is the whole range and dK/dV are written once, in bf16.
//                A key tile of a shared trunk is visible to EVERY later query tile: when its range is
//                longer than an even share of the launch (bwd_work_kernel), it is split over several
//                CTAs, which add their fp32 partial dK/dV into an accumulator with 16-byte vector reds
//                (converted by the epilogue kernel) -- otherwise one CTA would be the long pole of bushy
//                tries.  dQ tiles leave through shared memory + TMA tensor reduce-add into an fp32
//                accumulator and are converted to bf16 by the same epilogue kernel.
//   per step   = 5 GEMMs, all 128x128x128 bf16 -> fp32 on tcgen05:
//                  S^T  = K  Q^T     (SS, both K-major)            -> TMEM R0
//                  dP^T = V  dO^T    (SS, both K-major)            -> TMEM R1
//                  dV  += P^T dO     (A = P^T bf16 in TMEM, B = dO MN-major)
//                  dQ^T = K^T dS^T   (A = K read MN-major, B = dS^T read MN-major) -> TMEM R1
//                  dK  += dS^T Q     (A = dS^T smem K-major, B = Q MN-major)
//                computed transposed so that TMEM lanes = key rows: the ancestor mask is
//                then a per-thread column RANGE test [k, subtree_end[k]).
//   warp roles = warps 0-7 compute (lane = key row, two column halves) | warps 8-11 dQ drain
//                (lane = head-dim index) | warp 14 TMA producer | warps 15 and 13 MMA issuers (15 owns TMEM).
//                The issuer sits in the highest warp slot: the SM arbiter favours high warp ids
//                and the issuer is the one warp whose latency is on every critical path.
//   TMEM       = R0 S^T/P^T [0,128) (P^T bf16: queries 0-63 in cols [0,32), 64-127 in cols [64,96))
//                  R1 dP^T/dQ [128,256)  dV [256,384)  dK [384,512)
#include "dta_common.cuh"
#include "dta_ptx.cuh"
#include "../../include/dta_b200.h"
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace dta {

static constexpr int kT = 128;                       // tile edge (queries, keys) and head dim
static constexpr int kHalf = kT * 128;               // [128 rows][64 bf16] swizzled half tile, bytes
static constexpr int kTile = 2 * kHalf;              // 32 KB
static constexpr int kBwdThreads = 512;
static constexpr int kProducerWarp = 14, kMmaWarp = 15, kMmaWarpB = 13;
static constexpr int kQStages = 2;                   // Q ring depth
static constexpr int kDqBufs = 3;                    // staging buffers per drain warp (reduces in flight)
static constexpr int kDqRows = 16;                   // query rows per dQ reduce box (2 KB per warp and buffer)
static constexpr int kAug = kT * 16;                 // one un-swizzled [128][8 bf16] K-chunk, bytes
static constexpr float kLog2e = 1.4426950408889634f;
static constexpr bool kDefaultDqTma = true;
static constexpr int kDefaultL2Ahead = 0;
static constexpr int kDefaultRedGap = 0;             // see the drain warps
static constexpr int kDefaultBwdPoly = 0;            // of every 8 exponentials, on the FMA pipes

struct BwdSmem {
  static constexpr int k = 0;
  static constexpr int v = k + kTile;
  static constexpr int q = v + kTile;                 // kQStages stages
  static constexpr int d_o = q + kQStages * kTile;    // ONE stage (see the ring note in the kernel)
  static constexpr int ds = d_o + kTile;
  // un-swizzled K-major [128 rows][8 bf16] chunks (2 KB each) for the broadcast k-step (see kernel header)
  static constexpr int aug_q = ds + kTile;            // kQStages chunks: per-query (lse hi, mid, lo, 0...)
  static constexpr int aug_do = aug_q + kQStages * kAug;   // 1 chunk: per-query -(delta hi, mid, lo, 0...)
  static constexpr int aug_one = aug_do + kAug;       // 1 chunk: (1, 1, 1, 0...) in every row
  static constexpr int aug_zero = aug_one + kAug;     // 1 chunk of zeros = k columns 8..15 of all of the above
  // dQ staging for the bulk tensor reduce: per drain warp kDqBufs [kDqRows queries][32 floats] buffers
  static constexpr int dq_stage = aug_zero + kAug;
  static constexpr int bars = dq_stage + 4 * kDqBufs * kDqRows * 128;
  static constexpr int n_bars = 1 + 2 * kQStages + 2 + 8;
  static constexpr int misc = bars + n_bars * 8;
  static constexpr int total = misc + 16;
};
static_assert(BwdSmem::total <= 232448, "backward kernel shared memory exceeds 227 KB");

struct BwdParams {
  const __nv_bfloat16* aug_lse;    // [Hq, n_pad, 8]  3-term bf16 split of (log2(scale) - lse2) / scale_log2, then zeros
  const __nv_bfloat16* aug_delta;  // [Hq, n_pad, 8]  3-term bf16 split of -delta, then zeros   (both: delta pre-pass)
  float* dq_accum;         // [n_pad, Hq, 128] fp32
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
  const int32_t* subtree_end;
  const int32_t* kmin_end;
  const int32_t* k_order;
  const int32_t* k_qlast;
  const int4* work;        // [work_cap] {key tile, first query tile, query tiles, split?}  (bwd_work_kernel)
  const int32_t* n_work;   // [1] number of valid work items
  float* dk_acc;           // [n_pad, Hkv, 128] fp32: dK of key tiles whose query range is split over several CTAs
  float* dv_acc;           // same for dV
  int n_nodes, n_pad, n_q_heads, n_kv_heads, group;
  int64_t k_row_stride, k_head_stride, v_row_stride, v_head_stride;
  float scale, scale_log2, inv_scale;
  int l2_ahead;            // L2 prefetch distance of the Q / dO tiles, in steps (0 = off)
  int red_gap;             // dependent integer ops between two dQ reds of a drain warp (issue pacing)
  int debug_flags;         // timing ablations / trace modes (see the uses); 0 in production
  long long* trace;        // debug timeline (NULL in production): [step][16] clock64 stamps of CTA 0
};

#define DTA_TRACE(ev)                                                                         \
  do {                                                                                        \
    if (p.trace != nullptr && blockIdx.x == 0 && it < 48)                                     \
      p.trace[it * 16 + (ev)] = clock64();                                                    \
  } while (0)
// trace mode 2 (debug flag 32): per compute warp, slot w = P^T pass done, slot 8 + w = dS^T pass done
__device__ __forceinline__ void red_add_f32(float* addr, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}
__device__ __forceinline__ void red_add_v4_f32(float* addr, float a, float b, float c, float d) {   // 16-byte aligned
  asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// kPolyOf8: how many of every 8 exponentials of the P^T pass run on the FMA pipes (poly_exp2)
// kDqTma: dQ tiles leave through shared memory + cp.reduce.async.bulk.tensor (TMA) instead of per-lane reds
template <int kPolyOf8, bool kDqTma, bool kWorkList>
__global__ void __launch_bounds__(kBwdThreads, 1)
tree_attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                     const __grid_constant__ CUtensorMap tmap_dq, const __grid_constant__ CUtensorMap tmap_dk,
                     const __grid_constant__ CUtensorMap tmap_dv, const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem + BwdSmem::k;
  uint8_t* sV = smem + BwdSmem::v;
  uint8_t* sQ = smem + BwdSmem::q;
  uint8_t* sdO = smem + BwdSmem::d_o;
  uint8_t* sdS = smem + BwdSmem::ds;
  uint8_t* sAugQ = smem + BwdSmem::aug_q;
  uint8_t* sAugDO = smem + BwdSmem::aug_do;
  uint8_t* sAugOne = smem + BwdSmem::aug_one;
  uint8_t* sAugZero = smem + BwdSmem::aug_zero;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BwdSmem::bars);
  uint64_t* bar_kv_full = bars;
  uint64_t* bar_q_full = bars + 1;                    // [kQStages]
  uint64_t* bar_q_empty = bar_q_full + kQStages;      // [kQStages]
  uint64_t* bar_do_full = bar_q_empty + kQStages;     // single dO stage, one phase per step
  uint64_t* bar_do_empty = bar_do_full + 1;
  uint64_t* bar_st_full = bar_do_empty + 1;           // S^T ready
  uint64_t* bar_p_full = bar_st_full + 1;             // P^T written (8 warp arrivals)
  uint64_t* bar_dp_full = bar_p_full + 1;             // dP^T ready
  uint64_t* bar_ds_full = bar_dp_full + 1;            // dS^T in smem (8 warp arrivals)
  uint64_t* bar_ds_empty = bar_ds_full + 1;           // (unused: implied by dp_full, see the dS^T pass)
  uint64_t* bar_dq_full = bar_ds_empty + 1;           // dQ tile ready in R1
  uint64_t* bar_r1_free = bar_dq_full + 1;            // dQ tile copied out of R1 (4 warp arrivals)
  uint64_t* bar_dkv_full = bar_r1_free + 1;           // everything done
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + BwdSmem::misc);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // kWorkList: the CTA's job comes from the device-built work list (chunks of split key tiles), else it is the whole
  // live range of key tile k_order[blockIdx / Hkv].  The second form needs no list kernel, no spare CTAs and no
  // dependent loads before the K/V fetch; earlier forms of the list decode also cost 2-4 % of the kernel through
  // code generation alone (profiles/r01/exp14_bwd_work_decode.log).  The caller picks it through
  // dta_attn_shape_t::bwd_whole_tiles when the launch is provably balanced without splitting.
  int jt, q_first, n_qt;
  bool split = false;
  if constexpr (kWorkList) {
    const int widx = blockIdx.x / p.n_kv_heads;
    if (widx >= *p.n_work) return;                        // the grid is sized for the worst-case number of chunks
    const int32_t* item = reinterpret_cast<const int32_t*>(p.work) + 4 * widx;
    jt = item[0];                                         // key tile
    q_first = item[1];                                    // first query tile of this CTA's chunk
    n_qt = item[2];                                       // query tiles in the chunk
    split = item[3] != 0;                                 // other CTAs hold the rest of this key tile's range
  } else {
    jt = p.k_order[blockIdx.x / p.n_kv_heads];
    q_first = jt;
    n_qt = p.k_qlast[jt] - jt + 1;
  }
  const int kvh = blockIdx.x % p.n_kv_heads;
  const int k0 = jt * kT;
  const int n_iter = n_qt * p.group;
  const int kmin_end = p.kmin_end[jt];

  if (warp == kProducerWarp && lane == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("dta: dynamic shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_do);
    if (kDqTma) tma_prefetch_desc(&tmap_dq);
  }
  if (warp < 4) {
    // constant operands of the broadcast k-step: ones (three of them, matching the 3-term splits) and zeros
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
      mbar_init(bar_do_empty, 2);          // dP^T (issuer B) and dV (issuer A) both retired
      mbar_init(bar_st_full, 1);
      mbar_init(bar_p_full, 8);
      mbar_init(bar_dp_full, 1);
      mbar_init(bar_ds_full, 8);
      mbar_init(bar_ds_empty, 1);
      mbar_init(bar_dq_full, 1);
      mbar_init(bar_r1_free, 4);
      mbar_init(bar_dkv_full, 2);          // both issuing warps
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(s_tmem, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const uint32_t R0 = tmem, R1 = tmem + 128, RdV = tmem + 256, RdK = tmem + 384;

  if (warp == kProducerWarp) {
    // ================================ TMA producer =======================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_kv_full, 2 * kTile);
      for (int h = 0; h < 2; ++h) {
        tma_load_3d(sK + h * kHalf, &tmap_k, bar_kv_full, h * 64, kvh, k0);
        tma_load_3d(sV + h * kHalf, &tmap_v, bar_kv_full, h * 64, kvh, k0);
      }
      // Ring shape.  Q(it) is read by the first MMA of a step (S^T) and by the last (dK): two stages.  dO(it) is
      // read by dP^T(it) and dV(it) only -- both early -- and dO(it+1) is not needed before dP^T(it+1), the LAST
      // MMA of step it: one stage, reloaded as soon as dV(it) retires, has a whole step of slack.
      for (int it = 0; it < n_iter; ++it) {
        const int head = kvh * p.group + it / n_qt;
        const int q0 = (q_first + it % n_qt) * kT;
        const int st = it % kQStages;
        const uint32_t ph = (it / kQStages) & 1;
        mbar_wait(bar_q_empty + st, ph ^ 1);
        const int64_t aug_off = (static_cast<int64_t>(head) * p.n_pad + q0) * 8;
        mbar_arrive_expect_tx(bar_q_full + st, kTile + kAug);
        for (int h = 0; h < 2; ++h) tma_load_3d(sQ + st * kTile + h * kHalf, &tmap_q, bar_q_full + st, h * 64, head, q0);
        bulk_load_1d(sAugQ + st * kAug, p.aug_lse + aug_off, kAug, bar_q_full + st);
        mbar_wait(bar_do_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(bar_do_full, kTile + kAug);
        for (int h = 0; h < 2; ++h) tma_load_3d(sdO + h * kHalf, &tmap_do, bar_do_full, h * 64, head, q0);
        bulk_load_1d(sAugDO, p.aug_delta + aug_off, kAug, bar_do_full);
        // warm L2 for the tiles `l2_ahead` steps ahead: the ring cannot request them before a stage retires, so
        // what the eventual load costs must be an L2 hit (~1.0-1.4k cycles), not a DRAM round trip (~1.9-2.6k
        // with every SM streaming: tools/tma_bench.cu)
        if (p.l2_ahead > 0 && it + p.l2_ahead < n_iter) {
          const int it2 = it + p.l2_ahead;
          const int head2 = kvh * p.group + it2 / n_qt;
          const int q2 = (q_first + it2 % n_qt) * kT;
          for (int h = 0; h < 2; ++h) {
            tma_prefetch_l2_3d(&tmap_q, h * 64, head2, q2);
            tma_prefetch_l2_3d(&tmap_do, h * 64, head2, q2);
          }
        }
      }
    }
  } else if (warp == kMmaWarp || warp == kMmaWarpB) {
    // ================================ MMA issuers ========================================
    // TWO issuing warps, one per dependency loop of a step, so that a wait in one loop never holds back the
    // MMAs of the other (one in-order issuer spent ~1.3k of ~4.4k cycles per step blocked on the wrong loop's
    // barrier while the tensor pipe idled):
    //   warp A (kMmaWarp) : wait P^T(it)  -> dV(it)  -> wait Q(it+1) -> S^T(it+1)                  [TMEM R0 loop]
    //   warp B (kMmaWarpB): wait dS^T(it) -> dQ^T(it), dK(it) -> wait dO(it+1), R1 drained -> dP^T(it+1)  [R1 loop]
    // The tensor pipe executes tcgen05.mma in arrival order whichever warp sent them; the two streams touch
    // disjoint accumulators (A: R0, dV; B: R1, dK), and every hazard between them goes through a compute-warp
    // barrier (S^T -> P^T -> dS^T).  dO(it) is read by dP^T(it) (B) and dV(it) (A): bar_do_empty counts both.
    // In each warp the WHOLE warp runs the loop (waits are warp-uniform) and only the tcgen05 instructions are
    // predicated on one elected lane, so descriptor arithmetic stays on the uniform datapath.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    {
      constexpr uint32_t idesc_kk = make_idesc_bf16(kT, kT, false, false);   // S^T, dP^T
      constexpr uint32_t idesc_kn = make_idesc_bf16(kT, kT, false, true);    // dV, dK (B MN-major)
      constexpr uint32_t idesc_nn = make_idesc_bf16(kT, kT, true, true);     // dQ^T (A and B MN-major)
      // descriptor low words; k-step kk adds a compile-time constant (see dta_ptx.cuh)
      const uint32_t kK = sdesc_lo_kmajor(smem_u32(sK)), kV = sdesc_lo_kmajor(smem_u32(sV));
      const uint32_t kQ = sdesc_lo_kmajor(smem_u32(sQ)), kDO = sdesc_lo_kmajor(smem_u32(sdO));
      const uint32_t kDS = sdesc_lo_kmajor(smem_u32(sdS));
      const uint32_t nK = sdesc_lo_mnmajor(smem_u32(sK), kHalf), nQ = sdesc_lo_mnmajor(smem_u32(sQ), kHalf);
      const uint32_t nDO = sdesc_lo_mnmajor(smem_u32(sdO), kHalf), nDS = sdesc_lo_mnmajor(smem_u32(sdS), kHalf);
      constexpr uint32_t kStage = kTile >> 4;                       // next Q stage
      // broadcast k-step: un-swizzled K-major, SBO = 128 (8-row groups), LBO = distance to the shared zero chunk
      constexpr uint32_t kAugHi = (128u >> 4) | (1u << 14);
      const uint32_t zero_addr = smem_u32(sAugZero);
      auto aug_lo = [&](uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (((zero_addr - addr) >> 4) << 16); };
      const uint32_t aOne = aug_lo(smem_u32(sAugOne)), aDO = aug_lo(smem_u32(sAugDO)), aQ0 = aug_lo(smem_u32(sAugQ));
      constexpr uint32_t kAugStage = (kAug >> 4) - ((kAug >> 4) << 16);   // next chunk: address up, LBO down
#define DTA_KOFF(kk) ((((kk) >> 2) * (kHalf >> 4)) + (((kk) & 3) * 2))   /* K-major: 16 elements along the row */
#define DTA_NOFF(kk) ((kk) * (2048 >> 4))                                /* MN-major: 16 rows of the tile     */

      if (warp == kMmaWarp) {
        // ------------------------------ warp A: dV, S^T ------------------------------------
        auto issue_st = [&](uint32_t st) {       // S^T = K Q^T
          if (elect_one()) {
            const uint32_t b = kQ + st * kStage;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) umma_ss_lo(R0, kK + DTA_KOFF(kk), b + DTA_KOFF(kk), idesc_kk, kk > 0);
            umma_ss_lohi(R0, aOne, aQ0 + st * kAugStage, kAugHi, idesc_kk, 1u);   // S^T += 1 (x) lse'
            umma_commit(bar_st_full);
          }
          __syncwarp();
        };
        mbar_wait(bar_kv_full, 0);
        mbar_wait(bar_q_full + 0, 0);
        tc_fence_after();
        issue_st(0);
        uint32_t st = 0, stph = 0;               // Q stage of step it and its phase parity
        for (int it = 0; it < n_iter; ++it) {
          const bool has_next = it + 1 < n_iter;
          const uint32_t nst = st + 1 == kQStages ? 0u : st + 1;
          const uint32_t nph = st + 1 == kQStages ? stph ^ 1u : stph;
          const uint32_t not_first = it > 0 ? 1u : 0u;
          // dV += P^T dO
          mbar_wait(bar_p_full, it & 1);
          if (lane == 0) DTA_TRACE(0);
          mbar_wait(bar_do_full, it & 1);          // dO(it) landed (warp B checked it for dP^T(it); this warp has not)
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_ts_lo(RdV, R0 + (kk >> 2) * 64 + (kk & 3) * 8, nDO + DTA_NOFF(kk), idesc_kn, (kk > 0) ? 1u : not_first);
            umma_commit(bar_do_empty);             // this warp's half of "dO(it) has no reader left"
          }
          __syncwarp();
          if (lane == 0) DTA_TRACE(13);
          // S^T of the next step reuses R0 right behind dV (tensor pipe executes one warp's MMAs in issue order)
          if (has_next) {
            mbar_wait(bar_q_full + nst, nph);
            if (lane == 0) DTA_TRACE(14);
            tc_fence_after();
            issue_st(nst);
          }
          if (lane == 0) DTA_TRACE(1);
          st = nst;
          stph = nph;
        }
      } else {
        // ------------------------------ warp B: dP^T, dQ^T, dK -------------------------------
        auto issue_dp = [&]() {                  // dP^T = V dO^T
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) umma_ss_lo(R1, kV + DTA_KOFF(kk), kDO + DTA_KOFF(kk), idesc_kk, kk > 0);
            umma_ss_lohi(R1, aOne, aDO, kAugHi, idesc_kk, 1u);                                   // dP^T -= 1 (x) delta
            umma_commit(bar_dp_full);
            umma_commit(bar_do_empty);           // this warp's half of "dO(it) has no reader left"
          }
          __syncwarp();
        };
        mbar_wait(bar_kv_full, 0);
        mbar_wait(bar_do_full, 0);
        tc_fence_after();
        issue_dp();
        uint32_t st = 0;
        for (int it = 0; it < n_iter; ++it) {
          const bool has_next = it + 1 < n_iter;
          const uint32_t not_first = it > 0 ? 1u : 0u;
          // dQ^T = K^T dS^T (first, so the drain overlaps dK), dK += dS^T Q
          mbar_wait(bar_ds_full, it & 1);
          if (lane == 0) DTA_TRACE(2);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) umma_ss_lo(R1, nK + DTA_NOFF(kk), nDS + DTA_NOFF(kk), idesc_nn, kk > 0);
            umma_commit(bar_dq_full);
            const uint32_t b = nQ + st * kStage;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_ss_lo(RdK, kDS + DTA_KOFF(kk), b + DTA_NOFF(kk), idesc_kn, (kk > 0) ? 1u : not_first);
            umma_commit(bar_q_empty + st);         // S^T(it) (warp A) retired long ago: dK(it) depends on it through P^T / dS^T
          }
          __syncwarp();
          if (lane == 0) DTA_TRACE(3);
          // dP^T of the next step needs R1 back from the drain warps
          if (has_next) {
            mbar_wait(bar_do_full, (it + 1) & 1);
            mbar_wait(bar_r1_free, it & 1);
            if (lane == 0) DTA_TRACE(4);
            tc_fence_after();
            issue_dp();
            if (lane == 0) DTA_TRACE(5);
          }
          st = st + 1 == kQStages ? 0u : st + 1;
        }
      }
      if (elect_one()) umma_commit(bar_dkv_full);     // one arrival per issuing warp
      __syncwarp();
    }
  } else if (warp >= 12) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  } else if (warp < 8) {
    // ================================ compute warps ======================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 160;");
    const int r = ((warp & 3) << 5) | lane;          // key row inside the tile == TMEM lane
    const int ch = warp >> 2;                        // column (query) half
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) << 5) << 16;
    const int k_idx = k0 + r;
    const int k_end = k_idx < p.n_nodes ? p.subtree_end[k_idx] : 0;

    for (int it = 0; it < n_iter; ++it) {
      const int qt = q_first + it % n_qt;
      const int qbase = qt * kT + ch * 64;
      const bool partial = !(qt > jt && kmin_end >= (qt + 1) * kT);
      const int vis_lo = k_idx - qbase, vis_hi = k_end - qbase;      // visible columns of this row: [lo, hi)

      // ---- P^T = exp2(S^T * scale_log2 - lse * log2e), masked -------------------------
      mbar_wait(bar_st_full, it & 1);
      if (threadIdx.x == 0) DTA_TRACE(6);
      tc_fence_after();
      float pr[64];
      tmem_ld_x32(R0 + lane_addr + ch * 64, reinterpret_cast<uint32_t*>(pr));
      tmem_ld_wait();
      tmem_ld_x32(R0 + lane_addr + ch * 64 + 32, reinterpret_cast<uint32_t*>(pr) + 32);   // in flight under the first half
      // The MMA already added the per-query exponent offset (broadcast k-step), so
      // pr = scale * P^T = 2^(S'^T * scale_log2) with no per-column operand: the 64 lse values every thread used
      // to pull through LDS (4 wavefronts per LDS.128, broadcast or not) were ~1k data-pipe cycles per step.
      // The 1/scale is undone once, in the dV epilogue, so that dS^T = pr * dP'^T needs no further multiply.
      // Two code copies: tiles the ancestor mask leaves untouched pay nothing for it.
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
        tmem_st_x32(R0 + lane_addr + ch * 64, pk);   // inside this warp-half's own S^T columns
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p_full);          // one arrival per warp
      if (threadIdx.x == 0) DTA_TRACE(7);

      // ---- dS^T = (scale P^T) o dP'^T -> smem (bf16, 128B-swizzled) --------------------
      mbar_wait(bar_dp_full, it & 1);
      if (threadIdx.x == 0) DTA_TRACE(8);
      tc_fence_after();
      uint32_t dsp[32];
#pragma unroll
      {                                       // dP'^T = dP^T - delta already (broadcast k-step); two 32-column halves
        float dp0[32], dp1[32];
        tmem_ld_x32(R1 + lane_addr + ch * 64, reinterpret_cast<uint32_t*>(dp0));
        tmem_ld_wait();
        tmem_ld_x32(R1 + lane_addr + ch * 64 + 32, reinterpret_cast<uint32_t*>(dp1));   // in flight under the math below
#pragma unroll
        for (int i = 0; i < 16; ++i) dsp[i] = pack_bf16x2(pr[2 * i] * dp0[2 * i], pr[2 * i + 1] * dp0[2 * i + 1]);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          dsp[16 + i] = pack_bf16x2(pr[32 + 2 * i] * dp1[2 * i], pr[32 + 2 * i + 1] * dp1[2 * i + 1]);
      }
      tc_fence_before();                   // our R1 reads are done before the arrive below
      // sdS is free: dK(it-1), its last reader, was issued by the same warp before dP^T(it), and dp_full(it) -- which
      // this warp has already observed -- is a tcgen05.commit covering every earlier MMA of that warp.
      uint8_t* dst = sdS + ch * kHalf;
#pragma unroll
      for (int chunk = 0; chunk < 8; ++chunk) {
        *reinterpret_cast<uint4*>(dst + sw128_offset(r, chunk)) =
            make_uint4(dsp[4 * chunk], dsp[4 * chunk + 1], dsp[4 * chunk + 2], dsp[4 * chunk + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ds_full);
      if (threadIdx.x == 0) DTA_TRACE(9);
    }

    // ---- epilogue: dV, dK -> bf16 ----------------------------------------------------------
    mbar_wait(bar_dkv_full, 0);
    tc_fence_after();
    if (split) {
      // partial sums of a split key tile: fp32 vector reds (16 bytes per lane; a lane owns one key row)
      float* dv_acc = p.dv_acc + (static_cast<int64_t>(k_idx) * p.n_kv_heads + kvh) * kT + ch * 64;
      float* dk_acc = p.dk_acc + (static_cast<int64_t>(k_idx) * p.n_kv_heads + kvh) * kT + ch * 64;
      // (rare path: a rolled loop keeps it out of the instruction cache's way -- the kernel is ~110 KB of SASS)
      const float is = p.inv_scale;            // P^T was carried pre-multiplied by scale
#pragma unroll 1
      for (int c = 0; c < 64; c += 8) {
        uint32_t a[8], b[8];
        tmem_ld_x8(RdV + lane_addr + ch * 64 + c, a);
        tmem_ld_x8(RdK + lane_addr + ch * 64 + c, b);
        tmem_ld_wait();
        if (k_idx < p.n_nodes) {
          red_add_v4_f32(dv_acc + c, __uint_as_float(a[0]) * is, __uint_as_float(a[1]) * is, __uint_as_float(a[2]) * is,
                         __uint_as_float(a[3]) * is);
          red_add_v4_f32(dv_acc + c + 4, __uint_as_float(a[4]) * is, __uint_as_float(a[5]) * is,
                         __uint_as_float(a[6]) * is, __uint_as_float(a[7]) * is);
          red_add_v4_f32(dk_acc + c, __uint_as_float(b[0]), __uint_as_float(b[1]), __uint_as_float(b[2]),
                         __uint_as_float(b[3]));
          red_add_v4_f32(dk_acc + c + 4, __uint_as_float(b[4]), __uint_as_float(b[5]), __uint_as_float(b[6]),
                         __uint_as_float(b[7]));
        }
      }
    } else {
      // dV and dK leave through shared memory and a TMA store each: a thread owns one key row, so storing its 128
      // bytes straight to global memory makes every store instruction touch 32 rows (measured on the forward:
      // 7,100 cycles of LSU time per tile pair, csrc/attn_fwd_persist.cu).  Staging = the two Q stages, dead now that
      // every MMA has retired: [128 keys][64 d] halves with the 128-byte swizzle the tensor map expects; rows past
      // n_nodes are clipped by the store.
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
          const float is = p.inv_scale;      // P^T was carried pre-multiplied by scale
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
      named_bar_sync(1, 256);                  // the eight compute warps
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
    }
  } else {
    // ================================ dQ drain warps =====================================
    // R1 holds dQ^T: TMEM lane = head-dim index d, column = query.  A warp-level red therefore adds 32
    // consecutive floats of ONE dq row: 128 contiguous bytes = 4 full sectors (the untransposed layout
    // touched 32 half-used sectors per instruction and doubled the L2 atomic traffic).
    asm volatile("setmaxnreg.inc.sync.aligned.u32 136;");
    const int d = ((warp & 3) << 5) | lane;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) << 5) << 16;
    for (int it = 0; it < n_iter; ++it) {
      const int head = kvh * p.group + it / n_qt;
      const int q0 = (q_first + it % n_qt) * kT;
      mbar_wait(bar_dq_full, it & 1);
      if (threadIdx.x == 256) DTA_TRACE(10);
      tc_fence_after();
      float acc[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_x32(R1 + lane_addr + c * 32, reinterpret_cast<uint32_t*>(acc) + c * 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_r1_free);
      if (threadIdx.x == 256) DTA_TRACE(11);
      // NOTE: the running 64-bit pointer below makes the 128 reds one dependent IMAD.WIDE chain, which PACES
      // them.  Issuing them with independent addresses (measured) floods the LSU queue, starves the compute
      // warps' LDS/STS/TMEM traffic and slows the whole kernel by 50 %; a TMA bulk reduce from shared memory
      // is the proper fix but needs 16-64 KB of staging this kernel does not have.
      float* dst = p.dq_accum + (static_cast<int64_t>(q0) * p.n_q_heads + head) * kT + d;
      const int64_t row = static_cast<int64_t>(p.n_q_heads) * kT;
      const int n_valid = min(kT, p.n_nodes - q0);
      if (kDqTma) {
        // Each drain warp owns 32 head-dim lanes = one 128-byte column block of the dq rows.  It stages kDqRows
        // query rows at a time ([rows][32 floats], conflict-free STS.32) in one of its two private buffers and
        // hands them to the TMA unit as ONE tensor reduce-add; no other warp is involved, so the only
        // synchronisation is __syncwarp + the bulk-group read fence of the issuing lane.  512 reds per step and
        // SM (one 128-byte red per ~5 LSU cycles, all step long) become 512 single-wavefront stores + 32 TMA ops.
        float* stage = reinterpret_cast<float*>(smem + BwdSmem::dq_stage) + (warp & 3) * (kDqBufs * kDqRows * 32);
        const int c0 = head * kT + ((warp & 3) << 5);
#pragma unroll
        for (int ck = 0; ck < kT / kDqRows; ++ck) {
          float* buf = stage + ((it * (kT / kDqRows) + ck) % kDqBufs) * (kDqRows * 32);
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kDqBufs - 1) : "memory");   // oldest buffer is free
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
      } else if (p.debug_flags & 1) {
        // ablation: no reds
      } else if (n_valid == kT && p.red_gap > 0) {
        // rate-matched issue: the SM retires one 128-byte red per ~5 cycles (measured, tools/red_bench.cu), i.e.
        // one per ~20 cycles per drain warp.  Issuing faster only fills the per-scheduler MIO queue, and every
        // LDS / STS / MUFU of the compute warps on that scheduler then waits behind the queued reds.
        uint32_t pace = lane;
#pragma unroll
        for (int c = 0; c < kT; ++c) {
          red_add_f32(dst + c * row, acc[c]);
          for (int g = 0; g < p.red_gap; ++g) asm volatile("mad.lo.u32 %0, %0, %0, %0;" : "+r"(pace));
        }
      } else if (n_valid == kT) {
#pragma unroll
        for (int c = 0; c < kT; ++c) red_add_f32(dst + c * row, acc[c]);
      } else {
#pragma unroll
        for (int c = 0; c < kT; ++c)
          if (c < n_valid) red_add_f32(dst + c * row, acc[c]);
      }
      if (threadIdx.x == 256) DTA_TRACE(12);
    }
    if (kDqTma && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------
// Work list of the main kernel (one CTA, once per call; n_tiles <= a few thousand).  Key tile t has
// n_qt(t) = k_qlast[t] - t + 1 live query tiles and is one job per KV head.  With share = the (tile, head) steps an
// SM gets when the launch is perfectly balanced: if some tile is longer than 3/4 of a share, every tile longer than
//   L = max(kMinSplitTiles, ceil(share / 2)),   share = ceil(n_kv_heads * sum_t n_qt(t) / n_sm),
// is cut into ceil(n_qt / L) chunks of equal size (ordering: see below); shorter tiles cannot unbalance an LPT
// schedule by more than that and stay whole -- splitting is not free (chunks re-read K/V, leave through fp32 reds,
// and thin out the lock-step that keeps Q / dO in L2: 1-3 % on tries that do not need it).
// sum(chunks) <= n_tiles + 2 n_sm / n_kv_heads, which sizes the grid on the host without a device read-back.
// ------------------------------------------------------------------------------------
static constexpr int kMinSplitTiles = 16;

__global__ void __launch_bounds__(1024) bwd_work_kernel(const int32_t* __restrict__ k_order,
                                                        const int32_t* __restrict__ k_qlast, int n_tiles, int n_sm,
                                                        int n_kv_heads, int enable_split, int4* __restrict__ work, int32_t* __restrict__ n_work,
                                                        int32_t* __restrict__ k_split, int32_t* __restrict__ split_list) {
  __shared__ int s_warp[32];
  __shared__ int s_bcast;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto block_scan = [&](int v, int& total) {          // exclusive scan over the 1024 threads
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane], winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      s_warp[lane] = winc - w;
      if (lane == 31) s_bcast = winc;
    }
    __syncthreads();
    const int excl = s_warp[warp] + inc - v;
    total = s_bcast;
    __syncthreads();
    return excl;
  };
  __shared__ int s_max;
  if (tid == 0) s_max = 0;
  __syncthreads();
  int local = 0, local_max = 0;
  for (int i = tid; i < n_tiles; i += blockDim.x) {
    const int n = k_qlast[i] - i + 1;
    local += n;
    local_max = max(local_max, n);
  }
  atomicMax(&s_max, local_max);
  int total;
  (void)block_scan(local, total);                      // (its barriers also publish s_max)
  const int share = static_cast<int>((static_cast<int64_t>(total) * n_kv_heads + n_sm - 1) / n_sm);
  // split only when some tile exceeds 3/4 of a share (LPT absorbs shorter ones: splitting fan(4096, 16) at 1/2 costs
  // 3 % schedule efficiency in tools/bwd_schedule_sim.py); chunks are then at most half a share long
  const bool need_split = enable_split && s_max > max(kMinSplitTiles, (3 * share + 3) / 4);
  const int split_len = need_split ? max(kMinSplitTiles, (share + 1) / 2) : (1 << 30);
  const int max_parts = (s_max + split_len - 1) / split_len;
  // Order: chunk 0 of every split tile (k_order sequence), then chunk 1 of every split tile, ..., then the unsplit
  // tiles.  Heavy items first (LPT), and the CTAs resident at the same time walk the SAME query range, which is
  // what keeps their Q / dO loads and dQ reductions meeting in L2 (interleaving a tile's chunks costs 7 %).
  int base = 0;
  for (int c = 0; c <= max_parts; ++c) {               // c == max_parts: the unsplit tiles
    if (c < max_parts && max_parts == 1) continue;     // nothing is split
    for (int start = 0; start < n_tiles; start += blockDim.x) {
      const int pos = start + tid;
      int jt = 0, n_qt = 0, parts = 0;
      if (pos < n_tiles) {
        jt = k_order[pos];
        n_qt = k_qlast[jt] - jt + 1;
        parts = (n_qt + split_len - 1) / split_len;
      }
      const bool emit = pos < n_tiles && (c == max_parts ? parts == 1 : (parts > 1 && c < parts));
      int chunk_total;
      const int off = block_scan(emit ? 1 : 0, chunk_total);
      if (emit) {
        if (c == max_parts) {
          k_split[jt] = 0;
          work[base + off] = make_int4(jt, jt, n_qt, 0);
        } else {
          k_split[jt] = 1;
          if (c == 0) split_list[base + off] = jt;     // pass 0 emits exactly one item per split tile
          const int lo = static_cast<int>(static_cast<int64_t>(c) * n_qt / parts);
          const int hi = static_cast<int>(static_cast<int64_t>(c + 1) * n_qt / parts);
          work[base + off] = make_int4(jt, jt + lo, hi - lo, 1);
        }
      }
      base += chunk_total;
      if (c == 0 && start + static_cast<int>(blockDim.x) >= n_tiles && tid == 0) n_work[1] = base;   // number of split tiles
    }
  }
  if (tid == 0) {
    n_work[0] = base;
    if (max_parts == 1) n_work[1] = 0;
  }
}

// rows of split key tiles in the fp32 dK / dV accumulators ([n_pad, Hkv, 128] each); split_list / n_split: the compact
// list of those tiles (empty for every launch that needs no splitting, so the loop costs one load)
__device__ __forceinline__ void zero_split_rows(const int32_t* __restrict__ split_list, const int32_t* __restrict__ n_split,
                                                int n_kv_heads, float* __restrict__ dk_acc, float* __restrict__ dv_acc) {
  const int n = *n_split;
  const int64_t per_tile = static_cast<int64_t>(kT) * n_kv_heads * (kT / 4);       // float4 per tile and buffer
  const int64_t tid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t n_thr = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int s = 0; s < n; ++s) {
    const int t = split_list[s];
    float4* a = reinterpret_cast<float4*>(dk_acc) + t * per_tile;
    float4* b = reinterpret_cast<float4*>(dv_acc) + t * per_tile;
    for (int64_t i = tid; i < per_tile; i += n_thr) {
      a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------
// Pre-pass (one warp per (q, h) row; HBM-bound): delta[h, q] = sum_d dO[q,h,d] * O[q,h,d] and the exponent
// offset L = (log2(scale) - lse2[h, q]) / scale_log2, each written as a 3-term bf16 split into a 16-byte
// row {hi, mid, lo, 0, 0, 0, 0, 0} -- the B operand rows of the main kernel's broadcast k-step, which
// computes S'^T = S^T + L and dP'^T = dP^T - delta on the tensor core (products with 1.0 are exact and
// the fp32 accumulator keeps all 24 bits of the three terms).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 split3_bf16(float x) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(h);
  const __nv_bfloat16 m = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(m);
  const __nv_bfloat16 l = __float2bfloat16_rn(r2);
  const uint32_t w0 = static_cast<uint32_t>(__bfloat16_as_ushort(h)) | (static_cast<uint32_t>(__bfloat16_as_ushort(m)) << 16);
  return make_uint4(w0, static_cast<uint32_t>(__bfloat16_as_ushort(l)), 0u, 0u);
}

__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ out,
                                                         const __nv_bfloat16* __restrict__ dout, int n_nodes,
                                                         int n_pad, int n_heads, int64_t row_stride,
                                                         int64_t head_stride, const float* __restrict__ lse2,
                                                         float lshift, float inv_scale_log2,
                                                         uint4* __restrict__ aug_lse, uint4* __restrict__ aug_delta,
                                                         const int32_t* __restrict__ split_list,
                                                         const int32_t* __restrict__ n_split, int n_kv_heads,
                                                         float* __restrict__ dk_acc, float* __restrict__ dv_acc,
                                                         float4* __restrict__ dq_zero, int64_t n_dq_zero) {
  zero_split_rows(split_list, n_split, n_kv_heads, dk_acc, dv_acc);
  // the fp32 dQ accumulator is zeroed here, interleaved with this kernel's reads (a read + write stream uses HBM better
  // than a read-only pass followed by a memset: -40 us per layer at 67k tokens)
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n_dq_zero;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    __stcs(dq_zero + i, make_float4(0.f, 0.f, 0.f, 0.f));
  // half-warp per (q, h) row, 16 bytes per lane and tensor, four row pairs loaded before the first is reduced
  const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
  const int64_t n_rows = static_cast<int64_t>(n_pad) * n_heads;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t w0 = ((blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5) * 8; w0 < n_rows; w0 += n_warps * 8) {
    uint4 a[4], b[4];
    float l2[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t w = w0 + 2 * u + half;
      const int q = static_cast<int>(w / n_heads), h = static_cast<int>(w % n_heads);
      a[u] = b[u] = make_uint4(0u, 0u, 0u, 0u);
      l2[u] = 0.f;
      if (w < n_rows && q < n_nodes) {
        const int64_t off = q * row_stride + h * head_stride + hl * 8;
        a[u] = __ldg(reinterpret_cast<const uint4*>(out + off));
        b[u] = __ldg(reinterpret_cast<const uint4*>(dout + off));
        if (hl == 0) l2[u] = __ldg(lse2 + static_cast<int64_t>(h) * n_pad + q);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t w = w0 + 2 * u + half;
      const int q = static_cast<int>(w / n_heads), h = static_cast<int>(w % n_heads);
      const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a[u]);
      const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&b[u]);
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 x = __bfloat1622float2(a2[i]), y = __bfloat1622float2(b2[i]);
        acc = fmaf(x.x, y.x, acc);
        acc = fmaf(x.y, y.y, acc);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (hl == 0 && w < n_rows) {
        const int64_t o = static_cast<int64_t>(h) * n_pad + q;
        aug_delta[o] = split3_bf16(-acc);
        aug_lse[o] = split3_bf16(q < n_nodes ? (lshift - l2[u]) * inv_scale_log2 : 0.f);
      }
    }
  }
}

// dq (bf16, q strides) <- dq_accum (fp32 [n_pad, Hq, 128])
// + for key tiles whose query range was split: dk / dv (bf16, k / v strides) <- the fp32 partial-sum accumulators
__global__ void __launch_bounds__(256) attn_dq_convert_kernel(const float* __restrict__ acc, int n_nodes, int n_dq_rows, int n_heads,
                                                              int64_t row_stride, int64_t head_stride,
                                                              __nv_bfloat16* __restrict__ dq,
                                                              const int32_t* __restrict__ split_list,
                                                              const int32_t* __restrict__ n_split,
                                                              int n_kv_heads, const float* __restrict__ dk_acc,
                                                              const float* __restrict__ dv_acc, int64_t k_row_stride,
                                                              int64_t k_head_stride, int64_t v_row_stride,
                                                              int64_t v_head_stride, __nv_bfloat16* __restrict__ dk,
                                                              __nv_bfloat16* __restrict__ dv) {
  {
    const int64_t per_tile = static_cast<int64_t>(kT) * n_kv_heads * (kT / 8);     // 8-element vectors per tile
    const int64_t tid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int64_t n_thr = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int n_sp = *n_split;
    for (int sidx = 0; sidx < n_sp; ++sidx) {
      const int t = split_list[sidx];
      for (int64_t i = tid; i < per_tile; i += n_thr) {
        const int c = static_cast<int>(i % (kT / 8)) * 8;
        const int64_t rh = i / (kT / 8);
        const int h = static_cast<int>(rh % n_kv_heads);
        const int64_t k = static_cast<int64_t>(t) * kT + rh / n_kv_heads;
        if (k >= n_nodes) continue;
        const int64_t src = (k * n_kv_heads + h) * kT + c;
#pragma unroll
        for (int which = 0; which < 2; ++which) {
          const float* a_ = (which ? dv_acc : dk_acc) + src;
          const float4 a = __ldcs(reinterpret_cast<const float4*>(a_));
          const float4 b = __ldcs(reinterpret_cast<const float4*>(a_ + 4));
          uint4 w;
          w.x = pack_bf16x2(a.x, a.y);
          w.y = pack_bf16x2(a.z, a.w);
          w.z = pack_bf16x2(b.x, b.y);
          w.w = pack_bf16x2(b.z, b.w);
          __nv_bfloat16* dst = which ? dv + k * v_row_stride + h * v_head_stride + c : dk + k * k_row_stride + h * k_head_stride + c;
          *reinterpret_cast<uint4*>(dst) = w;
        }
      }
    }
  }
  // n_dq_rows = n_nodes, or 0 when the caller keeps dQ in fp32 (only the split tiles' dK / dV are converted then)
  const int64_t n_vec = static_cast<int64_t>(n_dq_rows) * n_heads * (kT / 8);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n_vec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % (kT / 8)) * 8;
    const int64_t rh = i / (kT / 8);
    const int h = static_cast<int>(rh % n_heads);
    const int64_t q = rh / n_heads;
    const float4 a = __ldcs(reinterpret_cast<const float4*>(acc + rh * kT + c));
    const float4 b = __ldcs(reinterpret_cast<const float4*>(acc + rh * kT + c + 4));
    uint4 w;
    w.x = pack_bf16x2(a.x, a.y);
    w.y = pack_bf16x2(a.z, a.w);
    w.z = pack_bf16x2(b.x, b.y);
    w.w = pack_bf16x2(b.z, b.w);
    *reinterpret_cast<uint4*>(dq + q * row_stride + h * head_stride + c) = w;
  }
}

}  // namespace dta

using namespace dta;

namespace dta {
namespace bwd_pair {   // attn_bwd_pair.cu: experimental CTA-pair main kernel (DTA_BWD_PAIR=1)
int launch(const void* q, const void* k, const void* v, const void* dout, const void* aug_lse, const void* aug_delta,
           float* dq_accum, void* dk, void* dv, const dta_attn_meta_t* meta, const dta_attn_shape_t* sh, int32_t* scratch,
           cudaEvent_t ev_start, cudaEvent_t ev_stop, long long* trace /* [48][2][16] or NULL */, cudaStream_t stream);
}
}  // namespace dta

static long long* g_bwd_trace = nullptr;

// Workspace: [dq_accum fp32 | lse operand rows | delta operand rows | work items | {n_work, n_split} | k_split flags |
//             split tile list | dk_acc | dv_acc]
struct BwdWorkspace {
  size_t dq, aug_lse, aug_delta, work, n_work, k_split, split_list, dk_acc, dv_acc, total;
  int work_cap;
};
static BwdWorkspace bwd_workspace_layout(const dta_attn_shape_t* sh, int n_sm) {
  auto up = [](size_t x) { return (x + 255) & ~static_cast<size_t>(255); };
  BwdWorkspace w;
  const int n_tiles = sh->n_pad / DTA_TILE;
  w.work_cap = n_tiles + (2 * n_sm + sh->n_kv_heads - 1) / sh->n_kv_heads + 1;   // see bwd_work_kernel
  const size_t aug = static_cast<size_t>(sh->n_pad) * sh->n_q_heads * 16;     // one 16-byte operand row per (head, query)
  const size_t dkv = static_cast<size_t>(sh->n_pad) * sh->n_kv_heads * kT * sizeof(float);
  size_t o = 0;
  w.dq = o;        o = up(o + static_cast<size_t>(sh->n_pad) * sh->n_q_heads * kT * sizeof(float));
  w.aug_lse = o;   o = up(o + aug);
  w.aug_delta = o; o = up(o + aug);
  w.work = o;      o = up(o + static_cast<size_t>(w.work_cap) * sizeof(int4));
  w.n_work = o;    o = up(o + 16);
  w.k_split = o;   o = up(o + static_cast<size_t>(n_tiles) * sizeof(int32_t));
  w.split_list = o; o = up(o + static_cast<size_t>(n_tiles) * sizeof(int32_t));
  w.dk_acc = o;    o = up(o + dkv);
  w.dv_acc = o;    o = up(o + dkv);
  w.total = o;
  return w;
}

extern "C" {

/* debug / test only (not in the public header): the work list dta_tree_attn_bwd builds for a launch.
 * work: int32[4 * cap] {key tile, first query tile, query tiles, split}, n_work: int32[2], k_split / split_list: int32[n_tiles];
 * returns the capacity the caller must provide (cap) when work == NULL. */
int dta_debug_bwd_work_list(const int32_t* k_order, const int32_t* k_qlast, int n_tiles, int n_sm, int n_kv_heads,
                            int32_t* work, int32_t* n_work /* [2]: items, split tiles */, int32_t* k_split,
                            int32_t* split_list /* [n_tiles] */, dta_stream_t stream) {
  const int cap = n_tiles + (2 * n_sm + n_kv_heads - 1) / n_kv_heads + 1;
  if (work == nullptr) return cap;
  bwd_work_kernel<<<1, 1024, 0, stream>>>(k_order, k_qlast, n_tiles, n_sm, n_kv_heads, 1, reinterpret_cast<int4*>(work),
                                          n_work, k_split, split_list);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

/* debug only (not in the public header): device buffer of 48*16 int64 receiving the step timeline of CTA 0
 * (48*32 int64 in pair mode: both CTAs of cluster 0) */
int dta_debug_set_bwd_trace(long long* buf) { g_bwd_trace = buf; return 0; }

size_t dta_tree_attn_bwd_workspace_bytes(const dta_attn_shape_t* sh) {
  if (!sh) return 0;
  return bwd_workspace_layout(sh, sm_count()).total;
}

}  // extern "C" (debug entry points, workspace size)

// dq_f32 != NULL: that buffer ([n_pad, Hq, 128] fp32, 16-byte aligned) IS the dQ accumulator and is left as the result; the
// bf16 dq tensor is not written (dq may be NULL) -- the consumer (dta_qknorm_rope_bwd_f32dq) reads the fp32 values.
static int tree_attn_bwd_impl(const void* q, const void* k, const void* v, const void* out, const void* dout,
                              const float* lse, void* dq, float* dq_f32, void* dk, void* dv, const dta_attn_meta_t* meta,
                              const dta_attn_shape_t* sh, void* workspace, dta_stream_t stream) {
  DTA_REQUIRE(q && k && v && out && dout && lse && (dq || dq_f32) && dk && dv && meta && sh && workspace, "NULL pointer argument");
  DTA_REQUIRE(sh->head_dim == kT, "head_dim must be %d (got %d)", kT, sh->head_dim);
  DTA_REQUIRE(sh->n_nodes >= 1 && sh->n_pad % DTA_TILE == 0 && sh->n_pad >= sh->n_nodes &&
                  sh->n_pad - sh->n_nodes < DTA_TILE,
              "n_pad must be n_nodes rounded up to %d", DTA_TILE);
  DTA_REQUIRE(sh->n_kv_heads >= 1 && sh->n_q_heads % sh->n_kv_heads == 0, "n_q_heads must be a multiple of n_kv_heads");
  DTA_REQUIRE(meta->subtree_end && meta->kmin_end && meta->k_order && meta->k_qlast, "NULL metadata pointer");
  DTA_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  DTA_REQUIRE((reinterpret_cast<uintptr_t>(lse) & 15) == 0, "lse must be 16-byte aligned");
  const uintptr_t align_or = reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dq_f32) | reinterpret_cast<uintptr_t>(dk) |
                             reinterpret_cast<uintptr_t>(dv) | reinterpret_cast<uintptr_t>(out) |
                             reinterpret_cast<uintptr_t>(dout);
  DTA_REQUIRE((align_or & 15) == 0, "tensors must be 16-byte aligned");
  DTA_REQUIRE(sh->q_row_stride % 8 == 0 && sh->q_head_stride % 8 == 0 && sh->k_row_stride % 8 == 0 &&
                  sh->k_head_stride % 8 == 0 && sh->v_row_stride % 8 == 0 && sh->v_head_stride % 8 == 0 &&
                  sh->o_row_stride % 8 == 0 && sh->o_head_stride % 8 == 0,
              "strides must be multiples of 8 elements");

  const BwdWorkspace wl = bwd_workspace_layout(sh, sm_count());
  uint8_t* ws8 = static_cast<uint8_t*>(workspace);
  float* dq_accum = dq_f32 ? dq_f32 : reinterpret_cast<float*>(ws8 + wl.dq);
  const size_t dq_bytes = static_cast<size_t>(sh->n_pad) * sh->n_q_heads * kT * sizeof(float);
  uint4* aug_lse = reinterpret_cast<uint4*>(ws8 + wl.aug_lse);
  uint4* aug_delta = reinterpret_cast<uint4*>(ws8 + wl.aug_delta);
  int4* work = reinterpret_cast<int4*>(ws8 + wl.work);
  int32_t* n_work = reinterpret_cast<int32_t*>(ws8 + wl.n_work);
  int32_t* k_split = reinterpret_cast<int32_t*>(ws8 + wl.k_split);
  int32_t* split_list = reinterpret_cast<int32_t*>(ws8 + wl.split_list);
  float* dk_acc = reinterpret_cast<float*>(ws8 + wl.dk_acc);
  float* dv_acc = reinterpret_cast<float*>(ws8 + wl.dv_acc);
  const int n_tiles_ = sh->n_pad / DTA_TILE;
  // (dq_accum is zeroed by attn_delta_kernel below)
  const int enable_split = getenv("DTA_BWD_SPLIT") ? atoi(getenv("DTA_BWD_SPLIT")) : 1;   // 0: never split (A/B knob)
  // CTA-pair main kernel (attn_bwd_pair.cu): the pair kernel has no split chunks and pairing wastes 22-33 % of the
  // steps on bushy tries, hence only for launches with bwd_whole_tiles
  // DEFAULT ("auto"): the CTA-pair kernel for launches the caller marked balanced, the single-CTA work-list kernel
  // otherwise; DTA_BWD_PAIR=0 / 1 force one of them (A/B, tests).  In the bench step the pair kernel runs the
  // tau2 launch at 915-925 TFLOP/s against 895-905 (same box: profiles/r02/visit5_*.log, visit1_*.log).
  const char* pair_env = getenv("DTA_BWD_PAIR");
  const bool use_pair = (!pair_env || strcmp(pair_env, "auto") == 0) ? sh->bwd_whole_tiles != 0 : atoi(pair_env) != 0;
  if (sh->bwd_whole_tiles != 0 || !enable_split || use_pair) {
    DTA_CHECK_CUDA(cudaMemsetAsync(n_work, 0, 16, stream));                      // n_split = 0: nothing to zero or convert
  } else {
    bwd_work_kernel<<<1, 1024, 0, stream>>>(meta->k_order, meta->k_qlast, n_tiles_, sm_count(), sh->n_kv_heads, 1, work, n_work,
                                            k_split, split_list);
  }
  DTA_CHECK_CUDA(cudaGetLastError());

  const int64_t n_rows = static_cast<int64_t>(sh->n_pad) * sh->n_q_heads;
  {
    int64_t grid = ((n_rows + 7) / 8 * 32 + 255) / 256;
    const int64_t cap = 16 * static_cast<int64_t>(sm_count());
    if (grid > cap) grid = cap;
    attn_delta_kernel<<<static_cast<int>(grid), 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(dout), sh->n_nodes, sh->n_pad,
        sh->n_q_heads, sh->o_row_stride, sh->o_head_stride, lse, log2f(sh->scale), 1.0f / (sh->scale * kLog2e),
        aug_lse, aug_delta, split_list, n_work + 1, sh->n_kv_heads, dk_acc, dv_acc, reinterpret_cast<float4*>(dq_accum),
        static_cast<int64_t>(dq_bytes / sizeof(float4)));
    DTA_CHECK_CUDA(cudaGetLastError());
  }

  CUtensorMap tq, tk, tv, tdo;
  int rc;
  if ((rc = make_tmap_bf16_rows(&tq, q, sh->n_nodes, sh->n_q_heads, kT, sh->q_row_stride, sh->q_head_stride, kT))) return rc;
  if ((rc = make_tmap_bf16_rows(&tk, k, sh->n_nodes, sh->n_kv_heads, kT, sh->k_row_stride, sh->k_head_stride, kT))) return rc;
  if ((rc = make_tmap_bf16_rows(&tv, v, sh->n_nodes, sh->n_kv_heads, kT, sh->v_row_stride, sh->v_head_stride, kT))) return rc;
  if ((rc = make_tmap_bf16_rows(&tdo, dout, sh->n_nodes, sh->n_q_heads, kT, sh->o_row_stride, sh->o_head_stride, kT))) return rc;
  CUtensorMap tdq;
  if ((rc = make_tmap_f32_2d(&tdq, dq_accum, static_cast<int64_t>(sh->n_q_heads) * kT, sh->n_pad, 32, kDqRows))) return rc;
  CUtensorMap tdk, tdv;     // dK / dV stores of the single-CTA kernel's whole tiles (dk: the k strides, dv: the v strides)
  if ((rc = make_tmap_bf16_rows(&tdk, dk, sh->n_nodes, sh->n_kv_heads, kT, sh->k_row_stride, sh->k_head_stride, kT))) return rc;
  if ((rc = make_tmap_bf16_rows(&tdv, dv, sh->n_nodes, sh->n_kv_heads, kT, sh->v_row_stride, sh->v_head_stride, kT))) return rc;

  BwdParams p;
  p.aug_lse = reinterpret_cast<const __nv_bfloat16*>(aug_lse);
  p.aug_delta = reinterpret_cast<const __nv_bfloat16*>(aug_delta);
  p.dq_accum = dq_accum;
  p.dk = static_cast<__nv_bfloat16*>(dk);
  p.dv = static_cast<__nv_bfloat16*>(dv);
  p.subtree_end = meta->subtree_end;
  p.kmin_end = meta->kmin_end;
  p.k_order = meta->k_order;
  p.k_qlast = meta->k_qlast;
  p.work = work;
  p.n_work = n_work;
  p.dk_acc = dk_acc;
  p.dv_acc = dv_acc;
  p.n_nodes = sh->n_nodes;
  p.n_pad = sh->n_pad;
  p.n_q_heads = sh->n_q_heads;
  p.n_kv_heads = sh->n_kv_heads;
  p.group = sh->n_q_heads / sh->n_kv_heads;
  p.k_row_stride = sh->k_row_stride;
  p.k_head_stride = sh->k_head_stride;
  p.v_row_stride = sh->v_row_stride;
  p.v_head_stride = sh->v_head_stride;
  p.scale = sh->scale;
  p.scale_log2 = sh->scale * kLog2e;
  p.inv_scale = 1.0f / sh->scale;
  p.l2_ahead = getenv("DTA_BWD_L2PF") ? atoi(getenv("DTA_BWD_L2PF")) : kDefaultL2Ahead;
  p.red_gap = getenv("DTA_BWD_REDGAP") ? atoi(getenv("DTA_BWD_REDGAP")) : kDefaultRedGap;
  p.debug_flags = getenv("DTA_DEBUG_FLAGS") ? atoi(getenv("DTA_DEBUG_FLAGS")) : 0;
  p.trace = g_bwd_trace;

  using Kernel = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                          const CUtensorMap, const CUtensorMap, const BwdParams);
  // tuning knobs: DTA_BWD_POLY = share (of 8) of the exponentials computed on the FMA pipes; DTA_BWD_DQ_TMA = 0
  // selects the per-lane red path for dQ instead of the TMA tensor reduce (work-list schedule only)
  static Kernel kernels[2] = {nullptr, nullptr};          // [whole tiles, work list]
  if (!kernels[0]) {
    const char* e = getenv("DTA_BWD_POLY");
    const int poly = e ? atoi(e) : kDefaultBwdPoly;
    const char* t = getenv("DTA_BWD_DQ_TMA");
    const bool dq_tma = t ? atoi(t) != 0 : kDefaultDqTma;
    kernels[0] = poly <= 0 ? tree_attn_bwd_kernel<0, true, false> : poly <= 2 ? tree_attn_bwd_kernel<2, true, false>
                                                                              : tree_attn_bwd_kernel<4, true, false>;
    kernels[1] = dq_tma ? (poly <= 0 ? tree_attn_bwd_kernel<0, true, true> : poly <= 2 ? tree_attn_bwd_kernel<2, true, true>
                                                                                       : tree_attn_bwd_kernel<4, true, true>)
                        : tree_attn_bwd_kernel<0, false, true>;
    for (Kernel k : kernels) DTA_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::total));
  }
  const bool whole_tiles = sh->bwd_whole_tiles != 0 || !enable_split;
  Kernel kernel = kernels[whole_tiles ? 0 : 1];
  const int n_tiles = sh->n_pad / DTA_TILE;
  const ProfileHook hook = take_profile_hook();
  if (use_pair) {
    // whole key-tile PAIRS, no split chunks; the work-list region of the workspace is this mode's scratch
    if ((rc = bwd_pair::launch(q, k, v, dout, aug_lse, aug_delta, dq_accum, dk, dv, meta, sh, reinterpret_cast<int32_t*>(work),
                               hook.start, hook.stop, g_bwd_trace, stream)))
      return rc;
  } else {
  if (hook.start) DTA_CHECK_CUDA(cudaEventRecord(hook.start, stream));
  kernel<<<(whole_tiles ? n_tiles : wl.work_cap) * sh->n_kv_heads, kBwdThreads, BwdSmem::total, stream>>>(tq, tk, tv, tdo, tdq, tdk, tdv, p);
  DTA_CHECK_CUDA(cudaGetLastError());
  if (hook.stop) DTA_CHECK_CUDA(cudaEventRecord(hook.stop, stream));
  }

  {
    // (fp32 dq requested: only the split key tiles' dk / dv are converted; n_nodes = 0 empties the dq loop)
    const int64_t n_vec = dq_f32 ? 1 : static_cast<int64_t>(sh->n_nodes) * sh->n_q_heads * (kT / 8);
    int64_t grid = (n_vec + 255) / 256;
    const int64_t cap = 16 * static_cast<int64_t>(sm_count());
    if (grid > cap) grid = cap;
    attn_dq_convert_kernel<<<static_cast<int>(dq_f32 ? 4 * sm_count() : grid), 256, 0, stream>>>(
        dq_accum, sh->n_nodes, dq_f32 ? 0 : sh->n_nodes, sh->n_q_heads, sh->q_row_stride, sh->q_head_stride, static_cast<__nv_bfloat16*>(dq),
        split_list, n_work + 1, sh->n_kv_heads, dk_acc, dv_acc, sh->k_row_stride, sh->k_head_stride, sh->v_row_stride,
        sh->v_head_stride, static_cast<__nv_bfloat16*>(dk), static_cast<__nv_bfloat16*>(dv));
    DTA_CHECK_CUDA(cudaGetLastError());
  }
  return DTA_OK;
}

extern "C" {

int dta_tree_attn_bwd(const void* q, const void* k, const void* v, const void* out, const void* dout,
                      const float* lse, void* dq, void* dk, void* dv, const dta_attn_meta_t* meta,
                      const dta_attn_shape_t* sh, void* workspace, dta_stream_t stream) {
  DTA_REQUIRE(dq != nullptr, "NULL pointer argument");
  return tree_attn_bwd_impl(q, k, v, out, dout, lse, dq, nullptr, dk, dv, meta, sh, workspace, stream);
}

int dta_tree_attn_bwd_f32dq(const void* q, const void* k, const void* v, const void* out, const void* dout,
                            const float* lse, float* dq_f32, void* dk, void* dv, const dta_attn_meta_t* meta,
                            const dta_attn_shape_t* sh, void* workspace, dta_stream_t stream) {
  DTA_REQUIRE(dq_f32 != nullptr, "NULL pointer argument");
  return tree_attn_bwd_impl(q, k, v, out, dout, lse, nullptr, dq_f32, dk, dv, meta, sh, workspace, stream);
}

}  // extern "C"
