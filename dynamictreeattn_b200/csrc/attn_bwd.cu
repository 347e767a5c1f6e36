// attn_bwd.cu — tree-masked flash attention backward for sm_100a (tcgen05 + TMEM + TMA).
//
// Replaces the autograd backward through HF attention that the reference runs once per
// popped block, including its O(prefix) grad-KV accumulation
// (tree_training_engine.py:404-465), by ONE pass per layer over the linearised trie.
//
// Work decomposition
//   CTA        = one 128-key tile x one KV head.  It loops over the query heads of the GQA
//                group and over the key tile's LIVE query tiles, which form one contiguous
//                range [tile, k_qlast[tile]] because every key's visible queries are the
//                interval [k, subtree_end[k]).  dK/dV accumulate in TMEM across the whole
//                loop (no atomics); dQ tiles are reduced into an fp32 accumulator with
//                red.global.add.f32 (coalesced: lanes = head-dim index) and converted to bf16 by a second kernel.
//   per step   = 5 GEMMs, all 128x128x128 bf16 -> fp32 on tcgen05:
//                  S^T  = K  Q^T     (SS, both K-major)            -> TMEM R0
//                  dP^T = V  dO^T    (SS, both K-major)            -> TMEM R1
//                  dV  += P^T dO     (A = P^T bf16 in TMEM, B = dO MN-major)
//                  dQ^T = K^T dS^T   (A = K read MN-major, B = dS^T read MN-major) -> TMEM R1
//                  dK  += dS^T Q     (A = dS^T smem K-major, B = Q MN-major)
//                computed transposed so that TMEM lanes = key rows: the ancestor mask is
//                then a per-thread column RANGE test [k, subtree_end[k]).
//   warp roles = warps 0-7 compute (lane = key row, two column halves) | warps 8-11 dQ drain
//                (lane = head-dim index) | warp 14 TMA producer | warp 15 MMA issuer + TMEM owner.
//                The issuer sits in the highest warp slot: the SM arbiter favours high warp ids
//                and the issuer is the one warp whose latency is on every critical path.
//   TMEM       = R0 S^T/P^T [0,128) (P^T bf16: queries 0-63 in cols [0,32), 64-127 in cols [64,96))
//                  R1 dP^T/dQ [128,256)  dV [256,384)  dK [384,512)
#include "dta_common.cuh"
#include "dta_ptx.cuh"
#include "../../include/dta_b200.h"
#include <cmath>
#include <cstdlib>

namespace dta {

static constexpr int kT = 128;                       // tile edge (queries, keys) and head dim
static constexpr int kHalf = kT * 128;               // [128 rows][64 bf16] swizzled half tile, bytes
static constexpr int kTile = 2 * kHalf;              // 32 KB
static constexpr int kBwdThreads = 512;
static constexpr int kProducerWarp = 14, kMmaWarp = 15;
static constexpr float kLog2e = 1.4426950408889634f;
static constexpr int kDefaultBwdPoly = 2;            // of every 8 exponentials, on the FMA pipes

struct BwdSmem {
  static constexpr int k = 0;
  static constexpr int v = k + kTile;
  static constexpr int q = v + kTile;                 // 2 stages
  static constexpr int d_o = q + 2 * kTile;           // 2 stages
  static constexpr int ds = d_o + 2 * kTile;
  static constexpr int lse = ds + kTile;              // float[2][128]
  static constexpr int delta = lse + 2 * kT * 4;      // float[2][128]
  static constexpr int bars = delta + 2 * kT * 4;
  static constexpr int n_bars = 1 + 8 + 8;
  static constexpr int misc = bars + n_bars * 8;
  static constexpr int total = misc + 16;
};
static_assert(BwdSmem::total <= 232448, "backward kernel shared memory exceeds 227 KB");

struct BwdParams {
  const float* lse;        // [Hq, n_pad]  log2(scale) - lse2  (written by the delta pre-pass)
  const float* delta;      // [Hq, n_pad]
  float* dq_accum;         // [n_pad, Hq, 128] fp32
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;
  const int32_t* subtree_end;
  const int32_t* kmin_end;
  const int32_t* k_order;
  const int32_t* k_qlast;
  int n_nodes, n_pad, n_q_heads, n_kv_heads, group;
  int64_t k_row_stride, k_head_stride, v_row_stride, v_head_stride;
  float scale, scale_log2, inv_scale;
  int debug_flags;         // bit 0: skip the dQ reds (timing experiments only)
  long long* trace;        // debug timeline (NULL in production): [step][16] clock64 stamps of CTA 0
};

#define DTA_TRACE(ev)                                                                         \
  do {                                                                                        \
    if (p.trace != nullptr && blockIdx.x == 0 && it < 48) p.trace[it * 16 + (ev)] = clock64(); \
  } while (0)

__device__ __forceinline__ void red_add_f32(float* addr, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}

// kPolyOf8: how many of every 8 exponentials of the P^T pass run on the FMA pipes (poly_exp2)
template <int kPolyOf8>
__global__ void __launch_bounds__(kBwdThreads, 1)
tree_attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                     const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem + BwdSmem::k;
  uint8_t* sV = smem + BwdSmem::v;
  uint8_t* sQ = smem + BwdSmem::q;
  uint8_t* sdO = smem + BwdSmem::d_o;
  uint8_t* sdS = smem + BwdSmem::ds;
  float* sLse = reinterpret_cast<float*>(smem + BwdSmem::lse);
  float* sDelta = reinterpret_cast<float*>(smem + BwdSmem::delta);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BwdSmem::bars);
  uint64_t* bar_kv_full = bars;
  uint64_t* bar_q_full = bars + 1;      // [2]
  uint64_t* bar_q_empty = bars + 3;     // [2]
  uint64_t* bar_do_full = bars + 5;     // [2]
  uint64_t* bar_do_empty = bars + 7;    // [2]
  uint64_t* bar_st_full = bars + 9;     // S^T ready
  uint64_t* bar_p_full = bars + 10;     // P^T written (256 arrivals)
  uint64_t* bar_dp_full = bars + 11;    // dP^T ready
  uint64_t* bar_ds_full = bars + 12;    // dS^T in smem (256 arrivals)
  uint64_t* bar_ds_empty = bars + 13;   // dK/dQ MMAs of the step done -> sdS reusable
  uint64_t* bar_dq_full = bars + 14;    // dQ tile ready in R1
  uint64_t* bar_r1_free = bars + 15;    // dQ tile copied out of R1 (128 arrivals)
  uint64_t* bar_dkv_full = bars + 16;   // everything done
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + BwdSmem::misc);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int jt = p.k_order[blockIdx.x / p.n_kv_heads];   // key tile
  const int kvh = blockIdx.x % p.n_kv_heads;
  const int k0 = jt * kT;
  const int n_qt = p.k_qlast[jt] - jt + 1;
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
  }
  if (warp == kMmaWarp) {
    if (lane == 0) {
      mbar_init(bar_kv_full, 1);
      for (int s = 0; s < 2; ++s) {
        mbar_init(bar_q_full + s, 1);
        mbar_init(bar_q_empty + s, 1);
        mbar_init(bar_do_full + s, 1);
        mbar_init(bar_do_empty + s, 1);
      }
      mbar_init(bar_st_full, 1);
      mbar_init(bar_p_full, 8);
      mbar_init(bar_dp_full, 1);
      mbar_init(bar_ds_full, 8);
      mbar_init(bar_ds_empty, 1);
      mbar_init(bar_dq_full, 1);
      mbar_init(bar_r1_free, 4);
      mbar_init(bar_dkv_full, 1);
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
      for (int it = 0; it < n_iter; ++it) {
        const int head = kvh * p.group + it / n_qt;
        const int q0 = (jt + it % n_qt) * kT;
        const int st = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        mbar_wait(bar_q_empty + st, ph ^ 1);
        mbar_arrive_expect_tx(bar_q_full + st, kTile + kT * 4);
        for (int h = 0; h < 2; ++h) tma_load_3d(sQ + st * kTile + h * kHalf, &tmap_q, bar_q_full + st, h * 64, head, q0);
        bulk_load_1d(sLse + st * kT, p.lse + static_cast<int64_t>(head) * p.n_pad + q0, kT * 4, bar_q_full + st);
        mbar_wait(bar_do_empty + st, ph ^ 1);
        mbar_arrive_expect_tx(bar_do_full + st, kTile + kT * 4);
        for (int h = 0; h < 2; ++h) tma_load_3d(sdO + st * kTile + h * kHalf, &tmap_do, bar_do_full + st, h * 64, head, q0);
        bulk_load_1d(sDelta + st * kT, p.delta + static_cast<int64_t>(head) * p.n_pad + q0, kT * 4, bar_do_full + st);
        // warm L2 for the tiles two steps ahead: with only two shared-memory stages the next load cannot be
        // issued before this step's dK retires, so its latency must be an L2 hit, not a DRAM round trip
        if ((p.debug_flags & 16) && it + 2 < n_iter) {
          const int head2 = kvh * p.group + (it + 2) / n_qt;
          const int q2 = (jt + (it + 2) % n_qt) * kT;
          for (int h = 0; h < 2; ++h) {
            tma_prefetch_l2_3d(&tmap_q, h * 64, head2, q2);
            tma_prefetch_l2_3d(&tmap_do, h * 64, head2, q2);
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ================================ MMA issuer =========================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    // The WHOLE warp runs this loop (waits are warp-uniform) and only the tcgen05 instructions are
    // predicated on one elected lane: descriptor arithmetic then stays on the uniform datapath instead
    // of costing a vector->uniform register move per operand, which made issue (not the tensor pipe)
    // the limiter when a single divergent thread ran the loop.
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
      constexpr uint32_t kStage = kTile >> 4;                       // next Q / dO stage
#define DTA_KOFF(kk) ((((kk) >> 2) * (kHalf >> 4)) + (((kk) & 3) * 2))   /* K-major: 16 elements along the row */
#define DTA_NOFF(kk) ((kk) * (2048 >> 4))                                /* MN-major: 16 rows of the tile     */

      auto issue_st = [&](uint32_t st) {       // S^T = K Q^T
        if (elect_one()) {
          const uint32_t b = kQ + st * kStage;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_ss_lo(R0, kK + DTA_KOFF(kk), b + DTA_KOFF(kk), idesc_kk, kk > 0);
          umma_commit(bar_st_full);
        }
        __syncwarp();
      };
      auto issue_dp = [&](uint32_t st) {       // dP^T = V dO^T
        if (elect_one()) {
          const uint32_t b = kDO + st * kStage;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_ss_lo(R1, kV + DTA_KOFF(kk), b + DTA_KOFF(kk), idesc_kk, kk > 0);
          umma_commit(bar_dp_full);
        }
        __syncwarp();
      };

      mbar_wait(bar_kv_full, 0);
      mbar_wait(bar_q_full + 0, 0);
      tc_fence_after();
      issue_st(0);
      mbar_wait(bar_do_full + 0, 0);
      tc_fence_after();
      issue_dp(0);

      for (int it = 0; it < n_iter; ++it) {
        const uint32_t st = it & 1;
        const bool has_next = it + 1 < n_iter;
        const uint32_t nst = (it + 1) & 1;
        const uint32_t nph = ((it + 1) >> 1) & 1;
        const uint32_t not_first = it > 0 ? 1u : 0u;
        // dV += P^T dO
        mbar_wait(bar_p_full, it & 1);
        if (lane == 0) DTA_TRACE(0);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b = nDO + st * kStage;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_ts_lo(RdV, R0 + (kk >> 2) * 64 + (kk & 3) * 8, b + DTA_NOFF(kk), idesc_kn, (kk > 0) ? 1u : not_first);
        }
        __syncwarp();
        // S^T of the next step reuses R0 right behind dV (tensor pipe executes in issue order)
        if (has_next) {
          mbar_wait(bar_q_full + nst, nph);
          tc_fence_after();
          issue_st(nst);
        }
        // dQ^T = K^T dS^T (first, so the drain overlaps dK), dK += dS^T Q
        if (lane == 0) DTA_TRACE(1);
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
          umma_commit(bar_ds_empty);
          umma_commit(bar_q_empty + st);
          umma_commit(bar_do_empty + st);
        }
        __syncwarp();
        if (lane == 0) DTA_TRACE(3);
        // dP^T of the next step needs R1 back from the drain warps
        if (has_next) {
          mbar_wait(bar_do_full + nst, nph);
          mbar_wait(bar_r1_free, it & 1);
          if (lane == 0) DTA_TRACE(4);
          tc_fence_after();
          issue_dp(nst);
          if (lane == 0) DTA_TRACE(5);
        }
      }
      if (elect_one()) umma_commit(bar_dkv_full);
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
      const int qt = jt + it % n_qt;
      const int qbase = qt * kT + ch * 64;
      const int st = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const bool partial = !(qt > jt && kmin_end >= (qt + 1) * kT);
      const float* lse = sLse + st * kT + ch * 64;
      const float* dlt = sDelta + st * kT + ch * 64;
      const int vis_lo = k_idx - qbase, vis_hi = k_end - qbase;      // visible columns of this row: [lo, hi)

      // ---- P^T = exp2(S^T * scale_log2 - lse * log2e), masked -------------------------
      mbar_wait(bar_q_full + st, ph);      // LSE tile landed
      mbar_wait(bar_st_full, it & 1);
      if (threadIdx.x == 0) DTA_TRACE(6);
      tc_fence_after();
      float pr[64];
      tmem_ld_x32(R0 + lane_addr + ch * 64, reinterpret_cast<uint32_t*>(pr));
      tmem_ld_x32(R0 + lane_addr + ch * 64 + 32, reinterpret_cast<uint32_t*>(pr) + 32);
      tmem_ld_wait();
      // pr = scale * P^T = 2^(s*scale_log2 + log2(scale) - lse2); the 1/scale is undone once, in the dV
      // epilogue, so that dS^T = pr * (dP^T - delta) needs no further multiply.  Two code copies: tiles
      // the ancestor mask leaves untouched pay nothing for it (a predicated select still costs its
      // issue slot, and most tiles of a deep trie are mask-free).
      if (!partial) {
#pragma unroll
        for (int c4 = 0; c4 < 16; ++c4) {
          const float4 l4 = reinterpret_cast<const float4*>(lse)[c4];
          const float lv[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c = c4 * 4 + u;
            const float x = fmaf(pr[c], p.scale_log2, lv[u]);
            pr[c] = exp2_slot_on_fma<kPolyOf8>(c) ? poly_exp2(x) : fast_exp2(x);
          }
        }
      } else {
#pragma unroll
        for (int c4 = 0; c4 < 16; ++c4) {
          const float4 l4 = reinterpret_cast<const float4*>(lse)[c4];
          const float lv[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c = c4 * 4 + u;
            const float x = fmaf(pr[c], p.scale_log2, lv[u]);
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

      // ---- dS^T = P^T o (dP^T - delta) * scale -> smem (bf16, 128B-swizzled) ----------
      mbar_wait(bar_do_full + st, ph);     // delta tile landed
      mbar_wait(bar_dp_full, it & 1);
      if (threadIdx.x == 0) DTA_TRACE(8);
      tc_fence_after();
      uint32_t dsp[32];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {        // two 32-column halves: delta loads overlap the TMEM load
        float dp[32];
        tmem_ld_x32(R1 + lane_addr + ch * 64 + hf * 32, reinterpret_cast<uint32_t*>(dp));
        float4 d4[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) d4[i] = reinterpret_cast<const float4*>(dlt)[hf * 8 + i];
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = hf * 32 + i * 4;
          dsp[hf * 16 + i * 2] = pack_bf16x2(pr[c] * (dp[i * 4] - d4[i].x), pr[c + 1] * (dp[i * 4 + 1] - d4[i].y));
          dsp[hf * 16 + i * 2 + 1] = pack_bf16x2(pr[c + 2] * (dp[i * 4 + 2] - d4[i].z), pr[c + 3] * (dp[i * 4 + 3] - d4[i].w));
        }
      }
      tc_fence_before();                   // our R1 reads are done before the arrive below
      if (it > 0) mbar_wait(bar_ds_empty, (it - 1) & 1);
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
    if (true) {
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
            const float is = p.inv_scale;      // P^T was carried pre-multiplied by scale
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
    // ================================ dQ drain warps =====================================
    // R1 holds dQ^T: TMEM lane = head-dim index d, column = query.  A warp-level red therefore adds 32
    // consecutive floats of ONE dq row: 128 contiguous bytes = 4 full sectors (the untransposed layout
    // touched 32 half-used sectors per instruction and doubled the L2 atomic traffic).
    asm volatile("setmaxnreg.inc.sync.aligned.u32 136;");
    const int d = ((warp & 3) << 5) | lane;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) << 5) << 16;
    for (int it = 0; it < n_iter; ++it) {
      const int head = kvh * p.group + it / n_qt;
      const int q0 = (jt + it % n_qt) * kT;
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
      const int n_valid = (p.debug_flags & 1) ? 0 : min(kT, p.n_nodes - q0);
      if (n_valid == kT) {
#pragma unroll
        for (int c = 0; c < kT; ++c) red_add_f32(dst + c * row, acc[c]);
      } else {
#pragma unroll
        for (int c = 0; c < kT; ++c)
          if (c < n_valid) red_add_f32(dst + c * row, acc[c]);
      }
      if (threadIdx.x == 256) DTA_TRACE(12);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------
// delta[h, q] = sum_d dO[q,h,d] * O[q,h,d]      (one warp per (q, h) row; HBM-bound)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ out,
                                                         const __nv_bfloat16* __restrict__ dout, int n_nodes,
                                                         int n_pad, int n_heads, int64_t row_stride,
                                                         int64_t head_stride, float* __restrict__ delta,
                                                         const float* __restrict__ lse2, float lshift,
                                                         float* __restrict__ lse_shift) {
  const int lane = threadIdx.x & 31;
  const int64_t n_rows = static_cast<int64_t>(n_pad) * n_heads;
  for (int64_t w = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; w < n_rows;
       w += (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5) {
    const int q = static_cast<int>(w / n_heads), h = static_cast<int>(w % n_heads);
    float acc = 0.f;
    if (q < n_nodes) {
      const int64_t off = q * row_stride + h * head_stride + lane * 4;
      const uint2 a = *reinterpret_cast<const uint2*>(out + off);
      const uint2 b = *reinterpret_cast<const uint2*>(dout + off);
      const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
      const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float2 x = __bfloat1622float2(a2[i]), y = __bfloat1622float2(b2[i]);
        acc = fmaf(x.x, y.x, acc);
        acc = fmaf(x.y, y.y, acc);
      }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      const int64_t o = static_cast<int64_t>(h) * n_pad + q;
      delta[o] = acc;
      lse_shift[o] = q < n_nodes ? lshift - lse2[o] : 0.f;      // exponent offset used by the main kernel
    }
  }
}

// dq (bf16, q strides) <- dq_accum (fp32 [n_pad, Hq, 128])
__global__ void __launch_bounds__(256) attn_dq_convert_kernel(const float* __restrict__ acc, int n_nodes, int n_heads,
                                                              int64_t row_stride, int64_t head_stride,
                                                              __nv_bfloat16* __restrict__ dq) {
  const int64_t n_vec = static_cast<int64_t>(n_nodes) * n_heads * (kT / 8);
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

static long long* g_bwd_trace = nullptr;

extern "C" {

/* debug only (not in the public header): device buffer of 48*16 int64 receiving the step timeline of CTA 0 */
int dta_debug_set_bwd_trace(long long* buf) { g_bwd_trace = buf; return 0; }

size_t dta_tree_attn_bwd_workspace_bytes(const dta_attn_shape_t* sh) {
  if (!sh) return 0;
  const size_t dq = static_cast<size_t>(sh->n_pad) * sh->n_q_heads * kT * sizeof(float);
  const size_t delta = static_cast<size_t>(sh->n_pad) * sh->n_q_heads * sizeof(float);
  return dq + 2 * delta + 256;
}

int dta_tree_attn_bwd(const void* q, const void* k, const void* v, const void* out, const void* dout,
                      const float* lse, void* dq, void* dk, void* dv, const dta_attn_meta_t* meta,
                      const dta_attn_shape_t* sh, void* workspace, dta_stream_t stream) {
  DTA_REQUIRE(q && k && v && out && dout && lse && dq && dk && dv && meta && sh && workspace, "NULL pointer argument");
  DTA_REQUIRE(sh->head_dim == kT, "head_dim must be %d (got %d)", kT, sh->head_dim);
  DTA_REQUIRE(sh->n_nodes >= 1 && sh->n_pad % DTA_TILE == 0 && sh->n_pad >= sh->n_nodes &&
                  sh->n_pad - sh->n_nodes < DTA_TILE,
              "n_pad must be n_nodes rounded up to %d", DTA_TILE);
  DTA_REQUIRE(sh->n_kv_heads >= 1 && sh->n_q_heads % sh->n_kv_heads == 0, "n_q_heads must be a multiple of n_kv_heads");
  DTA_REQUIRE(meta->subtree_end && meta->kmin_end && meta->k_order && meta->k_qlast, "NULL metadata pointer");
  DTA_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  DTA_REQUIRE((reinterpret_cast<uintptr_t>(lse) & 15) == 0, "lse must be 16-byte aligned");
  const uintptr_t align_or = reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) |
                             reinterpret_cast<uintptr_t>(dv) | reinterpret_cast<uintptr_t>(out) |
                             reinterpret_cast<uintptr_t>(dout);
  DTA_REQUIRE((align_or & 15) == 0, "tensors must be 16-byte aligned");
  DTA_REQUIRE(sh->q_row_stride % 8 == 0 && sh->q_head_stride % 8 == 0 && sh->k_row_stride % 8 == 0 &&
                  sh->k_head_stride % 8 == 0 && sh->v_row_stride % 8 == 0 && sh->v_head_stride % 8 == 0 &&
                  sh->o_row_stride % 8 == 0 && sh->o_head_stride % 8 == 0,
              "strides must be multiples of 8 elements");

  float* dq_accum = static_cast<float*>(workspace);
  const size_t dq_bytes = static_cast<size_t>(sh->n_pad) * sh->n_q_heads * kT * sizeof(float);
  float* delta = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + dq_bytes);
  float* lse_shift = delta + static_cast<size_t>(sh->n_pad) * sh->n_q_heads;
  DTA_CHECK_CUDA(cudaMemsetAsync(dq_accum, 0, dq_bytes, stream));

  const int64_t n_rows = static_cast<int64_t>(sh->n_pad) * sh->n_q_heads;
  {
    int64_t grid = (n_rows * 32 + 255) / 256;
    const int64_t cap = 16 * static_cast<int64_t>(sm_count());
    if (grid > cap) grid = cap;
    attn_delta_kernel<<<static_cast<int>(grid), 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(dout), sh->n_nodes, sh->n_pad,
        sh->n_q_heads, sh->o_row_stride, sh->o_head_stride, delta, lse, log2f(sh->scale), lse_shift);
    DTA_CHECK_CUDA(cudaGetLastError());
  }

  CUtensorMap tq, tk, tv, tdo;
  int rc;
  if ((rc = make_tmap_bf16_rows(&tq, q, sh->n_nodes, sh->n_q_heads, kT, sh->q_row_stride, sh->q_head_stride, kT))) return rc;
  if ((rc = make_tmap_bf16_rows(&tk, k, sh->n_nodes, sh->n_kv_heads, kT, sh->k_row_stride, sh->k_head_stride, kT))) return rc;
  if ((rc = make_tmap_bf16_rows(&tv, v, sh->n_nodes, sh->n_kv_heads, kT, sh->v_row_stride, sh->v_head_stride, kT))) return rc;
  if ((rc = make_tmap_bf16_rows(&tdo, dout, sh->n_nodes, sh->n_q_heads, kT, sh->o_row_stride, sh->o_head_stride, kT))) return rc;

  BwdParams p;
  p.lse = lse_shift;
  p.delta = delta;
  p.dq_accum = dq_accum;
  p.dk = static_cast<__nv_bfloat16*>(dk);
  p.dv = static_cast<__nv_bfloat16*>(dv);
  p.subtree_end = meta->subtree_end;
  p.kmin_end = meta->kmin_end;
  p.k_order = meta->k_order;
  p.k_qlast = meta->k_qlast;
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
  p.debug_flags = getenv("DTA_DEBUG_FLAGS") ? atoi(getenv("DTA_DEBUG_FLAGS")) : 0;
  p.trace = g_bwd_trace;

  using Kernel = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const BwdParams);
  static Kernel kernel = nullptr;
  if (!kernel) {
    // share of the exponentials computed on the FMA pipes; DTA_BWD_POLY (0..4 of every 8) is a tuning knob
    const char* e = getenv("DTA_BWD_POLY");
    const int poly = e ? atoi(e) : kDefaultBwdPoly;
    kernel = poly <= 0 ? tree_attn_bwd_kernel<0> : poly == 1 ? tree_attn_bwd_kernel<1>
           : poly == 2 ? tree_attn_bwd_kernel<2> : poly == 3 ? tree_attn_bwd_kernel<3> : tree_attn_bwd_kernel<4>;
    DTA_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::total));
  }
  const int n_tiles = sh->n_pad / DTA_TILE;
  const ProfileHook hook = take_profile_hook();
  if (hook.start) DTA_CHECK_CUDA(cudaEventRecord(hook.start, stream));
  kernel<<<n_tiles * sh->n_kv_heads, kBwdThreads, BwdSmem::total, stream>>>(tq, tk, tv, tdo, p);
  DTA_CHECK_CUDA(cudaGetLastError());
  if (hook.stop) DTA_CHECK_CUDA(cudaEventRecord(hook.stop, stream));

  {
    const int64_t n_vec = static_cast<int64_t>(sh->n_nodes) * sh->n_q_heads * (kT / 8);
    int64_t grid = (n_vec + 255) / 256;
    const int64_t cap = 16 * static_cast<int64_t>(sm_count());
    if (grid > cap) grid = cap;
    attn_dq_convert_kernel<<<static_cast<int>(grid), 256, 0, stream>>>(dq_accum, sh->n_nodes, sh->n_q_heads,
                                                                       sh->q_row_stride, sh->q_head_stride,
                                                                       static_cast<__nv_bfloat16*>(dq));
    DTA_CHECK_CUDA(cudaGetLastError());
  }
  return DTA_OK;
}

}  // extern "C"
