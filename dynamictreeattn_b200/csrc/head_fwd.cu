// head_fwd.cu — lm_head fused into the log-prob / entropy statistics.
//
// The reference materialises the logits of every block it scores: HF computes `logits = lm_head(hidden)`
// (HF models/qwen3/modeling_qwen3.py:505) and vocab_parallel.py:13-27 then runs log_softmax / gather / entropy over
// the [B, V] tensor (fp32 copies included).  Here the [rows, V] logits never exist: one kernel runs the
// hidden x W^T GEMM on tcgen05 and reduces every 128 x 256 accumulator tile straight out of TMEM into the 16-byte
// row statistics {max, sum e^(x-max), sum x e^(x-max), -} that the log-prob path (and the vocab-parallel exchange)
// is built on.  HBM traffic: hidden and W once -- not the 2 * rows * V bytes written and read again (20 GB + 20 GB per
// step at 67k rows x 151,936 columns).
//
// Structure (one cluster = 2 CTAs = 256 rows x a range of vocab tiles):
//   * tcgen05.mma.cta_group::2, M = 256 (128 rows per CTA), N = 256 vocab columns, K = 64 per stage: each CTA stages
//     its own 128 rows of `hidden` (A) and HALF of the W tile (B, split along N), so an SM fetches 32 KB per
//     256 x 256 x 64 block of MACs -- the operand traffic of a 256 x 256 tile, which is what keeps this shape off the
//     L2 -> SM bandwidth limit (a 128 x 256 single-CTA tile needs 1.5x the bytes per FLOP).
//   * 6-stage TMA ring (192 KB); both CTAs' loads complete on the LEADER's "full" barrier, the leader's
//     tcgen05.commit multicasts "empty" to both producers.
//   * the accumulator is double-buffered in TMEM (2 x 256 columns): the MMAs of tile n + 1 run while the four epilogue
//     warps of each CTA (thread = row = TMEM lane) fold tile n into their running statistics with the online-softmax
//     rule; one exponential per element (MUFU), ~1/4 of the tile's MMA time.
//   * rows x vocab is cut into (row pair, vocab split) work items; consecutive clusters take consecutive row pairs of
//     the SAME split, so the clusters resident at one time sweep the same W tiles (W is read from HBM about once per
//     wave).  The per-split statistics are merged by dta_logprob_finalize, exactly like vocab-parallel shards.
// Label logits (one per trie edge) are K-long dot products of rows that are already in L2: lm_head_label_kernel.
#include "dta_common.cuh"
#include "dta_ptx.cuh"
#include "../../include/dta_b200.h"
#include <cuda_bf16.h>

namespace dta {
namespace head {

static constexpr int kBM = 128;                      // rows per CTA (M = 256 per cluster)
static constexpr int kBN = 256;                      // vocab columns per tile (128 staged per CTA)
static constexpr int kBK = 64;                       // K per stage: one 128-byte swizzled row
static constexpr int kBlock = kBM * 128;             // [128 rows][64 bf16] swizzled block, bytes (16 KB)
static constexpr int kStage = 2 * kBlock;            // A block + B half
static constexpr int kStages = 6;
static constexpr int kThreads = 192;                 // warps 0-3 epilogue, 4 producer, 5 MMA issuer
static constexpr int kProducerWarp = 4, kMmaWarp = 5;
static constexpr float kLog2e = 1.4426950408889634f;

struct Smem {
  static constexpr int stages = 0;
  static constexpr int bars = stages + kStages * kStage;
  static constexpr int n_bars = 2 * kStages + 4;
  static constexpr int misc = bars + n_bars * 8;
  static constexpr int total = misc + 16;
};
static_assert(Smem::total + 1024 <= 232448, "fused head kernel shared memory exceeds 227 KB");

struct Params {
  float4* stats;            // [n_split][rows_pad] {m, s, t, 0}; rows_pad = n_row_pairs * 256
  int64_t n_rows, n_cols;   // rows of hidden, vocabulary size (rows of W)
  int k_chunks;             // hidden / 64
  int n_row_pairs, n_tiles, tiles_per_split;
  float inv_temp;
  // backward variant (kBackward): the tile is turned into d loss / d logits and stored in bf16
  const float* logz;        // [n_rows]
  const float* entropy;     // [n_rows]
  const float* g_entropy;   // [n_rows] or NULL
  const float* g_rowsum;    // [n_rows] or NULL: sum of g_lp over the labels of the row
  __nv_bfloat16* dlogits;   // [n_rows][d_stride]
  int64_t d_stride;
};
static constexpr float kLn2 = 0.6931471805599453f;

// ---- cta_group::2 forms (the same ones attn_bwd_pair.cu uses; verified on the B200 by tools/umma_2cta*.cu) --------
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
__device__ __forceinline__ void umma2_ss(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(kSdescHi)
      : "memory");
}
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {       // arrives on `bar` of BOTH CTAs of the pair
  const uint16_t mask = 3;
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
__device__ __forceinline__ void tma2_load_3d(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1,
                                             int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

// kMode 0: statistics epilogue (forward, no logits anywhere).
// kMode 1: the SAME mainloop recomputes the logits tile and the epilogue writes d loss / d logits = inv_temp * p *
//          (g_H * (-H - log p) - sum g_lp) in bf16 (the one-hot label terms are added by dta_logprob_bwd_scatter), i.e.
//          the recompute GEMM and logprob_bwd_kernel in one pass with no logits in HBM.
// kMode 2: statistics AND the bf16 logits (into p.dlogits): for a training step that keeps the logits for its backward
//          -- they are written once, under the MMAs, and not read again in the forward.  The statistics are taken from
//          the ROUNDED values, so they are exactly the statistics of the tensor the backward will read.
template <int kMode>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
lm_head_kernel(const __grid_constant__ CUtensorMap tmap_h, const __grid_constant__ CUtensorMap tmap_w, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
  uint64_t* bar_full = bars;                          // [kStages] leader: the stage of BOTH CTAs landed
  uint64_t* bar_empty = bars + kStages;               // [kStages] both: the stage's MMAs retired (multicast commit)
  uint64_t* bar_acc_full = bar_empty + kStages;       // [2] both: accumulator buffer complete (multicast commit)
  uint64_t* bar_acc_empty = bar_acc_full + 2;         // [2] leader: read by the epilogue warps of both CTAs (8 arrivals)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + Smem::misc);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cid = blockIdx.x >> 1;
  const int split = cid / p.n_row_pairs;              // consecutive clusters: consecutive row pairs, SAME vocab range
  const int mb = cid % p.n_row_pairs;
  const int row0 = (mb * 2 + static_cast<int>(rank)) * kBM;
  const int t0 = split * p.tiles_per_split;
  const int t1 = min(t0 + p.tiles_per_split, p.n_tiles);
  const int n_my_tiles = max(t1 - t0, 0);

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmap_h);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(bar_full + s, 1);
        mbar_init(bar_empty + s, 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(bar_acc_full + b, 1);
        mbar_init(bar_acc_empty + b, 8);
      }
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

  if (warp == kProducerWarp) {
    // ================================ TMA producer (one per CTA) ==========================
    if (lane == 0) {
      int it = 0;
      for (int ti = 0; ti < n_my_tiles; ++ti) {
        const int col0 = (t0 + ti) * kBN + static_cast<int>(rank) * (kBN / 2);     // this CTA's half of the W tile
        for (int kc = 0; kc < p.k_chunks; ++kc, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(bar_empty + s, ph ^ 1);
          const uint32_t full_l = mapa_u32(smem_u32(bar_full + s), 0);
          if (leader) mbar_arrive_expect_tx(bar_full + s, 2 * kStage);
          uint8_t* st = smem + Smem::stages + s * kStage;
          tma2_load_3d(st, &tmap_h, full_l, kc * kBK, 0, row0);
          tma2_load_3d(st + kBlock, &tmap_w, full_l, kc * kBK, 0, col0);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ================================ MMA issuer (leader only) ============================
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * kBM, kBN, false, false);       // M = 256, N = 256, K-major A and B
      const uint32_t a0 = sdesc_lo_kmajor(smem_u32(smem + Smem::stages));
      const uint32_t b0 = sdesc_lo_kmajor(smem_u32(smem + Smem::stages + kBlock));
      constexpr uint32_t kStageW = kStage >> 4;
      int it = 0;
      for (int ti = 0; ti < n_my_tiles; ++ti) {
        const int b = ti & 1;
        mbar_wait(bar_acc_empty + b, ((ti >> 1) & 1) ^ 1);        // first use of each buffer passes at once
        tc_fence_after();
        const uint32_t d = tmem + b * kBN;
        for (int kc = 0; kc < p.k_chunks; ++kc, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(bar_full + s, ph);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < kBK / 16; ++kk)
              umma2_ss(d, a0 + s * kStageW + kk * 2, b0 + s * kStageW + kk * 2, idesc, (kc > 0 || kk > 0) ? 1u : 0u);
            umma2_commit(bar_empty + s);
          }
          __syncwarp();
        }
        if (elect_one()) umma2_commit(bar_acc_full + b);
        __syncwarp();
      }
    }
  } else {
    const int r = (warp << 5) | lane;                              // TMEM lane == row inside this CTA's 128
    const uint32_t lane_addr = static_cast<uint32_t>(warp << 5) << 16;
    const int64_t row = static_cast<int64_t>(row0) + r;
    const float c = p.inv_temp * kLog2e;
    if constexpr (kMode == 1) {
      // ================================ epilogue: d loss / d logits, bf16 ===================
      const bool live = row < p.n_rows;
      const float lz2 = live ? p.logz[row] * kLog2e : 0.f;
      const float a = (live && p.g_entropy) ? p.g_entropy[row] : 0.f;
      const float G = (live && p.g_rowsum) ? p.g_rowsum[row] : 0.f;
      const float k0 = live ? -(a * p.entropy[row] + G) : 0.f;
      const float aln2 = -a * kLn2;
      __nv_bfloat16* drow = p.dlogits + (live ? row : 0) * p.d_stride;
      const bool vec_ok = (p.d_stride % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.dlogits) & 15) == 0);
      for (int ti = 0; ti < n_my_tiles; ++ti) {
        const int b = ti & 1;
        mbar_wait(bar_acc_full + b, (ti >> 1) & 1);
        tc_fence_after();
        const int64_t col0 = static_cast<int64_t>(t0 + ti) * kBN;
        const int64_t left = p.n_cols - col0;
        const int valid = left < kBN ? static_cast<int>(left) : kBN;
#pragma unroll 1
        for (int ch = 0; ch < kBN / 32; ++ch) {
          const int cnt = min(32, valid - ch * 32);
          if (cnt <= 0) break;
          float x[32];
          tmem_ld_x32(tmem + lane_addr + b * kBN + ch * 32, reinterpret_cast<uint32_t*>(x));
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float lp2 = fmaf(x[i], c, -lz2);                 // log2 p
            x[i] = p.inv_temp * fast_exp2(lp2) * fmaf(aln2, lp2, k0);
          }
          if (live) {
            __nv_bfloat16* dst = drow + col0 + ch * 32;
            if (cnt == 32 && vec_ok) {
#pragma unroll
              for (int v = 0; v < 4; ++v) {
                uint4 w4;
                w4.x = pack_bf16x2(x[8 * v + 0], x[8 * v + 1]);
                w4.y = pack_bf16x2(x[8 * v + 2], x[8 * v + 3]);
                w4.z = pack_bf16x2(x[8 * v + 4], x[8 * v + 5]);
                w4.w = pack_bf16x2(x[8 * v + 6], x[8 * v + 7]);
                __stcs(reinterpret_cast<uint4*>(dst) + v, w4);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < cnt) dst[i] = __float2bfloat16_rn(x[i]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(bar_acc_empty + b), 0));
      }
    } else {
    // ================================ epilogue: online softmax statistics =================
    float m2 = -INFINITY;     // running max of x * log2e (x = logit * inv_temp)
    float s = 0.f, t = 0.f;   // sum 2^(x2 - m2), sum logit * 2^(x2 - m2)
    for (int ti = 0; ti < n_my_tiles; ++ti) {
      const int b = ti & 1;
      mbar_wait(bar_acc_full + b, (ti >> 1) & 1);
      tc_fence_after();
      const int64_t col0 = static_cast<int64_t>(t0 + ti) * kBN;
      const int64_t left = p.n_cols - col0;
      const int valid = left < kBN ? static_cast<int>(left) : kBN;      // < 256 only in the last tile
#pragma unroll 1
      for (int ch = 0; ch < kBN / 32; ++ch) {
        const int cnt = min(32, valid - ch * 32);
        if (cnt <= 0) break;
        float x[32];
        tmem_ld_x32(tmem + lane_addr + b * kBN + ch * 32, reinterpret_cast<uint32_t*>(x));
        tmem_ld_wait();
        if constexpr (kMode == 2) {
          // keep the logits: round to bf16, store, and continue with the rounded values
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            pk[i] = pack_bf16x2(x[2 * i], x[2 * i + 1]);
            x[2 * i] = __uint_as_float(pk[i] << 16);
            x[2 * i + 1] = __uint_as_float(pk[i] & 0xffff0000u);
          }
          if (row < p.n_rows) {
            __nv_bfloat16* dst = p.dlogits + row * p.d_stride + col0 + ch * 32;
            if (cnt == 32 && (p.d_stride % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.dlogits) & 15) == 0)) {
#pragma unroll
              for (int v = 0; v < 4; ++v)
                __stcs(reinterpret_cast<uint4*>(dst) + v, make_uint4(pk[4 * v], pk[4 * v + 1], pk[4 * v + 2], pk[4 * v + 3]));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < cnt) dst[i] = __float2bfloat16_rn(x[i]);
            }
          }
        }
        float mx = -INFINITY;
        if (cnt == 32) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, x[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) if (i < cnt) mx = fmaxf(mx, x[i]);
        }
        mx *= c;                                                   // inv_temp > 0
        if (mx > m2) {
          const float f = fast_exp2(m2 - mx);                      // 2^(-inf) = 0 on the first chunk
          s *= f;
          t *= f;
          m2 = mx;
        }
        if (cnt == 32) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float e = fast_exp2(fmaf(x[i], c, -m2));
            s += e;
            t = fmaf(x[i], e, t);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i < cnt) {
              const float e = fast_exp2(fmaf(x[i], c, -m2));
              s += e;
              t = fmaf(x[i], e, t);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(bar_acc_empty + b), 0));
    }
    // same layout as row_stats_kernel: {max x, sum e^(x - max), sum x e^(x - max), 0}, natural-log domain; a split
    // without tiles writes the identity of the merge {-inf, 0, 0}
    const int64_t rows_pad = static_cast<int64_t>(p.n_row_pairs) * 2 * kBM;
    p.stats[static_cast<int64_t>(split) * rows_pad + row] =
        make_float4(m2 == -INFINITY ? -INFINITY : m2 / kLog2e, s, t * p.inv_temp, 0.f);
  }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                 // no arrival or MMA of the peer may target a CTA that has left
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem, 512);
  }
}

// x_label[m] = inv_temp * <hidden[row_of[m]], W[labels[m]]>  (fp32 accumulation of bf16 products, like the MMA);
// one warp per label, 16-byte loads; row_of[m] < 0 (a root: no incoming edge) or a label outside [0, n_cols) gives 0.
__global__ void __launch_bounds__(256) lm_head_label_kernel(const __nv_bfloat16* __restrict__ hidden, int64_t h_stride,
                                                            const __nv_bfloat16* __restrict__ w, int64_t w_stride, int k,
                                                            int64_t n_cols, const int64_t* __restrict__ labels,
                                                            const int32_t* __restrict__ row_of, int64_t n_labels,
                                                            float inv_temp, float* __restrict__ x_label) {
  const int lane = threadIdx.x & 31;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t mi = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; mi < n_labels; mi += n_warps) {
    const int64_t r = row_of ? row_of[mi] : mi;
    const int64_t lab = labels[mi];
    float acc = 0.f;
    if (r >= 0 && lab >= 0 && lab < n_cols) {
      const uint4* a = reinterpret_cast<const uint4*>(hidden + r * h_stride);
      const uint4* b = reinterpret_cast<const uint4*>(w + lab * w_stride);
      for (int v = lane; v < k / 8; v += 32) {
        const uint4 av = __ldg(a + v), bv = __ldg(b + v);
        const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&av);
        const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&bv);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 xa = __bfloat1622float2(a2[i]), xb = __bfloat1622float2(b2[i]);
          acc = fmaf(xa.x, xb.x, acc);
          acc = fmaf(xa.y, xb.y, acc);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) x_label[mi] = acc * inv_temp;
  }
}

static int plan_splits(int n_row_pairs, int n_tiles) {
  // ~4 waves of clusters (74 resident pairs on 148 SMs) so the tail is short; at most one tile per split
  int n_split = (4 * (sm_count() / 2) + n_row_pairs - 1) / n_row_pairs;
  if (n_split < 1) n_split = 1;
  if (n_split > 32) n_split = 32;
  if (n_split > n_tiles) n_split = n_tiles;
  return n_split;
}

}  // namespace head
}  // namespace dta

using namespace dta;

extern "C" {

int dta_lm_head_stats_splits(int64_t n_rows, int64_t n_cols) {
  if (n_rows < 1 || n_cols < 1) return 0;
  const int n_row_pairs = static_cast<int>((n_rows + 2 * head::kBM - 1) / (2 * head::kBM));
  const int n_tiles = static_cast<int>((n_cols + head::kBN - 1) / head::kBN);
  return head::plan_splits(n_row_pairs, n_tiles);
}

int64_t dta_lm_head_stats_rows_pad(int64_t n_rows) { return (n_rows + 2 * head::kBM - 1) / (2 * head::kBM) * (2 * head::kBM); }

static int launch_lm_head(int mode, const void* hidden, int64_t h_row_stride, const void* weight, int64_t w_row_stride,
                          int64_t n_rows, int64_t n_cols, int hidden_size, float inv_temperature, float* stats,
                          const float* logz, const float* entropy, const float* g_entropy, const float* g_rowsum,
                          void* dlogits, int64_t d_row_stride, dta_stream_t stream) {
  const bool backward = mode == 1;
  DTA_REQUIRE(hidden && weight && (backward ? (logz && entropy && dlogits) : stats != nullptr) && (mode != 2 || dlogits),
              "NULL pointer argument");
  DTA_REQUIRE(n_rows >= 1 && n_cols >= 1, "bad shape");
  DTA_REQUIRE(hidden_size >= head::kBK && hidden_size % head::kBK == 0, "hidden size must be a multiple of %d (got %d)",
              head::kBK, hidden_size);
  DTA_REQUIRE(inv_temperature > 0.f, "temperature must be positive");
  DTA_REQUIRE(backward || (reinterpret_cast<uintptr_t>(stats) & 15) == 0, "stats must be 16-byte aligned");
  DTA_REQUIRE(mode == 0 || d_row_stride >= n_cols, "logits / dlogits row stride shorter than a row");
  const int n_row_pairs = static_cast<int>((n_rows + 2 * head::kBM - 1) / (2 * head::kBM));
  const int n_tiles = static_cast<int>((n_cols + head::kBN - 1) / head::kBN);
  const int n_split = head::plan_splits(n_row_pairs, n_tiles);
  CUtensorMap th, tw;
  int rc;
  // 3-D maps (k, 1, rows): box (64, 1, 128), 128-byte swizzle; rows / columns past the end read as zeros
  if ((rc = make_tmap_bf16_rows(&th, hidden, n_rows, 1, hidden_size, h_row_stride, h_row_stride, head::kBM))) return rc;
  if ((rc = make_tmap_bf16_rows(&tw, weight, n_cols, 1, hidden_size, w_row_stride, w_row_stride, head::kBN / 2))) return rc;
  head::Params p;
  p.stats = reinterpret_cast<float4*>(stats);
  p.n_rows = n_rows;
  p.n_cols = n_cols;
  p.k_chunks = hidden_size / head::kBK;
  p.n_row_pairs = n_row_pairs;
  p.n_tiles = n_tiles;
  p.tiles_per_split = (n_tiles + n_split - 1) / n_split;
  p.inv_temp = inv_temperature;
  p.logz = logz;
  p.entropy = entropy;
  p.g_entropy = g_entropy;
  p.g_rowsum = g_rowsum;
  p.dlogits = static_cast<__nv_bfloat16*>(dlogits);
  p.d_stride = d_row_stride;
  static bool configured = false;
  const int smem_bytes = head::Smem::total + 1024;
  if (!configured) {
    DTA_CHECK_CUDA(cudaFuncSetAttribute(head::lm_head_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    DTA_CHECK_CUDA(cudaFuncSetAttribute(head::lm_head_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    DTA_CHECK_CUDA(cudaFuncSetAttribute(head::lm_head_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = true;
  }
  const ProfileHook hook = take_profile_hook();
  if (hook.start) DTA_CHECK_CUDA(cudaEventRecord(hook.start, stream));
  const int grid = 2 * n_row_pairs * n_split;
  if (mode == 1)
    head::lm_head_kernel<1><<<grid, head::kThreads, smem_bytes, stream>>>(th, tw, p);
  else if (mode == 2)
    head::lm_head_kernel<2><<<grid, head::kThreads, smem_bytes, stream>>>(th, tw, p);
  else
    head::lm_head_kernel<0><<<grid, head::kThreads, smem_bytes, stream>>>(th, tw, p);
  DTA_CHECK_CUDA(cudaGetLastError());
  if (hook.stop) DTA_CHECK_CUDA(cudaEventRecord(hook.stop, stream));
  return DTA_OK;
}

int dta_lm_head_stats(const void* hidden, int64_t h_row_stride, const void* weight, int64_t w_row_stride, int64_t n_rows,
                      int64_t n_cols, int hidden_size, float inv_temperature, float* stats, dta_stream_t stream) {
  return launch_lm_head(0, hidden, h_row_stride, weight, w_row_stride, n_rows, n_cols, hidden_size, inv_temperature, stats,
                        nullptr, nullptr, nullptr, nullptr, nullptr, 0, stream);
}

int dta_lm_head_stats_logits(const void* hidden, int64_t h_row_stride, const void* weight, int64_t w_row_stride, int64_t n_rows,
                             int64_t n_cols, int hidden_size, float inv_temperature, float* stats, void* logits,
                             int64_t l_row_stride, dta_stream_t stream) {
  return launch_lm_head(2, hidden, h_row_stride, weight, w_row_stride, n_rows, n_cols, hidden_size, inv_temperature, stats,
                        nullptr, nullptr, nullptr, nullptr, logits, l_row_stride, stream);
}

int dta_lm_head_dlogits(const void* hidden, int64_t h_row_stride, const void* weight, int64_t w_row_stride, int64_t n_rows,
                        int64_t n_cols, int hidden_size, float inv_temperature, const float* logz, const float* entropy,
                        const float* g_entropy, const float* g_lp_rowsum, void* dlogits, int64_t d_row_stride,
                        dta_stream_t stream) {
  return launch_lm_head(1, hidden, h_row_stride, weight, w_row_stride, n_rows, n_cols, hidden_size, inv_temperature, nullptr,
                        logz, entropy, g_entropy, g_lp_rowsum, dlogits, d_row_stride, stream);
}

int dta_lm_head_label_logits(const void* hidden, int64_t h_row_stride, const void* weight, int64_t w_row_stride,
                             int hidden_size, int64_t n_cols, const int64_t* labels, const int32_t* row_of,
                             int64_t n_labels, float inv_temperature, float* x_label, dta_stream_t stream) {
  if (n_labels == 0) return DTA_OK;
  DTA_REQUIRE(hidden && weight && labels && x_label, "NULL pointer argument");
  DTA_REQUIRE(hidden_size % 8 == 0 && h_row_stride % 8 == 0 && w_row_stride % 8 == 0 &&
                  ((reinterpret_cast<uintptr_t>(hidden) | reinterpret_cast<uintptr_t>(weight)) & 15) == 0,
              "hidden / weight rows must be 16-byte aligned");
  int64_t grid = (n_labels * 32 + 255) / 256;
  const int64_t cap = 16 * static_cast<int64_t>(sm_count());
  if (grid > cap) grid = cap;
  head::lm_head_label_kernel<<<static_cast<int>(grid), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(hidden), h_row_stride, static_cast<const __nv_bfloat16*>(weight), w_row_stride,
      hidden_size, n_cols, labels, row_of, n_labels, inv_temperature, x_label);
  DTA_CHECK_CUDA(cudaGetLastError());
  return DTA_OK;
}

}  // extern "C"
