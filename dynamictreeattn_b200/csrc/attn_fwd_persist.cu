// attn_fwd_persist.cu — the DEFAULT tree attention forward: a persistent kernel, one CTA per SM walking the work items
// of attn_fwd.cu's one-item-per-CTA kernel (which stays for tries of more than kPersistMaxTiles tiles and as the
// bit-equality reference, DTA_FWD_PERSIST=0).  Same mathematics, tiles, TMEM plan, warp roles and softmax -- read that
// file's header first.  Why: a (query tile, head pair) item of a bushy trie sees one to ten key tiles, and the other
// kernel then spends most of its time in CTA prologue and epilogue (fan(0, 64): 1.5 key tiles per item, 177 TFLOP/s;
// 318 here); on the 67k-token tau2 trie the hidden prologue / epilogue and the shared-memory epilogue are worth 13 %
// alone and 6 % inside the step (profiles/r02/attn_sweep_config5_persist.jsonl, bench_ab9_*.json).  What changes:
//
//   grid         = one CTA per SM; the producer warp draws items from a global counter (atomicAdd), in the same
//                  heaviest-first, head-major order the other kernel is launched in.  An item is claimed when the
//                  producer has issued the last load of the one before it -- claiming earlier (to hide the atomic's
//                  round trip) makes a CTA that starts a long item sit on the next long one: -10 % on a 16k chain.
//   item ring    = four slots {query tile, head pair, live-tile list}: the producer compacts the live key tiles of
//                  item i+1 and loads its Q / first K / V while the softmax warps are still in the last tile and the
//                  epilogue of item i; the MMA warp issues QK^T(0) of item i+1 right behind the last PV of item i
//                  (the tensor pipe executes one thread's MMAs in issue order, which attn_fwd.cu already relies on
//                  between PV(i) and QK^T(i+1) of one tile).  Four slots, not two: with two, the compaction of item
//                  i+1 waits for the epilogue of item i-1, and on one-tile items that wait is the item period.
//                  A slot's list holds kPersistMaxTiles entries; larger tries take the other kernel.
//   barriers     = the K / V ring and the S / P barriers run on across items (phases from running counters);
//                  bar_q_empty (commit behind the item's last QK^T) lets the next Q in; bar_item_full / _empty guard
//                  the item slots (empty: one arrival per softmax warp).
//   mask         = the 128 subtree ends of a key tile travel with its K load (512 B bulk copy on the same barrier)
//                  into a 4-slot ring, so a boundary tile's ancestor test reads shared memory: on a bushy trie two of
//                  three tiles are boundary tiles and their first touch was an L2 / DRAM round trip in the chain.
//                  Slot g & 3 is rewritten behind k_empty of tile g - 2, i.e. after QK^T(g - 2) retired, which is after
//                  PV(g - 3) retired, which is after the softmax of tiles g - 3 and g - 4 finished.
//                  A diagonal tile whose subtrees all reach past it (kmin_end >= q1: the inside of a branch) needs
//                  only the causal half of the test: no loads, one compare per element.
//   epilogue     = O leaves through shared memory and a TMA store.  One thread per row writing its 256 bytes straight
//                  to global memory makes every store instruction touch 32 rows: 7,100 cycles per item on the timeline
//                  (tools/fwd_persist_trace.py) with the LSU queue backed up into the next item.  The staging tiles
//                  are the item's LAST key tile's stages -- V for head A, K for head B, dead once the item's last
//                  PV retired; their "empty" arrivals then come from the store warps (warps 8 / 9: wait for the 128
//                  rows, issue the two half-tile stores, wait for the reads, arrive) instead of the MMA warp's commit.
//   scan         = the live-tile compaction loads 128 tiles' bounds per round trip instead of 32.
#include "attn_fwd.cuh"
#include <atomic>
#include <cstdlib>
#include <mutex>

namespace dta {
namespace fwd_persist {

static constexpr int kEndSlots = 4;
static constexpr int kEndBytes = kBN * 4;
static constexpr int kItemSlots = 4;
// Two softmax threads per row (Roles / kSplit below).  MEASURED SLOWER and therefore off: tau2 1018 against 1128 TFLOP/s
// alone, 967 against 1070 inside the step (profiles/r02/bench_ab11_*.json, attn_bushy_fwd_split{0,1}.jsonl).  The two
// softmax warps a scheduler already hosts (heads A and B) keep its issue slots and the MUFU pipe busy between them;
// four half-length warps do the same work plus the exchange, at 104 registers instead of 224.  DTA_FWD_SPLIT=1: A/B knob.
static constexpr bool kDefaultSplit = false;
static constexpr int kSlotLive = kPersistMaxTiles;

struct Smem {
  static constexpr int q = 0;                                  // 2 tiles
  static constexpr int k = q + 2 * kTileBytes;                 // kStages tiles
  static constexpr int v = k + kStages * kTileBytes;           // kStages tiles
  static constexpr int ends = v + kStages * kTileBytes;        // kEndSlots x int32[128]
  static constexpr int live = ends + kEndSlots * kEndBytes;    // kItemSlots x uint16[kSlotLive]: (key tile << 1) | boundary
  static constexpr int qorder = live + kItemSlots * kSlotLive * 2;   // uint16[kSlotLive]: q_order, read once
  static constexpr int meta = qorder + kSlotLive * 2;          // kItemSlots x {query tile, head pair, n_live (-1: no more work), -}
  static constexpr int xch = meta + kItemSlots * 16;           // kSplit: row maxima (two step parities) and row sums, [t][h][row] floats
  static constexpr int bars = xch + 3 * 2048;
  static constexpr int n_bars = 2 + 4 * kStages + 8 + 2 * kItemSlots;
  static constexpr int misc = bars + n_bars * 8;               // tmem addr
  static constexpr int total = misc + 16;
};
static_assert(Smem::total + 1024 <= 232448, "persistent forward kernel shared memory exceeds 227 KB");
static_assert((kSlotLive << 1) <= 65536, "live-tile entries are 16 bit");

// explicit shared-space accesses: the buffers are carved out of one aligned generic pointer, from which ptxas cannot
// tell the address space (it emitted generic LD.E for the mask loads)
__device__ __forceinline__ int4 lds_v4(uint32_t addr) {
  int4 r;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
  return r;
}
__device__ __forceinline__ int lds_s32(uint32_t addr) {
  int r;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(r) : "r"(addr) : "memory");
  return r;
}
__device__ __forceinline__ int lds_u16(uint32_t addr) {
  int r;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(r) : "r"(addr) : "memory");
  return r;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float r;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr) : "memory");
  return r;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint4 w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(smem_src),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// kTrace: debug timeline (tools/fwd_persist_trace.py): CTA 0 stamps clock64 at the hand-overs of its first kTraceItems
// items.  A separate instantiation, so the production kernel carries none of it.
static constexpr int kTraceItems = 24, kTraceEvents = 24;
#define DTA_FTRACE(ev)                                                                          \
  do {                                                                                          \
    if (kTrace && blockIdx.x == 0 && ic < kTraceItems) trace[ic * kTraceEvents + (ev)] = clock64(); \
  } while (0)

// kSplit: TWO softmax threads per query row (16 softmax warps; warps w and w + 8 share 32 rows and take the key halves
// [0, 64) / [64, 128) of every tile).  One thread per row makes the softmax of a tile an ~1,860-cycle dependency chain
// of ONE warp per scheduler, and the per-head chain S -> softmax -> PV -> S is what leaves the tensor pipe ~60 % busy
// (attn_fwd.cu's header); halving the chain needs only the row maximum exchanged per step (4 bytes per thread through
// shared memory + a 64-thread named barrier) and the row sums once per item.  P halves land in disjoint TMEM columns;
// the half that overwrites the other's S columns does so behind the barrier, i.e. after the other half has loaded them.
template <bool kSplit>
struct Roles {
  static constexpr int softmax_warps = kSplit ? 16 : 8;
  static constexpr int store_warp0 = softmax_warps;        // two O store warps
  static constexpr int producer = softmax_warps + 2;
  static constexpr int mma = softmax_warps + 3;            // highest warp slot: highest arbiter priority
  static constexpr int threads = (softmax_warps + 4) * 32;
  static constexpr int row_threads = kSplit ? 256 : 128;   // arrivals per tile on the P / O-staged barriers
};

template <int kPolyOf8, bool kTrace, bool kSplit>
__global__ void __launch_bounds__(Roles<kSplit>::threads, 1)
tree_attn_fwd_persist_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                             const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o,
                             const FwdParams p, int* counters, int n_items, long long* trace) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + Smem::q;
  uint8_t* sK = smem + Smem::k;
  uint8_t* sV = smem + Smem::v;
  uint8_t* sEnds = smem + Smem::ends;
  uint16_t* sLive = reinterpret_cast<uint16_t*>(smem + Smem::live);
  uint16_t* sQOrder = reinterpret_cast<uint16_t*>(smem + Smem::qorder);
  volatile int* sMeta = reinterpret_cast<volatile int*>(smem + Smem::meta);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
  uint64_t* bar_q_full = bars;
  uint64_t* bar_q_empty = bars + 1;
  uint64_t* bar_k_full = bars + 2;
  uint64_t* bar_k_empty = bar_k_full + kStages;
  uint64_t* bar_v_full = bar_k_empty + kStages;
  uint64_t* bar_v_empty = bar_v_full + kStages;
  uint64_t* bar_s_full = bar_v_empty + kStages;   // [2] per tile
  uint64_t* bar_p_full = bar_s_full + 2;          // [2]
  uint64_t* bar_o_final = bar_p_full + 2;         // [2]
  uint64_t* bar_o_staged = bar_o_final + 2;       // [2] the tile's 128 rows of O are in the staging tile
  uint64_t* bar_item_full = bar_o_staged + 2;              // [kItemSlots]
  uint64_t* bar_item_empty = bar_item_full + kItemSlots;   // [kItemSlots]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + Smem::misc);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  using R = Roles<kSplit>;
  constexpr int kProducerWarp = R::producer, kMmaWarp = R::mma;      // shadow attn_fwd.cuh's (the 12-warp layout)

  // ---- setup ---------------------------------------------------------------------------
  if (warp == kProducerWarp) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_q);
      tma_prefetch_desc(&tmap_k);
      tma_prefetch_desc(&tmap_v);
      tma_prefetch_desc(&tmap_o);
    }
  } else if (warp == kMmaWarp) {
    if (lane == 0) {
      mbar_init(bar_q_full, 1);
      mbar_init(bar_q_empty, 1);
      for (int s = 0; s < kStages; ++s) {
        mbar_init(bar_k_full + s, 1);
        mbar_init(bar_k_empty + s, 1);
        mbar_init(bar_v_full + s, 1);
        mbar_init(bar_v_empty + s, 1);
      }
      for (int t = 0; t < 2; ++t) {
        mbar_init(bar_s_full + t, 1);
        mbar_init(bar_p_full + t, R::row_threads);
        mbar_init(bar_o_final + t, 1);
        mbar_init(bar_o_staged + t, R::row_threads);
      }
      for (int s = 0; s < kItemSlots; ++s) {
        mbar_init(bar_item_full + s, 1);
        mbar_init(bar_item_empty + s, R::softmax_warps + 2);   // softmax warps + 2 store warps
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(s_tmem, 512);
  } else if (warp < 8) {
    for (int i = threadIdx.x; i < p.n_tiles; i += 256) sQOrder[i] = static_cast<uint16_t>(p.q_order[i]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp == kProducerWarp) {
    // ================================ work fetch + TMA producer ==========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    int ic = 0;                  // items published by this CTA
    uint32_t g = 0;              // key tiles loaded by this CTA (K / V ring and subtree-end ring position)
    int item = 0;
    if (lane == 0) item = atomicAdd(counters, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    for (;;) {
      const int buf = ic & (kItemSlots - 1);
      mbar_wait_timed(bar_item_empty + buf, ((ic / kItemSlots) & 1) ^ 1);  // softmax warps are done with this slot
      if (lane == 0) DTA_FTRACE(0);
      if (item >= n_items) {
        if (lane == 0) {
          sMeta[buf * 4 + 2] = -1;
          mbar_arrive(bar_item_full + buf);
        }
        break;
      }
      const int qtile = sQOrder[item % p.n_tiles];
      const int pair = item / p.n_tiles;
      const int kvh = pair / p.pairs_per_kv;
      const int head_a = kvh * p.group + 2 * (pair % p.pairs_per_kv);
      const bool valid_b = (2 * (pair % p.pairs_per_kv) + 1) < p.group;
      const int q0 = qtile * kBM, q1 = q0 + kBM;
      uint16_t* live = sLive + buf * kSlotLive;

      // ordered compaction of the live key tiles of this query tile, 128 tiles per round trip
      int count = 0;
      for (int base = 0; base <= qtile; base += 128) {
        int kmx[4], kmn[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = base + u * 32 + lane;
          kmx[u] = j <= qtile ? __ldg(p.kmax_end + j) : 0;
          kmn[u] = j <= qtile ? __ldg(p.kmin_end + j) : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = base + u * 32 + lane;
          const bool is_live = j <= qtile && kmx[u] > q0;
          const bool partial = !(j < qtile && kmn[u] >= q1);
          const bool causal_only = j == qtile && kmn[u] >= q1;       // kmn[u] of the diagonal tile is loaded below
          const unsigned ballot = __ballot_sync(0xffffffffu, is_live);
          if (is_live) {
            const int pos = count + __popc(ballot & ((1u << lane) - 1u));
            if (pos < kSlotLive) live[pos] = static_cast<uint16_t>((j << 2) | (causal_only ? 2 : 0) | (partial ? 1 : 0));
          }
          count += __popc(ballot);
        }
      }
      __syncwarp();
      if (lane == 0) {
        sMeta[buf * 4 + 0] = qtile;
        sMeta[buf * 4 + 1] = pair;
        sMeta[buf * 4 + 2] = count;
        mbar_arrive(bar_item_full + buf);
        DTA_FTRACE(1);
        if (kTrace && blockIdx.x == 0 && ic < kTraceItems) trace[ic * kTraceEvents + 19] = count;

        mbar_wait_timed(bar_q_empty, (ic & 1) ^ 1);                         // the previous item's last QK^T retired
        DTA_FTRACE(2);
        mbar_arrive_expect_tx(bar_q_full, (valid_b ? 2u : 1u) * kTileBytes);
        for (int t = 0; t < (valid_b ? 2 : 1); ++t)
          for (int h = 0; h < 2; ++h)
            tma_load_3d(sQ + t * kTileBytes + h * kHalfBytes, &tmap_q, bar_q_full, h * 64, head_a + t, q0);
        for (int it = 0; it < count; ++it, ++g) {
          const int j = live[it] >> 2;
          const int st = g % kStages;
          const uint32_t ph = (g / kStages) & 1;
          mbar_wait_timed(bar_k_empty + st, ph ^ 1);
          mbar_arrive_expect_tx(bar_k_full + st, kTileBytes + kEndBytes);
          bulk_load_1d(sEnds + (g & (kEndSlots - 1)) * kEndBytes, p.subtree_end + j * kBN, kEndBytes, bar_k_full + st);
          for (int h = 0; h < 2; ++h)
            tma_load_3d(sK + st * kTileBytes + h * kHalfBytes, &tmap_k, bar_k_full + st, h * 64, kvh, j * kBN);
          mbar_wait_timed(bar_v_empty + st, ph ^ 1);
          mbar_arrive_expect_tx(bar_v_full + st, kTileBytes);
          for (int h = 0; h < 2; ++h)
            tma_load_3d(sV + st * kTileBytes + h * kHalfBytes, &tmap_v, bar_v_full + st, h * 64, kvh, j * kBN);
        }
        DTA_FTRACE(3);
        item = atomicAdd(counters, 1);                                // claimed only now: see the header
        if (kTrace) asm volatile("" ::"r"(item) : "memory");          // the stamp below waits for the atomic's result
        DTA_FTRACE(4);
      }
      g = __shfl_sync(0xffffffffu, g, 0);
      item = __shfl_sync(0xffffffffu, item, 0);
      ++ic;
    }
    // the last CTA to run out of work re-arms the counters for the next launch that uses this slot
    if (lane == 0) {
      __threadfence();
      if (atomicAdd(counters + 1, 1) == static_cast<int>(gridDim.x) - 1) {
        counters[0] = 0;
        counters[1] = 0;
      }
    }
  } else if (warp == kMmaWarp) {
    // ================================ MMA issuer =========================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    constexpr uint32_t idesc_qk = make_idesc_bf16(kBM, kBN, false, false);
    constexpr uint32_t idesc_pv = make_idesc_bf16(kBM, kD, false, true);
    const uint32_t kQ = sdesc_lo_kmajor(smem_u32(sQ)), kK = sdesc_lo_kmajor(smem_u32(sK));
    const uint32_t nV = sdesc_lo_mnmajor(smem_u32(sV), kHalfBytes);
    constexpr uint32_t kTileW = kTileBytes >> 4;
    auto issue_qk = [&](int t, int st) {
      if (elect_one()) {
        const uint32_t a = kQ + t * kTileW, b = kK + st * kTileW, d = tmem + t * 128;
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk) {
          const uint32_t off = (kk >> 2) * (kHalfBytes >> 4) + (kk & 3) * 2;
          umma_ss_lo(d, a + off, b + off, idesc_qk, kk > 0 ? 1u : 0u);
        }
        umma_commit(bar_s_full + t);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int t, int st, bool first) {
      if (elect_one()) {
        const uint32_t b = nV + st * kTileW, a = tmem + t * 128, d = tmem + 256 + t * 128;
        const uint32_t acc0 = first ? 0u : 1u;
#pragma unroll
        for (int kk = 0; kk < kBN / 16; ++kk) umma_ts_lo(d, a + kk * 8, b + kk * (2048 >> 4), idesc_pv, kk > 0 ? 1u : acc0);
      }
      __syncwarp();
    };
    int ic = 0;
    uint32_t g = 0;              // key tiles consumed (ring position)
    uint32_t steps_b = 0;        // P hand-overs of tile B (it sits out the items of an odd group's last pair; tile A: g)
    for (;; ++ic) {
      const int buf = ic & (kItemSlots - 1);
      mbar_wait_timed(bar_item_full + buf, (ic / kItemSlots) & 1);
      const int n_live = lds_s32(smem_u32(smem + Smem::meta) + buf * 16 + 8);
      if (n_live < 0) break;
      if (lane == 0) DTA_FTRACE(5);
      const int pair = lds_s32(smem_u32(smem + Smem::meta) + buf * 16 + 4);
      const int n_t = ((2 * (pair % p.pairs_per_kv) + 1) < p.group) ? 2 : 1;
      mbar_wait_timed(bar_q_full, ic & 1);
      if (lane == 0) DTA_FTRACE(20);
      mbar_wait_timed(bar_k_full + g % kStages, (g / kStages) & 1);
      if (lane == 0) DTA_FTRACE(6);
      tc_fence_after();
      for (int t = 0; t < n_t; ++t) issue_qk(t, g % kStages);
      // The stages of the item's LAST tile become the O staging tiles (V: head A, K: head B) and are freed by the store
      // warps; every other stage is freed here, behind the MMAs that read it.
      if (elect_one()) {
        if (n_live > 1 || n_t == 1) umma_commit(bar_k_empty + g % kStages);
        if (n_live == 1) umma_commit(bar_q_empty);
      }
      __syncwarp();
      for (int it = 0; it < n_live; ++it, ++g) {
        const int st = g % kStages;
        const uint32_t ph = (g / kStages) & 1;
        const bool has_next = it + 1 < n_live;
        const int nst = (g + 1) % kStages;
        const uint32_t nph = ((g + 1) / kStages) & 1;
        mbar_wait_timed(bar_v_full + st, ph);
        if (has_next) mbar_wait_timed(bar_k_full + nst, nph);
        for (int t = 0; t < n_t; ++t) {
          mbar_wait_timed(bar_p_full + t, (t == 0 ? g : steps_b) & 1);
          if (lane == 0 && it == 0 && t == 0) DTA_FTRACE(7);
          tc_fence_after();
          issue_pv(t, st, it == 0);
          if (has_next) issue_qk(t, nst);
        }
        if (elect_one()) {
          if (has_next) {
            umma_commit(bar_v_empty + st);
            const bool last_qk = it + 2 == n_live;                   // that was the item's last QK^T
            if (!last_qk || n_t == 1) umma_commit(bar_k_empty + nst);
            if (last_qk) umma_commit(bar_q_empty);
          }
        }
        __syncwarp();
        if (n_t == 2) ++steps_b;
      }
      if (elect_one())
        for (int t = 0; t < n_t; ++t) umma_commit(bar_o_final + t);
      __syncwarp();
      if (lane == 0) DTA_FTRACE(8);
    }
  } else if (warp >= R::store_warp0) {
    // ================================ O store warps (head A, head B) ======================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    const int t = warp - R::store_warp0;
    uint32_t g = 0, done = 0;
    for (int ic = 0;; ++ic) {
      const int buf = ic & (kItemSlots - 1);
      mbar_wait_timed(bar_item_full + buf, (ic / kItemSlots) & 1);
      const uint32_t meta = smem_u32(smem + Smem::meta) + buf * 16;
      const int n_live = lds_s32(meta + 8);
      if (n_live < 0) break;
      const int qtile = lds_s32(meta), pair = lds_s32(meta + 4);
      const int head = (pair / p.pairs_per_kv) * p.group + 2 * (pair % p.pairs_per_kv) + t;
      const bool valid = t == 0 || (2 * (pair % p.pairs_per_kv) + 1) < p.group;
      if (valid) {
        const int st = (g + n_live - 1) % kStages;
        mbar_wait_timed(bar_o_staged + t, done & 1);
        ++done;
        if (lane == 0) {
          const uint32_t stage = smem_u32(t == 0 ? sV : sK) + st * kTileBytes;
          tma_store_3d(&tmap_o, stage, 0, head, qtile * kBM);
          tma_store_3d(&tmap_o, stage + kHalfBytes, 64, head, qtile * kBM);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          mbar_arrive(t == 0 ? bar_v_empty + st : bar_k_empty + st);
        }
      }
      g += static_cast<uint32_t>(n_live);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_item_empty + buf);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else {
    if constexpr (kSplit) {
    // ================================ softmax + epilogue, two threads per row ============
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");   // 640 threads start at 96; the four role warps release 4 x 32 x 40 = 5,120 registers, 16 x 32 x 8 = 4,096 are taken here
    const int t = (warp >> 2) & 1;                 // tile A / B
    const int h = warp >> 3;                       // key half of every tile: columns [64h, 64h + 64)
    const int row = ((warp & 3) << 5) | lane;      // TMEM lane == query row inside the tile
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) << 5) << 16;
    const uint32_t tS = tmem + lane_addr + t * 128;
    const uint32_t tO = tmem + lane_addr + 256 + t * 128 + h * 64;
    const uint32_t xch = smem_u32(smem + Smem::xch);
    const uint32_t x_mine = xch + ((t * 2 + h) * 128 + row) * 4, x_other = xch + ((t * 2 + (h ^ 1)) * 128 + row) * 4;
    const uint32_t pair_bar = 1 + t * 4 + (warp & 3);   // named barrier of the two warps that share these 32 rows
    uint32_t g = 0;                                // ring position of the item's first key tile
    uint32_t steps = 0, done = 0;                  // S tiles / items this warp has worked on (barrier phases)
    for (int ic = 0;; ++ic) {
      const int buf = ic & (kItemSlots - 1);
      mbar_wait_timed(bar_item_full + buf, (ic / kItemSlots) & 1);
      const uint32_t meta = smem_u32(smem + Smem::meta) + buf * 16;
      const int n_live = lds_s32(meta + 8);
      if (n_live < 0) break;
      const int qtile = lds_s32(meta);
      const int pair = lds_s32(meta + 4);
      const int kvh = pair / p.pairs_per_kv;
      const int head = kvh * p.group + 2 * (pair % p.pairs_per_kv) + t;
      const bool valid = t == 0 || (2 * (pair % p.pairs_per_kv) + 1) < p.group;
      const uint32_t live = smem_u32(sLive) + buf * (kSlotLive * 2);
      const bool tracer = kTrace && (warp & 3) == 0 && h == 0 && lane == 0;
      if (tracer) DTA_FTRACE(9 + 5 * t);
      if (valid) {
        const int q0 = qtile * kBM;
        const int q = q0 + row;
        float m2 = -INFINITY;   // running max of the WHOLE row (both halves keep the same value), log2 domain
        float l = 0.f;          // running sum of 2^(z - m2) over this thread's key half

        for (int it = 0; it < n_live; ++it, ++steps) {
          const int entry = lds_u16(live + it * 2);
          const int k0 = (entry >> 2) * kBN + h * 64;          // first key of this thread's half
          mbar_wait_timed(bar_s_full + t, steps & 1);
          if (tracer && it == 0) DTA_FTRACE(10 + 5 * t);
          tc_fence_after();
          float s[64];
          tmem_ld_x32(tS + h * 64, reinterpret_cast<uint32_t*>(s));
          tmem_ld_x32(tS + h * 64 + 32, reinterpret_cast<uint32_t*>(s) + 32);
          tmem_ld_wait();

          if (entry & 2) {  // diagonal tile inside a branch: causal
#pragma unroll
            for (int i = 0; i < 64; ++i)
              if (h * 64 + i > row) s[i] = -INFINITY;
          } else if (entry & 1) {  // boundary tile: per-element ancestor test against the tile's subtree ends
            const uint32_t ends = smem_u32(sEnds) + ((g + it) & (kEndSlots - 1)) * kEndBytes + h * 256;
#pragma unroll
            for (int c4 = 0; c4 < 16; ++c4) {
              const int4 e = lds_v4(ends + c4 * 16);
              const int kb = k0 + c4 * 4;
              if (!(kb + 0 <= q && q < e.x)) s[c4 * 4 + 0] = -INFINITY;
              if (!(kb + 1 <= q && q < e.y)) s[c4 * 4 + 1] = -INFINITY;
              if (!(kb + 2 <= q && q < e.z)) s[c4 * 4 + 2] = -INFINITY;
              if (!(kb + 3 <= q && q < e.w)) s[c4 * 4 + 3] = -INFINITY;
            }
          }
          float mx = s[0];
#pragma unroll
          for (int c = 1; c < 64; ++c) mx = fmaxf(mx, s[c]);
          // the row maximum over both halves.  The barrier also orders this half's S loads before the other half's
          // P stores into the same TMEM columns.
          const uint32_t par = (steps & 1) * 2048;
          sts_f32(x_mine + par, mx);
          named_bar_sync(pair_bar, 64);
          mx = fmaxf(mx, lds_f32(x_other + par));
          mx *= p.scale_log2;

          float alpha = 1.f;
          if (mx > m2 + kRescaleThreshold) {
            alpha = fast_exp2(m2 - mx);  // m2 == -inf -> 0
            m2 = mx;
          }
          if (it > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {   // same vote in both warps of the pair
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {          // rare path (lazy rescale): 16 columns at a time, rolled
              uint32_t o[16];
              tmem_ld_x16(tO + c * 16, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_x16(tO + c * 16, o);
            }
            tmem_st_wait();
          }
          l *= alpha;
          const float mref = (m2 == -INFINITY) ? 0.f : m2;
          const uint64_t scale2 = f2_bcast(p.scale_log2), nmref2 = f2_bcast(-mref);
          uint64_t lsum_a = f2_bcast(0.f), lsum_b = f2_bcast(0.f);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint64_t x2 = f2_fma(f2_pack(s[c * 16 + 2 * i], s[c * 16 + 2 * i + 1]), scale2, nmref2);
              uint64_t p2;
              if (exp2_pair_on_fma<kPolyOf8>(i)) {
                p2 = poly_exp2_x2(x2);
              } else {
                float x0, x1;
                f2_unpack(x2, x0, x1);
                p2 = f2_pack(fast_exp2(x0), fast_exp2(x1));
              }
              if (i & 1) lsum_b = f2_add(lsum_b, p2); else lsum_a = f2_add(lsum_a, p2);
              float p0, p1;
              f2_unpack(p2, p0, p1);
              pk[i] = pack_bf16x2(p0, p1);
            }
            tmem_st_x8(tS + h * 32 + c * 8, pk);
          }
          {
            float la, lb;
            f2_unpack(f2_add(lsum_a, lsum_b), la, lb);
            l += la + lb;
          }
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(bar_p_full + t);
          if (tracer && it + 1 == n_live) DTA_FTRACE(11 + 5 * t);
        }

        // ---- epilogue: O / l -> bf16, LSE -----------------------------------------------
        mbar_wait_timed(bar_o_final + t, done & 1);
        if (tracer) DTA_FTRACE(12 + 5 * t);
        ++done;
        tc_fence_after();
        sts_f32(x_mine + 4096, l);
        named_bar_sync(pair_bar, 64);
        const float l_other = lds_f32(x_other + 4096);
        const float lt = h == 0 ? l + l_other : l_other + l;        // the same sum, in the same order, in both threads
        const float inv_l = lt > 0.f ? 1.f / lt : 0.f;
        if (h == 0 && q < p.n_pad) {
          float lse = 0.f;
          if (q < p.n_nodes) lse = m2 + log2f(lt);     // base-2 log-sum-exp of the scaled scores
          p.lse[static_cast<int64_t>(head) * p.n_pad + q] = lse;
        }
        // this thread's 64 O columns into its half of the staging tile
        const uint32_t stage = smem_u32(t == 0 ? sV : sK) + ((g + n_live - 1) % kStages) * kTileBytes + h * kHalfBytes;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t o[32];
          tmem_ld_x32(tO + c * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(o[i + 0]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
            w.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
            w.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
            w.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
            sts_v4(stage + sw128_offset(row, c * 4 + (i >> 3)), w);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(bar_o_staged + t);
        if (tracer) DTA_FTRACE(13 + 5 * t);
      }
      g += static_cast<uint32_t>(n_live);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_item_empty + buf);
    }
    } else {
    // ================================ softmax + epilogue =================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int t = warp >> 2;                       // tile A / B
    const int row = ((warp & 3) << 5) | lane;      // TMEM lane == query row inside the tile
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) << 5) << 16;
    const uint32_t tS = tmem + lane_addr + t * 128;
    const uint32_t tO = tmem + lane_addr + 256 + t * 128;
    uint32_t g = 0;                                // ring position of the item's first key tile
    uint32_t steps = 0, done = 0;                  // S tiles / items this warp has worked on (barrier phases)
    for (int ic = 0;; ++ic) {
      const int buf = ic & (kItemSlots - 1);
      mbar_wait_timed(bar_item_full + buf, (ic / kItemSlots) & 1);
      const uint32_t meta = smem_u32(smem + Smem::meta) + buf * 16;
      const int n_live = lds_s32(meta + 8);
      if (n_live < 0) break;
      const int qtile = lds_s32(meta);
      const int pair = lds_s32(meta + 4);
      const int kvh = pair / p.pairs_per_kv;
      const int head = kvh * p.group + 2 * (pair % p.pairs_per_kv) + t;
      const bool valid = t == 0 || (2 * (pair % p.pairs_per_kv) + 1) < p.group;
      const uint32_t live = smem_u32(sLive) + buf * (kSlotLive * 2);
      const bool tracer = kTrace && (warp & 3) == 0 && lane == 0;
      if (tracer) DTA_FTRACE(9 + 5 * t);
      if (valid) {
        const int q0 = qtile * kBM;
        const int q = q0 + row;
        float m2 = -INFINITY;   // running max, log2 domain (already multiplied by scale_log2)
        float l = 0.f;          // running sum of 2^(z - m2)

        for (int it = 0; it < n_live; ++it, ++steps) {
          const int entry = lds_u16(live + it * 2);
          const int k0 = (entry >> 2) * kBN;
          mbar_wait_timed(bar_s_full + t, steps & 1);
          if (tracer && it == 0) DTA_FTRACE(10 + 5 * t);
          tc_fence_after();
          float s[kBN];
#pragma unroll
          for (int c = 0; c < 4; ++c) tmem_ld_x32(tS + c * 32, reinterpret_cast<uint32_t*>(s) + c * 32);
          tmem_ld_wait();

          if (entry & 2) {  // diagonal tile inside a branch: every subtree reaches past the tile, the test is causal
#pragma unroll
            for (int i = 1; i < kBN; ++i)
              if (i > row) s[i] = -INFINITY;
          } else if (entry & 1) {  // boundary tile: per-element ancestor test against the tile's subtree ends
            const uint32_t ends = smem_u32(sEnds) + ((g + it) & (kEndSlots - 1)) * kEndBytes;
#pragma unroll
            for (int c4 = 0; c4 < kBN / 4; ++c4) {
              const int4 e = lds_v4(ends + c4 * 16);
              const int kb = k0 + c4 * 4;
              if (!(kb + 0 <= q && q < e.x)) s[c4 * 4 + 0] = -INFINITY;
              if (!(kb + 1 <= q && q < e.y)) s[c4 * 4 + 1] = -INFINITY;
              if (!(kb + 2 <= q && q < e.z)) s[c4 * 4 + 2] = -INFINITY;
              if (!(kb + 3 <= q && q < e.w)) s[c4 * 4 + 3] = -INFINITY;
            }
          }
          float mx = s[0];
#pragma unroll
          for (int c = 1; c < kBN; ++c) mx = fmaxf(mx, s[c]);
          mx *= p.scale_log2;

          // lazy rescale: keep the stale max unless the new one is > 2^8 larger
          float alpha = 1.f;
          if (mx > m2 + kRescaleThreshold) {
            alpha = fast_exp2(m2 - mx);  // m2 == -inf -> 0
            m2 = mx;
          }
          if (it > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
            // O_t is stable: s_full of this step was committed after PV_t of the previous one
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t o[32];
              tmem_ld_x32(tO + c * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_x32(tO + c * 32, o);
            }
            tmem_st_wait();
          }
          l *= alpha;
          const float mref = (m2 == -INFINITY) ? 0.f : m2;
          const uint64_t scale2 = f2_bcast(p.scale_log2), nmref2 = f2_bcast(-mref);
          uint64_t lsum_a = f2_bcast(0.f), lsum_b = f2_bcast(0.f);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint64_t x2 = f2_fma(f2_pack(s[c * 32 + 2 * i], s[c * 32 + 2 * i + 1]), scale2, nmref2);
              uint64_t p2;
              if (exp2_pair_on_fma<kPolyOf8>(i)) {
                p2 = poly_exp2_x2(x2);
              } else {
                float x0, x1;
                f2_unpack(x2, x0, x1);
                p2 = f2_pack(fast_exp2(x0), fast_exp2(x1));
              }
              if (i & 1) lsum_b = f2_add(lsum_b, p2); else lsum_a = f2_add(lsum_a, p2);
              float p0, p1;
              f2_unpack(p2, p0, p1);
              pk[i] = pack_bf16x2(p0, p1);
            }
            tmem_st_x16(tS + c * 16, pk);
          }
          {
            float la, lb;
            f2_unpack(f2_add(lsum_a, lsum_b), la, lb);
            l += la + lb;
          }
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(bar_p_full + t);
          if (tracer && it + 1 == n_live) DTA_FTRACE(11 + 5 * t);
        }

        // ---- epilogue: O / l -> bf16, LSE -----------------------------------------------
        mbar_wait_timed(bar_o_final + t, done & 1);
        if (tracer) DTA_FTRACE(12 + 5 * t);
        ++done;
        tc_fence_after();
        const float inv_l = l > 0.f ? 1.f / l : 0.f;
        if (q < p.n_pad) {
          float lse = 0.f;
          if (q < p.n_nodes) lse = m2 + log2f(l);      // base-2 log-sum-exp of the scaled scores
          p.lse[static_cast<int64_t>(head) * p.n_pad + q] = lse;
        }
        // O rows into the staging tile ([128 rows][64 bf16] halves, 128-byte swizzle: what the TMA store expects)
        const uint32_t stage = smem_u32(t == 0 ? sV : sK) + ((g + n_live - 1) % kStages) * kTileBytes;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t o[32];
          tmem_ld_x32(tO + c * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(o[i + 0]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
            w.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
            w.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
            w.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
            const int col = c * 32 + i;
            sts_v4(stage + (col >> 6) * kHalfBytes + sw128_offset(row, (col & 63) >> 3), w);
          }
        }
        fence_proxy_async_smem();
        // the O columns are read: order the TMEM loads before the arrivals that let PV(0) of the next item overwrite them
        tc_fence_before();
        mbar_arrive(bar_o_staged + t);
        if (tracer) DTA_FTRACE(13 + 5 * t);
      }
      g += static_cast<uint32_t>(n_live);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_item_empty + buf);
    }
    }
  }

  // ---- teardown --------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// Work counters {next item, CTAs out of work}: a small pool per device, one slot per launch in rotation, re-armed by
// the launch's last CTA, so launches on different streams do not share a slot unless kSlots of them are in flight.
static constexpr int kSlots = 64;
static constexpr int kSlotInts = 32;    // 128 bytes apart
static constexpr int kMaxDevices = 64;

static int counter_slot(int** out) {
  static int* pool[kMaxDevices] = {};
  static std::atomic<unsigned> turn{0};
  static std::mutex mu;                  // first call per device from two host threads at once
  int dev = 0;
  DTA_CHECK_CUDA(cudaGetDevice(&dev));
  DTA_REQUIRE(dev >= 0 && dev < kMaxDevices, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lock(mu);
  if (!pool[dev]) {
    int* ptr = nullptr;
    DTA_CHECK_CUDA(cudaMalloc(&ptr, kSlots * kSlotInts * sizeof(int)));
    DTA_CHECK_CUDA(cudaMemset(ptr, 0, kSlots * kSlotInts * sizeof(int)));
    pool[dev] = ptr;
  }
  *out = pool[dev] + (turn.fetch_add(1) % kSlots) * kSlotInts;
  return DTA_OK;
}

}  // namespace fwd_persist

int launch_tree_attn_fwd_persist(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const FwdParams& p,
                                 int poly_of_8, cudaStream_t stream, const ProfileHook& hook) {
  using namespace fwd_persist;
  using Kernel = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const FwdParams, int*,
                          int, long long*);
  const int smem_bytes = Smem::total + 1024;
  static Kernel kernel = nullptr;
  static long long* trace = nullptr;
  static int threads = 0;
  if (!kernel) {
    // DTA_FWD_PERSIST_TRACE = device address (decimal) of kTraceItems x kTraceEvents int64: tools/fwd_persist_trace.py
    const char* te = getenv("DTA_FWD_PERSIST_TRACE");
    if (te && *te) trace = reinterpret_cast<long long*>(strtoull(te, nullptr, 10));
    const char* se = getenv("DTA_FWD_SPLIT");                // A/B knob: two softmax threads per row (kSplit)
    const bool split = se ? atoi(se) != 0 : kDefaultSplit;
    Kernel k;
    if (split)
      k = trace ? tree_attn_fwd_persist_kernel<kDefaultFwdPoly, true, true>
        : poly_of_8 <= 1 ? tree_attn_fwd_persist_kernel<0, false, true> : poly_of_8 <= 3 ? tree_attn_fwd_persist_kernel<2, false, true>
        : poly_of_8 <= 5 ? tree_attn_fwd_persist_kernel<4, false, true> : tree_attn_fwd_persist_kernel<6, false, true>;
    else
      k = trace ? tree_attn_fwd_persist_kernel<kDefaultFwdPoly, true, false>
        : poly_of_8 <= 1 ? tree_attn_fwd_persist_kernel<0, false, false> : poly_of_8 <= 3 ? tree_attn_fwd_persist_kernel<2, false, false>
        : poly_of_8 <= 5 ? tree_attn_fwd_persist_kernel<4, false, false> : tree_attn_fwd_persist_kernel<6, false, false>;
    DTA_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    threads = split ? Roles<true>::threads : Roles<false>::threads;
    kernel = k;
  }
  int* counters = nullptr;
  int rc;
  if ((rc = counter_slot(&counters))) return rc;
  CUtensorMap to;    // rows past n_nodes are clipped by the store
  if ((rc = make_tmap_bf16_rows(&to, p.out, p.n_nodes, p.n_q_heads, kD, p.o_row_stride, p.o_head_stride, kBM))) return rc;
  const int n_items = p.n_tiles * p.n_kv_heads * p.pairs_per_kv;
  const int grid = n_items < sm_count() ? n_items : sm_count();
  if (hook.start) DTA_CHECK_CUDA(cudaEventRecord(hook.start, stream));
  kernel<<<grid, threads, smem_bytes, stream>>>(tq, tk, tv, to, p, counters, n_items, trace);
  DTA_CHECK_CUDA(cudaGetLastError());
  if (hook.stop) DTA_CHECK_CUDA(cudaEventRecord(hook.stop, stream));
  return DTA_OK;
}

}  // namespace dta
