// attn_fwd.cuh — constants and launch parameters shared by the two forward kernels (attn_fwd.cu: one CTA per work item;
// attn_fwd_persist.cu: one CTA per SM walking the same items).
#pragma once
#include "dta_common.cuh"
#include "dta_ptx.cuh"
#include "../../include/dta_b200.h"

namespace dta {

static constexpr int kBM = 128;       // queries per tile
static constexpr int kBN = 128;       // keys per tile
static constexpr int kD = 128;        // head dim
static constexpr int kHalfBytes = kBM * 128;          // one [128 rows][64 bf16] swizzled half-tile
static constexpr int kTileBytes = 2 * kHalfBytes;     // 32 KB
static constexpr int kStages = 2;
static constexpr int kFwdThreads = 384;
static constexpr int kProducerWarp = 10, kMmaWarp = 11;
static constexpr int kMaxLive = 4096;
static constexpr int kPersistMaxTiles = 2048;     // attn_fwd_persist.cu: entries of one item slot's live-tile list
static constexpr float kLog2e = 1.4426950408889634f;
static constexpr float kLn2 = 0.6931471805599453f;
static constexpr float kRescaleThreshold = 8.0f;      // log2 units
static constexpr int kDefaultFwdPoly = 2;             // of every 8 exponentials, on the FMA pipes

struct FwdSmem {
  // offsets inside the 1024-aligned dynamic shared buffer
  static constexpr int q = 0;                                  // 2 tiles
  static constexpr int k = q + 2 * kTileBytes;                 // kStages tiles
  static constexpr int v = k + kStages * kTileBytes;           // kStages tiles
  static constexpr int live = v + kStages * kTileBytes;        // int32[kMaxLive]
  static constexpr int bars = live + kMaxLive * 4;             // mbarriers
  static constexpr int n_bars = 1 + 4 * kStages + 6;
  static constexpr int misc = bars + n_bars * 8;               // tmem addr, n_live
  static constexpr int total = misc + 16;
};

struct FwdParams {
  __nv_bfloat16* out;
  float* lse;
  const int32_t* subtree_end;
  const int32_t* kmax_end;
  const int32_t* kmin_end;
  const int32_t* q_order;
  int n_nodes, n_pad, n_tiles;
  int n_q_heads, n_kv_heads, group, pairs_per_kv;
  int64_t o_row_stride, o_head_stride;
  float scale_log2;
};

// attn_fwd_persist.cu: the persistent launch (same arguments the one-item-per-CTA kernel takes)
int launch_tree_attn_fwd_persist(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const FwdParams& p,
                                 int poly_of_8, cudaStream_t stream, const ProfileHook& hook);

}  // namespace dta
