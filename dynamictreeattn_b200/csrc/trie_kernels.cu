// trie_kernels.cu — integer kernels: adjacent LCP, DFS linearisation of (lens, lcp) into
// parent / depth / ancestor-interval arrays, root-to-leaf paths, attention tile metadata.
//
// All outputs are bit-exact with the reference's host Python (token_trie.py:6-30) and with
// the derived linearisation spec (validated against
// tree_training_engine.py:541-546, 588-594).  These kernels are HBM/latency-bound integer
// work: coalesced grid-stride loops, warp-shuffle scans, no tensor cores.
#include "dta_common.cuh"
#include <cstdlib>
#include "../../include/dta_b200.h"

namespace dta {

static constexpr int kMaxLevels = 20;  // sparse tables cover n_leaf < 2^20

__host__ __device__ inline int table_levels(int n_leaf) {
  int lv = 1;
  while ((1 << lv) < n_leaf) ++lv;
  return lv + 1;
}

// ------------------------------------------------------------------------------------
// Adjacent LCP: one CTA per pair, 128-bit loads when both sequences are 16-byte aligned.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lcp_adjacent_kernel(const int64_t* __restrict__ tokens,
                                                           const int64_t* __restrict__ seq_off, int n_pairs,
                                                           int32_t* __restrict__ lcp_out,
                                                           int32_t* __restrict__ unsorted_flag) {
  __shared__ int s_first;
  for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
    const int64_t a0 = seq_off[pair], a1 = seq_off[pair + 1], b1 = seq_off[pair + 2];
    const int64_t* a = tokens + a0;
    const int64_t* b = tokens + a1;
    const int la = static_cast<int>(a1 - a0), lb = static_cast<int>(b1 - a1);
    const int n = la < lb ? la : lb;
    if (threadIdx.x == 0) s_first = n;
    __syncthreads();
    // chunked so the scan stops at the first chunk holding a mismatch
    for (int base = 0; base < n; base += blockDim.x * 4) {
      int mine = n;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int i = base + u * blockDim.x + threadIdx.x;
        if (i < n && a[i] != b[i]) mine = mine < i ? mine : i;
      }
      // warp min, then one shared atomic per warp
      for (int o = 16; o > 0; o >>= 1) {
        int other = __shfl_xor_sync(0xffffffffu, mine, o);
        mine = mine < other ? mine : other;
      }
      if ((threadIdx.x & 31) == 0 && mine < n) atomicMin(&s_first, mine);
      __syncthreads();
      if (s_first < n) break;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const int l = s_first;
      lcp_out[pair] = l;
      if (l < n && a[l] > b[l]) atomicExch(unsorted_flag, 1);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------
// Pass 1 (one CTA): fork depths s[i], exclusive scan of new-token counts -> leaf_off,
// and the two sparse min-tables over s used by pass 2.
//   back[k][i] = min s[max(0,i-2^k+1) .. i]     fwd[k][i] = min s[i .. min(n-1,i+2^k-1)]
// ------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_inclusive_scan(int v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int up = __shfl_up_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) >= o) v += up;
  }
  return v;
}

__global__ void __launch_bounds__(1024) trie_scan_kernel(const int32_t* __restrict__ lens,
                                                         const int32_t* __restrict__ lcp, int n_leaf,
                                                         int32_t* __restrict__ leaf_off, int32_t* __restrict__ ws,
                                                         int levels) {
  int32_t* s = ws;                                  // back[0] == fwd[0] == s
  int32_t* back = ws;                               // [levels][n_leaf], level 0 aliases s
  int32_t* fwd = ws + static_cast<size_t>(levels) * n_leaf;  // [levels][n_leaf]
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();

  for (int base = 0; base < n_leaf; base += blockDim.x) {
    const int i = base + tid;
    int si = 0, cnt = 0;
    if (i < n_leaf) {
      si = i > 0 ? lcp[i - 1] : 0;
      cnt = lens[i] - si;
      s[i] = si;
      fwd[i] = si;
    }
    int incl = warp_inclusive_scan(cnt);
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int w = warp_sums[lane];
      w = warp_inclusive_scan(w);
      warp_sums[lane] = w;
    }
    __syncthreads();
    const int before = carry + (wid > 0 ? warp_sums[wid - 1] : 0);
    if (i < n_leaf) leaf_off[i] = before + incl - cnt;
    __syncthreads();
    if (tid == blockDim.x - 1) carry = before + incl;
    __syncthreads();
  }
  if (tid == 0) leaf_off[n_leaf] = carry;

  for (int k = 1; k < levels; ++k) {
    __syncthreads();
    const int half = 1 << (k - 1);
    const int32_t* bp = back + static_cast<size_t>(k - 1) * n_leaf;
    const int32_t* fp = fwd + static_cast<size_t>(k - 1) * n_leaf;
    int32_t* bk = back + static_cast<size_t>(k) * n_leaf;
    int32_t* fk = fwd + static_cast<size_t>(k) * n_leaf;
    for (int i = tid; i < n_leaf; i += blockDim.x) {
      int b = bp[i];
      if (i - half >= 0) { int o = bp[i - half]; b = b < o ? b : o; }
      bk[i] = b;
      int f = fp[i];
      if (i + half < n_leaf) { int o = fp[i + half]; f = f < o ? f : o; }
      fk[i] = f;
    }
  }
}

// largest j <= start with s[j] <= q   (s[0] == 0 guarantees existence for q >= 0)
__device__ __forceinline__ int last_leq(const int32_t* __restrict__ back, int n_leaf, int levels, int start, int q) {
  int pos = start;
  for (int k = levels - 1; k >= 0; --k) {
    const int step = 1 << k;
    if (pos - step + 1 >= 0 && back[static_cast<size_t>(k) * n_leaf + pos] > q) pos -= step;
  }
  return pos;
}
// smallest j >= start with s[j] <= q, or n_leaf
__device__ __forceinline__ int first_leq(const int32_t* __restrict__ fwd, int n_leaf, int levels, int start, int q) {
  int pos = start;
  for (int k = levels - 1; k >= 0; --k) {
    const int step = 1 << k;
    if (pos + step <= n_leaf && fwd[static_cast<size_t>(k) * n_leaf + pos] > q) pos += step;
  }
  return pos;
}
// largest i with off[i] <= x, over off[0..n] ascending (off[n] is the total)
template <typename T>
__device__ __forceinline__ int owner_of(const T* __restrict__ off, int n, T x) {
  int lo = 0, hi = n;  // answer in [lo, hi)
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= x) lo = mid; else hi = mid;
  }
  return lo;
}

// ------------------------------------------------------------------------------------
// Pass 2: one thread per node.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) trie_nodes_kernel(const int32_t* __restrict__ leaf_off,
                                                         const int32_t* __restrict__ ws, int n_leaf, int levels,
                                                         int n_nodes, int n_pad, int32_t* __restrict__ depth,
                                                         int32_t* __restrict__ parent,
                                                         int32_t* __restrict__ subtree_end,
                                                         int32_t* __restrict__ leaf_of) {
  const int32_t* s = ws;
  const int32_t* back = ws;
  const int32_t* fwd = ws + static_cast<size_t>(levels) * n_leaf;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < n_pad; n += gridDim.x * blockDim.x) {
    if (n >= n_nodes) {
      depth[n] = 0; parent[n] = -1; subtree_end[n] = 0; leaf_of[n] = -1;
      continue;
    }
    const int i = owner_of<int32_t>(leaf_off, n_leaf, n);
    const int off = leaf_off[i];
    const int d = s[i] + (n - off);
    int par;
    if (n > off) par = n - 1;
    else if (d == 0) par = -1;
    else {
      const int j = last_leq(back, n_leaf, levels, i - 1, d - 1);
      par = leaf_off[j] + (d - 1 - s[j]);
    }
    const int jn = first_leq(fwd, n_leaf, levels, i + 1, d);
    depth[n] = d;
    parent[n] = par;
    subtree_end[n] = jn >= n_leaf ? n_nodes : leaf_off[jn];
    leaf_of[n] = i;
  }
}

__global__ void __launch_bounds__(256) trie_paths_kernel(const int32_t* __restrict__ leaf_off,
                                                         const int32_t* __restrict__ ws, int n_leaf, int levels,
                                                         const int64_t* __restrict__ path_off, int64_t total,
                                                         int32_t* __restrict__ path) {
  const int32_t* s = ws;
  const int32_t* back = ws;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int i = owner_of<int64_t>(path_off, n_leaf, idx);
    const int d = static_cast<int>(idx - path_off[i]);
    int j = i;
    if (d < s[i]) j = last_leq(back, n_leaf, levels, i - 1, d);
    path[idx] = leaf_off[j] + (d - s[j]);
  }
}

// ------------------------------------------------------------------------------------
// Tile metadata
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tile_minmax_kernel(const int32_t* __restrict__ subtree_end,
                                                          const int32_t* __restrict__ depth, int n_nodes,
                                                          int n_tiles, int32_t* __restrict__ kmax_end,
                                                          int32_t* __restrict__ kmin_end,
                                                          unsigned long long* __restrict__ pair_count) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_tiles) return;
  int mx = 0, mn = 0x7fffffff;
  unsigned long long pairs = 0;
  for (int u = 0; u < DTA_TILE / 32; ++u) {
    const int n = warp * DTA_TILE + u * 32 + lane;
    if (n < n_nodes) {
      const int e = subtree_end[n];
      mx = mx > e ? mx : e;
      mn = mn < e ? mn : e;
      pairs += static_cast<unsigned long long>(depth[n]) + 1ull;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    int a = __shfl_xor_sync(0xffffffffu, mx, o); mx = mx > a ? mx : a;
    int b = __shfl_xor_sync(0xffffffffu, mn, o); mn = mn < b ? mn : b;
    pairs += __shfl_xor_sync(0xffffffffu, pairs, o);
  }
  if (lane == 0) {
    kmax_end[warp] = mx;
    kmin_end[warp] = mn;
    atomicAdd(pair_count, pairs);
  }
}

// One CTA: live counts + descending rank sort (stable on tile index).  n_tiles is small
// (N/128), the O(n_tiles^2) compare loop is a few hundred thousand integer ops.
__global__ void __launch_bounds__(1024) tile_order_kernel(const int32_t* __restrict__ kmax_end, int n_tiles,
                                                          int32_t* __restrict__ q_order,
                                                          int32_t* __restrict__ q_nlive,
                                                          int32_t* __restrict__ k_order,
                                                          int32_t* __restrict__ k_qlast, int k_bucket) {
  for (int i = threadIdx.x; i < n_tiles; i += blockDim.x) {
    const int q0 = i * DTA_TILE;
    int live = 0;
    for (int j = 0; j <= i; ++j) live += kmax_end[j] > q0 ? 1 : 0;
    q_nlive[i] = live;
    k_qlast[i] = (kmax_end[i] - 1) / DTA_TILE;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_tiles; i += blockDim.x) {
    const int wq = q_nlive[i], wk = (k_qlast[i] - i) / k_bucket;
    int rq = 0, rk = 0;
    for (int j = 0; j < n_tiles; ++j) {
      const int oq = q_nlive[j], ok = (k_qlast[j] - j) / k_bucket;
      rq += (oq > wq || (oq == wq && j < i)) ? 1 : 0;
      rk += (ok > wk || (ok == wk && j < i)) ? 1 : 0;
    }
    q_order[rq] = i;
    k_order[rk] = i;
  }
}

}  // namespace dta

// --------------------------------------------------------------------------------------
// C ABI
// --------------------------------------------------------------------------------------
using namespace dta;

extern "C" {

int dta_lcp_adjacent(const int64_t* tokens, const int64_t* seq_off, int n_seq, int32_t* lcp_out,
                     int32_t* unsorted_flag, dta_stream_t stream) {
  DTA_REQUIRE(n_seq >= 0, "n_seq must be >= 0");
  DTA_REQUIRE(unsorted_flag != nullptr, "unsorted_flag is NULL");
  DTA_CHECK_CUDA(cudaMemsetAsync(unsorted_flag, 0, sizeof(int32_t), stream));
  if (n_seq < 2) return DTA_OK;
  DTA_REQUIRE(tokens && seq_off && lcp_out, "NULL pointer argument");
  const int n_pairs = n_seq - 1;
  const int grid = n_pairs < 4 * sm_count() ? n_pairs : 4 * sm_count();
  lcp_adjacent_kernel<<<grid, 256, 0, stream>>>(tokens, seq_off, n_pairs, lcp_out, unsorted_flag);
  DTA_CHECK_CUDA(cudaGetLastError());
  return DTA_OK;
}

size_t dta_trie_workspace_bytes(int n_leaf) {
  if (n_leaf <= 0) return 16;
  return static_cast<size_t>(2) * table_levels(n_leaf) * n_leaf * sizeof(int32_t) + 16;
}

int dta_trie_linearize(const int32_t* lens, const int32_t* lcp, int n_leaf, int n_nodes, int n_pad,
                       int32_t* leaf_off, int32_t* depth, int32_t* parent, int32_t* subtree_end,
                       int32_t* leaf_of, void* workspace, dta_stream_t stream) {
  DTA_REQUIRE(n_leaf >= 1, "n_leaf must be >= 1 (got %d)", n_leaf);
  DTA_REQUIRE(n_leaf < (1 << (kMaxLevels - 1)), "n_leaf %d too large", n_leaf);
  DTA_REQUIRE(n_nodes >= 1 && n_pad >= n_nodes && n_pad % DTA_TILE == 0,
              "need 1 <= n_nodes <= n_pad, n_pad %% %d == 0 (got %d, %d)", DTA_TILE, n_nodes, n_pad);
  DTA_REQUIRE(lens && (lcp || n_leaf == 1) && leaf_off && depth && parent && subtree_end && leaf_of && workspace,
              "NULL pointer argument");
  const int levels = table_levels(n_leaf);
  trie_scan_kernel<<<1, 1024, 0, stream>>>(lens, lcp, n_leaf, leaf_off, static_cast<int32_t*>(workspace), levels);
  DTA_CHECK_CUDA(cudaGetLastError());
  int grid = (n_pad + 255) / 256;
  const int cap = 8 * sm_count();
  if (grid > cap) grid = cap;
  trie_nodes_kernel<<<grid, 256, 0, stream>>>(leaf_off, static_cast<const int32_t*>(workspace), n_leaf, levels,
                                              n_nodes, n_pad, depth, parent, subtree_end, leaf_of);
  DTA_CHECK_CUDA(cudaGetLastError());
  return DTA_OK;
}

int dta_trie_paths(const int32_t* lens, const int32_t* lcp, const int32_t* leaf_off, int n_leaf,
                   const int64_t* path_off, int64_t total, int32_t* path, const void* workspace,
                   dta_stream_t stream) {
  (void)lens; (void)lcp;
  DTA_REQUIRE(n_leaf >= 1 && total >= 0, "bad sizes");
  if (total == 0) return DTA_OK;
  DTA_REQUIRE(leaf_off && path_off && path && workspace, "NULL pointer argument");
  const int levels = table_levels(n_leaf);
  int64_t grid = (total + 255) / 256;
  const int cap = 8 * sm_count();
  if (grid > cap) grid = cap;
  trie_paths_kernel<<<static_cast<int>(grid), 256, 0, stream>>>(leaf_off, static_cast<const int32_t*>(workspace),
                                                                n_leaf, levels, path_off, total, path);
  DTA_CHECK_CUDA(cudaGetLastError());
  return DTA_OK;
}

int dta_tile_meta(const int32_t* subtree_end, const int32_t* depth, int n_nodes, int n_pad, int32_t* kmax_end,
                  int32_t* kmin_end, int32_t* q_order, int32_t* q_nlive, int32_t* k_order, int32_t* k_qlast,
                  int64_t* pair_count, dta_stream_t stream) {
  DTA_REQUIRE(n_nodes >= 1 && n_pad >= n_nodes && n_pad % DTA_TILE == 0 && n_pad - n_nodes < DTA_TILE,
              "n_pad must be n_nodes rounded up to %d", DTA_TILE);
  DTA_REQUIRE(subtree_end && depth && kmax_end && kmin_end && q_order && q_nlive && k_order && k_qlast && pair_count,
              "NULL pointer argument");
  const int n_tiles = n_pad / DTA_TILE;
  DTA_CHECK_CUDA(cudaMemsetAsync(pair_count, 0, sizeof(int64_t), stream));
  tile_minmax_kernel<<<(n_tiles * 32 + 255) / 256, 256, 0, stream>>>(
      subtree_end, depth, n_nodes, n_tiles, kmax_end, kmin_end, reinterpret_cast<unsigned long long*>(pair_count));
  DTA_CHECK_CUDA(cudaGetLastError());
  const char* kb = getenv("DTA_K_ORDER_BUCKET");      // tuning override of the weight-class width (tests use the default)
  const int k_bucket = kb && atoi(kb) > 0 ? atoi(kb) : DTA_K_ORDER_BUCKET;
  tile_order_kernel<<<1, 1024, 0, stream>>>(kmax_end, n_tiles, q_order, q_nlive, k_order, k_qlast, k_bucket);
  DTA_CHECK_CUDA(cudaGetLastError());
  return DTA_OK;
}

}  // extern "C"
