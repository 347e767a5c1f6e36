// trie_build.cu — device-side trie build: lexicographic STABLE sort of the rollouts,
// leafization and the forward / backward leaf orders, bit-exact with the reference's host code.
//
// Replaces, for CUDA-resident rollouts, the host work of the reference's TokenTrie:
//   token_trie.py:68-73   Python list.sort on .tolist() keys                     -> rank sort on pairwise compares
//   token_trie.py:6-10    _lcp_torch, one .item() sync per adjacent pair          -> the pairwise LCP matrix
//   token_trie.py:13-49   _leafization (merge full-prefix sequences into leaves)  -> flag + block scan
//   trie.py:93-133, 135-148, 171-188, 190-202, 225-233
//                         CompressedTrie + get_order_forward / get_order_backward -> trie_plan_kernel
//   token_trie.py:90-94   permute(): adjacent LCPs of the reordered leaves        -> LCP-matrix lookups
//
// Design.  S sequences of up to 16k tokens are a few MB: they sit in L2, so the natural GPU formulation is ALL-PAIRS:
// one warp per unordered pair (i, j) scans both sequences 32 tokens at a time until the first mismatch and records
// LCP(i, j) and the three-way comparison (shorter-prefix-first, signed ids — the order Python's list comparison
// gives).  The stable sort is then a rank sort: rank(i) = #{j : seq_j < seq_i or (seq_j == seq_i and j < i)}, which is
// exactly the permutation a stable sort produces.  Every later LCP the reference recomputes from tokens (adjacent
// pairs after sorting, after leafization, after each permute) is a lookup in the matrix.  The compressed trie and its
// two DFS orders are O(n_leaf log n_leaf) pointer work on <= 2 n_leaf nodes: one thread of one CTA walks it with the
// reference's exact rules (children sorted STABLY by chain_tail_depth, resp. (has_children, chain_tail_depth), the
// backward order reversed), while the rest of the CTA does the scans and lookups around it.
// All results land in ONE int32 buffer, so the host needs a single small D2H copy to size its allocations.
#include "dta_common.cuh"
#include "../../include/dta_b200.h"

namespace dta {

static constexpr int kPlanThreads = 1024;

// ------------------------------------------------------------------------------------
// all-pairs LCP + three-way compare: one warp per pair i < j
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) seq_pair_kernel(const int64_t* __restrict__ tokens, const int64_t* __restrict__ off,
                                                       int n_seq, int32_t* __restrict__ lcp, int8_t* __restrict__ cmp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x;
  const int j = blockIdx.y * 8 + warp;
  if (j >= n_seq) return;
  const int64_t len_i = off[i + 1] - off[i];
  if (j == i) {
    if (lane == 0) {
      lcp[static_cast<int64_t>(i) * n_seq + i] = static_cast<int32_t>(len_i);
      cmp[static_cast<int64_t>(i) * n_seq + i] = 0;
    }
    return;
  }
  if (j < i) return;
  const int64_t len_j = off[j + 1] - off[j];
  const int64_t n = len_i < len_j ? len_i : len_j;
  const int64_t* a = tokens + off[i];
  const int64_t* b = tokens + off[j];
  int64_t l = n;
  int c = len_i < len_j ? -1 : (len_i == len_j ? 0 : 1);          // one is a prefix of the other: shorter first
  for (int64_t pos = 0; pos < n; pos += 32) {
    const int64_t p = pos + lane;
    int64_t x = 0, y = 0;
    if (p < n) {
      x = a[p];
      y = b[p];
    }
    const unsigned diff = __ballot_sync(0xffffffffu, x != y);
    if (diff) {
      const int first = __ffs(diff) - 1;
      l = pos + first;
      const int64_t xf = __shfl_sync(0xffffffffu, x, first), yf = __shfl_sync(0xffffffffu, y, first);
      c = xf < yf ? -1 : 1;                                         // signed, like Python int comparison
      break;
    }
  }
  if (lane == 0) {
    lcp[static_cast<int64_t>(i) * n_seq + j] = static_cast<int32_t>(l);
    lcp[static_cast<int64_t>(j) * n_seq + i] = static_cast<int32_t>(l);
    cmp[static_cast<int64_t>(i) * n_seq + j] = static_cast<int8_t>(c);
    cmp[static_cast<int64_t>(j) * n_seq + i] = static_cast<int8_t>(-c);
  }
}

// rank sort: one warp per sequence; perm[rank] = original index (== stable sort)
__global__ void __launch_bounds__(256) seq_rank_kernel(const int8_t* __restrict__ cmp, int n_seq, int presorted,
                                                       int32_t* __restrict__ perm) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n_seq) return;
  if (presorted) {                                  // caller vouches for the order (sorted=True): identity
    if (lane == 0) perm[i] = i;
    return;
  }
  int r = 0;
  const int8_t* row = cmp + static_cast<int64_t>(i) * n_seq;       // row[j] = compare(seq_i, seq_j)
  for (int j = lane; j < n_seq; j += 32) {
    const int c = row[j];
    r += (c > 0 || (c == 0 && j < i)) ? 1 : 0;                     // seq_j sorts before seq_i
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  if (lane == 0) perm[r] = i;
}

// ------------------------------------------------------------------------------------
// plan: leafization + compressed trie + leaf orders (one CTA)
// ------------------------------------------------------------------------------------
struct PlanScratch {            // all int32[2 * n_seq + 2] unless noted
  int32_t *depth, *seq_id, *child_head, *child_tail, *next_sib, *tail, *stack_v, *stack_c, *kids;
};

struct PlanOut {                // views into the single output buffer (see dta_trie_build)
  int32_t *header, *perm, *slcp, *absorbed, *leaf_pos, *leaf_lcp, *order_fwd, *lcp_fwd, *order_bwd, *lcp_bwd;
};

__device__ __forceinline__ void node_append(const PlanScratch& s, int v, int c) {
  s.next_sib[c] = -1;
  if (s.child_tail[v] < 0) s.child_head[v] = c; else s.next_sib[s.child_tail[v]] = c;
  s.child_tail[v] = c;
}

// child ordering key; backward: leaves before forks, then by tail (reference trie.py:197-202)
__device__ __forceinline__ int64_t child_key(const PlanScratch& s, int c, bool backward) {
  const int64_t fork = (backward && s.child_head[c] >= 0) ? 1 : 0;
  return (fork << 32) | static_cast<uint32_t>(s.tail[c]);
}

// One DFS order of the leaves (forward: backward == false).  Sequential; mirrors trie.py:135-148 (tails, post-order,
// tail[v] = tail of the FIRST child under the active order) and trie.py:171-188 (pre-order emit with children sorted
// stably by key).
__device__ void plan_order(const PlanScratch& s, bool backward, int32_t* order, int n_leaf) {
  // ---- tails ----
  int sp = 0;
  s.stack_v[sp] = 0;
  s.stack_c[sp] = s.child_head[0];
  ++sp;
  while (sp > 0) {
    const int v = s.stack_v[sp - 1];
    if (s.seq_id[v] != -1) {
      s.tail[v] = s.depth[v];
      --sp;
      continue;
    }
    const int c = s.stack_c[sp - 1];
    if (c != -1) {
      s.stack_c[sp - 1] = s.next_sib[c];
      s.stack_v[sp] = c;
      s.stack_c[sp] = s.child_head[c];
      ++sp;
    } else {
      int best = s.child_head[v];
      for (int k = s.next_sib[best]; k != -1; k = s.next_sib[k])
        if (child_key(s, k, backward) < child_key(s, best, backward)) best = k;     // strict: first of equal keys wins
      s.tail[v] = s.tail[best];
      --sp;
    }
  }
  // ---- ordered pre-order walk ----
  int n_out = 0;
  sp = 0;
  s.stack_v[sp++] = 0;
  while (sp > 0) {
    const int v = s.stack_v[--sp];
    if (s.seq_id[v] != -1) {
      order[n_out++] = s.seq_id[v];
      continue;
    }
    int nk = 0;
    for (int k = s.child_head[v]; k != -1; k = s.next_sib[k]) {                     // stable insertion sort by key
      const int64_t key = child_key(s, k, backward);
      int p = nk++;
      while (p > 0 && child_key(s, s.kids[p - 1], backward) > key) {
        s.kids[p] = s.kids[p - 1];
        --p;
      }
      s.kids[p] = k;
    }
    for (int p = nk - 1; p >= 0; --p) s.stack_v[sp++] = s.kids[p];                  // first child on top
  }
  if (backward)                                                                     // "whole order reversed"
    for (int a = 0, b = n_leaf - 1; a < b; ++a, --b) {
      const int t = order[a];
      order[a] = order[b];
      order[b] = t;
    }
}

__global__ void __launch_bounds__(kPlanThreads) trie_plan_kernel(const int64_t* __restrict__ off, int n_seq,
                                                                 const int32_t* __restrict__ lcp,
                                                                 const int8_t* __restrict__ cmp, PlanScratch s,
                                                                 PlanOut o) {
  __shared__ int s_warp[32];
  __shared__ int s_total, s_unsorted, s_nleaf;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_unsorted = 0;
  __syncthreads();
  // ---- adjacent LCPs of the sorted order, absorbed flags, sortedness of the INPUT order -----------------------
  for (int r = tid; r < n_seq; r += kPlanThreads) {
    const int a = o.perm[r];
    int l = 0, ab = 0;
    if (r + 1 < n_seq) {
      const int b = o.perm[r + 1];
      l = lcp[static_cast<int64_t>(a) * n_seq + b];
      const int64_t la = off[a + 1] - off[a], lb = off[b + 1] - off[b];
      ab = l >= (la < lb ? la : lb) ? 1 : 0;                       // full prefix of its successor (token_trie.py:33)
      if (cmp[static_cast<int64_t>(r) * n_seq + r + 1] > 0) atomicAdd(&s_unsorted, 1);   // input pair (r, r+1) out of order
    }
    o.slcp[r] = l;
    o.absorbed[r] = ab;
  }
  __syncthreads();
  // ---- leaves = sorted sequences that are not absorbed: compaction by block scan ------------------------------
  int base = 0;
  for (int start = 0; start < n_seq; start += kPlanThreads) {
    const int r = start + tid;
    const int keep = (r < n_seq && !o.absorbed[r]) ? 1 : 0;
    int inc = keep;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const int w = s_warp[lane];
      int winc = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, winc, d);
        if (lane >= d) winc += t;
      }
      s_warp[lane] = winc - w;
      if (lane == 31) s_total = winc;
    }
    __syncthreads();
    if (keep) {
      const int j = base + s_warp[warp] + inc - 1;
      o.leaf_pos[j] = r;
      o.leaf_lcp[j] = o.slcp[r];                                   // LCP with the next sorted sequence (token_trie.py:41-42)
    }
    base += s_total;
    __syncthreads();
  }
  if (tid == 0) s_nleaf = base;
  __syncthreads();
  const int n_leaf = s_nleaf;
  // ---- compressed trie + the two leaf orders: one thread, the reference's rules verbatim -----------------------
  if (tid == 0) {
    int n_nodes = 0;
    auto add = [&](int d, int sid) {
      const int v = n_nodes++;
      s.depth[v] = d;
      s.seq_id[v] = sid;
      s.child_head[v] = s.child_tail[v] = -1;
      s.next_sib[v] = -1;
      s.tail[v] = 0;
      return v;
    };
    add(0, -1);
    int32_t* spine = s.stack_v;                                     // right spine (trie.py:93-133)
    int sp = 0;
    spine[sp++] = 0;
    for (int i = 0; i < n_leaf; ++i) {
      const int f = i > 0 ? o.leaf_lcp[i - 1] : 0;
      const int a = o.perm[o.leaf_pos[i]];
      const int length = static_cast<int>(off[a + 1] - off[a]);
      int closed = -1;
      while (sp > 1 && (s.depth[spine[sp - 1]] > f || s.seq_id[spine[sp - 1]] != -1)) {
        const int v = spine[--sp];
        if (closed >= 0) node_append(s, v, closed);
        closed = v;
      }
      if (s.depth[spine[sp - 1]] < f) spine[sp++] = add(f, -1);
      if (closed >= 0) node_append(s, spine[sp - 1], closed);
      spine[sp++] = add(length, i);
    }
    while (sp > 1) {
      const int v = spine[--sp];
      node_append(s, spine[sp - 1], v);
    }
    if (n_leaf > 0) {
      plan_order(s, false, o.order_fwd, n_leaf);
      plan_order(s, true, o.order_bwd, n_leaf);
    }
    o.header[0] = n_leaf;
    o.header[1] = s_unsorted;
  }
  __syncthreads();
  // ---- adjacent LCPs of the reordered leaves (TokenTrie.permute recomputes them from tokens, token_trie.py:94) ---
  for (int k = tid; k + 1 < n_leaf; k += kPlanThreads) {
    const int a = o.perm[o.leaf_pos[o.order_fwd[k]]], b = o.perm[o.leaf_pos[o.order_fwd[k + 1]]];
    o.lcp_fwd[k] = lcp[static_cast<int64_t>(a) * n_seq + b];
    const int c = o.perm[o.leaf_pos[o.order_bwd[k]]], d = o.perm[o.leaf_pos[o.order_bwd[k + 1]]];
    o.lcp_bwd[k] = lcp[static_cast<int64_t>(c) * n_seq + d];
  }
}

struct BuildWorkspace {
  size_t lcp, cmp, scratch, total;
};
static BuildWorkspace build_workspace_layout(int n_seq) {
  auto up = [](size_t x) { return (x + 255) & ~static_cast<size_t>(255); };
  BuildWorkspace w;
  const size_t n = static_cast<size_t>(n_seq);
  size_t o = 0;
  w.lcp = o;     o = up(o + n * n * sizeof(int32_t));
  w.cmp = o;     o = up(o + n * n);
  w.scratch = o; o = up(o + 9 * (2 * n + 2) * sizeof(int32_t));
  w.total = o;
  return w;
}

}  // namespace dta

using namespace dta;

extern "C" {

size_t dta_trie_build_workspace_bytes(int n_seq) { return n_seq < 1 ? 0 : build_workspace_layout(n_seq).total; }

int dta_trie_build(const int64_t* tokens, const int64_t* offsets, int n_seq, int presorted, void* workspace,
                   int32_t* out, dta_stream_t stream) {
  DTA_REQUIRE(n_seq >= 1 && n_seq <= DTA_TRIE_BUILD_MAX_SEQ, "n_seq must be in [1, %d] (got %d)", DTA_TRIE_BUILD_MAX_SEQ, n_seq);
  DTA_REQUIRE(tokens && offsets && workspace && out, "NULL pointer argument");
  DTA_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  const BuildWorkspace wl = build_workspace_layout(n_seq);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  int32_t* lcp = reinterpret_cast<int32_t*>(ws + wl.lcp);
  int8_t* cmp = reinterpret_cast<int8_t*>(ws + wl.cmp);
  int32_t* sc = reinterpret_cast<int32_t*>(ws + wl.scratch);
  const size_t m = 2 * static_cast<size_t>(n_seq) + 2;
  PlanScratch s{sc, sc + m, sc + 2 * m, sc + 3 * m, sc + 4 * m, sc + 5 * m, sc + 6 * m, sc + 7 * m, sc + 8 * m};
  const size_t n = static_cast<size_t>(n_seq);
  PlanOut o{out, out + 2, out + 2 + n, out + 2 + 2 * n, out + 2 + 3 * n, out + 2 + 4 * n,
            out + 2 + 5 * n, out + 2 + 6 * n, out + 2 + 7 * n, out + 2 + 8 * n};
  seq_pair_kernel<<<dim3(n_seq, (n_seq + 7) / 8), 256, 0, stream>>>(tokens, offsets, n_seq, lcp, cmp);
  DTA_CHECK_CUDA(cudaGetLastError());
  seq_rank_kernel<<<(n_seq + 7) / 8, 256, 0, stream>>>(cmp, n_seq, presorted, o.perm);
  DTA_CHECK_CUDA(cudaGetLastError());
  trie_plan_kernel<<<1, kPlanThreads, 0, stream>>>(offsets, n_seq, lcp, cmp, s, o);
  DTA_CHECK_CUDA(cudaGetLastError());
  return DTA_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------
// HOST-side planning helper for the partitioner (no device work).  LB_by_DFS_and_TM (reference data_parallel.py:58-107)
// probes a few hundred contiguous leaf ranges per call; every probe builds the range's compressed trie, orders its
// leaves (trie.py:190-202, 225-233) and evaluates _get_stats (trie.py:7-56).  In Python that is ~100 us per probe and the
// whole partition 20-40 ms of a ~800 ms step at 8 GPUs; here it is the same algorithm, verbatim, in C++ (a few us).
// Results are integers and must equal the Python code's bit for bit (tests/test_host_golden.py).
// ------------------------------------------------------------------------------------
#include <vector>

namespace dta {
namespace host_plan {

struct Trie {
  std::vector<int> depth, seq_id, head, tail_child, next, tail;
  int add(int d, int s) {
    depth.push_back(d); seq_id.push_back(s); head.push_back(-1); tail_child.push_back(-1); next.push_back(-1); tail.push_back(0);
    return static_cast<int>(depth.size()) - 1;
  }
  void append(int v, int c) {
    next[c] = -1;
    if (tail_child[v] < 0) head[v] = c; else next[tail_child[v]] = c;
    tail_child[v] = c;
  }
  long long key(int c, bool backward) const {
    const long long fork = (backward && head[c] >= 0) ? 1 : 0;
    return (fork << 32) | static_cast<unsigned>(tail[c]);
  }
};

static void build(Trie& t, const int32_t* lens, const int32_t* lcps, int n_leaf) {
  t.add(0, -1);
  std::vector<int> spine{0};
  for (int i = 0; i < n_leaf; ++i) {
    const int f = i > 0 ? lcps[i - 1] : 0;
    int closed = -1;
    while (spine.size() > 1 && (t.depth[spine.back()] > f || t.seq_id[spine.back()] != -1)) {
      const int v = spine.back();
      spine.pop_back();
      if (closed >= 0) t.append(v, closed);
      closed = v;
    }
    if (t.depth[spine.back()] < f) spine.push_back(t.add(f, -1));
    if (closed >= 0) t.append(spine.back(), closed);
    spine.push_back(t.add(lens[i], i));
  }
  while (spine.size() > 1) {
    const int v = spine.back();
    spine.pop_back();
    t.append(spine.back(), v);
  }
}

// leaves in DFS order with children sorted stably by key; lens / lcps as trie.py's _walk emits them (lcp = shallowest node
// depth crossed since the previous leaf)
static void order(Trie& t, bool backward, std::vector<int>& lens, std::vector<int>& lcps) {
  std::vector<std::pair<int, int>> st{{0, t.head[0]}};
  while (!st.empty()) {                      // tails, post-order (trie.py:135-148)
    const int v = st.back().first;
    if (t.seq_id[v] != -1) { t.tail[v] = t.depth[v]; st.pop_back(); continue; }
    const int c = st.back().second;
    if (c != -1) {
      st.back().second = t.next[c];
      st.push_back({c, t.head[c]});
    } else {
      int best = t.head[v];
      for (int k = t.next[best]; k != -1; k = t.next[k])
        if (t.key(k, backward) < t.key(best, backward)) best = k;
      t.tail[v] = t.tail[best];
      st.pop_back();
    }
  }
  int lca = 0;
  std::vector<std::pair<int, int>> todo{{0, -1}};   // (node, parent depth)
  std::vector<int> kids;
  while (!todo.empty()) {
    const int v = todo.back().first, pd = todo.back().second;
    todo.pop_back();
    if (pd >= 0 && pd < lca) lca = pd;
    if (t.seq_id[v] != -1) {
      lens.push_back(t.depth[v]);
      lcps.push_back(lca);
      lca = t.depth[v];
      continue;
    }
    kids.clear();
    for (int k = t.head[v]; k != -1; k = t.next[k]) {   // stable insertion sort by key
      kids.push_back(k);
      size_t p = kids.size() - 1;
      while (p > 0 && t.key(kids[p - 1], backward) > t.key(k, backward)) { kids[p] = kids[p - 1]; --p; }
      kids[p] = k;
    }
    for (size_t p = kids.size(); p-- > 0;) todo.push_back({kids[p], t.depth[v]});
  }
}

}  // namespace host_plan
}  // namespace dta

extern "C" int dta_host_subtrie_stats(const int32_t* lens, const int32_t* lcps, int n_leaf, int backward, int64_t block_size,
                                      int64_t* out5) {
  using namespace dta::host_plan;
  if (!lens || !out5 || n_leaf < 1 || (n_leaf > 1 && !lcps)) return 1;
  Trie t;
  build(t, lens, lcps, n_leaf);
  std::vector<int> L, C;
  order(t, backward != 0, L, C);
  // get_order_forward: (lens, lcp[1:]); get_order_backward: (lens[::-1], lcp[1:][::-1])   (trie.py:225-233)
  std::vector<int> lcp(C.begin() + 1, C.end());
  if (backward) {
    std::vector<int> Lr(L.rbegin(), L.rend()), Cr(lcp.rbegin(), lcp.rend());
    L.swap(Lr);
    lcp.swap(Cr);
  }
  // _get_stats (trie.py:7-56), with the last-leaf lcp read as 0 where the reference raises IndexError
  long long n_tree = 0, sum_depth = 0, sum_prefix = 0, n_f1 = 0;
  for (int i = 0; i < n_leaf; ++i) n_tree += L[i];
  for (size_t i = 0; i < lcp.size(); ++i) n_tree -= lcp[i];
  for (int i = 0; i < n_leaf; ++i) {
    const long long lo = i > 0 ? lcp[i - 1] : 0, hi = L[i];
    sum_depth += (lo + hi - 1) * (hi - lo) / 2;
  }
  if (!backward) {
    for (size_t i = 0; i < lcp.size(); ++i) sum_prefix += lcp[i];
  } else {
    for (int i = 0; i < n_leaf; ++i) {
      const long long keep = i < static_cast<int>(lcp.size()) ? lcp[i] : 0, hi = L[i];
      const long long popped = hi - keep, pushed_from = i > 0 ? lcp[i - 1] : 0;
      long long f1_end;
      if (block_size <= 0 || popped <= block_size) {
        f1_end = keep;
        sum_prefix += keep;
      } else {
        const long long n_blocks = (popped + block_size - 1) / block_size;
        const long long blk = (popped + n_blocks - 1) / n_blocks;
        f1_end = hi - blk;
        for (long long b = 0; b < n_blocks; ++b) {
          const long long x = hi - (b + 1) * blk;
          sum_prefix += x > keep ? x : keep;
        }
      }
      const long long d = f1_end - pushed_from;
      n_f1 += d > 0 ? d : 0;
    }
  }
  out5[0] = n_leaf;
  out5[1] = n_tree;
  out5[2] = sum_prefix;
  out5[3] = sum_depth;
  out5[4] = n_f1;
  return 0;
}
