/* dta_b200.h — C ABI of the B200-native DynamicTreeAttn hot path (libdta_b200.so).
 *
 * The reference (Whisper-6/DynamicTreeAttn) is pure Python and has no FFI layer of its
 * own; each entry point below names the reference code whose work it
 * replaces.  Conventions:
 *   - every pointer is a DEVICE pointer on the current CUDA device unless marked HOST;
 *   - the caller owns all memory, including workspaces (sizes from *_workspace_bytes);
 *   - every call is asynchronous on `stream` and returns 0 on success, else a non-zero
 *     code whose text is available from dta_last_error() (thread-local);
 *   - no torch types, no C++ types, no exceptions cross this boundary;
 *   - there is NO CPU fallback: without a sm_100a device the calls fail.
 *
 * Linearised trie ("layout") vocabulary, SURVEY.md §9:
 *   node n = one trie token, numbered in DFS push order; N = n_nodes; n_pad = N rounded up
 *   to a multiple of DTA_TILE; depth[n] = position id; parent[n] = node one level up (-1 at a
 *   root); subtree_end[n] = first node after n's subtree, so key k is visible to query q
 *   iff k <= q < subtree_end[k].
 */
#ifndef DTA_B200_H
#define DTA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DTA_ABI_VERSION 8
#define DTA_TILE 128 /* attention tile edge (queries and keys) */

typedef struct CUstream_st* dta_stream_t;

const char* dta_last_error(void);
int dta_abi_version(void);

/* ---------------------------------------------------------------------------------------
 * Trie indexing (integer, bit-exact with the reference)
 * ------------------------------------------------------------------------------------- */

/* Longest common prefix of every adjacent pair of n_seq sequences stored back to back in
 * `tokens` (sequence i = tokens[seq_off[i] .. seq_off[i+1])), plus a sortedness check:
 * *unsorted_flag is set to 1 if some pair is out of lexicographic order.
 * Replaces _lcp_torch / the LCP loop of _leafization (reference token_trie.py:6-10, 22-30). */
int dta_lcp_adjacent(const int64_t* tokens, const int64_t* seq_off, int n_seq, int32_t* lcp_out,
                     int32_t* unsorted_flag, dta_stream_t stream);

/* Workspace for dta_trie_linearize / dta_trie_paths (sparse tables over the fork depths). */
size_t dta_trie_workspace_bytes(int n_leaf);

/* DFS linearisation of the trie given by (lens[n_leaf], lcp[n_leaf-1]) — the push order of
 * the reference engine (reference tree_training_engine.py:541-546, 588-594).
 * Outputs: leaf_off[n_leaf+1] (exclusive warp-scan of new tokens per leaf; last = N),
 * depth[n_pad], parent[n_pad], subtree_end[n_pad], leaf_of[n_pad]; entries >= N are
 * padding (depth 0, parent -1, subtree_end 0, leaf_of -1).  n_nodes must equal
 * sum(lens) - sum(lcp). */
int dta_trie_linearize(const int32_t* lens, const int32_t* lcp, int n_leaf, int n_nodes, int n_pad,
                       int32_t* leaf_off, int32_t* depth, int32_t* parent, int32_t* subtree_end,
                       int32_t* leaf_of, void* workspace, dta_stream_t stream);

/* Root-to-leaf node ids: path[path_off[i] + d] = node at depth d on leaf i's path, d < lens[i]
 * (the stack slots the reference slices as logprobs[:L-1] / entropy[:L],
 * reference tree_training_engine.py:225-228, 397-398).  path_off = exclusive scan of lens.
 * Must run after dta_trie_linearize with the same workspace. */
int dta_trie_paths(const int32_t* lens, const int32_t* lcp, const int32_t* leaf_off, int n_leaf,
                   const int64_t* path_off, int64_t total, int32_t* path, const void* workspace,
                   dta_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Device-side trie build for CUDA-resident rollouts: the STABLE lexicographic sort
 * (reference token_trie.py:68-73), _leafization (token_trie.py:13-49) and the leaf orders of forward_permute /
 * backward_permute (trie.py:190-202, 225-233) with the LCPs TokenTrie.permute recomputes (token_trie.py:94) —
 * bit-exact with the reference's host code.  tokens: the sequences back to back (int64, signed compare);
 * offsets[n_seq + 1]: exclusive scan of their lengths.  presorted != 0: keep the given order (TokenTrie(sorted=True)).
 * out: int32[2 + 9 * n_seq], n = n_seq:
 *   out[0] n_leaf          out[1] number of ADJACENT INPUT pairs that are out of order (sorted=True check)
 *   +2      perm[n]        sorted position -> input index
 *   +2+n    slcp[n]        LCP of sorted sequence r with r + 1 (last: 0)
 *   +2+2n   absorbed[n]    1: sorted sequence r is a full prefix of its successor (folded into a later leaf)
 *   +2+3n   leaf_pos[n]    leaf j -> its sorted position          +2+4n  leaf_lcp[n]  LCP of leaf j with leaf j + 1
 *   +2+5n   order_fwd[n]   get_order_forward() over the leaves    +2+6n  lcp_fwd[n]   LCPs after forward_permute
 *   +2+7n   order_bwd[n]   get_order_backward()                   +2+8n  lcp_bwd[n]   LCPs after backward_permute
 * (leaf-indexed arrays hold n_leaf resp. n_leaf - 1 valid entries).  One small D2H copy of `out` gives the host
 * everything it needs to size the linearisation. */
#define DTA_TRIE_BUILD_MAX_SEQ 2048
size_t dta_trie_build_workspace_bytes(int n_seq);
int dta_trie_build(const int64_t* tokens, const int64_t* offsets, int n_seq, int presorted, void* workspace,
                   int32_t* out, dta_stream_t stream);

/* HOST function (no device work; every pointer is a HOST pointer): cost-model features of the sub-trie spanned by the
 * DFS-ordered leaves (lens[n_leaf], lcps[n_leaf - 1]) in the order get_order_forward (backward = 0) or
 * get_order_backward (backward != 0) visits them — reference trie.py:190-202, 225-233 followed by _get_stats
 * (trie.py:7-56; block_size <= 0 means None).  out5 = {n_leaf_sequences, n_tree_tokens, sum_prefix_len, sum_depth,
 * n_f1_tokens (0 in forward mode)}.  Used by the partitioner's probes (LB_by_DFS_and_TM, data_parallel.py:58-79):
 * the same integers as the Python code, ~100x faster. */
int dta_host_subtrie_stats(const int32_t* lens, const int32_t* lcps, int n_leaf, int backward, int64_t block_size,
                           int64_t* out5);

/* Per-tile mask summaries and work orders for the attention kernels (n_tiles = n_pad/DTA_TILE):
 *   kmax_end[t], kmin_end[t] : max / min of subtree_end over the VALID keys of tile t
 *   q_order[n_tiles]         : query tiles sorted by number of live key tiles, descending
 *   q_nlive[n_tiles]         : that number, indexed by tile
 *   k_order[n_tiles]         : key tiles sorted by weight CLASS = (live query tiles - 1) / DTA_K_ORDER_BUCKET,
 *                              descending, ties in tile (= DFS) order: heavy tiles first for load balance, but
 *                              neighbours of one subtree stay adjacent, so the CTAs resident at the same time
 *                              walk the same query tiles and their Q / dO loads and dQ reductions meet in L2
 *   k_qlast[n_tiles]         : last live query tile of key tile t (live range is [t, k_qlast[t]])
 *   pair_count[1] (int64)    : number of visible (q,k) pairs = sum(depth)+N */
#define DTA_K_ORDER_BUCKET 64
int dta_tile_meta(const int32_t* subtree_end, const int32_t* depth, int n_nodes, int n_pad,
                  int32_t* kmax_end, int32_t* kmin_end, int32_t* q_order, int32_t* q_nlive, int32_t* k_order,
                  int32_t* k_qlast, int64_t* pair_count, dta_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Tree attention (bf16 in, fp32 accumulate) — replaces the per-push/per-pop HF attention
 * calls against the KV stack (reference tree_training_engine.py:182-186, 248-252, 351-353;
 * HF models/qwen3/modeling_qwen3.py:273-287) by ONE ancestor-masked pass per layer.
 * Layout: q/out/dq [N, Hq, D], k/v/dk/dv [N, Hkv, D] with arbitrary row/head strides
 * (elements), D contiguous; D must be 128; Hq % Hkv == 0.
 * ------------------------------------------------------------------------------------- */
typedef struct {
  int32_t n_nodes, n_pad;
  int32_t n_q_heads, n_kv_heads, head_dim;
  int64_t q_row_stride, q_head_stride;
  int64_t k_row_stride, k_head_stride;
  int64_t v_row_stride, v_head_stride;
  int64_t o_row_stride, o_head_stride;
  float scale; /* softmax scale, normally head_dim^-0.5 */
  /* Backward schedule.  0: work list built on the device (any trie; a key tile whose query range would be the
   * launch's long pole is split over several CTAs).  1: one CTA per (key tile, KV head), no list kernel, for
   * callers that know the launch is balanced without it: n_tiles <= (pairs / 128^2) * n_kv_heads / (2 * n_sm) is
   * sufficient (pairs = visible query-key pairs, host-computable from lens / lcp).  Ignored by the forward. */
  int32_t bwd_whole_tiles;
} dta_attn_shape_t;

typedef struct {
  const int32_t* subtree_end; /* [n_pad] */
  const int32_t* kmax_end;    /* [n_tiles] */
  const int32_t* kmin_end;    /* [n_tiles] */
  const int32_t* q_order;     /* [n_tiles] */
  const int32_t* k_order;     /* [n_tiles] */
  const int32_t* k_qlast;     /* [n_tiles] */
} dta_attn_meta_t;

/* out[N,Hq,D] bf16 (o strides), lse[Hq, n_pad] fp32 = log2(sum_k 2^(scale*log2(e)*q.k)), i.e. the
 * log-sum-exp of the scaled scores in BASE 2 (multiply by ln 2 for the natural-log value). */
int dta_tree_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse,
                      const dta_attn_meta_t* meta, const dta_attn_shape_t* shape, dta_stream_t stream);

/* Measurement hook (bench.py roofline): the NEXT dta_tree_attn_fwd / dta_tree_attn_bwd call on this
 * host thread records `start` immediately before and `stop` immediately after its MAIN kernel
 * (cudaEvent_t handles, on the call's stream), then the hook clears itself.  NULLs disarm it. */
int dta_profile_next_attn(void* start_event, void* stop_event);

/* Workspace (256-byte aligned, contents need not be initialised): fp32 dq accumulator [n_pad,Hq,D], the lse / delta
 * operand rows [Hq,n_pad] x 16 B each, the launch's work list (key tile x chunk of its live query tiles; a trunk tile
 * that would be the launch's long pole is split over several CTAs), and fp32 dk / dv accumulators [n_pad,Hkv,D] for
 * the split tiles.  About (4*Hq + 8*Hkv) * D + 32*Hq bytes per node. */
size_t dta_tree_attn_bwd_workspace_bytes(const dta_attn_shape_t* shape);

/* dout uses the o strides; dq uses the q strides, dk the k strides, dv the v strides. */
int dta_tree_attn_bwd(const void* q, const void* k, const void* v, const void* out, const void* dout,
                      const float* lse, void* dq, void* dk, void* dv, const dta_attn_meta_t* meta,
                      const dta_attn_shape_t* shape, void* workspace, dta_stream_t stream);

/* Same backward with dQ left in fp32: dq_f32[n_pad, Hq, D] (16-byte aligned, contents need not be initialised) is used
 * as the accumulator of the kernel's tensor reduce-adds and handed back as is — no bf16 dq tensor, no conversion pass.
 * Consumer: dta_qknorm_rope_bwd_f32dq (the q/k-norm + RoPE backward sits right behind attention in the model,
 * HF models/qwen3/modeling_qwen3.py:263-268). */
int dta_tree_attn_bwd_f32dq(const void* q, const void* k, const void* v, const void* out, const void* dout,
                            const float* lse, float* dq_f32, void* dk, void* dv, const dta_attn_meta_t* meta,
                            const dta_attn_shape_t* shape, void* workspace, dta_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Log-prob + entropy over a vocab shard — replaces _gather_logprobs[_entropy] and the
 * vocab-parallel autograd functions (reference vocab_parallel.py:13-27, 117-166, 249-321,
 * 324-370).  x = logit * inv_temperature.  dtype: 0 = bf16, 1 = fp32.
 * ------------------------------------------------------------------------------------- */

/* One pass over logits[n_rows, n_cols] (row_stride elements): stats[r] = {m, sum e^(x-m), sum x e^(x-m), x_label}
 * for this shard, x_label = x[r][labels[r] - vocab_start] if labels != NULL and the label falls in this shard,
 * else 0 (the reference's masked label gather, vocab_parallel.py:145-156, in the same launch).  These 16 bytes per
 * row are ALL a vocab-parallel rank exchanges. */
int dta_logprob_row_stats(const void* logits, int dtype, int64_t n_rows, int64_t n_cols, int64_t row_stride,
                          float inv_temperature, const int64_t* labels, int64_t vocab_start, float* stats,
                          dta_stream_t stream);

/* x_label[m * x_stride] = x[row_of[m]][labels[m] - vocab_start] if the label falls in this shard
 * else 0.  row_of == NULL means row_of[m] = m.  row_of[m] < 0 yields 0 (root nodes have no
 * edge).  For SEVERAL labels per row (trie edges: a fork row has one label per child); with one label per row
 * dta_logprob_row_stats gathers it itself. */
int dta_logprob_gather_labels(const void* logits, int dtype, int64_t n_cols, int64_t row_stride,
                              float inv_temperature, const int64_t* labels, const int32_t* row_of,
                              int64_t n_labels, int64_t vocab_start, float* x_label, int64_t x_stride,
                              dta_stream_t stream);

/* Merge `world` shard statistics with the online-softmax rule (the reference's MAX + 3 SUM all-reduces,
 * vocab_parallel.py:262-304, after ONE all-gather): rank w's rows are stats_all[w * world_stride + 4 r .. + 4)
 * (world_stride in floats, >= 4 n_rows; the gathered block may carry extra words behind the row statistics).
 * logz[r] = log sum_v e^x, entropy[r] = logz - E_p[x]; logprob (may be NULL) = sum over shards of the 4th
 * statistics word - logz = the log-prob of the row's label. */
int dta_logprob_finalize(const float* stats_all, int world, int64_t world_stride, int64_t n_rows, float* logz,
                         float* entropy, float* logprob, dta_stream_t stream);

/* dlogits[r,c] = inv_temperature * p * ( g_entropy[r] * (-entropy[r] - log p) - g_lp_rowsum[r] ),
 * p = e^(x - logz[r]); written in `dtype`; dlogits may alias logits (in place).
 * g_lp_rowsum[r] = sum of g_lp over the labels attached to row r.  Either grad may be NULL.
 * labels != NULL (ONE label per row; g_lp_rowsum[r] is then that label's g_lp): the one-hot term
 * dlogits[r][labels[r] - vocab_start] += inv_temperature * g_lp_rowsum[r] is added in the same pass for labels in
 * this shard (reference vocab_parallel.py:337-368), and no dta_logprob_bwd_scatter call is needed. */
int dta_logprob_bwd(const void* logits, int dtype, int64_t n_rows, int64_t n_cols, int64_t row_stride,
                    float inv_temperature, const float* logz, const float* entropy, const float* g_entropy,
                    const float* g_lp_rowsum, const int64_t* labels, int64_t vocab_start, void* dlogits,
                    int64_t d_row_stride, dta_stream_t stream);

/* dlogits[row_of[m]][labels[m]-vocab_start] += inv_temperature * g_lp[m] for labels in the shard.
 * (row, label) pairs must be unique — true for trie edges and for one label per row. */
int dta_logprob_bwd_scatter(void* dlogits, int dtype, int64_t n_cols, int64_t d_row_stride,
                            float inv_temperature, const int64_t* labels, const int32_t* row_of,
                            const float* g_lp, int64_t n_labels, int64_t vocab_start, dta_stream_t stream);

/* g_lp_rowsum[row_of[m]] += g_lp[m] (caller zero-fills g_lp_rowsum). */
int dta_logprob_rowsum_grad(const float* g_lp, const int32_t* row_of, int64_t n_labels, float* g_lp_rowsum,
                            dta_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * lm_head fused into the log-prob / entropy statistics: replaces
 * `logits = lm_head(hidden)` (HF models/qwen3/modeling_qwen3.py:505) followed by the fp32 log_softmax / entropy of
 * reference vocab_parallel.py:13-27 WITHOUT materialising the [rows, vocab] logits: a tcgen05 (cta_group::2,
 * 256 x 256 x 64) GEMM whose accumulator tiles are reduced in the epilogue, straight out of TMEM.
 * hidden [n_rows, hidden_size], weight [n_cols, hidden_size]: bf16, unit inner stride, row strides in elements
 * (multiples of 8), 16-byte aligned.  stats: float[n_splits][rows_pad][4] = {m, sum e^(x-m), sum x e^(x-m), 0} per
 * vocab split and row, n_splits = dta_lm_head_stats_splits(n_rows, n_cols), rows_pad =
 * dta_lm_head_stats_rows_pad(n_rows); merge with dta_logprob_finalize(stats, n_splits, 4 * rows_pad, n_rows, ...).
 * ------------------------------------------------------------------------------------- */
int dta_lm_head_stats_splits(int64_t n_rows, int64_t n_cols);
int64_t dta_lm_head_stats_rows_pad(int64_t n_rows);
int dta_lm_head_stats(const void* hidden, int64_t h_row_stride, const void* weight, int64_t w_row_stride, int64_t n_rows,
                      int64_t n_cols, int hidden_size, float inv_temperature, float* stats, dta_stream_t stream);

/* Same launch that ALSO keeps the logits (bf16, logits[n_rows][l_row_stride]) for a backward that wants them: written
 * once under the MMAs, never read in the forward; the statistics are those of the ROUNDED values (= of the stored
 * tensor, as if it had been produced by a library GEMM and reduced by dta_logprob_row_stats). */
int dta_lm_head_stats_logits(const void* hidden, int64_t h_row_stride, const void* weight, int64_t w_row_stride, int64_t n_rows,
                             int64_t n_cols, int hidden_size, float inv_temperature, float* stats, void* logits,
                             int64_t l_row_stride, dta_stream_t stream);

/* Backward of the same fusion: the SAME GEMM mainloop recomputes each logits tile and the epilogue turns it into
 * d loss / d logits (formula of dta_logprob_bwd, without the one-hot label terms: add them with
 * dta_logprob_bwd_scatter), stored in bf16: dlogits[n_rows][d_row_stride].  The reference keeps the fp32 softmax for
 * this (vocab_parallel.py:312-318); here neither logits nor softmax exist between the passes.  The caller then forms
 * d hidden = dlogits W and d W = dlogits^T hidden with library GEMMs, row chunk by row chunk. */
int dta_lm_head_dlogits(const void* hidden, int64_t h_row_stride, const void* weight, int64_t w_row_stride, int64_t n_rows,
                        int64_t n_cols, int hidden_size, float inv_temperature, const float* logz, const float* entropy,
                        const float* g_entropy, const float* g_lp_rowsum, void* dlogits, int64_t d_row_stride,
                        dta_stream_t stream);

/* x_label[m] = inv_temperature * <hidden[row_of[m]], weight[labels[m]]>: the logit of one (row, label) edge without
 * the logits row (fp32 accumulation of bf16 products).  row_of == NULL: row_of[m] = m; row_of[m] < 0 or a label
 * outside [0, n_cols) yields 0. */
int dta_lm_head_label_logits(const void* hidden, int64_t h_row_stride, const void* weight, int64_t w_row_stride,
                             int hidden_size, int64_t n_cols, const int64_t* labels, const int32_t* row_of,
                             int64_t n_labels, float inv_temperature, float* x_label, dta_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Fused layer kernels either side of attention — replace the ATen
 * elementwise chains of the HF reference model code (HF models/qwen3/modeling_qwen3.py:50-64 RMSNorm,
 * :81-83 + :137-181 RoPE, :263-268 q/k norm, :81-83 SwiGLU MLP).  bf16 tensors, fp32 math.
 * ------------------------------------------------------------------------------------- */

/* s = bf16(x + residual) (residual may be NULL: s = x); y = bf16(s * rstd * weight),
 * rstd[r] = rsqrt(mean(s^2) + eps).  x/residual/y/sum_out: contiguous [n_rows, hidden]. */
int dta_rmsnorm_fwd(const void* x, const void* residual, const void* weight, int64_t n_rows, int hidden, float eps,
                    void* y, void* sum_out, float* rstd, dta_stream_t stream);

/* dx = rstd * (g - xh * mean(g * xh)) + dresidual (NULL = 0), g = dy * weight, xh = x * rstd;
 * dweight[hidden] (fp32) += sum_rows dy * xh.  `x` is the tensor that was normalised (sum_out). */
int dta_rmsnorm_bwd(const void* dy, const void* x, const void* weight, const float* rstd, const void* dresidual,
                    int64_t n_rows, int hidden, void* dx, float* dweight, dta_stream_t stream);

/* q_out[n,h,:] = RoPE_{depth[n]}(RMSNorm(q[n,h,:]) * wq), same for k.  q/k: strided [n_rows, H, 128] views
 * of the projection output; q_out/k_out contiguous; cos_sin[pos][0:64] = cos, [64:128] = sin (fp32);
 * rstd[n_rows, Hq+Hkv] is saved for the backward. */
int dta_qknorm_rope_fwd(const void* q, int64_t q_row_stride, int64_t q_head_stride, const void* k,
                        int64_t k_row_stride, int64_t k_head_stride, const void* wq, const void* wk,
                        const int32_t* depth, const float* cos_sin, int64_t n_rows, int n_q_heads, int n_kv_heads,
                        int head_dim, float eps, void* q_out, void* k_out, float* rstd, dta_stream_t stream);

int dta_qknorm_rope_bwd(const void* dq_out, const void* dk_out, const void* q, int64_t q_row_stride,
                        int64_t q_head_stride, const void* k, int64_t k_row_stride, int64_t k_head_stride,
                        const void* wq, const void* wk, const int32_t* depth, const float* cos_sin, const float* rstd,
                        int64_t n_rows, int n_q_heads, int n_kv_heads, int head_dim, void* dq, int64_t dq_row_stride,
                        int64_t dq_head_stride, void* dk, int64_t dk_row_stride, int64_t dk_head_stride, float* dwq,
                        float* dwk, dta_stream_t stream);

/* dta_qknorm_rope_bwd with d q_out given in fp32, contiguous [n_rows, Hq, 128] (the accumulator dta_tree_attn_bwd_f32dq
 * leaves behind); d k_out stays bf16.  Needs 16-byte aligned tensors and strides that are multiples of 8 elements. */
int dta_qknorm_rope_bwd_f32dq(const float* dq_out_f32, const void* dk_out, const void* q, int64_t q_row_stride,
                              int64_t q_head_stride, const void* k, int64_t k_row_stride, int64_t k_head_stride,
                              const void* wq, const void* wk, const int32_t* depth, const float* cos_sin, const float* rstd,
                              int64_t n_rows, int n_q_heads, int n_kv_heads, int head_dim, void* dq, int64_t dq_row_stride,
                              int64_t dq_head_stride, void* dk, int64_t dk_row_stride, int64_t dk_head_stride, float* dwq,
                              float* dwk, dta_stream_t stream);

/* out = silu(gate) * up;  backward returns d gate and d up. */
int dta_swiglu_fwd(const void* gate, const void* up, int64_t n_elems, void* out, dta_stream_t stream);
int dta_swiglu_bwd(const void* dout, const void* gate, const void* up, int64_t n_elems, void* dgate, void* dup,
                   dta_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DTA_B200_H */
