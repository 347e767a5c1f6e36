// layer_kernels.cu — fused elementwise/normalisation kernels either side of tree attention
// (SURVEY.md §8f row N3): residual-add + RMSNorm, per-head q/k RMSNorm + RoPE(depth), SwiGLU, each
// forward and backward.  They replace ~40 ATen elementwise launches per decoder layer of the HF
// reference model code (HF models/qwen3/modeling_qwen3.py:50-64, 81-83, 137-181, 263-268), which
// cost more than the attention kernels once attention runs on tcgen05.  All are HBM-bound: bf16 in
// and out, fp32 math, 16-byte accesses, one pass over each tensor.
#include "dta_common.cuh"
#include <cstdlib>
#include "../../include/dta_b200.h"
#include <cuda_bf16.h>

namespace dta {

__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float* f) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float* f) {
  uint4 v;
  uint32_t* w = &v.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = v;
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ float block_sum(float v, float* smem) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  float t = lane < nw ? smem[lane] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  __syncthreads();
  return t;
}

// ------------------------------------------------------------------------------------
// (residual add +) RMSNorm.  blockDim.x == H / 8, one row per CTA iteration.
//   s = bf16(x + res)   (written to `sum_out` when res != NULL; HF adds the residual in bf16)
//   y = bf16( s * rstd * w ),  rstd = rsqrt(mean(s^2) + eps)
// ------------------------------------------------------------------------------------
// Both kernels are latency-bound per row (load -> block reduction -> store), so the loads of the CTA's NEXT row are
// issued before the reduction of the current one: two rows per CTA in flight, 8 CTAs per SM.
__device__ __forceinline__ void cvt8(const uint4 v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

__global__ void __launch_bounds__(1024) rmsnorm_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                           const __nv_bfloat16* __restrict__ res,
                                                           const __nv_bfloat16* __restrict__ w, int64_t n_rows, int H,
                                                           float eps, __nv_bfloat16* __restrict__ y,
                                                           __nv_bfloat16* __restrict__ sum_out,
                                                           float* __restrict__ rstd_out) {
  __shared__ float red[32];
  const int c = threadIdx.x * 8;
  float wf[8];
  ld8(w + c, wf);
  const bool has_res = res != nullptr;
  int64_t row = blockIdx.x;
  uint4 xr = make_uint4(0, 0, 0, 0), rr = make_uint4(0, 0, 0, 0);
  if (row < n_rows) {
    xr = ldg16(x + row * H + c);
    if (has_res) rr = ldg16(res + row * H + c);
  }
  while (row < n_rows) {
    const int64_t next = row + gridDim.x;
    uint4 xn = make_uint4(0, 0, 0, 0), rn = make_uint4(0, 0, 0, 0);
    if (next < n_rows) {
      xn = ldg16(x + next * H + c);
      if (has_res) rn = ldg16(res + next * H + c);
    }
    float v[8];
    cvt8(xr, v);
    if (has_res) {
      float r[8];
      cvt8(rr, r);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = bf16_round(v[i] + r[i]);
      st8(sum_out + row * H + c, v);
    }
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) ss = fmaf(v[i], v[i], ss);
    ss = block_sum(ss, red);
    const float rstd = rsqrtf(ss / H + eps);
    if (threadIdx.x == 0) rstd_out[row] = rstd;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = v[i] * rstd * wf[i];
    st8(y + row * H + c, v);
    xr = xn;
    rr = rn;
    row = next;
  }
}

// dx = rstd * (g - xh * mean(g * xh)) [+ dres],  g = dy * w,  xh = x * rstd;   dw += sum_rows dy * xh
__global__ void __launch_bounds__(1024) rmsnorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy,
                                                           const __nv_bfloat16* __restrict__ x,
                                                           const __nv_bfloat16* __restrict__ w,
                                                           const float* __restrict__ rstd_in,
                                                           const __nv_bfloat16* __restrict__ dres, int64_t n_rows,
                                                           int H, __nv_bfloat16* __restrict__ dx,
                                                           float* __restrict__ dw) {
  __shared__ float red[32];
  const int c = threadIdx.x * 8;
  float wf[8], dwacc[8];
  ld8(w + c, wf);
#pragma unroll
  for (int i = 0; i < 8; ++i) dwacc[i] = 0.f;
  const bool has_res = dres != nullptr;
  int64_t row = blockIdx.x;
  uint4 gr = make_uint4(0, 0, 0, 0), xr = gr, rr = gr;
  float rstd = 0.f;
  if (row < n_rows) {
    gr = ldg16(dy + row * H + c);
    xr = ldg16(x + row * H + c);
    if (has_res) rr = ldg16(dres + row * H + c);
    rstd = __ldg(rstd_in + row);
  }
  while (row < n_rows) {
    const int64_t next = row + gridDim.x;
    uint4 gn = make_uint4(0, 0, 0, 0), xn = gn, rn = gn;
    float rstd_n = 0.f;
    if (next < n_rows) {
      gn = ldg16(dy + next * H + c);
      xn = ldg16(x + next * H + c);
      if (has_res) rn = ldg16(dres + next * H + c);
      rstd_n = __ldg(rstd_in + next);
    }
    float g[8], xv[8];
    cvt8(gr, g);
    cvt8(xr, xv);
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      xv[i] *= rstd;
      dwacc[i] = fmaf(g[i], xv[i], dwacc[i]);
      g[i] *= wf[i];
      dot = fmaf(g[i], xv[i], dot);
    }
    dot = block_sum(dot, red) / H;
    float out[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = rstd * (g[i] - xv[i] * dot);
    if (has_res) {
      float r[8];
      cvt8(rr, r);
#pragma unroll
      for (int i = 0; i < 8; ++i) out[i] += r[i];
    }
    st8(dx + row * H + c, out);
    gr = gn;
    xr = xn;
    rr = rn;
    rstd = rstd_n;
    row = next;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) atomicAdd(dw + c + i, dwacc[i]);
}

// Warp-per-row backward for hidden <= 1024 (kVec = hidden / 256 16-byte vectors per lane and tensor): no block-wide
// reduction on the row's critical path, every load of a row issued at once, 16-24 rows in flight per SM.
// dw: per-lane fp32 partials -> shared-memory reduction over the CTA's warps -> one global atomic per column and CTA.
template <int kVec>
__global__ void __launch_bounds__(256, 2) rmsnorm_bwd_warp_kernel(const __nv_bfloat16* __restrict__ dy,
                                                              const __nv_bfloat16* __restrict__ x,
                                                              const __nv_bfloat16* __restrict__ w,
                                                              const float* __restrict__ rstd_in,
                                                              const __nv_bfloat16* __restrict__ dres, int64_t n_rows,
                                                              __nv_bfloat16* __restrict__ dx, float* __restrict__ dw) {
  constexpr int H = kVec * 256;
  __shared__ float sdw[H];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < H; i += blockDim.x) sdw[i] = 0.f;
  float dwacc[kVec][8];
#pragma unroll
  for (int v = 0; v < kVec; ++v)
#pragma unroll
    for (int i = 0; i < 8; ++i) dwacc[v][i] = 0.f;
  const bool has_res = dres != nullptr;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t row = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; row < n_rows; row += n_warps) {
    uint4 gr[kVec], xr[kVec], rr[kVec];
#pragma unroll
    for (int v = 0; v < kVec; ++v) {
      const int64_t off = row * H + v * 256 + lane * 8;
      gr[v] = ldg16(dy + off);
      xr[v] = ldg16(x + off);
      rr[v] = has_res ? ldg16(dres + off) : make_uint4(0, 0, 0, 0);
    }
    const float rstd = __ldg(rstd_in + row);
    float dot = 0.f;
    // two passes over the raw registers (the bf16 -> fp32 unpack is two ALU ops) instead of keeping fp32 copies live
#pragma unroll
    for (int v = 0; v < kVec; ++v) {
      float wf[8], g[8], xv[8];
      ld8(w + v * 256 + lane * 8, wf);
      cvt8(gr[v], g);
      cvt8(xr[v], xv);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xh = xv[i] * rstd;
        dwacc[v][i] = fmaf(g[i], xh, dwacc[v][i]);
        dot = fmaf(g[i] * wf[i], xh, dot);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    dot *= (1.f / H);
#pragma unroll
    for (int v = 0; v < kVec; ++v) {
      float wf[8], g[8], xv[8], r[8], out[8];
      ld8(w + v * 256 + lane * 8, wf);
      cvt8(gr[v], g);
      cvt8(xr[v], xv);
      cvt8(rr[v], r);
#pragma unroll
      for (int i = 0; i < 8; ++i) out[i] = rstd * (g[i] * wf[i] - xv[i] * rstd * dot) + r[i];
      st8(dx + row * H + v * 256 + lane * 8, out);
    }
  }
  __syncthreads();
#pragma unroll
  for (int v = 0; v < kVec; ++v)
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&sdw[v * 256 + lane * 8 + i], dwacc[v][i]);
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += blockDim.x) atomicAdd(dw + i, sdw[i]);
}

// ------------------------------------------------------------------------------------
// per-head RMSNorm + RoPE(depth) on q and k, head_dim 128.  One warp per (node, head); lane l owns
// elements 4l..4l+3, its rotate_half partner is lane l^16.  cos_sin[pos][0..63] = cos, [64..127] = sin.
// ------------------------------------------------------------------------------------
struct QkArgs {
  const __nv_bfloat16* q; int64_t q_row, q_head;       // raw projections (strided views)
  const __nv_bfloat16* k; int64_t k_row, k_head;
  __nv_bfloat16* q_out; __nv_bfloat16* k_out;           // contiguous [N, H, 128]
  const __nv_bfloat16* wq; const __nv_bfloat16* wk;
  const int32_t* depth; const float* cos_sin;
  float* rstd;                                          // [N, Hq + Hkv]
  int64_t n_rows; int hq, hkv; float eps;
};

__global__ void __launch_bounds__(256) qknorm_rope_fwd_kernel(const QkArgs a) {
  const int lane = threadIdx.x & 31;
  const int heads = a.hq + a.hkv;
  const int64_t total = a.n_rows * heads;
  float wq[4], wk[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    wq[j] = __bfloat162float(a.wq[lane * 4 + j]);
    wk[j] = __bfloat162float(a.wk[lane * 4 + j]);
  }
  for (int64_t wi = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; wi < total;
       wi += (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5) {
    const int64_t n = wi / heads;
    const int h = static_cast<int>(wi % heads);
    const bool is_q = h < a.hq;
    const __nv_bfloat16* src = is_q ? a.q + n * a.q_row + h * a.q_head : a.k + n * a.k_row + (h - a.hq) * a.k_head;
    __nv_bfloat16* dst = is_q ? a.q_out + (n * a.hq + h) * 128 : a.k_out + (n * a.hkv + (h - a.hq)) * 128;
    const uint2 raw = *reinterpret_cast<const uint2*>(src + lane * 4);
    float v[4] = {__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u), __uint_as_float(raw.y << 16),
                  __uint_as_float(raw.y & 0xffff0000u)};
    float ss = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss * (1.f / 128.f) + a.eps);
    if (lane == 0) a.rstd[wi] = rstd;
    const float4 cs = *reinterpret_cast<const float4*>(a.cos_sin + static_cast<int64_t>(a.depth[n]) * 128 + (lane & 15) * 4);
    const float4 sn = *reinterpret_cast<const float4*>(a.cos_sin + static_cast<int64_t>(a.depth[n]) * 128 + 64 + (lane & 15) * 4);
    const float c4[4] = {cs.x, cs.y, cs.z, cs.w}, s4[4] = {sn.x, sn.y, sn.z, sn.w};
    float out[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float xh = v[j] * rstd * (is_q ? wq[j] : wk[j]);
      const float partner = __shfl_xor_sync(0xffffffffu, xh, 16);
      out[j] = lane < 16 ? xh * c4[j] - partner * s4[j] : xh * c4[j] + partner * s4[j];
    }
    uint2 o;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(out[0], out[1]), h1 = __floats2bfloat162_rn(out[2], out[3]);
    o.x = *reinterpret_cast<uint32_t*>(&h0);
    o.y = *reinterpret_cast<uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(dst + lane * 4) = o;
  }
}

struct QkBwdArgs {
  const __nv_bfloat16* dq_out; const __nv_bfloat16* dk_out;   // contiguous [N, H, 128]
  const float* dq_out_f32;                                    // kDqF32: the attention backward's fp32 dQ accumulator, read directly
  const __nv_bfloat16* q; int64_t q_row, q_head;              // raw projections
  const __nv_bfloat16* k; int64_t k_row, k_head;
  __nv_bfloat16* dq; int64_t dq_row, dq_head;                 // gradients of the raw projections
  __nv_bfloat16* dk; int64_t dk_row, dk_head;
  const __nv_bfloat16* wq; const __nv_bfloat16* wk;
  const int32_t* depth; const float* cos_sin; const float* rstd;
  float* dwq; float* dwk;                                     // fp32 [128], accumulated
  int64_t n_rows; int hq, hkv;
};

__global__ void __launch_bounds__(256) qknorm_rope_bwd_kernel(const QkBwdArgs a) {
  const int lane = threadIdx.x & 31;
  const int heads = a.hq + a.hkv;
  const int64_t total = a.n_rows * heads;
  float wq[4], wk[4], accq[4] = {0, 0, 0, 0}, acck[4] = {0, 0, 0, 0};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    wq[j] = __bfloat162float(a.wq[lane * 4 + j]);
    wk[j] = __bfloat162float(a.wk[lane * 4 + j]);
  }
  for (int64_t wi = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; wi < total;
       wi += (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5) {
    const int64_t n = wi / heads;
    const int h = static_cast<int>(wi % heads);
    const bool is_q = h < a.hq;
    const int hh = is_q ? h : h - a.hq;
    const __nv_bfloat16* gsrc = is_q ? a.dq_out + (n * a.hq + hh) * 128 : a.dk_out + (n * a.hkv + hh) * 128;
    const __nv_bfloat16* xsrc = is_q ? a.q + n * a.q_row + hh * a.q_head : a.k + n * a.k_row + hh * a.k_head;
    __nv_bfloat16* dst = is_q ? a.dq + n * a.dq_row + hh * a.dq_head : a.dk + n * a.dk_row + hh * a.dk_head;
    const uint2 graw = *reinterpret_cast<const uint2*>(gsrc + lane * 4);
    const uint2 xraw = *reinterpret_cast<const uint2*>(xsrc + lane * 4);
    const float go[4] = {__uint_as_float(graw.x << 16), __uint_as_float(graw.x & 0xffff0000u),
                         __uint_as_float(graw.y << 16), __uint_as_float(graw.y & 0xffff0000u)};
    const float xv[4] = {__uint_as_float(xraw.x << 16), __uint_as_float(xraw.x & 0xffff0000u),
                         __uint_as_float(xraw.y << 16), __uint_as_float(xraw.y & 0xffff0000u)};
    const float rstd = a.rstd[wi];
    const float4 cs = *reinterpret_cast<const float4*>(a.cos_sin + static_cast<int64_t>(a.depth[n]) * 128 + (lane & 15) * 4);
    const float4 sn = *reinterpret_cast<const float4*>(a.cos_sin + static_cast<int64_t>(a.depth[n]) * 128 + 64 + (lane & 15) * 4);
    const float c4[4] = {cs.x, cs.y, cs.z, cs.w}, s4[4] = {sn.x, sn.y, sn.z, sn.w};
    float g[4], xh[4], dot = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float partner = __shfl_xor_sync(0xffffffffu, go[j], 16);
      // transpose of the rotation: d xh[i] = do[i] c + do[i+64] s ; d xh[i+64] = do[i+64] c - do[i] s
      const float dxh = lane < 16 ? go[j] * c4[j] + partner * s4[j] : go[j] * c4[j] - partner * s4[j];
      xh[j] = xv[j] * rstd;
      if (is_q) accq[j] = fmaf(dxh, xh[j], accq[j]); else acck[j] = fmaf(dxh, xh[j], acck[j]);
      g[j] = dxh * (is_q ? wq[j] : wk[j]);
      dot = fmaf(g[j], xh[j], dot);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    dot *= (1.f / 128.f);
    float out[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = rstd * (g[j] - xh[j] * dot);
    uint2 o;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(out[0], out[1]), h1 = __floats2bfloat162_rn(out[2], out[3]);
    o.x = *reinterpret_cast<uint32_t*>(&h0);
    o.y = *reinterpret_cast<uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(dst + lane * 4) = o;
  }
  // weight gradients: reduce over the CTA's warps in shared memory, one atomic per column per CTA
  __shared__ float sq[128], sk[128];
  if (threadIdx.x < 128) { sq[threadIdx.x] = 0.f; sk[threadIdx.x] = 0.f; }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    atomicAdd(&sq[lane * 4 + j], accq[j]);
    atomicAdd(&sk[lane * 4 + j], acck[j]);
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    atomicAdd(a.dwq + threadIdx.x, sq[threadIdx.x]);
    atomicAdd(a.dwk + threadIdx.x, sk[threadIdx.x]);
  }
}

// ------------------------------------------------------------------------------------
// 16-byte versions of the two kernels above (strides multiples of 8 elements, 16-byte aligned bases): one warp per
// NODE, one half-warp per head row (lane hl of the half owns elements 8hl..8hl+7, its rotate_half partner is lane
// hl ^ 8), four head pairs loaded before the first is reduced.  The node's depth / cos / sin are fetched once for
// all of its heads, and 8 rows x 256 B are in flight per warp instead of one.
// ------------------------------------------------------------------------------------
static constexpr int kQkUnroll = 4;   // head PAIRS in flight per warp

// kMinBlocks: resident blocks per SM the register allocation must allow (occupancy vs registers: A/B knob DTA_QK_OCC)
template <int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks) qknorm_rope_fwd_v8_kernel(const QkArgs a) {
  const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
  const int heads = a.hq + a.hkv;
  float wq[8], wk[8];
  ld8(a.wq + hl * 8, wq);
  ld8(a.wk + hl * 8, wk);
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t n = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; n < a.n_rows; n += n_warps) {
    const float* cs_row = a.cos_sin + static_cast<int64_t>(__ldg(a.depth + n)) * 128 + (hl & 7) * 8;
    float cs[8], sn[8];
    *reinterpret_cast<float4*>(cs) = __ldg(reinterpret_cast<const float4*>(cs_row));
    *reinterpret_cast<float4*>(cs + 4) = __ldg(reinterpret_cast<const float4*>(cs_row + 4));
    *reinterpret_cast<float4*>(sn) = __ldg(reinterpret_cast<const float4*>(cs_row + 64));
    *reinterpret_cast<float4*>(sn + 4) = __ldg(reinterpret_cast<const float4*>(cs_row + 68));
    if (hl < 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sn[j] = -sn[j];          // first half of the head: x c - partner s
    }
    for (int h0 = 0; h0 < heads; h0 += 2 * kQkUnroll) {
      uint4 raw[kQkUnroll];
#pragma unroll
      for (int u = 0; u < kQkUnroll; ++u) {
        const int h = h0 + 2 * u + half;
        raw[u] = make_uint4(0, 0, 0, 0);
        if (h < heads) {
          const __nv_bfloat16* src = h < a.hq ? a.q + n * a.q_row + h * a.q_head : a.k + n * a.k_row + (h - a.hq) * a.k_head;
          raw[u] = ldg16(src + hl * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < kQkUnroll; ++u) {
        const int h = h0 + 2 * u + half;
        const bool is_q = h < a.hq;
        float v[8];
        cvt8(raw[u], v);
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) ss = fmaf(v[j], v[j], ss);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float rstd = rsqrtf(ss * (1.f / 128.f) + a.eps);
        float out[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = v[j] * rstd * (is_q ? wq[j] : wk[j]);
          const float partner = __shfl_xor_sync(0xffffffffu, xh, 8);
          out[j] = fmaf(partner, sn[j], xh * cs[j]);
        }
        if (h < heads) {
          if (hl == 0) a.rstd[n * heads + h] = rstd;
          __nv_bfloat16* dst = is_q ? a.q_out + (n * a.hq + h) * 128 : a.k_out + (n * a.hkv + (h - a.hq)) * 128;
          st8(dst + hl * 8, out);
        }
      }
    }
  }
}

// kDqF32: d q_out comes as the fp32 accumulator of the attention backward ([N, Hq, 128] contiguous): the bf16 dq tensor and
// the kernel that used to write it (attn_dq_convert) drop out of the step
template <int kMinBlocks, bool kDqF32, int kUnroll>
__global__ void __launch_bounds__(256, kMinBlocks) qknorm_rope_bwd_v8_kernel(const QkBwdArgs a) {
  const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
  const int heads = a.hq + a.hkv;
  float wq[8], wk[8], accq[8], acck[8];
  ld8(a.wq + hl * 8, wq);
  ld8(a.wk + hl * 8, wk);
#pragma unroll
  for (int j = 0; j < 8; ++j) accq[j] = acck[j] = 0.f;
  const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t n = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5; n < a.n_rows; n += n_warps) {
    const float* cs_row = a.cos_sin + static_cast<int64_t>(__ldg(a.depth + n)) * 128 + (hl & 7) * 8;
    float cs[8], sn[8];
    *reinterpret_cast<float4*>(cs) = __ldg(reinterpret_cast<const float4*>(cs_row));
    *reinterpret_cast<float4*>(cs + 4) = __ldg(reinterpret_cast<const float4*>(cs_row + 4));
    *reinterpret_cast<float4*>(sn) = __ldg(reinterpret_cast<const float4*>(cs_row + 64));
    *reinterpret_cast<float4*>(sn + 4) = __ldg(reinterpret_cast<const float4*>(cs_row + 68));
    if (hl >= 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sn[j] = -sn[j];          // transpose of the rotation: second half gets - do[i] s
    }
    for (int h0 = 0; h0 < heads; h0 += 2 * kUnroll) {
      uint4 graw[kUnroll], xraw[kUnroll];
      float4 gf[kDqF32 ? kUnroll : 1][2];
      float rstd_u[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int h = h0 + 2 * u + half;
        graw[u] = xraw[u] = make_uint4(0, 0, 0, 0);
        rstd_u[u] = 0.f;
        if (h < heads) {
          const bool is_q = h < a.hq;
          const int hh = is_q ? h : h - a.hq;
          if (kDqF32 && is_q) {
            const float4* src = reinterpret_cast<const float4*>(a.dq_out_f32 + (n * a.hq + hh) * 128 + hl * 8);
            gf[kDqF32 ? u : 0][0] = __ldg(src);
            gf[kDqF32 ? u : 0][1] = __ldg(src + 1);
          } else {
            graw[u] = ldg16((is_q ? a.dq_out + (n * a.hq + hh) * 128 : a.dk_out + (n * a.hkv + hh) * 128) + hl * 8);
          }
          xraw[u] = ldg16((is_q ? a.q + n * a.q_row + hh * a.q_head : a.k + n * a.k_row + hh * a.k_head) + hl * 8);
          rstd_u[u] = __ldg(a.rstd + n * heads + h);
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int h = h0 + 2 * u + half;
        const bool is_q = h < a.hq;
        const int hh = is_q ? h : h - a.hq;
        float go[8], xh[8], g[8];
        if (kDqF32 && is_q) {
          const float4 g0 = gf[kDqF32 ? u : 0][0], g1 = gf[kDqF32 ? u : 0][1];
          go[0] = g0.x; go[1] = g0.y; go[2] = g0.z; go[3] = g0.w;
          go[4] = g1.x; go[5] = g1.y; go[6] = g1.z; go[7] = g1.w;
        } else {
          cvt8(graw[u], go);
        }
        cvt8(xraw[u], xh);
        const float rstd = rstd_u[u];
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float partner = __shfl_xor_sync(0xffffffffu, go[j], 8);
          // d xh[i] = do[i] c + do[i+64] s ; d xh[i+64] = do[i+64] c - do[i] s
          const float dxh = fmaf(partner, sn[j], go[j] * cs[j]);
          xh[j] *= rstd;
          if (is_q) accq[j] = fmaf(dxh, xh[j], accq[j]); else acck[j] = fmaf(dxh, xh[j], acck[j]);
          g[j] = dxh * (is_q ? wq[j] : wk[j]);
          dot = fmaf(g[j], xh[j], dot);
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        dot *= (1.f / 128.f);
        if (h < heads) {
          float out[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) out[j] = rstd * (g[j] - xh[j] * dot);
          __nv_bfloat16* dst = is_q ? a.dq + n * a.dq_row + hh * a.dq_head : a.dk + n * a.dk_row + hh * a.dk_head;
          st8(dst + hl * 8, out);
        }
      }
    }
  }
  // weight gradients: reduce over the CTA's warps in shared memory, one atomic per column per CTA
  __shared__ float sq[128], sk[128];
  if (threadIdx.x < 128) { sq[threadIdx.x] = 0.f; sk[threadIdx.x] = 0.f; }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&sq[hl * 8 + j], accq[j]);
    atomicAdd(&sk[hl * 8 + j], acck[j]);
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    atomicAdd(a.dwq + threadIdx.x, sq[threadIdx.x]);
    atomicAdd(a.dwk + threadIdx.x, sk[threadIdx.x]);
  }
}

// ------------------------------------------------------------------------------------
// SwiGLU: m = silu(g) * u ;  dg = dm * u * sig(g) * (1 + g * (1 - sig(g))), du = dm * silu(g)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) swiglu_fwd_kernel(const __nv_bfloat16* __restrict__ g,
                                                         const __nv_bfloat16* __restrict__ u, int64_t n_vec,
                                                         __nv_bfloat16* __restrict__ m) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n_vec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float gv[8], uv[8];
    ld8(g + i * 8, gv);
    ld8(u + i * 8, uv);
#pragma unroll
    for (int j = 0; j < 8; ++j) gv[j] = gv[j] / (1.f + __expf(-gv[j])) * uv[j];
    st8(m + i * 8, gv);
  }
}
__global__ void __launch_bounds__(256) swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ dm,
                                                         const __nv_bfloat16* __restrict__ g,
                                                         const __nv_bfloat16* __restrict__ u, int64_t n_vec,
                                                         __nv_bfloat16* __restrict__ dg,
                                                         __nv_bfloat16* __restrict__ du) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n_vec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float dv[8], gv[8], uv[8], og[8], ou[8];
    ld8(dm + i * 8, dv);
    ld8(g + i * 8, gv);
    ld8(u + i * 8, uv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sg = 1.f / (1.f + __expf(-gv[j]));
      ou[j] = dv[j] * gv[j] * sg;
      og[j] = dv[j] * uv[j] * sg * (1.f + gv[j] * (1.f - sg));
    }
    st8(dg + i * 8, og);
    st8(du + i * 8, ou);
  }
}

static int elem_grid(int64_t n, int threads) {
  int64_t g = (n + threads - 1) / threads;
  const int64_t cap = 16 * static_cast<int64_t>(sm_count());
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace dta

namespace dta {
// measured at the bench shape (tools/layer_bench.py, round 2): forward 0.230 / 0.206 / 0.224 ms at 2 / 3 / 4 blocks per SM
// (80 registers at 3: the loads of more warps in flight win over 12 bytes of spills); backward 0.299 / 0.398 / 0.494 (spills)
static constexpr int kDefaultQkFwdOcc = 3, kDefaultQkBwdOcc = 2;
static int qk_occupancy(int dflt) {
  const char* e = getenv("DTA_QK_OCC");
  return e ? atoi(e) : dflt;
}
}  // namespace dta

using namespace dta;

extern "C" {

int dta_rmsnorm_fwd(const void* x, const void* residual, const void* weight, int64_t n_rows, int hidden, float eps,
                    void* y, void* sum_out, float* rstd, dta_stream_t stream) {
  DTA_REQUIRE(hidden % 8 == 0 && hidden >= 256 && hidden <= 8192 && (hidden / 8) % 32 == 0,
              "hidden must be a multiple of 256 in [256, 8192] (got %d)", hidden);
  if (n_rows == 0) return DTA_OK;
  DTA_REQUIRE(x && weight && y && rstd && (residual == nullptr || sum_out != nullptr), "NULL pointer argument");
  DTA_REQUIRE(aligned16(x) && aligned16(weight) && aligned16(y) && aligned16(residual) && aligned16(sum_out),
              "tensors must be 16-byte aligned");
  const int64_t cap = 8 * static_cast<int64_t>(sm_count());
  rmsnorm_fwd_kernel<<<static_cast<int>(n_rows < cap ? n_rows : cap), hidden / 8, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(residual),
      static_cast<const __nv_bfloat16*>(weight), n_rows, hidden, eps, static_cast<__nv_bfloat16*>(y),
      static_cast<__nv_bfloat16*>(sum_out), rstd);
  DTA_CHECK_CUDA(cudaGetLastError());
  return DTA_OK;
}

int dta_rmsnorm_bwd(const void* dy, const void* x, const void* weight, const float* rstd, const void* dresidual,
                    int64_t n_rows, int hidden, void* dx, float* dweight, dta_stream_t stream) {
  DTA_REQUIRE(hidden % 8 == 0 && hidden >= 256 && hidden <= 8192 && (hidden / 8) % 32 == 0,
              "hidden must be a multiple of 256 in [256, 8192] (got %d)", hidden);
  if (n_rows == 0) return DTA_OK;
  DTA_REQUIRE(dy && x && weight && rstd && dx && dweight, "NULL pointer argument");
  DTA_REQUIRE(aligned16(dy) && aligned16(x) && aligned16(weight) && aligned16(dx) && aligned16(dresidual),
              "tensors must be 16-byte aligned");
  const __nv_bfloat16* dy_ = static_cast<const __nv_bfloat16*>(dy);
  const __nv_bfloat16* x_ = static_cast<const __nv_bfloat16*>(x);
  const __nv_bfloat16* w_ = static_cast<const __nv_bfloat16*>(weight);
  const __nv_bfloat16* dr_ = static_cast<const __nv_bfloat16*>(dresidual);
  __nv_bfloat16* dx_ = static_cast<__nv_bfloat16*>(dx);
  if (hidden <= 1024) {
    int64_t grid = (n_rows + 7) / 8;                      // 8 warps = 8 rows per CTA iteration
    const int64_t cap2 = 2 * static_cast<int64_t>(sm_count());  // 2 resident CTAs per SM (128 registers)
    if (grid > cap2) grid = cap2;
    const int g = static_cast<int>(grid);
    switch (hidden / 256) {
      case 1: rmsnorm_bwd_warp_kernel<1><<<g, 256, 0, stream>>>(dy_, x_, w_, rstd, dr_, n_rows, dx_, dweight); break;
      case 2: rmsnorm_bwd_warp_kernel<2><<<g, 256, 0, stream>>>(dy_, x_, w_, rstd, dr_, n_rows, dx_, dweight); break;
      case 3: rmsnorm_bwd_warp_kernel<3><<<g, 256, 0, stream>>>(dy_, x_, w_, rstd, dr_, n_rows, dx_, dweight); break;
      default: rmsnorm_bwd_warp_kernel<4><<<g, 256, 0, stream>>>(dy_, x_, w_, rstd, dr_, n_rows, dx_, dweight); break;
    }
  } else {
    const int64_t cap = 8 * static_cast<int64_t>(sm_count());
    rmsnorm_bwd_kernel<<<static_cast<int>(n_rows < cap ? n_rows : cap), hidden / 8, 0, stream>>>(dy_, x_, w_, rstd, dr_,
                                                                                                  n_rows, hidden, dx_, dweight);
  }
  DTA_CHECK_CUDA(cudaGetLastError());
  return DTA_OK;
}

int dta_qknorm_rope_fwd(const void* q, int64_t q_row_stride, int64_t q_head_stride, const void* k,
                        int64_t k_row_stride, int64_t k_head_stride, const void* wq, const void* wk,
                        const int32_t* depth, const float* cos_sin, int64_t n_rows, int n_q_heads, int n_kv_heads,
                        int head_dim, float eps, void* q_out, void* k_out, float* rstd, dta_stream_t stream) {
  DTA_REQUIRE(head_dim == 128, "head_dim must be 128 (got %d)", head_dim);
  if (n_rows == 0) return DTA_OK;
  DTA_REQUIRE(q && k && wq && wk && depth && cos_sin && q_out && k_out && rstd, "NULL pointer argument");
  DTA_REQUIRE(q_row_stride % 4 == 0 && q_head_stride % 4 == 0 && k_row_stride % 4 == 0 && k_head_stride % 4 == 0 &&
                  (reinterpret_cast<uintptr_t>(q) & 7) == 0 && (reinterpret_cast<uintptr_t>(k) & 7) == 0,
              "q/k must be 8-byte aligned with strides that are multiples of 4 elements");
  QkArgs a{static_cast<const __nv_bfloat16*>(q), q_row_stride, q_head_stride,
           static_cast<const __nv_bfloat16*>(k), k_row_stride, k_head_stride,
           static_cast<__nv_bfloat16*>(q_out), static_cast<__nv_bfloat16*>(k_out),
           static_cast<const __nv_bfloat16*>(wq), static_cast<const __nv_bfloat16*>(wk), depth, cos_sin, rstd,
           n_rows, n_q_heads, n_kv_heads, eps};
  const bool v8 = q_row_stride % 8 == 0 && q_head_stride % 8 == 0 && k_row_stride % 8 == 0 && k_head_stride % 8 == 0 &&
                  aligned16(q) && aligned16(k) && aligned16(q_out) && aligned16(k_out) && aligned16(wq) && aligned16(wk) &&
                  aligned16(cos_sin);
  if (v8) {
    const int occ = qk_occupancy(kDefaultQkFwdOcc);
    if (occ >= 4) qknorm_rope_fwd_v8_kernel<4><<<elem_grid(n_rows * 32, 256), 256, 0, stream>>>(a);
    else if (occ == 3) qknorm_rope_fwd_v8_kernel<3><<<elem_grid(n_rows * 32, 256), 256, 0, stream>>>(a);
    else qknorm_rope_fwd_v8_kernel<2><<<elem_grid(n_rows * 32, 256), 256, 0, stream>>>(a);
  } else
    qknorm_rope_fwd_kernel<<<elem_grid(n_rows * (n_q_heads + n_kv_heads) * 32, 256), 256, 0, stream>>>(a);
  DTA_CHECK_CUDA(cudaGetLastError());
  return DTA_OK;
}

static int launch_qknorm_rope_bwd(const void* dq_out, const float* dq_out_f32, const void* dk_out, const void* q,
                                  int64_t q_row_stride, int64_t q_head_stride, const void* k, int64_t k_row_stride,
                                  int64_t k_head_stride, const void* wq, const void* wk, const int32_t* depth,
                                  const float* cos_sin, const float* rstd, int64_t n_rows, int n_q_heads, int n_kv_heads,
                                  int head_dim, void* dq, int64_t dq_row_stride, int64_t dq_head_stride, void* dk,
                                  int64_t dk_row_stride, int64_t dk_head_stride, float* dwq, float* dwk, dta_stream_t stream) {
  DTA_REQUIRE(head_dim == 128, "head_dim must be 128 (got %d)", head_dim);
  if (n_rows == 0) return DTA_OK;
  DTA_REQUIRE((dq_out || dq_out_f32) && dk_out && q && k && wq && wk && depth && cos_sin && rstd && dq && dk && dwq && dwk,
              "NULL pointer argument");
  QkBwdArgs a{static_cast<const __nv_bfloat16*>(dq_out), static_cast<const __nv_bfloat16*>(dk_out), dq_out_f32,
              static_cast<const __nv_bfloat16*>(q), q_row_stride, q_head_stride,
              static_cast<const __nv_bfloat16*>(k), k_row_stride, k_head_stride,
              static_cast<__nv_bfloat16*>(dq), dq_row_stride, dq_head_stride,
              static_cast<__nv_bfloat16*>(dk), dk_row_stride, dk_head_stride,
              static_cast<const __nv_bfloat16*>(wq), static_cast<const __nv_bfloat16*>(wk), depth, cos_sin, rstd,
              dwq, dwk, n_rows, n_q_heads, n_kv_heads};
  const bool v8 = q_row_stride % 8 == 0 && q_head_stride % 8 == 0 && k_row_stride % 8 == 0 && k_head_stride % 8 == 0 &&
                  dq_row_stride % 8 == 0 && dq_head_stride % 8 == 0 && dk_row_stride % 8 == 0 && dk_head_stride % 8 == 0 &&
                  aligned16(q) && aligned16(k) && aligned16(dq) && aligned16(dk) && (dq_out_f32 ? aligned16(dq_out_f32) : aligned16(dq_out)) &&
                  aligned16(dk_out) && aligned16(wq) && aligned16(wk) && aligned16(cos_sin);
  DTA_REQUIRE(v8 || !dq_out_f32, "the fp32-dq variant needs 16-byte aligned tensors and strides that are multiples of 8");
  const int64_t cap = 8 * static_cast<int64_t>(sm_count());
  int64_t grid = ((v8 ? n_rows : n_rows * (n_q_heads + n_kv_heads)) * 32 + 255) / 256;
  if (grid > cap) grid = cap;
  if (v8 && dq_out_f32) {
    qknorm_rope_bwd_v8_kernel<2, true, 2><<<static_cast<int>(grid), 256, 0, stream>>>(a);
  } else if (v8) {
    const int occ = qk_occupancy(kDefaultQkBwdOcc);
    if (occ >= 4) qknorm_rope_bwd_v8_kernel<4, false, kQkUnroll><<<static_cast<int>(grid), 256, 0, stream>>>(a);
    else if (occ == 3) qknorm_rope_bwd_v8_kernel<3, false, kQkUnroll><<<static_cast<int>(grid), 256, 0, stream>>>(a);
    else qknorm_rope_bwd_v8_kernel<2, false, kQkUnroll><<<static_cast<int>(grid), 256, 0, stream>>>(a);
  } else {
    qknorm_rope_bwd_kernel<<<static_cast<int>(grid), 256, 0, stream>>>(a);
  }
  DTA_CHECK_CUDA(cudaGetLastError());
  return DTA_OK;
}

int dta_qknorm_rope_bwd(const void* dq_out, const void* dk_out, const void* q, int64_t q_row_stride,
                        int64_t q_head_stride, const void* k, int64_t k_row_stride, int64_t k_head_stride,
                        const void* wq, const void* wk, const int32_t* depth, const float* cos_sin, const float* rstd,
                        int64_t n_rows, int n_q_heads, int n_kv_heads, int head_dim, void* dq, int64_t dq_row_stride,
                        int64_t dq_head_stride, void* dk, int64_t dk_row_stride, int64_t dk_head_stride, float* dwq,
                        float* dwk, dta_stream_t stream) {
  return launch_qknorm_rope_bwd(dq_out, nullptr, dk_out, q, q_row_stride, q_head_stride, k, k_row_stride, k_head_stride, wq, wk,
                                depth, cos_sin, rstd, n_rows, n_q_heads, n_kv_heads, head_dim, dq, dq_row_stride, dq_head_stride,
                                dk, dk_row_stride, dk_head_stride, dwq, dwk, stream);
}

int dta_qknorm_rope_bwd_f32dq(const float* dq_out_f32, const void* dk_out, const void* q, int64_t q_row_stride,
                              int64_t q_head_stride, const void* k, int64_t k_row_stride, int64_t k_head_stride,
                              const void* wq, const void* wk, const int32_t* depth, const float* cos_sin, const float* rstd,
                              int64_t n_rows, int n_q_heads, int n_kv_heads, int head_dim, void* dq, int64_t dq_row_stride,
                              int64_t dq_head_stride, void* dk, int64_t dk_row_stride, int64_t dk_head_stride, float* dwq,
                              float* dwk, dta_stream_t stream) {
  DTA_REQUIRE(dq_out_f32 != nullptr, "NULL fp32 dq");
  return launch_qknorm_rope_bwd(nullptr, dq_out_f32, dk_out, q, q_row_stride, q_head_stride, k, k_row_stride, k_head_stride, wq,
                                wk, depth, cos_sin, rstd, n_rows, n_q_heads, n_kv_heads, head_dim, dq, dq_row_stride,
                                dq_head_stride, dk, dk_row_stride, dk_head_stride, dwq, dwk, stream);
}

int dta_swiglu_fwd(const void* gate, const void* up, int64_t n_elems, void* out, dta_stream_t stream) {
  DTA_REQUIRE(n_elems % 8 == 0, "element count must be a multiple of 8");
  if (n_elems == 0) return DTA_OK;
  DTA_REQUIRE(gate && up && out && aligned16(gate) && aligned16(up) && aligned16(out), "NULL or misaligned pointer");
  swiglu_fwd_kernel<<<elem_grid(n_elems / 8, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(gate),
                                                                    static_cast<const __nv_bfloat16*>(up),
                                                                    n_elems / 8, static_cast<__nv_bfloat16*>(out));
  DTA_CHECK_CUDA(cudaGetLastError());
  return DTA_OK;
}

int dta_swiglu_bwd(const void* dout, const void* gate, const void* up, int64_t n_elems, void* dgate, void* dup,
                   dta_stream_t stream) {
  DTA_REQUIRE(n_elems % 8 == 0, "element count must be a multiple of 8");
  if (n_elems == 0) return DTA_OK;
  DTA_REQUIRE(dout && gate && up && dgate && dup && aligned16(dout) && aligned16(gate) && aligned16(up) &&
                  aligned16(dgate) && aligned16(dup),
              "NULL or misaligned pointer");
  swiglu_bwd_kernel<<<elem_grid(n_elems / 8, 256), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(dout), static_cast<const __nv_bfloat16*>(gate),
      static_cast<const __nv_bfloat16*>(up), n_elems / 8, static_cast<__nv_bfloat16*>(dgate),
      static_cast<__nv_bfloat16*>(dup));
  DTA_CHECK_CUDA(cudaGetLastError());
  return DTA_OK;
}

}  // extern "C"
