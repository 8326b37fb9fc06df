// Non-GEMM kernels of the TokenPacker path: point-query stencil, local-window attention, weight packing,
// and the HD front end (tiling, separator rows).  All HBM-bound: 16-byte vector accesses along the channel dim.
#pragma once

#include "tp_ptx.cuh"

namespace tp {

constexpr int kGrid = 24;       // builder.py:42 raw_grid
constexpr int kTokens = 576;    // 24*24
constexpr int kC = 1024;        // embed_dim / kv_dim (builder.py:43,45)
constexpr int kCm = 4096;       // multi-level stack width (builder.py:61,67)
constexpr int kHeads = 8;       // builder.py:44
constexpr int kHeadDim = 128;

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}

__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// ------------------------------------------------------------------------------------------------
// Point queries (builder.py:117-118): bilinear 24x24 -> g x g with align_corners=False is a fixed stencil per
// s x s window: the source coordinate of output i is s*i + s/2 - 0.5, so an odd s reads the window's centre token
// (weight exactly 1) and an even s the mean of its centre 2x2 (weights exactly 0.5 each) — s=2: mean of the 2x2;
// s=3: the centre; s=4: mean of the centre 2x2; s=1: the token itself.  Computed in fp32 (the reference upcasts
// with .float()) and rounded once to bf16 (.to(x.dtype)).
// One thread per 8 channels of one query.
// ------------------------------------------------------------------------------------------------
template <int S>
__global__ void point_query_kernel(const __nv_bfloat16* __restrict__ x0, long long crop_stride, __nv_bfloat16* __restrict__ q,
                                   long long n_queries) {
  constexpr int G = kGrid / S;
  constexpr int M = G * G;
  grid_dependency_wait();        // PDL: inputs may come from the previous kernel on the stream
  grid_launch_dependents();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long query = idx >> 7;          // 128 vectors of 8 channels per query
  const int vec = static_cast<int>(idx & 127);
  if (query >= n_queries) return;
  const long long n = query / M;
  const int m = static_cast<int>(query - n * M);
  const int hb = m / G, wb = m - hb * G;
  const __nv_bfloat16* base = x0 + n * crop_stride + vec * 8;
  auto tok = [&](int r, int c) { return __ldg(reinterpret_cast<const uint4*>(base + static_cast<long long>(r * kGrid + c) * kC)); };
  uint4 out;
  if (S % 2 == 1) {
    out = tok(hb * S + (S - 1) / 2, wb * S + (S - 1) / 2);
  } else {
    const int r0 = hb * S + S / 2 - 1;
    const int c0 = wb * S + S / 2 - 1;
    float a[8], b[8], c[8], d[8], o[8];
    unpack8(tok(r0, c0), a);
    unpack8(tok(r0, c0 + 1), b);
    unpack8(tok(r0 + 1, c0), c);
    unpack8(tok(r0 + 1, c0 + 1), d);
    // 0.5*(0.5a+0.5b) + 0.5*(0.5c+0.5d): scaling by powers of two is exact, so this association is bit-identical
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = 0.25f * ((a[i] + b[i]) + (c[i] + d[i]));
    out = pack8(o);
  }
  *reinterpret_cast<uint4*>(q + query * kC + vec * 8) = out;
}

// ------------------------------------------------------------------------------------------------
// Local-window cross attention core (builder.py:122-130 == nn.MultiheadAttention with L=1, S=s*s):
//   per query (n, hb, wb) and head h:  p = softmax_j( q'_h . k'_{j,h} ),  ctx_h = sum_j p_j v'_{j,h}
// q' is already scaled by 1/sqrt(128) (fused into the in_proj_q GEMM epilogue).  The window gather
// (divide_feature, builder.py:96-105) is pure address arithmetic here: fine token (hb*s+hi, wb*s+wi).
// One warp per query; lane l owns channels {256*i + 8*l .. +7 : i=0..3}; channel block i of lanes 0-15 is head 2i,
// of lanes 16-31 head 2i+1, so a head's dot product is a 16-lane shuffle reduction.
// ------------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(256) window_attn_kernel(const __nv_bfloat16* __restrict__ qp, const __nv_bfloat16* __restrict__ kp,
                                                          const __nv_bfloat16* __restrict__ vp, __nv_bfloat16* __restrict__ ctx,
                                                          long long n_queries) {
  constexpr int G = kGrid / S;
  constexpr int M = G * G;
  constexpr int W = S * S;
  grid_dependency_wait();        // PDL: q', k', v' come from the previous GEMM launch
  grid_launch_dependents();
  const long long query = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (query >= n_queries) return;
  const long long n = query / M;
  const int m = static_cast<int>(query - n * M);
  const int hb = m / G, wb = m - hb * G;
  const long long tok0 = n * kTokens + static_cast<long long>(hb * S) * kGrid + wb * S;

  float qf[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i) unpack8(__ldg(reinterpret_cast<const uint4*>(qp + query * kC + i * 256 + lane * 8)), qf[i]);

  float sc[4][W];
#pragma unroll
  for (int j = 0; j < W; ++j) {
    const long long tok = tok0 + (j / S) * kGrid + (j % S);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float kf[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(kp + tok * kC + i * 256 + lane * 8)), kf);
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) d = fmaf(qf[i][e], kf[e], d);
      sc[i][j] = d;
    }
  }
  // 16-lane reductions (lanes 0-15 and 16-31 hold different heads)
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < W; ++j) {
      float d = sc[i][j];
      d += __shfl_xor_sync(0xffffffffu, d, 8);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      sc[i][j] = d;
    }
  // softmax over the W keys, fp32
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float mx = sc[i][0];
#pragma unroll
    for (int j = 1; j < W; ++j) mx = fmaxf(mx, sc[i][j]);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      sc[i][j] = __expf(sc[i][j] - mx);
      sum += sc[i][j];
    }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int j = 0; j < W; ++j) sc[i][j] *= inv;
  }
  float of[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) of[i][e] = 0.f;
#pragma unroll
  for (int j = 0; j < W; ++j) {
    const long long tok = tok0 + (j / S) * kGrid + (j % S);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float vf[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(vp + tok * kC + i * 256 + lane * 8)), vf);
#pragma unroll
      for (int e = 0; e < 8; ++e) of[i][e] = fmaf(sc[i][j], vf[e], of[i][e]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(ctx + query * kC + i * 256 + lane * 8) = pack8(of[i]);
}

// Same computation for any window size (scale_factor 1, 6, 8, 12, 24: 1 ... 576 keys per query — constructor-valid upstream,
// builder.py:51-52, though no released model uses them): the keys are streamed through an online softmax instead of being
// held in registers.  Same warp / lane ownership as window_attn_kernel.
__global__ void __launch_bounds__(256) window_attn_stream_kernel(const __nv_bfloat16* __restrict__ qp, const __nv_bfloat16* __restrict__ kp,
                                                                 const __nv_bfloat16* __restrict__ vp, __nv_bfloat16* __restrict__ ctx,
                                                                 long long n_queries, int s) {
  const int G = kGrid / s;
  const int M = G * G;
  const int W = s * s;
  grid_dependency_wait();
  grid_launch_dependents();
  const long long query = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (query >= n_queries) return;
  const long long n = query / M;
  const int m = static_cast<int>(query - n * M);
  const int hb = m / G, wb = m - hb * G;
  const long long tok0 = n * kTokens + static_cast<long long>(hb * s) * kGrid + wb * s;

  float qf[4][8], of[4][8], mx[4], den[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    unpack8(__ldg(reinterpret_cast<const uint4*>(qp + query * kC + i * 256 + lane * 8)), qf[i]);
    mx[i] = -INFINITY;
    den[i] = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) of[i][e] = 0.f;
  }
  for (int j = 0; j < W; ++j) {
    const int hi = j / s;
    const long long tok = tok0 + static_cast<long long>(hi) * kGrid + (j - hi * s);
    float sc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float kf[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(kp + tok * kC + i * 256 + lane * 8)), kf);
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) d = fmaf(qf[i][e], kf[e], d);
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
      sc[i] = d;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float m_new = fmaxf(mx[i], sc[i]);
      const float corr = __expf(mx[i] - m_new);       // 0 on the first key (mx = -inf)
      const float pj = __expf(sc[i] - m_new);
      den[i] = fmaf(den[i], corr, pj);
      mx[i] = m_new;
      float vf[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(vp + tok * kC + i * 256 + lane * 8)), vf);
#pragma unroll
      for (int e = 0; e < 8; ++e) of[i][e] = fmaf(pj, vf[e], of[i][e] * corr);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float inv = 1.0f / den[i];
#pragma unroll
    for (int e = 0; e < 8; ++e) of[i][e] *= inv;
    *reinterpret_cast<uint4*>(ctx + query * kC + i * 256 + lane * 8) = pack8(of[i]);
  }
}

// ------------------------------------------------------------------------------------------------
// Weight packing
// ------------------------------------------------------------------------------------------------
__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __bfloat162float(src[i]);
}

// Several bf16 -> fp32 vectors in one launch (the biases of a training-step repack): blockIdx.y = segment
struct CastSeg { const __nv_bfloat16* src; float* dst; int n; };
struct CastSegs { CastSeg s[8]; };
__global__ void bf16_to_f32_multi_kernel(CastSegs segs) {
  const CastSeg g = segs.s[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < g.n; i += gridDim.x * blockDim.x) g.dst[i] = __bfloat162float(g.src[i]);
}

// LayerNorm folded into the following linear (one warp per output feature o):
//   LN(y) W^T + b = rstd * ( y (gamma.W)^T - mu * rowsum(gamma.W) ) + ( W beta + b )
//   w_out[o,:] = bf16(W[o,:] * gamma),  wsum[o] = sum_i w_out[o,i] (of the ROUNDED values),  cst[o] = W[o,:].beta + b[o]
__global__ void fold_layernorm_kernel(const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* __restrict__ bias,
                                      const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta,
                                      __nv_bfloat16* __restrict__ w_out, float* __restrict__ wsum, float* __restrict__ cst,
                                      int out_dim, int in_dim) {
  const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (o >= out_dim) return;
  float s = 0.f, c = 0.f;
  for (int i = lane; i < in_dim; i += 32) {
    const float wv = __bfloat162float(w[static_cast<long long>(o) * in_dim + i]);
    const __nv_bfloat16 folded = __float2bfloat16_rn(wv * __bfloat162float(gamma[i]));
    w_out[static_cast<long long>(o) * in_dim + i] = folded;
    s += __bfloat162float(folded);
    c = fmaf(wv, __bfloat162float(beta[i]), c);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, off);
    c += __shfl_xor_sync(0xffffffffu, c, off);
  }
  if (lane == 0) {
    wsum[o] = s;
    cst[o] = c + __bfloat162float(bias[o]);
  }
}

// out[c, r] = in[r, c] for a square bf16 matrix (pack time only: W_o^T as the K-major B operand of the fold GEMM)
__global__ void transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int n) {
  __shared__ __nv_bfloat16 tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) tile[i][threadIdx.x] = in[static_cast<long long>(by + i) * n + bx + threadIdx.x];
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) out[static_cast<long long>(bx + i) * n + by + threadIdx.x] = tile[threadIdx.x][i];
}

// out[o] = sum_i w[o, i] * x[i] + b[o]   (fp32 out; one warp per output; pack time only)
__global__ void matvec_bias_kernel(const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* __restrict__ x,
                                   const __nv_bfloat16* __restrict__ b, float* __restrict__ out, int out_dim, int in_dim) {
  const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (o >= out_dim) return;
  float acc = 0.f;
  for (int i = lane; i < in_dim; i += 32) acc = fmaf(__bfloat162float(w[static_cast<long long>(o) * in_dim + i]), __bfloat162float(x[i]), acc);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) out[o] = acc + __bfloat162float(b[o]);
}

// ------------------------------------------------------------------------------------------------
// HD front end
// ------------------------------------------------------------------------------------------------
constexpr int kBlockPx = 336;   // train.py:699 block_size

struct LinearTap {
  int i0, i1;
  float w0, w1;
};

// ATen upsample_bilinear2d, align_corners=False, scales derived from sizes: src = (in/out)*(dst+0.5)-0.5 clamped at 0.
__device__ __forceinline__ LinearTap linear_tap_scaled(int dst, int in_size, float scale);

__device__ __forceinline__ LinearTap linear_tap(int dst, int in_size, int out_size) {
  return linear_tap_scaled(dst, in_size, static_cast<float>(in_size) / static_cast<float>(out_size));
}

// same with the scale (in_size / out_size as one IEEE float division) supplied by the caller
__device__ __forceinline__ LinearTap linear_tap_scaled(int dst, int in_size, float scale) {
  float src = __fmaf_rn(scale, static_cast<float>(dst) + 0.5f, -0.5f);   // ONE fused multiply-add, like ATen's CPU kernel (see oracle/hd_oracle.py)
  src = fmaxf(src, 0.f);
  LinearTap t;
  t.i0 = min(static_cast<int>(src), in_size - 1);
  t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
  t.w1 = fminf(fmaxf(src - static_cast<float>(t.i0), 0.f), 1.f);
  t.w0 = 1.f - t.w1;
  return t;
}

__device__ __forceinline__ float bilerp(float a, float b, float c, float d, const LinearTap& ty, const LinearTap& tx) {
  const float top = __fadd_rn(__fmul_rn(tx.w0, a), __fmul_rn(tx.w1, b));
  const float bot = __fadd_rn(__fmul_rn(tx.w0, c), __fmul_rn(tx.w1, d));
  return __fadd_rn(__fmul_rn(ty.w0, top), __fmul_rn(ty.w1, bot));
}

// Pass 1 (train.py:709-717): crops[(i*wb+j), ch, y, x] = canvas[ch, 336 i + y, 336 j + x], canvas = zero-padded
// bilinear resize of image[3,h,w] to (h_r, w_r).
__global__ void hd_tile_main_kernel(const float* __restrict__ image, int h, int w, int hb, int wb, int h_r, int w_r,
                                    float* __restrict__ crops) {
  const long long total = 3ll * hb * kBlockPx * wb * kBlockPx;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cw = wb * kBlockPx, chh = hb * kBlockPx;
  const int X = static_cast<int>(idx % cw);
  const int Y = static_cast<int>((idx / cw) % chh);
  const int ch = static_cast<int>(idx / (static_cast<long long>(cw) * chh));
  float v = 0.f;
  if (Y < h_r && X < w_r) {
    const LinearTap ty = linear_tap(Y, h, h_r), tx = linear_tap(X, w, w_r);
    const float* p = image + static_cast<long long>(ch) * h * w;
    v = bilerp(p[static_cast<long long>(ty.i0) * w + tx.i0], p[static_cast<long long>(ty.i0) * w + tx.i1],
               p[static_cast<long long>(ty.i1) * w + tx.i0], p[static_cast<long long>(ty.i1) * w + tx.i1], ty, tx);
  }
  const int ci = Y / kBlockPx, cj = X / kBlockPx;
  const int y = Y - ci * kBlockPx, x = X - cj * kBlockPx;
  crops[((static_cast<long long>(ci * wb + cj) * 3 + ch) * kBlockPx + y) * kBlockPx + x] = v;
}

// Pass 2 (train.py:718-730): thumbnail = zero-padded bilinear resize of the PADDED canvas (read back from the crop
// layout written by pass 1) to (h_t, w_t); stored as crop index hb*wb.
__global__ void hd_tile_thumb_kernel(int hb, int wb, int h_t, int w_t, float* __restrict__ crops) {
  const int total = 3 * kBlockPx * kBlockPx;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = idx % kBlockPx;
  const int y = (idx / kBlockPx) % kBlockPx;
  const int ch = idx / (kBlockPx * kBlockPx);
  float v = 0.f;
  if (y < h_t && x < w_t) {
    const LinearTap ty = linear_tap(y, hb * kBlockPx, h_t), tx = linear_tap(x, wb * kBlockPx, w_t);
    auto canvas = [&](int Y, int X) {
      const int ci = Y / kBlockPx, cj = X / kBlockPx;
      return crops[((static_cast<long long>(ci * wb + cj) * 3 + ch) * kBlockPx + (Y - ci * kBlockPx)) * kBlockPx + (X - cj * kBlockPx)];
    };
    v = bilerp(canvas(ty.i0, tx.i0), canvas(ty.i0, tx.i1), canvas(ty.i1, tx.i0), canvas(ty.i1, tx.i1), ty, tx);
  }
  crops[((static_cast<long long>(hb * wb) * 3 + ch) * kBlockPx + y) * kBlockPx + x] = v;
}

// ------------------------------------------------------------------------------------------------
// Batched form of the tiling block: ONE launch tiles a whole batch of variable-size images (the collator concatenates the
// crops of a batch, train.py:797-800), thumbnails included, float4 stores.  Every thread produces 4 consecutive pixels of one
// crop row.  The thumbnail is resized from the PADDED canvas (train.py:710,718-730); instead of reading pass-1 output back, each
// canvas tap is recomputed from the source image with exactly the arithmetic the main crops use, so the bits are those of
// the two-pass kernels above.
// ------------------------------------------------------------------------------------------------
struct HdImage {            // mirrors tp_hd_image (include/tokenpacker_b200.h)
  const float* image;       // [3, h, w] fp32, normalised
  int h, w, hb, wb;
  int h_r, w_r;             // resized content of the main canvas
  int h_t, w_t;             // resized content of the thumbnail (0 when hb*wb == 1)
  long long crop0;          // index of this image's first crop in the batch output
  float sy, sx;             // h / h_r, w / w_r as float divisions (ATen's scale), computed once per image on the host
  float ty, tx;             // (336 hb) / h_t, (336 wb) / w_t
};

__device__ __forceinline__ float hd_canvas_value(const HdImage& im, const float* __restrict__ plane, int Y, int X) {
  if (Y >= im.h_r || X >= im.w_r) return 0.f;
  const LinearTap ty = linear_tap_scaled(Y, im.h, im.sy), tx = linear_tap_scaled(X, im.w, im.sx);
  const float* r0 = plane + static_cast<long long>(ty.i0) * im.w;
  const float* r1 = plane + static_cast<long long>(ty.i1) * im.w;
  return bilerp(__ldg(r0 + tx.i0), __ldg(r0 + tx.i1), __ldg(r1 + tx.i0), __ldg(r1 + tx.i1), ty, tx);
}

// crop_table[c] = (image index, grid row, grid column); grid column -1 marks the image's thumbnail.
// One CTA = kHdRows rows of one crop; thread = one pixel COLUMN: its column tap is computed once and reused by the kHdRows x 3 (row,
// channel) pairs, and the lanes of a warp read neighbouring source pixels (a warp's load touches 4-6 sectors; with 4 pixels per
// thread it was 16 sectors of which 6.5 bytes each were used, and the kernel sat at 18 % of DRAM waiting for L1).  Scalar stores of
// 32 consecutive floats per warp.  The tap arithmetic is unchanged, so the bits are.
#ifndef TP_HD_ROWS
#define TP_HD_ROWS 8
#endif
#ifndef TP_HD_UNROLL
#define TP_HD_UNROLL 2
#endif
constexpr int kHdRows = TP_HD_ROWS;                            // rows per thread; 336 = 42 x 8
constexpr int kHdUnroll = TP_HD_UNROLL;                        // rows in flight per thread (loads of the next row issue under the math of this one)
static_assert(kBlockPx % kHdRows == 0, "rows per CTA must divide the crop height");
__global__ void __launch_bounds__(kBlockPx) hd_tile_batch_kernel(const HdImage* __restrict__ images, const int* __restrict__ crop_table,
                                                                 long long n_crops, float* __restrict__ crops) {
  constexpr int kRowGroups = kBlockPx / kHdRows;
  const long long crop = blockIdx.x / kRowGroups;
  const int y0 = static_cast<int>(blockIdx.x % kRowGroups) * kHdRows;
  const int x = threadIdx.x;
  if (crop >= n_crops) return;
  const int img = crop_table[crop * 3], ci = crop_table[crop * 3 + 1], cj = crop_table[crop * 3 + 2];
  const HdImage im = images[img];
  const long long plane_sz = static_cast<long long>(im.h) * im.w;
  float* out_base = crops + (crop * 3 * kBlockPx + y0) * kBlockPx + x;            // channel stride 336 * 336, row stride 336
  if (cj >= 0) {
    const int X = cj * kBlockPx + x;
    const bool okx = X < im.w_r;
    const LinearTap tx = linear_tap_scaled(okx ? X : 0, im.w, im.sx);
#pragma unroll kHdUnroll
    for (int r = 0; r < kHdRows; ++r) {
      const int Y = ci * kBlockPx + y0 + r;
      const bool ok = okx && Y < im.h_r;
      const LinearTap ty = linear_tap_scaled(ok ? Y : 0, im.h, im.sy);
      const float* r0 = im.image + static_cast<long long>(ty.i0) * im.w;
      const float* r1 = im.image + static_cast<long long>(ty.i1) * im.w;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        out_base[(static_cast<long long>(ch) * kBlockPx + r) * kBlockPx] =
            ok ? bilerp(__ldg(r0 + tx.i0), __ldg(r0 + tx.i1), __ldg(r1 + tx.i0), __ldg(r1 + tx.i1), ty, tx) : 0.f;
        r0 += plane_sz;
        r1 += plane_sz;
      }
    }
  } else {
    // thumbnail: resized from the PADDED canvas; a canvas tap = the main crops' arithmetic at that canvas pixel (zero outside the content)
    const int ch_h = im.hb * kBlockPx, ch_w = im.wb * kBlockPx;
    const bool okx = x < im.w_t;
    const LinearTap cx = linear_tap_scaled(okx ? x : 0, ch_w, im.tx);                 // canvas column tap ...
    const bool okx0 = cx.i0 < im.w_r, okx1 = cx.i1 < im.w_r;
    const LinearTap sx0 = linear_tap_scaled(okx0 ? cx.i0 : 0, im.w, im.sx);           // ... and the source taps of its two canvas columns
    const LinearTap sx1 = linear_tap_scaled(okx1 ? cx.i1 : 0, im.w, im.sx);
#pragma unroll kHdUnroll
    for (int r = 0; r < kHdRows; ++r) {
      const int y = y0 + r;
      const bool ok = okx && y < im.h_t;
      const LinearTap cy = linear_tap_scaled(ok ? y : 0, ch_h, im.ty);
      const bool oky0 = cy.i0 < im.h_r, oky1 = cy.i1 < im.h_r;
      const LinearTap sy0 = linear_tap_scaled(oky0 ? cy.i0 : 0, im.h, im.sy), sy1 = linear_tap_scaled(oky1 ? cy.i1 : 0, im.h, im.sy);
      const float* a0 = im.image + static_cast<long long>(sy0.i0) * im.w;   // source rows of canvas row cy.i0
      const float* a1 = im.image + static_cast<long long>(sy0.i1) * im.w;
      const float* b0 = im.image + static_cast<long long>(sy1.i0) * im.w;   // ... of canvas row cy.i1
      const float* b1 = im.image + static_cast<long long>(sy1.i1) * im.w;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        float o = 0.f;
        if (ok) {
          auto canvas = [&](const float* q0, const float* q1, const LinearTap& sy, bool rowok, const LinearTap& sx, bool colok) {
            return (rowok && colok) ? bilerp(__ldg(q0 + sx.i0), __ldg(q0 + sx.i1), __ldg(q1 + sx.i0), __ldg(q1 + sx.i1), sy, sx) : 0.f;
          };
          o = bilerp(canvas(a0, a1, sy0, oky0, sx0, okx0), canvas(a0, a1, sy0, oky0, sx1, okx1),
                     canvas(b0, b1, sy1, oky1, sx0, okx0), canvas(b0, b1, sy1, oky1, sx1, okx1), cy, cx);
        }
        out_base[(static_cast<long long>(ch) * kBlockPx + r) * kBlockPx] = o;
        a0 += plane_sz; a1 += plane_sz; b0 += plane_sz; b1 += plane_sz;
      }
    }
  }
}

// out[seg_row_offset[c] + m, :] = feats[c, m, :]  (bf16; one thread per 8 channels): crop token blocks -> packed rows.
__global__ void scatter_crops_kernel(const __nv_bfloat16* __restrict__ feats, long long n_crops, int tokens, int hidden,
                                     const long long* __restrict__ seg_row_offset, __nv_bfloat16* __restrict__ out) {
  const int vecs = hidden / 8;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long rows = n_crops * tokens;
  if (idx >= rows * vecs) return;
  const long long r = idx / vecs;
  const int v = static_cast<int>(idx - r * vecs);
  const long long c = r / tokens;
  const long long dst = seg_row_offset[c] + (r - c * tokens);
  *reinterpret_cast<uint4*>(out + dst * hidden + v * 8) = __ldg(reinterpret_cast<const uint4*>(feats + r * hidden + v * 8));
}

// Text/vision splice (llava_arch.py:119-233 as one gather): out[i,:] = table[src[i]] if src[i] >= 0, zeros if src[i] == -1,
// visual[-src[i]-2] otherwise.  bf16 rows, one thread per 8 channels.
__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ table, const __nv_bfloat16* __restrict__ visual, int hidden,
                                   const long long* __restrict__ src, long long n_rows, __nv_bfloat16* __restrict__ out) {
  const int vecs = hidden / 8;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n_rows * vecs) return;
  const long long r = idx / vecs;
  const int v = static_cast<int>(idx - r * vecs);
  const long long sidx = src[r];
  uint4 val = make_uint4(0u, 0u, 0u, 0u);
  if (sidx >= 0) val = __ldg(reinterpret_cast<const uint4*>(table + sidx * hidden + v * 8));
  else if (sidx <= -2) val = __ldg(reinterpret_cast<const uint4*>(visual + (-sidx - 2) * hidden + v * 8));
  *reinterpret_cast<uint4*>(out + r * hidden + v * 8) = val;
}

// out[rows[i], :] = row (bf16 [hidden]); one thread per 8 channels.
__global__ void fill_rows_kernel(__nv_bfloat16* __restrict__ out, int hidden, const long long* __restrict__ rows, long long n_rows,
                                 const __nv_bfloat16* __restrict__ row) {
  const int vecs = hidden / 8;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n_rows * vecs) return;
  const long long r = idx / vecs;
  const int v = static_cast<int>(idx - r * vecs);
  *reinterpret_cast<uint4*>(out + rows[r] * hidden + v * 8) = __ldg(reinterpret_cast<const uint4*>(row + v * 8));
}

}  // namespace tp
