// Persistent warp-specialised tcgen05 GEMMs for sm_100a with the fused epilogues the TokenPacker path needs.
//
//   C[M,N] (bf16) = epilogue( A[M,K] (bf16, K-major) . B[N,K]^T (bf16, K-major) ),  fp32 accumulation in TMEM.
//
// Every nn.Linear on the reference hot path (builder.py:59-83; MHA in/out projections builder.py:77) is an
// instance of these kernels: activations are [rows, in] and weights are [out, in], both K-major, which is exactly the
// operand form tcgen05.mma takes from shared memory, so no transposes exist anywhere.
//
// Two kernels share one epilogue:
//   tp_gemm_kernel<BN>   one CTA per SM, 128 x BN tiles,  tcgen05.mma cta_group::1 (small problems, BN = 128 | 256)
//   tp_gemm2_kernel      CTA pairs (cluster 2x1x1) on the two SMs of a TPC, 256 x 256 tiles, cta_group::2: each CTA
//                        stages its own 128 rows of A and HALF of the B tile, so shared-memory fill traffic per FLOP
//                        drops by a third; 4-stage mbarrier ring + double-buffered output slabs for the TMA stores.
//                        Also the home of the TN form (MN-major operands, wgrad), multi-part A (four CLIP layers side by
//                        side along K), grouped launches and the peer (all-gather) stores.
//
// CTA = 384 threads, persistent over output tiles (default role layout):
//   warps 0-7   epilogue: tcgen05.ld 32x32b (thread == output row), fused per-row / per-column math on the packed fp32
//               pipe, swizzled shared-memory slabs (direct 16-byte stores in the one-CTA kernels / arbitrary row scatter)
//   warps 8, 9  store warps of the pair kernel (one per column half): wait for a finished slab on an mbarrier, issue its TMA
//               store(s), hand the buffer back, and — for GEMMs that other GEMMs of the same launch depend on — publish each
//               finished tile to a global counter.  The epilogue warps never wait for a store.  Warp 9 also allocates TMEM
//               (2 accumulator buffers: the epilogue of tile i overlaps tile i+1's MMAs)
//   warp 10     TMA producer   (warp-uniform loop, elect.sync around the issue): cp.async.bulk.tensor boxes, 128B swizzle, mbarrier
//               ring; spins on the producer GEMM's tile counter before the first load of a dependent tile
//   warp 11     MMA issuer     (leader CTA only in pair mode): fp32 accumulators in TMEM
//
// Fused epilogue (all optional, selected at run time, warp-uniform branches):
//   v = acc
//   v = fma(rstd_r, fma(-mu_r, col_a[c], v), col_b[c])    LayerNorm folded into the NEXT linear: (mu, rstd) from per-row sums,
//                                                col_b = the folded constant W.beta + b    -- or, without a fold --
//   v = v + col_b[c]                             bias
//   v = gelu_erf(v)                              exact erf GELU (nn.GELU default)
//   v = alpha * v                                1/sqrt(head_dim) query scaling
//   y = bf16(v);  stats_out[r][slot] = (mean, M2) of y over this warp's 128 columns  -> next LayerNorm fold (deterministic:
//                                                one slot per 128-column block, combined Chan-style in fixed order by the consumer)
//   C[dst_row(r), c] = y                         optional segment scatter (HD packed output)
#pragma once

#include "tp_ptx.cuh"

namespace tp {

struct GemmEpilogue {
  __nv_bfloat16* c;        // output
  long long ldc;           // elements between output rows
  const float* col_a;      // [N]  LN fold: row-sum of the gamma-folded weight      (nullptr: no LN fold)
  const float* col_b;      // [N]  bias                                             (nullptr: none)
  const float* stats_in;   // [M, stats_in_slots, 2] per-block (mean, M2) of the A rows (required with col_a)
  float* stats_out;        // [M, stats_out_slots, 2]                               (nullptr: none)
  const long long* seg_row_offset;  // [M / seg_len] destination row of each segment's first row (nullptr: identity / uniform stride)
  int seg_len;             // rows per segment (one crop's tokens)
  int seg_stride;          // uniform-stride form (seg_row_offset == nullptr): segment i starts at output row i * seg_stride
                           // (0: contiguous).  The HD packed layout (llava_arch.py:139-155) is exactly this with stride = seg_len + 1:
                           // every crop is followed by ONE separator row (',' or '\n').  Kept on the TMA-store path (3-D C map).
  int stats_in_slots;
  int stats_out_slots;     // = N / 128 (host-checked; statistics need 256-column tiles): slot = col / 128
  float ln_inv_dim;        // 1 / ln_dim
  float ln_eps;
  float alpha;
  int gelu;
  int wm_s;                // != 0: rows are tokens in raster order (24 x 24 per crop) and C is stored WINDOW-MAJOR for scale factor
                           // wm_s: row (crop, hb, wb, hi, wi) of [crops * 576] — the s x s keys of a window become wm_s^2 consecutive
                           // rows (divide_feature, builder.py:96-105, done by the store instead of five permute copies).  The row
                           // statistics go to the permuted row too.  Pair kernel / TMA stores only; wm_s in {2, 4, 8}.
  int out_f32;             // 1: C is float [M, ldc] (split-K partial sums of the wgrads): fp32 direct stores, no bf16 rounding
  int dual;                // 1 (with gelu): the bf16-rounded PRE-activation (after bias / LN fold, before GELU and alpha) is stored too,
                           // through GemmProblem::tmap_cx[0] (the training forward keeps z for GELU'(z); builder.py:66-75 under autograd).
                           // Pair kernel, plain (unsegmented) TMA-store output, two 64-column staging buffers per half.
  long long* prof;         // TP_GEMM_PROFILE builds only: [grid][16] cycle counters (nullptr otherwise)
};

#ifndef TP_EPI_SUB_PAIRS
#define TP_EPI_SUB_PAIRS 8      // 4 / 8 / 16 measured: 8 = most interleaving that still fits the register budget without spills
#endif

#ifdef TP_GEMM_PROFILE
#define TP_PROF_T0() const long long prof_t0__ = clock64()
#define TP_PROF_ADD(var) (var) += clock64() - prof_t0__
#else
#define TP_PROF_T0() do {} while (0)
#define TP_PROF_ADD(var) do {} while (0)
#endif

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;     // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 384;
// Warp roles.  The warp scheduler favours higher warp ids when several warps of an SM sub-partition are eligible, and
// the single-lane TMA / MMA warps must never lose an issue slot to the (instruction-heavy) epilogue warps: they get
// the HIGHEST ids.  Epilogue warp w reads TMEM lanes 32*(w % 4)..+31 (hardware restriction), so 8 epilogue warps =
// 4 lane quarters x 2 column halves.
constexpr int kEpiWarp0 = 0;
constexpr int kTmaWarp = 10;
constexpr int kMmaWarp = 11;
constexpr int kAllocWarp = 9;
constexpr int kStoreWarp0 = 8;      // pair kernel: warps 8 and 9 issue the TMA stores of column half 0 / 1 (role layout 1 only)
constexpr int kNumEpiWarps = 8;
constexpr int kEpiThreads = kNumEpiWarps * 32;
constexpr int kEpiBarrierId = 1;

// ------------------------------------------------------------------------------------------------
// Shared epilogue for one 128-row x kTileN-column accumulator tile held in this CTA's TMEM.
//   tmem_acc : TMEM address of the accumulator buffer (column offset applied, lane 0)
//   row      : global output row of this thread (lane of TMEM == row inside the tile)
//   col_tile0: global column of the tile's first column
//   s_col    : shared staging for this tile's col_a / col_b slices, [2][kTileN] floats (already filled + synced)
// `release()` is invoked as soon as the last tcgen05.ld of this warp has landed in registers, so the MMA warp gets
// the TMEM buffer back before the math / stores of the final chunk.
// ------------------------------------------------------------------------------------------------
// Output staging for TMA stores: each column half of the tile (4 warps) owns two 16 KiB buffers holding a 128-row x 64-col
// slab in the 128B-swizzled layout; a slab is written with conflict-free 16-byte st.shared, then ONE thread hands it
// to the TMA unit (full 128-byte row segments instead of 16-byte scattered stores: less L1/L2 work and less power).
constexpr int kMaxPeers = 8;
// Fused all-gather: tensor maps of the SAME output slot in every peer GPU's gathered buffer (peer-mapped over NVLink);
// each finished slab is TMA-stored to all of them instead of to one local matrix.
constexpr int kBoxLevels = 4;   // exact-size store boxes: rows = unit << level
constexpr int kWholeLevels = 3; // ... and boxes of 1, 2 or 3 WHOLE segments (crops that lie entirely inside a 128-row slab)
struct PeerStores {
  CUtensorMap m[kMaxPeers][kBoxLevels + kWholeLevels];   // [destination][level]; plain (unsegmented) output uses level 0 only
  int count;                 // 0: ordinary local output through GemmProblem::tmap_c
  int whole;                 // 1: the whole-segment maps m[.][kBoxLevels + k - 1] (k segments) are valid (seg_len <= 128)
};

struct OutStage {
  uint8_t* buf;              // this half's n_bufs x 16 KiB staging buffers (nullptr: direct 16-byte global stores)
  uint64_t* full_bar;        // [n_bufs] slab written (count 4: one arrive per epilogue warp of the half) -> store warp
  uint64_t* empty_bar;       // [n_bufs] slab's TMA store has finished reading the buffer (count 1, store warp) -> epilogue warps
  int n_bufs;                // 2: slabs alternate buffers; 1: single buffer
  uint32_t slab_seq;         // running slab number of this half (buffer = seq % n_bufs, mbarrier phase = seq / n_bufs)
  bool swizzle;              // slab rows in the TMA swizzle of the row width (false: plain rows, for C maps built without swizzle)
};
// Output slab = 128 rows x kSlabCols columns of bf16, in the TMA swizzle of that row width (64 columns: 128-byte rows,
// SWIZZLE_128B; 32 columns: 64-byte rows, SWIZZLE_64B).  The narrow form halves the staging memory, which buys two more
// stages of the operand ring (-DTP_SLAB_COLS=32 -DTP_PAIR_STAGES=6): the ring depth is what hides operand-fetch latency.
#ifndef TP_SLAB_COLS
#define TP_SLAB_COLS 64
#endif
constexpr int kSlabCols = TP_SLAB_COLS;
static_assert(kSlabCols == 64 || kSlabCols == 32, "slab width: 64 (128B swizzle) or 32 (64B swizzle) columns");
constexpr int kSlabRowBytes = kSlabCols * 2;
constexpr int kChunksPerSlab = kSlabCols / 32;
constexpr int kOutSlabBytes = 128 * kSlabRowBytes;

// raster row (crop n, token row tr, token column tc) -> window-major row (crop n, window (hb, wb), key (hi, wi)) for windows of s x s
__device__ __forceinline__ long long window_major_row(long long row, int s) {
  const long long n = row / 576;
  const int t = static_cast<int>(row - n * 576);
  const int tr = t / 24, tc = t - tr * 24;
  const int hb = tr / s, hi = tr - hb * s, wb = tc / s, wi = tc - wb * s;
  return n * 576 + ((hb * (24 / s) + wb) * s + hi) * s + wi;
}

// (mean, rstd) of a LayerNorm row from its per-128-column (mean, M2) pairs, combined Chan-style in fixed order
__device__ __forceinline__ void ln_row_stats(const float* stats, long long row, int slots, float inv_dim, float eps, float& mu, float& rstd) {
  const float2* st = reinterpret_cast<const float2*>(stats) + row * slots;
  float t1 = 0.f, m2 = 0.f;
  for (int i = 0; i < slots; ++i) t1 = __fadd_rn(t1, st[i].x);
  const float inv_slots = __frcp_rn(static_cast<float>(slots));
  mu = __fmul_rn(t1, inv_slots);
  float between = 0.f;
  for (int i = 0; i < slots; ++i) {
    const float2 v = st[i];
    const float d = __fsub_rn(v.x, mu);
    between = fmaf(d, d, between);
    m2 = __fadd_rn(m2, v.y);
  }
  const float var = __fmul_rn(fmaf(between, __fmul_rn(inv_slots, __frcp_rn(inv_dim)), m2), inv_dim);
  rstd = rsqrtf(__fadd_rn(var, eps));
}

template <int kTileN, typename ReleaseFn>
__device__ __forceinline__ void epilogue_tile(const GemmEpilogue& ep, int M, int N, uint32_t tmem_acc, int row, int col_tile0,
                                              int quarter, int half, const float* s_col, const OutStage& out, ReleaseFn release,
                                              [[maybe_unused]] long long* pc = nullptr, long long c_extra = 0) {
  constexpr int kColsPerWarp = kTileN / 2;
  constexpr int kChunks = kColsPerWarp / 32;
  const bool ln_fold = ep.col_a != nullptr;
  const bool row_ok = row < M;
  float mu = 0.f, rstd = 1.f;
  // Row statistics arrive as one (mean_i, M2_i) pair per 128-column block (M2 = sum of squared deviations from the block's own
  // mean) and are combined Chan-style in a fixed order: no E[y^2] - mu^2 cancellation when |mean| >> std, bitwise reproducible.
  // (explicit intrinsics throughout the epilogue math: no FMA-contraction freedom for the compiler, so the one-CTA and CTA-pair
  // instantiations produce the same bits for the same row)
  if (ln_fold && row_ok) ln_row_stats(ep.stats_in, row, ep.stats_in_slots, ep.ln_inv_dim, ep.ln_eps, mu, rstd);
  long long dst_row = row;
  if (ep.seg_row_offset != nullptr && row_ok) {
    const int seg = row / ep.seg_len;
    dst_row = ep.seg_row_offset[seg] + (row - seg * ep.seg_len);
  } else if (ep.seg_stride != 0 && row_ok) {
    const int seg = row / ep.seg_len;
    dst_row = static_cast<long long>(seg) * ep.seg_stride + (row - seg * ep.seg_len);
  }
  __nv_bfloat16* c_row = ep.c + dst_row * ep.ldc;
  float* c_row32 = reinterpret_cast<float*>(ep.c) + c_extra + dst_row * ep.ldc;     // out_f32 only (c_extra: split-K slice)
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(half * kColsPerWarp);
  const uint32_t sa_addr = smem_u32(s_col + half * kColsPerWarp), sb_addr = smem_u32(s_col + kTileN + half * kColsPerWarp);
  const uint32_t out_addr = out.buf != nullptr ? smem_u32(out.buf) : 0u;

  float s1 = 0.f, s2 = 0.f, shift = 0.f;     // statistics of (y - shift), shift = the block's first value: sums stay small
  const bool do_stats = ep.stats_out != nullptr;
  const bool scale = ep.alpha != 1.0f;
  const uint64_t rstd2 = pk2(rstd), nmu2 = pk2(-mu), alpha2 = pk2(ep.alpha);
  uint32_t r[2][32];
  tmem_ld_32x32b_x32(taddr, r[0]);
#pragma unroll
  for (int chunk = 0; chunk < kChunks; ++chunk) {
    {
      TP_PROF_T0();
      tmem_ld_wait();
      TP_PROF_ADD(pc[0]);
    }
    if (chunk + 1 < kChunks) tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>((chunk + 1) * 32), r[(chunk + 1) & 1]);
    else release();                                   // every TMEM read of this warp has landed in registers
    const int col0 = col_tile0 + half * kColsPerWarp + chunk * 32;
    // dual output: every 64-column slab exists twice — pre-activation (even slab number, buffer 0) and activation (odd, buffer 1)
    const bool dual = ep.dual != 0 && out.buf != nullptr;
    const uint32_t slab_pre = out.slab_seq + 2u * static_cast<uint32_t>(chunk / kChunksPerSlab);
    const uint32_t slab_q = dual ? slab_pre + 1u : out.slab_seq + static_cast<uint32_t>(chunk / kChunksPerSlab);
    const uint32_t slab_buf = slab_q & static_cast<uint32_t>(out.n_bufs - 1);
    if (out.buf != nullptr && (chunk % kChunksPerSlab) == 0) {
      // the TMA store that last used this staging buffer must have finished READING it (signalled by the store warp)
      TP_PROF_T0();
      if (dual) mbar_wait(&out.empty_bar[slab_pre & 1u], ((slab_pre >> 1) & 1u) ^ 1u);
      mbar_wait(&out.empty_bar[slab_buf], ((slab_q >> (out.n_bufs - 1)) & 1u) ^ 1u);
      TP_PROF_ADD(pc[2]);
    }
    if (col0 < N) {        // N is a multiple of 32 (checked on the host) -> whole chunk in or out
      // kSubPairs packed pairs (2 columns each) go through every step together: each run-time option is ONE warp-uniform branch
      // around a basic block of kSubPairs independent dependency chains for the scheduler to interleave.
      constexpr int kSubPairs = TP_EPI_SUB_PAIRS;
      const int rloc = quarter * 32 + static_cast<int>(lane_id());
#pragma unroll
      for (int sub = 0; sub < 16 / kSubPairs; ++sub) {
        const int lc = chunk * 32 + sub * kSubPairs * 2;             // first column of this sub-block inside the warp's slice
        uint64_t v[kSubPairs];
#pragma unroll
        for (int j = 0; j < kSubPairs; ++j)
          v[j] = pk2(__uint_as_float(r[chunk & 1][sub * kSubPairs * 2 + 2 * j]), __uint_as_float(r[chunk & 1][sub * kSubPairs * 2 + 2 * j + 1]));
        const uint32_t sb4 = sb_addr + static_cast<uint32_t>(lc) * 4u;
        if (ln_fold) {     // v = fma(rstd, fma(-mu, col_a, v), col_b): explicit fmas — ptxas contracts adjacent mul.f32x2 / add.f32x2
          const uint32_t sa4 = sa_addr + static_cast<uint32_t>(lc) * 4u;    // pairs when it sees them, and not in every instantiation alike
#pragma unroll
          for (int q = 0; q < kSubPairs / 2; ++q) {
            const float4 a = lds_f4(sa4 + q * 16), b = lds_f4(sb4 + q * 16);
            v[2 * q] = fma2(rstd2, fma2(nmu2, pk2(a.x, a.y), v[2 * q]), pk2(b.x, b.y));
            v[2 * q + 1] = fma2(rstd2, fma2(nmu2, pk2(a.z, a.w), v[2 * q + 1]), pk2(b.z, b.w));
          }
        } else {
#pragma unroll
          for (int q = 0; q < kSubPairs / 2; ++q) {
            const float4 b = lds_f4(sb4 + q * 16);
            v[2 * q] = add2(v[2 * q], pk2(b.x, b.y));
            v[2 * q + 1] = add2(v[2 * q + 1], pk2(b.z, b.w));
          }
        }
        if (dual) {        // the pre-activation slab (same swizzled position in the other staging buffer)
          const uint32_t row_pre = out_addr + static_cast<uint32_t>((slab_pre & 1u) * kOutSlabBytes + rloc * kSlabRowBytes);
          const int swz_pre = !out.swizzle ? 0 : (kSlabCols == 64 ? (rloc & 7) : ((rloc >> 1) & 3));
#pragma unroll
          for (int g = 0; g < kSubPairs / 4; ++g) {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float lo, hi;
              upk2(v[4 * g + j], lo, hi);
              w[j] = pack_bf16x2(lo, hi);
            }
            const int ci = (chunk % kChunksPerSlab) * 4 + sub * (kSubPairs / 4) + g;
            sts_u4(row_pre + static_cast<uint32_t>((ci ^ swz_pre) << 4), w[0], w[1], w[2], w[3]);
          }
        }
        if (ep.gelu) {
#pragma unroll
          for (int j = 0; j < kSubPairs; ++j) v[j] = gelu_erf_pk(v[j]);
        }
        if (scale) {       // alpha == 1 (every GEMM but in_proj_q): x * 1 is x, skip the multiply
#pragma unroll
          for (int j = 0; j < kSubPairs; ++j) v[j] = mul2(v[j], alpha2);
        }
        uint32_t pk[kSubPairs];
#pragma unroll
        for (int j = 0; j < kSubPairs; ++j) {
          float lo, hi;
          upk2(v[j], lo, hi);
          pk[j] = pack_bf16x2(lo, hi);
        }
        if (do_stats) {    // LayerNorm statistics of the ROUNDED values the next GEMM will read (column order: deterministic)
          if (chunk == 0 && sub == 0) shift = bf16_lo(pk[0]);
#pragma unroll
          for (int j = 0; j < kSubPairs; ++j) {
            const float y0 = __fsub_rn(bf16_lo(pk[j]), shift), y1 = __fsub_rn(bf16_hi(pk[j]), shift);
            s1 = __fadd_rn(s1, __fadd_rn(y0, y1));
            s2 = fmaf(y0, y0, fmaf(y1, y1, s2));
          }
        }
        if (out.buf != nullptr) {
          // 16-byte piece index inside the slab row, XOR-swizzled like TMA does: 128-byte rows (SWIZZLE_128B) with (row & 7),
          // 64-byte rows (SWIZZLE_64B) with ((row >> 1) & 3) — address bits [7,9/10) folded into bits [4,6/7)
          const uint32_t row_base = out_addr + static_cast<uint32_t>(slab_buf * kOutSlabBytes + rloc * kSlabRowBytes);
          const int swz = !out.swizzle ? 0 : (kSlabCols == 64 ? (rloc & 7) : ((rloc >> 1) & 3));
#pragma unroll
          for (int g = 0; g < kSubPairs / 4; ++g) {
            const int ci = (chunk % kChunksPerSlab) * 4 + sub * (kSubPairs / 4) + g;
            sts_u4(row_base + static_cast<uint32_t>((ci ^ swz) << 4), pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
          }
        } else if (row_ok) {
          if (ep.out_f32) {
#pragma unroll
            for (int g = 0; g < kSubPairs / 2; ++g) {
              float4 f;
              upk2(v[2 * g], f.x, f.y);
              upk2(v[2 * g + 1], f.z, f.w);
              *reinterpret_cast<float4*>(c_row32 + col0 + sub * kSubPairs * 2 + g * 4) = f;
            }
          } else {
#pragma unroll
            for (int g = 0; g < kSubPairs / 4; ++g)
              *reinterpret_cast<uint4*>(c_row + col0 + sub * kSubPairs * 2 + g * 8) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
          }
        }
      }
    }
    if (chunk == kChunks - 1 && ep.stats_out != nullptr && row_ok) {
      // statistics are final once the last chunk's values exist; written BEFORE the last slab is handed over so that the store
      // warp's tile-done release (dependent GEMMs of the same launch read them) covers these stores too
      const int slot = (col_tile0 + half * kColsPerWarp) / kColsPerWarp;
      long long srow = row;
      if (ep.wm_s != 0) srow = window_major_row(row, ep.wm_s);
      if (slot < ep.stats_out_slots) {
        // (mean, M2) of this 128-column block: mean = shift + s1/n, M2 = s2 - s1^2/n  (deviations from `shift` are O(std): no cancellation)
        constexpr float inv_n = 1.0f / kColsPerWarp;
        const float dm = __fmul_rn(s1, inv_n);
        reinterpret_cast<float2*>(ep.stats_out)[srow * ep.stats_out_slots + slot] =
            make_float2(__fadd_rn(shift, dm), fmaxf(fmaf(-s1, dm, s2), 0.f));
      }
    }
    if (out.buf != nullptr && (chunk % kChunksPerSlab) == kChunksPerSlab - 1) {
      // slab complete: make the generic-proxy writes visible to the async proxy, then one arrive per warp hands it to the store warp
      TP_PROF_T0();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane_id() == 0) {
        if (dual) mbar_arrive(&out.full_bar[slab_pre & 1u]);
        mbar_arrive(&out.full_bar[slab_buf]);
      }
      TP_PROF_ADD(pc[1]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// KV-attention tiles (problem kind 1): the MHA in-projections of the keys and values FUSED with the local-window attention core
// (builder.py:122-130 == nn.MultiheadAttention, L = 1 query, S = s*s keys per window, 8 heads x 128).
//   A tile = 256 window-major rows (CTA pair) x TWO heads, in two accumulator phases of the ordinary 256 x 256 x K pipeline:
//     phase K: k' = y_k . (gamma_k W_ik)^T for the two heads  -> epilogue: folded LayerNorm, dot with the window's q' (already scaled
//              by 1/sqrt 128), softmax over the s*s consecutive lanes of the window (warp shuffles)  -> p stays in a register
//     phase V: v' = y_v . (gamma_v W_iv)^T                     -> epilogue: folded LayerNorm, p * v', halving exchange over the window's
//              lanes, store the window's 128 context channels of the head
//   The two phases use the two TMEM accumulator buffers alternately, so the MMAs of phase V run under the score epilogue and the
//   next tile's phase K under the P.V epilogue: the tensor pipe sees the same back-to-back 256-wide k-loops as a plain GEMM.
//   Thread == key row; column half (warps 0-3 / 4-7) == head of the pair.  k' and v' never exist in memory (fp32, unrounded, in
//   registers): the [R,1024] x 2 round trip through HBM and the separate attention kernel are gone.
// ------------------------------------------------------------------------------------------------
struct AttnParams {
  const __nv_bfloat16* qp;      // [Q, 1024] q', row = window index (= query index), scaled
  __nv_bfloat16* ctx;           // [Q, 1024]
  const float* stats_k;         // [R, slots, 2] (window-major rows) per-block (mean, M2) of y_k / y_v
  const float* stats_v;
  const float* wsum_k;          // [1024] LayerNorm-fold column vectors of the two in-projections
  const float* cst_k;
  const float* wsum_v;
  const float* cst_v;
  int s;                        // scale factor: W = s*s consecutive rows per window (2 or 4)
  int stats_slots;
  float ln_inv_dim, ln_eps;
  int* done_counter;            // ctx row blocks of 256 queries: counter[(m_blk * 256 / W) / 256] += 1 per (CTA, head pair)
  // dependencies of a tile: the raster row blocks of y_k / y_v covering the crops it touches, and its queries' q' row block
  const int* k_counter;
  const int* v_counter;
  int kv_target;
  const int* q_counter;
  int q_target;
};

__device__ __forceinline__ float bf16x2_get(const uint4& v, int i) {      // i in [0, 8)
  const uint32_t w = i < 4 ? (i < 2 ? v.x : v.y) : (i < 6 ? v.z : v.w);
  return (i & 1) ? bf16_hi(w) : bf16_lo(w);
}

// Phase K epilogue: returns this row's softmax weight p for head `head`.  s_vec: [wsum | cst][256] of the tile's two heads.
template <typename ReleaseFn>
__device__ __forceinline__ float attn_scores(const AttnParams& at, int M, uint32_t tmem_acc, int row, int head, int quarter, int half,
                                             const float* s_vec, ReleaseFn release) {
  const int W = at.s * at.s;
  const bool row_ok = row < M;
  float mu = 0.f, rstd = 0.f;
  if (row_ok) ln_row_stats(at.stats_k, row, at.stats_slots, at.ln_inv_dim, at.ln_eps, mu, rstd);
  const long long window = row / W;
  const float* wsum = s_vec + half * 128;
  const float* cst = s_vec + 256 + half * 128;
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(half * 128);
  const uint4* qrow = reinterpret_cast<const uint4*>(at.qp + window * 1024 + head * 128);
  uint32_t r[2][32];
  tmem_ld_32x32b_x32(taddr, r[0]);
  float score = 0.f;
#pragma unroll
  for (int chunk = 0; chunk < 4; ++chunk) {
    tmem_ld_wait();
    if (chunk + 1 < 4) tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>((chunk + 1) * 32), r[(chunk + 1) & 1]);
    else release();                                   // every TMEM read of this warp has landed in registers
    uint4 q4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) q4[i] = row_ok ? __ldg(qrow + chunk * 4 + i) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const float kf = fmaf(rstd, fmaf(-mu, wsum[chunk * 32 + c], __uint_as_float(r[chunk & 1][c])), cst[chunk * 32 + c]);
      score = fmaf(bf16x2_get(q4[c >> 3], c & 7), kf, score);
    }
  }
  // softmax over the W keys of my window = W consecutive lanes (W divides 32, windows never straddle a warp)
  float mx = score;
  for (int off = 1; off < W; off <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  const float e = __expf(score - mx);
  float den = e;
  for (int off = 1; off < W; off <<= 1) den += __shfl_xor_sync(0xffffffffu, den, off);
  return e * (1.0f / den);
}

// Phase V epilogue: ctx = sum over the window's lanes of p * v'.  Halving exchange: after log2(W) steps each lane holds 32 / W
// channels of the chunk.
template <typename ReleaseFn>
__device__ __forceinline__ void attn_pv(const AttnParams& at, int M, uint32_t tmem_acc, int row, int head, int quarter, int half, float p,
                                        const float* s_vec, ReleaseFn release) {
  const int W = at.s * at.s;
  const bool row_ok = row < M;
  const uint32_t lane = lane_id();
  float mu = 0.f, rstd = 0.f;
  if (row_ok) ln_row_stats(at.stats_v, row, at.stats_slots, at.ln_inv_dim, at.ln_eps, mu, rstd);
  const long long window = row / W;
  const float* wsum = s_vec + half * 128;
  const float* cst = s_vec + 256 + half * 128;
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(half * 128);
  uint32_t r[2][32];
  tmem_ld_32x32b_x32(taddr, r[0]);
#pragma unroll
  for (int chunk = 0; chunk < 4; ++chunk) {
    tmem_ld_wait();
    if (chunk + 1 < 4) tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>((chunk + 1) * 32), r[(chunk + 1) & 1]);
    else release();
    float v[32];
#pragma unroll
    for (int c = 0; c < 32; ++c)
      v[c] = p * fmaf(rstd, fmaf(-mu, wsum[chunk * 32 + c], __uint_as_float(r[chunk & 1][c])), cst[chunk * 32 + c]);
    int first = 0;                                    // my live values cover channels [first, first + 32 >> steps) of the chunk
#pragma unroll
    for (int step = 0; step < 4; ++step) {
      const int off = 1 << step;
      const int hn = 16 >> step;                      // steps run as a prefix (off < W), so the live count before step k is 32 >> k
      if (off < W) {
        const bool up = (lane & off) != 0;            // upper lane of the pair keeps the upper half
#pragma unroll
        for (int i = 0; i < hn; ++i) {
          const float send = up ? v[i] : v[i + hn];
          const float keep = up ? v[i + hn] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
        if (up) first += hn;
      }
    }
    if (row_ok) {
      __nv_bfloat16* dst = at.ctx + window * 1024 + head * 128 + chunk * 32 + first;
      if (W == 4) {
        *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
      } else {                                        // W == 16: two channels per lane
        *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(v[0], v[1]);
      }
    }
  }
}

// Stage this tile's slices of col_a / col_b into shared memory (coalesced), then sync the 256 epilogue threads.
template <int kTileN>
__device__ __forceinline__ void stage_col_vectors(const GemmEpilogue& ep, int N, int col_tile0, float* s_col, int epi_tid,
                                                  bool single_buffer = false) {
  if (single_buffer) named_bar_sync(kEpiBarrierId, kEpiThreads);      // everyone is done reading the previous tile's vectors
  for (int c = epi_tid; c < kTileN; c += kEpiThreads) {
    const int col = col_tile0 + c;
    const bool ok = col < N;
    s_col[c] = (ok && ep.col_a != nullptr) ? __ldg(ep.col_a + col) : 0.f;
    s_col[kTileN + c] = (ok && ep.col_b != nullptr) ? __ldg(ep.col_b + col) : 0.f;
  }
  named_bar_sync(kEpiBarrierId, kEpiThreads);
}

// ================================================================================================
// One-CTA kernel: 128 x kBlockN tiles
// ================================================================================================
template <int kBlockN>
struct GemmConfig {
  static constexpr int kStages = (kBlockN == 256) ? 4 : 6;
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = kBlockN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * kBlockN;  // double-buffered accumulator
  static constexpr int kColStageBytes = 2 * 2 * kBlockN * 4;   // [2 buffers][col_a | col_b][kBlockN] floats
  static constexpr int kBarrierBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kSmemBytes = kStages * kStageBytes + kColStageBytes + kBarrierBytes + 1024;  // +1024: manual alignment
};

template <int kBlockN>
__global__ void __launch_bounds__(kGemmThreads, 1)
tp_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int M, int N, int K,
               int a_seg_rows, GemmEpilogue ep) {
  using Cfg = GemmConfig<kBlockN>;
  constexpr int kStages = Cfg::kStages;
  static_assert(kBlockN == 128 || kBlockN == 256, "BLOCK_N");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  float* s_col_base = reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes + Cfg::kColStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const uint32_t lane = lane_id();

  const int num_m_blocks = (M + kBlockM - 1) / kBlockM;
  const int num_n_blocks = (N + kBlockN - 1) / kBlockN;
  const int num_tiles = num_m_blocks * num_n_blocks;
  const int num_k_blocks = (K + kBlockK - 1) / kBlockK;

  if (warp_idx == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  } else if (warp_idx == kMmaWarp && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], kNumEpiWarps);
    }
    fence_barrier_init();
  } else if (warp_idx == kAllocWarp) {
    tmem_alloc<Cfg::kTmemCols>(tmem_base_smem);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  grid_dependency_wait();                  // PDL: the prologue above overlapped the previous kernel's tail
  grid_launch_dependents();

  if (warp_idx == kTmaWarp) {
    // ======================================= TMA producer =======================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / num_n_blocks;
        const int n_blk = tile - m_blk * num_n_blocks;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          if (a_seg_rows == 0) {
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * kBlockK, m_blk * kBlockM);
          } else {
            // segmented A (3-D map, 64-row boxes): global row g -> (segment g / seg_rows, row g % seg_rows)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int g = m_blk * kBlockM + h * 64;
              const int seg = g / a_seg_rows;
              tma_load_3d(sa + h * (Cfg::kABytes / 2), &tmap_a, &full_bar[stage], kb * kBlockK, g - seg * a_seg_rows, seg);
            }
          }
          tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * kBlockK, n_blk * kBlockN);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp_idx == kMmaWarp) {
    // ======================================= MMA issuer =========================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(kBlockM, kBlockN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);   // epilogue has drained this accumulator buffer
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * kBlockN);
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);                // TMA bytes have landed
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t desc_a = make_smem_desc_kmajor_sw128(sa);
          const uint64_t desc_b = make_smem_desc_kmajor_sw128(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advance 16 elements = 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
            umma_bf16(tmem_d, desc_a + static_cast<uint64_t>(k * 2), desc_b + static_cast<uint64_t>(k * 2), idesc,
                      static_cast<uint32_t>((kb | k) != 0));
          }
          umma_commit(&empty_bar[stage]);                    // smem slot reusable once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tmem_full_bar[acc]);                    // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp_idx >= kEpiWarp0 && warp_idx < kEpiWarp0 + kNumEpiWarps) {
    // ======================================= epilogue ===========================================
    const int e = warp_idx - kEpiWarp0;
    const int quarter = warp_idx & 3;            // TMEM lane quarter this warp may access
    const int half = e >> 2;                     // which half of the tile's columns
    const int epi_tid = e * 32 + static_cast<int>(lane);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n_blocks;
      const int n_blk = tile - m_blk * num_n_blocks;
      float* s_col = s_col_base + acc * 2 * kBlockN;
      stage_col_vectors<kBlockN>(ep, N, n_blk * kBlockN, s_col, epi_tid);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      uint64_t* release_bar = &tmem_empty_bar[acc];
      const OutStage no_stage{nullptr, nullptr, nullptr, 1, 0u, true};
      epilogue_tile<kBlockN>(ep, M, N, tmem_base + static_cast<uint32_t>(acc * kBlockN),
                             m_blk * kBlockM + quarter * 32 + static_cast<int>(lane), n_blk * kBlockN, quarter, half, s_col, no_stage, [&]() {
                               tcgen05_fence_before();
                               __syncwarp();
                               if (lane == 0) mbar_arrive(release_bar);
                             });
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == kAllocWarp) {
    tcgen05_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ================================================================================================
// CTA-pair kernel: 256 x 256 tiles, cta_group::2, GROUPED: one launch runs up to kMaxGroup independent GEMM problems
// (e.g. k_proj_1.2 | v_proj_1.2 | q_proj_1, which share no data but would each leave the machine with a partial last
// wave and pay launch + prologue + drain on their own).  Tiles are numbered across the problems of the group; every warp
// role walks the same sequence.
// ================================================================================================
struct Gemm2Config {
  static constexpr int kTileM = 256;
  static constexpr int kTileN = 256;
#ifndef TP_PAIR_STAGES
#define TP_PAIR_STAGES 5
#endif
#ifndef TP_OUT_BUFS
#define TP_OUT_BUFS ((TP_PAIR_STAGES <= 5 || TP_SLAB_COLS == 32) ? 2 : 1)
#endif
  // Ring depth is what hides the operand-fetch latency (measured on the configs[1] step, same box: 4 / 5 / 6 stages = 1.001 /
  // 0.967 / 0.973 ms); since the store warps took the TMA stores off the epilogue warps' path one staging slab per column half is
  // enough, which is what pays for the fifth stage.
  static constexpr int kStages = TP_PAIR_STAGES;
  static constexpr int kOutBufs = TP_OUT_BUFS;                   // staging buffers per column half (1 or 2)
  static constexpr int kABytes = kBlockM * kBlockK * 2;          // this CTA's 128 rows of A
  static constexpr int kBBytes = (kTileN / 2) * kBlockK * 2;     // this CTA's half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;          // 32 KiB
  static constexpr int kTmemCols = 2 * kTileN;
  static constexpr int kOutBytes = 2 * kOutBufs * kOutSlabBytes; // [2 column halves][kOutBufs] output slabs for TMA stores
  static constexpr int kColStageBytes = 2 * kTileN * 4;          // [col_a | col_b][kTileN] floats, ONE buffer (barrier before it is rewritten)
  static constexpr int kBarrierBytes = (2 * kStages + 4 + 4 * kOutBufs) * 8 + 16;   // ring + accumulators + slab full/empty per half
  // 5 stages x 32 KiB + 4 x 16 KiB slabs + 2 KiB + barriers = 226.2 KiB of the 227 KiB an SM offers: the dynamic shared memory is
  // declared 1024-byte aligned (no alignment slack), and the per-column vectors are single-buffered
  static constexpr int kSmemBytes = kStages * kStageBytes + kOutBytes + kColStageBytes + kBarrierBytes;
  static_assert(kSmemBytes <= 232448, "shared memory budget of an sm_100 SM (227 KiB)");
};

constexpr int kMaxGroup = 8;

constexpr int kMaxAParts = 4;

struct GemmProblem {
  CUtensorMap tmap_a, tmap_b, tmap_c;
  CUtensorMap tmap_cx[kBoxLevels - 1];       // further C maps of the segmented / window-major stores (see the store warps): boxes of
                                             // c_unit << level rows (segmented) or 8 * (level + 1) tokens (window-major); level 0 = tmap_c
  CUtensorMap tmap_a2, tmap_b2;              // kind 1: the value operands (tmap_a / tmap_b: the key operands)
  int kind;              // 0: GEMM with the fused epilogue; 1: KV-attention tile (see attn_epilogue_tile)
  int c_wm_s;            // != 0: tmap_c is the 5-D window-major map of GemmEpilogue::wm_s
  int c_noswz;           // 1: tmap_c (and the peer maps) were built WITHOUT swizzle: the epilogue writes plain slab rows
  AttnParams attn;
  CUtensorMap tmap_a_more[kMaxAParts - 1];   // A given as several tensors side by side along K (e.g. the four CLIP hidden states
                                             // that the reference concatenates, clip_encoder.py:28-44): part p covers k-blocks
                                             // [p * a_kblocks_per_part, (p+1) * a_kblocks_per_part)
  int a_parts;           // 1: a single A tensor
  int a_kblocks_per_part;
  int M, N, K;
  int a_seg_rows;        // 0: plain 2-D A; else rows per segment of the 3-D (crop-strided) A map
  int ab_mn_major;       // 1: BOTH operands are given as row-major [K, M] / [K, N] matrices (wgrad: C = A^T . B, contraction over rows)
                         // 2: only B is ([K, N] row-major: dgrad C = A . B with the weight as stored); A is the usual K-major [M, K]
  int use_tma_store;     // C through TMA stores (0 when rows are scattered to arbitrary segment offsets)
  int c_seg_len;         // != 0: tmap_c (and the peer maps) are 3-D (cols, row in segment, segment): uniform-stride segmented output
  int c_unit;            //       gcd(c_seg_len, 128): every piece of a slab that belongs to one segment is a multiple of it
  int num_n_blocks;
  int num_tiles;         // = tiles_mn * k_splits
  int num_k_blocks;
  // split-K (wgrads whose output is a few tiles but whose contraction runs over every row of the batch): tile = (split, m, n),
  // split s covers k-blocks [s * kb_per_split, (s+1) * kb_per_split) and writes its fp32 partial to slice s of C (ep.out_f32)
  int k_splits;          // >= 1
  int kb_per_split;
  int tiles_mn;
  long long c_split_stride;   // floats between consecutive split slices of C
  // Dependencies between GEMMs of ONE launch (a chain of linears runs as a single persistent kernel: no ramp / drain / partial
  // last wave per layer).  Tiles are numbered problem after problem and every CTA pair walks its tiles in increasing order, so a
  // tile only ever waits for lower-numbered tiles: no deadlock as long as all pairs are co-resident (grid <= 74 pairs).
  int* done_counter;     // != nullptr: [ceil(M/256)] tile counter of THIS problem's output row blocks, +1 per (CTA, column half)
                         //             once that part of a tile is in global memory (bumped by the store warps)
  const int* dep_counter;// != nullptr: the A operand's row block m_blk is ready when dep_counter[m_blk >> dep_shift] >= dep_target
  int dep_target;        //             (= 4 * num_n_blocks of the producing problem: 2 CTAs x 2 column halves per tile)
  int dep_shift;         //             0 for GEMM -> GEMM (same row blocks); 31 for a single launch-wide counter (front work)
  int dep_span;          // != 0: the producer is a KV-attention problem: row block m_blk (256 queries) is complete after dep_per arrivals
  int dep_src_blocks;    //       from each of its source tiles [m_blk * dep_span, min((m_blk + 1) * dep_span, dep_src_blocks))
  int dep_per;
  int peer_out;          // C of this problem goes to the PeerStores maps (fused all-gather) instead of tmap_c
  GemmEpilogue ep;
};

// Optional prologue work of a chained launch: the point queries (builder.py:117-118: bilinear 24x24 -> g x g, align_corners=False
// == a fixed stencil per s x s window: the centre token for odd s, the mean of the centre 2x2 for even s; fp32, one bf16 rounding).
// Done by the epilogue warps of every CTA BEFORE their first tile — that time is otherwise idle (the first accumulator of the
// K=4096 GEMM takes ~33k cycles to appear), so the stencil costs nothing and needs no launch of its own.  Every CTA handles a
// strided share of the (query, 8-channel vector) items and then bumps done_counter once; the GEMM that reads q waits for
// gridDim.x arrivals.
struct FrontWork {
  const __nv_bfloat16* x0;   // nullptr: no front work in this launch
  __nv_bfloat16* q;          // [n_queries, 1024]
  long long crop_stride;     // elements between crops of x0
  long long n_queries;
  int s;                     // scale factor
  int* done_counter;
};

// Optional tile schedule of a chained launch: instead of "stage after stage", the tile numbers run through SEGMENTS (problem,
// first row block, number of row blocks) in the order the host lists them.  The fused forward interleaves the stages by groups of
// row blocks with a fixed lag between producer and consumer stages (a software wavefront): intermediates are consumed while they
// are still in L2, and epilogue-heavy tiles (GELU, attention) alternate with K=4096 tiles on every CTA pair, so neither the
// tensor pipe nor the epilogue warps idle through a whole stage.  Any order in which every tile's producers have lower tile
// numbers is deadlock-free.
struct TileSeg {
  int prob;       // index into GemmGroup::p
  int m_lo;       // first 256-row block
  int tile0;      // first tile number of the segment
  int n_tiles;    // row blocks x n-blocks of the problem
};
constexpr int kMaxSegs = 384;

struct GemmGroup {
  GemmProblem p[kMaxGroup];
  int count;
  int total_tiles;
  FrontWork front;
  int n_segs;                 // 0: tiles are numbered problem after problem
  TileSeg segs[kMaxSegs];
};

struct TileRef {
  const GemmProblem* pr;
  int m_blk, n_blk;
  int kb0, kb1;          // k-block range of this tile (the whole K unless the problem is split)
  int split;
};

__device__ __forceinline__ TileRef decode_tile(const GemmGroup& g, int tile, int& cursor) {
  if (g.n_segs != 0) {
    // scheduled launch: every role walks its tiles in increasing order, so the segment cursor only moves forward
    while (tile >= g.segs[cursor].tile0 + g.segs[cursor].n_tiles) ++cursor;
    const TileSeg& sg = g.segs[cursor];
    TileRef t;
    t.pr = &g.p[sg.prob];
    const int local = tile - sg.tile0;
    const int mm = local / t.pr->num_n_blocks;
    t.m_blk = sg.m_lo + mm;
    t.n_blk = local - mm * t.pr->num_n_blocks;
    t.split = 0;
    t.kb0 = 0;
    t.kb1 = t.pr->num_k_blocks;
    return t;
  }
  int p = 0;
  while (p + 1 < g.count && tile >= g.p[p].num_tiles) {
    tile -= g.p[p].num_tiles;
    ++p;
  }
  TileRef t;
  t.pr = &g.p[p];
  t.split = 0;
  if (t.pr->k_splits > 1) {
    t.split = tile / t.pr->tiles_mn;
    tile -= t.split * t.pr->tiles_mn;
  }
  t.m_blk = tile / t.pr->num_n_blocks;
  t.n_blk = tile - t.m_blk * t.pr->num_n_blocks;
  t.kb0 = t.split * t.pr->kb_per_split;
  t.kb1 = t.kb0 + t.pr->kb_per_split < t.pr->num_k_blocks ? t.kb0 + t.pr->kb_per_split : t.pr->num_k_blocks;
  return t;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
tp_gemm2_kernel(const __grid_constant__ GemmGroup grp, const __grid_constant__ PeerStores peers) {
  using Cfg = Gemm2Config;
  constexpr int kStages = Cfg::kStages;
  constexpr int kTileN = Cfg::kTileN;

  extern __shared__ __align__(1024) uint8_t smem_raw[];     // 1 KiB aligned: swizzle atoms of the operand tiles and output slabs
  uint8_t* smem = smem_raw;
  uint8_t* s_out = smem + kStages * Cfg::kStageBytes;                                  // 1 KiB aligned (swizzle atoms)
  float* s_col_base = reinterpret_cast<float*>(s_out + Cfg::kOutBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_out + Cfg::kOutBytes + Cfg::kColStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* slab_full_bar = tmem_empty_bar + 2;                 // [2 halves][kOutBufs]
  uint64_t* slab_empty_bar = slab_full_bar + 2 * Cfg::kOutBufs; // [2 halves][kOutBufs]
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(slab_empty_bar + 2 * Cfg::kOutBufs);

  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const uint32_t lane = lane_id();
  const uint32_t cta_rank = cluster_ctarank();
  const bool is_leader = cta_rank == 0;
  const int num_tiles = grp.total_tiles;
  const int pair_idx = static_cast<int>(blockIdx.x >> 1);
  const int num_pairs = static_cast<int>(gridDim.x >> 1);

  if (warp_idx == kTmaWarp && lane == 0) {
    for (int i = 0; i < grp.count; ++i) {
      tma_prefetch_desc(&grp.p[i].tmap_a);
      for (int q = 1; q < grp.p[i].a_parts; ++q) tma_prefetch_desc(&grp.p[i].tmap_a_more[q - 1]);
      tma_prefetch_desc(&grp.p[i].tmap_b);
      if (grp.p[i].kind == 1) {
        tma_prefetch_desc(&grp.p[i].tmap_a2);
        tma_prefetch_desc(&grp.p[i].tmap_b2);
      }
      if (grp.p[i].use_tma_store) tma_prefetch_desc(&grp.p[i].tmap_c);
    }
  } else if (warp_idx == kMmaWarp && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 2);          // one arrival per CTA's producer (the leader's carries the expected bytes of both)
      mbar_init(&empty_bar[i], 1);         // multicast tcgen05.commit from the leader
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);                   // multicast tcgen05.commit from the leader
      mbar_init(&tmem_empty_bar[i], 2 * kNumEpiWarps);   // epilogue warps of BOTH CTAs (waited on in the leader only)
    }
    for (int i = 0; i < 2 * Cfg::kOutBufs; ++i) {
      mbar_init(&slab_full_bar[i], kNumEpiWarps / 2);    // one arrive per epilogue warp of the column half
      mbar_init(&slab_empty_bar[i], 1);                  // the half's store warp
    }
    fence_barrier_init();
  } else if (warp_idx == kAllocWarp) {
    tmem_alloc_pair<Cfg::kTmemCols>(tmem_base_smem);
  }
  tcgen05_fence_before();
  cluster_sync_all();                      // barriers of both CTAs initialised before any remote arrive / multicast commit
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  // Programmatic dependent launch: everything above overlapped the tail of the previous kernel on the stream; from here
  // on we read what it wrote.  (No-op when the launch carries no PDL attribute.)
  grid_dependency_wait();
  grid_launch_dependents();                // the next kernel's CTAs may take over SMs as ours exit (they block in their own wait)

  if (warp_idx == kTmaWarp) {
    // ======================================= TMA producer (both CTAs) ============================
    // Whole warp runs the loop (uniform control flow); one elected lane issues the arrive + TMA instructions.
    int stage = 0, cursor = 0;
    uint32_t phase = 0;
    [[maybe_unused]] long long w_empty = 0;
    [[maybe_unused]] const long long t_begin = clock64();
    for (int tile = pair_idx; tile < num_tiles; tile += num_pairs) {
      const TileRef t = decode_tile(grp, tile, cursor);
      const GemmProblem& pr = *t.pr;
      const int row0 = t.m_blk * Cfg::kTileM + static_cast<int>(cta_rank) * kBlockM;          // my 128 rows of A
      const int brow0 = t.n_blk * kTileN + static_cast<int>(cta_rank) * (kTileN / 2);        // my half of the B tile
      // segmented A (3-D map, 64-row boxes): global row g -> (segment g / seg_rows, row g % seg_rows); hoisted per tile
      if (pr.dep_counter != nullptr) {
        // A's row block is written by an earlier problem of this launch: wait until all its tiles have been published (acquire),
        // then order the TMA (async proxy) reads after the acquire
        int target = pr.dep_target;
        if (pr.dep_span != 0) {
          const int lo = t.m_blk * pr.dep_span;
          const int hi = lo + pr.dep_span < pr.dep_src_blocks ? lo + pr.dep_span : pr.dep_src_blocks;
          target = (hi - lo) * pr.dep_per;
        }
        wait_counter_at_least(pr.dep_counter + (t.m_blk >> pr.dep_shift), target);
        fence_proxy_async_all();
      }
      if (pr.kind == 1) {
        // KV-attention tile: y_k . W_ik(head pair)^T, then y_v . W_iv(head pair)^T, through the same ring
        const AttnParams& at = pr.attn;
        if (at.k_counter != nullptr) {
          // y_k / y_v were stored window-major by raster-ordered GEMMs of this launch: wait for every raster row block of the
          // crops this tile touches, and for the q' row block of its windows
          const long long r_lo = static_cast<long long>(t.m_blk) * Cfg::kTileM;
          const long long r_hi = r_lo + Cfg::kTileM < pr.M ? r_lo + Cfg::kTileM : pr.M;
          const int n_blocks = (pr.M + Cfg::kTileM - 1) / Cfg::kTileM;
          int b_lo = static_cast<int>((r_lo / 576) * 576 / Cfg::kTileM);
          int b_hi = static_cast<int>((((r_hi - 1) / 576 + 1) * 576 + Cfg::kTileM - 1) / Cfg::kTileM);
          if (b_hi > n_blocks) b_hi = n_blocks;
          for (int b = b_lo; b < b_hi; ++b) {
            wait_counter_at_least(at.k_counter + b, at.kv_target);
            wait_counter_at_least(at.v_counter + b, at.kv_target);
          }
          if (at.q_counter != nullptr) wait_counter_at_least(at.q_counter + ((r_lo / (at.s * at.s)) >> 8), at.q_target);
          fence_proxy_async_all();
        }
        for (int op = 0; op < 2; ++op) {            // phase K, then phase V: two ordinary 256-wide k-loops (brow0: my half of the head pair's weight rows)
          const CUtensorMap* ta = op == 0 ? &pr.tmap_a : &pr.tmap_a2;
          const CUtensorMap* tb = op == 0 ? &pr.tmap_b : &pr.tmap_b2;
          for (int kb = 0; kb < pr.num_k_blocks; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            if (elect_one()) {
              uint8_t* sa = smem + stage * Cfg::kStageBytes;
              if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
              else mbar_arrive_cluster(&full_bar[stage], 0);
              tma_load_2d_pair(sa, ta, &full_bar[stage], kb * kBlockK, row0);
              tma_load_2d_pair(sa + Cfg::kABytes, tb, &full_bar[stage], kb * kBlockK, brow0);
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
        continue;
      }
      int seg0 = 0, srow0 = 0, seg1 = 0, srow1 = 0;
      if (pr.a_seg_rows != 0) {
        seg0 = row0 / pr.a_seg_rows;
        srow0 = row0 - seg0 * pr.a_seg_rows;
        seg1 = (row0 + 64) / pr.a_seg_rows;
        srow1 = row0 + 64 - seg1 * pr.a_seg_rows;
      }
      for (int kb = t.kb0; kb < t.kb1; ++kb) {
        {
          TP_PROF_T0();
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          TP_PROF_ADD(w_empty);
        }
        if (elect_one()) {
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
          else mbar_arrive_cluster(&full_bar[stage], 0);
          if (pr.ab_mn_major == 1) {
            // boxes of [64 K-rows x 64 MN-elements]: coordinates (mn, k); two MN atoms per operand per CTA
            tma_load_2d_pair(sa, &pr.tmap_a, &full_bar[stage], row0, kb * kBlockK);
            tma_load_2d_pair(sa + Cfg::kABytes / 2, &pr.tmap_a, &full_bar[stage], row0 + 64, kb * kBlockK);
          } else {
            const int part = pr.a_parts > 1 ? kb / pr.a_kblocks_per_part : 0;
            const CUtensorMap* ta = part == 0 ? &pr.tmap_a : &pr.tmap_a_more[part - 1];
            const int ka = (kb - part * pr.a_kblocks_per_part) * kBlockK;      // K coordinate inside this part
            if (pr.a_seg_rows == 0) {
              tma_load_2d_pair(sa, ta, &full_bar[stage], ka, row0);
            } else {
              tma_load_3d_pair(sa, ta, &full_bar[stage], ka, srow0, seg0);
              tma_load_3d_pair(sa + Cfg::kABytes / 2, ta, &full_bar[stage], ka, srow1, seg1);
            }
          }
          if (pr.ab_mn_major == 0) {
            tma_load_2d_pair(sb, &pr.tmap_b, &full_bar[stage], kb * kBlockK, brow0);
          } else {                                       // 1 (TN) and 2 (NN): B is a row-major [K, N] matrix
            tma_load_2d_pair(sb, &pr.tmap_b, &full_bar[stage], brow0, kb * kBlockK);
            tma_load_2d_pair(sb + Cfg::kBBytes / 2, &pr.tmap_b, &full_bar[stage], brow0 + 64, kb * kBlockK);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
#ifdef TP_GEMM_PROFILE
    if (grp.p[0].ep.prof != nullptr && lane == 0) {
      grp.p[0].ep.prof[blockIdx.x * 16 + 0] = w_empty;
      grp.p[0].ep.prof[blockIdx.x * 16 + 1] = clock64() - t_begin;
    }
#endif
  } else if (warp_idx == kMmaWarp) {
    // ======================================= MMA issuer (leader CTA only) ========================
    if (is_leader) {
      int stage = 0, cursor = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      [[maybe_unused]] long long w_full = 0, w_tmem = 0;
      [[maybe_unused]] const long long t_begin = clock64();
      for (int tile = pair_idx; tile < num_tiles; tile += num_pairs) {
        const TileRef mt = decode_tile(grp, tile, cursor);
        const GemmProblem& mpr = *mt.pr;
        const int mn_major = mpr.ab_mn_major;
        const uint32_t idesc = mn_major == 1 ? make_idesc_bf16_f32(Cfg::kTileM, kTileN, 1, 1)
                             : mn_major == 2 ? make_idesc_bf16_f32(Cfg::kTileM, kTileN, 0, 1) : make_idesc_bf16_f32(Cfg::kTileM, kTileN);
        {
          TP_PROF_T0();
          mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);   // both epilogues have drained this accumulator buffer
          TP_PROF_ADD(w_tmem);
        }
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * kTileN);
        if (mpr.kind == 1) {
          // phase K into this accumulator buffer, phase V into the other one: two ordinary 256 x 256 k-loops, each handed to the
          // epilogue on its own (the wait on tmem_empty above covered phase K's buffer)
          constexpr uint32_t idesc_kv = make_idesc_bf16_f32(Cfg::kTileM, kTileN);
          for (int op = 0; op < 2; ++op) {
            if (op == 1) {
              mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
              tcgen05_fence_after();
            }
            const uint32_t tmem_kv = tmem_base + static_cast<uint32_t>(acc * kTileN);
            for (int kb = 0; kb < mpr.num_k_blocks; ++kb) {
              mbar_wait(&full_bar[stage], phase);
              tcgen05_fence_after();
              if (elect_one()) {
                const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
                const uint64_t desc_a = make_smem_desc_kmajor_sw128(sa);
                const uint64_t desc_b = make_smem_desc_kmajor_sw128(sa + Cfg::kABytes);
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                  umma_bf16_pair(tmem_kv, desc_a + static_cast<uint64_t>(k * 2), desc_b + static_cast<uint64_t>(k * 2), idesc_kv,
                                 static_cast<uint32_t>((kb | k) != 0));
                }
                umma_commit_pair(&empty_bar[stage], 0x3);
                if (kb == mpr.num_k_blocks - 1) umma_commit_pair(&tmem_full_bar[acc], 0x3);
              }
              __syncwarp();
              if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
          }
          continue;
        }
        for (int kb = mt.kb0; kb < mt.kb1; ++kb) {
          {
            TP_PROF_T0();
            mbar_wait(&full_bar[stage], phase);              // both CTAs' boxes have landed
            TP_PROF_ADD(w_full);
          }
          tcgen05_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
            if (mn_major == 2) {
              // NN (dgrad: B = the weight as stored, row-major [K, N]): K-major A tile, MN-major B tile
              const uint64_t desc_a = make_smem_desc_kmajor_sw128(sa);
              const uint64_t desc_b = make_smem_desc_mnmajor_sw128(sa + Cfg::kABytes, Cfg::kBBytes / 2);
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                umma_bf16_pair(tmem_d, desc_a + static_cast<uint64_t>(k * 2), desc_b + static_cast<uint64_t>(k * (2048 >> 4)), idesc,
                               static_cast<uint32_t>(((kb - mt.kb0) | k) != 0));
              }
            } else if (!mn_major) {
              const uint64_t desc_a = make_smem_desc_kmajor_sw128(sa);
              const uint64_t desc_b = make_smem_desc_kmajor_sw128(sa + Cfg::kABytes);
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                umma_bf16_pair(tmem_d, desc_a + static_cast<uint64_t>(k * 2), desc_b + static_cast<uint64_t>(k * 2), idesc,
                               static_cast<uint32_t>(((kb - mt.kb0) | k) != 0));
              }
            } else {
              // MN-major tiles: two 64-wide MN atoms 8 KiB apart per operand; one UMMA (K = 16) consumes two 8-row K groups = 2 KiB
              const uint64_t desc_a = make_smem_desc_mnmajor_sw128(sa, Cfg::kABytes / 2);
              const uint64_t desc_b = make_smem_desc_mnmajor_sw128(sa + Cfg::kABytes, Cfg::kBBytes / 2);
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                umma_bf16_pair(tmem_d, desc_a + static_cast<uint64_t>(k * (2048 >> 4)), desc_b + static_cast<uint64_t>(k * (2048 >> 4)), idesc,
                               static_cast<uint32_t>(((kb - mt.kb0) | k) != 0));
              }
            }
            umma_commit_pair(&empty_bar[stage], 0x3);        // frees the slot in BOTH CTAs
            if (kb == mt.kb1 - 1) umma_commit_pair(&tmem_full_bar[acc], 0x3);   // accumulator complete -> both epilogues
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
#ifdef TP_GEMM_PROFILE
      if (grp.p[0].ep.prof != nullptr && lane == 0) {
        grp.p[0].ep.prof[blockIdx.x * 16 + 2] = w_full;
        grp.p[0].ep.prof[blockIdx.x * 16 + 3] = w_tmem;
        grp.p[0].ep.prof[blockIdx.x * 16 + 4] = clock64() - t_begin;
      }
#endif
    }
  } else if (warp_idx >= kEpiWarp0 && warp_idx < kEpiWarp0 + kNumEpiWarps) {
    // ======================================= epilogue (both CTAs, own 128 rows) ==================
    const int e = warp_idx - kEpiWarp0;
    const int quarter = warp_idx & 3;
    const int half = e >> 2;
    const int epi_tid = e * 32 + static_cast<int>(lane);
    int acc = 0, cursor = 0;
    uint32_t acc_phase = 0;
    uint32_t slab_seq = 0;
    [[maybe_unused]] long long w_acc = 0, t_work = 0;
    [[maybe_unused]] long long pc[3] = {0, 0, 0};
    if (grp.front.x0 != nullptr) {
      // point queries: this CTA's share of the (query, 8-channel vector) items, 128 vectors per query
      const FrontWork& fw = grp.front;
      const int g = 24 / fw.s, mq = g * g;
      const int lo = (fw.s & 1) ? (fw.s - 1) / 2 : fw.s / 2 - 1;          // first tap inside the window (row and column)
      const long long items = fw.n_queries * 128;
      for (long long idx = static_cast<long long>(blockIdx.x) * kEpiThreads + epi_tid; idx < items; idx += static_cast<long long>(gridDim.x) * kEpiThreads) {
        const long long query = idx >> 7;
        const int vec = static_cast<int>(idx & 127);
        const long long n = query / mq;
        const int m = static_cast<int>(query - n * mq);
        const int hb = m / g, wb = m - hb * g;
        const __nv_bfloat16* base = fw.x0 + n * fw.crop_stride + vec * 8 + static_cast<long long>((hb * fw.s + lo) * 24 + wb * fw.s + lo) * 1024;
        uint4 o = __ldg(reinterpret_cast<const uint4*>(base));
        if (!(fw.s & 1)) {
          const uint4 b = __ldg(reinterpret_cast<const uint4*>(base + 1024)), c = __ldg(reinterpret_cast<const uint4*>(base + 24 * 1024)),
                      d = __ldg(reinterpret_cast<const uint4*>(base + 25 * 1024));
          // 0.25 * ((a + b) + (c + d)): the association of point_query_kernel (bit-identical; power-of-two scaling is exact)
          auto mean4 = [](uint32_t a_, uint32_t b_, uint32_t c_, uint32_t d_) {
            const float l = 0.25f * ((bf16_lo(a_) + bf16_lo(b_)) + (bf16_lo(c_) + bf16_lo(d_)));
            const float h = 0.25f * ((bf16_hi(a_) + bf16_hi(b_)) + (bf16_hi(c_) + bf16_hi(d_)));
            return pack_bf16x2(l, h);
          };
          o = make_uint4(mean4(o.x, b.x, c.x, d.x), mean4(o.y, b.y, c.y, d.y), mean4(o.z, b.z, c.z, d.z), mean4(o.w, b.w, c.w, d.w));
        }
        *reinterpret_cast<uint4*>(fw.q + query * 1024 + vec * 8) = o;
      }
      named_bar_sync(kEpiBarrierId, kEpiThreads);           // every epilogue thread's stores are issued ...
      if (epi_tid == 0) {
        __threadfence();                                     // ... and ordered (cumulatively) before the release below
        red_release_gpu_add(fw.done_counter, 1);
      }
    }
    for (int tile = pair_idx; tile < num_tiles; tile += num_pairs) {
      const TileRef t = decode_tile(grp, tile, cursor);
      const GemmProblem& pr = *t.pr;
      float* s_col = s_col_base;
      uint64_t* release_bar = &tmem_empty_bar[acc];
      const int row_tile0 = t.m_blk * Cfg::kTileM + static_cast<int>(cta_rank) * kBlockM;
      const int row = row_tile0 + quarter * 32 + static_cast<int>(lane);
      if (pr.kind == 1) {
        const AttnParams& at = pr.attn;
        const int head = t.n_blk * 2 + half;                 // column half == head of the tile's pair
        auto release_acc = [&](uint64_t* bar) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (is_leader) mbar_arrive(bar);
            else mbar_arrive_cluster(bar, 0);
          }
        };
        GemmEpilogue vec;                                    // only col_a / col_b are read by the staging helper
        vec.col_a = at.wsum_k;
        vec.col_b = at.cst_k;
        stage_col_vectors<kTileN>(vec, pr.N, t.n_blk * kTileN, s_col, epi_tid, true);
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tcgen05_fence_after();
        const float p = attn_scores(at, pr.M, tmem_base + static_cast<uint32_t>(acc * kTileN), row, head, quarter, half, s_col,
                                    [&]() { release_acc(release_bar); });
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        float* s_col_v = s_col_base;
        vec.col_a = at.wsum_v;
        vec.col_b = at.cst_v;
        stage_col_vectors<kTileN>(vec, pr.N, t.n_blk * kTileN, s_col_v, epi_tid, true);
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        tcgen05_fence_after();
        uint64_t* release_v = &tmem_empty_bar[acc];
        attn_pv(at, pr.M, tmem_base + static_cast<uint32_t>(acc * kTileN), row, head, quarter, half, p, s_col_v, [&]() { release_acc(release_v); });
        if (at.done_counter != nullptr) {
          named_bar_sync(kEpiBarrierId, kEpiThreads);       // every epilogue thread's ctx stores are issued ...
          if (epi_tid == 0) {
            __threadfence();                                 // ... and ordered (cumulatively) before the release
            red_release_gpu_add(at.done_counter + ((static_cast<long long>(t.m_blk) * Cfg::kTileM / (at.s * at.s)) >> 8), 1);
          }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        continue;
      }
      stage_col_vectors<kTileN>(pr.ep, pr.N, t.n_blk * kTileN, s_col, epi_tid, true);
      {
        TP_PROF_T0();
        mbar_wait(&tmem_full_bar[acc], acc_phase);
        TP_PROF_ADD(w_acc);
      }
      TP_PROF_T0();
      tcgen05_fence_after();
      const OutStage out{pr.use_tma_store ? s_out + half * Cfg::kOutBufs * kOutSlabBytes : nullptr, slab_full_bar + half * Cfg::kOutBufs,
                         slab_empty_bar + half * Cfg::kOutBufs, Cfg::kOutBufs, slab_seq, pr.c_noswz == 0};
      if (pr.use_tma_store) slab_seq += (kTileN / 2 / kSlabCols) * (pr.ep.dual ? 2 : 1);     // slabs per tile and column half
      epilogue_tile<kTileN>(pr.ep, pr.M, pr.N, tmem_base + static_cast<uint32_t>(acc * kTileN), row, t.n_blk * kTileN, quarter, half, s_col,
                            out, [&]() {
                              tcgen05_fence_before();
                              __syncwarp();
                              if (lane == 0) {
                                if (is_leader) mbar_arrive(release_bar);
                                else mbar_arrive_cluster(release_bar, 0);
                              }
                            }, pc, static_cast<long long>(t.split) * pr.c_split_stride);
      TP_PROF_ADD(t_work);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
#ifdef TP_GEMM_PROFILE
    if (grp.p[0].ep.prof != nullptr && e == 0 && lane == 0) {
      grp.p[0].ep.prof[blockIdx.x * 16 + 5] = w_acc;
      grp.p[0].ep.prof[blockIdx.x * 16 + 6] = t_work;
      grp.p[0].ep.prof[blockIdx.x * 16 + 8] = pc[0];          // epilogue warp 0: cycles in tcgen05.wait::ld
      grp.p[0].ep.prof[blockIdx.x * 16 + 9] = pc[1];         // ... in fence.proxy.async
      grp.p[0].ep.prof[blockIdx.x * 16 + 10] = pc[2];          // ... in wait_group.read + named barrier
    }
#endif
  }

  if (warp_idx == kStoreWarp0 || warp_idx == kStoreWarp0 + 1) {
    // ======================================= store warps (both CTAs, one per column half) =========
    // Walks the same tile sequence as the epilogue warps of its half.  Per slab: wait until the 4 epilogue warps have written
    // it (mbarrier), issue the TMA store(s) — plain 2-D box, clipped 3-D boxes for segmented rows, one per peer GPU for the
    // fused all-gather —, wait until the copy engine has READ the buffer and hand it back.  Per tile of a GEMM that others in
    // this launch depend on: wait for the stores to be PERFORMED, then publish the tile (release) on its row block's counter.
    const int half = warp_idx - kStoreWarp0;
    uint64_t* full = slab_full_bar + half * Cfg::kOutBufs;
    uint64_t* empty = slab_empty_bar + half * Cfg::kOutBufs;
    uint32_t q = 0;
    int cursor = 0;
    for (int tile = pair_idx; tile < num_tiles; tile += num_pairs) {
      const TileRef t = decode_tile(grp, tile, cursor);
      const GemmProblem& pr = *t.pr;
      if (!pr.use_tma_store) continue;
      const int row_tile0 = t.m_blk * Cfg::kTileM + static_cast<int>(cta_rank) * kBlockM;
      const bool to_peers = pr.peer_out != 0 && peers.count > 0;
      const int n_maps = to_peers ? peers.count : 1;
      const int dual = pr.ep.dual != 0 ? 1 : 0;     // every column slab twice: pre-activation (tmap_cx[0]) then activation (tmap_c)
      for (int slab = 0; slab < (kTileN / 2 / kSlabCols) << dual; ++slab, ++q) {
        const uint32_t buf = q & static_cast<uint32_t>(Cfg::kOutBufs - 1);
        mbar_wait(&full[buf], (q >> (Cfg::kOutBufs - 1)) & 1u);
        {
          // Every TMA store of this slab is a JOB; lane l issues jobs l, l + 32, ...  All lanes walk the same (cheap) enumeration of the
          // slab's pieces and keep the parameters of their own jobs, then issue them together: the up to ~64 stores of a packed-row
          // slab that goes to eight GPUs leave the warp in two or three instructions instead of one lane issuing them one by one.
          const uint8_t* src = s_out + (half * Cfg::kOutBufs + static_cast<int>(buf)) * kOutSlabBytes;
          const int col = t.n_blk * kTileN + half * (kTileN / 2) + (slab >> dual) * kSlabCols;
          constexpr int kJobsPerLane = 3;                  // 96 jobs per slab at most (host-checked: pieces x destinations)
          const CUtensorMap* jmap[kJobsPerLane];
          int jlo[kJobsPerLane], jc1[kJobsPerLane], jc2[kJobsPerLane], jc3[kJobsPerLane], jc4[kJobsPerLane];
#pragma unroll
          for (int k = 0; k < kJobsPerLane; ++k) jmap[k] = nullptr;
          int job = 0;
          auto add = [&](const CUtensorMap* mp, int lo, int c1, int c2, int c3, int c4) {
#pragma unroll
            for (int k = 0; k < kJobsPerLane; ++k)
              if (job == static_cast<int>(lane) + 32 * k) { jmap[k] = mp; jlo[k] = lo; jc1[k] = c1; jc2[k] = c2; jc3[k] = c3; jc4[k] = c4; }
            ++job;
          };
          // TMA stores must lie entirely inside the tensor (a box that sticks out of a segment faults: measured), so every piece of
          // a slab goes out through boxes of EXACTLY its size: a few maps per destination with box heights unit << level.
          int kind = 2;                                    // dimensionality of this problem's stores: 2, 3 or 5
          if (pr.c_wm_s != 0) {
            // Raster rows -> window-major rows: the slab is cut at token-row boundaries (24 tokens; crops are 24 token rows, so
            // token row R24 = global row / 24 = (crop * g + hb) * s + hi).  Slab edges fall on multiples of 8 tokens — a whole
            // number of windows for s in {2, 4, 8} —, so a piece holds 8, 16 or 24 tokens: map (tokens / 8 - 1), a
            // (channel, wi, -, wb) box of that many windows at (hi, crop-and-hb).
            kind = 5;
            const int sf = pr.c_wm_s;
            const int n_r24 = pr.M / 24;
            int r24 = row_tile0 / 24;
            for (int a = r24 * 24 - row_tile0; a < kBlockM && r24 < n_r24; a += 24, ++r24) {
              const int lo = max(a, 0), hi = min(a + 24, kBlockM);
              const int lvl = (hi - lo) / 8 - 1;
              add(lvl == 0 ? &pr.tmap_c : &pr.tmap_cx[lvl - 1], lo, 0, r24 % sf, (lo - a) / sf, r24 / sf);    // (c, wi, hi, wb, crop-and-hb)
            }
          } else if (pr.c_seg_len == 0) {
            if (dual && (slab & 1) == 0) add(&pr.tmap_cx[0], 0, row_tile0, 0, 0, 0);
            else
              for (int p = 0; p < n_maps; ++p) add(to_peers ? &peers.m[p][0] : &pr.tmap_c, 0, row_tile0, 0, 0, 0);
          } else {
            // Segmented output rows (global row g = seg * seg_len + r  ->  map coordinate (col, r, seg)): the slab's 128 rows are
            // cut at segment boundaries.  Runs of up to 3 WHOLE segments leave as one (cols, seg_len, k) box per destination; a
            // partial piece of L = n * unit rows leaves as one box per set bit of n (largest first).
            kind = 3;
            const int n_segs = pr.M / pr.c_seg_len;
            const bool whole_ok = to_peers && peers.whole != 0;
            int seg = row_tile0 / pr.c_seg_len;
            int a = seg * pr.c_seg_len - row_tile0;
            while (a < kBlockM && seg < n_segs) {
              if (whole_ok && a >= 0 && a + pr.c_seg_len <= kBlockM) {
                int k = 1;
                while (k < kWholeLevels && a + (k + 1) * pr.c_seg_len <= kBlockM && seg + k < n_segs) ++k;
                for (int p = 0; p < n_maps; ++p) add(&peers.m[p][kBoxLevels + k - 1], a, 0, seg, 0, 0);
                a += k * pr.c_seg_len;
                seg += k;
                continue;
              }
              int lo = max(a, 0);
              const int hi = min(a + pr.c_seg_len, kBlockM);
              for (int lvl = kBoxLevels - 1; lvl >= 0; --lvl) {
                const int rows = pr.c_unit << lvl;
                while (hi - lo >= rows) {
                  for (int p = 0; p < n_maps; ++p)
                    add(to_peers ? &peers.m[p][lvl] : (lvl == 0 ? &pr.tmap_c : &pr.tmap_cx[lvl - 1]), lo, lo - a, seg, 0, 0);
                  lo += rows;
                }
              }
              a += pr.c_seg_len;
              ++seg;
            }
          }
          if (job > 32 * kJobsPerLane) __trap();       // cannot happen: the host rejects shapes whose worst slab needs more jobs
#pragma unroll
          for (int k = 0; k < kJobsPerLane; ++k) {
            if (jmap[k] != nullptr) {
              const uint8_t* from = src + jlo[k] * kSlabRowBytes;
              if (kind == 2) tma_store_2d(jmap[k], from, col, jc1[k]);
              else if (kind == 3) tma_store_3d(jmap[k], from, col, jc1[k], jc2[k]);
              else tma_store_5d(jmap[k], from, col, jc1[k], jc2[k], jc3[k], jc4[k]);
            }
          }
          bulk_commit_group();
          bulk_wait_group_read<0>();               // this lane's stores have read the buffer ...
          __syncwarp();                            // ... and so have everybody else's: the epilogue warps may overwrite it
          if (lane == 0) mbar_arrive(&empty[buf]);
        }
        __syncwarp();
      }
      if (pr.done_counter != nullptr) {
        bulk_wait_group<0>();                      // every lane: its stores of this CTA-half's part of the tile are in global memory ...
        fence_proxy_async_all();                   // ... (async-proxy writes) ordered before the generic-proxy release below
        __syncwarp();
        if (lane == 0) red_release_gpu_add(pr.done_counter + t.m_blk, 1);
        __syncwarp();
      }
    }
    bulk_wait_group<0>();                          // my half's last TMA stores have been performed
    if (peers.count > 0) __threadfence_system();   // ... and are ordered before the cross-GPU barrier that follows the kernel
    __syncwarp();
  }

  tcgen05_fence_before();
  cluster_sync_all();                      // the peer may read my smem / signal my barriers until here
  if (warp_idx == kAllocWarp) {
    tcgen05_fence_after();
    tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace tp
