// Persistent warp-specialised tcgen05 GEMM for sm_100a with the fused epilogues the TokenPacker path needs.
//
//   C[M,N] (bf16) = epilogue( A[M,K] (bf16, K-major) . B[N,K]^T (bf16, K-major) ),  fp32 accumulation in TMEM.
//
// Every nn.Linear on the reference hot path (builder.py:59-83; MHA in/out projections builder.py:77) is an
// instance of this kernel: activations are [rows, in] and weights are [out, in], both K-major, which is exactly the
// operand form tcgen05.mma takes from shared memory, so no transposes exist anywhere.
//
// CTA = 384 threads, one CTA per SM, persistent over 128 x BLOCK_N output tiles:
//   warp 0      TMA producer   (one lane): cp.async.bulk.tensor 2-D boxes, 128B swizzle, kStages-deep mbarrier ring
//   warp 1      MMA issuer     (one lane): tcgen05.mma cta_group::1 kind::f16, 128 x BLOCK_N x 16, fp32 accum in TMEM
//   warp 2      TMEM allocator (2 accumulator buffers of BLOCK_N columns: the epilogue of tile i overlaps tile i+1's MMAs)
//   warps 4-11  epilogue: tcgen05.ld 32x32b (thread == output row), fused per-row / per-column math, 16-byte stores
//
// Fused epilogue (all optional, selected at run time, warp-uniform branches):
//   v = acc
//   v = rstd_r * (v - mu_r * col_a[c])           LayerNorm folded into the NEXT linear: (mu, rstd) from per-row sums
//   v = v + col_b[c]                             bias (or the folded constant W.beta + b)
//   v = gelu_erf(v)                              exact erf GELU (nn.GELU default)
//   v = alpha * v                                1/sqrt(head_dim) query scaling
//   y = bf16(v);  stats_out[r] += (y, y*y)       per-row sums of the ROUNDED values for the next LayerNorm fold
//   C[dst_row(r), c] = y                         optional segment scatter (HD packed output)
#pragma once

#include "tp_ptx.cuh"

namespace tp {

struct GemmEpilogue {
  __nv_bfloat16* c;        // output
  long long ldc;           // elements between output rows
  const float* col_a;      // [N]  LN fold: row-sum of the gamma-folded weight      (nullptr: no LN fold)
  const float* col_b;      // [N]  bias                                             (nullptr: none)
  const float* stats_in;   // [M,2] (sum, sum of squares) of the A rows over ln_dim (required with col_a)
  float* stats_out;        // [M,2] accumulated with atomics                        (nullptr: none)
  const long long* seg_row_offset;  // [M / seg_len] destination row of each segment's first row (nullptr: identity)
  int seg_len;
  float ln_inv_dim;        // 1 / ln_dim
  float ln_eps;
  float alpha;
  int gelu;
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;     // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 384;
constexpr int kEpiWarp0 = 4;
constexpr int kNumEpiWarps = 8;

template <int kBlockN>
struct GemmConfig {
  static constexpr int kStages = (kBlockN == 256) ? 4 : 6;
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = kBlockN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * kBlockN;  // double-buffered accumulator
  static constexpr int kBarrierBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarrierBytes + 1024;  // +1024: manual 1 KiB alignment
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
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
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

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  } else if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], kNumEpiWarps);
    }
    fence_barrier_init();
  } else if (warp_idx == 2) {
    tmem_alloc<Cfg::kTmemCols>(tmem_base_smem);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp_idx == 0) {
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
  } else if (warp_idx == 1) {
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
  } else if (warp_idx >= kEpiWarp0) {
    // ======================================= epilogue ===========================================
    const int e = warp_idx - kEpiWarp0;
    const int quarter = warp_idx & 3;            // TMEM lane quarter this warp may access
    const int half = e >> 2;                     // which half of the tile's columns
    constexpr int kColsPerWarp = kBlockN / 2;
    const bool ln_fold = ep.col_a != nullptr;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n_blocks;
      const int n_blk = tile - m_blk * num_n_blocks;
      const int row = m_blk * kBlockM + quarter * 32 + static_cast<int>(lane);
      const bool row_ok = row < M;
      float mu = 0.f, rstd = 1.f;
      if (ln_fold && row_ok) {
        const float2 st = *reinterpret_cast<const float2*>(ep.stats_in + 2ll * row);
        mu = st.x * ep.ln_inv_dim;
        const float var = fmaxf(st.y * ep.ln_inv_dim - mu * mu, 0.f);
        rstd = rsqrtf(var + ep.ln_eps);
      }
      long long dst_row = row;
      if (ep.seg_row_offset != nullptr && row_ok) {
        const int seg = row / ep.seg_len;
        dst_row = ep.seg_row_offset[seg] + (row - seg * ep.seg_len);
      }
      __nv_bfloat16* c_row = ep.c + dst_row * ep.ldc;

      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int chunk = 0; chunk < kColsPerWarp / 32; ++chunk) {
        const int col_in_tile = half * kColsPerWarp + chunk * 32;
        const int col0 = n_blk * kBlockN + col_in_tile;
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * kBlockN + col_in_tile), r);
        tmem_ld_wait();
        if (col0 < N) {        // N is a multiple of 32 (checked on the host) -> whole chunk in or out
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {     // 8 columns = one 16-byte store
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g8 * 8 + j]);
            const int c8 = col0 + g8 * 8;
            if (ln_fold) {
              const float4 a0 = __ldg(reinterpret_cast<const float4*>(ep.col_a + c8));
              const float4 a1 = __ldg(reinterpret_cast<const float4*>(ep.col_a + c8 + 4));
              const float ca[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = rstd * (v[j] - mu * ca[j]);
            }
            if (ep.col_b != nullptr) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.col_b + c8));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(ep.col_b + c8 + 4));
              const float cb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] += cb[j];
            }
            if (ep.gelu) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
            }
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              pk[j] = pack_bf16x2(v[2 * j] * ep.alpha, v[2 * j + 1] * ep.alpha);
              const float y0 = bf16_lo(pk[j]), y1 = bf16_hi(pk[j]);
              s1 += y0 + y1;
              s2 = fmaf(y0, y0, fmaf(y1, y1, s2));
            }
            if (row_ok) *reinterpret_cast<uint4*>(c_row + c8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
      // all TMEM reads of this warp for this accumulator are complete -> hand the buffer back to the MMA warp
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if (ep.stats_out != nullptr && row_ok) {
        atomicAdd(ep.stats_out + 2ll * row, s1);
        atomicAdd(ep.stats_out + 2ll * row + 1, s2);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tcgen05_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace tp
