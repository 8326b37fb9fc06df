// Thin inline-PTX wrappers for the sm_100a features the TokenPacker kernels use:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / fences).
// Written against the PTX ISA for CUDA 12.9; no CUTLASS/CuTe dependency.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tp {

// ------------------------------------------------------------------------------------------------
// misc
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// One lane of a fully converged warp (the same lane every time).  Keeping the surrounding loop warp-uniform and
// electing only around the single-thread instructions (TMA, tcgen05.mma, commits) lets ptxas keep loop state in uniform
// registers instead of wrapping every UTMALDG / UTCHMMA in an R2UR "waterfall" loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// Every spin on an mbarrier is bounded: a protocol bug must trap (visible error), never hang the GPU.
#ifndef TP_SPIN_LIMIT_CYCLES
#define TP_SPIN_LIMIT_CYCLES (4000000000ll)   // ~2 s at 1.9 GHz
#endif

// ------------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > TP_SPIN_LIMIT_CYCLES) {
      printf("tokenpacker_b200: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// TMA
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

// 2-D tiled load: coordinates are (c0 = innermost element index, c1 = row index).
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// 3-D tiled load: (c0 = element, c1 = row inside a segment, c2 = segment).  Used for activations whose crops are
// not contiguous (CLIP [:,1:] views): rows are fetched as 64-row boxes that never straddle a crop (576 = 9 * 64).
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// L2 prefetch of a box (no shared-memory destination, no barrier): lets the producer run further ahead of the
// shared-memory ring than its capacity allows, hiding HBM latency of first-touch activation tiles.
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1)
               : "memory");
}

__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap* m, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0),
               "r"(c1), "r"(c2)
               : "memory");
}

// TMA store of a shared-memory box to global memory (bulk async group semantics, tracked per issuing thread).
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
// 3-D form: (c0 = column, c1 = row inside a segment, c2 = segment).  Coordinates are SIGNED and the box is clipped against the
// tensor bounds on both sides: rows with c1 + i < 0 or >= dim1 are simply not written (their shared-memory rows are skipped),
// which is what lets one fixed-size box store the head or the tail of a segment (see the packed-row stores in tp_gemm.cuh).
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// 5-D form, used to store rows that arrive in natural token order (24 x 24 raster per crop) in WINDOW-MAJOR order: the map views
// the destination as (channel, wi, hi, wb, crop-and-hb) with strides that put the s x s tokens of a window next to each other.
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* m, const void* smem_src, int32_t c0, int32_t c1, int32_t c2, int32_t c3,
                                             int32_t c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most kPending of this thread's bulk groups still READ their shared-memory source
template <int kPending>
__device__ __forceinline__ void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}
// wait until at most kPending of this thread's bulk groups are incomplete (writes performed)
template <int kPending>
__device__ __forceinline__ void bulk_wait_group() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(kPending) : "memory");
}

// Same with an L2 cache-policy hint (createpolicy-style 64-bit immediate policies below).
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], "
      "[%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}

constexpr uint64_t kPolicyEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kPolicyEvictLast = 0x14F0000000000000ull;
constexpr uint64_t kPolicyEvictNormal = 0x1000000000000000ull;

// ------------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation
// ------------------------------------------------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
  static_assert(kCols >= 32 && kCols <= 512 && (kCols & (kCols - 1)) == 0, "TMEM columns: power of two in [32,512]");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t tmem_addr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "n"(kCols) : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// tcgen05: descriptors
// ------------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor for a K-major operand tile stored as rows of 128 bytes (64 bf16) with the
// 128-byte swizzle TMA writes (CU_TENSOR_MAP_SWIZZLE_128B): 8-row groups are 1024 B apart (SBO), the leading
// offset is unused for swizzled K-major layouts (encoded 1), descriptor version 1 (Blackwell), layout type 2.
__device__ __forceinline__ uint64_t make_smem_desc_kmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);       // [0,14)  start address >> 4
  d |= static_cast<uint64_t>(1) << 16;                           // [16,30) leading byte offset >> 4 (ignored)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                   // [32,46) stride byte offset >> 4
  d |= static_cast<uint64_t>(1) << 46;                           // [46,48) descriptor version = 1
  d |= static_cast<uint64_t>(2) << 61;                           // [61,64) SWIZZLE_128B
  return d;
}

// Same for an MN-major operand tile (the contraction index K is the SLOW dimension in shared memory: what a TMA box of
// [64 K-rows x 64 MN-elements] of a row-major [K, MN] matrix produces).  Canonical layout (uint128 units):
// Swizzle<3,4,3> o ((8,n),(8,k)) : ((1,LBO),(8,SBO)) — 64 MN-elements contiguous per 128-byte row, 8 K-rows per swizzle
// atom; SBO = distance between consecutive 8-row K groups (1024 B), LBO = distance between 64-wide MN atoms.
__device__ __forceinline__ uint64_t make_smem_desc_mnmajor_sw128(uint32_t smem_addr, uint32_t mn_atom_stride_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(mn_atom_stride_bytes >> 4) << 16;   // leading byte offset: next 64-element MN atom
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                   // stride byte offset: next group of 8 K-rows
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Instruction descriptor for kind::f16: A,B = bf16 (format 1), D = fp32 (format 1), both operands K-major.
__host__ __device__ constexpr uint32_t make_idesc_bf16_f32(uint32_t m, uint32_t n, uint32_t a_mn_major = 0, uint32_t b_mn_major = 0) {
  return (1u << 4)            // [4,6)   D format  : 1 = F32
         | (1u << 7)          // [7,10)  A format  : 1 = BF16
         | (1u << 10)         // [10,13) B format  : 1 = BF16
         | (a_mn_major << 15) // [15]    A major   : 0 = K, 1 = MN
         | (b_mn_major << 16) // [16]    B major   : 0 = K, 1 = MN
         | ((n >> 3) << 17)   // [17,23) N >> 3
         | ((m >> 4) << 24);  // [24,29) M >> 4
}

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread for the whole CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Arrive on an mbarrier once all tcgen05.mma issued so far by this thread have completed
// (implicitly performs tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------------
// tcgen05: TMEM -> registers.  32 lanes x 32 columns of 32 bit: thread t of the warp receives lane
// (lane_base + t), columns col .. col+31.  A warp may only touch lanes 32*(warp_id % 4) .. +31.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// CTA-pair (cta_group::2) variants: two CTAs of a cluster on the two SMs of a TPC act as one 256-row MMA.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// In the shared::cluster window a CTA's own shared memory sits at (cta rank in pair) << 24 | offset: clearing bit 24
// turns a local barrier address into the address of the SAME barrier in the pair's leader (even) CTA.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;

// Arrive (count 1) on the mbarrier at the same offset in CTA `target_cta` of the cluster.
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t target_cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(target_cta)
      : "memory");
}

// TMA loads issued by either CTA of a pair; completion bytes are signalled on the LEADER CTA's mbarrier.
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}

template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t tmem_addr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "n"(kCols) : "memory");
}

// D[tmem of both CTAs] (+)= A (256 rows: 128 from each CTA's smem) * B^T (N columns: N/2 from each CTA's smem).
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Commit: arrive on the mbarrier at this offset in every CTA of `cta_mask` once the pair's MMAs so far have retired.
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// Programmatic dependent launch (PDL).  wait: block until the preceding kernel on the stream has completed and its
// memory is visible (returns immediately if this launch has no programmatic dependency).  launch_dependents: allow the
// NEXT kernel's CTAs to be scheduled (they run their prologue, then block in their own wait).
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// Cross-CTA tile counters in global memory (dependencies between GEMMs of one persistent launch)
// ------------------------------------------------------------------------------------------------
// Orders async-proxy accesses (TMA loads / stores) with the generic-proxy ones around it, all state spaces.
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Bounded like every other spin of the library: a protocol bug must trap, never hang the GPU.
__device__ __forceinline__ void wait_counter_at_least(const int* p, int target) {
  if (ld_acquire_gpu(p) >= target) return;
  const long long t0 = clock64();
  while (ld_acquire_gpu(p) < target) {
    __nanosleep(64);
    if (clock64() - t0 > TP_SPIN_LIMIT_CYCLES) {
      printf("tokenpacker_b200: tile-counter wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ------------------------------------------------------------------------------------------------
// small math / packing helpers
// ------------------------------------------------------------------------------------------------
// Exact-erf GELU (nn.GELU default, builder.py:63,69,81).  erf by Abramowitz & Stegun 7.1.28,
//   erf(t) = 1 - (1 + a1 t + ... + a6 t^6)^-16,  |error| <= 3e-7 analytically, <= 2e-6 evaluated in fp32,
// i.e. a GELU error below 1e-6 absolute — three orders of magnitude under the bf16 rounding of the stored result —
// instead of libdevice erff's two divergent branches (~40 instructions): the epilogue of the K=1024 GEMMs is
// instruction-bound, so this is on the critical path.
// The reciprocal is ONE MUFU.RCP: p >= 1, so none of div.approx's range scaling (FSETP/FSEL/2xFMUL per element) is needed;
// for p > 2^126 (|x| > ~14) both give 1 - tiny = 1.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ float gelu_erf(float x) {
  const float t = __fmul_rn(fabsf(x), 0.70710678118654752440f);
  float p = 0.0000430638f;
  p = fmaf(p, t, 0.0002765672f);
  p = fmaf(p, t, 0.0001520143f);
  p = fmaf(p, t, 0.0092705272f);
  p = fmaf(p, t, 0.0422820123f);
  p = fmaf(p, t, 0.0705230784f);
  p = fmaf(p, t, 1.0f);
  p = __fmul_rn(p, p); p = __fmul_rn(p, p); p = __fmul_rn(p, p); p = __fmul_rn(p, p);   // ^16 (+inf for |x| > ~24 -> erf = 1)
  const float e = __fsub_rn(1.0f, rcp_approx(p));       // erf(|x| / sqrt 2)
  return __fmul_rn(0.5f, fmaf(fabsf(x), e, x));         // 0.5 x (1 + sign(x) erf(|x|/sqrt 2)) = 0.5 (x + |x| e)
}

// Blackwell's packed-fp32 pipe (add/mul/fma .f32x2 = FADD2/FMUL2/FFMA2): two independent IEEE round-to-nearest operations
// per instruction, so every result is bit-identical to the two scalar operations it replaces.
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ uint64_t pk2(float a) { return pk2(a, a); }
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// Two GELUs at once; same bits as two gelu_erf calls (1 - r = fma(r, -1, 1) exactly).
__device__ __forceinline__ uint64_t gelu_erf_pk(uint64_t x) {
  constexpr uint64_t kAbs = 0x7fffffff7fffffffull;
  const uint64_t ax = x & kAbs;
  const uint64_t t = mul2(ax, pk2(0.70710678118654752440f));
  uint64_t p = pk2(0.0000430638f);
  p = fma2(p, t, pk2(0.0002765672f));
  p = fma2(p, t, pk2(0.0001520143f));
  p = fma2(p, t, pk2(0.0092705272f));
  p = fma2(p, t, pk2(0.0422820123f));
  p = fma2(p, t, pk2(0.0705230784f));
  p = fma2(p, t, pk2(1.0f));
  p = mul2(p, p); p = mul2(p, p); p = mul2(p, p); p = mul2(p, p);
  float p0, p1;
  upk2(p, p0, p1);
  const uint64_t e = fma2(pk2(rcp_approx(p0), rcp_approx(p1)), pk2(-1.0f), pk2(1.0f));
  return mul2(pk2(0.5f), fma2(ax, e, x));
}

__device__ __forceinline__ void gelu_erf_x2(float& x0, float& x1) { upk2(gelu_erf_pk(pk2(x0, x1)), x0, x1); }

// Explicit shared-space vector accesses for the epilogue (pointers that travel through structs reach ptxas as generic
// addresses: LD.E/ST.E with an address-space check instead of LDS/STS).  volatile: ordered with the barrier / mbarrier asm.
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

}  // namespace tp
