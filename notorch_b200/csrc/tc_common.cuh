// tcgen05 / TMEM / mbarrier / bulk-TMA PTX wrappers and UMMA descriptor builders shared by the
// tensor-core kernels (gemm_tc.cu, wgrad_tc.cu). sm_100a only.
#pragma once

#include "common.cuh"

namespace nt {
namespace tc {

constexpr int TILE_M = 128;

// byte offset of (row r, 16-byte chunk c) inside a 128B-swizzled K-major tile (Swizzle<3,4,3>)
__host__ __device__ __forceinline__ uint32_t swz128(uint32_t r, uint32_t c) {
  return (r >> 3) * 1024u + (r & 7u) * 128u + ((c ^ (r & 7u)) << 4);
}

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU. The spin loop lives out of line so
// that the dozens of wait sites in the warp-specialised kernels do not each carry the clock / printf / trap code (the hot
// loops have to fit the instruction cache).
static __device__ __noinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("notorch_b200: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_spin(bar, parity);
}
// Same, for waits that are expected to be LONG (epilogue warps waiting for an accumulator, copy warps waiting for a free
// stage): after a few polls the warp sleeps between polls, so that its spin loop stops competing for issue slots with the
// producer warps (ncu: the four epilogue warps of the weight-gradient kernel spent the whole kernel polling, ~28 % of samples).
static __device__ __noinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
  long long t0 = clock64();
  int polls = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++polls > 8) __nanosleep(polls > 64 ? 200 : 40);
    if ((polls & 255) == 0 && clock64() - t0 > 4000000000LL) {
      printf("notorch_b200: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_sleep(bar, parity);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// Asynchronous L2 prefetch of `bytes` (multiple of 16) starting at a 16-byte-aligned global address (TMA engine).
__device__ __forceinline__ void l2_prefetch_bulk(const void* gptr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// One lane of a converged warp. ptxas recognises the ELECT-derived predicate, so code under `if (elect_one())` may feed
// tcgen05 / bulk-copy instructions from uniform registers directly; with `if (lane == 0)` it wraps every such instruction
// in a vote/elect loop (~100+ cycles each, which made the MMA issuer the bottleneck of the fused kernels).
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}
// tcgen05.mma.kind::tf32 with both shared-memory descriptors given as (low word, shared high word): the high words of
// all descriptors of one layout are identical, the low word is base + (byte offset >> 4).
__device__ __forceinline__ void umma_tf32_lo(uint32_t tmem_d, uint32_t adesc_lo, uint32_t bdesc_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(adesc_lo), "r"(bdesc_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate) : "memory");
}
constexpr uint32_t KMAJOR_SW128_DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1 KiB, version 1, SWIZZLE_128B
constexpr uint32_t MNMAJOR_SW128B32_DESC_HI = (512u >> 4) | (1u << 14) | (1u << 29);  // SBO = 512 B, version 1, SWIZZLE_128B_BASE32B
__device__ __forceinline__ uint32_t kmajor_desc_lo(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint32_t mnmajor_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) { return ((smem_addr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16); }

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// what kind::tf32 makes of an fp32 word left as is in shared memory: the low 13 mantissa bits are ignored
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// hi / lo split of a streamed operand for the three-product TF32 scheme. kind::tf32 reads the upper 19 bits of a shared-memory word,
// i.e. the fp32 value TRUNCATED to TF32, so a value can also be left as it is and serve as its own hi part.
//   SPLIT_RNA      hi = rna(v), lo = rna(v - hi): two conversions per value, hi has to be written; smallest error (K2: not ALU-bound)
//   SPLIT_INPLACE  hi = v as it is (never written when the tile holds the raw rows), lo = rna(v - trunc(v)): one conversion (K4a)
//   SPLIT_NOCVT    hi = v as it is, lo = v - trunc(v) left to the tensor core's truncation: no conversion. cvt.rna.tf32 issues at a
//                  fraction of the FP32 rate and K4b's producers are ALU / shared-memory bound: 267 -> 241 us at d = 300, with the
//                  weight-gradient error unchanged to two digits (4.1e-6 -> 4.2e-6 at d = 1024, measured)
enum { SPLIT_RNA = 0, SPLIT_INPLACE = 1, SPLIT_NOCVT = 2 };
template <int KIND>
__device__ __forceinline__ void tf32_split4(const float4& v, float4& hi, float4& lo) {
  if constexpr (KIND == SPLIT_RNA) {
    hi = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
    lo = make_float4(tf32_rna(v.x - hi.x), tf32_rna(v.y - hi.y), tf32_rna(v.z - hi.z), tf32_rna(v.w - hi.w));
  } else {
    const float4 t = make_float4(tf32_trunc(v.x), tf32_trunc(v.y), tf32_trunc(v.z), tf32_trunc(v.w));
    hi = v;
    if constexpr (KIND == SPLIT_INPLACE) lo = make_float4(tf32_rna(v.x - t.x), tf32_rna(v.y - t.y), tf32_rna(v.z - t.z), tf32_rna(v.w - t.w));
    else lo = make_float4(v.x - t.x, v.y - t.y, v.z - t.z, v.w - t.w);
  }
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format: version = 1 at bit 46,
// layout type 2 at bits 61-63, SBO = 1024 B between 8-row groups, LBO unused for swizzled K-major).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::tf32 instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128.
__host__ __device__ __forceinline__ uint32_t make_idesc_tf32(int n, bool mn_major = false) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (mn_major ? (3u << 15) : 0u) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}
// MN-major TF32 operands: the only legal shared-memory layout is SWIZZLE_128B_BASE32B (descriptor layout type 1;
// CUTLASS: "for mn-major tf32 operands, SW128_32B is the only available smem layout"). 128-byte rows hold 32
// consecutive M/N elements of one k; the swizzle atom is 4 k-rows (512 B) and XORs the 32-byte unit index
// (address bits 5-6) with the k-row index mod 4 (address bits 7-8): Swizzle<2,5,2>.
//   LBO = byte stride between successive 32-element M/N chunks, SBO = byte stride between groups of 4 k-rows.
__device__ __forceinline__ uint64_t make_mnmajor_sw128b32_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) |
         (1ull << 46) | (1ull << 61);
}
// byte offset of (k-row r, 16-byte chunk c16 in [0,8)) inside one 32-feature MN-major chunk whose k-rows are 128 B apart
__host__ __device__ __forceinline__ uint32_t mn_swz32(uint32_t r, uint32_t c16) {
  return r * 128u + ((((c16 >> 1) ^ (r & 3u)) << 5) | ((c16 & 1u) << 4));
}
// ---- cluster / cta_group::2 PTX ----------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (.release.cta) as in CUTLASS ClusterBarrier::arrive(cta_id): the .release.cluster form compiles to
  // MEMBAR + ERRBAR and its .acquire.cluster counterpart to CCTL.IVALL (an L1 flush per wait) - ncu showed both as top stalls
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma2_tf32_lo(uint32_t tmem_d, uint32_t adesc_lo, uint32_t bdesc_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %4, p;\n\t}"
      ::"r"(tmem_d), "r"(adesc_lo), "r"(bdesc_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives (once) on the barrier at the same shared-memory offset in BOTH CTAs of the pair when all prior MMAs retire
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}

}  // namespace tc
}  // namespace nt
