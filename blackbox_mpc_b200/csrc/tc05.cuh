// tc05.cuh — thin inline-PTX layer over the sm_100a tensor-core / TMEM / mbarrier /
// bulk-copy instructions used by the rollout kernel.  No CUTLASS dependency.
//
// Conventions
//   * smem addresses are 32-bit shared-window addresses (cvta.to.shared).
//   * TMEM addresses are 32-bit: bits[31:16] = lane (data path), bits[15:0] = column.
//   * every spin-wait carries a watchdog so a protocol bug traps instead of hanging the GPU.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tc05 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ----------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy smem writes -> visible to the async proxy (UMMA operand reads, bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// try_wait with an explicit suspend-time hint (ns): the thread sleeps in hardware until the phase completes or the
// hint elapses.  Without a hint the default limit is short and a waiting warp burns issue slots in its retry loop
// (ncu, profiles/r2*: tens of millions of TRYWAIT iterations per launch from warps that wait for whole layers).
__device__ __forceinline__ uint32_t mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(ns)
      : "memory");
  return ok;
}
// Long waits (a whole layer or step): sleeps up to ~20 us per attempt; traps after ~1 s.
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, volatile uint32_t* dbg = nullptr, uint32_t tag = 0) {
  uint32_t spins = 0;
  while (!mbar_try_wait_hint(bar, parity, 20000u)) {
    if (++spins > (1u << 16)) {
      if (dbg && (threadIdx.x & 31) == 0) {
        const uint32_t slot = 8u * (threadIdx.x >> 5) + 256u * (blockIdx.x & 1);
        dbg[slot + 0] = 0xDEAD0000u | threadIdx.x; dbg[slot + 1] = bar; dbg[slot + 2] = parity; dbg[slot + 3] = tag;
        __threadfence_system();
      }
      asm volatile("trap;");
    }
  }
}
#ifndef TC05_WATCHDOG_SPINS
#define TC05_WATCHDOG_SPINS (1u << 24)
#endif
// Blocks until the phase with the given parity completes.  Traps after a bounded number of
// polls so a broken pipeline aborts the launch (sticky error on the host) instead of hanging.
// When TC05_DEBUG_BUF is defined it names a `volatile uint32_t*` in scope (host-mapped memory): the
// watchdog records which barrier/parity/thread starved before trapping.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, volatile uint32_t* dbg = nullptr,
                                          uint32_t tag = 0) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    ++spins;
    if (dbg && spins == TC05_WATCHDOG_SPINS / 2 && (threadIdx.x & 31) == 0) {
      const uint32_t slot = 8u * (threadIdx.x >> 5) + 256u * (blockIdx.x & 1);
      dbg[slot + 0] = 0xDEAD0000u | threadIdx.x; dbg[slot + 1] = bar; dbg[slot + 2] = parity; dbg[slot + 3] = tag;
      __threadfence_system();
    }
    if (spins > TC05_WATCHDOG_SPINS) { asm volatile("trap;"); }
  }
}

// One lane of a fully converged warp (the same lane every time): predicate for single-thread issue.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------- bulk copy (UBLKCP)
// 1-D global -> shared copy on the async proxy; completion is signalled as tx-bytes on `bar`.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst_smem), "l"(src_gmem), "r"(bytes), "r"(bar)
      : "memory");
}

// ----------------------------------------------------------------------------- TMEM management
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor, K-major, no swizzle ("interleave"): the operand is a grid of
// 8-row x 16-byte core matrices, each 128 contiguous bytes.
//   lbo = byte distance between the two core matrices adjacent along K inside one K=16 MMA
//   sbo = byte distance between core matrices adjacent along M/N (next 8 rows)
__device__ __forceinline__ uint64_t smem_desc_kmajor_noswz(uint32_t saddr, uint32_t lbo,
                                                           uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;  // descriptor version 1 (Blackwell)
  return d;         // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}
// Instruction descriptor for kind::f16 with BF16 A/B, FP32 accumulate, both operands K-major.
__host__ __device__ constexpr uint32_t idesc_bf16_f32(uint32_t M, uint32_t N) {
  return (1u << 4) /*D=f32*/ | (1u << 7) /*A=bf16*/ | (1u << 10) /*B=bf16*/ | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}

// ----------------------------------------------------------------------------- MMA
// D[tmem] (+)= A[tmem] * B[smem]^T   (A: 128 lanes x K=16 bf16 packed 2/column; B: N x 16, K-major)
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                       uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                       uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One ring UNIT of an SS contraction (two K-chunks, the second optional; bf16x3 = three MMAs per chunk, else one),
// issued by one elected lane of a CONVERGED warp, followed by up to two commits, with two non-blocking mbarrier
// probes issued FIRST and read back LAST.  A satisfied try_wait / test_wait costs ~150-200 cycles of latency in the
// issuing warp's in-order instruction stream; inside one asm block the probes' latency overlaps the MMA issue, and the
// caller only blocks on a barrier whose probe failed (trace: 1300 -> ~500 cycles per unit in the issuer, profiles/r2*).
//   a0 / b0: descriptors of chunk 0 (hi image); *_lo: offset of the bf16-lo image; *_chunk: offset of chunk 1 (all in
//   descriptor units of 16 bytes, added to the 64-bit descriptor).  commit1 == 0: no second commit.
// (A variant of this block specialised at compile time on three / two / commit1, every instruction guarded by the elect
// predicate only, measured SLOWER: 1.352 vs 1.237 ms per C4 rollout — eight copies of the block in the issuer's loop.)
__device__ __forceinline__ void mma_unit_ss_probe(uint32_t d_tmem, uint64_t a0, uint64_t b0, uint32_t a_lo, uint32_t b_lo,
                                                  uint32_t a_chunk, uint32_t b_chunk, uint32_t idesc, uint32_t acc,
                                                  uint32_t three, uint32_t two, uint32_t commit0, uint32_t commit1,
                                                  uint32_t probe0, uint32_t parity0, uint32_t probe1, uint32_t parity1,
                                                  uint32_t& ok0, uint32_t& ok1) {
  // operands: %0 ok0, %1 ok1 | %2 d, %3 a0, %4 b0, %5 a_lo, %6 b_lo, %7 a_chunk, %8 b_chunk, %9 idesc, %10 acc, %11 three,
  //           %12 two, %13 commit0, %14 commit1, %15 probe0, %16 parity0, %17 probe1, %18 parity1
  asm volatile(
      "{\n\t"
      ".reg .pred q0, q1, pe, pacc, p3, p2, p23, pc1, pt;\n\t"
      ".reg .b64 al, bl, a1, b1, a1l, b1l, t;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q0, [%15], %16;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q1, [%17], %18;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pacc, %10, 0;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "setp.ne.b32 p3, %11, 0;\n\t"
      "and.pred p3, p3, pe;\n\t"
      "setp.ne.b32 p2, %12, 0;\n\t"
      "and.pred p2, p2, pe;\n\t"
      "and.pred p23, p2, p3;\n\t"
      "setp.ne.b32 pc1, %14, 0;\n\t"
      "and.pred pc1, pc1, pe;\n\t"
      "cvt.u64.u32 t, %5;\n\t add.u64 al, %3, t;\n\t"
      "cvt.u64.u32 t, %6;\n\t add.u64 bl, %4, t;\n\t"
      "cvt.u64.u32 t, %7;\n\t add.u64 a1, %3, t;\n\t add.u64 a1l, al, t;\n\t"
      "cvt.u64.u32 t, %8;\n\t add.u64 b1, %4, t;\n\t add.u64 b1l, bl, t;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%2], %3, %4, %9, pacc;\n\t"
      "@p3 tcgen05.mma.cta_group::1.kind::f16 [%2], al, %4, %9, pt;\n\t"
      "@p3 tcgen05.mma.cta_group::1.kind::f16 [%2], %3, bl, %9, pt;\n\t"
      "@p2 tcgen05.mma.cta_group::1.kind::f16 [%2], a1, b1, %9, pt;\n\t"
      "@p23 tcgen05.mma.cta_group::1.kind::f16 [%2], a1l, b1, %9, pt;\n\t"
      "@p23 tcgen05.mma.cta_group::1.kind::f16 [%2], a1, b1l, %9, pt;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%13];\n\t"
      "@pc1 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%14];\n\t"
      "selp.u32 %0, 1, 0, q0;\n\t"
      "selp.u32 %1, 1, 0, q1;\n\t"
      "}"
      : "=r"(ok0), "=r"(ok1)
      : "r"(d_tmem), "l"(a0), "l"(b0), "r"(a_lo), "r"(b_lo), "r"(a_chunk), "r"(b_chunk), "r"(idesc), "r"(acc), "r"(three),
        "r"(two), "r"(commit0), "r"(commit1), "r"(probe0), "r"(parity0), "r"(probe1), "r"(parity1)
      : "memory");
}
#ifndef TC05_UNIT_BRANCH
#define TC05_UNIT_BRANCH 0   // 1: branch around the unit instead of predicating every tcgen05 instruction (A/B)
#endif
// The common case of mma_unit_ss_probe — three passes, both K-chunks present — with every tcgen05 instruction guarded by
// the elect predicate alone (no per-instruction runtime predicates: each one costs a VOTEU + predicate logic in SASS) and
// the descriptors built from their 32-bit low words (the high word, SBO / version / swizzle, is one constant).
__device__ __forceinline__ void mma_unit_ss_probe_full(uint32_t d_tmem, uint32_t a0_lo, uint32_t b0_lo, uint32_t desc_hi, uint32_t a_lo,
                                                       uint32_t b_lo, uint32_t a_chunk, uint32_t b_chunk, uint32_t idesc, uint32_t acc,
                                                       uint32_t commit0, uint32_t commit1, uint32_t probe0, uint32_t parity0,
                                                       uint32_t probe1, uint32_t parity1, uint32_t& ok0, uint32_t& ok1) {
  // %0 ok0, %1 ok1 | %2 d, %3 a0_lo, %4 b0_lo, %5 desc_hi, %6 a_lo, %7 b_lo, %8 a_chunk, %9 b_chunk, %10 idesc, %11 acc,
  // %12 commit0, %13 commit1, %14 probe0, %15 parity0, %16 probe1, %17 parity1
  asm volatile(
      "{\n\t"
      ".reg .pred q0, q1, pe, pacc, pc1, pt;\n\t"
      ".reg .b32 t;\n\t"
      ".reg .b64 a0, b0, al, bl, a1, b1, a1l, b1l;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q0, [%14], %15;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q1, [%16], %17;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pacc, %11, 0;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "setp.ne.b32 pc1, %13, 0;\n\t"
      "and.pred pc1, pc1, pe;\n\t"
      "mov.b64 a0, {%3, %5};\n\t"
      "mov.b64 b0, {%4, %5};\n\t"
      "add.u32 t, %3, %6;\n\t mov.b64 al, {t, %5};\n\t"
      "add.u32 t, %4, %7;\n\t mov.b64 bl, {t, %5};\n\t"
      "add.u32 t, %3, %8;\n\t mov.b64 a1, {t, %5};\n\t"
      "add.u32 t, t, %6;\n\t mov.b64 a1l, {t, %5};\n\t"
      "add.u32 t, %4, %9;\n\t mov.b64 b1, {t, %5};\n\t"
      "add.u32 t, t, %7;\n\t mov.b64 b1l, {t, %5};\n\t"
#if TC05_UNIT_BRANCH
      "@!pe bra UNIT_DONE;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%2], a0, b0, %10, pacc;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%2], al, b0, %10, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%2], a0, bl, %10, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%2], a1, b1, %10, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%2], a1l, b1, %10, pt;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%2], a1, b1l, %10, pt;\n\t"
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%12];\n\t"
      "@pc1 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%13];\n\t"
      "UNIT_DONE:\n\t"
#else
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%2], a0, b0, %10, pacc;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%2], al, b0, %10, pt;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%2], a0, bl, %10, pt;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%2], a1, b1, %10, pt;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%2], a1l, b1, %10, pt;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%2], a1, b1l, %10, pt;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%12];\n\t"
      "@pc1 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%13];\n\t"
#endif
      "selp.u32 %0, 1, 0, q0;\n\t"
      "selp.u32 %1, 1, 0, q1;\n\t"
      "}"
      : "=r"(ok0), "=r"(ok1)
      : "r"(d_tmem), "r"(a0_lo), "r"(b0_lo), "r"(desc_hi), "r"(a_lo), "r"(b_lo), "r"(a_chunk), "r"(b_chunk), "r"(idesc), "r"(acc),
        "r"(commit0), "r"(commit1), "r"(probe0), "r"(parity0), "r"(probe1), "r"(parity1)
      : "memory");
}
// Same for a unit that holds ONE K-chunk (the last unit of a layer with an odd number of chunks): three MMAs.
__device__ __forceinline__ void mma_unit_ss_probe_half(uint32_t d_tmem, uint32_t a0_lo, uint32_t b0_lo, uint32_t desc_hi, uint32_t a_lo,
                                                       uint32_t b_lo, uint32_t idesc, uint32_t acc, uint32_t commit0, uint32_t commit1,
                                                       uint32_t probe0, uint32_t parity0, uint32_t probe1, uint32_t parity1,
                                                       uint32_t& ok0, uint32_t& ok1) {
  // %0 ok0, %1 ok1 | %2 d, %3 a0_lo, %4 b0_lo, %5 desc_hi, %6 a_lo, %7 b_lo, %8 idesc, %9 acc, %10 commit0, %11 commit1,
  // %12 probe0, %13 parity0, %14 probe1, %15 parity1
  asm volatile(
      "{\n\t"
      ".reg .pred q0, q1, pe, pacc, pc1, pt;\n\t"
      ".reg .b32 t;\n\t"
      ".reg .b64 a0, b0, al, bl;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q0, [%12], %13;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 q1, [%14], %15;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "setp.ne.b32 pacc, %9, 0;\n\t"
      "setp.eq.b32 pt, 0, 0;\n\t"
      "setp.ne.b32 pc1, %11, 0;\n\t"
      "and.pred pc1, pc1, pe;\n\t"
      "mov.b64 a0, {%3, %5};\n\t"
      "mov.b64 b0, {%4, %5};\n\t"
      "add.u32 t, %3, %6;\n\t mov.b64 al, {t, %5};\n\t"
      "add.u32 t, %4, %7;\n\t mov.b64 bl, {t, %5};\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%2], a0, b0, %8, pacc;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%2], al, b0, %8, pt;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%2], a0, bl, %8, pt;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%10];\n\t"
      "@pc1 tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%11];\n\t"
      "selp.u32 %0, 1, 0, q0;\n\t"
      "selp.u32 %1, 1, 0, q1;\n\t"
      "}"
      : "=r"(ok0), "=r"(ok1)
      : "r"(d_tmem), "r"(a0_lo), "r"(b0_lo), "r"(desc_hi), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(commit0), "r"(commit1),
        "r"(probe0), "r"(parity0), "r"(probe1), "r"(parity1)
      : "memory");
}
// All previously issued MMAs of this thread arrive (count 1) on `bar` when they complete.
// Implies tcgen05.fence::before_thread_sync.
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}

// ----------------------------------------------------------------------------- TMEM <-> registers
// 32x32b shape: thread i of the warp touches lane (lane_base + i); .xN moves N consecutive columns.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
      "r"(r[7])
      : "memory");
}
// hi[8] | lo[8] of one K-chunk occupy 16 consecutive columns: one instruction instead of two x8 stores
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&a)[8], const uint32_t (&b)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
      "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7])
      : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3])
               : "memory");
}

// ----------------------------------------------------------------------------- bf16 split helpers
// Round-to-nearest-even fp32 -> bf16 pair packed {hi16 = b, lo16 = a}: element `a` sits in the low half.
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float bf16_lo_as_f32(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16_hi_as_f32(uint32_t packed) {
  return __uint_as_float(packed & 0xFFFF0000u);
}
// x = hi + lo (+ O(2^-18 |x|)):  hi = bf16(x), lo = bf16(x - hi).  Two elements at a time.
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16x2(a, b);
  lo = pack_bf16x2(a - bf16_lo_as_f32(hi), b - bf16_hi_as_f32(hi));
}

// Same split with the bf16 rounding of `hi` done on the FMA pipe (Veltkamp: c = (2^16+1) x,
// hi = c - (c - x) is x rounded to 8 significant bits, exactly a bf16), so that only the `lo` pair
// needs a cvt (which shares the MUFU pipe).  Requires |x| < 2^111 (no overflow of c); used for
// activations and normalised inputs.
__device__ __forceinline__ void split_bf16x2_veltkamp(float a, float b, uint32_t& hi, uint32_t& lo) {
  const float ca = __fmul_rn(a, 65537.0f), cb = __fmul_rn(b, 65537.0f);
  const float ha = __fsub_rn(ca, __fsub_rn(ca, a)), hb = __fsub_rn(cb, __fsub_rn(cb, b));
  asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(hi) : "r"(__float_as_uint(ha)), "r"(__float_as_uint(hb)));
  lo = pack_bf16x2(__fsub_rn(a, ha), __fsub_rn(b, hb));
}

// ----------------------------------------------------------------------------- packed fp32x2 (FADD2 / FMUL2 / FFMA2)
// Two fp32 lanes per instruction: halves the issue slots of the epilogue's FMA-pipe work.
__device__ __forceinline__ uint64_t pk2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

#ifndef BBMPC_LO_CVT
#define BBMPC_LO_CVT 1   // 0: integer-pipe rounding (measured slower: 1.416 vs 1.375 ms per C4 rollout together with the FMA-pipe reciprocal)
#endif
// Veltkamp hi/lo bf16 split of a packed pair (see split_bf16x2_veltkamp): 4 packed ops + prmt + one cvt.
// Written as z = 2^16 y (exact), s = fl(y + z), hi = s - z (exact): ptxas contracts mul.rn.f32x2 + sub.rn.f32x2
// into FFMA2 (it does not for the scalar .rn forms); in the textbook form c = 65537 y, hi = c - (c - y) that
// contraction cancels the rounding (hi == y, lo == 0, measured), in this form every contraction is exact.
__device__ __forceinline__ void split_bf16x2_packed(uint64_t y, uint32_t& hi, uint32_t& lo) {
  const uint64_t z = mul2(y, pk2(65536.0f, 65536.0f));
  const uint64_t h = sub2(add2(y, z), z);
  const uint64_t l = sub2(y, h);
  float h0, h1, l0, l1;
  upk2(h, h0, h1); upk2(l, l0, l1);
  asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(hi) : "r"(__float_as_uint(h0)), "r"(__float_as_uint(h1)));
#if BBMPC_LO_CVT == 1
  lo = pack_bf16x2(l0, l1);
#elif BBMPC_LO_CVT == 2
  // residual TRUNCATED to bf16 (one PRMT, no MUFU-pipe cvt): |y - hi - lo| <= 2^-8 |lo| <= 2^-17 |y| instead of 2^-18
  asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(lo) : "r"(__float_as_uint(l0)), "r"(__float_as_uint(l1)));
#else
  // bf16 rounding of the residual on the integer pipe (round half away from zero: +0x8000 on the magnitude bits, keep
  // the high half): cvt.rn.bf16x2 runs on the MUFU pipe, which bounds the conversion warps (8 of 28 MUFU-pipe
  // operations per 16-column chunk).  Differs from round-to-nearest-even only on exact ties of the residual.
  asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(lo) : "r"(__float_as_uint(l0) + 0x8000u), "r"(__float_as_uint(l1) + 0x8000u));
#endif
}

}  // namespace tc05
