// tc_ptx.cuh -- thin inline-PTX wrappers shared by the tcgen05 kernels (rms_tc.cu, data_tc.cu):
// mbarrier, TMA (cp.async.bulk.tensor, CTA-pair form), tcgen05 MMA / commit / ld, UMMA descriptors.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>

namespace mdsctk {

// ---------------------------------------------------------------- PTX wrappers ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Strictly non-blocking probe (try_wait may suspend the thread for a system-dependent time when the phase is incomplete)
__device__ __forceinline__ bool mbar_test_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must trap, not hang the GPU.  try_wait suspends the thread in
// hardware for up to the hint (ns) per poll, so the loop costs few issue slots.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, int tag)
{
    uint32_t polls = 0;
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
            : "memory");
        if (ok) return;
        if (++polls > 400000u) {
            printf("tcgen05 sweep: mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, blockIdx.x, threadIdx.x,
                   parity);
            __trap();
        }
    }
}
// Latency-critical variant for the MMA issuer: plain try_wait polling (no suspend hint).
__device__ __forceinline__ void mbar_wait_spin(uint64_t *bar, uint32_t parity, int tag)
{
    uint32_t polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++polls > 200000000u) {
            printf("tcgen05 sweep: mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, blockIdx.x, threadIdx.x, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void quarter_sync(int quarter)
{
    asm volatile("bar.sync %0, 128;" ::"r"(quarter + 1) : "memory");
}
// quarter_sync that also returns the OR of `pred` over the quarter's 128 threads (one barrier, no shared-memory flag)
__device__ __forceinline__ bool quarter_any(int quarter, bool pred)
{
    uint32_t r;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.or.pred p, %2, 128, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(r) : "r"((uint32_t)pred), "r"(quarter + 1) : "memory");
    return r != 0;
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// CTA-pair copy: lands in THIS CTA's shared memory, completes on the mbarrier `bar_cluster_addr`
// (a shared::cluster address; the leader's "full" barrier for both CTAs of the pair).  `policy` is
// an L2 eviction-priority descriptor (kEvictLast for the fit tile, re-read every pass).
constexpr uint64_t kEvictNormal = 0x1000000000000000ull, kEvictFirst = 0x12F0000000000000ull, kEvictLast = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_3d_2sm(void *dst, const CUtensorMap *map, uint32_t bar_cluster_addr, int c0, int c1,
                                                int c2, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
        : "memory");
}
// L2 prefetch of a tensor box (no shared-memory destination, no barrier): the sweep's producer warms the reference
// tile it will stream a few passes later, so the ring's copies hit L2 instead of paying an HBM round trip each
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *map, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *map, int c0, int c1)
{
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void *dst, const CUtensorMap *map, uint32_t bar_cluster_addr, int c0, int c1,
                                                uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}
// shared::cluster address of the same variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(const void *p, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
// Hand-off of TMEM accumulators to the MMA issuer of the leader CTA: the reads are ordered by
// tcgen05.wait::ld + tcgen05.fence::before_thread_sync, no global/shared data is published, so the
// arrive needs no cluster-scope release fence (measured: the MEMBAR of the .release.cluster form was 11 % of
// the epilogue's stall samples, on the critical path of every pass).  Same form as cutlass ClusterBarrier::arrive.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t bar_cluster_addr)
{
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_nofence(uint32_t bar_cluster_addr)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
// one lane of a converged warp (the others fall through); keeps the operands of the guarded
// instruction in uniform registers instead of a per-lane "waterfall" loop
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// pair MMA bookkeeping: arrives on the mbarrier at the same offset in every CTA of `mask` once the
// cta_group::2 MMAs issued so far are done
__device__ __forceinline__ void tc_commit2_mc(uint64_t *bar, uint16_t mask)
{
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"(mask)
        : "memory");
}
// D[256 x N] (+)= A[256 x K] * B[N x K]^T over the CTA pair; descriptors are shared-memory offsets
// valid in both CTAs, d_tmem likewise
template <bool BF16>
__device__ __forceinline__ void tc_mma2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    if constexpr (BF16) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
            : "memory");
    }
}
// Same with the descriptors given as their LOW words (start address >> 4 | LBO field); the high word of
// every SWIZZLE_64B K-major descriptor here is the constant kDescHiSw64, so stepping through a tile is one
// 32-bit add per operand instead of rebuilding the 64-bit descriptor.  ACC: accumulate flag known at compile time.
constexpr uint32_t kDescHiSw64 = (uint32_t)((512u >> 4) | (1u << 14) | (4u << 29));   // SBO=512 B, version 1, SWIZZLE_64B
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
// ... and of every SWIZZLE_128B K-major one (128-byte rows, 8-row groups 1024 B apart; a 16-element fp16 k-step = +32 B)
constexpr uint32_t kDescHiSw128 = (uint32_t)((1024u >> 4) | (1u << 14) | (2u << 29));   // SBO=1024 B, version 1, SWIZZLE_128B
// fp16 CTA-pair MMA from descriptor low words with the shared high word given
__device__ __forceinline__ void tc_mma2_f16_lo_hi(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(hi)
        : "memory");
}
template <bool BF16>
__device__ __forceinline__ void tc_mma2_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc)
{
    if constexpr (BF16) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "mov.b64 da, {%1, %5};\n\t"
            "mov.b64 db, {%2, %5};\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}"
            ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(kDescHiSw64)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "mov.b64 da, {%1, %5};\n\t"
            "mov.b64 db, {%2, %5};\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %3, p;\n\t}"
            ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(kDescHiSw64)
            : "memory");
    }
}
// A operand in TENSOR MEMORY (the .ts form): a_tmem = TMEM address of this k-step's [128 lanes x 8 columns] block of
// packed 16-bit pairs (lane = row of A in each CTA of the pair, column j = K elements 2j, 2j+1), B from shared memory.
__device__ __forceinline__ void tc_mma2_ts_lo(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(acc), "r"(kDescHiSw64)
        : "memory");
}
// One shared-memory stage of the resident-tile sweep (rms_tc2.cu) as ONE straight-line block: two k-steps x three
// fit planes = six M=256 N=144 MMAs.  a0..a2: descriptor low words of the planes' A tiles at the first k-step (the second
// is 32 bytes further, + 2), b: of the reference half-tile; d0: accumulator region of plane 0 (planes 1, 2 at + 144,
// + 288 columns); acc0 = 0 starts a new accumulation.  Everything the tensor pipe needs differs from these by constants,
// so the issuing thread spends ~2 instructions per MMA (it has 72 clk per MMA before it becomes the bottleneck).
// bhi: high word of the B descriptor (kDescHiSw64 for 64-byte reference rows, kDescHiSw128 for 128-byte ones); A is SWIZZLE_64B.
__device__ __forceinline__ void tc2_issue_stage_ss(uint32_t d0, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t b, uint32_t idesc,
                                                   uint32_t acc0, uint32_t bhi)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .b32 hi, d1, d2, x0, x1, x2, y;\n\t"
        ".reg .b64 da, db;\n\t"
        "mov.b32 hi, 0x80004020;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.u32 q, %0, %0;\n\t"
        "add.u32 d1, %0, 144;\n\t"
        "add.u32 d2, %0, 288;\n\t"
        "add.u32 y, %4, 2;\n\t"
        "add.u32 x0, %1, 2;\n\t"
        "add.u32 x1, %2, 2;\n\t"
        "add.u32 x2, %3, 2;\n\t"
        "mov.b64 db, {%4, %7};\n\t"
        "mov.b64 da, {%1, hi};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t"
        "mov.b64 da, {%2, hi};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [d1], da, db, %5, p;\n\t"
        "mov.b64 da, {%3, hi};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [d2], da, db, %5, p;\n\t"
        "mov.b64 db, {y, %7};\n\t"
        "mov.b64 da, {x0, hi};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, q;\n\t"
        "mov.b64 da, {x1, hi};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [d1], da, db, %5, q;\n\t"
        "mov.b64 da, {x2, hi};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [d2], da, db, %5, q;\n\t"
        "}"
        ::"r"(d0), "r"(a0), "r"(a1), "r"(a2), "r"(b), "r"(idesc), "r"(acc0), "r"(bhi)
        : "memory");
}
// Same with the three A operands in tensor memory (t0..t2: TMEM addresses of the first k-step, the second 8 columns
// further); two = 0: only the first k-step exists (odd number of k-steps).
__device__ __forceinline__ void tc2_issue_stage_ts(uint32_t d0, uint32_t t0, uint32_t t1, uint32_t t2, uint32_t b, uint32_t idesc,
                                                   uint32_t acc0, uint32_t two, uint32_t bhi)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .b32 hi, d1, d2, x0, x1, x2, y;\n\t"
        ".reg .b64 db;\n\t"
        "mov.b32 hi, 0x80004020;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.ne.b32 q, %7, 0;\n\t"
        "add.u32 d1, %0, 144;\n\t"
        "add.u32 d2, %0, 288;\n\t"
        "add.u32 y, %4, 2;\n\t"
        "add.u32 x0, %1, 8;\n\t"
        "add.u32 x1, %2, 8;\n\t"
        "add.u32 x2, %3, 8;\n\t"
        "mov.b64 db, {%4, %8};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %5, p;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [d1], [%2], db, %5, p;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [d2], [%3], db, %5, p;\n\t"
        "mov.b64 db, {y, %8};\n\t"
        "setp.eq.u32 p, %0, %0;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], [x0], db, %5, p;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [d1], [x1], db, %5, p;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [d2], [x2], db, %5, p;\n\t"
        "}"
        ::"r"(d0), "r"(t0), "r"(t1), "r"(t2), "r"(b), "r"(idesc), "r"(acc0), "r"(two), "r"(bhi)
        : "memory");
}
// eight consecutive 32-bit columns of this thread's TMEM lane <- registers (operand staging; complete after tc_wait_st)
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&v)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// cluster-scope acquire wait: pairs with mbarrier.arrive.release.cluster of the peer CTA (data published through
// distributed shared memory before the arrive is visible after the wait)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity, int tag)
{
    uint32_t polls = 0;
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
        if (++polls > 200000000u) {
            printf("tcgen05 sweep: mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, blockIdx.x, threadIdx.x, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void st_cluster_u32(uint32_t cluster_addr, uint32_t v)
{
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}

// four consecutive accumulator columns of this thread's TMEM lane -> v[c][0..3] (valid after tc_wait_ld)
__device__ __forceinline__ void tc_ld4(uint32_t taddr, float (&v)[9][4], int c)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v[c][0]), "=f"(v[c][1]), "=f"(v[c][2]), "=f"(v[c][3])
                 : "r"(taddr));
}
// two consecutive accumulator columns -> v[c][0..1]
__device__ __forceinline__ void tc_ld2(uint32_t taddr, float (&v)[9][2], int c)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=f"(v[c][0]), "=f"(v[c][1]) : "r"(taddr));
}
// sixteen consecutive accumulator columns of this thread's TMEM lane (valid after tc_wait_ld)
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
                   "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// tc_wait_ld that names the sixteen registers as read-write operands: in a software-pipelined epilogue (the next batch's
// tcgen05.ld issued before this batch's arithmetic) nothing else stops the compiler from scheduling a use above the wait
__device__ __forceinline__ void tc_wait_ld16(float (&v)[16])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]),
                   "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])
                 :: "memory");
}

// K-major, SWIZZLE_64B shared-memory operand descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (8 rows x 64 B = 512 B)
//   [46,48) version=1 | [61,64) layout type 4 = SWIZZLE_64B
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 (1<<4), A/B format at [7,10)/[10,13)
// (BF16 = 1, TF32 = 2), both K-major, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t umma_idesc(int fmt, int M, int N)
{
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Driver entry point for tensor-map encoding (no -lcuda link dependency).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_tensor_map_encoder()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

}  // namespace mdsctk
