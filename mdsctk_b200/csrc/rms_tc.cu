// rms_tc.cu -- all-pairs superposed-RMSD sweep on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
// Same job as rms_simt.cu (the row blocks of knn_rms.cpp:231-293) with the nine per-pair dot
// products S_ab = sum_n x_na y_nb issued as a batched dense contraction:
//
//   work item    (fit tile of 128 frames, reference segment); persistent CTAs, one per SM.
//   CTA tile     128 fit frames (M) x 48 reference frames per accumulator pass.
//   operands     K-major tiles staged by TMA (cp.async.bulk.tensor.3d, SWIZZLE_64B, 64-byte rows)
//                from the frames x 3 x atoms SoA planes through a 3-stage mbarrier ring.
//                A_p = plane p (x|y|z) of the 128 fit frames, B = [x | y | z] planes of the 48
//                reference frames (144 rows).
//   MMA          tcgen05.mma.cta_group::1, M=128 N=144, issued by one elected thread.
//                Precision recovery by operand splitting (the pack kernel writes both parts):
//                  3xBF16: x ~= h + m (bf16), D += Ah*Bh + Ah*Bm + Am*Bh   kind::f16,  K=16, 32 atoms/stage
//                  3xTF32: x ~= h + l (tf32), D += Ah*Bh + Ah*Bl + Al*Bh   kind::tf32, K=8,  16 atoms/stage
//                  1xTF32: D += Ah*Bh only (coarse filter).
//   accumulators TMEM, 3 regions of 144 fp32 columns: D_p[lane q][b*48 + j] = S_pb(q, ref j).
//                TMEM lane = fit frame: an epilogue thread reads the nine S values of ITS fit row
//                with tcgen05.ld (no shuffles), solves QCP in registers (qcp.cuh) and appends the
//                survivors to the row's list; cursor and threshold live in shared memory
//                (select.cuh), thresholds are carried across reference segments through HBM.
//   roles        warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2-17 epilogue
//                (four warps per TMEM lane quarter, 12 of the 48 reference columns each; the QCP
//                chains are latency-bound, so the epilogue is spread over many warps).
// The accumulators are single-buffered (432 of 512 TMEM columns): the MMA of the next pass starts
// when the epilogue has READ the previous accumulators, and runs under the epilogue's arithmetic.
#include "common.cuh"
#include "qcp.cuh"
#include "select.cuh"

#include <cuda.h>
#include <cstdio>
#include <cstdlib>

namespace mdsctk {

namespace tc {
constexpr int TQ = 128;
constexpr int TR = 48;
constexpr int NST = 3;
constexpr int ROW_BYTES = 64;                 // one operand row per stage (SWIZZLE_64B)
constexpr int A_TILE = TQ * ROW_BYTES;        // 8192 B : one plane of the fit tile
constexpr int B_TILE = TR * ROW_BYTES;        // 3072 B : one plane of the reference tile
constexpr int OFF_AHI = 0;
constexpr int OFF_ALO = 3 * A_TILE;
constexpr int OFF_BHI = 6 * A_TILE;
constexpr int OFF_BLO = 6 * A_TILE + 3 * B_TILE;
constexpr int STAGE_BYTES = 6 * A_TILE + 6 * B_TILE;   // 67584
constexpr int SUBS = 4;                       // epilogue warps per TMEM lane quarter
constexpr int EPI_WARPS = 4 * SUBS;           // 16
constexpr int NTHR = 64 + EPI_WARPS * 32;     // 576
constexpr int UMMA_N = 3 * TR;                // 144
constexpr int TMEM_COLS = 512;
constexpr int SMEM_BYTES = NST * STAGE_BYTES + 1024;   // + alignment slack
constexpr int SUBW = TR / SUBS;               // 12 reference columns per epilogue warp
constexpr int EB = 4;                         // pairs per epilogue batch
constexpr int MERGE_EVERY = 4;                // passes between quarter-wide list merges (power of two)
constexpr int SUB_APP = 2 * MERGE_EVERY * SUBW;   // 96: private append area per epilogue warp and row
}  // namespace tc

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
        if (++polls > 4000000u) {
            printf("rms_sweep_tc: mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, blockIdx.x, threadIdx.x,
                   parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void quarter_sync(int quarter)
{
    asm volatile("bar.sync %0, 128;" ::"r"(quarter + 1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
template <bool BF16>
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    if constexpr (BF16) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
            : "memory");
    }
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, float (&v)[9][8], int c)
{
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
                 : "r"(taddr));
    v[c][0] = __uint_as_float(r0); v[c][1] = __uint_as_float(r1); v[c][2] = __uint_as_float(r2);
    v[c][3] = __uint_as_float(r3); v[c][4] = __uint_as_float(r4); v[c][5] = __uint_as_float(r5);
    v[c][6] = __uint_as_float(r6); v[c][7] = __uint_as_float(r7);
}
__device__ __forceinline__ void tc_ld4(uint32_t taddr, float (&v)[9][4], int c)
{
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(taddr));
    v[c][0] = __uint_as_float(r0); v[c][1] = __uint_as_float(r1); v[c][2] = __uint_as_float(r2);
    v[c][3] = __uint_as_float(r3);
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

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

struct TcArgs {
    const float *q_G, *r_G;
    long long q_begin, n_q, n_r;
    int A_pad, do_fit, n_seg;
    CandLists<float> cl;
    float *debug_tile;          // optional [128][9][48]: raw accumulators of (fit tile 0, ref tile 0)
    float *row_tau;             // [n_q] running admission threshold per fit row (+inf before the first segment)
    int dbg;                    // MDSCTK_TC_DEBUG bits: 1 skip QCP, 2 skip MMA issue, 4 skip TMA (timing experiments)
};

// MODE: 1 = 3xTF32, 2 = 1xTF32, 3 = 3xBF16
template <int MODE>
__global__ void __launch_bounds__(tc::NTHR, 1)
rms_sweep_tc_kernel(const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
                    const __grid_constant__ CUtensorMap map_r_hi, const __grid_constant__ CUtensorMap map_r_lo,
                    TcArgs a)
{
    using namespace tc;
    constexpr bool BF16 = MODE == 3;
    constexpr bool SPLIT = MODE != 2;
    constexpr int KC = BF16 ? 32 : 16;                 // atoms per stage (64-byte rows)
    constexpr int KSTEPS = 2;                          // UMMA_K = 32 bytes; two per 64-byte row
    constexpr uint32_t IDESC = umma_idesc(BF16 ? 1 : 2, TQ, UMMA_N);
    constexpr uint32_t STAGE_TX = SPLIT ? STAGE_BYTES : (3 * A_TILE + 3 * B_TILE);

    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[NST], bar_empty[NST], bar_tmem_full, bar_tmem_empty;
    __shared__ uint32_t s_tmem_base;
    __shared__ unsigned s_hist[EPI_WARPS][256];
    __shared__ int s_cnt[EPI_WARPS][32];
    __shared__ float s_tau[TQ];
    __shared__ int s_mcnt[TQ];

    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk = (a.A_pad + KC - 1) / KC;
    const long long n_qt = (a.n_q + TQ - 1) / TQ;
    const long long n_rt = (a.n_r + TR - 1) / TR;
    const long long n_items = n_qt * a.n_seg;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_tmem_full, 1);
        mbar_init(&bar_tmem_empty, EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                     "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    // item -> (fit tile, reference tile range); segment-major so that concurrently running CTAs
    // stream the same part of the reference set (L2 reuse)
    auto item_range = [&](long long it, long long &qt, long long &rt0, long long &rt1, int &seg) {
        seg = (int)(it / n_qt);
        qt = it - (long long)seg * n_qt;
        rt0 = n_rt * seg / a.n_seg;
        rt1 = n_rt * (seg + 1) / a.n_seg;
    };

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (long long it = blockIdx.x; it < n_items; it += gridDim.x) {
                long long qt, rt0, rt1; int seg;
                item_range(it, qt, rt0, rt1, seg);
                const int q0 = (int)(a.q_begin + qt * TQ);
                for (long long rt = rt0; rt < rt1; ++rt) {
                    const int r0 = (int)(rt * TR);
                    for (int kc = 0; kc < nk; ++kc) {
                        mbar_wait(&bar_empty[s], ph ^ 1, 1);
                        unsigned char *st = smem + s * STAGE_BYTES;
                        if (a.dbg & 4) { mbar_arrive(&bar_full[s]); if (++s == NST) { s = 0; ph ^= 1; } continue; }
                        mbar_expect_tx(&bar_full[s], STAGE_TX);
                        // one box = 64 bytes of atoms x rows frames x 3 planes, landing as [plane][frame][atoms]
                        tma_load_3d(st + OFF_AHI, &map_q_hi, &bar_full[s], kc * KC, q0, 0);
                        tma_load_3d(st + OFF_BHI, &map_r_hi, &bar_full[s], kc * KC, r0, 0);
                        if constexpr (SPLIT) {
                            tma_load_3d(st + OFF_ALO, &map_q_lo, &bar_full[s], kc * KC, q0, 0);
                            tma_load_3d(st + OFF_BLO, &map_r_lo, &bar_full[s], kc * KC, r0, 0);
                        }
                        if (++s == NST) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer =================================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0, tph = 0;
            bool first_pass = true;
            for (long long it = blockIdx.x; it < n_items; it += gridDim.x) {
                long long qt, rt0, rt1; int seg;
                item_range(it, qt, rt0, rt1, seg);
                for (long long rt = rt0; rt < rt1; ++rt) {
                    if (!first_pass) {  // the epilogue must have read the previous accumulators
                        mbar_wait(&bar_tmem_empty, tph, 2);
                        tph ^= 1;
                    }
                    first_pass = false;
                    tc_fence_after();
                    for (int kc = 0; kc < nk; ++kc) {
                        mbar_wait(&bar_full[s], ph, 3);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                        // atoms beyond A_pad are zero-filled by TMA; skip a k-step that is all padding
                        const int ksteps = (a.A_pad - kc * KC) * (BF16 ? 2 : 4) > 32 ? KSTEPS : 1;
                        for (int ks = 0; ks < ((a.dbg & 2) ? 0 : ksteps); ++ks) {
                            const uint32_t koff = ks * 32;
                            const uint64_t bhi = umma_desc_sw64(sa + OFF_BHI + koff);
                            const uint64_t blo = umma_desc_sw64(sa + OFF_BLO + koff);
#pragma unroll
                            for (int p = 0; p < 3; ++p) {
                                const uint32_t d = tmem_base + p * UMMA_N;
                                const uint64_t ahi = umma_desc_sw64(sa + OFF_AHI + p * A_TILE + koff);
                                tc_mma<BF16>(d, ahi, bhi, IDESC, (kc | ks) != 0);
                                if constexpr (SPLIT) {
                                    const uint64_t alo = umma_desc_sw64(sa + OFF_ALO + p * A_TILE + koff);
                                    tc_mma<BF16>(d, ahi, blo, IDESC, 1);
                                    tc_mma<BF16>(d, alo, bhi, IDESC, 1);
                                }
                            }
                        }
                        tc_commit(&bar_empty[s]);  // frees the smem stage once these MMAs have read it
                        if (++s == NST) { s = 0; ph ^= 1; }
                    }
                    tc_commit(&bar_tmem_full);     // accumulators complete -> epilogue
                }
            }
        }
    } else {
        // =============================== epilogue ===================================
        // List of a (fit row, reference segment): [0, keep) merged candidates, then one private append
        // area of SUB_APP entries per epilogue warp of the lane quarter (cursor in the warp's own
        // shared-memory slot), so the pass loop needs no barrier.  Every MERGE_EVERY passes the four
        // warps of a quarter meet; rows with a filling append area are merged, reduced to the `keep`
        // smallest by radix select and their admission threshold s_tau is lowered.  Pairs that
        // survive the straight-line filter are queued per warp and refined with all lanes busy.
        const int ew = warp - 2;                      // 0..15
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
        const int sub = ew >> 2;                      // which 12 reference columns of every tile
        const int e_of_quarter = (quarter + 2) & 3;   // ew = sub * 4 + e_of_quarter
        const int row_in_tile = quarter * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const unsigned lt_mask = (1u << lane) - 1u;
        unsigned *hist = s_hist[ew];                  // radix histogram during merges, refine queue otherwise
        float *q_c2 = reinterpret_cast<float *>(hist), *q_c1 = q_c2 + 32, *q_c0 = q_c2 + 64, *q_e0 = q_c2 + 96,
              *q_x = q_c2 + 128;
        int *q_ref = reinterpret_cast<int *>(hist) + 160, *q_own = reinterpret_cast<int *>(hist) + 192;
        int *wcnt = s_cnt[ew];                        // fill of this warp's append area of row (quarter*32 + l)
        const size_t row_stride = (size_t)a.cl.H * a.cl.cap;
        uint32_t tph = 0;
        int qn = 0;                                   // queue fill (warp-uniform)
        size_t lbase0 = 0;                            // list of the quarter's row 0 in the current item
        float *lkeys = a.cl.key;
        int *lidxs = a.cl.idx;

        auto drain = [&]() {
            __syncwarp();
            if (lane < qn) {
                const int own = q_own[lane];
                const int l = own & 31;
                const float e0 = q_e0[lane], x1 = q_x[lane];
                const float tau_r = *reinterpret_cast<volatile float *>(&s_tau[quarter * 32 + l]);
                float d2;
                if (own & 32) {                       // --nofit: already final
                    d2 = fmaxf(2.0f * (e0 - x1), 0.0f);
                } else {
                    QcpCoef c; c.c2 = q_c2[lane]; c.c1 = q_c1[lane]; c.c0 = q_c0[lane];
                    d2 = qcp_refine(c, e0, e0, x1, tau_r);
                }
                if (d2 < tau_r) {
                    const int pos = atomicAdd(&wcnt[l], 1);
                    if (pos < SUB_APP) {
                        const size_t at = lbase0 + (size_t)l * row_stride + a.cl.keep + sub * SUB_APP + pos;
                        lkeys[at] = d2;
                        lidxs[at] = q_ref[lane];
                    }
                }
            }
            qn = 0;
            __syncwarp();
        };
        auto push = [&](bool pred, float c2, float c1, float c0, float e0, float x1, int ridx, int own) {
            const unsigned m = __ballot_sync(0xffffffffu, pred);
            if (!m) return;
            const int n = __popc(m);
            if (qn + n > 32) drain();
            if (pred) {
                const int p = qn + __popc(m & lt_mask);
                q_c2[p] = c2; q_c1[p] = c1; q_c0[p] = c0; q_e0[p] = e0; q_x[p] = x1; q_ref[p] = ridx; q_own[p] = own;
            }
            qn += n;
        };
        // Quarter-wide merge of the rows this warp is responsible for (8 per warp).  final: every row.
        auto merge_rows = [&](long long qt, int seg, bool final) {
            quarter_sync(quarter);
            for (int r8 = 0; r8 < 8; ++r8) {
                const int l = sub * 8 + r8;
                const int row = quarter * 32 + l;
                const long long qr = qt * TQ + row;
                if (qr >= a.n_q) break;
                int cs[SUBS], cmax = 0;
#pragma unroll
                for (int s2 = 0; s2 < SUBS; ++s2) {
                    cs[s2] = min(s_cnt[s2 * 4 + e_of_quarter][l], SUB_APP);
                    cmax = max(cmax, cs[s2]);
                }
                if (!final && cmax < SUB_APP - MERGE_EVERY * SUBW) continue;   // cannot overflow before the next merge
                const size_t at = lbase0 + (size_t)l * row_stride;
                int total = s_mcnt[row];
#pragma unroll
                for (int s2 = 0; s2 < SUBS; ++s2) {
                    const size_t src = at + a.cl.keep + s2 * SUB_APP;
                    for (int base = 0; base < cs[s2]; base += 32) {    // dest <= src: forward move is safe
                        const int i = base + lane;
                        float kv = 0.f; int iv = 0;
                        if (i < cs[s2]) { kv = lkeys[src + i]; iv = lidxs[src + i]; }
                        __syncwarp();
                        if (i < cs[s2]) { lkeys[at + total + i] = kv; lidxs[at + total + i] = iv; }
                        __syncwarp();
                    }
                    total += cs[s2];
                }
                float tl = s_tau[row];
                if (total > a.cl.keep) {
                    tl = fminf(tl, warp_compact_list<float>(lkeys + at, lidxs + at, total, a.cl.keep, hist));
                    total = a.cl.keep;
                }
                __syncwarp();
                if (lane == 0) {
                    s_tau[row] = tl;
                    s_mcnt[row] = total;
#pragma unroll
                    for (int s2 = 0; s2 < SUBS; ++s2) s_cnt[s2 * 4 + e_of_quarter][l] = 0;
                    if (final) {
                        const size_t lid = (size_t)qr * a.cl.H + seg;
                        a.cl.cnt[lid] = total;
                        a.cl.tau[lid] = tl;
                        atomicMin(reinterpret_cast<unsigned *>(a.row_tau + qr), __float_as_uint(tl));  // tau >= +0
                    }
                }
            }
            quarter_sync(quarter);
        };

        for (long long it = blockIdx.x; it < n_items; it += gridDim.x) {
            long long qt, rt0, rt1; int seg;
            item_range(it, qt, rt0, rt1, seg);
            const long long qrow = qt * TQ + row_in_tile;        // row within the query range
            const bool qvalid = qrow < a.n_q;
            const float hgq = 0.5f * a.q_G[a.q_begin + (qvalid ? qrow : a.n_q - 1)];
            lbase0 = ((size_t)(qt * TQ + quarter * 32) * a.cl.H + seg) * a.cl.cap;
            wcnt[lane] = 0;
            if (sub == 0) {
                s_mcnt[row_in_tile] = 0;
                // admission threshold carried over from segments of this row that already finished
                s_tau[row_in_tile] = qvalid ? __ldcg(a.row_tau + qrow) : 0.0f;
            }
            quarter_sync(quarter);
            for (long long rt = rt0; rt < rt1; ++rt) {
                const long long r0 = rt * TR + sub * SUBW;
                float4 gv[SUBW / EB];
#pragma unroll
                for (int jb = 0; jb < SUBW / EB; ++jb) gv[jb] = __ldg(reinterpret_cast<const float4 *>(a.r_G + r0) + jb);
                if (lane == 0) mbar_wait(&bar_tmem_full, tph, 4);
                tph ^= 1;
                __syncwarp();
                tc_fence_after();
                const float tau = *reinterpret_cast<volatile float *>(&s_tau[row_in_tile]);
                const float htau = 0.5f * tau;
#pragma unroll
                for (int jb = 0; jb < SUBW / EB; ++jb) {
                    float sv[9][EB];
#pragma unroll
                    for (int p = 0; p < 3; ++p)
#pragma unroll
                        for (int b = 0; b < 3; ++b)
                            tc_ld4(t_lane + p * UMMA_N + b * TR + sub * SUBW + jb * EB, sv, p * 3 + b);
                    tc_wait_ld();
                    if (jb == SUBW / EB - 1) {  // last TMEM read of this pass: hand the accumulators back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bar_tmem_empty);
                    }
                    if (a.debug_tile && it == 0 && rt == 0) {
#pragma unroll
                        for (int c = 0; c < 9; ++c)
#pragma unroll
                            for (int j = 0; j < EB; ++j)
                                a.debug_tile[(size_t)row_in_tile * (9 * TR) + c * TR + sub * SUBW + jb * EB + j] = sv[c][j];
                    }
                    const float4 g = gv[jb];
                    const float e0[EB] = {fmaf(0.5f, g.x, hgq), fmaf(0.5f, g.y, hgq), fmaf(0.5f, g.z, hgq),
                                          fmaf(0.5f, g.w, hgq)};
                    const long long rb = r0 + jb * EB;
                    if (a.dbg & 1) {
                        float acc = 0.f;
#pragma unroll
                        for (int j = 0; j < EB; ++j) acc += sv[0][j] + sv[4][j] + sv[8][j] + e0[j];
                        if (acc == 12345.f) wcnt[lane] = 1;
                        continue;
                    }
                    if (!a.do_fit) {                  // --nofit: lambda = trace S
#pragma unroll
                        for (int j = 0; j < EB; ++j) {
                            const float tr = sv[0][j] + sv[4][j] + sv[8][j];
                            push(qvalid && rb + j < a.n_r && 2.0f * (e0[j] - tr) < tau, 0.f, 0.f, 0.f, e0[j], tr,
                                 (int)(rb + j), lane | 32);
                        }
                        continue;
                    }
                    // (1) cheap bound: lambda_max <= sqrt(3) |S|_F, so RMSD^2 >= 2 (E0 - sqrt(3 F)).  Far
                    //     pairs (other conformational basins) are rejected for 9 FMAs; a batch whose
                    //     128 pairs are all far skips the characteristic polynomial altogether.
                    float f[EB];
                    bool far = true;
#pragma unroll
                    for (int j = 0; j < EB; ++j) {
                        f[j] = qcp_frob2(sv, j);
                        const float t = e0[j] - htau;
                        far = far && (t > 0.0f) && (3.0001f * f[j] < t * t);
                    }
                    if (!(a.dbg & 8) && __all_sync(0xffffffffu, far)) continue;
                    // (2) QCP coefficients + one Newton step from E0: a valid lower bound on RMSD^2
                    QcpCoef c[EB];
                    float x1[EB];
#pragma unroll
                    for (int j = 0; j < EB; ++j) {
                        const float s9[9] = {sv[0][j], sv[1][j], sv[2][j], sv[3][j], sv[4][j], sv[5][j], sv[6][j], sv[7][j], sv[8][j]};
                        c[j] = qcp_coefficients(s9, f[j]);
                    }
#pragma unroll
                    for (int j = 0; j < EB; ++j) {
                        const float xn = qcp_newton_step(c[j], e0[j]);
                        x1[j] = (xn == xn) ? xn : e0[j];
                    }
                    // (3) survivors go to the warp's refine queue
#pragma unroll
                    for (int j = 0; j < EB; ++j)
                        push(qvalid && rb + j < a.n_r && !(2.0f * (e0[j] - x1[j]) > tau), c[j].c2, c[j].c1, c[j].c0, e0[j],
                             x1[j], (int)(rb + j), lane);
                }
                drain();
                // an append area takes at most SUBW entries per pass
                if ((((rt - rt0) & (MERGE_EVERY - 1)) == MERGE_EVERY - 1) && rt + 1 < rt1) merge_rows(qt, seg, false);
            }
            merge_rows(qt, seg, true);                // leaves one list of <= keep candidates per row
        }
    }

    // ---- teardown --------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

// ---------------------------------------------------------------- host side ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode()
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

// planes[n][3][A_pad] as a 3-D tensor ordered (atom, frame, plane); box = 64 bytes of atoms x `rows`
// frames x 3 planes, so one TMA op lands the three plane tiles back to back as [plane][frame][atoms].
static bool make_plane_map(CUtensorMap *m, const void *planes, long long n, int A_pad, int rows, bool bf16)
{
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    const int esz = bf16 ? 2 : 4;
    cuuint64_t dims[3] = {(cuuint64_t)A_pad, (cuuint64_t)n, 3};
    cuuint64_t strides[2] = {(cuuint64_t)A_pad * 3 * esz, (cuuint64_t)A_pad * esz};
    cuuint32_t box[3] = {(cuuint32_t)(tc::ROW_BYTES / esz), (cuuint32_t)rows, 3};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(planes),
               dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Reference segments per fit tile: enough work items for even waves over the SMs, but every
// segment keeps at least 16 reference tiles (its lists warm up once per segment).
int rms_tc_choose_segments(long long n_fit, long long n_ref, int n_sms)
{
    const long long n_qt = (n_fit + tc::TQ - 1) / tc::TQ;
    const long long n_rt = (n_ref + tc::TR - 1) / tc::TR;
    int best = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= 8; ++s) {
        if (s > 1 && n_rt / s < 16) break;
        const long long items = n_qt * s;
        const double eff = (double)items / (double)(((items + n_sms - 1) / n_sms) * n_sms);
        if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
    }
    return best;
}

int rms_tc_lists_per_segment() { return 1; }

// Entries reserved per (fit row, reference segment): keep merged + one append area per epilogue warp.
int rms_tc_list_stride(int keep) { return (keep + tc::SUBS * tc::SUB_APP + 31) / 32 * 32; }

cudaError_t launch_rms_sweep_tc(int mode, const FrameSetView &fit, const void *fit_hi, const void *fit_lo,
                                long long fit_begin, long long n_fit, const FrameSetView &ref, const void *ref_hi,
                                const void *ref_lo, int do_fit, int n_seg, CandLists<float> cl, float *row_tau,
                                float *debug_tile, int n_sms, cudaStream_t st)
{
    if (n_fit <= 0) return cudaSuccess;
    if (cl.H != n_seg) return cudaErrorInvalidValue;
    const bool bf16 = mode == 3;
    CUtensorMap mq_hi, mq_lo, mr_hi, mr_lo;
    if (!make_plane_map(&mq_hi, fit_hi, fit.n, fit.A_pad, tc::TQ, bf16) ||
        !make_plane_map(&mq_lo, fit_lo, fit.n, fit.A_pad, tc::TQ, bf16) ||
        !make_plane_map(&mr_hi, ref_hi, ref.n, ref.A_pad, tc::TR, bf16) ||
        !make_plane_map(&mr_lo, ref_lo, ref.n, ref.A_pad, tc::TR, bf16))
        return cudaErrorInvalidValue;
    TcArgs a;
    a.q_G = fit.G; a.r_G = ref.G; a.q_begin = fit_begin; a.n_q = n_fit; a.n_r = ref.n;
    a.A_pad = ref.A_pad; a.do_fit = do_fit; a.n_seg = n_seg; a.cl = cl; a.debug_tile = debug_tile; a.row_tau = row_tau;
    const char *dbg = getenv("MDSCTK_TC_DEBUG");
    a.dbg = dbg ? atoi(dbg) : 0;
    const long long n_items = ((n_fit + tc::TQ - 1) / tc::TQ) * n_seg;
    const unsigned grid = (unsigned)(n_items < n_sms ? n_items : n_sms);
    cudaError_t e;
#define MDSCTK_LAUNCH_TC(M)                                                                                           \
    e = cudaFuncSetAttribute(rms_sweep_tc_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES);   \
    if (e != cudaSuccess) return e;                                                                                  \
    rms_sweep_tc_kernel<M><<<grid, tc::NTHR, tc::SMEM_BYTES, st>>>(mq_hi, mq_lo, mr_hi, mr_lo, a)
    switch (mode) {
    case 1: MDSCTK_LAUNCH_TC(1); break;
    case 2: MDSCTK_LAUNCH_TC(2); break;
    case 3: MDSCTK_LAUNCH_TC(3); break;
    default: return cudaErrorInvalidValue;
    }
#undef MDSCTK_LAUNCH_TC
    return cudaGetLastError();
}

}  // namespace mdsctk
