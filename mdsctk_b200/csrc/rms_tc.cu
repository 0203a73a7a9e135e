// rms_tc.cu -- all-pairs superposed-RMSD sweep on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
// Same job as rms_simt.cu (the row blocks of knn_rms.cpp:231-293) with the nine per-pair dot
// products S_ab = sum_n x_na y_nb issued as a batched dense contraction:
//
//   work item    (fit super-tile of 256 frames, reference segment), taken by a CTA PAIR (cluster of 2,
//                tcgen05 cta_group::2); persistent, one CTA per SM.  CTA r of the pair owns fit rows
//                [128 r, 128 r + 128) of the super-tile and loads HALF of every reference tile; the
//                pair's MMA (M = 256) reads the other half from the peer's shared memory, which is
//                what keeps a single SM's shared-memory read bandwidth from capping the tensor pipe
//                (measured: one-SM M=128 MMAs with both operands in shared memory run at ~82 B/clk of
//                operand reads, i.e. 104 instead of 72 clk for N=144).
//   pass         256 fit frames x 48 reference frames.
//   operands     K-major tiles staged by TMA (cp.async.bulk.tensor.3d, SWIZZLE_64B, 64-byte rows)
//                from the frames x 3 x atoms SoA planes through a 3-stage mbarrier ring; both CTAs'
//                copies complete on the leader's "full" barrier.
//                A_p = plane p (x|y|z) of the fit frames; B = [x|y|z of refs 0-23 | x|y|z of refs 24-47]
//                (144 rows, the first 72 in CTA 0, the last 72 in CTA 1).
//   MMA          tcgen05.mma.cta_group::2, M=256 N=144, issued by one elected lane of the leader CTA.
//                Precision recovery by operand splitting (the pack kernel writes both parts):
//                  3xBF16: x ~= h + m (bf16), D += Ah*Bh + Ah*Bm + Am*Bh   kind::f16,  K=16, 32 atoms/stage
//                  3xTF32: x ~= h + l (tf32), D += Ah*Bh + Ah*Bl + Al*Bh   kind::tf32, K=8,  16 atoms/stage
//                  1xTF32: D += Ah*Bh only (coarse filter).
//                  3xFP16: as 3xBF16 on an fp16 split of 64 x (22 bits kept instead of 16)
//                  2xFP16: D += Ah*Bh + Ah*Bl, fit operand rounded to fp16 (11 bits): two MMAs and
//                          half the fit-operand traffic
//                  1xFP16: D += Ah*Bh only: ONE MMA per k-step.  These two modes compute, up to fp32
//                          accumulation noise, the exact min-RMSD between the ROUNDED structures
//                          (E0 from Gh / G2, the norms of what the tensor cores see); min-RMSD over
//                          rotations is a metric, so it is within g_q + g_r (the rounding residual
//                          norms, pack.cu) of the true distance -- the re-score certificate uses
//                          that bound instead of an empirical noise constant (rms_rescore.cu)
//   residency    1xFP16 only (one 16-bit part per operand), atoms <= 352: the y and z planes of the CTA's fit
//                tile (2 x 8 KB per 32-atom chunk, 160 KB at 300 atoms) are loaded ONCE per work item and
//                stay in shared memory; the ring then carries only the x plane of the fit tile and the
//                reference half-tile (12.5 KB per stage instead of 28.5 KB), which takes the L2->SM operand
//                stream (the limiter once a pass needs a third of the MMAs) from 291 to 131 KB per pass.
//   accumulators TMEM of each CTA: its 128 fit rows x 3 regions of 144 fp32 columns (432 of 512);
//                TMEM lane = fit frame, so an epilogue thread reads the nine S values of ITS fit row
//                with tcgen05.ld (no shuffles), bounds RMSD^2 from below (Frobenius bound, then QCP
//                coefficients + one Newton step, qcp.cuh) and queues the survivors.
//   roles        warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2-17 epilogue
//                (four warps per TMEM lane quarter, 12 of the 48 reference columns each).
#include "common.cuh"
#include "qcp.cuh"
#include "select.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>
#include <cstdio>
#include <cstdlib>

#ifndef MDSCTK_TC_PROF_BUILD
#define MDSCTK_TC_PROF_BUILD 0        // 1: compile the clock counters read by MDSCTK_TC_PROF=1 (costs registers)
#endif

namespace mdsctk {

namespace tc {
constexpr int TQ = 128;                       // fit frames per CTA
constexpr int TR = 48;                        // reference frames per pass (24 loaded by each CTA)
constexpr int TRH = TR / 2;
constexpr int ROW_BYTES = 64;                 // one operand row per stage (SWIZZLE_64B)
constexpr int UMMA_M = 2 * TQ;                // 256 across the pair
constexpr int UMMA_N = 3 * TR;                // 144
constexpr int A_TILE = TQ * ROW_BYTES;        // 8192 B : one plane of the CTA's fit tile
constexpr int A_PART = 3 * A_TILE;            // 24576 B: x|y|z planes of one split part
constexpr int B_PART = 3 * TRH * ROW_BYTES;   // 4608 B : this CTA's 72 of the 144 reference operand rows
constexpr int MAX_NST = 8;
constexpr int RES_CHUNK = 2 * A_TILE;         // 16384 B: resident y|z planes of one 32-atom chunk (fp16)
constexpr int RES_STAGE = A_TILE + B_PART;    // 12800 B: ring stage in resident mode (fit x plane + reference half-tile)
constexpr int RES_MAX_NK = 11;                // resident chunks that still leave room for a 2-stage ring
constexpr int STATIC_SMEM = 20 * 1024;        // bound on the kernel's static shared memory (checked at launch)

// MODE: 1 = 3xTF32, 2 = 1xTF32, 3 = 3xBF16, 4 = 3xFP16, 5 = 2xFP16 (fit operand: hi part only), 6 = 1xFP16
template <int MODE> struct Mode {
    static constexpr bool K16 = MODE >= 3;                         // 16-bit operands: K = 16 per MMA, 32 atoms per stage
    static constexpr int FMT = MODE == 3 ? 1 : (MODE >= 4 ? 0 : 2); // instruction-descriptor operand format
    static constexpr bool A_LO = MODE == 1 || MODE == 3 || MODE == 4;
    static constexpr bool B_LO = MODE != 2 && MODE != 6;
    static constexpr int OFF_AHI = 0;
    static constexpr int OFF_ALO = A_PART;
    static constexpr int OFF_BHI = (A_LO ? 2 : 1) * A_PART;
    static constexpr int OFF_BLO = OFF_BHI + B_PART;
    static constexpr int STAGE_BYTES = OFF_BHI + (B_LO ? 2 : 1) * B_PART;   // 58368 / 33792 / 29184
    static constexpr int NST_FIT = (227 * 1024 - STATIC_SMEM - 1024) / STAGE_BYTES;
    static constexpr int NST = NST_FIT < MAX_NST ? NST_FIT : MAX_NST;       // 3 / 6 / 6
    static constexpr int SMEM_BYTES = NST * STAGE_BYTES + 1024;             // + alignment slack
    static constexpr float SCALE = MODE >= 4 ? kRmsHalfScale * kRmsHalfScale : 1.0f;   // accumulators hold SCALE * S
};
constexpr int SUBS = 4;                       // epilogue warps per TMEM lane quarter
constexpr int EPI_WARPS = 4 * SUBS;           // 16
constexpr int NTHR = 64 + EPI_WARPS * 32;     // 576
constexpr int TMEM_COLS = 512;
constexpr int SUBW = TR / SUBS;               // 12 reference columns per epilogue warp
constexpr int EB = 2;                         // pairs per lane and epilogue batch (two batches in flight: TMEM reads of the next one run under the arithmetic of this one)
constexpr int MERGE_EVERY = 4;                // passes between quarter-wide list merges (power of two)
constexpr int SUB_APP = 2 * MERGE_EVERY * SUBW;   // 96: private append area per epilogue warp and row
}  // namespace tc

struct TcArgs {
    const float *q_G, *r_G;
    long long q_begin, n_q, n_r;
    int A_pad, do_fit, n_seg;
    CandLists<float> cl;        // H = n_seg lists per fit row
    float *debug_tile;          // optional [128][9][48]: raw accumulators of (fit tile 0, ref tile 0)
    float *row_tau;             // [n_q] running admission threshold per fit row (+inf before the first segment)
    const float4 *q_sig, *r_sig;   // singular values of the weighted frames (pack.cu): von Neumann pre-bound
    float pre_rel, pre_sqrt_gmax;  // operand rounding allowance of the pre-bound: g <= pre_rel (sqrt(Gq) + sqrt(max Gr)); < 0: off
    const int *own_tile;        // out-of-sample queries: per fit super-tile, the reference tile to start from (else NULL)
    int res, res_nst;           // resident fit planes (1xFP16): on/off, ring stages that fit beside them
    // MDSCTK_TC_DEBUG bits (timing experiments; results are only valid with 0, 16-128, 256, 512, 1024, 2048, 4096, 16384, 131072):
    //   1 skip QCP   2 skip MMA issue   4 skip TMA   8 no Frobenius batch skip   16 / 32 eviction hints off / evict-first
    //   64 suspending waits in the issuer   128 spinning waits in the epilogue   256 resident fit planes (1xFP16)
    //   512 release.cluster arrive on the hand-back   1024 read every batch (pre-bound computed, not used)
    //   2048 pre-bound off   4096 spinning waits in the producer   8192 / 65536 alternative clock counters (prof build)
    //   16384 relaxed.cluster arrive   32768 no start-tile guess for out-of-sample queries   131072 forward the peer
    //   CTA's hand-back through one remote arrival
    int dbg;
    long long *prof;            // MDSCTK_TC_PROF=1: [grid][8] clock sums (see launch_rms_sweep_tc)
};

// Launched as clusters of 2 CTAs.  With the FP16 modes everything downstream of the accumulators
// (E0, thresholds, Newton iterates) is kept in units of SCALE nm^2 and converted when a candidate
// is stored.
template <int MODE>
// 96 registers: 576 threads x 112 is refused at launch ("too many resources") although it is below 64 K
__global__ void __launch_bounds__(tc::NTHR, 1)
rms_sweep_tc_kernel(const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
                    const __grid_constant__ CUtensorMap map_r_hi, const __grid_constant__ CUtensorMap map_r_lo,
                    TcArgs a)
{
    using namespace tc;
    using MD = Mode<MODE>;
    constexpr bool BF16 = MD::K16;                     // 16-bit operand kinds (bf16 / fp16) share kind::f16
    constexpr bool A_LO = MD::A_LO, B_LO = MD::B_LO;
    constexpr int NST = MD::NST, STAGE_BYTES = MD::STAGE_BYTES;
    constexpr int OFF_AHI = MD::OFF_AHI, OFF_ALO = MD::OFF_ALO, OFF_BHI = MD::OFF_BHI, OFF_BLO = MD::OFF_BLO;
    constexpr float SC = MD::SCALE, INV_SC = 1.0f / MD::SCALE;
    constexpr int KC = BF16 ? 32 : 16;                 // atoms per stage (64-byte rows)
    constexpr int KSTEPS = 2;                          // UMMA_K = 32 bytes; two per 64-byte row
    constexpr uint32_t IDESC = umma_idesc(MD::FMT, UMMA_M, UMMA_N);
    constexpr uint32_t STAGE_TX = 2u * STAGE_BYTES;    // both CTAs' bytes land on the leader's barrier

    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[MAX_NST], bar_empty[MAX_NST], bar_tmem_full, bar_tmem_empty, bar_res_full,
        bar_res_empty, bar_epi_done;
    __shared__ uint32_t s_tmem_base;
    __shared__ long long s_tcommit;      // clock-counter build: when the issuer committed the pass (leader CTA)
    __shared__ unsigned s_hist[EPI_WARPS][256];
    __shared__ int s_cnt[EPI_WARPS][32];
    __shared__ float s_tau[TQ];
    __shared__ int s_mcnt[TQ];

    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();           // 0 = leader (issues the pair's MMAs)
    const int nk = (a.A_pad + KC - 1) / KC;
    // resident mode: smem = [nk chunks x (y|z planes)] [ring of nst stages x (x plane | reference half-tile)]
    const bool res = (MODE == 6) && a.res != 0;
    const int nst = res ? a.res_nst : NST;
    const int stage_bytes = res ? RES_STAGE : STAGE_BYTES;
    const int ring_off = res ? nk * RES_CHUNK : 0;
    const int off_b = res ? A_TILE : OFF_BHI;          // reference operand inside a stage
    const long long n_qt = (a.n_q + UMMA_M - 1) / UMMA_M;          // fit super-tiles of 256 rows
    const long long n_rt = (a.n_r + TR - 1) / TR;
    const long long n_items = n_qt * a.n_seg;
    const long long pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_tmem_full, 1);
        mbar_init(&bar_res_full, 1);
        mbar_init(&bar_res_empty, 1);
        // accumulator hand-back (used on the leader only): the epilogue warps of BOTH CTAs arrive on it.  With
        // MDSCTK_TC_DEBUG bit 131072 the peer's warps arrive on their own bar_epi_done and its otherwise idle warp 1
        // forwards ONE remote arrival (measured: no difference, 79.9 vs 81.5 ms -- the 16 remote arrivals are not
        // what the hand-back costs).
        mbar_init(&bar_tmem_empty, (a.dbg & 131072) ? EPI_WARPS + 1 : 2 * EPI_WARPS);
        mbar_init(&bar_epi_done, EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                     "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                  // the peer's mbarriers and TMEM exist before anything is signalled to them
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    // item -> (fit super-tile, reference segment).  Diagonal first: the first n_qt items pair every fit
    // super-tile with the reference segment that holds its own frames (trajectory frames close in
    // time are close in space), so a row's admission threshold is tight before the other segments
    // run; those then reject almost every pair on the cheap bounds.  The remaining items are
    // segment-major so that concurrently running pairs stream the same part of the reference set.
    // Within the diagonal segment the tile order is rotated to start at the row's own tile.
    auto item_range = [&](long long it, long long &qt, long long &rt0, long long &rt1, int &seg, long long &rot) {
        int s_rest = -1;
        if (it < n_qt) qt = it;
        else { s_rest = (int)((it - n_qt) / n_qt); qt = (it - n_qt) - (long long)s_rest * n_qt; }
        // reference tile of the first fit frame (fit rows that ARE reference frames), else the guess of
        // rms_guess_own_tile_kernel
        const long long own = min(a.own_tile ? (long long)a.own_tile[qt] : (a.q_begin + qt * UMMA_M) / TR, n_rt - 1);
        int sd = (int)(own * a.n_seg / n_rt);
        while (sd + 1 < a.n_seg && n_rt * (sd + 1) / a.n_seg <= own) ++sd;
        while (sd > 0 && n_rt * sd / a.n_seg > own) --sd;
        seg = s_rest < 0 ? sd : (s_rest < sd ? s_rest : s_rest + 1);
        rt0 = n_rt * seg / a.n_seg;
        rt1 = n_rt * (seg + 1) / a.n_seg;
        rot = s_rest < 0 ? own - rt0 : 0;
    };
    // i-th reference tile of an item
    auto tile_at = [](long long i, long long rt0, long long rt1, long long rot) {
        const long long t = rt0 + rot + i;
        return t < rt1 ? t : t - (rt1 - rt0);
    };

    if (warp == 0) {
        // =============================== TMA producer ===============================
        // converged warp; one elected lane issues.  Each CTA loads its own fit rows and its half of
        // the reference tile; every copy completes on the LEADER's full barrier.
        int s = 0;
        uint32_t ph = 0, rph = 0;
        bool first_item = true;
        const uint64_t pol_a = (a.dbg & 16) ? kEvictNormal : kEvictLast, pol_b = (a.dbg & 32) ? kEvictFirst : kEvictNormal;
        for (long long it = pair_id; it < n_items; it += n_pairs) {
            long long qt, rt0, rt1, rot; int seg;
            item_range(it, qt, rt0, rt1, seg, rot);
            const int q0 = (int)(a.q_begin + qt * UMMA_M + rank * TQ);
            if (res) {
                // the y|z planes of this item's fit tile, once: the previous item's MMAs must be done with theirs
                if (!first_item) { mbar_wait(&bar_res_empty, rph, 5); rph ^= 1; }
                first_item = false;
                const uint32_t res_leader = map_to_cta(&bar_res_full, 0);
                if (elect_one()) {
                    if (a.dbg & 4) {
                        if (rank == 0) mbar_arrive(&bar_res_full);
                    } else {
                        if (rank == 0) mbar_expect_tx(&bar_res_full, 2u * (uint32_t)(nk * RES_CHUNK));
                        for (int kc = 0; kc < nk; ++kc)          // map_q_lo: the single-plane box of the fit hi planes
                            for (int p = 1; p < 3; ++p)
                                tma_load_3d_2sm(smem + kc * RES_CHUNK + (p - 1) * A_TILE, &map_q_lo, res_leader, kc * KC, q0, p, pol_a);
                    }
                }
                __syncwarp();
            }
            for (long long ti = 0; ti < rt1 - rt0; ++ti) {
                const int r0 = (int)(tile_at(ti, rt0, rt1, rot) * TR + rank * TRH);
                for (int kc = 0; kc < nk; ++kc) {
                    if (a.dbg & 4096) mbar_wait_spin(&bar_empty[s], ph ^ 1, 1); else mbar_wait(&bar_empty[s], ph ^ 1, 1);   // the pair's MMAs have read stage s (both CTAs)
                    unsigned char *st = smem + ring_off + s * stage_bytes;
                    const uint32_t full_leader = map_to_cta(&bar_full[s], 0);
                    if (elect_one()) {
                        if (a.dbg & 4) {
                            if (rank == 0) mbar_arrive(&bar_full[s]);
                        } else if (res) {
                            if (rank == 0) mbar_expect_tx(&bar_full[s], 2u * RES_STAGE);
                            tma_load_3d_2sm(st, &map_q_lo, full_leader, kc * KC, q0, 0, pol_a);            // fit x plane
                            tma_load_3d_2sm(st + A_TILE, &map_r_hi, full_leader, kc * KC, r0, 0, pol_b);   // reference half-tile
                        } else {
                            if (rank == 0) mbar_expect_tx(&bar_full[s], STAGE_TX);
                            // one box = 64 bytes of atoms x rows frames x 3 planes, landing as [plane][frame][atoms]
                            tma_load_3d_2sm(st + OFF_AHI, &map_q_hi, full_leader, kc * KC, q0, 0, pol_a);
                            tma_load_3d_2sm(st + OFF_BHI, &map_r_hi, full_leader, kc * KC, r0, 0, pol_b);
                            if constexpr (A_LO) tma_load_3d_2sm(st + OFF_ALO, &map_q_lo, full_leader, kc * KC, q0, 0, pol_a);
                            if constexpr (B_LO) tma_load_3d_2sm(st + OFF_BLO, &map_r_lo, full_leader, kc * KC, r0, 0, pol_b);
                        }
                    }
                    __syncwarp();
                    if (++s == nst) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer (leader CTA) ====================
        // converged warp, descriptors in uniform registers, one elected lane issues
        if (rank == 0) {
            int s = 0;
            uint32_t ph = 0, tph = 0, rfph = 0;
            bool first_pass = true;
            long long t_wait_empty = 0, t_wait_full = 0, t_total0 = MDSCTK_TC_PROF_BUILD ? clock64() : 0, n_pass_done = 0;
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t smem_u = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
            const uint32_t res_lo = umma_desc_lo(smem_u), ring_lo = umma_desc_lo(smem_u + ring_off);
            const uint32_t stage_lo = (uint32_t)stage_bytes >> 4;
            const bool skip_mma = (a.dbg & 2) != 0;
            const int grp = nst >= 6 ? 2 : 1;          // a shallow ring (3x modes: 3 stages) cannot wait for two stages at once
            for (long long it = pair_id; it < n_items; it += n_pairs) {
                long long qt, rt0, rt1, rot; int seg;
                item_range(it, qt, rt0, rt1, seg, rot);
                for (long long ti = 0; ti < rt1 - rt0; ++ti) {
                    if (!first_pass) {  // both epilogues must have read the previous accumulators
                        const long long t0 = (MDSCTK_TC_PROF_BUILD && a.prof) ? clock64() : 0;
                        if (a.dbg & 64) mbar_wait(&bar_tmem_empty, tph, 2); else mbar_wait_spin(&bar_tmem_empty, tph, 2);
                        tph ^= 1;
                        if (MDSCTK_TC_PROF_BUILD && a.prof) t_wait_empty += clock64() - t0;
                    }
                    ++n_pass_done;
                    first_pass = false;
                    tc_fence_after();
                    // Stages are taken in groups of GRP: one wait/fence/elect round per group.  tcgen05.mma issue
                    // back-pressures (measured: the issuing thread's busy time equals the MMA execution time), so
                    // whatever runs between two stages leaves the tensor pipe idle -- with one MMA per k-step a stage
                    // is only 6 MMAs (430 clk) and a per-stage round cost a third of the MMA phase.
                    for (int kc0 = 0; kc0 < nk; kc0 += grp) {
                        const int ng = min(grp, nk - kc0);
                        {
                            const long long t0 = (MDSCTK_TC_PROF_BUILD && a.prof) ? clock64() : 0;
                            int s2 = s; uint32_t ph2 = ph;
                            for (int u = 0; u < ng; ++u) {
                                mbar_wait_spin(&bar_full[s2], ph2, 3);
                                if (++s2 == nst) { s2 = 0; ph2 ^= 1; }
                            }
                            if (MDSCTK_TC_PROF_BUILD && a.prof) t_wait_full += clock64() - t0;
                        }
                        if (res && ti == 0 && kc0 == 0) { mbar_wait_spin(&bar_res_full, rfph, 6); rfph ^= 1; }   // this item's y|z planes
                        tc_fence_after();
                        if (elect_one()) {
                            // descriptor LOW words: ring slot / resident chunk base + constant offsets (one add per operand)
                            int s2 = s;
                            for (int u = 0; u < ng; ++u) {
                                const int kc = kc0 + u;
                                const uint32_t st_lo = ring_lo + (uint32_t)s2 * stage_lo;
                                const uint32_t rs_lo = res_lo + (uint32_t)kc * (RES_CHUNK >> 4);
                                // atoms beyond A_pad are zero-filled by TMA; skip a k-step that is all padding
                                const int ksteps = skip_mma ? 0 : ((a.A_pad - kc * KC) * (BF16 ? 2 : 4) > 32 ? KSTEPS : 1);
#pragma unroll
                                for (int ks = 0; ks < KSTEPS; ++ks) {
                                    if (ks < ksteps) {
                                        const uint32_t koff = ks * 2;                      // 32 bytes >> 4
                                        const uint32_t bhi = st_lo + (uint32_t)(off_b >> 4) + koff;
                                        const uint32_t blo = st_lo + (OFF_BLO >> 4) + koff;
#pragma unroll
                                        for (int p = 0; p < 3; ++p) {
                                            const uint32_t d = tmem_u + p * UMMA_N;
                                            const uint32_t ahi = (res && p > 0) ? rs_lo + (p - 1) * (A_TILE >> 4) + koff
                                                                                : st_lo + (OFF_AHI >> 4) + p * (A_TILE >> 4) + koff;
                                            tc_mma2_lo<BF16>(d, ahi, bhi, IDESC, ks == 0 ? (uint32_t)(kc != 0) : 1u);
                                            if constexpr (B_LO) tc_mma2_lo<BF16>(d, ahi, blo, IDESC, 1);
                                            if constexpr (A_LO) {
                                                const uint32_t alo = st_lo + (OFF_ALO >> 4) + p * (A_TILE >> 4) + koff;
                                                tc_mma2_lo<BF16>(d, alo, bhi, IDESC, 1);
                                            }
                                        }
                                    }
                                }
                                tc_commit2_mc(&bar_empty[s2], 3);                      // frees the stage in both CTAs
                                if (kc == nk - 1) {
                                    tc_commit2_mc(&bar_tmem_full, 3);                  // accumulators complete -> both epilogues
                                    if (MDSCTK_TC_PROF_BUILD && a.prof) *reinterpret_cast<volatile long long *>(&s_tcommit) = clock64();
                                }
                                if (res && kc == nk - 1 && ti == rt1 - rt0 - 1) tc_commit2_mc(&bar_res_empty, 3);   // item done with its y|z planes
                                if (++s2 == nst) s2 = 0;
                            }
                        }
                        __syncwarp();
                        s += ng;                                  // ng <= nst
                        if (s >= nst) { s -= nst; ph ^= 1; }
                    }
                }
            }
            if (MDSCTK_TC_PROF_BUILD && a.prof && lane == 0) {
                long long *pr = a.prof + (size_t)blockIdx.x * 8;
                pr[0] = clock64() - t_total0; pr[1] = t_wait_empty; pr[2] = t_wait_full; pr[3] = n_pass_done;
            }
        } else if (a.dbg & 131072) {
            // =============================== hand-back forwarder (peer CTA, experiment) =
            uint32_t fph = 0;
            const uint32_t empty_leader = map_to_cta(&bar_tmem_empty, 0);
            for (long long it = pair_id; it < n_items; it += n_pairs) {
                long long qt, rt0, rt1, rot; int seg;
                item_range(it, qt, rt0, rt1, seg, rot);
                for (long long ti = 0; ti < rt1 - rt0; ++ti) {
                    if (lane == 0) {
                        mbar_wait_spin(&bar_epi_done, fph, 7);       // this CTA's 16 epilogue warps have read their accumulators
                        mbar_arrive_cluster_nofence(empty_leader);
                    }
                    fph ^= 1;
                    __syncwarp();
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
        const int sub = ew >> 2;                      // which 12 reference frames of every tile
        const int e_of_quarter = (quarter + 2) & 3;   // ew = sub * 4 + e_of_quarter
        const int row_in_tile = quarter * 32 + lane;
        // accumulator column of (plane b, ref 12 sub + j): refs 0-23 come from CTA 0's operand rows
        // (columns [0,72) = x|y|z x 24), refs 24-47 from CTA 1's (columns [72,144))
        const uint32_t t_warp = tmem_base + ((uint32_t)(quarter * 32) << 16) + (sub >> 1) * (3 * TRH) + (sub & 1) * SUBW;
        const uint32_t empty_leader = map_to_cta(&bar_tmem_empty, 0);
        const unsigned lt_mask = (1u << lane) - 1u;
        unsigned *hist = s_hist[ew];                  // radix histogram during merges, refine queue otherwise
        float *q_c2 = reinterpret_cast<float *>(hist), *q_c1 = q_c2 + 32, *q_c0 = q_c2 + 64, *q_e0 = q_c2 + 96,
              *q_x = q_c2 + 128;
        int *q_ref = reinterpret_cast<int *>(hist) + 160, *q_own = reinterpret_cast<int *>(hist) + 192;
        int *wcnt = s_cnt[ew];                        // fill of this warp's append area of row (quarter*32 + l)
        const size_t row_stride = (size_t)a.cl.H * a.cl.cap;
        uint32_t tph = 0;
        long long t_wait_tmem = 0, t_hold = 0, t_post = 0, t_merge = 0, n_live = 0, t_commit_lat = 0;
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
                    d2 = qcp_refine(c, e0, e0, x1, SC * tau_r);
                }
                d2 *= INV_SC;                         // accumulator units -> nm^2 (exact power of two)
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
        auto merge_rows = [&](long long row0, int seg, bool final) {
            quarter_sync(quarter);
            for (int r8 = 0; r8 < 8; ++r8) {
                const int l = sub * 8 + r8;
                const int row = quarter * 32 + l;
                const long long qr = row0 + row;
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

        for (long long it = pair_id; it < n_items; it += n_pairs) {
            long long qt, rt0, rt1, rot; int seg;
            item_range(it, qt, rt0, rt1, seg, rot);
            const long long row0 = qt * UMMA_M + rank * TQ;      // first fit row of this CTA within the query range
            const long long qrow = row0 + row_in_tile;
            const bool qvalid = qrow < a.n_q;
            const float hgq = (0.5f * SC) * a.q_G[a.q_begin + (qvalid ? qrow : a.n_q - 1)];   // accumulator units
            lbase0 = ((size_t)(row0 + quarter * 32) * a.cl.H + seg) * a.cl.cap;
            // pre-bound (nm units): RMSD >= |sigma(x) - sigma(y)|_2 - g, g = rounding residuals of what the sweep contracts
            // The warp's 32 fit rows are consecutive frames: their singular values span a small box, and the largest
            // rounding allowance / G of the 32 serve all of them, so ONE lane can test a reference frame for the warp.
            const float4 sq = __ldg(a.q_sig + a.q_begin + (qvalid ? qrow : a.n_q - 1));
            const float gq_nm = 2.0f * INV_SC * hgq;
            auto wmax = [](float v) { return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(v))); };   // v >= 0
            auto wmin = [](float v) { return __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(v))); };
            const float bx0 = wmin(sq.x), bx1 = wmax(sq.x), by0 = wmin(sq.y), by1 = wmax(sq.y), bz0 = wmin(sq.z), bz1 = wmax(sq.z);
            const float gq_max = wmax(gq_nm);
            const float pre_g = a.pre_rel * (sqrtf(gq_max) + a.pre_sqrt_gmax) + 1e-6f;
            wcnt[lane] = 0;
            if (sub == 0) {
                s_mcnt[row_in_tile] = 0;
                // admission threshold carried over from segments of this row that already finished
                s_tau[row_in_tile] = qvalid ? __ldcg(a.row_tau + qrow) : 0.0f;
            }
            quarter_sync(quarter);
            // per-reference scalars of a pass (G and the singular values of the warp's 12 frames): lane l < 12 fetches
            // those of frame l one pass AHEAD, the warp broadcasts them by shuffle -- twelve dependent-latency
            // broadcast loads per pass and lane cost more than the whole TMEM drain when registers are short
            float4 nx_sig = make_float4(0.f, 0.f, 0.f, 0.f);
            float nx_g = 0.f;
            auto fetch_refs = [&](long long ti2) {
                const long long rb2 = tile_at(ti2, rt0, rt1, rot) * TR + sub * SUBW;
                if (lane < SUBW) { nx_sig = __ldg(a.r_sig + rb2 + lane); nx_g = __ldg(a.r_G + rb2 + lane); }
            };
            fetch_refs(0);
            for (long long ti = 0; ti < rt1 - rt0; ++ti) {
                const long long rt = tile_at(ti, rt0, rt1, rot);
                const long long rb = rt * TR + sub * SUBW;       // first reference frame of this warp's columns
                const float4 cs = nx_sig;
                const float cg = nx_g;
                if (ti + 1 < rt1 - rt0) fetch_refs(ti + 1);
                // Cheapest test first, while the MMAs of this pass still run: min-RMSD^2 >= sum_i (sigma_i(x) - sigma_i(y))^2
                // (von Neumann).  Lane l < 12 tests reference frame l against the box of the warp's rows and the
                // largest threshold among them; a batch whose frames all fail is not even read from TMEM -- for
                // frames of another conformational basin that is nearly every batch, and the hold shrinks to the
                // hand-back.
                unsigned live = (1u << (SUBW / EB)) - 1u;          // batches whose accumulators must be read
                if (a.pre_rel >= 0.0f && a.do_fit) {
                    const float tau_max = wmax(*reinterpret_cast<volatile float *>(&s_tau[row_in_tile]));   // only decreases: stale is safe
                    const float d0 = fmaxf(fmaxf(bx0 - cs.x, cs.x - bx1), 0.0f), d1 = fmaxf(fmaxf(by0 - cs.y, cs.y - by1), 0.0f),
                                d2 = fmaxf(fmaxf(bz0 - cs.z, cs.z - bz1), 0.0f);
                    const float lb = sqrtf(fmaf(d0, d0, fmaf(d1, d1, d2 * d2))) - pre_g;
                    // key = rounded-structure RMSD^2 + accumulation bias/noise (< 2e-5 E0): keep clear of it
                    const bool far = lane >= SUBW || (lb > 0.0f && lb * lb > tau_max + 2e-5f * (gq_max + cg));
                    const unsigned near = ~__ballot_sync(0xffffffffu, far);      // bit l: frame l may hold a neighbour
                    live = 0;
#pragma unroll
                    for (int hb = 0; hb < SUBW / EB; ++hb)
                        if (near & (((1u << EB) - 1u) << (hb * EB))) live |= 1u << hb;
                    if (a.dbg & 1024) live = (1u << (SUBW / EB)) - 1u;
                }
                if (MDSCTK_TC_PROF_BUILD && a.prof && !(a.dbg & 65536)) n_live += __popc(live);
                const long long tp0 = (MDSCTK_TC_PROF_BUILD && a.prof) ? clock64() : 0;
                if (lane == 0) { if (a.dbg & 128) mbar_wait_spin(&bar_tmem_full, tph, 4); else mbar_wait(&bar_tmem_full, tph, 4); }
                tph ^= 1;
                __syncwarp();
                tc_fence_after();
                const long long tp1 = (MDSCTK_TC_PROF_BUILD && a.prof) ? clock64() : 0;
                if (MDSCTK_TC_PROF_BUILD && a.prof && rank == 0) t_commit_lat += tp1 - *reinterpret_cast<volatile long long *>(&s_tcommit);
                long long tp2 = tp1;
                const float tau = SC * *reinterpret_cast<volatile float *>(&s_tau[row_in_tile]);   // accumulator units
                const float htau = 0.5f * tau;
                // Software-pipelined drain: the tcgen05.ld of batch hb+1 are issued before the arithmetic of batch
                // hb, so the TMEM read port (the floor of the hold: 128 lanes x 432 columns per pass) stays busy.
                float svb[2][9][EB];
                auto load_batch = [&](int h, float (&dst)[9][EB]) {
#pragma unroll
                    for (int p = 0; p < 3; ++p)
#pragma unroll
                        for (int b = 0; b < 3; ++b)
                            tc_ld2(t_warp + p * UMMA_N + b * TRH + h, dst, p * 3 + b);
                };
                bool released = false;
                auto release = [&]() {                // hand the accumulators back to the MMA issuer
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (a.dbg & 512) mbar_arrive_cluster(empty_leader);
                        else if (a.dbg & 16384) mbar_arrive_cluster_relaxed(empty_leader);
                        else if (rank == 0 || !(a.dbg & 131072)) mbar_arrive_cluster_nofence(empty_leader);
                        else mbar_arrive(&bar_epi_done);              // peer CTA: local arrival, forwarded once by warp 1
                    }
                    if (MDSCTK_TC_PROF_BUILD && a.prof) tp2 = clock64();
                    released = true;
                };
                if (a.debug_tile) live = (1u << (SUBW / EB)) - 1u;
                if (live == 0) release();
                if (live & 1u) load_batch(0, svb[0]);
#pragma unroll
                for (int hb = 0; hb < SUBW / EB; ++hb) {
                    const int h = hb * EB;
                    float (&sv)[9][EB] = svb[hb & 1];
                    const bool cur = (live >> hb) & 1u;
                    if (cur) tc_wait_ld();
                    if (hb + 1 < SUBW / EB && ((live >> (hb + 1)) & 1u)) load_batch(h + EB, svb[(hb + 1) & 1]);
                    if (!released && (live >> (hb + 1)) == 0) release();     // last TMEM read of this pass is complete
                    if (!cur) continue;
                    if (a.debug_tile && it == 0 && rt == 0 && rank == 0) {
#pragma unroll
                        for (int c = 0; c < 9; ++c)
#pragma unroll
                            for (int j = 0; j < EB; ++j)
                                a.debug_tile[(size_t)row_in_tile * (9 * TR) + c * TR + sub * SUBW + h + j] = sv[c][j] * INV_SC;
                    }
                    float e0[EB];
#pragma unroll
                    for (int j = 0; j < EB; ++j) e0[j] = fmaf(0.5f * SC, __shfl_sync(0xffffffffu, cg, h + j), hgq);
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
                            push(qvalid && rb + h + j < a.n_r && 2.0f * (e0[j] - tr) < tau, 0.f, 0.f, 0.f, e0[j], tr,
                                 (int)(rb + h + j), lane | 32);
                        }
                        continue;
                    }
                    // (1) cheap bound: lambda_max <= sqrt(3) |S|_F, so RMSD^2 >= 2 (E0 - sqrt(3 F)).  Far
                    //     pairs (other conformational basins) are rejected for 9 FMAs; a batch whose
                    //     64 pairs are all far skips the characteristic polynomial altogether.
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
                        push(qvalid && rb + h + j < a.n_r && !(2.0f * (e0[j] - x1[j]) > tau), c[j].c2, c[j].c1, c[j].c0, e0[j],
                             x1[j], (int)(rb + h + j), lane);
                }
                drain();
                const long long tp3 = (MDSCTK_TC_PROF_BUILD && a.prof) ? clock64() : 0;
                // an append area takes at most SUBW entries per pass
                if (((ti & (MERGE_EVERY - 1)) == MERGE_EVERY - 1) && ti + 1 < rt1 - rt0) merge_rows(row0, seg, false);
                if (MDSCTK_TC_PROF_BUILD && a.prof) {
                    // dbg bit 65536: count only the passes that read accumulators (live != 0), and how many there were
                    const bool cnt = !(a.dbg & 65536) || live != 0;
                    if (cnt) { t_wait_tmem += tp1 - tp0; t_hold += tp2 - tp1; t_post += tp3 - tp2; t_merge += clock64() - tp3; }
                    if ((a.dbg & 65536) && live != 0) n_live += 1;
                }
            }
            merge_rows(row0, seg, true);              // leaves one list of <= keep candidates per row
        }
        if (MDSCTK_TC_PROF_BUILD && a.prof && ew == 0 && lane == 0) {
            long long *pr = a.prof + (size_t)blockIdx.x * 8 + 4;
            pr[0] = (a.dbg & 65536) ? n_live : t_wait_tmem; pr[1] = t_hold; pr[2] = (a.dbg & 8192) ? t_commit_lat : t_post; pr[3] = (a.dbg & 8192) ? n_live : t_merge;   // dbg 8192: live batches instead of merge clocks
        }
    }

    // ---- teardown --------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                  // no CTA leaves (or frees TMEM) while the peer may still use it
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

// Timing-experiment switches (MDSCTK_TC_DEBUG bits, TcArgs::dbg / Tc2Args::dbg).  Several of them give INVALID results, so the
// production library never reads the environment: they exist only in builds with -DMDSCTK_TC_EXPERIMENTS=1
// (scripts/build_variant.sh exp -DMDSCTK_TC_EXPERIMENTS=1; scripts/build_prof.sh adds the clock counters).  Bits the
// version-2 sweep (rms_tc2.cu) honours: 1 skip the QCP bounds (invalid), 2 skip the MMAs (invalid), 8 no early-out of far
// batches, 64 every pass light (invalid), 512 short mailbox polls, 1024 no scout (every pass live), 2048 no pre-bound,
// 8192 generic per-MMA issue path, 32768 no own-tile guess; MDSCTK_TC_N overrides the MMA's N (invalid), MDSCTK_TC_WIDE=0
// forces 32-atom ring stages.
int tc_experiment_bits()
{
#if MDSCTK_TC_EXPERIMENTS
    const char *dbg = getenv("MDSCTK_TC_DEBUG");
    return dbg ? atoi(dbg) : 0;
#else
    return 0;
#endif
}

// Out-of-sample queries (knn_rms -f): the fit rows are not reference frames, so "start at the row's own tile" has
// no meaning -- but the frames that can be near a fit frame have nearly its singular values (the von Neumann bound
// again), so every fit super-tile starts at the reference tile whose mid frame is closest to the super-tile's mean
// in singular-value space.  Trajectory frames of one conformational basin are contiguous, so the admission
// thresholds are tight after a few passes, as in the in-sample order.  One block per super-tile.
__global__ void __launch_bounds__(256) rms_guess_own_tile_kernel(const float4 *q_sig, long long q_begin, long long n_q,
                                                                 const float4 *r_sig, long long n_r, int *own_tile)
{
    __shared__ float s_sum[3][8];
    __shared__ float s_best[8];
    __shared__ int s_arg[8], s_cnt[8];
    const long long qt = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long row = qt * tc::UMMA_M + tid;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    int have = 0;
    if (row < n_q) { v = q_sig[q_begin + row]; have = 1; }
    float sx = v.x, sy = v.y, sz = v.z;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sz += __shfl_xor_sync(0xffffffffu, sz, o); have += __shfl_xor_sync(0xffffffffu, have, o);
    }
    if (lane == 0) { s_sum[0][warp] = sx; s_sum[1][warp] = sy; s_sum[2][warp] = sz; s_cnt[warp] = have; }
    __syncthreads();
    float mx = 0.f, my = 0.f, mz = 0.f;
    int cnt = 0;
    for (int w = 0; w < 8; ++w) { mx += s_sum[0][w]; my += s_sum[1][w]; mz += s_sum[2][w]; cnt += s_cnt[w]; }
    const float inv = 1.0f / (float)max(cnt, 1);
    mx *= inv; my *= inv; mz *= inv;
    const long long n_rt = (n_r + tc::TR - 1) / tc::TR;
    float best = __uint_as_float(0x7f800000u);
    int arg = 0;
    for (long long t = tid; t < n_rt; t += 256) {
        const float4 r = r_sig[min(t * tc::TR + tc::TR / 2, n_r - 1)];
        const float d0 = r.x - mx, d1 = r.y - my, d2 = r.z - mz;
        const float d = fmaf(d0, d0, fmaf(d1, d1, d2 * d2));
        if (d < best) { best = d; arg = (int)t; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ob < best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    if (lane == 0) { s_best[warp] = best; s_arg[warp] = arg; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; ++w)
            if (s_best[w] < best || (s_best[w] == best && s_arg[w] < arg)) { best = s_best[w]; arg = s_arg[w]; }
        own_tile[qt] = arg;
    }
}

cudaError_t launch_rms_guess_own_tile(const float4 *q_sig, long long q_begin, long long n_q, const float4 *r_sig, long long n_r,
                                      int *own_tile, cudaStream_t st)
{
    const long long n_qt = (n_q + tc::UMMA_M - 1) / tc::UMMA_M;
    if (n_qt <= 0) return cudaSuccess;
    rms_guess_own_tile_kernel<<<(unsigned)n_qt, 256, 0, st>>>(q_sig, q_begin, n_q, r_sig, n_r, own_tile);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- host side ----
// planes[n][3][A_pad] as a 3-D tensor ordered (atom, frame, plane); box = 64 bytes of atoms x `rows`
// frames x 3 planes, so one TMA op lands the three plane tiles back to back as [plane][frame][atoms].
static bool make_plane_map(CUtensorMap *m, const void *planes, long long n, int A_pad, int rows, int mode, int box_planes = 3)
{
    EncodeTiledFn enc = get_tensor_map_encoder();
    if (!enc) return false;
    const int esz = mode >= 3 ? 2 : 4;
    const CUtensorMapDataType dt = mode == 3 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                   : (mode >= 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
    cuuint64_t dims[3] = {(cuuint64_t)A_pad, (cuuint64_t)n, 3};
    cuuint64_t strides[2] = {(cuuint64_t)A_pad * 3 * esz, (cuuint64_t)A_pad * esz};
    cuuint32_t box[3] = {(cuuint32_t)(tc::ROW_BYTES / esz), (cuuint32_t)rows, (cuuint32_t)box_planes};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, dt, 3, const_cast<void *>(planes),
               dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Reference segments per fit super-tile.  A work item is (super-tile, segment), dealt statically to the CTA pairs, so two
// things decide: (1) whole waves of items over the pairs, and (2) -- measured in round 2 -- the length of an item: with one
// segment an item sweeps ALL reference frames (110 ms at 1M frames) and a launch ends with the pairs finishing up to one
// item apart (C4 block, 131 072 rows x 1M frames: 864 ms with 1 segment, 727 ms with 4, 704 ms with 12; C3: 72.8 / 63 / 60 ms
// with 1 / 3 / 4; clk per pass and DRAM traffic are the same -- it is the tail of the launch that shrinks).  More segments
// cost candidate lists (one per row and segment) and re-score merging (C4: 22 ms with 1, 29 ms with 4, 50 ms with 8), so 4
// is the sweet spot; every segment keeps at least 16 reference tiles.
int rms_tc_choose_segments(long long n_fit, long long n_ref, int n_sms)
{
    const int n_pairs = n_sms / 2 > 0 ? n_sms / 2 : 1;
    const long long n_qt = (n_fit + tc::UMMA_M - 1) / tc::UMMA_M;
    const long long n_rt = (n_ref + tc::TR - 1) / tc::TR;
    static const double balance[9] = {0.0, 0.84, 0.93, 0.97, 1.0, 0.97, 0.95, 0.94, 0.93};   // measured sweep + re-score time relative to 4 segments
    int best = 1;
    double best_score = 0.0;
    for (int s = 1; s <= 8; ++s) {
        if (s > 1 && n_rt / s < 16) break;
        const long long items = n_qt * s;
        const double eff = (double)items / (double)(((items + n_pairs - 1) / n_pairs) * n_pairs);
        const double score = eff * balance[s];
        if (score > best_score + 1e-9) { best_score = score; best = s; }
    }
    return best;
}

int rms_tc_lists_per_segment() { return 1; }

// Entries reserved per list: keep merged + one append area per epilogue warp of a lane quarter.
int rms_tc_list_stride(int keep) { return (keep + tc::SUBS * tc::SUB_APP + 31) / 32 * 32; }

// resident mode (1xFP16): ring stages that fit beside the y|z planes of the fit tile; 0 = does not fit
static int res_ring_stages(int A_pad)
{
    const int nk = (A_pad + 31) / 32;
    if (nk > tc::RES_MAX_NK) return 0;
    const int room = 227 * 1024 - tc::STATIC_SMEM - 1024 - nk * tc::RES_CHUNK;
    const int nst = room / tc::RES_STAGE;
    return nst < 2 ? 0 : (nst < tc::MAX_NST ? nst : tc::MAX_NST);
}

template <int M>
static cudaError_t launch_tc_mode(const CUtensorMap &mq_hi, const CUtensorMap &mq_lo, const CUtensorMap &mr_hi,
                                  const CUtensorMap &mr_lo, const TcArgs &a, long long n_items, int n_sms, cudaStream_t st)
{
    const int smem_bytes = a.res ? ((a.A_pad + 31) / 32) * tc::RES_CHUNK + a.res_nst * tc::RES_STAGE + 1024 : tc::Mode<M>::SMEM_BYTES;
    cudaError_t e = cudaFuncSetAttribute(rms_sweep_tc_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, rms_sweep_tc_kernel<M>);
    if (e == cudaSuccess && fa.sharedSizeBytes > (size_t)tc::STATIC_SMEM) e = cudaErrorInvalidConfiguration;
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.blockDim = dim3(tc::NTHR); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
    int max_pairs = n_sms / 2;
    cfg.gridDim = dim3((unsigned)(max_pairs * 2));
    int q = 0;
    if (cudaOccupancyMaxActiveClusters(&q, rms_sweep_tc_kernel<M>, &cfg) == cudaSuccess && q > 0 && q < max_pairs)
        max_pairs = q;                              // a GPC with an odd SM count cannot host every pair
    (void)cudaGetLastError();
#if MDSCTK_TC_EXPERIMENTS
    if (const char *mp = getenv("MDSCTK_TC_MAX_PAIRS")) { const int v = atoi(mp); if (v > 0 && v < max_pairs) max_pairs = v; }
#endif
    const long long n_pairs = n_items < max_pairs ? n_items : max_pairs;
    cfg.gridDim = dim3((unsigned)(n_pairs * 2));
    return cudaLaunchKernelEx(&cfg, rms_sweep_tc_kernel<M>, mq_hi, mq_lo, mr_hi, mr_lo, a);
}

cudaError_t launch_rms_sweep_tc(int mode, const FrameSetView &fit, const void *fit_hi, const void *fit_lo,
                                long long fit_begin, long long n_fit, const FrameSetView &ref, const void *ref_hi,
                                const void *ref_lo, int do_fit, int n_seg, CandLists<float> cl, float *row_tau,
                                float g_ref_max, int *own_tile_scratch, float *debug_tile, int n_sms, cudaStream_t st)
{
    if (n_fit <= 0) return cudaSuccess;
    if (cl.H != n_seg || cl.cap < cl.keep + tc::SUBS * tc::SUB_APP) return cudaErrorInvalidValue;
    const int dbg_bits = tc_experiment_bits();     // 0 unless built with -DMDSCTK_TC_EXPERIMENTS=1 (scripts/build_prof.sh)
    // 1xFP16, MDSCTK_TC_DEBUG bit 256: keep two of the three fit planes resident in shared memory when they fit.
    // Off by default: it cuts the L2->SM operand stream 2.2x but leaves room for only 3 ring stages at 300 atoms,
    // and the sweep is bound by the MMA <-> epilogue hand-over, not by the stream (measured 92-101 ms vs 84-89 ms).
    const int res_nst = (mode == 6 && (dbg_bits & 256)) ? res_ring_stages(ref.A_pad) : 0;
    CUtensorMap mq_hi, mq_lo, mr_hi, mr_lo;
    if (!make_plane_map(&mq_hi, fit_hi, fit.n, fit.A_pad, tc::TQ, mode) ||
        // resident mode has no lo operand: the slot carries the single-plane box of the fit hi planes
        !(res_nst ? make_plane_map(&mq_lo, fit_hi, fit.n, fit.A_pad, tc::TQ, mode, 1)
                  : make_plane_map(&mq_lo, fit_lo, fit.n, fit.A_pad, tc::TQ, mode)) ||
        !make_plane_map(&mr_hi, ref_hi, ref.n, ref.A_pad, tc::TRH, mode) ||
        !make_plane_map(&mr_lo, ref_lo, ref.n, ref.A_pad, tc::TRH, mode))
        return cudaErrorInvalidValue;
    TcArgs a;
    // the reduced FP16 modes measure distances between the rounded structures: E0 from their norms
    a.q_G = mode >= 5 ? fit.Gh : fit.G;
    a.r_G = mode == 6 ? ref.Gh : (mode == 5 ? ref.G2 : ref.G);
    a.q_begin = fit_begin; a.n_q = n_fit; a.n_r = ref.n;
    a.A_pad = ref.A_pad; a.do_fit = do_fit; a.n_seg = n_seg; a.cl = cl; a.debug_tile = debug_tile; a.row_tau = row_tau;
    a.dbg = dbg_bits;
    a.res = res_nst > 0; a.res_nst = res_nst;
    // pre-bound: relative rounding error of one operand of this mode (what ties the contracted structures to the
    // true ones): tf32/fp16 hi only 2^-11, bf16 hi+mid 2^-16, two-part splits 2^-20; MDSCTK_TC_DEBUG bit 2048: off
    static const float kPreRel[7] = {0.f, 1.0e-6f, 4.9e-4f, 1.6e-5f, 1.0e-6f, 4.9e-4f, 4.9e-4f};
    a.q_sig = reinterpret_cast<const float4 *>(fit.sig); a.r_sig = reinterpret_cast<const float4 *>(ref.sig);
    a.pre_rel = (dbg_bits & 2048) ? -1.0f : kPreRel[mode];
    a.pre_sqrt_gmax = sqrtf(g_ref_max > 0.f ? g_ref_max : 0.f);
    a.own_tile = nullptr;
    if (own_tile_scratch && !(dbg_bits & 32768)) {     // out-of-sample: where to start each fit super-tile (bit 32768: off)
        const long long n_qt = (n_fit + tc::UMMA_M - 1) / tc::UMMA_M;
        rms_guess_own_tile_kernel<<<(unsigned)n_qt, 256, 0, st>>>(a.q_sig, fit_begin, n_fit, a.r_sig, ref.n, own_tile_scratch);
        a.own_tile = own_tile_scratch;
    }
    // MDSCTK_TC_PROF=1: per-CTA clock sums {MMA warp: total, wait tmem_empty, wait full, passes |
    // epilogue warp 0: wait tmem_full, TMEM hold, post-release compute, merges}, printed to stderr
    static long long *d_prof = nullptr;
    bool prof = false;
#if MDSCTK_TC_PROF_BUILD
    if (const char *pe = getenv("MDSCTK_TC_PROF")) prof = atoi(pe) != 0;
#endif
    a.prof = nullptr;
    if (prof) {
        if (!d_prof && cudaMalloc(&d_prof, 1024 * 8 * sizeof(long long)) != cudaSuccess) return cudaErrorMemoryAllocation;
        cudaMemsetAsync(d_prof, 0, 1024 * 8 * sizeof(long long), st);
        a.prof = d_prof;
    }
    const long long n_items = ((n_fit + tc::UMMA_M - 1) / tc::UMMA_M) * n_seg;
    cudaError_t e;
    switch (mode) {
    case 1: e = launch_tc_mode<1>(mq_hi, mq_lo, mr_hi, mr_lo, a, n_items, n_sms, st); break;
    case 2: e = launch_tc_mode<2>(mq_hi, mq_lo, mr_hi, mr_lo, a, n_items, n_sms, st); break;
    case 3: e = launch_tc_mode<3>(mq_hi, mq_lo, mr_hi, mr_lo, a, n_items, n_sms, st); break;
    case 4: e = launch_tc_mode<4>(mq_hi, mq_lo, mr_hi, mr_lo, a, n_items, n_sms, st); break;
    case 5: e = launch_tc_mode<5>(mq_hi, mq_lo, mr_hi, mr_lo, a, n_items, n_sms, st); break;
    case 6: e = launch_tc_mode<6>(mq_hi, mq_lo, mr_hi, mr_lo, a, n_items, n_sms, st); break;
    default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    if (prof) {
        static long long h[1024 * 8];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost);
        double sum[8] = {0}; int nb = 0;
        for (int b = 0; b < 1024; b += 1) {
            if (h[b * 8 + 4] == 0 && h[b * 8] == 0) continue;
            nb++;
            for (int k = 0; k < 8; ++k) sum[k] += (double)h[b * 8 + k];
        }
        // leaders carry the MMA columns (half of the CTAs); epilogue columns come from every CTA
        fprintf(stderr, "[tc prof] ctas=%d  mma: total %.0f  wait_tmem_empty %.0f  wait_full %.0f  passes %.0f (clk, per leader) | "
                        "epi warp0: wait_tmem_full %.0f  hold %.0f  post %.0f  merge %.0f (clk per CTA)\n",
                nb, sum[0] / (nb / 2.0), sum[1] / (nb / 2.0), sum[2] / (nb / 2.0), sum[3] / (nb / 2.0), sum[4] / nb, sum[5] / nb,
                sum[6] / nb, sum[7] / nb);
    }
    return cudaGetLastError();
}

}  // namespace mdsctk
