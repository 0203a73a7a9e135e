// rms_tc2.cu -- second-generation tcgen05 sweep for the default 1xFP16 contraction (round 2).
//
// Same job, tile shapes, epilogue arithmetic and candidate lists as rms_tc.cu (the row blocks of
// knn_rms.cpp:231-293: fit super-tile of 256 frames x reference segment per CTA pair, pass = 256 x 48 pairs,
// M=256 N=144 cta_group::2 MMAs into 432 TMEM columns, QCP bounds + streaming top-k in 16 epilogue warps).
// What changed is everything AROUND the MMAs, because the round-1 counters showed a pass spending two thirds of
// its time outside them (12.4k clk per pass against 4.1k clk of tensor time):
//
//   resident fit tile   The fit operand of a work item is pass-invariant, yet round 1 re-streamed it from L2 every
//                       pass (240 of 291 KB per pass and CTA, 6.4 TB/s of L2->SM traffic -- the pass could not go
//                       below ~7k clk on bandwidth alone).  Here the CTA's 128 x 3 x A_pad fp16 fit tile is loaded
//                       ONCE per item: 64-byte-row chunks (two k-steps of one plane, 8 KB) in shared memory, and
//                       what does not fit (233 KB at 300 atoms against 227 KB of shared memory) in the 80 TMEM
//                       columns the accumulators leave free -- tcgen05.mma takes its A operand from tensor memory
//                       as readily as from shared memory (the .ts form), so those k-steps also cost no shared-memory
//                       read bandwidth.  The ring then carries only the reference half-tiles: 46 KB per pass.
//   pass director       In a pass whose 48 reference frames all fail the singular-value pre-bound against the
//                       super-tile (frames of another conformational basin: nearly every pass once the rows'
//                       thresholds are tight) nobody reads the accumulators, so nobody needs to hand them back.  The
//                       MMA warp tests that bound (a scout warp evaluates it ahead of the tensor pipe), and tells the
//                       epilogue warps of both CTAs only about the passes that can hold a neighbour ("heavy" passes,
//                       through a mailbox word in shared memory).  Light passes run back to back on the tensor
//                       pipe with no TMEM hand-over at all; the epilogue warps sleep through them.  The contraction
//                       itself stays dense: every MMA of every pass is issued.
//   few, fat rounds     What then paced a light pass was neither the MMAs nor the copies but the two single threads
//                       that drive them: a round of the producer (barrier wait, expect-tx, one bulk-tensor copy) and
//                       of the issuer (barrier test, MMAs, commit) costs each ~500 clk whatever it moves (measured:
//                       the pass took 6.1k clk with the MMAs skipped, 7.9k with them, whether the ring was 5, 6 or
//                       8 stages deep; one commit per two stages changed nothing; issuing from ONE thread in a
//                       divergent branch instead of an elected lane of the converged warp made every tcgen05.mma
//                       block for its 72 clk).  So the reference operand moves in 64-atom stages of 128-byte rows
//                       (SWIZZLE_128B against the fit tile's SWIZZLE_64B; 9 KB, twelve MMAs per round) wherever
//                       three such stages fit beside the fit tile -- at 300 atoms the refine queues give up 12 of
//                       their 32 entries for the third -- the producer is one thread, and the scout's arithmetic
//                       has a warp of its own: 304-atom light pass 7.9k -> 5.9k clk, C4 block 856 -> 704 ms.
//
// Warps: 0 copy producer (one thread, both CTAs), 1 MMA issuer + director (leader CTA), 2-17 epilogue, 18 pass scout (leader).
//
// Synchronisation (per CTA pair; "leader" = CTA rank 0, which issues the pair's MMAs):
//   bar_full[s] / bar_empty[s]   reference ring: TMA complete_tx on the leader / tcgen05.commit multicast to both CTAs
//   bar_res_full / bar_res_empty resident fit tile of an item loaded / every MMA of the item retired (both CTAs)
//   bar_item_ready (leader)      32 epilogue warps: TMEM-resident k-steps written, quarter thresholds published
//   mtile[4] / pass_now (leader) scout -> director: pass number + 1 | threshold that makes the tile heavy; director -> scout: pass
//                                being issued (the scout stays at most three passes ahead)
//   mailbox (both CTAs)          (item seq << 20) | (heavy pass index + 1), or | 0xFFFFF = end of item; written by the
//                                director only after every epilogue warp has handed back the previous heavy pass,
//                                so one word per CTA is enough
//   bar_tmem_full / bar_tmem_empty  accumulators of a HEAVY pass complete / read by all 32 epilogue warps
#include "common.cuh"
#include "qcp.cuh"
#include "select.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>

#ifndef MDSCTK_TC_PROF_BUILD
#define MDSCTK_TC_PROF_BUILD 0
#endif

namespace mdsctk {

namespace tc2 {
constexpr int TQ = 128;                       // fit frames per CTA
constexpr int TR = 48;                        // reference frames per pass (24 loaded by each CTA)
constexpr int TRH = TR / 2;
constexpr int UMMA_M = 2 * TQ;                // 256 across the pair
constexpr int UMMA_N = 3 * TR;                // 144
constexpr int A_CHUNK = TQ * 64;              // 8192 B: 32 atoms (two k-steps) of one plane of the CTA's fit tile
constexpr int B_STAGE = 3 * TRH * 64;         // 4608 B: 32 atoms of this CTA's 72 reference operand rows
constexpr int SUBS = 4;                       // epilogue warps per TMEM lane quarter
constexpr int EPI_WARPS = 4 * SUBS;           // 16
constexpr int SCOUT_WARP = 2 + EPI_WARPS;       // warp 18: the pass scout (leader CTA)
constexpr int NTHR = 64 + EPI_WARPS * 32 + 32; // 608
constexpr int SUBW = TR / SUBS;               // 12 reference columns per epilogue warp
constexpr int EB = 2;                         // pairs per lane and epilogue batch
constexpr int MERGE_EVERY = 4;                // heavy passes between quarter-wide list merges
constexpr int SUB_APP = 2 * MERGE_EVERY * SUBW;   // 96: private append area per epilogue warp and row
constexpr int ACC_COLS = 3 * UMMA_N;          // 432 accumulator columns
constexpr int TMEM_COLS = 512;
constexpr int MAX_TMEM_UNITS = (TMEM_COLS - ACC_COLS) / 8;   // 10 k-steps of one plane (8 columns of fp16 pairs each)
constexpr int MAX_NST = 8;
constexpr int SMEM_MAX = 227 * 1024;
constexpr uint32_t END_PASS = 0xFFFFFu;
#ifndef MDSCTK_TC2_QCAP
#define MDSCTK_TC2_QCAP 32
#endif
constexpr int QCAP = MDSCTK_TC2_QCAP;         // refine-queue entries per epilogue warp (20 B per entry and warp of the ring's shared memory);
constexpr int QCAP_SMALL = 20;                // ... and the size that buys a third wide ring stage at 300 atoms (>= 13: the histogram)
constexpr int B_STAGE_WIDE = 2 * B_STAGE;     // 9216 B: 64 atoms (128-byte rows, SWIZZLE_128B) per stage
constexpr float SC = kRmsHalfScale * kRmsHalfScale, INV_SC = 1.0f / SC;   // accumulators hold SC * S
constexpr uint32_t IDESC = umma_idesc(0, UMMA_M, UMMA_N);                  // fp16 x fp16 -> fp32

// control block, carved from dynamic shared memory behind the operands (no static shared memory: the operand area
// must start on a 1024-byte boundary and every byte counts)
struct Ctl {
    uint64_t bar_full[MAX_NST], bar_empty[MAX_NST];
    uint64_t bar_tmem_full, bar_tmem_empty, bar_res_full, bar_res_empty, bar_item_ready;
    uint32_t tmem_base;
    uint32_t mailbox;
    float qtau[8];                    // leader's copy is what the director reads: [cta rank][quarter] largest threshold
#if MDSCTK_TC_PROF_BUILD
    long long t_commit[MAX_NST], t_issue[MAX_NST];   // clock counters of the ring's round trip (leader CTA)
#endif
    unsigned long long mtile[4];      // scout -> director, per pass (mod 4): pass number + 1 | (smallest threshold that makes the tile heavy) << 32
    uint32_t pass_now;                // director -> scout: the pass being issued (the scout stays at most 3 ahead)
    float tau[TQ];                    // running admission threshold per fit row of this CTA
    unsigned short mcnt[TQ];          // merged entries per row
    unsigned cnt8[EPI_WARPS][8];      // fill of each warp's private append area, one BYTE per row of its quarter
    unsigned scratch[4];              // really [EPI_WARPS][5 * qcap]: per warp, refine queue (5 x qcap words) / 64-bin radix histogram during merges
};
}  // namespace tc2

struct Tc2Args {
    const float *q_G, *r_G;       // norms of the ROUNDED structures (Gh)
    long long q_begin, n_q, n_r;
    int A_pad, do_fit, n_seg;
    CandLists<float> cl;
    float *debug_tile;
    float *row_tau;
    const float4 *q_sig, *r_sig;
    float pre_rel, pre_sqrt_gmax;
    const int *own_tile;
    const __half *q_fh;           // fit fp16 planes [n][3][A_pad]: source of the TMEM-resident k-steps
    int nkc, nks;                 // reference chunks per pass (ceil(A_pad/32)), k-steps (A_pad/16)
    int chunks[3], chunk_base[3]; // plane p: `chunks` 64-byte chunks resident in shared memory from slot chunk_base
    int tmem_unit0[3];            // plane p: its k-steps [2 chunks[p], nks) live in TMEM units tmem_unit0[p]..
    int n_tmem_units, n_res_chunks, nst;
    int qcap;                     // refine-queue entries per epilogue warp
    int wide, bst;                // 128-byte reference rows: 64 atoms (four k-steps) per ring stage; bytes per stage
    int dbg;
    uint32_t idesc;               // instruction descriptor (M=256, N=144, fp16 -> fp32; other N only in timing experiments)
    long long *prof;
};

namespace tc2 {

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p)
{
    return *reinterpret_cast<const volatile uint32_t *>(p);
}
__device__ __forceinline__ float wmaxf(float v) { return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(v))); }   // v >= 0
__device__ __forceinline__ float wminf(float v) { return __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(v))); }

}  // namespace tc2

__global__ void __launch_bounds__(tc2::NTHR, 1)
rms_sweep_tc2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_r, Tc2Args a)
{
    using namespace tc2;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *ring = smem + (size_t)a.n_res_chunks * A_CHUNK;
    const int bst = a.bst;
    Ctl &c = *reinterpret_cast<Ctl *>(ring + (size_t)a.nst * bst);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int nkc = a.nkc, nst = a.nst;
    const int katoms = a.wide ? 64 : 32;          // atoms per ring stage
    const long long n_qt = (a.n_q + UMMA_M - 1) / UMMA_M;
    const long long n_rt = (a.n_r + TR - 1) / TR;
    const long long n_items = n_qt * a.n_seg;
    const long long pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) { printf("rms_sweep_tc2: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
        for (int s = 0; s < MAX_NST; ++s) { mbar_init(&c.bar_full[s], 1); mbar_init(&c.bar_empty[s], 1); }
        mbar_init(&c.bar_tmem_full, 1);
        mbar_init(&c.bar_tmem_empty, 2 * EPI_WARPS);
        mbar_init(&c.bar_res_full, 1);
        mbar_init(&c.bar_res_empty, 1);
        mbar_init(&c.bar_item_ready, 2 * EPI_WARPS);
        c.mailbox = 0;
        c.pass_now = 0;
        for (int i = 0; i < 4; ++i) c.mtile[i] = 0ull;
#if MDSCTK_TC_PROF_BUILD
        for (int i = 0; i < MAX_NST; ++i) { c.t_commit[i] = 0; c.t_issue[i] = 0; }
#endif
        for (int i = 0; i < 8; ++i) c.qtau[i] = __uint_as_float(0x7f800000u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&c.tmem_base)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = c.tmem_base;

    // item -> (fit super-tile, reference segment): diagonal first, then segment-major (see rms_tc.cu)
    auto item_range = [&](long long it, long long &qt, long long &rt0, long long &rt1, int &seg, long long &rot) {
        int s_rest = -1;
        if (it < n_qt) qt = it;
        else { s_rest = (int)((it - n_qt) / n_qt); qt = (it - n_qt) - (long long)s_rest * n_qt; }
        const long long own = min(a.own_tile ? (long long)a.own_tile[qt] : (a.q_begin + qt * UMMA_M) / TR, n_rt - 1);
        int sd = (int)(own * a.n_seg / n_rt);
        while (sd + 1 < a.n_seg && n_rt * (sd + 1) / a.n_seg <= own) ++sd;
        while (sd > 0 && n_rt * sd / a.n_seg > own) --sd;
        seg = s_rest < 0 ? sd : (s_rest < sd ? s_rest : s_rest + 1);
        rt0 = n_rt * seg / a.n_seg;
        rt1 = n_rt * (seg + 1) / a.n_seg;
        rot = s_rest < 0 ? own - rt0 : 0;
    };
    auto tile_at = [](long long i, long long rt0, long long rt1, long long rot) {
        const long long t = rt0 + rot + i;
        return t < rt1 ? t : t - (rt1 - rt0);
    };

    const bool use_pre = a.pre_rel >= 0.0f && a.do_fit && !a.debug_tile && !(a.dbg & 1024);
    if (warp == 0) {
        // =============================== TMA producer (both CTAs): ONE thread ====================
        // The ring's copies are issued by a single thread, as the MMAs are: a round of the loop is a barrier test, an
        // expect-tx and one bulk-tensor copy, and with the whole warp in it (32 lanes polling the barrier, an election and a
        // warp barrier per round) a round cost ~700 clk -- the producer, not the tensor pipe, paced the light passes
        // (measured: the pass took as long with the MMAs skipped).  The scout's arithmetic lives in a warp of its own.
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0, rph = 0;
            bool first_item = true;
            long long p_wake = 0, p_n = 0, p_wait = 0, p_issue = 0, p_rounds = 0, p_total0 = MDSCTK_TC_PROF_BUILD ? clock64() : 0;
            const bool p_on = MDSCTK_TC_PROF_BUILD && a.prof && rank == 0;
            const uint32_t res_leader = map_to_cta(&c.bar_res_full, 0), full_leader0 = map_to_cta(&c.bar_full[0], 0);
            for (long long it = pair_id; it < n_items; it += n_pairs) {
                long long qt, rt0, rt1, rot; int seg;
                item_range(it, qt, rt0, rt1, seg, rot);
                const int q0 = (int)(a.q_begin + qt * UMMA_M + rank * TQ);
                // the fit tile of this item, once: the previous item's MMAs must have retired
                if (!first_item) { mbar_wait(&c.bar_res_empty, rph, 5); rph ^= 1; }
                first_item = false;
                if (rank == 0) mbar_expect_tx(&c.bar_res_full, 2u * (uint32_t)(a.n_res_chunks * A_CHUNK));
                for (int p = 0; p < 3; ++p)
                    for (int ch = 0; ch < a.chunks[p]; ++ch)
                        tma_load_3d_2sm(smem + (size_t)(a.chunk_base[p] + ch) * A_CHUNK, &map_q, res_leader, ch * 32, q0, p, kEvictNormal);
                const long long n_tiles = rt1 - rt0;
                for (long long ti = 0; ti < n_tiles; ++ti) {
                    const int r0 = (int)(tile_at(ti, rt0, rt1, rot) * TR + rank * TRH);
                    for (int kc = 0; kc < nkc; ++kc) {
                        const long long pw0 = p_on ? clock64() : 0;
                        mbar_wait(&c.bar_empty[s], ph ^ 1, 1);       // suspending wait: the hardware sleeps the thread until the barrier flips
                        const long long pw1 = p_on ? clock64() : 0;
#if MDSCTK_TC_PROF_BUILD
                        if (p_on) {
                            const long long tc0 = *reinterpret_cast<volatile long long *>(&c.t_commit[s]);
                            if (tc0 > 0 && pw1 > tc0 && pw1 - tc0 < 1000000) { p_wake += pw1 - tc0; ++p_n; }
                            *reinterpret_cast<volatile long long *>(&c.t_issue[s]) = pw1;
                        }
#endif
                        if (rank == 0) mbar_expect_tx(&c.bar_full[s], 2u * (uint32_t)bst);
                        tma_load_3d_2sm(ring + (size_t)s * bst, &map_r, full_leader0 + 8u * (uint32_t)s, kc * katoms, r0, 0, kEvictNormal);
                        if (p_on) { p_wait += pw1 - pw0; p_issue += clock64() - pw1; ++p_rounds; }
                        if (++s == nst) { s = 0; ph ^= 1; }
                    }
                }
            }
#if MDSCTK_TC_PROF_BUILD
            if (p_on) {
                a.prof[(size_t)blockIdx.x * 8 + 5] = p_wake; a.prof[(size_t)blockIdx.x * 8 + 6] = p_n;
                long long *pr = a.prof + (size_t)(blockIdx.x + 1) * 8;      // the peer CTA's row is unused: producer counters
                pr[0] = clock64() - p_total0; pr[1] = 0; pr[2] = p_wait; pr[3] = p_issue; pr[4] = p_rounds;
            }
#endif
        }
        __syncwarp();
    } else if (warp == SCOUT_WARP) {
        // =============================== pass scout (leader CTA) ==================================
        // The director's arithmetic, done ahead of it: for every reference tile,
        //     m = min over its 48 frames of  lb^2 - 2e-5 (Gq_max + G_r)      (-inf where lb <= 0)
        // with lb the von Neumann lower bound of the frame against the box of the super-tile's singular values.  A pass
        // can hold a neighbour of one of the 256 rows only if m <= the largest admission threshold among them, which is
        // all the MMA warp has to test per pass.  One 8-byte word per pass, tagged with the pass number; the scout stays at
        // most three passes ahead of the director (four slots).
        if (rank == 0 && use_pre) {
            const float kNegInf = __uint_as_float(0xff800000u), kPosInf = __uint_as_float(0x7f800000u);
            uint32_t g = 0;                               // pass number, as the director counts them
            for (long long it = pair_id; it < n_items; it += n_pairs) {
                long long qt, rt0, rt1, rot; int seg;
                item_range(it, qt, rt0, rt1, seg, rot);
                // box of the super-tile's singular values and its largest rounded norm
                float bx0 = 3.0e38f, bx1 = 0.f, by0 = 3.0e38f, by1 = 0.f, bz0 = 3.0e38f, bz1 = 0.f, gq_max = 0.f;
                for (int j = 0; j < UMMA_M / 32; ++j) {
                    const long long row = min(qt * UMMA_M + lane + 32 * j, a.n_q - 1);
                    const float4 sq = __ldg(a.q_sig + a.q_begin + row);
                    bx0 = fminf(bx0, sq.x); bx1 = fmaxf(bx1, sq.x); by0 = fminf(by0, sq.y); by1 = fmaxf(by1, sq.y);
                    bz0 = fminf(bz0, sq.z); bz1 = fmaxf(bz1, sq.z);
                    gq_max = fmaxf(gq_max, __ldg(a.q_G + a.q_begin + row));
                }
                bx0 = wminf(bx0); bx1 = wmaxf(bx1); by0 = wminf(by0); by1 = wmaxf(by1); bz0 = wminf(bz0); bz1 = wmaxf(bz1);
                gq_max = wmaxf(gq_max);
                const float pre_g = a.pre_rel * (sqrtf(gq_max) + a.pre_sqrt_gmax) + 1e-6f;
                // reference scalars one tile ahead: lane l holds frames l and 32 + l (< 48) of the tile
                float4 nsA = make_float4(0.f, 0.f, 0.f, 0.f), nsB = nsA;
                float ngA = 0.f, ngB = 0.f;
                auto fetch_refs = [&](long long ti2) {
                    const long long rb = tile_at(ti2, rt0, rt1, rot) * TR;
                    nsA = __ldg(a.r_sig + rb + lane); ngA = __ldg(a.r_G + rb + lane);
                    if (lane < TR - 32) { nsB = __ldg(a.r_sig + rb + 32 + lane); ngB = __ldg(a.r_G + rb + 32 + lane); }
                };
                const long long n_tiles = rt1 - rt0;
                fetch_refs(0);
                for (long long ti = 0; ti < n_tiles; ++ti, ++g) {
                    const float4 sA = nsA, sB = nsB;
                    const float gA = ngA, gB = ngB;
                    if (ti + 1 < n_tiles) fetch_refs(ti + 1);
                    auto score = [&](const float4 &sg, float gr) {
                        const float d0 = fmaxf(fmaxf(bx0 - sg.x, sg.x - bx1), 0.0f), d1 = fmaxf(fmaxf(by0 - sg.y, sg.y - by1), 0.0f),
                                    d2 = fmaxf(fmaxf(bz0 - sg.z, sg.z - bz1), 0.0f);
                        const float lb = sqrtf(fmaf(d0, d0, fmaf(d1, d1, d2 * d2))) - pre_g;
                        // key = rounded-structure RMSD^2 + accumulation bias/noise (< 2e-5 E0): keep clear of it
                        const float v = fmaf(lb, lb, -2e-5f * (gq_max + gr));
                        return (lb > 0.0f && v == v) ? v : kNegInf;
                    };
                    float m = fminf(score(sA, gA), lane < TR - 32 ? score(sB, gB) : kPosInf);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
                    // slot g & 3 held pass g - 4: the director is done with it once it has begun pass g - 3
                    if (lane == 0) {
                        uint32_t polls = 0;
                        while ((int32_t)(g - ld_volatile_u32(&c.pass_now)) > 3) {
                            __nanosleep(100);
                            if (++polls > 100000000u) { printf("rms_sweep_tc2: scout timeout block=%d pass=%u\n", blockIdx.x, g); __trap(); }
                        }
                        *reinterpret_cast<volatile unsigned long long *>(&c.mtile[g & 3]) = (unsigned long long)(g + 1u) | ((unsigned long long)__float_as_uint(m) << 32);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer + pass director (leader CTA) ================
        // One warp is the whole control path of the tensor pipe, and a single warp retires an instruction every few
        // clocks at best: whatever it executes between two groups of MMAs is time the pipe idles once its short queue has
        // drained (measured: a pass cost 6.7k clk with the MMAs skipped, 10.7k with them -- purely additive).  So this
        // loop is kept to the bone: the director is one comparison (the scout above did the arithmetic), the readiness
        // test of the NEXT group's stages is issued before the MMAs of the current one, descriptors advance by constant
        // offsets, and nothing but the MMAs and their commits sits between two groups.
        // (The MMAs are issued by an elected lane of the CONVERGED warp: with the loop run by one thread inside a divergent
        // branch every tcgen05.mma blocks its thread for the ~72 clk it executes and a commit for ~220 -- measured, C3 62 -> 90 ms.)
        if (rank == 0) {
            int s = 0;
            uint32_t ph = 0, tph = 0, rfph = 0, iph = 0, seq = 0, gpass = 0;
            bool rdy = false;
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t res_lo = umma_desc_lo(__shfl_sync(0xffffffffu, smem_u32(smem), 0));
            const uint32_t ring_lo = umma_desc_lo(__shfl_sync(0xffffffffu, smem_u32(ring), 0));
            const uint32_t peer_mailbox = map_to_cta(&c.mailbox, 1);
            // per plane: shared-memory k-steps, descriptor / TMEM address of k-step 0 (k-step ks: + (ks >> 1) * 512 + (ks & 1) * 2,
            // resp. + 8 ks)
            const int lim0 = 2 * a.chunks[0], lim1 = 2 * a.chunks[1], lim2 = 2 * a.chunks[2];
            const uint32_t ss0 = res_lo + (uint32_t)a.chunk_base[0] * (A_CHUNK >> 4), ss1 = res_lo + (uint32_t)a.chunk_base[1] * (A_CHUNK >> 4),
                           ss2 = res_lo + (uint32_t)a.chunk_base[2] * (A_CHUNK >> 4);
            const uint32_t ts0 = tmem_u + ACC_COLS + 8u * (uint32_t)(a.tmem_unit0[0] - lim0), ts1 = tmem_u + ACC_COLS + 8u * (uint32_t)(a.tmem_unit0[1] - lim1),
                           ts2 = tmem_u + ACC_COLS + 8u * (uint32_t)(a.tmem_unit0[2] - lim2);
            // layouts that give every plane the same split (what tc2_layout prefers) have only two kinds of stage: all three
            // A tiles in shared memory, or all three in TMEM -- the straight-line blocks of tc_ptx.cuh
            const bool uniform_split = lim0 == lim1 && lim1 == lim2 && !(a.dbg & 8192);
            const int nks = (a.dbg & 2) ? 0 : a.nks;
            const uint32_t idesc = a.idesc;
            const uint32_t bhi = a.wide ? kDescHiSw128 : kDescHiSw64;
            const bool no_director = (a.dbg & 64) != 0;
            long long t_total0 = MDSCTK_TC_PROF_BUILD ? clock64() : 0, t_wait_full = 0, t_wait_empty = 0, n_pass = 0, n_heavy = 0, t_tma = 0, n_tma = 0;
            for (long long it = pair_id; it < n_items; it += n_pairs) {
                long long qt, rt0, rt1, rot; int seg;
                item_range(it, qt, rt0, rt1, seg, rot);
                seq = (seq + 1) & 2047u;
                // this item's TMEM-resident k-steps are written and the thresholds published; its shared-memory chunks landed
                mbar_wait_cluster(&c.bar_item_ready, iph, 8); iph ^= 1;
                mbar_wait_spin(&c.bar_res_full, rfph, 6); rfph ^= 1;
                tc_fence_after();
                bool pending_release = false;
                const long long n_tiles = rt1 - rt0;
                for (long long ti = 0; ti < n_tiles; ++ti) {
                    // the first stage of the pass must have landed before the scout's word for this tile is read
                    if (!rdy) {
                        const long long t0 = (MDSCTK_TC_PROF_BUILD && a.prof) ? clock64() : 0;
                        mbar_wait_spin(&c.bar_full[s], ph, 3);
                        if (MDSCTK_TC_PROF_BUILD && a.prof) t_wait_full += clock64() - t0;
                        rdy = true;
                    }
                    // ---- director: can any of the 48 frames hold a neighbour of any of the 256 rows? ----
                    const float tq = lane < 8 ? *reinterpret_cast<volatile float *>(&c.qtau[lane]) : 0.0f;   // only decreases: stale is safe
                    if (lane == 0) *reinterpret_cast<volatile uint32_t *>(&c.pass_now) = gpass;
                    float m = __uint_as_float(0xff800000u);          // no scout: every pass is live
                    if (use_pre) {
                        unsigned long long w;
                        uint32_t polls = 0;
                        while ((uint32_t)(w = *reinterpret_cast<volatile unsigned long long *>(&c.mtile[gpass & 3])) != gpass + 1u)
                            if (++polls > 200000000u) { printf("rms_sweep_tc2: director timeout block=%d pass=%u\n", blockIdx.x, gpass); __trap(); }
                        m = __uint_as_float((uint32_t)(w >> 32));
                    }
                    ++gpass;
                    const bool heavy = !no_director && m <= wmaxf(tq);
                    // ---- the previous heavy pass must have been handed back before anything touches TMEM or the mailbox ----
                    if (pending_release) {
                        const long long t0 = (MDSCTK_TC_PROF_BUILD && a.prof) ? clock64() : 0;
                        mbar_wait_spin(&c.bar_tmem_empty, tph, 2); tph ^= 1;
                        pending_release = false;
                        tc_fence_after();
                        if (MDSCTK_TC_PROF_BUILD && a.prof) t_wait_empty += clock64() - t0;
                    }
                    if (heavy) {
                        const uint32_t msg = (seq << 20) | (uint32_t)(ti + 1);
                        if (lane == 0) {
                            *reinterpret_cast<volatile uint32_t *>(&c.mailbox) = msg;
                            st_cluster_u32(peer_mailbox, msg);
                        }
                        ++n_heavy;
                    }
                    ++n_pass;
                    // ---- MMAs of the pass, one ring stage (two k-steps, or four with 128-byte rows) per round ----
                    for (int kc0 = 0; kc0 < nkc; ++kc0) {
                        if (!rdy) {
                            const long long t0 = (MDSCTK_TC_PROF_BUILD && a.prof) ? clock64() : 0;
                            mbar_wait_spin(&c.bar_full[s], ph, 3);
#if MDSCTK_TC_PROF_BUILD
                            if (a.prof) { t_tma += clock64() - *reinterpret_cast<volatile long long *>(&c.t_issue[s]); ++n_tma; t_wait_full += clock64() - t0; }
#endif
                        }
                        // this round's stage, and the readiness test of the next one (issued now, looked at afterwards)
                        const int sA = s;
                        if (++s == nst) { s = 0; ph ^= 1; }
                        const bool rdy_next = mbar_test_wait(&c.bar_full[s], ph);
                        if (elect_one()) {
                            // Which operand form a plane uses can only change between pairs of k-steps (the shared-memory part of
                            // a plane is whole 64-byte chunks), so there is one uniform branch per pair and the MMAs themselves
                            // are straight-line code whose descriptors differ by constants: the issuing thread must not need
                            // more than the ~72 clk an MMA executes for (measured: 114 clk per MMA with per-MMA address
                            // arithmetic and predication).
                            const int ks0 = (a.wide ? 4 : 2) * kc0;
                            auto issue_stage = [&](uint32_t b, int ks, uint32_t acc0, bool two) {
                                const uint32_t so = (uint32_t)(ks >> 1) * (A_CHUNK >> 4);
                                if (ks < lim0) { tc_mma2_lo<true>(tmem_u, ss0 + so, b, idesc, acc0); if (two) tc_mma2_lo<true>(tmem_u, ss0 + so + 2u, b + 2u, idesc, 1u); }
                                else { tc_mma2_ts_lo(tmem_u, ts0 + 8u * (uint32_t)ks, b, idesc, acc0); if (two) tc_mma2_ts_lo(tmem_u, ts0 + 8u * (uint32_t)ks + 8u, b + 2u, idesc, 1u); }
                                if (ks < lim1) { tc_mma2_lo<true>(tmem_u + UMMA_N, ss1 + so, b, idesc, acc0); if (two) tc_mma2_lo<true>(tmem_u + UMMA_N, ss1 + so + 2u, b + 2u, idesc, 1u); }
                                else { tc_mma2_ts_lo(tmem_u + UMMA_N, ts1 + 8u * (uint32_t)ks, b, idesc, acc0); if (two) tc_mma2_ts_lo(tmem_u + UMMA_N, ts1 + 8u * (uint32_t)ks + 8u, b + 2u, idesc, 1u); }
                                if (ks < lim2) { tc_mma2_lo<true>(tmem_u + 2 * UMMA_N, ss2 + so, b, idesc, acc0); if (two) tc_mma2_lo<true>(tmem_u + 2 * UMMA_N, ss2 + so + 2u, b + 2u, idesc, 1u); }
                                else { tc_mma2_ts_lo(tmem_u + 2 * UMMA_N, ts2 + 8u * (uint32_t)ks, b, idesc, acc0); if (two) tc_mma2_ts_lo(tmem_u + 2 * UMMA_N, ts2 + 8u * (uint32_t)ks + 8u, b + 2u, idesc, 1u); }
                            };
                            auto issue_fast = [&](uint32_t b, int ks, uint32_t acc0) {
                                if (ks + 1 < lim0) {
                                    const uint32_t so = (uint32_t)(ks >> 1) * (A_CHUNK >> 4);
                                    tc2_issue_stage_ss(tmem_u, ss0 + so, ss1 + so, ss2 + so, b, idesc, acc0, bhi);
                                } else {
                                    const uint32_t to = 8u * (uint32_t)ks;
                                    tc2_issue_stage_ts(tmem_u, ts0 + to, ts1 + to, ts2 + to, b, idesc, acc0, ks + 1 < nks, bhi);
                                }
                            };
                            if (ks0 < nks) {
                                const uint32_t b0 = ring_lo + (uint32_t)sA * ((uint32_t)bst >> 4);
                                if (uniform_split) issue_fast(b0, ks0, kc0 != 0);
                                else issue_stage(b0, ks0, kc0 != 0, ks0 + 1 < nks);
                                // wide stage: k-steps 2 and 3 of its 64 atoms sit 64 bytes into the 128-byte rows
                                if (a.wide && ks0 + 2 < nks) issue_fast(b0 + 4u, ks0 + 2, 1u);
                            }
                            tc_commit2_mc(&c.bar_empty[sA], 3);
#if MDSCTK_TC_PROF_BUILD
                            if (a.prof) *reinterpret_cast<volatile long long *>(&c.t_commit[sA]) = clock64();
#endif
                            if (kc0 == nkc - 1) {
                                if (heavy) tc_commit2_mc(&c.bar_tmem_full, 3);
                                if (ti == n_tiles - 1) tc_commit2_mc(&c.bar_res_empty, 3);      // item done with its fit tile
                            }
                        }
                        __syncwarp();
                        rdy = rdy_next;
                    }
                    pending_release = heavy;
                }
                if (pending_release) { mbar_wait_spin(&c.bar_tmem_empty, tph, 2); tph ^= 1; tc_fence_after(); }
                if (lane == 0) {
                    const uint32_t msg = (seq << 20) | END_PASS;
                    *reinterpret_cast<volatile uint32_t *>(&c.mailbox) = msg;
                    st_cluster_u32(peer_mailbox, msg);
                }
                __syncwarp();
            }
            if (MDSCTK_TC_PROF_BUILD && a.prof && lane == 0) {
                long long *pr = a.prof + (size_t)blockIdx.x * 8;
                pr[0] = clock64() - t_total0; pr[1] = t_wait_empty; pr[2] = t_wait_full; pr[3] = n_pass; pr[4] = n_heavy;
                pr[7] = n_tma > 0 ? t_tma / n_tma : 0;
            }
        }
    } else {
        // =============================== epilogue (both CTAs) ===================================
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int sub = ew >> 2;
        const int e_of_quarter = (quarter + 2) & 3;   // ew = sub * 4 + e_of_quarter
        const int row_in_tile = quarter * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const uint32_t t_warp = t_lane + (sub >> 1) * (3 * TRH) + (sub & 1) * SUBW;
        const uint32_t empty_leader = map_to_cta(&c.bar_tmem_empty, 0);
        const uint32_t ready_leader = map_to_cta(&c.bar_item_ready, 0);
        const uint32_t qtau_leader = map_to_cta(&c.qtau[rank * 4 + quarter], 0);
        const unsigned lt_mask = (1u << lane) - 1u;
        const int QC = a.qcap;
        unsigned *hist = c.scratch + (size_t)ew * 5 * QC;
        float *q_c2 = reinterpret_cast<float *>(hist), *q_c1 = q_c2 + QC, *q_c0 = q_c2 + 2 * QC, *q_e0 = q_c2 + 3 * QC;
        int *q_tag = reinterpret_cast<int *>(hist) + 4 * QC;          // (reference column within the warp's 12) << 8 | own lane | 32 (nofit)
        unsigned *wcnt8 = c.cnt8[ew];
        const size_t row_stride = (size_t)a.cl.H * a.cl.cap;
        uint32_t hph = 0, eph = 0, last_msg = 0;
        bool first_item = true;
        int qn = 0;
        size_t lbase0 = 0;
        long long rb_cur = 0;
        float *lkeys = a.cl.key;
        int *lidxs = a.cl.idx;

        auto drain = [&]() {
            __syncwarp();
            if (lane < qn) {
                const int tag = q_tag[lane];
                const int l = tag & 31;
                const float e0 = q_e0[lane];
                const float tau_r = *reinterpret_cast<volatile float *>(&c.tau[quarter * 32 + l]);
                float d2;
                if (tag & 32) {                       // --nofit: q_c2 holds trace S
                    d2 = fmaxf(2.0f * (e0 - q_c2[lane]), 0.0f);
                } else {
                    QcpCoef cf; cf.c2 = q_c2[lane]; cf.c1 = q_c1[lane]; cf.c0 = q_c0[lane];
                    const float xn = qcp_newton_step(cf, e0);
                    d2 = qcp_refine(cf, e0, e0, (xn == xn) ? xn : e0, SC * tau_r);
                }
                d2 *= INV_SC;
                if (d2 < tau_r) {
                    const unsigned old = atomicAdd(&wcnt8[l >> 2], 1u << (8 * (l & 3)));
                    const int pos = (int)((old >> (8 * (l & 3))) & 255u);
                    if (pos < SUB_APP) {
                        const size_t at = lbase0 + (size_t)l * row_stride + a.cl.keep + sub * SUB_APP + pos;
                        lkeys[at] = d2;
                        lidxs[at] = (int)(rb_cur + (tag >> 8));
                    }
                }
            }
            qn = 0;
            __syncwarp();
        };
        auto push = [&](bool pred, float c2, float c1, float c0, float e0, int tag) {
            unsigned m = __ballot_sync(0xffffffffu, pred);
            while (m) {                               // warp-uniform: at most two rounds (32 candidates, QCAP slots)
                if (qn == QC) drain();
                const int rank = __popc(m & lt_mask);
                const bool now = pred && rank < QC - qn;
                if (now) {
                    const int p = qn + rank;
                    q_c2[p] = c2; q_c1[p] = c1; q_c0[p] = c0; q_e0[p] = e0; q_tag[p] = tag;
                }
                const unsigned took = __ballot_sync(0xffffffffu, now);
                qn += __popc(took);
                m &= ~took;
                pred = pred && !now;
            }
        };
        auto cnt_of = [&](int e, int l) { return (int)((c.cnt8[e][l >> 2] >> (8 * (l & 3))) & 255u); };
        // Quarter-wide merge of the rows this warp is responsible for (8 per warp); afterwards the quarter's largest
        // threshold is published to the director.
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
                    cs[s2] = min(cnt_of(s2 * 4 + e_of_quarter, l), SUB_APP);
                    cmax = max(cmax, cs[s2]);
                }
                if (!final && cmax < SUB_APP - MERGE_EVERY * SUBW) continue;
                const size_t at = lbase0 + (size_t)l * row_stride;
                int total = c.mcnt[row];
#pragma unroll
                for (int s2 = 0; s2 < SUBS; ++s2) {
                    const size_t src = at + a.cl.keep + s2 * SUB_APP;
                    for (int base = 0; base < cs[s2]; base += 32) {
                        const int i = base + lane;
                        float kv = 0.f; int iv = 0;
                        if (i < cs[s2]) { kv = lkeys[src + i]; iv = lidxs[src + i]; }
                        __syncwarp();
                        if (i < cs[s2]) { lkeys[at + total + i] = kv; lidxs[at + total + i] = iv; }
                        __syncwarp();
                    }
                    total += cs[s2];
                }
                float tl = c.tau[row];
                if (total > a.cl.keep) {
                    tl = fminf(tl, warp_compact_list6(lkeys + at, lidxs + at, total, a.cl.keep, hist));
                    total = a.cl.keep;
                }
                __syncwarp();
                if (lane == 0) {
                    c.tau[row] = tl;
                    c.mcnt[row] = (unsigned short)total;
#pragma unroll
                    for (int s2 = 0; s2 < SUBS; ++s2)
                        atomicAnd(&c.cnt8[s2 * 4 + e_of_quarter][l >> 2], ~(255u << (8 * (l & 3))));
                    if (final) {
                        const size_t lid = (size_t)qr * a.cl.H + seg;
                        a.cl.cnt[lid] = total;
                        a.cl.tau[lid] = tl;
                        atomicMin(reinterpret_cast<unsigned *>(a.row_tau + qr), __float_as_uint(tl));
                    }
                }
            }
            quarter_sync(quarter);
            if (sub == 0 && !final) {
                const float tq = wmaxf(c.tau[row_in_tile]);
                if (lane == 0) st_cluster_u32(qtau_leader, __float_as_uint(tq));
            }
        };

        for (long long it = pair_id; it < n_items; it += n_pairs) {
            long long qt, rt0, rt1, rot; int seg;
            item_range(it, qt, rt0, rt1, seg, rot);
            const long long row0 = qt * UMMA_M + rank * TQ;
            const long long qrow = row0 + row_in_tile;
            const bool qvalid = qrow < a.n_q;
            const long long qabs = a.q_begin + (qvalid ? qrow : a.n_q - 1);
            const float hgq = (0.5f * SC) * a.q_G[qabs];
            lbase0 = ((size_t)(row0 + quarter * 32) * a.cl.H + seg) * a.cl.cap;
            const float4 sq = __ldg(a.q_sig + qabs);
            const float gq_nm = 2.0f * INV_SC * hgq;
            const float bx0 = wminf(sq.x), bx1 = wmaxf(sq.x), by0 = wminf(sq.y), by1 = wmaxf(sq.y), bz0 = wminf(sq.z), bz1 = wmaxf(sq.z);
            const float gq_max = wmaxf(gq_nm);
            const float pre_g = a.pre_rel * (sqrtf(gq_max) + a.pre_sqrt_gmax) + 1e-6f;
            if (lane < 8) wcnt8[lane] = 0;
            if (sub == 0) {
                c.mcnt[row_in_tile] = 0;
                c.tau[row_in_tile] = qvalid ? __ldcg(a.row_tau + qrow) : 0.0f;
            }
            quarter_sync(quarter);
            if (sub == 0) {
                const float tq = wmaxf(c.tau[row_in_tile]);
                if (lane == 0) st_cluster_u32(qtau_leader, __float_as_uint(tq));
            }
            // TMEM-resident k-steps of this item's fit tile: the previous item's MMAs must have retired first
            if (!first_item) {
                if (lane == 0) mbar_wait(&c.bar_res_empty, eph, 9);
                eph ^= 1;
                __syncwarp();
            }
            first_item = false;
            tc_fence_after();
            for (int u = sub; u < a.n_tmem_units; u += SUBS) {
                int p = 2;
                while (p > 0 && u < a.tmem_unit0[p]) --p;
                const int ks = 2 * a.chunks[p] + (u - a.tmem_unit0[p]);
                uint32_t v[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                if (qvalid) {
                    const uint4 *src = reinterpret_cast<const uint4 *>(a.q_fh + ((size_t)qabs * 3 + p) * a.A_pad + ks * 16);
                    const uint4 x0 = __ldg(src), x1 = __ldg(src + 1);
                    v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w; v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
                }
                tc_st8(t_lane + ACC_COLS + 8u * (uint32_t)u, v);
            }
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(ready_leader);      // release.cluster: also publishes the quarter threshold

            int hcount = 0;
            const long long n_tiles = rt1 - rt0;
            while (true) {
                // ---- wait for the director: next heavy pass of this item, or its end ----
                uint32_t msg = 0;
                if (lane == 0) {
                    uint32_t polls = 0;
                    while ((msg = ld_volatile_u32(&c.mailbox)) == last_msg) {
                        __nanosleep(a.dbg & 512 ? 64 : 400);      // a heavy pass starts with >= 4k clk of MMAs: no hurry, and idle warps should not burn power
                        if (++polls > 400000000u) { printf("rms_sweep_tc2: mailbox timeout block=%d warp=%d last=%x\n", blockIdx.x, warp, last_msg); __trap(); }
                    }
                }
                msg = __shfl_sync(0xffffffffu, msg, 0);
                last_msg = msg;
                if ((msg & END_PASS) == END_PASS) break;
                const long long ti = (long long)(msg & END_PASS) - 1;
                const long long rt = tile_at(ti, rt0, rt1, rot);
                const long long rb = rt * TR + sub * SUBW;
                rb_cur = rb;
                float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
                float cg = 0.f;
                if (lane < SUBW) { cs = __ldg(a.r_sig + rb + lane); cg = __ldg(a.r_G + rb + lane); }
                // per-warp pre-bound (finer than the director's: this warp's 32 rows, its 12 frames): which batches to read
                unsigned live = (1u << (SUBW / EB)) - 1u;
                if (a.pre_rel >= 0.0f && a.do_fit && !a.debug_tile && !(a.dbg & 1024)) {
                    const float tau_max = wmaxf(*reinterpret_cast<volatile float *>(&c.tau[row_in_tile]));
                    const float d0 = fmaxf(fmaxf(bx0 - cs.x, cs.x - bx1), 0.0f), d1 = fmaxf(fmaxf(by0 - cs.y, cs.y - by1), 0.0f),
                                d2 = fmaxf(fmaxf(bz0 - cs.z, cs.z - bz1), 0.0f);
                    const float lb = sqrtf(fmaf(d0, d0, fmaf(d1, d1, d2 * d2))) - pre_g;
                    const bool far = lane >= SUBW || (lb > 0.0f && lb * lb > tau_max + 2e-5f * (gq_max + cg));
                    const unsigned near = ~__ballot_sync(0xffffffffu, far);
                    live = 0;
#pragma unroll
                    for (int hb = 0; hb < SUBW / EB; ++hb)
                        if (near & (((1u << EB) - 1u) << (hb * EB))) live |= 1u << hb;
                }
                if (lane == 0) mbar_wait(&c.bar_tmem_full, hph, 4);
                hph ^= 1;
                __syncwarp();
                tc_fence_after();
                const float tau = SC * *reinterpret_cast<volatile float *>(&c.tau[row_in_tile]);
                const float htau = 0.5f * tau;
                float svb[2][9][EB];
                auto load_batch = [&](int h, float (&dst)[9][EB]) {
#pragma unroll
                    for (int p = 0; p < 3; ++p)
#pragma unroll
                        for (int b = 0; b < 3; ++b)
                            tc_ld2(t_warp + p * UMMA_N + b * TRH + h, dst, p * 3 + b);
                };
                bool released = false;
                auto release = [&]() {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster_nofence(empty_leader);
                    released = true;
                };
                if (live == 0) release();
                if (live & 1u) load_batch(0, svb[0]);
#pragma unroll
                for (int hb = 0; hb < SUBW / EB; ++hb) {
                    const int h = hb * EB;
                    float (&sv)[9][EB] = svb[hb & 1];
                    const bool cur = (live >> hb) & 1u;
                    if (cur) tc_wait_ld();
                    if (hb + 1 < SUBW / EB && ((live >> (hb + 1)) & 1u)) load_batch(h + EB, svb[(hb + 1) & 1]);
                    if (!released && (live >> (hb + 1)) == 0) release();
                    if (!cur) continue;
                    if (a.debug_tile && it == 0 && rt == 0 && rank == 0) {
#pragma unroll
                        for (int cc = 0; cc < 9; ++cc)
#pragma unroll
                            for (int j = 0; j < EB; ++j)
                                a.debug_tile[(size_t)row_in_tile * (9 * TR) + cc * TR + sub * SUBW + h + j] = sv[cc][j] * INV_SC;
                    }
                    float e0[EB];
#pragma unroll
                    for (int j = 0; j < EB; ++j) e0[j] = fmaf(0.5f * SC, __shfl_sync(0xffffffffu, cg, h + j), hgq);
                    if (a.dbg & 1) {
                        float acc = 0.f;
#pragma unroll
                        for (int j = 0; j < EB; ++j) acc += sv[0][j] + sv[4][j] + sv[8][j] + e0[j];
                        if (acc == 12345.f) wcnt8[lane & 7] = 1;
                        continue;
                    }
                    if (!a.do_fit) {
#pragma unroll
                        for (int j = 0; j < EB; ++j) {
                            const float tr = sv[0][j] + sv[4][j] + sv[8][j];
                            push(qvalid && rb + h + j < a.n_r && 2.0f * (e0[j] - tr) < tau, tr, 0.f, 0.f, e0[j], ((h + j) << 8) | lane | 32);
                        }
                        continue;
                    }
                    float f[EB];
                    bool far = true;
#pragma unroll
                    for (int j = 0; j < EB; ++j) {
                        f[j] = qcp_frob2(sv, j);
                        const float t = e0[j] - htau;
                        far = far && (t > 0.0f) && (3.0001f * f[j] < t * t);
                    }
                    if (!(a.dbg & 8) && __all_sync(0xffffffffu, far)) continue;
                    QcpCoef cf[EB];
                    float x1[EB];
#pragma unroll
                    for (int j = 0; j < EB; ++j) {
                        const float s9[9] = {sv[0][j], sv[1][j], sv[2][j], sv[3][j], sv[4][j], sv[5][j], sv[6][j], sv[7][j], sv[8][j]};
                        cf[j] = qcp_coefficients(s9, f[j]);
                    }
#pragma unroll
                    for (int j = 0; j < EB; ++j) {
                        const float xn = qcp_newton_step(cf[j], e0[j]);
                        x1[j] = (xn == xn) ? xn : e0[j];
                    }
#pragma unroll
                    for (int j = 0; j < EB; ++j)
                        push(qvalid && rb + h + j < a.n_r && !(2.0f * (e0[j] - x1[j]) > tau), cf[j].c2, cf[j].c1, cf[j].c0, e0[j],
                             ((h + j) << 8) | lane);
                }
                drain();
                ++hcount;
                if ((hcount & (MERGE_EVERY - 1)) == 0) merge_rows(row0, seg, false);
            }
            (void)n_tiles;
            merge_rows(row0, seg, true);
        }
    }

    // ---- teardown --------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(tc2::TMEM_COLS));
    }
}

// ---------------------------------------------------------------- host side ----
namespace {

struct Tc2Layout {
    int nks, nkc, chunks[3], chunk_base[3], tmem_unit0[3], n_tmem_units, n_res_chunks, nst, smem_bytes;
    int qcap, wide, bst;
    bool ok;
};

// Where the CTA's fit tile lives: whole 64-byte chunks (two k-steps) per plane in shared memory, the trailing
// k-steps of each plane in the spare TMEM columns (at most 10 k-steps in all), and a ring of >= 4 reference stages.
static int tc2_ctl_bytes(int qcap)
{
    return (int)((offsetof(tc2::Ctl, scratch) + (size_t)tc2::EPI_WARPS * 5 * qcap * sizeof(unsigned) + 15) / 16 * 16);
}

Tc2Layout tc2_layout(int A_pad, int want_wide)
{
    Tc2Layout L = {};
    L.nks = A_pad / 16;
    L.nkc = (A_pad + 31) / 32;
    const int ctl = tc2_ctl_bytes(tc2::QCAP);
    int best_nst = 0;
    bool best_equal = false;
    // z gives up k-steps to TMEM first (then y, x): try every split and keep the one with the deepest ring
    for (int tz = 0; tz <= tc2::MAX_TMEM_UNITS && tz <= L.nks; ++tz)
        for (int ty = 0; ty <= tz && tz + ty <= tc2::MAX_TMEM_UNITS; ++ty)
            for (int tx = 0; tx <= ty && tz + ty + tx <= tc2::MAX_TMEM_UNITS; ++tx) {
                const int t[3] = {tx, ty, tz};
                int chunks = 0;
                bool even = true;
                for (int p = 0; p < 3; ++p) {
                    const int in_smem = L.nks - t[p];
                    if (in_smem < 0 || (in_smem & 1)) { even = false; break; }
                    chunks += in_smem / 2;
                }
                if (!even) continue;
                const int room = tc2::SMEM_MAX - chunks * tc2::A_CHUNK - ctl;
                int nst = room / tc2::B_STAGE;
                if (nst > tc2::MAX_NST) nst = tc2::MAX_NST;
                // prefer a deeper ring, then the same split for every plane (two kinds of stage only: straight-line issue)
                const bool equal = tx == ty && ty == tz;
                if (nst > best_nst || (nst == best_nst && equal && !best_equal)) {
                    best_equal = equal;
                    best_nst = nst;
                    L.n_tmem_units = tx + ty + tz;
                    L.n_res_chunks = chunks;
                    L.nst = nst;
                    int base = 0, ubase = 0;
                    for (int p = 0; p < 3; ++p) {
                        L.chunks[p] = (L.nks - t[p]) / 2; L.chunk_base[p] = base; base += L.chunks[p];
                        L.tmem_unit0[p] = ubase; ubase += t[p];
                    }
                }
            }
    L.ok = best_nst >= 4 && A_pad % 16 == 0;
    L.qcap = tc2::QCAP; L.wide = 0; L.bst = tc2::B_STAGE;
    L.smem_bytes = L.n_res_chunks * tc2::A_CHUNK + L.nst * tc2::B_STAGE + ctl;
    // Wide stages (64 atoms, 128-byte rows): half the producer / issuer rounds per pass -- a round costs each of the two
    // single threads ~500 clk whatever it moves, which is what paced the light passes -- if three of them fit (a smaller
    // refine queue may have to pay for the third).  Needs the same split for every plane (straight-line issue blocks), so
    // the splits are searched again: t k-steps of every plane in TMEM; a ring of up to five stages is worth more than the
    // full-size refine queue (heavy passes drain it in batches of that size), the queue more than a still deeper ring.
    if (L.ok && want_wide) {
        long best_score = -1;
        for (int t = 0; 3 * t <= tc2::MAX_TMEM_UNITS && t <= L.nks; ++t) {
            if ((L.nks - t) & 1) continue;
            const int chunks = 3 * ((L.nks - t) / 2);
            for (int qc : {tc2::QCAP, tc2::QCAP_SMALL}) {
                const int room = tc2::SMEM_MAX - chunks * tc2::A_CHUNK - tc2_ctl_bytes(qc);
                int nw = room / tc2::B_STAGE_WIDE;
                if (nw > tc2::MAX_NST) nw = tc2::MAX_NST;
                const long score = 1000L * (nw < 5 ? nw : 5) + 100L * (qc == tc2::QCAP) + nw;
                if (nw >= 3 && score > best_score) {
                    best_score = score;
                    L.wide = 1; L.bst = tc2::B_STAGE_WIDE; L.nst = nw; L.qcap = qc; L.nkc = (A_pad + 63) / 64;
                    L.n_tmem_units = 3 * t; L.n_res_chunks = chunks;
                    for (int p = 0; p < 3; ++p) { L.chunks[p] = chunks / 3; L.chunk_base[p] = p * (chunks / 3); L.tmem_unit0[p] = p * t; }
                    L.smem_bytes = chunks * tc2::A_CHUNK + nw * tc2::B_STAGE_WIDE + tc2_ctl_bytes(qc);
                }
            }
        }
    }
    return L;
}

bool make_map(CUtensorMap *m, const void *planes, long long n, int A_pad, int rows, int box_planes, bool wide = false)
{
    EncodeTiledFn enc = get_tensor_map_encoder();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)A_pad, (cuuint64_t)n, 3};
    cuuint64_t strides[2] = {(cuuint64_t)A_pad * 3 * 2, (cuuint64_t)A_pad * 2};
    cuuint32_t box[3] = {wide ? 64u : 32u, (cuuint32_t)rows, (cuuint32_t)box_planes};      // wide: 128-byte rows, reaching past A_pad is zero-filled
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(planes), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               wide ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

bool rms_tc2_supported(int A_pad) { return tc2_layout(A_pad, 0).ok; }

// host-side description of the layout the launcher would pick (no GPU involved): tests/test_abi_cpu.py sweeps every atom count
void rms_tc2_layout_info(int A_pad, int want_wide, int out[12])
{
    const Tc2Layout L = tc2_layout(A_pad, want_wide);
    out[0] = L.ok; out[1] = L.wide; out[2] = L.nst; out[3] = L.bst; out[4] = L.qcap; out[5] = L.smem_bytes;
    out[6] = L.n_res_chunks; out[7] = L.n_tmem_units; out[8] = L.nkc; out[9] = L.nks;
    out[10] = tc2_ctl_bytes(L.qcap); out[11] = tc2::SMEM_MAX;
}

cudaError_t launch_rms_sweep_tc2(const FrameSetView &fit, long long fit_begin, long long n_fit, const FrameSetView &ref, int do_fit,
                                 int n_seg, CandLists<float> cl, float *row_tau, float g_ref_max, int *own_tile_scratch,
                                 float *debug_tile, int wide_stages, int n_sms, cudaStream_t st)
{
    if (n_fit <= 0) return cudaSuccess;
    if (cl.H != n_seg || cl.cap < cl.keep + tc2::SUBS * tc2::SUB_APP) return cudaErrorInvalidValue;
    int want_wide = wide_stages && !(tc_experiment_bits() & 8192);      // (not with the generic-issue experiment)
#if MDSCTK_TC_EXPERIMENTS
    if (const char *we = getenv("MDSCTK_TC_WIDE")) want_wide = want_wide && atoi(we) != 0;
#endif
    const Tc2Layout L = tc2_layout(ref.A_pad, want_wide);
    if (!L.ok) return cudaErrorInvalidConfiguration;
    CUtensorMap mq, mr;
    if (!make_map(&mq, fit.fh, fit.n, fit.A_pad, tc2::TQ, 1)) return cudaErrorInvalidValue;
    if (!make_map(&mr, ref.fh, ref.n, ref.A_pad, tc2::TRH, 3, L.wide != 0)) return cudaErrorInvalidValue;
    Tc2Args a = {};
    a.q_G = fit.Gh; a.r_G = ref.Gh;
    a.q_begin = fit_begin; a.n_q = n_fit; a.n_r = ref.n;
    a.A_pad = ref.A_pad; a.do_fit = do_fit; a.n_seg = n_seg; a.cl = cl; a.debug_tile = debug_tile; a.row_tau = row_tau;
    a.q_sig = reinterpret_cast<const float4 *>(fit.sig); a.r_sig = reinterpret_cast<const float4 *>(ref.sig);
    a.dbg = tc_experiment_bits();
    a.pre_rel = (a.dbg & 2048) ? -1.0f : 4.9e-4f;           // relative rounding error of one fp16 operand (2^-11)
    a.pre_sqrt_gmax = sqrtf(g_ref_max > 0.f ? g_ref_max : 0.f);
    a.q_fh = static_cast<const __half *>(fit.fh);
    a.nkc = L.nkc; a.nks = L.nks; a.n_tmem_units = L.n_tmem_units; a.n_res_chunks = L.n_res_chunks; a.nst = L.nst;
    a.qcap = L.qcap; a.wide = L.wide; a.bst = L.bst;
    for (int p = 0; p < 3; ++p) { a.chunks[p] = L.chunks[p]; a.chunk_base[p] = L.chunk_base[p]; a.tmem_unit0[p] = L.tmem_unit0[p]; }
    a.idesc = tc2::IDESC;
#if MDSCTK_TC_EXPERIMENTS
    if (const char *en = getenv("MDSCTK_TC_N")) a.idesc = umma_idesc(0, tc2::UMMA_M, atoi(en));     // MMA time vs N (results invalid)
#endif
    a.own_tile = nullptr;
    if (own_tile_scratch && !(a.dbg & 32768)) {
        const long long n_qt = (n_fit + tc2::UMMA_M - 1) / tc2::UMMA_M;
        launch_rms_guess_own_tile(a.q_sig, fit_begin, n_fit, a.r_sig, ref.n, own_tile_scratch, st);
        (void)n_qt;
        a.own_tile = own_tile_scratch;
    }
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
    cudaError_t e = cudaFuncSetAttribute(rms_sweep_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.smem_bytes);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.blockDim = dim3(tc2::NTHR); cfg.dynamicSmemBytes = (size_t)L.smem_bytes; cfg.stream = st;
    int max_pairs = n_sms / 2;
    cfg.gridDim = dim3((unsigned)(max_pairs * 2));
    int q = 0;
    if (cudaOccupancyMaxActiveClusters(&q, rms_sweep_tc2_kernel, &cfg) == cudaSuccess && q > 0 && q < max_pairs) max_pairs = q;
    (void)cudaGetLastError();
    const long long n_items = ((n_fit + tc2::UMMA_M - 1) / tc2::UMMA_M) * n_seg;
    const long long n_pairs = n_items < max_pairs ? n_items : max_pairs;
    cfg.gridDim = dim3((unsigned)(n_pairs * 2));
    e = cudaLaunchKernelEx(&cfg, rms_sweep_tc2_kernel, mq, mr, a);
    if (e != cudaSuccess) return e;
    if (prof) {
        static long long h[1024 * 8];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost);
        double sum[8] = {0}, psum[8] = {0}; int nb = 0;
        for (int b = 0; b < 1024; b += 2) {
            if (h[b * 8] == 0) continue;
            nb++;
            for (int k = 0; k < 8; ++k) { sum[k] += (double)h[b * 8 + k]; psum[k] += (double)h[(b + 1) * 8 + k]; }
        }
        if (nb && sum[3] > 0)
            fprintf(stderr, "[tc2 prof] producer (leader CTA) clk per pass: total %.0f = scout %.0f + wait for free stages %.0f + issue copies %.0f + rest; "
                            "rounds per pass %.2f | issuer per pass: MMA issue %.0f, commits %.0f, next-stage tests %.0f, director %.0f\n",
                    psum[0] / sum[3], 0.0, psum[2] / sum[3], psum[3] / sum[3], psum[4] / sum[3], psum[5] / sum[3], psum[6] / sum[3], psum[7] / sum[3], psum[1] / sum[3]);
        if (nb)
            fprintf(stderr, "[tc2 prof] pairs=%d passes %.0f heavy %.1f%% | clk per pass: total %.0f = wait hand-back %.0f + wait operands %.0f + "
                            "issue and rest | ring round trip: commit -> producer %.0f clk, copy issued -> seen full %.0f clk | layout: %d chunks "
                            "in smem, %d k-steps in TMEM, %d ring stages, %d B smem\n",
                    nb, sum[3] / nb, 100.0 * sum[4] / (sum[3] > 0 ? sum[3] : 1), sum[0] / sum[3], sum[1] / sum[3], sum[2] / sum[3],
                    sum[5] / (sum[6] > 0 ? sum[6] : 1), sum[7] / nb, L.n_res_chunks, L.n_tmem_units, L.nst, L.smem_bytes);
    }
    return cudaGetLastError();
}

}  // namespace mdsctk
