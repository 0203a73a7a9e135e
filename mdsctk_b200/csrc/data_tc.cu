// data_tc.cu -- knn_data (Euclidean) on the 5th-gen tensor cores: the same tile machinery as the RMSD
// sweep (rms_tc.cu) applied to  d^2(x, y) = |x|^2 + |y|^2 - 2 x.y.
//
// Replaces the row loop of knn_data.cpp:195-250 with ::distance = euclidean_distance
// (mdsctk.cpp:330-335) for large inputs.  The tensor-core contraction is only a FILTER: per fit row it
// keeps the k1 + slack smallest approximate distances (streaming top-k, select.cuh); the survivors are
// then re-computed in FP64 with the reference's own operation order (sequential sum of (a-b)^2, no FMA),
// so the distances written are BIT-IDENTICAL to the CPU tool's, and a per-row certificate (same
// argument as rms_rescore.cu) proves that no dropped pair could belong to the k1 nearest.  Rows that
// cannot be certified are recomputed by the exact FP64 sweep (data_knn.cu).
//
//   pack       double rows -> fp16 split of s*x (hi = rn(s x), lo = rn(s x - hi), s a power of two chosen
//              from max|x|), [n][D_pad] row-major = K-major operand rows; |s x|^2 in fp32.
//   sweep      CTA pair (cta_group::2), M = 256 fit rows x N = 256 reference rows per pass, K = 16 per
//              MMA, three MMAs per k-step (hh + hl + lh); TMA-fed 6-stage ring; TMEM double-buffered
//              (2 x 256 columns) so the MMAs of pass i+1 run under the epilogue of pass i.
//   epilogue   16 warps, TMEM lane = fit row: key = nq + nr - 2 dot, compared with the row's admission
//              threshold; survivors are appended to the row's private append area, merged and
//              radix-selected every pass by the quarter's warps.
#include "common.cuh"
#include "select.cuh"
#include "sort.cuh"
#include "tc_ptx.cuh"

#include <cuda_fp16.h>
#include <algorithm>
#include <cstdlib>
#include <cstdio>

#ifndef MDSCTK_TC_PROF_BUILD
#define MDSCTK_TC_PROF_BUILD 0        // 1: compile the clock counters read by MDSCTK_TC_PROF=1 (scripts/build_prof.sh)
#endif

namespace mdsctk {

namespace dtc {
constexpr int TQ = 128;                       // fit rows per CTA
constexpr int TR = 256;                       // reference rows per pass (128 loaded by each CTA)
constexpr int TRH = TR / 2;
constexpr int KC = 32;                        // dims per stage (64-byte rows of fp16)
constexpr int ROW_BYTES = 64;
constexpr int UMMA_M = 2 * TQ, UMMA_N = TR;
constexpr int A_PART = TQ * ROW_BYTES;        // 8192
constexpr int B_PART = TRH * ROW_BYTES;       // 8192
constexpr int OFF_AHI = 0, OFF_ALO = A_PART, OFF_BHI = 2 * A_PART, OFF_BLO = 2 * A_PART + B_PART;
constexpr int STAGE_BYTES = 2 * A_PART + 2 * B_PART;   // 32768
constexpr int NST = 6;
constexpr int SMEM_BYTES = NST * STAGE_BYTES + 1024;
constexpr int MAX_NST = 10;                   // barrier slots (resident mode uses up to 10 ring stages)
constexpr int RES_KC = 64;                    // resident mode: dims per chunk = 128-byte rows (SWIZZLE_128B): every L2 request is a whole line
constexpr int RES_CHUNK = TQ * 2 * RES_KC;    // 16384 B: 128 fit rows x 64 dims
constexpr int RES_STAGE = TRH * 2 * RES_KC;   // 16384 B: one ring stage = 64 dims of this CTA's 128 reference rows
constexpr int STATIC_SMEM = 12 * 1024;        // bound on the kernel's static shared memory (checked at launch)
constexpr int SUBS = 4, EPI_WARPS = 16, NTHR = 64 + EPI_WARPS * 32;
constexpr int SUBW = TR / SUBS;               // 64 reference columns per epilogue warp
constexpr int EB = 16;                        // columns per tcgen05.ld
constexpr int SUB_APP = 2 * SUBW;             // 128: private append area per epilogue warp and row
constexpr int TMEM_COLS = 512;
static_assert(SUBW == 4 * EB, "the epilogue's pass loop is written out for four batches");
}  // namespace dtc

// ------------------------------------------------------------------------------ pack ----
__global__ void data_maxabs_kernel(const double *__restrict__ v, size_t n, unsigned long long *out)
{
    double m = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double a = fabs(v[i]);
        if (a == a) m = fmax(m, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));   // m >= 0
}

cudaError_t launch_data_maxabs(const double *v, size_t n, double *out, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(out, 0, 8, st);
    if (e != cudaSuccess || n == 0) return e;
    data_maxabs_kernel<<<296, 256, 0, st>>>(v, n, reinterpret_cast<unsigned long long *>(out));
    return cudaGetLastError();
}

// one warp per row
// stats != NULL (knn_data -c): the packed row is the STANDARDISED row z = (v - mean) / (spread sqrt(dim - 1)) -- a unit
// vector whose Euclidean distances are the correlation distances: |z_x - z_y|^2 = 2 - 2 r = 4 ((1 - r) / 2), the square
// of twice what correlation_distance (mdsctk.cpp:337-360) returns.  So the Euclidean tensor filter serves both metrics.
__global__ void __launch_bounds__(256) data_pack_kernel(const double *__restrict__ rows, long long n, int dim, int D_pad,
                                                        double scale, const double *__restrict__ stats, __half *__restrict__ hi,
                                                        __half *__restrict__ lo, float *__restrict__ norm, float *__restrict__ norm1,
                                                        float *__restrict__ gres)
{
    const int lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n) return;
    const double *v = rows + (size_t)r * dim;
    double nn = 0.0, n1 = 0.0, r1 = 0.0;
    const double mean = stats ? stats[2 * r] : 0.0;
    const double mul = stats ? scale / (stats[2 * r + 1] * sqrt((double)dim - 1.0)) : scale;
    for (int x = lane; x < D_pad; x += 32) {
        float h = 0.0f, l = 0.0f;
        if (x < dim) {
            const double sv = (v[x] - mean) * mul;
            nn += sv * sv;
            const __half hh = __float2half_rn((float)sv);
            h = __half2float(hh);
            n1 += (double)h * (double)h;                       // the one-part filter contracts hi only: its norm ...
            r1 += (sv - (double)h) * (sv - (double)h);         // ... and how far the rounded row is from the true one
            l = (float)(sv - (double)h);
            hi[(size_t)r * D_pad + x] = hh;
            lo[(size_t)r * D_pad + x] = __float2half_rn(l);
        } else {
            hi[(size_t)r * D_pad + x] = __float2half_rn(0.0f);
            lo[(size_t)r * D_pad + x] = __float2half_rn(0.0f);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nn += __shfl_xor_sync(0xffffffffu, nn, o);
        n1 += __shfl_xor_sync(0xffffffffu, n1, o);
        r1 += __shfl_xor_sync(0xffffffffu, r1, o);
    }
    if (lane == 0) {
        norm[r] = (float)nn;
        if (norm1) { norm1[r] = (float)n1; gres[r] = __double2float_ru(sqrt(r1) / scale); }   // residual norm in INPUT units, rounded up
    }
}

cudaError_t launch_data_pack(const double *rows, long long n, int dim, int D_pad, double scale, const double *stats, void *hi,
                             void *lo, float *norm, float *norm1, float *gres, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    data_pack_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(rows, n, dim, D_pad, scale, stats, static_cast<__half *>(hi),
                                                               static_cast<__half *>(lo), norm, norm1, gres);
    return cudaGetLastError();
}

// rows whose spread is zero or not finite (constant rows, NaNs): correlation_distance divides by it, and such inputs
// are left to the exact FP64 sweep, which reproduces whatever the reference's arithmetic makes of them
__global__ void data_stats_check_kernel(const double *__restrict__ stats, long long n, int *bad)
{
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const double sp = stats[2 * r + 1], mean = stats[2 * r];
        if (!(sp > 0.0) || !(sp < 1.0e300) || !(fabs(mean) < 1.0e300)) atomicAdd(bad, 1);
    }
}

cudaError_t launch_data_stats_check(const double *stats, long long n, int *bad, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(bad, 0, sizeof(int), st);
    if (e != cudaSuccess || n <= 0) return e;
    data_stats_check_kernel<<<296, 256, 0, st>>>(stats, n, bad);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------ sweep ----
struct DataTcArgs {
    const float *q_norm, *r_norm;     // |s x|^2 per row (r_norm readable 256 floats past n_r)
    long long q_begin, n_q, n_r;      // q_begin: reference index of fit row 0 when the fit rows ARE reference rows, else -1
    int D_pad, n_seg;
    float inv_scale2;                 // 1 / s^2: accumulator units -> input units^2
    CandLists<float> cl;              // H = n_seg lists per fit row; key = approximate d^2 in input units
    float *row_tau;
    int one;                          // one-part filter: hi x hi only (q_norm / r_norm are then the norms of the hi parts)
    int res, res_nst;                 // one-part filter with the fit tile RESIDENT in shared memory: on/off, ring stages beside it
    long long *prof;                  // clock counters per pair (prof builds with MDSCTK_TC_PROF=1 only), else NULL
};

__global__ void __launch_bounds__(dtc::NTHR, 1)      // 96 registers: 18 warps put 5 on one scheduler, 16 384 / (5 x 32) = 102
data_sweep_tc_kernel(const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
                     const __grid_constant__ CUtensorMap map_r_hi, const __grid_constant__ CUtensorMap map_r_lo, DataTcArgs a)
{
    using namespace dtc;
    constexpr uint32_t IDESC = umma_idesc(0, UMMA_M, UMMA_N);          // fp16 x fp16 -> fp32
    constexpr uint32_t STAGE_TX = 2u * STAGE_BYTES;

    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[MAX_NST], bar_empty[MAX_NST], bar_tmem_full[2], bar_tmem_empty[2], bar_res_full, bar_res_empty;
    __shared__ uint32_t s_tmem_base;
    __shared__ unsigned s_hist[EPI_WARPS][64];        // 64-bin radix histogram (select.cuh warp_compact_list6): every KB here is ring
    __shared__ int s_cnt[EPI_WARPS][32];
    __shared__ __align__(16) float s_rn[EPI_WARPS][SUBW];      // reference norms of the current pass, per epilogue warp
    __shared__ float s_tau[TQ];
    __shared__ int s_mcnt[TQ];

    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int nk = a.D_pad / KC;
    const long long n_qt = (a.n_q + UMMA_M - 1) / UMMA_M;
    const long long n_rt = (a.n_r + TR - 1) / TR;
    const long long n_items = n_qt * a.n_seg;
    const long long pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < MAX_NST; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&bar_tmem_full[b], 1); mbar_init(&bar_tmem_empty[b], 2 * EPI_WARPS); }
        mbar_init(&bar_res_full, 1);
        mbar_init(&bar_res_empty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                     "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    // diagonal first (see rms_tc.cu): a fit super-tile meets the segment holding its own rows first
    auto item_range = [&](long long it, long long &qt, long long &rt0, long long &rt1, int &seg, long long &rot) {
        int s_rest = -1;
        if (it < n_qt) qt = it;
        else { s_rest = (int)((it - n_qt) / n_qt); qt = (it - n_qt) - (long long)s_rest * n_qt; }
        const long long own = a.q_begin >= 0 ? min((a.q_begin + qt * UMMA_M) / TR, n_rt - 1) : 0;
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

    if (warp == 0) {
        // =============================== TMA producer ===============================
        int s = 0;
        uint32_t ph = 0, rph = 0;
        bool first_item = true;
        for (long long it = pair_id; it < n_items; it += n_pairs) {
            long long qt, rt0, rt1, rot; int seg;
            item_range(it, qt, rt0, rt1, seg, rot);
            const int q0 = (int)(qt * UMMA_M + rank * TQ);
            if (a.res) {
                // Resident mode: the CTA's 128 fit rows (all D_pad dims, 16 KB per 64-dim chunk) are loaded ONCE per work item;
                // the ring carries the reference rows only -- half the L2 -> SM stream of the streaming mode, which is what
                // bounds this kernel (ncu: 1.06 TB per 131 072 x 1M block = 5.7 TB/s) -- in 128-byte rows.
                if (!first_item) { mbar_wait(&bar_res_empty, rph, 5); rph ^= 1; }
                first_item = false;
                const uint32_t res_leader = map_to_cta(&bar_res_full, 0);
                const int nstage = (a.D_pad + RES_KC - 1) / RES_KC;      // a box reaching past D_pad is zero-filled (and counted in full)
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(&bar_res_full, 2u * (uint32_t)(nstage * RES_CHUNK));
                    for (int kc = 0; kc < nstage; ++kc) tma_load_2d_2sm(smem + kc * RES_CHUNK, &map_q_hi, res_leader, kc * RES_KC, q0, kEvictNormal);
                }
                __syncwarp();
                unsigned char *ring = smem + nstage * RES_CHUNK;
                for (long long ti = 0; ti < rt1 - rt0; ++ti) {
                    const int r0 = (int)(tile_at(ti, rt0, rt1, rot) * TR + rank * TRH);
                    for (int j = 0; j < nstage; ++j) {
                        mbar_wait(&bar_empty[s], ph ^ 1, 1);
                        const uint32_t full_leader = map_to_cta(&bar_full[s], 0);
                        if (elect_one()) {
                            if (rank == 0) mbar_expect_tx(&bar_full[s], 2u * (uint32_t)RES_STAGE);
                            tma_load_2d_2sm(ring + s * RES_STAGE, &map_r_hi, full_leader, j * RES_KC, r0, kEvictNormal);
                        }
                        __syncwarp();
                        if (++s == a.res_nst) { s = 0; ph ^= 1; }
                    }
                }
                continue;
            }
            for (long long ti = 0; ti < rt1 - rt0; ++ti) {
                const int r0 = (int)(tile_at(ti, rt0, rt1, rot) * TR + rank * TRH);
                for (int kc = 0; kc < nk; ++kc) {
                    mbar_wait(&bar_empty[s], ph ^ 1, 1);
                    unsigned char *st = smem + s * STAGE_BYTES;
                    const uint32_t full_leader = map_to_cta(&bar_full[s], 0);
                    if (elect_one()) {
                        if (rank == 0) mbar_expect_tx(&bar_full[s], a.one ? STAGE_TX / 2 : STAGE_TX);
                        tma_load_2d_2sm(st + OFF_AHI, &map_q_hi, full_leader, kc * KC, q0, kEvictLast);
                        tma_load_2d_2sm(st + OFF_BHI, &map_r_hi, full_leader, kc * KC, r0, kEvictNormal);
                        if (!a.one) {
                            tma_load_2d_2sm(st + OFF_ALO, &map_q_lo, full_leader, kc * KC, q0, kEvictLast);
                            tma_load_2d_2sm(st + OFF_BLO, &map_r_lo, full_leader, kc * KC, r0, kEvictNormal);
                        }
                    }
                    __syncwarp();
                    if (++s == NST) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer (leader CTA) ====================
        if (rank == 0) {
            int s = 0;
            uint32_t ph = 0, eph = 0;
            long long gp = 0;
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint32_t smem_u = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
            uint32_t rfph = 0;
            long long p_full = 0, p_empty = 0;
            const long long p_t0 = (MDSCTK_TC_PROF_BUILD && a.prof) ? clock64() : 0;
            const uint32_t res_lo = umma_desc_lo(smem_u), ring_lo = umma_desc_lo(smem_u + (uint32_t)((a.D_pad + RES_KC - 1) / RES_KC * RES_CHUNK));
            for (long long it = pair_id; it < n_items; it += n_pairs) {
                long long qt, rt0, rt1, rot; int seg;
                item_range(it, qt, rt0, rt1, seg, rot);
                if (a.res) { mbar_wait_spin(&bar_res_full, rfph, 6); rfph ^= 1; }      // this item's fit tile has landed
                for (long long ti = 0; ti < rt1 - rt0; ++ti, ++gp) {
                    const int buf = (int)(gp & 1);
                    if (gp >= 2) {
                        const long long t0 = (MDSCTK_TC_PROF_BUILD && a.prof) ? clock64() : 0;
                        mbar_wait_spin(&bar_tmem_empty[buf], (eph >> buf) & 1u, 2);
                        eph ^= 1u << buf;
                        if (MDSCTK_TC_PROF_BUILD && a.prof) p_empty += clock64() - t0;
                    }
                    tc_fence_after();
                    const uint32_t d = tmem_u + buf * UMMA_N;
                    if (a.res) {
                        // resident mode: A from the resident tile, B from the ring; four MMAs (k-steps of 16 dims = 32 B inside the
                        // 128-byte swizzle atom) per stage, descriptors as low words differing by constants
                        const int nstage = (a.D_pad + RES_KC - 1) / RES_KC;
                        for (int j = 0; j < nstage; ++j) {
                            const long long t0 = (MDSCTK_TC_PROF_BUILD && a.prof) ? clock64() : 0;
                            mbar_wait_spin(&bar_full[s], ph, 3);
                            if (MDSCTK_TC_PROF_BUILD && a.prof) p_full += clock64() - t0;
                            if (elect_one()) {
                                const uint32_t a0 = res_lo + (uint32_t)j * (RES_CHUNK >> 4), b0 = ring_lo + (uint32_t)s * (RES_STAGE >> 4);
                                tc_mma2_f16_lo_hi(d, a0, b0, kDescHiSw128, IDESC, j != 0);
                                tc_mma2_f16_lo_hi(d, a0 + 2u, b0 + 2u, kDescHiSw128, IDESC, 1);
                                if (j * RES_KC + 32 < a.D_pad) {                                   // (D_pad is a multiple of 32)
                                    tc_mma2_f16_lo_hi(d, a0 + 4u, b0 + 4u, kDescHiSw128, IDESC, 1);
                                    tc_mma2_f16_lo_hi(d, a0 + 6u, b0 + 6u, kDescHiSw128, IDESC, 1);
                                }
                                tc_commit2_mc(&bar_empty[s], 3);
                                if (j == nstage - 1) {
                                    tc_commit2_mc(&bar_tmem_full[buf], 3);
                                    if (ti == rt1 - rt0 - 1) tc_commit2_mc(&bar_res_empty, 3);      // item done with its fit tile
                                }
                            }
                            __syncwarp();
                            if (++s == a.res_nst) { s = 0; ph ^= 1; }
                        }
                        continue;
                    }
                    for (int kc = 0; kc < nk; ++kc) {
                        mbar_wait_spin(&bar_full[s], ph, 3);
                        tc_fence_after();
                        const uint32_t sa = smem_u + s * STAGE_BYTES;
                        if (elect_one()) {
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks) {
                                const uint32_t koff = ks * 32;
                                const uint64_t ahi = umma_desc_sw64(sa + OFF_AHI + koff), alo = umma_desc_sw64(sa + OFF_ALO + koff);
                                const uint64_t bhi = umma_desc_sw64(sa + OFF_BHI + koff), blo = umma_desc_sw64(sa + OFF_BLO + koff);
                                tc_mma2<true>(d, ahi, bhi, IDESC, (kc | ks) != 0);
                                if (!a.one) {
                                    tc_mma2<true>(d, ahi, blo, IDESC, 1);
                                    tc_mma2<true>(d, alo, bhi, IDESC, 1);
                                }
                            }
                            tc_commit2_mc(&bar_empty[s], 3);
                            if (kc == nk - 1) tc_commit2_mc(&bar_tmem_full[buf], 3);
                        }
                        __syncwarp();
                        if (++s == NST) { s = 0; ph ^= 1; }
                    }
                }
            }
            if (MDSCTK_TC_PROF_BUILD && a.prof && lane == 0) {
                long long *pr = a.prof + pair_id * 8;
                pr[0] = clock64() - p_t0; pr[1] = p_full; pr[2] = p_empty; pr[3] = gp;
            }
        }
    } else {
        // =============================== epilogue ===================================
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int sub = ew >> 2;                      // reference columns [64 sub, 64 sub + 64) of every tile
        const int e_of_quarter = (quarter + 2) & 3;
        const int row_in_tile = quarter * 32 + lane;
        const uint32_t t_warp = tmem_base + ((uint32_t)(quarter * 32) << 16) + sub * SUBW;
        unsigned *hist = s_hist[ew];
        int *wcnt = s_cnt[ew];
        float *rn = s_rn[ew];
        float rn0 = 0.f, rn1 = 0.f;
        const size_t row_stride = (size_t)a.cl.H * a.cl.cap;
        float *lkeys = a.cl.key;
        int *lidxs = a.cl.idx;
        size_t lbase0 = 0;
        uint32_t fph = 0;
        long long gp = 0;
        long long p_wait = 0, p_batch = 0, p_merge = 0;
        const bool p_on = MDSCTK_TC_PROF_BUILD && a.prof && ew == 0 && rank == 0;
        const uint32_t empty_leader0 = map_to_cta(&bar_tmem_empty[0], 0), empty_leader1 = map_to_cta(&bar_tmem_empty[1], 0);

        // need: this thread's private area cannot take another full pass.  A non-final call costs one barrier (which also
        // carries the OR of `need` over the quarter) unless some row of the quarter has to be compacted.
        auto merge_rows = [&](long long row0, int seg, bool final, bool need) {
            if (final) quarter_sync(quarter);
            else if (!quarter_any(quarter, need)) return;
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
                if (!final && cmax <= SUB_APP - SUBW) continue;      // room for another full pass
                const size_t at = lbase0 + (size_t)l * row_stride;
                int total = s_mcnt[row];
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
                float tl = s_tau[row];
                if (total > a.cl.keep) {
                    tl = fminf(tl, warp_compact_list6(lkeys + at, lidxs + at, total, a.cl.keep, hist));
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
                        atomicMin(reinterpret_cast<unsigned *>(a.row_tau + qr), __float_as_uint(fmaxf(tl, 0.0f)));
                    }
                }
            }
            quarter_sync(quarter);
        };

        for (long long it = pair_id; it < n_items; it += n_pairs) {
            long long qt, rt0, rt1, rot; int seg;
            item_range(it, qt, rt0, rt1, seg, rot);
            const long long row0 = qt * UMMA_M + rank * TQ;
            const long long qrow = row0 + row_in_tile;
            const bool qvalid = qrow < a.n_q;
            const float nq = a.q_norm[qvalid ? qrow : a.n_q - 1];
            lbase0 = ((size_t)(row0 + quarter * 32) * a.cl.H + seg) * a.cl.cap;
            wcnt[lane] = 0;
            if (sub == 0) {
                s_mcnt[row_in_tile] = 0;
                s_tau[row_in_tile] = qvalid ? __ldcg(a.row_tau + qrow) : -1.0f;
            }
            quarter_sync(quarter);
            {
                const float *nx = a.r_norm + tile_at(0, rt0, rt1, rot) * TR + sub * SUBW;
                rn0 = __ldg(nx + lane); rn1 = __ldg(nx + 32 + lane);
            }
            float *my_keys = lkeys + lbase0 + (size_t)lane * row_stride + a.cl.keep + sub * SUB_APP;
            int *my_idxs = lidxs + lbase0 + (size_t)lane * row_stride + a.cl.keep + sub * SUB_APP;
            for (long long ti = 0; ti < rt1 - rt0; ++ti, ++gp) {
                const int buf = (int)(gp & 1);
                const long long rt = tile_at(ti, rt0, rt1, rot);
                const long long rb = rt * TR + sub * SUBW;
                const long long t0 = p_on ? clock64() : 0;
                // this pass's reference norms (prefetched into registers a pass ago) -> the warp's shared-memory row; the
                // batches below read them as broadcast float4 (a global load per batch sat on the pass's critical path)
                rn[lane] = rn0; rn[lane + 32] = rn1;
                if (ti + 1 < rt1 - rt0) {
                    const float *nx = a.r_norm + tile_at(ti + 1, rt0, rt1, rot) * TR + sub * SUBW;
                    rn0 = __ldg(nx + lane); rn1 = __ldg(nx + 32 + lane);
                }
                if (lane == 0) mbar_wait(&bar_tmem_full[buf], (fph >> buf) & 1u, 4);
                fph ^= 1u << buf;
                __syncwarp();
                tc_fence_after();
                const long long t1 = p_on ? clock64() : 0;
                // accumulator units: s^2 * input units^2
                const float tau_s = s_tau[row_in_tile] / a.inv_scale2;
                int cnt = wcnt[lane];
                // Nearly every batch holds no candidate once the row's threshold is tight: one fused multiply-add and one min
                // per key decide that, the per-key tests and the append run only for a batch whose smallest key passes.  Rows
                // beyond n_q carry tau = -1 and never pass; reference columns beyond n_r are sorted out on the slow path.
                // The next batch's TMEM load is in flight while this one is tested.
                auto test_batch = [&](const float (&dv)[EB], int h) {
                    float kmin = __uint_as_float(0x7f800000u);
#pragma unroll
                    for (int j = 0; j < EB / 4; ++j) {
                        const float4 g = reinterpret_cast<const float4 *>(rn + h)[j];
                        kmin = fminf(fminf(kmin, fmaf(-2.0f, dv[4 * j], g.x)), fmaf(-2.0f, dv[4 * j + 1], g.y));
                        kmin = fminf(fminf(kmin, fmaf(-2.0f, dv[4 * j + 2], g.z)), fmaf(-2.0f, dv[4 * j + 3], g.w));
                    }
                    if (!(kmin + nq > tau_s * 1.000001f)) {      // (the two evaluation orders differ by an ulp: err on the side of the slow path)
#pragma unroll
                        for (int j = 0; j < EB; ++j) {
                            const float key = fmaf(-2.0f, dv[j], nq + rn[h + j]);
                            if (key < tau_s && qvalid && rb + h + j < a.n_r) {
                                if (cnt < SUB_APP) {
                                    my_keys[cnt] = fmaxf(key, 0.0f) * a.inv_scale2;
                                    my_idxs[cnt] = (int)(rb + h + j);
                                }
                                ++cnt;
                            }
                        }
                    }
                };
                float dva[EB], dvb[EB];
                const uint32_t t_pass = t_warp + buf * UMMA_N;
                tc_ld16(t_pass, dva);
                tc_wait_ld16(dva);
                tc_ld16(t_pass + EB, dvb);
                test_batch(dva, 0);
                tc_wait_ld16(dvb);
                tc_ld16(t_pass + 2 * EB, dva);
                test_batch(dvb, EB);
                tc_wait_ld16(dva);
                tc_ld16(t_pass + 3 * EB, dvb);
                test_batch(dva, 2 * EB);
                tc_wait_ld16(dvb);
                // the accumulator buffer goes back to the issuer before the last batch is tested
                // (no cluster-scope release fence: the TMEM reads are ordered by tcgen05.wait::ld + fence::before_thread_sync)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_nofence(buf ? empty_leader1 : empty_leader0);
                test_batch(dvb, 3 * EB);
                wcnt[lane] = cnt;
                const long long t2 = p_on ? clock64() : 0;
                merge_rows(row0, seg, false, cnt > SUB_APP - SUBW);
                if (p_on) { const long long t3 = clock64(); p_wait += t1 - t0; p_batch += t2 - t1; p_merge += t3 - t2; }
            }
            merge_rows(row0, seg, true, true);
        }
        if (p_on && lane == 0) {
            long long *pr = a.prof + pair_id * 8;
            pr[4] = p_wait; pr[5] = p_batch; pr[6] = p_merge;
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

// rows[n][D_pad] fp16 as a 2-D tensor (dim, row); box = 32 dims (64 bytes) x `rows` rows
static bool make_row_map(CUtensorMap *m, const void *base, long long n, int D_pad, int rows, bool wide = false)
{
    EncodeTiledFn enc = get_tensor_map_encoder();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)D_pad, (cuuint64_t)n};
    cuuint64_t strides[1] = {(cuuint64_t)D_pad * 2};
    cuuint32_t box[2] = {(cuuint32_t)(wide ? dtc::RES_KC : dtc::KC), (cuuint32_t)rows};      // wide: 128-byte rows for the resident mode
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, wide ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool g_data_force_streaming = false;
void data_tc_set_streaming(bool on) { g_data_force_streaming = on; }      // test hook: the round-1 streaming mode of the one-part filter
static bool data_tc_force_streaming() { return g_data_force_streaming; }
int data_tc_pad_dim(int dim) { return (dim + dtc::KC - 1) / dtc::KC * dtc::KC; }
int data_tc_list_stride(int keep) { return (keep + dtc::SUBS * dtc::SUB_APP + 31) / 32 * 32; }

int data_tc_choose_segments(long long n_fit, long long n_ref, int n_sms)
{
    const int n_pairs = n_sms / 2 > 0 ? n_sms / 2 : 1;
    const long long n_qt = (n_fit + dtc::UMMA_M - 1) / dtc::UMMA_M;
    const long long n_rt = (n_ref + dtc::TR - 1) / dtc::TR;
    int best = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= 8; ++s) {
        if (s > 1 && n_rt / s < 8) break;
        const long long items = n_qt * s;
        const double eff = (double)items / (double)(((items + n_pairs - 1) / n_pairs) * n_pairs);
        if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
    }
    return best;
}

cudaError_t launch_data_sweep_tc(const void *fit_hi, const void *fit_lo, const float *fit_norm, long long n_fit,
                                 long long fit_begin_in_ref, const void *ref_hi, const void *ref_lo, const float *ref_norm,
                                 long long n_ref, int D_pad, double scale, int n_seg, CandLists<float> cl, float *row_tau,
                                 int one_part, int n_sms, cudaStream_t st)
{
    if (n_fit <= 0) return cudaSuccess;
    if (cl.H != n_seg || cl.cap < cl.keep + dtc::SUBS * dtc::SUB_APP || D_pad % dtc::KC) return cudaErrorInvalidValue;
    CUtensorMap mq_hi, mq_lo, mr_hi, mr_lo;
    if (!make_row_map(&mq_hi, fit_hi, n_fit, D_pad, dtc::TQ) || !make_row_map(&mq_lo, fit_lo, n_fit, D_pad, dtc::TQ) ||
        !make_row_map(&mr_hi, ref_hi, n_ref, D_pad, dtc::TRH) || !make_row_map(&mr_lo, ref_lo, n_ref, D_pad, dtc::TRH))
        return cudaErrorInvalidValue;
    DataTcArgs a;
    a.q_norm = fit_norm; a.r_norm = ref_norm; a.q_begin = fit_begin_in_ref; a.n_q = n_fit; a.n_r = n_ref;
    a.D_pad = D_pad; a.n_seg = n_seg; a.inv_scale2 = (float)(1.0 / (scale * scale)); a.cl = cl; a.row_tau = row_tau;
    a.one = one_part;
    a.prof = nullptr;
#if MDSCTK_TC_PROF_BUILD
    static long long *d_prof = nullptr;
    const bool prof = getenv("MDSCTK_TC_PROF") && atoi(getenv("MDSCTK_TC_PROF")) != 0;
    if (prof) {
        if (!d_prof && cudaMalloc(&d_prof, 128 * 8 * sizeof(long long)) != cudaSuccess) return cudaErrorMemoryAllocation;
        cudaMemsetAsync(d_prof, 0, 128 * 8 * sizeof(long long), st);
        a.prof = d_prof;
    }
#endif
    // resident fit tile (one-part filter): when 128 x D_pad fp16 fit rows leave room for >= 3 ring stages of reference rows
    const int nk = D_pad / dtc::KC;
    const int res_chunks = (D_pad + dtc::RES_KC - 1) / dtc::RES_KC;
    int res_nst = (227 * 1024 - dtc::STATIC_SMEM - 1024 - res_chunks * dtc::RES_CHUNK) / dtc::RES_STAGE;
    if (res_nst > dtc::MAX_NST) res_nst = dtc::MAX_NST;
    a.res = one_part && res_nst >= 3 && !data_tc_force_streaming();
    a.res_nst = res_nst;
    const int smem_bytes = a.res ? res_chunks * dtc::RES_CHUNK + res_nst * dtc::RES_STAGE + 1024 : dtc::SMEM_BYTES;
    if (a.res && (!make_row_map(&mq_hi, fit_hi, n_fit, D_pad, dtc::TQ, true) || !make_row_map(&mr_hi, ref_hi, n_ref, D_pad, dtc::TRH, true)))
        return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(data_sweep_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, data_sweep_tc_kernel);
    if (e == cudaSuccess && fa.sharedSizeBytes > (size_t)dtc::STATIC_SMEM) e = cudaErrorInvalidConfiguration;
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.blockDim = dim3(dtc::NTHR); cfg.dynamicSmemBytes = (size_t)smem_bytes; cfg.stream = st;
    int max_pairs = n_sms / 2;
    cfg.gridDim = dim3((unsigned)(max_pairs * 2));
    int q = 0;
    if (cudaOccupancyMaxActiveClusters(&q, data_sweep_tc_kernel, &cfg) == cudaSuccess && q > 0 && q < max_pairs) max_pairs = q;
    (void)cudaGetLastError();
    const long long n_items = ((n_fit + dtc::UMMA_M - 1) / dtc::UMMA_M) * n_seg;
    const long long n_pairs = n_items < max_pairs ? n_items : max_pairs;
    cfg.gridDim = dim3((unsigned)(n_pairs * 2));
    e = cudaLaunchKernelEx(&cfg, data_sweep_tc_kernel, mq_hi, mq_lo, mr_hi, mr_lo, a);
    if (e != cudaSuccess) return e;
#if MDSCTK_TC_PROF_BUILD
    if (a.prof) {
        static long long h[128 * 8];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, a.prof, sizeof(h), cudaMemcpyDeviceToHost);
        double t[7] = {0, 0, 0, 0, 0, 0, 0};
        for (long long p2 = 0; p2 < n_pairs; ++p2) for (int j = 0; j < 7; ++j) t[j] += (double)h[p2 * 8 + j];
        if (t[3] > 0)
            fprintf(stderr, "[data prof] pairs=%lld passes %.0f res=%d nst=%d | issuer clk per pass: total %.0f = wait operands %.0f + wait hand-back %.0f + issue %.0f"
                            " | epilogue warp 0: wait accumulators %.0f, batches %.0f, merge %.0f\n",
                    n_pairs, t[3], a.res, a.res_nst, t[0] / t[3], t[1] / t[3], t[2] / t[3], (t[0] - t[1] - t[2]) / t[3], t[4] / t[3], t[5] / t[3], t[6] / t[3]);
    }
#endif
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------- re-score ----
// One block per fit row: merge the row's lists, order by approximate key, recompute the candidates'
// squared distances in FP64 exactly as euclidean_distance does (mdsctk.cpp:330-335: d = a[x]-b[x];
// sum += d*d, left to right, no FMA), 64 candidates per round, then (distance, index) order and the
// certificate  max approx key of the exact top-k1 + 2 eps < smallest approx key not yet re-scored.
// One-part filter (q_g != NULL): its key is the exact squared distance of the fp16-ROUNDED rows plus bias and noise,
// and |d~ - d| <= g = g_q + max g_r (triangle inequality), so the rule becomes, exactly as in rms_rescore.cu,
//   min_i [key_i - max(d_i - g, 0)^2] + (d_k + g)^2 + 2 eps < a_next,  with the brackets of all candidates intersecting.
struct DataRescoreArgs {
    const double *fit, *ref;          // fit rows of this query (row 0 = fit row 0), reference rows
    long long n_fit;
    int dim, k1, P, round0;           // round0: candidates of the first round = min(128, roundup32(k1))
    CandLists<float> cl;
    double eps_rel;                   // noise bound of the filter as a fraction of (|x|^2 + |y|^2)
    const float *q_norm;              // scaled norms; inv_scale2 converts
    float inv_scale2, r_norm_max;
    const float *q_g;                 // one-part filter: residual norm |x - hi/s| of every fit row (input units), else NULL
    float g_ref_max;                  // ... and the largest one of the reference set
    const double *fit_stats, *ref_stats;   // correlation metric: (mean, spread) per row, else NULL (Euclidean)
    double *out_dist;
    int *out_idx, *flags, *n_bad, *bad_rows;
    double *err_stats;
};

constexpr int DRESCORE_ROUND = 128;   // upper bound of a round (one candidate per thread)
constexpr int DCHUNK = 16;            // dims staged per step: [round][17] doubles (13 KB for k = 64)

__global__ void __launch_bounds__(128) data_rescore_kernel(DataRescoreArgs a)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    const int P = a.P, D = a.dim;
    double *fq = reinterpret_cast<double *>(dsm);           // [D] fit row
    double *s_d = fq + D;                                    // [P]
    double *u_key = s_d + P;                                 // [P] exact key (squared distance), candidate order
    int *s_i = reinterpret_cast<int *>(u_key + P);           // [P]
    int *u_i = s_i + P;                                      // [P]
    float *u_apx = reinterpret_cast<float *>(u_i + P);       // [P]
    double *s_chunk = reinterpret_cast<double *>(dsm + (((size_t)D * 8 + (size_t)P * 28 + 15) & ~(size_t)15));   // [128][DCHUNK + 1]
    __shared__ int s_total, s_ok;
    __shared__ float s_taumin;
    __shared__ unsigned s_dtil;
    __shared__ unsigned long long s_emin, s_emax;

    const long long q = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int x = threadIdx.x; x < D; x += blockDim.x) fq[x] = a.fit[(size_t)q * D + x];
    const double kInfD = __longlong_as_double(0x7ff0000000000000LL);
    const float kInfF = __uint_as_float(0x7f800000u);
    const int k1 = a.k1;
    if (threadIdx.x == 0) { s_total = 0; s_taumin = kInfF; s_ok = 0; s_emin = ~0ull; s_emax = 0ull; }
    for (int i = threadIdx.x; i < P; i += blockDim.x) { s_d[i] = kInfD; s_i[i] = 0x7fffffff; }
    __syncthreads();
    for (int h = 0; h < a.cl.H; ++h) {
        const size_t lid = (size_t)q * a.cl.H + h;
        const int c = min(a.cl.cnt[lid], a.cl.keep);
        if (threadIdx.x == 0) { s_total += c; s_taumin = fminf(s_taumin, a.cl.tau[lid]); }
        __syncthreads();
        const int base = s_total - c;
        for (int i = threadIdx.x; i < c; i += blockDim.x) {
            s_d[base + i] = (double)a.cl.key[lid * a.cl.cap + i];
            s_i[base + i] = a.cl.idx[lid * a.cl.cap + i];
        }
        __syncthreads();
    }
    const int total = s_total;
    int Pr = 32;
    while (Pr < total) Pr <<= 1;
    const float tau_row = s_taumin;
    block_bitonic_sort(s_d, s_i, Pr);
    for (int i = threadIdx.x; i < Pr; i += blockDim.x) {
        u_apx[i] = i < total ? (float)s_d[i] : kInfF;
        u_i[i] = i < total ? s_i[i] : 0x7fffffff;
        u_key[i] = kInfD;
    }
    __syncthreads();

    const double nq = (double)a.q_norm[q] * (double)a.inv_scale2;
    const double grd = a.q_g ? (double)a.q_g[q] + (double)a.g_ref_max : 0.0;
    // correlation metric: exact key = 4 * correlation_distance^2 (the squared distance of the standardised rows the filter
    // contracted; the factor 4 is exact), computed in the reference's own order: sum (reference[x] - rmean) (fitting[x] - fmean)
    // left to right, then (1 - sum / ((n - 1) rs fs)) / 2 clamped at 0 (mdsctk.cpp:353-358; "reference" = the fit row)
    const bool corr = a.fit_stats != nullptr;
    const double fm = corr ? a.fit_stats[2 * q] : 0.0, fs = corr ? a.fit_stats[2 * q + 1] : 1.0;
    const double eps_max = a.eps_rel * (nq + (double)a.r_norm_max * (double)a.inv_scale2);
    double eps = eps_max;
    int done = 0;
    bool certified = false;
    while (done < total) {
        // first round: just enough candidates to fill the k-list; then 32 more at a time
        const int nb = min(done == 0 ? a.round0 : 32, total - done);
        // candidate rows are staged through shared memory in chunks of DCHUNK dims with coalesced loads
        // (a warp per candidate); thread c then adds its candidate's terms in the reference's order
        double sum = 0.0;
        for (int x0 = 0; x0 < D; x0 += DCHUNK) {
            const int w = min(DCHUNK, D - x0);
            for (int c = warp; c < nb; c += 4) {
                const double *rv = a.ref + (size_t)u_i[done + c] * D + x0;
                for (int x = lane; x < w; x += 32) s_chunk[c * (DCHUNK + 1) + x] = rv[x];
            }
            __syncthreads();
            if (threadIdx.x < nb) {
                const double *rc = s_chunk + threadIdx.x * (DCHUNK + 1);
                if (!corr) {
                    for (int x = 0; x < w; ++x) {
                        const double d = __dsub_rn(fq[x0 + x], rc[x]);
                        sum = __dadd_rn(sum, __dmul_rn(d, d));
                    }
                } else {
                    const double rm = a.ref_stats[2 * (size_t)u_i[done + threadIdx.x]];
                    for (int x = 0; x < w; ++x)
                        sum = __dadd_rn(sum, __dmul_rn(__dsub_rn(fq[x0 + x], fm), __dsub_rn(rc[x], rm)));
                }
            }
            __syncthreads();
        }
        if (corr && threadIdx.x < nb) {
            const double rs = a.ref_stats[2 * (size_t)u_i[done + threadIdx.x] + 1];
            const double den = __dmul_rn(__dmul_rn(__dsub_rn((double)D, 1.0), fs), rs);
            double v = __ddiv_rn(__dsub_rn(1.0, __ddiv_rn(sum, den)), 2.0);
            if (v < 0.0) v = 0.0;
            sum = 4.0 * v;
        }
        if (threadIdx.x < nb) {
            u_key[done + threadIdx.x] = sum;
            // bracket of the row-common bias: [key - (d+g)^2, key - max(d-g,0)^2]; a point (approx - exact) when g = 0
            const double dn = sqrt(sum), dm = fmax(dn - grd, 0.0);
            const double err_hi = (double)u_apx[done + threadIdx.x] - (grd > 0.0 ? dm * dm : sum);
            const double err_lo = (double)u_apx[done + threadIdx.x] - (grd > 0.0 ? (dn + grd) * (dn + grd) : sum);
            auto enc = [](double v) {
                unsigned long long bits = (unsigned long long)__double_as_longlong(v);
                return (bits >> 63) ? ~bits : (bits | 0x8000000000000000ull);
            };
            atomicMin(&s_emin, enc(err_hi));
            atomicMax(&s_emax, enc(err_lo));
        }
        __syncthreads();
        done += nb;
        int Ps = 32;                       // only the re-scored prefix carries finite keys
        while (Ps < done) Ps <<= 1;
        for (int i = threadIdx.x; i < Ps; i += blockDim.x) { s_d[i] = u_key[i]; s_i[i] = u_i[i]; }
        if (threadIdx.x == 0) s_dtil = 0u;
        __syncthreads();
        block_bitonic_sort(s_d, s_i, Ps);
        if (done >= k1) {
            const double dk = s_d[k1 - 1];
            const int ik = s_i[k1 - 1];
            float dt = 0.0f;
            for (int i = threadIdx.x; i < done; i += blockDim.x)
                if (pair_less(u_key[i], u_i[i], dk, ik) || (u_key[i] == dk && u_i[i] == ik)) dt = fmaxf(dt, u_apx[i]);
            atomicMax(&s_dtil, __float_as_uint(fmaxf(dt, 0.0f)));
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            auto dec = [](unsigned long long b) {
                b = (b >> 63) ? (b & 0x7fffffffffffffffull) : ~b;
                return __longlong_as_double((long long)b);
            };
            const double spread = dec(s_emax) - dec(s_emin);
            const float a_next = done < total ? fminf(u_apx[done], tau_row) : tau_row;
            bool ok = done >= k1;
            if (ok) {   // |y| <= |x| + d for every pair that could still matter
                const double reach = sqrt(nq) + sqrt(s_d[k1 - 1] + 4.0 * eps_max);
                eps = a.eps_rel * (nq + fmin((double)a.r_norm_max * (double)a.inv_scale2, reach * reach));
            }
            if (ok && a_next != kInfF) {
                if (a.q_g) {
                    const double dkg = sqrt(s_d[k1 - 1]) + grd;
                    ok = 0.5 * spread <= eps_max && dec(s_emin) + dkg * dkg + 2.0 * eps < (double)a_next;
                } else {
                    ok = 0.5 * spread <= eps_max && (double)__uint_as_float(s_dtil) + 2.0 * eps < (double)a_next;
                }
            }
            s_ok = ok ? 1 : 0;
        }
        __syncthreads();
        certified = s_ok != 0;
        if (certified) break;
    }
    if (total == 0 && threadIdx.x == 0) s_ok = 0;
    // the CPU tool takes sqrt of every entry before it sorts; sqrt is monotone, ties keep index order
    for (int j = threadIdx.x; j < k1; j += blockDim.x) {
        a.out_dist[(size_t)q * k1 + j] = sqrt(corr ? 0.25 * s_d[j] : s_d[j]);
        a.out_idx[(size_t)q * k1 + j] = s_i[j];
    }
    if (threadIdx.x == 0) {
        auto dec = [](unsigned long long b) {
            b = (b >> 63) ? (b & 0x7fffffffffffffffull) : ~b;
            return __longlong_as_double((long long)b);
        };
        if (total > 0) {
            const double lo = dec(s_emin), hi = dec(s_emax);
            atomic_max_nonneg(a.err_stats + 0, fmax(fabs(lo), fabs(hi)));
            atomic_max_nonneg(a.err_stats + 1, fmax(hi - lo, 0.0));
            atomic_max_nonneg(a.err_stats + 2, (double)done);
        }
        a.flags[q] = certified ? 1 : 0;
        if (!certified) {
            const int pos = atomicAdd(a.n_bad, 1);
            a.bad_rows[pos] = (int)q;
        }
    }
}

cudaError_t launch_data_rescore(const double *fit, const double *ref, long long n_fit, int dim, int k1, CandLists<float> cl,
                                double eps_rel, const float *q_norm, double scale, float r_norm_max, const float *q_g,
                                float g_ref_max, const double *fit_stats, const double *ref_stats, double *out_dist, int *out_idx,
                                int *flags, double *err_stats, int *n_bad, int *bad_rows, cudaStream_t st)
{
    if (n_fit <= 0) return cudaSuccess;
    DataRescoreArgs a;
    a.fit_stats = fit_stats; a.ref_stats = ref_stats;
    a.fit = fit; a.ref = ref; a.n_fit = n_fit; a.dim = dim; a.k1 = k1; a.cl = cl; a.eps_rel = eps_rel; a.q_norm = q_norm;
    a.inv_scale2 = (float)(1.0 / (scale * scale)); a.r_norm_max = r_norm_max; a.out_dist = out_dist; a.out_idx = out_idx;
    a.q_g = q_g; a.g_ref_max = g_ref_max;
    a.flags = flags; a.err_stats = err_stats; a.n_bad = n_bad; a.bad_rows = bad_rows;
    int P = 32;
    while (P < cl.keep * cl.H) P <<= 1;
    a.P = P;
    a.round0 = std::min(DRESCORE_ROUND, (k1 + 31) / 32 * 32);
    const size_t smem = (((size_t)dim * 8 + (size_t)P * 28 + 15) & ~(size_t)15) + (size_t)std::max(a.round0, 32) * (DCHUNK + 1) * 8;
    if (smem > 220 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(data_rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    data_rescore_kernel<<<(unsigned)n_fit, 128, smem, st>>>(a);
    return cudaGetLastError();
}

// ------------------------------------------------- helpers for the exact fallback of bad rows ----
__global__ void data_gather_rows_kernel(const double *__restrict__ src, const int *__restrict__ ids, int n_rows, int dim,
                                        double *__restrict__ dst)
{
    const int r = blockIdx.x;
    if (r >= n_rows) return;
    const double *s = src + (size_t)ids[r] * dim;
    for (int x = threadIdx.x; x < dim; x += blockDim.x) dst[(size_t)r * dim + x] = s[x];
}

__global__ void data_scatter_out_kernel(const double *__restrict__ d_src, const int *__restrict__ i_src,
                                        const int *__restrict__ ids, int n_rows, int k1, double *__restrict__ d_dst,
                                        int *__restrict__ i_dst)
{
    const int r = blockIdx.x;
    if (r >= n_rows) return;
    for (int j = threadIdx.x; j < k1; j += blockDim.x) {
        d_dst[(size_t)ids[r] * k1 + j] = d_src[(size_t)r * k1 + j];
        i_dst[(size_t)ids[r] * k1 + j] = i_src[(size_t)r * k1 + j];
    }
}

cudaError_t launch_data_gather_rows(const double *src, const int *ids, int n_rows, int dim, double *dst, cudaStream_t st)
{
    if (n_rows <= 0) return cudaSuccess;
    data_gather_rows_kernel<<<n_rows, 128, 0, st>>>(src, ids, n_rows, dim, dst);
    return cudaGetLastError();
}

cudaError_t launch_data_scatter_out(const double *d_src, const int *i_src, const int *ids, int n_rows, int k1, double *d_dst,
                                    int *i_dst, cudaStream_t st)
{
    if (n_rows <= 0) return cudaSuccess;
    data_scatter_out_kernel<<<n_rows, 64, 0, st>>>(d_src, i_src, ids, n_rows, k1, d_dst, i_dst);
    return cudaGetLastError();
}

}  // namespace mdsctk
