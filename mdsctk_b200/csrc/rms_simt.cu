// rms_simt.cu -- all-pairs superposed-RMSD sweep, FP32 CUDA-core contraction.
//
// First correct GPU path for knn_rms.cpp:231-293 (the OpenMP row blocks): every CTA owns
// 64 fit frames for the whole reference sweep, streams 32-frame reference tiles through a
// 3-stage cp.async pipeline, accumulates the nine S_ab = sum_n x_na y_nb per pair in FP32
// registers (2 fit x 4 reference frames per thread), solves QCP in the epilogue (qcp.cuh)
// and feeds the streaming top-k (select.cuh).  The N x N matrix never exists.
//
// This kernel is the precision-safe baseline and cross-check for the tcgen05 kernel
// (rms_tc.cu); both read the same `planes` layout.
#include "common.cuh"
#include "qcp.cuh"
#include "select.cuh"

namespace mdsctk {

namespace simt {
constexpr int TQ = 64;       // fit frames per CTA
constexpr int TR = 32;       // reference frames per tile
constexpr int KC = 16;       // atoms per pipeline stage
constexpr int ROWP = 20;     // smem row pitch in floats (16 + 4): conflict-free LDS.128
constexpr int NST = 3;       // pipeline stages
constexpr int NTHR = 256;
constexpr int ROWS = (TQ + TR) * 3;             // frame-plane rows per stage
constexpr int STAGE_FLOATS = ROWS * ROWP;
constexpr int SMEM_BYTES = NST * STAGE_FLOATS * 4;
}  // namespace simt

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

struct SimtArgs {
    const float *q_planes, *r_planes, *q_G, *r_G;
    long long q_begin, n_q, n_r;
    int A_pad, do_fit;
    CandLists<float> cl;
};

__global__ void __launch_bounds__(simt::NTHR, 2) rms_sweep_simt_kernel(SimtArgs a)
{
    using namespace simt;
    extern __shared__ __align__(16) float smem[];
    __shared__ int s_cnt[TQ];
    __shared__ float s_tau[TQ];
    __shared__ float s_gr[TR];
    __shared__ unsigned s_hist[NTHR / 32][256];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tq = lane;   // fit frames tq and tq+32 of the CTA tile
    const int tr = warp;   // reference frames 4*tr .. 4*tr+3 of the tile
    const long long q0 = (long long)blockIdx.x * TQ;  // row within the query range
    const int nk = a.A_pad / KC;
    const long long n_rt = (a.n_r + TR - 1) / TR;
    const long long total = n_rt * nk;
    const size_t plane_pitch = (size_t)a.A_pad;

    if (tid < TQ) { s_cnt[tid] = 0; s_tau[tid] = KeyBits<float>::inf(); }

    // cp.async producer: one stage = KC atoms of (TQ fit + TR reference) frames x 3 planes
    auto issue = [&](long long rt, int kc, int stage) {
        float *dst = smem + stage * STAGE_FLOATS;
        for (int i = tid; i < ROWS * 4; i += NTHR) {
            const int row = i >> 2, seg = i & 3;
            const int frame = row / 3, plane = row - frame * 3;
            const float *src;
            if (frame < TQ) {
                long long f = q0 + frame;
                if (f >= a.n_q) f = a.n_q - 1;
                src = a.q_planes + ((size_t)(a.q_begin + f) * 3 + plane) * plane_pitch;
            } else {
                long long f = rt * TR + (frame - TQ);
                if (f >= a.n_r) f = a.n_r - 1;
                src = a.r_planes + ((size_t)f * 3 + plane) * plane_pitch;
            }
            cp_async16(dst + row * ROWP + seg * 4, src + kc * KC + seg * 4);
        }
    };

    long long irt = 0;
    int ikc = 0;
    for (int s = 0; s < NST - 1; ++s) {
        if ((long long)s < total) {
            issue(irt, ikc, s);
            if (++ikc == nk) { ikc = 0; ++irt; }
        }
        cp_async_commit();
    }

    float qg[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        long long f = q0 + tq + 32 * i;
        qg[i] = a.q_G[a.q_begin + (f < a.n_q ? f : a.n_q - 1)];
    }

    long long t = 0;
    for (long long rt = 0; rt < n_rt; ++rt) {
        const long long r0 = rt * TR;
        if (tid < TR) s_gr[tid] = a.r_G[(r0 + tid < a.n_r) ? r0 + tid : a.n_r - 1];

        float acc[2][4][9];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < 9; ++c) acc[i][j][c] = 0.0f;

        for (int kc = 0; kc < nk; ++kc, ++t) {
            cp_async_wait<NST - 2>();
            __syncthreads();
            if (t + NST - 1 < total) {
                issue(irt, ikc, (int)((t + NST - 1) % NST));
                if (++ikc == nk) { ikc = 0; ++irt; }
            }
            cp_async_commit();

            const float *st = smem + (int)(t % NST) * STAGE_FLOATS;
            const float *sq = st;
            const float *sr = st + TQ * 3 * ROWP;
#pragma unroll
            for (int k4 = 0; k4 < KC / 4; ++k4) {
                float4 qv[2][3];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int p = 0; p < 3; ++p)
                        qv[i][p] = *reinterpret_cast<const float4 *>(sq + ((tq + 32 * i) * 3 + p) * ROWP + k4 * 4);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float4 rv[3];
#pragma unroll
                    for (int p = 0; p < 3; ++p)
                        rv[p] = *reinterpret_cast<const float4 *>(sr + ((tr * 4 + j) * 3 + p) * ROWP + k4 * 4);
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int pa = 0; pa < 3; ++pa)
#pragma unroll
                            for (int pb = 0; pb < 3; ++pb) {
                                float v = acc[i][j][pa * 3 + pb];
                                v = fmaf(qv[i][pa].x, rv[pb].x, v);
                                v = fmaf(qv[i][pa].y, rv[pb].y, v);
                                v = fmaf(qv[i][pa].z, rv[pb].z, v);
                                v = fmaf(qv[i][pa].w, rv[pb].w, v);
                                acc[i][j][pa * 3 + pb] = v;
                            }
                }
            }
        }

        // ---- epilogue: QCP + threshold-gated append -----------------------------------
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int ql = tq + 32 * i;
            const long long qrow = q0 + ql;
            const float tau = s_tau[ql];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int rl = tr * 4 + j;
                const long long ridx = r0 + rl;
                const float e0 = 0.5f * (qg[i] + s_gr[rl]);
                const float d2 = qcp_msd_bounded(acc[i][j], e0, tau, a.do_fit);
                if (d2 < tau && qrow < a.n_q && ridx < a.n_r) {
                    const int pos = atomicAdd(&s_cnt[ql], 1);
                    if (pos < a.cl.cap) {
                        a.cl.key[(size_t)qrow * a.cl.cap + pos] = d2;
                        a.cl.idx[(size_t)qrow * a.cl.cap + pos] = (int)ridx;
                    }
                }
            }
        }
        __syncthreads();
        // ---- compaction of lists that could overflow during the next tile ------------
        for (int ql = warp; ql < TQ; ql += NTHR / 32) {
            const int c = min(s_cnt[ql], a.cl.cap);
            if (c > a.cl.cap - TR && q0 + ql < a.n_q) {
                const size_t base = (size_t)(q0 + ql) * a.cl.cap;
                float nt = warp_compact_list<float>(a.cl.key + base, a.cl.idx + base, c, a.cl.keep, s_hist[warp]);
                if (lane == 0) { s_tau[ql] = nt; s_cnt[ql] = a.cl.keep; }
            }
        }
        // the __syncthreads at the top of the next chunk orders s_tau / s_cnt / s_gr
    }
    cp_async_wait<0>();
    __syncthreads();
    // ---- final compaction: leave at most `keep` candidates per row ---------------------
    for (int ql = warp; ql < TQ; ql += NTHR / 32) {
        if (q0 + ql >= a.n_q) continue;
        int c = min(s_cnt[ql], a.cl.cap);
        float tau = s_tau[ql];
        const size_t base = (size_t)(q0 + ql) * a.cl.cap;
        if (c > a.cl.keep) {
            tau = warp_compact_list<float>(a.cl.key + base, a.cl.idx + base, c, a.cl.keep, s_hist[warp]);
            c = a.cl.keep;
        }
        if (lane == 0) {
            a.cl.cnt[q0 + ql] = c;
            a.cl.tau[q0 + ql] = tau;
        }
    }
}

cudaError_t launch_rms_sweep_simt(const FrameSetView &fit, long long fit_begin, long long n_fit,
                                  const FrameSetView &ref, int do_fit, CandLists<float> cl, cudaStream_t st)
{
    if (n_fit <= 0) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(rms_sweep_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         simt::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    SimtArgs a;
    a.q_planes = fit.planes; a.r_planes = ref.planes; a.q_G = fit.G; a.r_G = ref.G;
    a.q_begin = fit_begin; a.n_q = n_fit; a.n_r = ref.n;
    a.A_pad = ref.A_pad; a.do_fit = do_fit; a.cl = cl;
    const unsigned grid = (unsigned)((n_fit + simt::TQ - 1) / simt::TQ);
    rms_sweep_simt_kernel<<<grid, simt::NTHR, simt::SMEM_BYTES, st>>>(a);
    return cudaGetLastError();
}

}  // namespace mdsctk
