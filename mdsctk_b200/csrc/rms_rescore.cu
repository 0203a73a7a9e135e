// rms_rescore.cu -- FP64 decision stage of the RMSD path.
//
// The sweep (rms_simt.cu / rms_tc.cu) only FILTERS: it leaves k1+slack candidates per fit
// row with approximate d^2.  This file makes the result exact:
//   rms_rescore_kernel   exact FP64 min-RMSD of every candidate from the RAW frames
//                        (double centring, double cross-covariance, FP64 QCP), final
//                        (distance, index) sort, distances in Angstrom as
//                        distance() returns them (knn_rms.cpp:40: rmsdev * 10.0), and a
//                        per-row certificate that no non-candidate can belong to the top k1.
//   rms_exact_rows_kernel / select_rows_f64_kernel
//                        exact FP64 full rows + exact selection for rows whose certificate
//                        failed (and the diagnostic mdsctk_knn_rms_rows entry point).
#include "common.cuh"

#include <algorithm>
#include <cuda_fp16.h>
#include "sort.cuh"

namespace mdsctk {

// ---- FP64 QCP: largest root of l^4 + c2 l^2 + c1 l + c0 from l0 = e0 ----------------------
__device__ inline double qcp_lambda_f64(const double *s, double e0)
{
    const double sxx = s[0], sxy = s[1], sxz = s[2], syx = s[3], syy = s[4], syz = s[5], szx = s[6], szy = s[7],
                 szz = s[8];
    const double c2 = -2.0 * (sxx * sxx + sxy * sxy + sxz * sxz + syx * syx + syy * syy + syz * syz + szx * szx +
                              szy * szy + szz * szz);
    const double det = sxx * (syy * szz - syz * szy) - sxy * (syx * szz - syz * szx) + sxz * (syx * szy - syy * szx);
    const double c1 = -8.0 * det;
    const double k00 = sxx + syy + szz, k11 = sxx - syy - szz, k22 = syy - sxx - szz, k33 = szz - sxx - syy;
    const double k01 = syz - szy, k02 = szx - sxz, k03 = sxy - syx, k12 = sxy + syx, k13 = szx + sxz, k23 = syz + szy;
    const double a0 = k00 * k11 - k01 * k01, a1 = k00 * k12 - k02 * k01, a2 = k00 * k13 - k03 * k01;
    const double a3 = k01 * k12 - k02 * k11, a4 = k01 * k13 - k03 * k11, a5 = k02 * k13 - k03 * k12;
    const double b0 = k02 * k13 - k12 * k03, b1 = k02 * k23 - k22 * k03, b2 = k02 * k33 - k23 * k03;
    const double b3 = k12 * k23 - k22 * k13, b4 = k12 * k33 - k23 * k13, b5 = k22 * k33 - k23 * k23;
    const double c0 = a0 * b5 - a1 * b4 + a2 * b3 + a3 * b2 - a4 * b1 + a5 * b0;
    double x = e0;
    for (int it = 0; it < 200; ++it) {
        const double x2 = x * x;
        const double b = (x2 + c2) * x;
        const double a = b + c1;
        const double p = a * x + c0;
        const double dp = 2.0 * x2 * x + b + a;
        if (dp == 0.0) break;
        const double xn = x - p / dp;
        if (!(xn == xn)) break;
        const double step = fabs(xn - x);
        x = xn;
        if (step <= 1e-15 * fabs(x)) break;
    }
    return x;
}

// Cross-covariance of the block's fit frame (wq in shared memory: w_a * (x_a - c)) with
// reference frame r, by one warp; every lane returns the 9 totals.
__device__ __forceinline__ void warp_cross_cov(const double *wq, const float *raw_r, const double *cen_r, int A,
                                               int lane, double *tot)
{
#pragma unroll
    for (int c = 0; c < 9; ++c) tot[c] = 0.0;
    const double cx = cen_r[0], cy = cen_r[1], cz = cen_r[2];
    for (int n = lane; n < A; n += 32) {
        const double y0 = (double)raw_r[3 * n + 0] - cx, y1 = (double)raw_r[3 * n + 1] - cy,
                     y2 = (double)raw_r[3 * n + 2] - cz;
        const double x0 = wq[3 * n + 0], x1 = wq[3 * n + 1], x2 = wq[3 * n + 2];
        tot[0] += x0 * y0; tot[1] += x0 * y1; tot[2] += x0 * y2;
        tot[3] += x1 * y0; tot[4] += x1 * y1; tot[5] += x1 * y2;
        tot[6] += x2 * y0; tot[7] += x2 * y1; tot[8] += x2 * y2;
    }
#pragma unroll
    for (int c = 0; c < 9; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot[c] += __shfl_xor_sync(0xffffffffu, tot[c], o);
}

__device__ __forceinline__ void load_fit_frame(double *wq, const float *raw_q, const double *cen_q,
                                               const double *wnorm, int A)
{
    for (int i = threadIdx.x; i < 3 * A; i += blockDim.x) {
        const int a = i / 3, d = i - 3 * a;
        wq[i] = wnorm[a] * ((double)raw_q[i] - cen_q[d]);
    }
}


struct RescoreArgs {
    FrameSetView fit, ref;
    long long fit_begin, n_fit;
    const double *wnorm;
    int do_fit, k1, P;  // P = pow2 >= H * keep
    CandLists<float> cl;
    double eps_scale;
    float g_ref_max;
    int fit_part;            // column of fit.gres the sweep's fit operand corresponds to, -1: no operand-rounding term
    int ref_parts;           // fp16 parts of the sweep's reference operand (1: fh, 2: fh + fl)
    float gres_ref_max;      // largest rounding residual norm of the reference operand (nm)
    double *out_dist;
    int *out_idx, *flags;
    double *err_stats;
    int *n_bad, *bad_rows;
};

// One block per fit row:
//   0. merge the row's H candidate lists and order them by approximate key
//   1. exact FP64 distance of the candidates in approximate order, one round (RESCORE_ROUND
//      candidates) at a time: warps build the cross-covariances cooperatively, then one thread
//      per candidate solves the FP64 QCP
//   2. after each round: (distance, index) order of everything re-scored so far and the
//      certificate.  Every pair not yet re-scored has approximate d^2 >= a_next (the next
//      candidate's key, or the lists' admission threshold tau_row).  Write the filter error of
//      a pair as (row-common bias) + noise with |noise| <= eps; then no such pair can beat the
//      exact k1-th neighbour if   max approx key of the exact top-k1 + 2 eps < a_next
//      (the bias -- fp32 accumulation truncation in the tensor core -- cancels).  eps is a
//      model bound (eps_scale * E0), re-checked against the spread observed on the candidates.
//      Rows that exhaust their candidates uncertified go to the exact FP64 fallback.
//   2'. 2xFP16 / 1xFP16 sweeps (fit_part >= 0).  Their keys are, up to the same bias + noise, the
//      exact min-RMSD^2 between the ROUNDED structures x~, y~ the tensor cores saw.  min_R |x - R y|
//      is a metric on centred-or-not point sets modulo rotation, so the rounded distance d~ obeys
//      |d~ - d| <= |x - x~| + |y - y~| <= g  (g = gres_q + max gres_r, FP64 norms from pack.cu):
//            max(d - g, 0)^2 + B - eps  <=  key  <=  (d + g)^2 + B + eps.
//      Every re-scored candidate i therefore brackets the bias, B <= key_i - max(d_i - g, 0)^2 + eps,
//      and a pair that was not re-scored (key >= a_next) has (d + g)^2 >= a_next - B - eps.  It
//      cannot beat the exact k1-th neighbour d_k if
//            min_i [key_i - max(d_i - g, 0)^2] + (d_k + g)^2 + 2 eps < a_next.
//      The brackets of all re-scored candidates must intersect (model check, replaces the spread
//      test).  With g = 0 this is the rule of step 2 with min_i err_i in place of the top-k1 maximum.
//      For the first ROUNDED_EXACT candidates of a row the kernel also computes d~ itself -- the FP64
//      min-RMSD of the two ROUNDED structures, from the fp16 planes -- so their bracket of B is exact
//      (key - d~^2, no g on that side): only the g that ties an UNSEEN pair's d to its d~ remains.
constexpr int RESCORE_ROUND = 64;
constexpr int ROUNDED_EXACT = 8;
constexpr int RESCORE_PMAX = 2048;   // working-set cap (candidates with key <= tau_row); a row that exceeds it is redone exactly

__global__ void __launch_bounds__(128) rms_rescore_kernel(RescoreArgs a)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    const int A = a.fit.A, P = a.P;
    double *wq = reinterpret_cast<double *>(dsm);          // [3A]
    double *wqh = wq + 3 * A;                              // [3A] the fit row as the sweep saw it: fp16 planes / 64
    double *s_d = wqh + 3 * A;                             // [P] sort keys
    double *u_dist = s_d + P;                              // [P] exact distance, candidate order
    double *s_S = u_dist + P;                              // [RESCORE_ROUND][10] cross-covariance + Gr
    double *s_Sh = s_S + RESCORE_ROUND * 10;               // [ROUNDED_EXACT][10] the same between the rounded structures
    int *s_i = reinterpret_cast<int *>(s_Sh + ROUNDED_EXACT * 10);  // [P]
    int *u_i = s_i + P;                                    // [P] candidate reference index
    float *u_apx = reinterpret_cast<float *>(u_i + P);     // [P] approximate d^2
    __shared__ int s_total, s_ok;
    __shared__ float s_taumin;
    __shared__ double s_gqh;
    __shared__ unsigned s_dtil;
    __shared__ unsigned long long s_emin, s_emax;   // order-preserving encodings of the error range
    __shared__ unsigned long long s_nmin, s_nmax;   // range of key - d~^2 over the candidates with an exact d~ (pure accumulation error)

    const long long q = blockIdx.x;
    const long long qf = a.fit_begin + q;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double *cen_q = a.fit.cen + 4 * qf;
    load_fit_frame(wq, a.fit.raw + (size_t)qf * A * 3, cen_q, a.wnorm, A);
    const double gq = cen_q[3];
    const int A_pad = a.fit.A_pad;
    if (a.fit_part >= 0) {     // rounded fit row, atom-major like wq
        const __half *fh = static_cast<const __half *>(a.fit.fh) + (size_t)qf * 3 * A_pad;
        for (int i = threadIdx.x; i < 3 * A; i += blockDim.x) {
            const int n = i / 3, d = i - 3 * n;
            wqh[i] = (double)__half2float(fh[d * A_pad + n]) * (1.0 / kRmsHalfScale);
        }
    }
    const double kInfD = __longlong_as_double(0x7ff0000000000000LL);
    const float kInfF = __uint_as_float(0x7f800000u);
    const int k1 = a.k1;

    // ---- 0. gather + order by approximate key ---------------------------------------------------
    if (threadIdx.x == 0) { s_total = 0; s_taumin = kInfF; s_ok = 0; s_emin = ~0ull; s_emax = 0ull; s_nmin = ~0ull; s_nmax = 0ull; s_gqh = 0.0; }
    for (int i = threadIdx.x; i < P; i += blockDim.x) { s_d[i] = kInfD; s_i[i] = 0x7fffffff; }
    __syncthreads();
    // admission threshold of the row = the smallest of its lists' thresholds; entries above it carry no
    // information (every pair outside the lists has key >= its list's tau >= tau_row)
    if (threadIdx.x == 0) {
        float t = kInfF;
        for (int h = 0; h < a.cl.H; ++h) t = fminf(t, a.cl.tau[(size_t)q * a.cl.H + h]);
        s_taumin = t;
    }
    __syncthreads();
    for (int h = 0; h < a.cl.H; ++h) {
        const size_t lid = (size_t)q * a.cl.H + h;
        const int c = min(a.cl.cnt[lid], a.cl.keep);
        for (int i = threadIdx.x; i < c; i += blockDim.x) {
            const float kv = a.cl.key[lid * a.cl.cap + i];
            if (kv <= s_taumin) {
                const int pos = atomicAdd(&s_total, 1);
                if (pos < P) { s_d[pos] = (double)kv; s_i[pos] = a.cl.idx[lid * a.cl.cap + i]; }
            }
        }
    }
    __syncthreads();
    const bool overflow = s_total > P;         // cannot happen unless H * keep > RESCORE_PMAX
    __syncthreads();
    if (threadIdx.x == 0 && overflow) s_total = 0;
    __syncthreads();
    const int total = s_total;
    int Pr = 32;                       // this row's working size: the lists are mostly short
    while (Pr < total) Pr <<= 1;
    const float tau_row = s_taumin;
    block_bitonic_sort(s_d, s_i, Pr);
    for (int i = threadIdx.x; i < Pr; i += blockDim.x) {
        u_apx[i] = i < total ? (float)s_d[i] : kInfF;
        u_i[i] = i < total ? s_i[i] : 0x7fffffff;
        u_dist[i] = kInfD;
    }
    __syncthreads();

    // Noise bound: eps_scale * E0 of the pair.  A pair that could still enter the top k1 has
    // RMSD^2 <= d2_k + margin, and RMSD >= |sqrt(G_q) - sqrt(G_r)| (both frames centred), so its
    // G_r <= (sqrt(G_q) + sqrt(d2_k + margin))^2 -- usually far below the largest G of the set.
    const double eps_max = a.eps_scale * 0.5 * (gq + (double)a.g_ref_max);
    // operand-rounding allowance of the reduced FP16 sweeps (0 for the full-precision modes)
    const double grd = a.fit_part >= 0 ? (double)a.fit.gres[2 * qf + a.fit_part] + (double)a.gres_ref_max : 0.0;
    double eps = eps_max;
    int done = 0;
    bool certified = false;
    while (done < total) {
        // ---- 1. one round of exact distances ----------------------------------------------------
        const int nb = min(RESCORE_ROUND, total - done);
        for (int c = warp; c < nb; c += 4) {
            const int r = u_i[done + c];
            double tot[9];
            warp_cross_cov(wq, a.ref.raw + (size_t)r * A * 3, a.ref.cen + 4 * (size_t)r, A, lane, tot);
            if (lane < 9) s_S[c * 10 + lane] = tot[lane];
            if (lane == 9) s_S[c * 10 + 9] = a.ref.cen[4 * (size_t)r + 3];
        }
        const bool rounded_round = a.fit_part >= 0 && done == 0;
        if (rounded_round) {
            // the same contraction between the ROUNDED structures (what the sweep's tensor cores multiplied), in FP64
            for (int c = warp; c < min(nb, ROUNDED_EXACT); c += 4) {
                const size_t rb = (size_t)u_i[c] * 3 * A_pad;
                const __half *rh = static_cast<const __half *>(a.ref.fh) + rb;
                const __half *rl = static_cast<const __half *>(a.ref.fl) + rb;
                double tot[10];
#pragma unroll
                for (int e = 0; e < 10; ++e) tot[e] = 0.0;
                for (int n = lane; n < A; n += 32) {
                    double y0 = (double)__half2float(rh[n]), y1 = (double)__half2float(rh[A_pad + n]),
                           y2 = (double)__half2float(rh[2 * A_pad + n]);
                    if (a.ref_parts == 2) {
                        y0 += (double)__half2float(rl[n]); y1 += (double)__half2float(rl[A_pad + n]);
                        y2 += (double)__half2float(rl[2 * A_pad + n]);
                    }
                    y0 *= 1.0 / kRmsHalfScale; y1 *= 1.0 / kRmsHalfScale; y2 *= 1.0 / kRmsHalfScale;
                    const double x0 = wqh[3 * n + 0], x1 = wqh[3 * n + 1], x2 = wqh[3 * n + 2];
                    tot[0] += x0 * y0; tot[1] += x0 * y1; tot[2] += x0 * y2;
                    tot[3] += x1 * y0; tot[4] += x1 * y1; tot[5] += x1 * y2;
                    tot[6] += x2 * y0; tot[7] += x2 * y1; tot[8] += x2 * y2;
                    tot[9] += y0 * y0 + y1 * y1 + y2 * y2;
                }
#pragma unroll
                for (int e = 0; e < 10; ++e)
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) tot[e] += __shfl_xor_sync(0xffffffffu, tot[e], o);
                if (lane == 0)
#pragma unroll
                    for (int e = 0; e < 10; ++e) s_Sh[c * 10 + e] = tot[e];
            }
            if (warp == 0) {               // |x~|^2 of the fit row (one warp: deterministic order)
                double gs = 0.0;
                for (int i = lane; i < 3 * A; i += 32) gs += wqh[i] * wqh[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) gs += __shfl_xor_sync(0xffffffffu, gs, o);
                if (lane == 0) s_gqh = gs;
            }
        }
        __syncthreads();
        if (threadIdx.x < nb) {
            const int c = threadIdx.x;
            double S[9];
#pragma unroll
            for (int e = 0; e < 9; ++e) S[e] = s_S[c * 10 + e];
            const double e0 = 0.5 * (gq + s_S[c * 10 + 9]);
            const double lam = a.do_fit ? qcp_lambda_f64(S, e0) : (S[0] + S[4] + S[8]);
            const double d2 = fmax(2.0 * (e0 - lam), 0.0);
            // distance exactly as distance() reports it: sqrt(msd) [nm] * 10.0 -> Angstrom
            u_dist[done + c] = sqrt(d2) * 10.0;
            // bracket of the row-common bias given this candidate: [key - (d+g)^2, key - max(d-g,0)^2];
            // both ends are approx - exact when g = 0
            const double dn = sqrt(d2), dm = fmax(dn - grd, 0.0);
            double err_hi = (double)u_apx[done + c] - dm * dm;
            double err_lo = (double)u_apx[done + c] - (dn + grd) * (dn + grd);
            // monotone map double -> u64 so that atomicMin/Max order like the doubles
            auto enc = [](double v) {
                unsigned long long bits = (unsigned long long)__double_as_longlong(v);
                return (bits >> 63) ? ~bits : (bits | 0x8000000000000000ull);
            };
            if (rounded_round && c < ROUNDED_EXACT) {   // exact d~ of this candidate: its bracket is a point
                double Sh[9];
#pragma unroll
                for (int e = 0; e < 9; ++e) Sh[e] = s_Sh[c * 10 + e];
                const double e0h = 0.5 * (s_gqh + s_Sh[c * 10 + 9]);
                const double lamh = a.do_fit ? qcp_lambda_f64(Sh, e0h) : (Sh[0] + Sh[4] + Sh[8]);
                err_hi = err_lo = (double)u_apx[done + c] - fmax(2.0 * (e0h - lamh), 0.0);
                atomicMin(&s_nmin, enc(err_hi));
                atomicMax(&s_nmax, enc(err_hi));
            }
            atomicMin(&s_emin, enc(err_hi));
            atomicMax(&s_emax, enc(err_lo));
        }
        __syncthreads();
        done += nb;
        // ---- 2. order + certificate ----------------------------------------------------------------
        int Ps = 32;                       // only the re-scored prefix carries finite keys
        while (Ps < done) Ps <<= 1;
        for (int i = threadIdx.x; i < Ps; i += blockDim.x) { s_d[i] = u_dist[i]; s_i[i] = u_i[i]; }
        if (threadIdx.x == 0) s_dtil = 0u;
        __syncthreads();
        block_bitonic_sort(s_d, s_i, Ps);
        if (done >= k1) {
            const double dk = s_d[k1 - 1];
            const int ik = s_i[k1 - 1];
            float dt = 0.0f;
            for (int i = threadIdx.x; i < done; i += blockDim.x)
                if (pair_less(u_dist[i], u_i[i], dk, ik) || (u_dist[i] == dk && u_i[i] == ik)) dt = fmaxf(dt, u_apx[i]);
            atomicMax(&s_dtil, __float_as_uint(dt));  // keys are >= +0
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
            if (ok) {
                const double dk_nm = s_d[k1 - 1] * 0.1;                       // exact k1-th distance, nm
                const double reach = sqrt(gq) + sqrt(dk_nm * dk_nm + 4.0 * eps_max);
                eps = a.eps_scale * 0.5 * (gq + fmin((double)a.g_ref_max, reach * reach));
            }
            if (ok && a_next != kInfF) {  // a_next == inf: nothing was ever dropped, every pair has been re-scored
                if (a.fit_part >= 0) {
                    const double dkg = s_d[k1 - 1] * 0.1 + grd;
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

    for (int j = threadIdx.x; j < k1; j += blockDim.x) {
        a.out_dist[(size_t)q * k1 + j] = s_d[j];
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
            // spread statistic: range of approx - exact; for the reduced FP16 sweeps the range of key - d~^2 over
            // the candidates whose rounded-structure distance was computed (the accumulation error alone)
            const double sp = a.fit_part >= 0 ? (s_nmax >= s_nmin ? dec(s_nmax) - dec(s_nmin) : 0.0) : hi - lo;
            atomic_max_nonneg(a.err_stats + 1, fmax(sp, 0.0));
            atomic_max_nonneg(a.err_stats + 2, (double)done);
        }
        a.flags[q] = certified ? 1 : 0;
        if (!certified) {
            const int pos = atomicAdd(a.n_bad, 1);
            a.bad_rows[pos] = (int)q;
        }
    }
}

cudaError_t launch_rms_rescore(const FrameSetView &fit, long long fit_begin, long long n_fit,
                               const FrameSetView &ref, const double *wnorm, int do_fit, CandLists<float> cl, int k1,
                               double eps_scale, float g_ref_max, int fit_part, int ref_parts, float gres_ref_max,
                               double *out_dist,
                               int *out_idx, int *flags,
                               double *err_stats, int *n_bad, int *bad_rows, cudaStream_t st)
{
    if (n_fit <= 0) return cudaSuccess;
    RescoreArgs a;
    a.fit = fit; a.ref = ref; a.fit_begin = fit_begin; a.n_fit = n_fit; a.wnorm = wnorm;
    a.do_fit = do_fit; a.k1 = k1; a.cl = cl; a.eps_scale = eps_scale; a.g_ref_max = g_ref_max;
    a.fit_part = fit_part; a.ref_parts = ref_parts; a.gres_ref_max = gres_ref_max;
    a.out_dist = out_dist; a.out_idx = out_idx; a.flags = flags; a.err_stats = err_stats; a.n_bad = n_bad;
    a.bad_rows = bad_rows;
    int P = 1;
    while (P < cl.keep * cl.H && P < RESCORE_PMAX) P <<= 1;
    while (P < cl.keep) P <<= 1;           // one list alone must always fit
    a.P = P;
    const size_t smem = (size_t)fit.A * 3 * 16 + (size_t)P * 28 + (RESCORE_ROUND + ROUNDED_EXACT) * 80;
    cudaError_t e = cudaFuncSetAttribute(rms_rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    rms_rescore_kernel<<<(unsigned)n_fit, 128, smem, st>>>(a);
    return cudaGetLastError();
}

// ---- exact FP64 full rows -------------------------------------------------------------------
struct ExactRowsArgs {
    FrameSetView fit, ref;
    const int *row_ids;
    long long fit_begin;
    const double *wnorm;
    int do_fit;
    double *out;  // [n_rows][n_ref] d^2 in nm^2
    long long ref_block0;
};

__global__ void __launch_bounds__(128) rms_exact_rows_kernel(ExactRowsArgs a)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    double *wq = reinterpret_cast<double *>(dsm);
    const int A = a.fit.A;
    // rows vary fastest across the grid: the blocks that share a block of reference frames run together, so
    // the frames come from L2 for all but the first row (a row alone streams the whole raw set from HBM)
    const int row = blockIdx.x;
    const long long qf = a.fit_begin + (a.row_ids ? a.row_ids[row] : row);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double *cen_q = a.fit.cen + 4 * qf;
    load_fit_frame(wq, a.fit.raw + (size_t)qf * A * 3, cen_q, a.wnorm, A);
    __syncthreads();
    const double gq = cen_q[3];
    const long long base = ((a.ref_block0 + (long long)blockIdx.y) * 4 + warp) * 32;
    if (base >= a.ref.n) return;
    const int nb = (int)min((long long)32, a.ref.n - base);
    double S[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) S[c] = 0.0;
    for (int c = 0; c < nb; ++c) {
        const long long r = base + c;
        double tot[9];
        warp_cross_cov(wq, a.ref.raw + (size_t)r * A * 3, a.ref.cen + 4 * (size_t)r, A, lane, tot);
        if (lane == c) {
#pragma unroll
            for (int e = 0; e < 9; ++e) S[e] = tot[e];
        }
    }
    if (lane < nb) {
        const long long r = base + lane;
        const double e0 = 0.5 * (gq + a.ref.cen[4 * (size_t)r + 3]);
        const double lam = a.do_fit ? qcp_lambda_f64(S, e0) : (S[0] + S[4] + S[8]);
        a.out[(size_t)row * a.ref.n + r] = fmax(2.0 * (e0 - lam), 0.0);
    }
}

cudaError_t launch_rms_exact_rows(const FrameSetView &fit, const int *row_ids, long long fit_begin, int n_rows,
                                  const FrameSetView &ref, const double *wnorm, int do_fit, double *out_d2,
                                  cudaStream_t st)
{
    if (n_rows <= 0 || ref.n <= 0) return cudaSuccess;
    ExactRowsArgs a;
    a.fit = fit; a.ref = ref; a.row_ids = row_ids; a.fit_begin = fit_begin; a.wnorm = wnorm; a.do_fit = do_fit;
    a.out = out_d2;
    const size_t smem = (size_t)fit.A * 3 * 8;
    cudaError_t e = cudaFuncSetAttribute(rms_exact_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long ref_blocks = (ref.n + 127) / 128;
    for (long long b0 = 0; b0 < ref_blocks; b0 += 65535) {      // gridDim.y limit
        ExactRowsArgs c = a;
        c.ref_block0 = b0;
        dim3 grid((unsigned)n_rows, (unsigned)std::min<long long>(65535, ref_blocks - b0));
        rms_exact_rows_kernel<<<grid, 128, smem, st>>>(c);
    }
    return cudaGetLastError();
}

// ---- exact top-k1 of full FP64 rows ------------------------------------------------------------
// One block per row.  Radix select on the 64-bit key for the k1-th smallest, then on the index
// among the entries equal to it, so the kept set is exactly the k1 smallest by (key, index).
template <typename U, int PASSES, typename LoadKey, typename Match>
__device__ inline U block_radix_select(long long n, int rank, int *taken_before, LoadKey load, Match match,
                                       unsigned *hist, U *s_bcast, int *s_rank)
{
    U prefix = 0, mask = 0;
    int remaining = rank;
    for (int pass = 0; pass < PASSES; ++pass) {
        const int shift = (PASSES - 1 - pass) * 8;
        for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0;
        __syncthreads();
        for (long long i = threadIdx.x; i < n; i += blockDim.x) {
            if (!match(i)) continue;
            const U k = load(i);
            if ((k & mask) == prefix) atomicAdd(&hist[(unsigned)(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned run = 0;
            int bin = 255;
            for (int b = 0; b < 256; ++b) {
                if (run + hist[b] >= (unsigned)remaining) { bin = b; break; }
                run += hist[b];
            }
            *s_bcast = prefix | ((U)bin << shift);
            *s_rank = remaining - (int)run;
        }
        __syncthreads();
        prefix = *s_bcast;
        remaining = *s_rank;
        mask |= (U)255 << shift;
        __syncthreads();
    }
    *taken_before = rank - remaining;  // entries strictly below the selected key
    return prefix;
}

struct SelectRowsArgs {
    const double *rows;
    long long n_ref;
    int k1, P;
    const int *row_ids;
    double scale;
    double *out_dist;
    int *out_idx;
};

__global__ void __launch_bounds__(1024) select_rows_f64_kernel(SelectRowsArgs a)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    double *s_d = reinterpret_cast<double *>(dsm);
    int *s_i = reinterpret_cast<int *>(s_d + a.P);
    __shared__ unsigned hist[256];
    __shared__ unsigned long long s_b64;
    __shared__ unsigned s_b32;
    __shared__ int s_rank, s_fill;
    const double *row = a.rows + (size_t)blockIdx.x * a.n_ref;
    using U64 = unsigned long long;
    int below = 0;
    const U64 kth = block_radix_select<U64, 8>(
        a.n_ref, a.k1, &below, [&](long long i) { return (U64)__double_as_longlong(row[i]); },
        [&](long long) { return true; }, hist, &s_b64, &s_rank);
    const int need = a.k1 - below;  // ties at kth to take, smallest indices first
    int below_i = 0;
    const unsigned ith = block_radix_select<unsigned, 4>(
        a.n_ref, need, &below_i, [&](long long i) { return (unsigned)i; },
        [&](long long i) { return (U64)__double_as_longlong(row[i]) == kth; }, hist, &s_b32, &s_rank);
    if (threadIdx.x == 0) s_fill = 0;
    for (int i = threadIdx.x; i < a.P; i += blockDim.x) {
        s_d[i] = __longlong_as_double(0x7ff0000000000000LL);
        s_i[i] = 0x7fffffff;
    }
    __syncthreads();
    for (long long i = threadIdx.x; i < a.n_ref; i += blockDim.x) {
        const U64 k = (U64)__double_as_longlong(row[i]);
        if (k < kth || (k == kth && (unsigned)i <= ith)) {
            const int pos = atomicAdd(&s_fill, 1);
            if (pos < a.P) { s_d[pos] = sqrt(row[i]) * a.scale; s_i[pos] = (int)i; }
        }
    }
    __syncthreads();
    block_bitonic_sort(s_d, s_i, a.P);
    const long long orow = a.row_ids ? a.row_ids[blockIdx.x] : blockIdx.x;
    for (int j = threadIdx.x; j < a.k1; j += blockDim.x) {
        a.out_dist[(size_t)orow * a.k1 + j] = s_d[j];
        a.out_idx[(size_t)orow * a.k1 + j] = s_i[j];
    }
}

cudaError_t launch_select_rows_f64(const double *rows_d2, int n_rows, long long n_ref, int k1, const int *row_ids,
                                   double scale, double *out_dist, int *out_idx, cudaStream_t st)
{
    if (n_rows <= 0) return cudaSuccess;
    SelectRowsArgs a;
    a.rows = rows_d2; a.n_ref = n_ref; a.k1 = k1; a.row_ids = row_ids; a.scale = scale;
    a.out_dist = out_dist; a.out_idx = out_idx;
    int P = 1;
    while (P < k1) P <<= 1;
    a.P = P;
    const size_t smem = (size_t)P * 12;
    cudaError_t e = cudaFuncSetAttribute(select_rows_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    select_rows_f64_kernel<<<(unsigned)n_rows, 1024, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace mdsctk
