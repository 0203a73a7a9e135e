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

__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v)
{
    atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}

struct RescoreArgs {
    FrameSetView fit, ref;
    long long fit_begin, n_fit;
    const double *wnorm;
    int do_fit, k1, P;  // P = pow2 >= keep
    CandLists<float> cl;
    double eps_scale;
    float g_ref_max;
    double *out_dist;
    int *out_idx, *flags;
    double *err_max;
    int *n_bad, *bad_rows;
};

__global__ void __launch_bounds__(128) rms_rescore_kernel(RescoreArgs a)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    const int A = a.fit.A;
    double *wq = reinterpret_cast<double *>(dsm);          // [3A]
    double *s_d = wq + 3 * A;                              // [P] exact d^2, then distance
    int *s_i = reinterpret_cast<int *>(s_d + a.P);         // [P]
    __shared__ double s_err[4];

    const long long q = blockIdx.x;
    const long long qf = a.fit_begin + q;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double *cen_q = a.fit.cen + 4 * qf;
    load_fit_frame(wq, a.fit.raw + (size_t)qf * A * 3, cen_q, a.wnorm, A);
    const double gq = cen_q[3];
    const int cnt = min(a.cl.cnt[q], a.cl.keep);
    const size_t lbase = (size_t)q * a.cl.cap;
    for (int i = threadIdx.x; i < a.P; i += blockDim.x) {
        s_d[i] = __longlong_as_double(0x7ff0000000000000LL);
        s_i[i] = 0x7fffffff;
    }
    __syncthreads();

    double err = 0.0;
    for (int b = warp * 32; b < cnt; b += 4 * 32) {
        double S[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) S[c] = 0.0;
        const int nb = min(32, cnt - b);
        for (int c = 0; c < nb; ++c) {
            const int r = a.cl.idx[lbase + b + c];
            double tot[9];
            warp_cross_cov(wq, a.ref.raw + (size_t)r * A * 3, a.ref.cen + 4 * (size_t)r, A, lane, tot);
            if (lane == c) {
#pragma unroll
                for (int e = 0; e < 9; ++e) S[e] = tot[e];
            }
        }
        if (lane < nb) {
            const int r = a.cl.idx[lbase + b + lane];
            const double e0 = 0.5 * (gq + a.ref.cen[4 * (size_t)r + 3]);
            const double lam = a.do_fit ? qcp_lambda_f64(S, e0) : (S[0] + S[4] + S[8]);
            const double d2 = fmax(2.0 * (e0 - lam), 0.0);
            s_d[b + lane] = d2;
            s_i[b + lane] = r;
            err = fmax(err, fabs(d2 - (double)a.cl.key[lbase + b + lane]));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) err = fmax(err, __shfl_xor_sync(0xffffffffu, err, o));
    if (lane == 0) s_err[warp] = err;
    __syncthreads();
    // distance exactly as distance() reports it: sqrt(msd) [nm] * 10.0 -> Angstrom
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) s_d[i] = sqrt(s_d[i]) * 10.0;
    __syncthreads();
    block_bitonic_sort(s_d, s_i, a.P);

    const int k1 = a.k1;
    for (int j = threadIdx.x; j < k1; j += blockDim.x) {
        a.out_dist[(size_t)q * k1 + j] = s_d[j];
        a.out_idx[(size_t)q * k1 + j] = s_i[j];
    }
    if (threadIdx.x == 0) {
        const double row_err = fmax(fmax(s_err[0], s_err[1]), fmax(s_err[2], s_err[3]));
        atomic_max_nonneg(a.err_max, row_err);
        const float tau = a.cl.tau[q];
        bool ok;
        if (tau == __uint_as_float(0x7f800000u)) {
            ok = cnt >= k1;  // nothing was ever dropped: the list holds every pair
        } else {
            const double eps = a.eps_scale * 0.5 * (gq + (double)a.g_ref_max);
            const double dk = s_d[k1 - 1] * 0.1;
            // a non-candidate has approx d^2 >= tau, hence exact d^2 >= tau - eps: it cannot enter
            // the top k1 if the exact k1-th candidate distance is below that.
            ok = cnt >= k1 && (dk * dk + eps < (double)tau) && row_err <= eps;
        }
        a.flags[q] = ok ? 1 : 0;
        if (!ok) {
            const int pos = atomicAdd(a.n_bad, 1);
            a.bad_rows[pos] = (int)q;
        }
    }
}

cudaError_t launch_rms_rescore(const FrameSetView &fit, long long fit_begin, long long n_fit,
                               const FrameSetView &ref, const double *wnorm, int do_fit, CandLists<float> cl, int k1,
                               double eps_scale, float g_ref_max, double *out_dist, int *out_idx, int *flags,
                               double *err_max, int *n_bad, int *bad_rows, cudaStream_t st)
{
    if (n_fit <= 0) return cudaSuccess;
    RescoreArgs a;
    a.fit = fit; a.ref = ref; a.fit_begin = fit_begin; a.n_fit = n_fit; a.wnorm = wnorm;
    a.do_fit = do_fit; a.k1 = k1; a.cl = cl; a.eps_scale = eps_scale; a.g_ref_max = g_ref_max;
    a.out_dist = out_dist; a.out_idx = out_idx; a.flags = flags; a.err_max = err_max; a.n_bad = n_bad;
    a.bad_rows = bad_rows;
    int P = 1;
    while (P < cl.keep) P <<= 1;
    a.P = P;
    const size_t smem = (size_t)fit.A * 3 * 8 + (size_t)P * 12;
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
};

__global__ void __launch_bounds__(128) rms_exact_rows_kernel(ExactRowsArgs a)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    double *wq = reinterpret_cast<double *>(dsm);
    const int A = a.fit.A;
    const int row = blockIdx.y;
    const long long qf = a.fit_begin + (a.row_ids ? a.row_ids[row] : row);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double *cen_q = a.fit.cen + 4 * qf;
    load_fit_frame(wq, a.fit.raw + (size_t)qf * A * 3, cen_q, a.wnorm, A);
    __syncthreads();
    const double gq = cen_q[3];
    const long long base = ((long long)blockIdx.x * 4 + warp) * 32;
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
    dim3 grid((unsigned)((ref.n + 127) / 128), (unsigned)n_rows);
    rms_exact_rows_kernel<<<grid, 128, smem, st>>>(a);
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
