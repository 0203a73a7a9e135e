// data_knn.cu -- knn_data on the GPU: exact FP64 all-pairs vector distances + streaming top-k.
//
// Replaces the row loop of knn_data.cpp:195-250 with ::distance = euclidean_distance
// (mdsctk.cpp:330-335) or correlation_distance (mdsctk.cpp:337-360, knn_data -c).
// The arithmetic is kept operation-for-operation in FP64 -- separate multiply and add
// (__dmul_rn/__dadd_rn, no FMA contraction), the same left-to-right summation over the
// vector -- so distances are BIT-IDENTICAL to the CPU tools and no re-score is needed.
// Selection keys are the pre-sqrt values; sqrt is applied to the survivors only.
#include "common.cuh"
#include "select.cuh"
#include "sort.cuh"

namespace mdsctk {

namespace dk {
constexpr int TQ = 64, TR = 64, KC = 16, NTHR = 256, PITCH = 65;
}

// Per-row mean and spread exactly as correlation_distance computes them for each argument:
//   sum += v[x]; sq += v[x]*v[x];  spread = sqrt((n*sq - sum*sum) / (n*(n-1)));  mean = sum/n
__global__ void data_rowstats_kernel(const double *__restrict__ rows, long long n, int dim, double *__restrict__ stats)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const double *v = rows + (size_t)r * dim;
    double sum = 0.0, sq = 0.0;
    for (int x = 0; x < dim; ++x) {
        sum = __dadd_rn(sum, v[x]);
        sq = __dadd_rn(sq, __dmul_rn(v[x], v[x]));
    }
    const double dn = (double)dim;
    const double spread =
        sqrt(__ddiv_rn(__dsub_rn(__dmul_rn(dn, sq), __dmul_rn(sum, sum)), __dmul_rn(dn, __dsub_rn(dn, 1.0))));
    stats[2 * r + 0] = __ddiv_rn(sum, dn);
    stats[2 * r + 1] = spread;
}

cudaError_t launch_data_rowstats(const double *rows, long long n, int dim, double *stats, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    data_rowstats_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(rows, n, dim, stats);
    return cudaGetLastError();
}

struct DataArgs {
    const double *fit, *fit_stats, *ref, *ref_stats;
    long long n_fit, n_ref;
    int dim, metric;
    CandLists<double> cl;
};

__global__ void __launch_bounds__(dk::NTHR, 2) data_sweep_kernel(DataArgs a)
{
    using namespace dk;
    __shared__ double s_f[KC][PITCH];
    __shared__ double s_r[KC][PITCH];
    __shared__ int s_cnt[TQ];
    __shared__ double s_tau[TQ];
    __shared__ unsigned s_hist[NTHR / 32][256];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;  // reference rows tx+16j, fit rows ty+16i
    const long long q0 = (long long)blockIdx.x * TQ;
    if (tid < TQ) { s_cnt[tid] = 0; s_tau[tid] = KeyBits<double>::inf(); }

    double fm[4] = {0, 0, 0, 0}, fs[4] = {1, 1, 1, 1};
    if (a.metric == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            long long f = q0 + ty + 16 * i;
            if (f >= a.n_fit) f = a.n_fit - 1;
            fm[i] = a.fit_stats[2 * f];
            fs[i] = a.fit_stats[2 * f + 1];
        }
    }
    const long long n_rt = (a.n_ref + TR - 1) / TR;
    for (long long rt = 0; rt < n_rt; ++rt) {
        const long long r0 = rt * TR;
        double rm[4] = {0, 0, 0, 0}, rs[4] = {1, 1, 1, 1};
        if (a.metric == 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                long long r = r0 + tx + 16 * j;
                if (r >= a.n_ref) r = a.n_ref - 1;
                rm[j] = a.ref_stats[2 * r];
                rs[j] = a.ref_stats[2 * r + 1];
            }
        }
        double acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

        for (int k0 = 0; k0 < a.dim; k0 += KC) {
            const int kv = min(KC, a.dim - k0);
            __syncthreads();
            for (int i = tid; i < TQ * KC; i += NTHR) {
                const int row = i / KC, k = i - row * KC;
                long long f = q0 + row, r = r0 + row;
                if (f >= a.n_fit) f = a.n_fit - 1;
                if (r >= a.n_ref) r = a.n_ref - 1;
                s_f[k][row] = (k < kv) ? a.fit[(size_t)f * a.dim + k0 + k] : 0.0;
                s_r[k][row] = (k < kv) ? a.ref[(size_t)r * a.dim + k0 + k] : 0.0;
            }
            __syncthreads();
            if (a.metric == 0) {
                for (int k = 0; k < kv; ++k) {
                    double fv[4], rv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) fv[i] = s_f[k][ty + 16 * i];
#pragma unroll
                    for (int j = 0; j < 4; ++j) rv[j] = s_r[k][tx + 16 * j];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const double d = __dsub_rn(fv[i], rv[j]);  // (reference[x] - fitting[x]), reference = fit row
                            acc[i][j] = __dadd_rn(acc[i][j], __dmul_rn(d, d));
                        }
                }
            } else {
                for (int k = 0; k < kv; ++k) {
                    double fv[4], rv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) fv[i] = __dsub_rn(s_f[k][ty + 16 * i], fm[i]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) rv[j] = __dsub_rn(s_r[k][tx + 16 * j], rm[j]);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = __dadd_rn(acc[i][j], __dmul_rn(fv[i], rv[j]));
                }
            }
        }
        // ---- keys + threshold-gated append ------------------------------------------------
        const double dn1 = __dsub_rn((double)a.dim, 1.0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int ql = ty + 16 * i;
            const long long qrow = q0 + ql;
            const double tau = s_tau[ql];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long ridx = r0 + tx + 16 * j;
                double key = acc[i][j];
                if (a.metric == 1) {
                    // (1.0 - (value / ((dsize - 1.0) * rsvalue * fsvalue))) / 2.0 ; clamp at 0
                    const double den = __dmul_rn(__dmul_rn(dn1, fs[i]), rs[j]);
                    key = __ddiv_rn(__dsub_rn(1.0, __ddiv_rn(key, den)), 2.0);
                    if (key < 0.0) key = 0.0;
                }
                if (key < tau && qrow < a.n_fit && ridx < a.n_ref) {
                    const int pos = atomicAdd(&s_cnt[ql], 1);
                    if (pos < a.cl.cap) {
                        a.cl.key[(size_t)qrow * a.cl.cap + pos] = key;
                        a.cl.idx[(size_t)qrow * a.cl.cap + pos] = (int)ridx;
                    }
                }
            }
        }
        __syncthreads();
        for (int ql = warp; ql < TQ; ql += NTHR / 32) {
            const int c = min(s_cnt[ql], a.cl.cap);
            if (c > a.cl.cap - TR && q0 + ql < a.n_fit) {
                const size_t base = (size_t)(q0 + ql) * a.cl.cap;
                double nt = warp_compact_list<double>(a.cl.key + base, a.cl.idx + base, c, a.cl.keep, s_hist[warp]);
                if (lane == 0) { s_tau[ql] = nt; s_cnt[ql] = a.cl.keep; }
            }
        }
        __syncthreads();
    }
    for (int ql = warp; ql < TQ; ql += NTHR / 32) {
        if (q0 + ql >= a.n_fit) continue;
        int c = min(s_cnt[ql], a.cl.cap);
        double tau = s_tau[ql];
        const size_t base = (size_t)(q0 + ql) * a.cl.cap;
        if (c > a.cl.keep) {
            tau = warp_compact_list<double>(a.cl.key + base, a.cl.idx + base, c, a.cl.keep, s_hist[warp]);
            c = a.cl.keep;
        }
        if (lane == 0) { a.cl.cnt[q0 + ql] = c; a.cl.tau[q0 + ql] = tau; }
    }
}

cudaError_t launch_data_sweep(const double *fit, const double *fit_stats, long long n_fit, const double *ref,
                              const double *ref_stats, long long n_ref, int dim, int metric, CandLists<double> cl,
                              cudaStream_t st)
{
    if (n_fit <= 0) return cudaSuccess;
    DataArgs a;
    a.fit = fit; a.fit_stats = fit_stats; a.ref = ref; a.ref_stats = ref_stats;
    a.n_fit = n_fit; a.n_ref = n_ref; a.dim = dim; a.metric = metric; a.cl = cl;
    data_sweep_kernel<<<(unsigned)((n_fit + dk::TQ - 1) / dk::TQ), dk::NTHR, 0, st>>>(a);
    return cudaGetLastError();
}

// ---- full distance rows (knn_data --sort false, knn_data.cpp:198-216): one thread per pair ----------
__global__ void data_exact_rows_kernel(const double *__restrict__ fit, const double *__restrict__ fit_stats, long long n_fit,
                                       const double *__restrict__ ref, const double *__restrict__ ref_stats, long long n_ref,
                                       int dim, int metric, double *__restrict__ out)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long f = blockIdx.y;
    if (r >= n_ref || f >= n_fit) return;
    const double *fv = fit + (size_t)f * dim, *rv = ref + (size_t)r * dim;
    double acc = 0.0;
    if (metric == 0) {
        for (int x = 0; x < dim; ++x) {
            const double d = __dsub_rn(fv[x], rv[x]);
            acc = __dadd_rn(acc, __dmul_rn(d, d));
        }
    } else {
        const double fm = fit_stats[2 * f], fs = fit_stats[2 * f + 1], rm = ref_stats[2 * r], rs = ref_stats[2 * r + 1];
        for (int x = 0; x < dim; ++x) acc = __dadd_rn(acc, __dmul_rn(__dsub_rn(fv[x], fm), __dsub_rn(rv[x], rm)));
        const double den = __dmul_rn(__dmul_rn(__dsub_rn((double)dim, 1.0), fs), rs);
        acc = __ddiv_rn(__dsub_rn(1.0, __ddiv_rn(acc, den)), 2.0);
        if (acc < 0.0) acc = 0.0;
    }
    out[(size_t)f * n_ref + r] = sqrt(acc);
}

cudaError_t launch_data_exact_rows(const double *fit, const double *fit_stats, long long n_fit, const double *ref,
                                   const double *ref_stats, long long n_ref, int dim, int metric, double *out, cudaStream_t st)
{
    if (n_fit <= 0 || n_ref <= 0) return cudaSuccess;
    dim3 grid((unsigned)((n_ref + 127) / 128), (unsigned)n_fit);
    data_exact_rows_kernel<<<grid, 128, 0, st>>>(fit, fit_stats, n_fit, ref, ref_stats, n_ref, dim, metric, out);
    return cudaGetLastError();
}

// ---- final (distance, index) sort of the survivors; distance = sqrt(key) --------------------
__global__ void __launch_bounds__(128) data_finalize_kernel(CandLists<double> cl, int k1, int P, double *out_dist,
                                                            int *out_idx)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    double *s_d = reinterpret_cast<double *>(dsm);
    int *s_i = reinterpret_cast<int *>(s_d + P);
    const long long q = blockIdx.x;
    const int cnt = min(cl.cnt[q], cl.keep);
    const size_t base = (size_t)q * cl.cap;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        if (i < cnt) { s_d[i] = sqrt(cl.key[base + i]); s_i[i] = cl.idx[base + i]; }
        else { s_d[i] = __longlong_as_double(0x7ff0000000000000LL); s_i[i] = 0x7fffffff; }
    }
    __syncthreads();
    block_bitonic_sort(s_d, s_i, P);
    for (int j = threadIdx.x; j < k1; j += blockDim.x) {
        out_dist[(size_t)q * k1 + j] = s_d[j];
        out_idx[(size_t)q * k1 + j] = s_i[j];
    }
}

cudaError_t launch_data_finalize(CandLists<double> cl, long long n_fit, int k1, double *out_dist, int *out_idx,
                                 cudaStream_t st)
{
    if (n_fit <= 0) return cudaSuccess;
    int P = 1;
    while (P < cl.keep) P <<= 1;
    const size_t smem = (size_t)P * 12;
    cudaError_t e = cudaFuncSetAttribute(data_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    data_finalize_kernel<<<(unsigned)n_fit, 128, smem, st>>>(cl, k1, P, out_dist, out_idx);
    return cudaGetLastError();
}

}  // namespace mdsctk
