// pack.cu -- frame loading on the GPU: centre, mass-scale and pack decoded frames.
//
// Replaces the per-frame reset_x(natoms,NULL,natoms,NULL,x,weights) of the load loops at
// knn_rms.cpp:186-206 (GROMACS: float centre-of-mass removal) and sets up the operand
// layout of the contraction: AoS float[n][A][3] (nm) -> planes float[n][3][A_pad] holding
// sqrt(m_a/M) * (x_a - c), G = sum (m_a/M)|x_a - c|^2, and the FP64 centroid for the re-score.
// plus the TF32 hi/lo split planes of the tensor-core sweep.
// plus the BF16 and FP16 split planes -- each operand family only when its pointer is given, so a run
// writes just what its sweep kernel reads (the default 1xFP16 sweep: fh, 6*A_pad bytes per frame).
// One warp per frame; HBM-bound: 12*A bytes read + 6*A_pad (default) .. 60*A_pad (all families) written.
// For the reduced-precision FP16 sweeps (2xFP16 / 1xFP16) the kernel also records, in FP64, what the
// rounding did to each frame: Gh = |fh/64|^2 and G2 = |(fh+fl)/64|^2 (norms of the structures the
// tensor cores actually see) and the residual norms g1 = |x - fh/64|, g2 = |x - (fh+fl)/64| (nm,
// rounded up).  min-RMSD over rotations is a metric, so the distance between two ROUNDED structures
// differs from the true one by at most g_q + g_r: the certificate's rigorous operand-rounding term.
#include "common.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace mdsctk {

__device__ __forceinline__ float tf32_rn(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__global__ void __launch_bounds__(256) pack_frames_kernel(const float *__restrict__ raw,
                                                          const double *__restrict__ wnorm,  // [2A]: m_a / M, then sqrt(m_a / M) (FP64)
                                                          long long n, int A, int A_pad,
                                                          float *__restrict__ planes, float *__restrict__ hi,
                                                          float *__restrict__ lo, __nv_bfloat16 *__restrict__ bh,
                                                          __nv_bfloat16 *__restrict__ bm, __half *__restrict__ fh,
                                                          __half *__restrict__ fl, float *__restrict__ G,
                                                          double *__restrict__ cen, float *__restrict__ Gh,
                                                          float *__restrict__ G2, float2 *__restrict__ gres,
                                                          float4 *__restrict__ sig)
{
    const int lane = threadIdx.x & 31;
    const long long f = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= n) return;
    const float *x = raw + (size_t)f * A * 3;
    double cx = 0.0, cy = 0.0, cz = 0.0;
    for (int a = lane; a < A; a += 32) {
        const double w = wnorm[a];
        cx += w * (double)x[3 * a + 0];
        cy += w * (double)x[3 * a + 1];
        cz += w * (double)x[3 * a + 2];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cx += __shfl_xor_sync(0xffffffffu, cx, o);
        cy += __shfl_xor_sync(0xffffffffu, cy, o);
        cz += __shfl_xor_sync(0xffffffffu, cz, o);
    }
    const size_t pbase = (size_t)f * 3 * A_pad;
    float *px = planes ? planes + pbase : nullptr;   // every operand family is optional: only what the chosen sweep reads is written
    float *py = px ? px + A_pad : nullptr, *pz = px ? px + 2 * A_pad : nullptr;
    double g = 0.0, gh = 0.0, g2n = 0.0, r1 = 0.0, r2 = 0.0;
    double t00 = 0.0, t01 = 0.0, t02 = 0.0, t11 = 0.0, t12 = 0.0, t22 = 0.0;   // gyration tensor of the weighted frame
    for (int a = lane; a < A_pad; a += 32) {
        float ox = 0.0f, oy = 0.0f, oz = 0.0f;
        double tx = 0.0, ty = 0.0, tz = 0.0;      // the FP64 operand sqrt(w) (x - c)
        if (a < A) {
            const double w = wnorm[a];
            const double dx = (double)x[3 * a + 0] - cx, dy = (double)x[3 * a + 1] - cy,
                         dz = (double)x[3 * a + 2] - cz;
            g += w * (dx * dx + dy * dy + dz * dz);
            const double s = wnorm[A + a];               // sqrt(w), precomputed (abi.cu upload_weights)
            tx = s * dx; ty = s * dy; tz = s * dz;
            ox = (float)tx; oy = (float)ty; oz = (float)tz;
            t00 += tx * tx; t01 += tx * ty; t02 += tx * tz; t11 += ty * ty; t12 += ty * tz; t22 += tz * tz;
        }
        if (planes) { px[a] = ox; py[a] = oy; pz[a] = oz; }
        if (hi) {  // 3xTF32 operand split for the tensor-core sweep: x ~= hi + lo, both exact TF32 values
            const float hx = tf32_rn(ox), hy = tf32_rn(oy), hz = tf32_rn(oz);
            hi[pbase + a] = hx; hi[pbase + A_pad + a] = hy; hi[pbase + 2 * A_pad + a] = hz;
            lo[pbase + a] = tf32_rn(ox - hx); lo[pbase + A_pad + a] = tf32_rn(oy - hy);
            lo[pbase + 2 * A_pad + a] = tf32_rn(oz - hz);
        }
        if (bh) {  // 3xBF16 split: x ~= bh + bm with a 2^-18 relative residual
            const float o[3] = {ox, oy, oz};
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const __nv_bfloat16 h = __float2bfloat16_rn(o[d]);
                bh[pbase + d * A_pad + a] = h;
                bm[pbase + d * A_pad + a] = __float2bfloat16_rn(o[d] - __bfloat162float(h));
            }
        }
        if (fh) {  // FP16 split of kRmsHalfScale * x: fh + fl carries 22 bits (fl is exact in fp32 before rounding)
            const float o[3] = {ox * kRmsHalfScale, oy * kRmsHalfScale, oz * kRmsHalfScale};
            const double t[3] = {tx, ty, tz};
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const __half h = __float2half_rn(o[d]);
                fh[pbase + d * A_pad + a] = h;
                const double v1 = (double)__half2float(h) * (1.0 / kRmsHalfScale);
                gh += v1 * v1;
                r1 += (t[d] - v1) * (t[d] - v1);
                if (fl) {   // second part: only the 3xFP16 / 2xFP16 sweeps read it
                    const __half l = __float2half_rn(o[d] - __half2float(h));
                    fl[pbase + d * A_pad + a] = l;
                    const double v2 = v1 + (double)__half2float(l) * (1.0 / kRmsHalfScale);
                    g2n += v2 * v2;
                    r2 += (t[d] - v2) * (t[d] - v2);
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        g += __shfl_xor_sync(0xffffffffu, g, o);
        gh += __shfl_xor_sync(0xffffffffu, gh, o);
        g2n += __shfl_xor_sync(0xffffffffu, g2n, o);
        r1 += __shfl_xor_sync(0xffffffffu, r1, o);
        r2 += __shfl_xor_sync(0xffffffffu, r2, o);
        t00 += __shfl_xor_sync(0xffffffffu, t00, o); t01 += __shfl_xor_sync(0xffffffffu, t01, o);
        t02 += __shfl_xor_sync(0xffffffffu, t02, o); t11 += __shfl_xor_sync(0xffffffffu, t11, o);
        t12 += __shfl_xor_sync(0xffffffffu, t12, o); t22 += __shfl_xor_sync(0xffffffffu, t22, o);
    }
    if (lane == 0) {
        if (fh) {
            // residual norms rounded up: they are subtracted from lower bounds
            Gh[f] = (float)gh;
            gres[f].x = __double2float_ru(sqrt(r1));
            if (fl) { G2[f] = (float)g2n; gres[f].y = __double2float_ru(sqrt(r2)); }
        }
        if (sig) {
            // singular values of the 3 x A frame matrix = sqrt of the eigenvalues of its gyration tensor (closed form
            // for a symmetric 3x3, FP64).  min-RMSD^2(x, y) >= sum_i (sigma_i(x) - sigma_i(y))^2 by von Neumann's
            // trace inequality: the sweep's cheapest test, applied before the accumulators are even read.
            const double q = (t00 + t11 + t22) / 3.0;
            const double p1 = t01 * t01 + t02 * t02 + t12 * t12;
            const double p2 = (t00 - q) * (t00 - q) + (t11 - q) * (t11 - q) + (t22 - q) * (t22 - q) + 2.0 * p1;
            double e1 = t00, e2 = t11, e3 = t22;
            if (p2 > 0.0) {
                const double p = sqrt(p2 / 6.0), ip = 1.0 / p;
                const double b00 = (t00 - q) * ip, b11 = (t11 - q) * ip, b22 = (t22 - q) * ip;
                const double b01 = t01 * ip, b02 = t02 * ip, b12 = t12 * ip;
                double r = 0.5 * (b00 * (b11 * b22 - b12 * b12) - b01 * (b01 * b22 - b12 * b02) + b02 * (b01 * b12 - b11 * b02));
                r = fmin(1.0, fmax(-1.0, r));
                const double phi = acos(r) / 3.0;
                e1 = q + 2.0 * p * cos(phi);
                e3 = q + 2.0 * p * cos(phi + 2.0943951023931953);
                e2 = 3.0 * q - e1 - e3;
            }
            // descending order (the closed form gives e1 >= e2 >= e3; the diagonal case may not)
            if (e1 < e2) { const double t = e1; e1 = e2; e2 = t; }
            if (e2 < e3) { const double t = e2; e2 = e3; e3 = t; }
            if (e1 < e2) { const double t = e1; e1 = e2; e2 = t; }
            sig[f] = make_float4((float)sqrt(fmax(e1, 0.0)), (float)sqrt(fmax(e2, 0.0)), (float)sqrt(fmax(e3, 0.0)), 0.0f);
        }
        G[f] = (float)g;
        cen[4 * f + 0] = cx; cen[4 * f + 1] = cy; cen[4 * f + 2] = cz; cen[4 * f + 3] = g;
    }
}

cudaError_t launch_pack_frames(const float *raw, const double *wnorm, long long n, int A, int A_pad, float *planes,
                               float *hi, float *lo, void *bh, void *bm, void *fh, void *fl, float *G, double *cen, float *Gh, float *G2,
                               float *gres, float *sig, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    const int warps = 8;
    const unsigned grid = (unsigned)((n + warps - 1) / warps);
    pack_frames_kernel<<<grid, warps * 32, 0, st>>>(raw, wnorm, n, A, A_pad, planes, hi, lo, static_cast<__nv_bfloat16 *>(bh),
                                                    static_cast<__nv_bfloat16 *>(bm), static_cast<__half *>(fh),
                                                    static_cast<__half *>(fl), G, cen, Gh, G2,
                                                    reinterpret_cast<float2 *>(gres), reinterpret_cast<float4 *>(sig));
    return cudaGetLastError();
}

__global__ void max_float_kernel(const float *v, long long n, int stride, float *out)
{
    float m = 0.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmaxf(m, v[i * stride]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int *>(out), __float_as_int(m));  // m >= 0
}

__global__ void fill_u32_kernel(uint32_t *p, size_t n, uint32_t v)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void iota_i32_kernel(int *p, int n, int start, int stride)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = start + i * stride;
}

cudaError_t launch_iota_i32(int *p, int n, int start, int stride, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    iota_i32_kernel<<<148, 256, 0, st>>>(p, n, start, stride);
    return cudaGetLastError();
}

// Audit of certified rows against the exact path: counts rows whose neighbour indices differ or whose distances
// differ by more than 1e-12 relative (the two FP64 paths sum the cross-covariance in different orders).
__global__ void audit_compare_kernel(const double *out_dist, const int *out_idx, const int *row_ids, const double *ex_dist,
                                     const int *ex_idx, int n_rows, int k1, int *mismatches)
{
    const int r = blockIdx.x;
    if (r >= n_rows) return;
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    const size_t o = (size_t)row_ids[r] * k1, e = (size_t)r * k1;
    for (int j = threadIdx.x; j < k1; j += blockDim.x) {
        const double a = out_dist[o + j], b = ex_dist[e + j];
        if (out_idx[o + j] != ex_idx[e + j] || !(fabs(a - b) <= 1e-12 * fmax(fabs(b), 1e-300) + 1e-300)) bad = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0 && bad) atomicAdd(mismatches, 1);
}

cudaError_t launch_audit_compare(const double *out_dist, const int *out_idx, const int *row_ids, const double *ex_dist,
                                 const int *ex_idx, int n_rows, int k1, int *mismatches, cudaStream_t st)
{
    if (n_rows <= 0) return cudaSuccess;
    audit_compare_kernel<<<n_rows, 128, 0, st>>>(out_dist, out_idx, row_ids, ex_dist, ex_idx, n_rows, k1, mismatches);
    return cudaGetLastError();
}

cudaError_t launch_fill_u32(void *p, size_t n, uint32_t v, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    fill_u32_kernel<<<148, 256, 0, st>>>(static_cast<uint32_t *>(p), n, v);
    return cudaGetLastError();
}

cudaError_t launch_max_float_strided(const float *v, long long n, int stride, float *out, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float), st);
    if (e != cudaSuccess) return e;
    if (n <= 0) return cudaSuccess;
    max_float_kernel<<<148, 256, 0, st>>>(v, n, stride, out);
    return cudaGetLastError();
}

cudaError_t launch_max_float(const float *v, long long n, float *out, cudaStream_t st)
{
    return launch_max_float_strided(v, n, 1, out, st);
}

}  // namespace mdsctk
