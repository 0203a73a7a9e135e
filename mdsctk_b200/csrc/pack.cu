// pack.cu -- frame loading on the GPU: centre, mass-scale and pack decoded frames.
//
// Replaces the per-frame reset_x(natoms,NULL,natoms,NULL,x,weights) of the load loops at
// knn_rms.cpp:186-206 (GROMACS: float centre-of-mass removal) and sets up the operand
// layout of the contraction: AoS float[n][A][3] (nm) -> planes float[n][3][A_pad] holding
// sqrt(m_a/M) * (x_a - c), G = sum (m_a/M)|x_a - c|^2, and the FP64 centroid for the re-score.
// One warp per frame; HBM-bound: 12*A bytes read + 12*A_pad written per frame.
#include "common.cuh"

namespace mdsctk {

__global__ void __launch_bounds__(256) pack_frames_kernel(const float *__restrict__ raw,
                                                          const double *__restrict__ wnorm,  // m_a / M (FP64)
                                                          long long n, int A, int A_pad,
                                                          float *__restrict__ planes, float *__restrict__ G,
                                                          double *__restrict__ cen)
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
    float *px = planes + (size_t)f * 3 * A_pad;
    float *py = px + A_pad, *pz = py + A_pad;
    double g = 0.0;
    for (int a = lane; a < A_pad; a += 32) {
        float ox = 0.0f, oy = 0.0f, oz = 0.0f;
        if (a < A) {
            const double w = wnorm[a];
            const double dx = (double)x[3 * a + 0] - cx, dy = (double)x[3 * a + 1] - cy,
                         dz = (double)x[3 * a + 2] - cz;
            g += w * (dx * dx + dy * dy + dz * dz);
            const double s = sqrt(w);
            ox = (float)(s * dx); oy = (float)(s * dy); oz = (float)(s * dz);
        }
        px[a] = ox; py[a] = oy; pz[a] = oz;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
    if (lane == 0) {
        G[f] = (float)g;
        cen[4 * f + 0] = cx; cen[4 * f + 1] = cy; cen[4 * f + 2] = cz; cen[4 * f + 3] = g;
    }
}

cudaError_t launch_pack_frames(const float *raw, const double *wnorm, long long n, int A, int A_pad, float *planes,
                               float *G, double *cen, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    const int warps = 8;
    const unsigned grid = (unsigned)((n + warps - 1) / warps);
    pack_frames_kernel<<<grid, warps * 32, 0, st>>>(raw, wnorm, n, A, A_pad, planes, G, cen);
    return cudaGetLastError();
}

__global__ void max_float_kernel(const float *v, long long n, float *out)
{
    float m = 0.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmaxf(m, v[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int *>(out), __float_as_int(m));  // m >= 0
}

cudaError_t launch_max_float(const float *v, long long n, float *out, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float), st);
    if (e != cudaSuccess) return e;
    if (n <= 0) return cudaSuccess;
    max_float_kernel<<<148, 256, 0, st>>>(v, n, out);
    return cudaGetLastError();
}

}  // namespace mdsctk
