// featurize.cu -- backbone phi/psi angles and their sin/cos embedding on the GPU: the producers of
// knn_data's input in the reference workflow (examples/cluster_phipsi.bash).
//
// Replaces the frame loops of bb_xtc_to_phipsi.cpp:106-122 (torsion(), mdsctk.cpp:643-676, over the
// N-CA-C backbone atoms of every frame: 2*(natoms/3)-2 angles per frame, float arithmetic widened to
// double on output) and angles_to_sincos.cpp:107-118 (each angle a -> sin a, cos a in double).
// HBM-bound: 12 bytes read per atom, 8 (angles) + 16 (sin/cos) bytes written per angle; one thread per
// (frame, angle).  The float operations are kept unfused and in the reference's order
// (__fmul_rn / __fsub_rn ...); acos is evaluated in double and rounded to float (correctly rounded).
#include "common.cuh"

namespace mdsctk {

__device__ __forceinline__ void crossprod_f(float (&c)[3], float x1, float y1, float z1, float x2, float y2, float z2)
{
    c[0] = __fsub_rn(__fmul_rn(y1, z2), __fmul_rn(z1, y2));      // mdsctk.cpp:634-641
    c[1] = __fsub_rn(__fmul_rn(z1, x2), __fmul_rn(x1, z2));
    c[2] = __fsub_rn(__fmul_rn(x1, y2), __fmul_rn(y1, x2));
}

__device__ __forceinline__ float dot3_f(const float (&a)[3], float bx, float by, float bz)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(a[0], bx), __fmul_rn(a[1], by)), __fmul_rn(a[2], bz));
}

// torsion(pos1, pos2, pos3, pos4, false), mdsctk.cpp:643-676
__device__ float torsion_f(const float *p1, const float *p2, const float *p3, const float *p4)
{
    float L[3], R[3], S[3];
    const float b1x = __fsub_rn(p2[0], p1[0]), b1y = __fsub_rn(p2[1], p1[1]), b1z = __fsub_rn(p2[2], p1[2]);
    const float b2x = __fsub_rn(p3[0], p2[0]), b2y = __fsub_rn(p3[1], p2[1]), b2z = __fsub_rn(p3[2], p2[2]);
    crossprod_f(L, b1x, b1y, b1z, b2x, b2y, b2z);
    crossprod_f(R, __fsub_rn(p4[0], p3[0]), __fsub_rn(p4[1], p3[1]), __fsub_rn(p4[2], p3[2]),
                __fsub_rn(p2[0], p3[0]), __fsub_rn(p2[1], p3[1]), __fsub_rn(p2[2], p3[2]));
    const float Lnorm = __fsqrt_rn(dot3_f(L, L[0], L[1], L[2]));
    const float Rnorm = __fsqrt_rn(dot3_f(R, R[0], R[1], R[2]));
    crossprod_f(S, L[0], L[1], L[2], R[0], R[1], R[2]);
    float angle = __fdiv_rn(dot3_f(L, R[0], R[1], R[2]), __fmul_rn(Lnorm, Rnorm));
    if (angle > 1.0f) angle = 1.0f;
    if (angle < -1.0f) angle = -1.0f;
    angle = (float)acos((double)angle);
    if (dot3_f(S, b2x, b2y, b2z) < 0.0f) angle = -angle;
    return angle;
}

// xyz [n][A][3] float (nm); phipsi [n][T] double, sincos [n][2T] double (either may be NULL), T = 2*(A/3)-2.
__global__ void phipsi_kernel(const float *__restrict__ xyz, long long n, int A, int T, double *__restrict__ phipsi,
                              double *__restrict__ sincos)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n * T) return;
    const long long f = g / T;
    const int t = (int)(g - f * T);
    // angle 2r uses atoms 3r..3r+3 (psi of residue r), angle 2r+1 atoms 3r+2..3r+5 (phi of residue r+1)
    const int x = 3 * (t >> 1) + ((t & 1) ? 2 : 0);
    const float *p = xyz + ((size_t)f * A + x) * 3;
    const double a = (double)torsion_f(p, p + 3, p + 6, p + 9);
    if (phipsi) phipsi[g] = a;
    if (sincos) { sincos[2 * g] = sin(a); sincos[2 * g + 1] = cos(a); }     // angles_to_sincos.cpp:109-110
}

__global__ void sincos_kernel(const double *__restrict__ angles, long long n, double *__restrict__ out)
{
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const double a = angles[g];
    out[2 * g] = sin(a);
    out[2 * g + 1] = cos(a);
}

cudaError_t launch_phipsi(const float *xyz, long long n, int A, double *phipsi, double *sincos, cudaStream_t st)
{
    const int T = 2 * (A / 3) - 2;
    if (n <= 0 || T <= 0) return cudaSuccess;
    const long long total = n * T;
    phipsi_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(xyz, n, A, T, phipsi, sincos);
    return cudaGetLastError();
}

cudaError_t launch_sincos(const double *angles, long long n, double *out, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    sincos_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(angles, n, out);
    return cudaGetLastError();
}

}  // namespace mdsctk
