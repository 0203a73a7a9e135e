// qcp.cuh -- per-pair rotation: Theobald's quaternion characteristic polynomial (QCP).
//
// Replaces GROMACS do_fit (calc_fit_R + 6x6 Jacobi + rotate) followed by rmsdev, as called
// from distance() at knn_rms.cpp:38-41.  For centred, sqrt(m/M)-scaled frames x, y with
// S_ab = sum_n x_na y_nb, G = sum |x|^2 and E0 = (Gx+Gy)/2, the minimum over proper
// rotations is  RMSD^2 = 2 (E0 - lambda_max(K))  where K is the 4x4 traceless key matrix of
// S.  lambda_max is the largest root of
//     P(l) = l^4 + C2 l^2 + C1 l + C0,   C2 = -2 |S|_F^2,  C1 = -8 det S,  C0 = det K = |S|_F^4 - 4 |cof S|_F^2,
// found by Newton from l0 = E0 >= lambda_max.  All roots are real, so the iterates decrease
// monotonically onto lambda_max and EVERY iterate gives a valid lower bound on RMSD^2 --
// a pair is rejected as soon as that bound exceeds the row's admission threshold.
// No rotation matrix is ever formed.
#pragma once
#include <cuda_runtime.h>

namespace mdsctk {

struct QcpCoef {
    float c2, c1, c0;
};

// |S|_F^2 of pair j of a batch (sv[3a+b][j] = S_ab).  lambda_max <= sqrt(3 |S|_F^2) (the four
// eigenvalues of K sum to 0 and their squares to 4 |S|_F^2), the cheapest lower bound on RMSD^2.
template <int B>
__device__ __forceinline__ float qcp_frob2(const float (&sv)[9][B], int j)
{
    float f = sv[0][j] * sv[0][j];
#pragma unroll
    for (int c = 1; c < 9; ++c) f = fmaf(sv[c][j], sv[c][j], f);
    return f;
}

// s[3*a+b] = S_ab, f = |S|_F^2
__device__ __forceinline__ QcpCoef qcp_coefficients(const float *s, float f)
{
    const float sxx = s[0], sxy = s[1], sxz = s[2];
    const float syx = s[3], syy = s[4], syz = s[5];
    const float szx = s[6], szy = s[7], szz = s[8];
    QcpCoef c;
    c.c2 = -2.0f * f;
    // det S by cofactors of the first row
    const float m0 = fmaf(syy, szz, -syz * szy);
    const float m1 = fmaf(syx, szz, -syz * szx);
    const float m2 = fmaf(syx, szy, -syy * szx);
    const float det = fmaf(sxx, m0, fmaf(-sxy, m1, sxz * m2));
    c.c1 = -8.0f * det;
    // det K = |S|_F^4 - 4 |cof S|_F^2: the eigenvalues of K are (s1+s2+s3), (s1-s2-s3), (-s1+s2-s3), (-s1-s2+s3)
    // in the (signed) singular values of S, and their product is  sum s_i^4 - 2 sum s_i^2 s_j^2
    // = (sum s_i^2)^2 - 4 sum_{i<j} s_i^2 s_j^2, where sum_{i<j} s_i^2 s_j^2 is the squared Frobenius norm of
    // the cofactor matrix.  29 operations instead of the ~60 of the 2x2-minor expansion of the 4x4
    // determinant (the identity is checked numerically in the CPU test suite: test_qcp_c0_identity).
    const float m3 = fmaf(sxy, szz, -sxz * szy);
    const float m4 = fmaf(sxx, szz, -sxz * szx);
    const float m5 = fmaf(sxx, szy, -sxy * szx);
    const float m6 = fmaf(sxy, syz, -sxz * syy);
    const float m7 = fmaf(sxx, syz, -sxz * syx);
    const float m8 = fmaf(sxx, syy, -sxy * syx);
    float cc = m0 * m0;
    cc = fmaf(m1, m1, cc); cc = fmaf(m2, m2, cc);
    cc = fmaf(m3, m3, cc); cc = fmaf(m4, m4, cc); cc = fmaf(m5, m5, cc);
    cc = fmaf(m6, m6, cc); cc = fmaf(m7, m7, cc); cc = fmaf(m8, m8, cc);
    c.c0 = fmaf(f, f, -4.0f * cc);
    return c;
}

__device__ __forceinline__ QcpCoef qcp_coefficients(const float *s)
{
    float f = s[0] * s[0];
#pragma unroll
    for (int c = 1; c < 9; ++c) f = fmaf(s[c], s[c], f);
    return qcp_coefficients(s, f);
}

// One Newton step on P from x; returns the new iterate.
__device__ __forceinline__ float qcp_newton_step(const QcpCoef &c, float x)
{
    const float x2 = x * x;
    const float b = (x2 + c.c2) * x;
    const float a = b + c.c1;
    const float p = fmaf(a, x, c.c0);
    const float dp = fmaf(2.0f * x2, x, b + a);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(dp));
    return fmaf(-p, r, x);
}

// Approximate min-RMSD^2 (nm^2) of one pair, or +inf as soon as a Newton iterate proves
// RMSD^2 > tau.  e0 = (Gq+Gr)/2 with weights normalised to sum 1.
__device__ __forceinline__ float qcp_msd_bounded(const float *s, float e0, float tau, int do_fit)
{
    const float kInf = __uint_as_float(0x7f800000u);
    if (!do_fit) {  // --nofit: lambda = trace S
        const float d2 = fmaxf(2.0f * (e0 - (s[0] + s[4] + s[8])), 0.0f);
        return d2 < tau ? d2 : kInf;
    }
    const QcpCoef c = qcp_coefficients(s);
    float x = e0;
    float xn = qcp_newton_step(c, x);
    if (!(xn == xn)) xn = x;  // 0/0 when both frames are degenerate
    if (2.0f * (e0 - xn) > tau) return kInf;
#pragma unroll 1
    for (int it = 0; it < 24; ++it) {
        if (fabsf(xn - x) <= 4e-7f * fabsf(xn)) break;
        x = xn;
        xn = qcp_newton_step(c, x);
        if (!(xn == xn)) { xn = x; break; }
        if (2.0f * (e0 - xn) > tau) return kInf;
    }
    const float d2 = fmaxf(2.0f * (e0 - xn), 0.0f);
    return d2 < tau ? d2 : kInf;
}

// Newton refinement to convergence from a first iterate xn (previous iterate x); +inf as soon as
// a bound exceeds tau.  Rarely executed: only pairs whose first bound is below the threshold.
static __device__ __noinline__ float qcp_refine(QcpCoef c, float e0, float x, float xn, float tau)
{
    const float kInf = __uint_as_float(0x7f800000u);
#pragma unroll 1
    for (int it = 0; it < 24; ++it) {
        if (fabsf(xn - x) <= 4e-7f * fabsf(xn)) break;
        x = xn;
        xn = qcp_newton_step(c, x);
        if (!(xn == xn)) { xn = x; break; }
        if (2.0f * (e0 - xn) > tau) return kInf;
    }
    const float d2 = fmaxf(2.0f * (e0 - xn), 0.0f);
    return d2 < tau ? d2 : kInf;
}

// Batch form for the tensor-core epilogue: B independent pairs in straight-line code so the
// coefficient and first-Newton chains of different pairs interleave (ILP).  sv[3a+b][j] = S_ab of pair j.
template <int B>
__device__ __forceinline__ void qcp_msd_batch(const float (&sv)[9][B], const float (&e0)[B], float tau, int do_fit,
                                              float (&d2)[B])
{
    const float kInf = __uint_as_float(0x7f800000u);
    if (!do_fit) {
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const float v = fmaxf(2.0f * (e0[j] - (sv[0][j] + sv[4][j] + sv[8][j])), 0.0f);
            d2[j] = v < tau ? v : kInf;
        }
        return;
    }
    QcpCoef c[B];
    float x1[B];
#pragma unroll
    for (int j = 0; j < B; ++j) {
        const float s9[9] = {sv[0][j], sv[1][j], sv[2][j], sv[3][j], sv[4][j], sv[5][j], sv[6][j], sv[7][j], sv[8][j]};
        c[j] = qcp_coefficients(s9);
    }
#pragma unroll
    for (int j = 0; j < B; ++j) {
        float xn = qcp_newton_step(c[j], e0[j]);
        x1[j] = (xn == xn) ? xn : e0[j];
    }
#pragma unroll
    for (int j = 0; j < B; ++j) {
        if (2.0f * (e0[j] - x1[j]) > tau) d2[j] = kInf;
        else d2[j] = qcp_refine(c[j], e0[j], e0[j], x1[j], tau);
    }
}

}  // namespace mdsctk
