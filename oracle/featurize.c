/* featurize.c -- CPU restatement of bb_xtc_to_phipsi / angles_to_sincos (TEST INFRASTRUCTURE, see oracle.h).
 *
 * torsion() follows mdsctk.cpp:634-676 operation for operation in float (`::real`), with the float
 * overloads of sqrt / acos the C++ reference resolves to; the frame loop follows
 * bb_xtc_to_phipsi.cpp:108-119 (angle 2r over atoms 3r..3r+3, angle 2r+1 over atoms 3r+2..3r+5) and
 * the embedding angles_to_sincos.cpp:107-118 (sin, cos in double, interleaved).
 * Built with -ffp-contract=off (oracle/Makefile): the reference's default build has no FMA either. */
#include "oracle.h"

#include <math.h>
#include <stddef.h>

static void crossprod(float C[3], float x1, float y1, float z1, float x2, float y2, float z2)
{
    C[0] = ((y1 * z2) - (z1 * y2));
    C[1] = ((z1 * x2) - (x1 * z2));
    C[2] = ((x1 * y2) - (y1 * x2));
}

float oracle_torsion(const float *pos1, const float *pos2, const float *pos3, const float *pos4)
{
    float L[3], R[3], S[3], Lnorm, Rnorm, angle;
    crossprod(L, (pos2[0] - pos1[0]), (pos2[1] - pos1[1]), (pos2[2] - pos1[2]),
              (pos3[0] - pos2[0]), (pos3[1] - pos2[1]), (pos3[2] - pos2[2]));
    crossprod(R, (pos4[0] - pos3[0]), (pos4[1] - pos3[1]), (pos4[2] - pos3[2]),
              (pos2[0] - pos3[0]), (pos2[1] - pos3[1]), (pos2[2] - pos3[2]));
    Lnorm = sqrtf(L[0] * L[0] + L[1] * L[1] + L[2] * L[2]);
    Rnorm = sqrtf(R[0] * R[0] + R[1] * R[1] + R[2] * R[2]);
    crossprod(S, L[0], L[1], L[2], R[0], R[1], R[2]);
    angle = (L[0] * R[0] + L[1] * R[1] + L[2] * R[2]) / (Lnorm * Rnorm);
    if (angle > 1.0) angle = 1.0;
    if (angle < -1.0) angle = -1.0;
    angle = acosf(angle);
    if ((S[0] * (pos3[0] - pos2[0]) + S[1] * (pos3[1] - pos2[1]) + S[2] * (pos3[2] - pos2[2])) < 0) angle = -angle;
    return angle;
}

/* xyz [n][natoms][3]; phipsi [n][2*(natoms/3)-2] */
void oracle_phipsi(const float *xyz, long long n, int natoms, double *phipsi)
{
    const int T = 2 * (natoms / 3) - 2;
    for (long long f = 0; f < n; f++) {
        const float *c = xyz + (size_t)f * natoms * 3;
        int i_mat = 0;
        for (int x = 0; x < natoms - 3;) {
            phipsi[f * T + i_mat++] = (double)oracle_torsion(c + 3 * x, c + 3 * (x + 1), c + 3 * (x + 2), c + 3 * (x + 3));
            x += 2;
            phipsi[f * T + i_mat++] = (double)oracle_torsion(c + 3 * x, c + 3 * (x + 1), c + 3 * (x + 2), c + 3 * (x + 3));
            x += 1;
        }
    }
}

void oracle_sincos(const double *angles, long long n, double *out)
{
    for (long long y = 0; y < n; y++) {
        out[2 * y] = sin(angles[y]);
        out[(2 * y) + 1] = cos(angles[y]);
    }
}
