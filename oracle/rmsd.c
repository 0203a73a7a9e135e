/*
 * rmsd.c -- ORACLE (test infrastructure, see oracle.h): per-frame and per-pair
 * arithmetic of the kNN stage.
 *
 * The reference's distance() (knn_rms.cpp:38-41) is
 *     if (dofit) do_fit(natoms, weights, reference, fitting);
 *     return (double) rmsdev(natoms, weights, reference, fitting) * 10.0;
 * with do_fit / rmsdev / reset_x from GROMACS 5.0-5.1 (mixed precision,
 * real = float; not vendored).  The three functions are restated below from
 * the published algorithm: float 3x3 cross matrix,
 * McLachlan's symmetric 6x6 eigenproblem solved by cyclic Jacobi in double,
 * third axis by cross product (proper rotation), float rotation in place,
 * float mass-weighted deviation.  oracle_rmsd_f64 is the independent FP64
 * statement of the same minimum (Kabsch / quaternion key matrix), used to
 * bound the float chain's noise and as the index-exact target.
 *
 * euclidean_distance / correlation_distance follow mdsctk.cpp:330-335 and
 * :337-360 operation by operation (build with -ffp-contract=off so the
 * sequential sums are not fused).
 */
#include "oracle.h"
#include <math.h>
#include <string.h>

/* ---- cyclic Jacobi for a small dense symmetric matrix (double) ------------ */
#define JMAX 6
static void jacobi_sym(int n, double a[JMAX][JMAX], double eval[JMAX], double evec[JMAX][JMAX])
{
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < n; j++) evec[i][j] = (i == j);
    }
    for (int sweep = 0; sweep < 64; sweep++) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < n; i++) {
            diag += a[i][i] * a[i][i];
            for (int j = i + 1; j < n; j++) off += a[i][j] * a[i][j];
        }
        if (off <= 1e-32 * (diag + off) || off == 0.0) break;
        for (int p = 0; p < n - 1; p++) {
            for (int q = p + 1; q < n; q++) {
                double apq = a[p][q];
                if (apq == 0.0) continue;
                /* an off-diagonal element below the rounding unit of both diagonals is dropped */
                double g = 100.0 * fabs(apq);
                if (sweep > 2 && fabs(a[p][p]) + g == fabs(a[p][p]) && fabs(a[q][q]) + g == fabs(a[q][q])) {
                    a[p][q] = a[q][p] = 0.0;
                    continue;
                }
                double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                /* A <- J^T A J, touching only rows/columns p and q of the symmetric matrix */
                a[p][p] -= t * apq;
                a[q][q] += t * apq;
                a[p][q] = a[q][p] = 0.0;
                for (int k = 0; k < n; k++) {
                    if (k == p || k == q) continue;
                    double akp = a[k][p], akq = a[k][q];
                    a[k][p] = a[p][k] = c * akp - s * akq;
                    a[k][q] = a[q][k] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {
                    double vkp = evec[k][p], vkq = evec[k][q];
                    evec[k][p] = c * vkp - s * vkq;
                    evec[k][q] = s * vkp + c * vkq;
                }
            }
        }
    }
    for (int i = 0; i < n; i++) eval[i] = a[i][i];
}

/* ---- reset_x(natoms, NULL, natoms, NULL, x, mass) -- knn_rms.cpp:189,199 --- */
void oracle_reset_x(int natoms, float *x, const float *mass)
{
    float xcm[3] = {0.0f, 0.0f, 0.0f};
    float tm = 0.0f;
    for (int i = 0; i < natoms; i++) {
        float mm = mass[i];
        for (int d = 0; d < 3; d++) xcm[d] += mm * x[3 * i + d];
        tm += mm;
    }
    for (int d = 0; d < 3; d++) xcm[d] /= tm;
    for (int i = 0; i < natoms; i++)
        for (int d = 0; d < 3; d++) x[3 * i + d] -= xcm[d];
}

/* ---- do_fit(natoms, w, xp, x): rotate x onto xp in place -- knn_rms.cpp:39 -- */
void oracle_do_fit(int natoms, const float *w, const float *xp, float *x)
{
    float u[3][3];
    memset(u, 0, sizeof u);
    for (int n = 0; n < natoms; n++) {
        float mn = w[n];
        if (mn == 0.0f) continue;
        for (int c = 0; c < 3; c++) {
            double xpc = xp[3 * n + c];
            for (int r = 0; r < 3; r++) {
                double xnr = x[3 * n + r];
                u[c][r] += (float)(mn * xnr * xpc);
            }
        }
    }
    double omega[JMAX][JMAX], om[JMAX][JMAX], d[JMAX];
    memset(omega, 0, sizeof omega);
    for (int r = 3; r < 6; r++)
        for (int c = 0; c < 3; c++) {
            omega[r][c] = u[r - 3][c];
            omega[c][r] = u[r - 3][c];
        }
    jacobi_sym(6, omega, d, om);

    float vh[3][3], vk[3][3];
    for (int j = 0; j < 2; j++) {
        int index = 0;
        double max_d = -1000;
        for (int i = 0; i < 6; i++)
            if (d[i] > max_d) { max_d = d[i]; index = i; }
        d[index] = -10000;
        for (int i = 0; i < 3; i++) {
            vh[j][i] = (float)(M_SQRT2 * om[i][index]);
            vk[j][i] = (float)(M_SQRT2 * om[i + 3][index]);
        }
    }
    /* third axes as cross products: never a mirror image, safe for flat structures */
    vh[2][0] = vh[0][1] * vh[1][2] - vh[0][2] * vh[1][1];
    vh[2][1] = vh[0][2] * vh[1][0] - vh[0][0] * vh[1][2];
    vh[2][2] = vh[0][0] * vh[1][1] - vh[0][1] * vh[1][0];
    vk[2][0] = vk[0][1] * vk[1][2] - vk[0][2] * vk[1][1];
    vk[2][1] = vk[0][2] * vk[1][0] - vk[0][0] * vk[1][2];
    vk[2][2] = vk[0][0] * vk[1][1] - vk[0][1] * vk[1][0];

    float R[3][3];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            float acc = 0.0f;
            for (int s = 0; s < 3; s++) acc += vk[s][r] * vh[s][c];
            R[r][c] = acc;
        }
    for (int j = 0; j < natoms; j++) {
        float old[3] = {x[3 * j], x[3 * j + 1], x[3 * j + 2]};
        for (int r = 0; r < 3; r++) {
            float acc = 0.0f;
            for (int c = 0; c < 3; c++) acc += R[r][c] * old[c];
            x[3 * j + r] = acc;
        }
    }
}

/* ---- rmsdev(natoms, mass, x, xp) -- knn_rms.cpp:40 -------------------------- */
float oracle_rmsdev(int natoms, const float *mass, const float *x, const float *xp)
{
    float tm = 0.0f, rd = 0.0f;
    for (int i = 0; i < natoms; i++) {
        float mm = mass[i];
        tm += mm;
        for (int d = 0; d < 3; d++) {
            float xd = x[3 * i + d] - xp[3 * i + d];
            rd += mm * xd * xd;
        }
    }
    return (float)sqrt(rd / tm);
}

/* ---- independent FP64 statement of the same minimum ------------------------- */
double oracle_rmsd_f64(int natoms, const float *mass, const float *a, const float *b, int dofit)
{
    double ca[3] = {0, 0, 0}, cb[3] = {0, 0, 0}, M = 0.0;
    for (int i = 0; i < natoms; i++) {
        double m = mass[i];
        M += m;
        for (int d = 0; d < 3; d++) {
            ca[d] += m * (double)a[3 * i + d];
            cb[d] += m * (double)b[3 * i + d];
        }
    }
    for (int d = 0; d < 3; d++) { ca[d] /= M; cb[d] /= M; }
    double S[3][3] = {{0}}, Ga = 0.0, Gb = 0.0;
    for (int i = 0; i < natoms; i++) {
        double m = mass[i], x[3], y[3];
        for (int d = 0; d < 3; d++) {
            x[d] = (double)a[3 * i + d] - ca[d];
            y[d] = (double)b[3 * i + d] - cb[d];
            Ga += m * x[d] * x[d];
            Gb += m * y[d] * y[d];
        }
        for (int p = 0; p < 3; p++)
            for (int q = 0; q < 3; q++) S[p][q] += m * x[p] * y[q];
    }
    double lambda;
    if (!dofit) {
        lambda = S[0][0] + S[1][1] + S[2][2];
    } else {
        double K[JMAX][JMAX], ev[JMAX], vec[JMAX][JMAX];
        memset(K, 0, sizeof K);
        K[0][0] = S[0][0] + S[1][1] + S[2][2];
        K[1][1] = S[0][0] - S[1][1] - S[2][2];
        K[2][2] = -S[0][0] + S[1][1] - S[2][2];
        K[3][3] = -S[0][0] - S[1][1] + S[2][2];
        K[0][1] = K[1][0] = S[1][2] - S[2][1];
        K[0][2] = K[2][0] = S[2][0] - S[0][2];
        K[0][3] = K[3][0] = S[0][1] - S[1][0];
        K[1][2] = K[2][1] = S[0][1] + S[1][0];
        K[1][3] = K[3][1] = S[2][0] + S[0][2];
        K[2][3] = K[3][2] = S[1][2] + S[2][1];
        jacobi_sym(4, K, ev, vec);
        lambda = ev[0];
        for (int i = 1; i < 4; i++)
            if (ev[i] > lambda) lambda = ev[i];
    }
    double msd = (Ga + Gb - 2.0 * lambda) / M;
    return msd > 0.0 ? sqrt(msd) : 0.0;
}

/* ---- mdsctk.cpp:330-335 ------------------------------------------------------ */
double oracle_euclidean_distance(int size, const double *reference, const double *fitting)
{
    double value = 0.0;
    for (int x = 0; x < size; x++) value += (reference[x] - fitting[x]) * (reference[x] - fitting[x]);
    return sqrt(value);
}

/* ---- mdsctk.cpp:337-360 ------------------------------------------------------ */
double oracle_correlation_distance(int size, const double *reference, const double *fitting)
{
    double rsum = 0.0, rsq = 0.0, fsum = 0.0, fsq = 0.0, acc = 0.0;
    double n = (double)size;
    for (int x = 0; x < size; x++) {
        rsum += reference[x];
        rsq += reference[x] * reference[x];
        fsum += fitting[x];
        fsq += fitting[x] * fitting[x];
    }
    rsq = sqrt(((n * rsq) - (rsum * rsum)) / (n * (n - 1.0)));
    rsum /= n;
    fsq = sqrt(((n * fsq) - (fsum * fsum)) / (n * (n - 1.0)));
    fsum /= n;
    for (int x = 0; x < size; x++) acc += (reference[x] - rsum) * (fitting[x] - fsum);
    acc = (1.0 - (acc / ((n - 1.0) * rsq * fsq))) / 2.0;
    if (acc < 0.0) acc = 0.0;
    return sqrt(acc);
}
