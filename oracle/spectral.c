/* spectral.c -- CPU restatement of auto_decomp_sparse (TEST INFRASTRUCTURE, see oracle.h).
 *
 * Affinity stage: auto_decomp_sparse.cpp:150-198, loop for loop (including its sigma quirk: the values of a
 * frame are collected in CSC traversal order, truncated to the FIRST k_a, and the sum is divided by k_a even
 * when fewer are present).  The -K branch (entropic affinities) is restated below, warm-start chain included.
 * Eigen-solve: the reference calls ARPACK (dsaupd/dseupd, mode 1, which = "LA", tol = machine eps, ncv =
 * 10*nev+1; mdsctk.cpp:857-924) -- a third-party Fortran library that is not vendored and not installed here
 * (CMakeLists.txt finds it with find_library, no version pin; arpack-ng 3.x API).  ARPACK's implicitly
 * restarted Lanczos converges to the nev algebraically largest eigenpairs of the symmetric matrix; the oracle
 * computes the same quantities by a dense cyclic Jacobi diagonalisation (small n only).  Eigenvector signs are
 * arbitrary in ARPACK; callers compare up to sign. */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* M (upper-triangular CSC values, distances) is replaced by the normalised affinities; returns the average sigma. */
double oracle_affinity(int n, const int *pcol, const int *irow, double *M, int k_a)
{
    double *sigma_a = (double *)calloc((size_t)n, sizeof(double));
    double *d_a = (double *)calloc((size_t)n, sizeof(double));
    int *cnt = (int *)calloc((size_t)n, sizeof(int));
    /* sorted_A[x]: the values pushed for frame x in traversal order (:156-161), truncated to the FIRST k_a (:163-164),
     * sorted ascending (:165) and added left to right (:166-168) */
    for (int x = 0; x < n; x++)
        for (int y = pcol[x]; y < pcol[x + 1]; y++) {
            if (cnt[x] < k_a) cnt[x]++;
            if (cnt[irow[y]] < k_a) cnt[irow[y]]++;
        }
    {
        double *buf = (double *)malloc(sizeof(double) * (size_t)(k_a > 0 ? k_a : 1));
        int *fill = (int *)calloc((size_t)n, sizeof(int));
        double **vals = (double **)malloc(sizeof(double *) * (size_t)n);
        for (int x = 0; x < n; x++) vals[x] = (double *)malloc(sizeof(double) * (size_t)(cnt[x] > 0 ? cnt[x] : 1));
        for (int x = 0; x < n; x++)
            for (int y = pcol[x]; y < pcol[x + 1]; y++) {
                if (fill[x] < k_a) vals[x][fill[x]++] = M[y];
                if (fill[irow[y]] < k_a) vals[irow[y]][fill[irow[y]]++] = M[y];
            }
        for (int x = 0; x < n; x++) {
            /* insertion sort ascending (std::sort result), then the left-to-right sum of :166-168 */
            for (int a = 1; a < fill[x]; a++) {
                const double v = vals[x][a];
                int b = a - 1;
                while (b >= 0 && vals[x][b] > v) { vals[x][b + 1] = vals[x][b]; b--; }
                vals[x][b + 1] = v;
            }
            sigma_a[x] = 0;
            for (int y = 0; y < fill[x]; y++) sigma_a[x] += vals[x][y];
            sigma_a[x] /= (double)k_a;
            free(vals[x]);
        }
        free(vals); free(fill); free(buf);
    }
    for (int x = 0; x < n; x++)
        for (int y = pcol[x]; y < pcol[x + 1]; y++)
            M[y] = exp(-(M[y] * M[y]) / (2.0 * sigma_a[x] * sigma_a[irow[y]]));      /* :176-178 */
    for (int x = 0; x < n; x++)
        for (int y = pcol[x]; y < pcol[x + 1]; y++) {                                 /* :183-188 */
            d_a[x] += M[y];
            d_a[irow[y]] += M[y];
        }
    for (int x = 0; x < n; x++) d_a[x] = 1.0 / sqrt(d_a[x]);
    for (int x = 0; x < n; x++)
        for (int y = pcol[x]; y < pcol[x + 1]; y++) M[y] *= d_a[irow[y]] * d_a[x];   /* :193-197 */
    double avg = 0.0;
    for (int x = 0; x < n; x++) avg += sigma_a[x];
    avg /= (double)n;
    free(sigma_a); free(d_a); free(cnt);
    return avg;
}

/* entropic_affinity_sigma, mdsctk.cpp:388-496: root of e(b) = beta m1 + log m0 - log K in b = log beta by Newton steps
 * safeguarded with bisection of the bracket [B_lower, B_upper]; A = the frame's k sorted distances. */
static double entropic_sigma(const double *A, int k, double b0, double logK, double logN, double B_lower, double B_upper)
{
    const int maxit = 20;
    const double tol = 1e-10, realmin = 2.225074e-308;
    double eps = 1.0;
    do { eps /= 2.0; } while (1.0 + (eps / 2.0) != 1.0);
    eps = sqrt(eps);
    double b = (b0 < B_lower || b0 > B_upper) ? (B_lower + B_upper) / 2.0 : b0;
    double *ed2 = (double *)malloc(sizeof(double) * (size_t)k), *m1v = (double *)malloc(sizeof(double) * (size_t)k);
    int i = 1;
    for (;;) {
        const double bE = exp(b);
        int pbm = 0;
        double e, g = 0.0, m0 = 0.0, m1 = 0.0, m2 = 0.0;
        for (int x = 0; x < k; x++) ed2[x] = exp(-(A[x] * A[x]) * bE);
        for (int x = 0; x < k; x++) m0 += ed2[x];
        if (m0 < realmin) {
            e = -logK;
            pbm = 1;
        } else {
            for (int x = 0; x < k; x++) m1v[x] = ed2[x] * ((A[x] * A[x]) / m0);
            for (int x = 0; x < k; x++) m1 += m1v[x];
            e = bE * m1 + log(m0) - logK;
        }
        if (fabs(e) < tol) break;
        if (B_upper - B_lower < 10.0 * eps) break;
        if (e < 0.0 && b <= B_upper) B_upper = b;
        else if (e > 0.0 && b >= B_lower) B_lower = b;
        pbm = pbm || e < -logK || e > logN - logK;
        if (!pbm) {
            if (i == maxit) { b = (B_lower + B_upper) / 2.0; i = 1; continue; }
            for (int x = 0; x < k; x++) m2 += m1v[x] * (A[x] * A[x]);
            g = (bE * bE) * (m1 * m1 - m2);
            if (g == 0) pbm = 1;
        }
        if (pbm) {
            double esqd_sum = 0.0;
            for (int x = 0; x < k; x++) esqd_sum += exp(-(A[x] * A[x]) * exp(B_lower)) + exp(-(A[x] * A[x]) * exp(B_upper));
            if (esqd_sum < 2.0 * sqrt(realmin)) break;
            b = (B_lower + B_upper) / 2.0;
            i = 1;
            continue;
        }
        b += -e / g;
        if (b < B_lower || b > B_upper) { b = (B_lower + B_upper) / 2.0; i = 0; }
        i++;
    }
    free(ed2); free(m1v);
    return 1.0 / sqrt(2.0 * exp(b));
}

static int cmp_kth(const void *a, const void *b)
{
    const double *x = (const double *)a, *y = (const double *)b;
    if (x[0] < y[0]) return -1;
    if (x[0] > y[0]) return 1;
    return (x[1] > y[1]) - (x[1] < y[1]);
}

/* entropic_affinity_sigmas, mdsctk.cpp:498-565.  A: n rows of k sorted distances (row-major); s[n] out.  The frames are
 * visited in ascending order of their ceil(K)-th distance, each started from the previous frame's solution. */
void oracle_entropic_affinity_sigmas(int n, int k, double K, const double *A, double *s)
{
    const int Ki = (int)ceil(K);
    const double N = (double)k, logK = log(K), logN = log(N), logNK = logN - logK;
    double p1;
    if (logK > log(sqrt(2.0 * N))) {
        p1 = 3.0 / 4.0;
    } else {
        p1 = 1.0 / 4.0;
        for (int x = 0; x < 100; x++) p1 -= (-p1 * log(p1 / N) - logK) / (-log(p1 / N) + 1.0);
        p1 = 1.0 - (p1 / 2.0);
    }
    double *BL = (double *)malloc(sizeof(double) * (size_t)n), *BU = (double *)malloc(sizeof(double) * (size_t)n);
    double *order = (double *)malloc(sizeof(double) * 2 * (size_t)n);
    for (int x = 0; x < n; x++) {
        const double *a = A + (size_t)x * k;
        BU[x] = log((2.0 * log(p1 * (N - 1.0) / (1.0 - p1))) / (a[1] * a[1] - a[0] * a[0]));
        const double bL1 = log((2.0 * logNK / (1.0 - (1.0 / N))) / (a[k - 1] * a[k - 1] - a[0] * a[0]));
        /* SQR(x)*SQR(x) of the reference expands (unparenthesised macro, mdsctk.h:84) to x*x*x*x = ((x*x)*x)*x: keep that order
         * (pinned against the reference's own code, tests/test_ref_slice.py) */
        const double bL2 = log((2.0 * sqrt(logNK)) / sqrt(a[k - 1] * a[k - 1] * a[k - 1] * a[k - 1] - a[0] * a[0] * a[0] * a[0]));
        BL[x] = bL1 > bL2 ? bL1 : bL2;
        order[2 * x] = a[Ki - 1];
        order[2 * x + 1] = (double)x;
    }
    qsort(order, (size_t)n, 2 * sizeof(double), cmp_kth);
    int j = (int)order[1];
    double b0 = (BL[j] + BU[j]) / 2.0;
    for (int t = 0; t < n; t++) {
        j = (int)order[2 * t + 1];
        s[j] = entropic_sigma(A + (size_t)j * k, k, b0, logK, logN, BL[j], BU[j]);
        b0 = log((1.0 / s[j]) * (1.0 / s[j]) / 2.0);
    }
    free(BL); free(BU); free(order);
}

/* w = A v for the symmetric matrix stored as its upper triangle in CSC (sp_dsymv, mdsctk.cpp:293-313) */
void oracle_sp_dsymv(int n, const int *irow, const int *pcol, const double *A, const double *v, double *w)
{
    for (int i = 0; i < n; i++) w[i] = 0.0;
    for (int i = 0; i < n; i++) {
        const double t = v[i];
        int k = pcol[i];
        if ((k != pcol[i + 1]) && (irow[k] == i)) { w[i] += t * A[k]; k++; }
        for (int j = k; j < pcol[i + 1]; j++) {
            w[irow[j]] += t * A[j];
            w[i] += v[irow[j]] * A[j];
        }
    }
}

/* Dense cyclic Jacobi: a (n x n, symmetric, destroyed) -> eigenvalues w[n] ascending, eigenvectors in the rows of z */
static void jacobi_eig(int n, double *a, double *w, double *z)
{
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) z[(size_t)i * n + j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0;
        for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) off += a[(size_t)p * n + q] * a[(size_t)p * n + q];
        if (off < 1e-300) break;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                const double apq = a[(size_t)p * n + q];
                if (fabs(apq) < 1e-300) continue;
                const double theta = (a[(size_t)q * n + q] - a[(size_t)p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; k++) {
                    const double akp = a[(size_t)k * n + p], akq = a[(size_t)k * n + q];
                    a[(size_t)k * n + p] = c * akp - s * akq;
                    a[(size_t)k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {
                    const double apk = a[(size_t)p * n + k], aqk = a[(size_t)q * n + k];
                    a[(size_t)p * n + k] = c * apk - s * aqk;
                    a[(size_t)q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; k++) {
                    const double zpk = z[(size_t)p * n + k], zqk = z[(size_t)q * n + k];
                    z[(size_t)p * n + k] = c * zpk - s * zqk;
                    z[(size_t)q * n + k] = s * zpk + c * zqk;
                }
            }
    }
    for (int i = 0; i < n; i++) w[i] = a[(size_t)i * n + i];
    for (int i = 0; i < n - 1; i++) {            /* selection sort ascending, rows of z follow */
        int m = i;
        for (int j = i + 1; j < n; j++) if (w[j] < w[m]) m = j;
        if (m != i) {
            const double t = w[i]; w[i] = w[m]; w[m] = t;
            for (int k = 0; k < n; k++) { const double u = z[(size_t)i * n + k]; z[(size_t)i * n + k] = z[(size_t)m * n + k]; z[(size_t)m * n + k] = u; }
        }
    }
}

/* The nev algebraically largest eigenpairs of the symmetric upper-CSC matrix, in the order auto_decomp_sparse
 * writes them (largest first, :207-218); evecs [nev][n]; residuals |A z - d z| / |d| (:221-229). */
int oracle_sym_eigs_largest(int n, const int *pcol, const int *irow, const double *M, int nev, double *evals, double *evecs,
                            double *residuals)
{
    if (nev < 1 || nev > n) return -1;
    double *a = (double *)calloc((size_t)n * n, sizeof(double));
    double *w = (double *)malloc(sizeof(double) * (size_t)n), *z = (double *)malloc(sizeof(double) * (size_t)n * n);
    double *ax = (double *)malloc(sizeof(double) * (size_t)n);
    if (!a || !w || !z || !ax) return -1;
    for (int x = 0; x < n; x++)
        for (int y = pcol[x]; y < pcol[x + 1]; y++) { a[(size_t)x * n + irow[y]] = M[y]; a[(size_t)irow[y] * n + x] = M[y]; }
    jacobi_eig(n, a, w, z);
    for (int e = 0; e < nev; e++) {
        const int src = n - 1 - e;
        evals[e] = w[src];
        memcpy(evecs + (size_t)e * n, z + (size_t)src * n, sizeof(double) * (size_t)n);
        oracle_sp_dsymv(n, irow, pcol, M, evecs + (size_t)e * n, ax);
        double nrm = 0.0;
        for (int i = 0; i < n; i++) { const double r = ax[i] - evals[e] * evecs[(size_t)e * n + i]; nrm += r * r; }
        residuals[e] = sqrt(nrm) / fabs(evals[e]);
    }
    free(a); free(w); free(z); free(ax);
    return 0;
}
