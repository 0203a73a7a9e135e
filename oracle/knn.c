/*
 * knn.c -- ORACLE (test infrastructure, see oracle.h): whole-stage drivers.
 *
 * Row loop, selection and output follow the reference tools:
 *   knn_rms.cpp:231-293  / knn_data.cpp:195-250
 *     #pragma omp parallel for over the fit frames of a block; per fit frame
 *     the distance to EVERY reference frame is stored in a row of doubles,
 *     permutation<double>::sort(k+1) (mdsctk.h:177-199, std::partial_sort on
 *     indices then on data) and entries [1..k] are written -- sorted position 0
 *     is dropped unconditionally.
 * Output order is frame order whatever the block size, so blocks are not
 * restated; ties (unspecified in libstdc++'s heap select) are broken by
 * (distance, index) ascending.
 *
 * mode 0 keeps the reference's float chain including its side effect: the fit
 * frame is copied once per row (knn_rms.cpp:272) and do_fit then rotates that
 * same copy in place for every reference frame of the sweep (:273-276).
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_max_threads(void)
{
#ifdef _OPENMP
    int a = omp_get_max_threads(), b = omp_get_num_procs();
    return a > b ? b : a; /* knn_rms.cpp:76 default */
#else
    return 1;
#endif
}

/* keep the k1 smallest (value, index) pairs of row[0..n), ascending */
typedef struct { double v; int i; } cand;
static int cand_less(const cand *a, const cand *b) { return a->v < b->v || (a->v == b->v && a->i < b->i); }

static void sift_down(cand *h, int n, int p)
{
    for (;;) {
        int l = 2 * p + 1, r = l + 1, m = p;
        if (l < n && cand_less(&h[m], &h[l])) m = l;
        if (r < n && cand_less(&h[m], &h[r])) m = r;
        if (m == p) return;
        cand t = h[p]; h[p] = h[m]; h[m] = t;
        p = m;
    }
}

static void select_smallest(const double *row, long long n, int k1, cand *heap)
{
    int filled = 0;
    for (long long j = 0; j < n; j++) {
        cand c = {row[j], (int)j};
        if (filled < k1) {
            heap[filled++] = c;
            if (filled == k1)
                for (int p = k1 / 2 - 1; p >= 0; p--) sift_down(heap, k1, p);
        } else if (cand_less(&c, &heap[0])) {
            heap[0] = c;
            sift_down(heap, k1, 0);
        }
    }
    if (filled < k1)
        for (int p = filled / 2 - 1; p >= 0; p--) sift_down(heap, filled, p);
    for (int m = filled - 1; m > 0; m--) { /* heap sort, ascending */
        cand t = heap[0]; heap[0] = heap[m]; heap[m] = t;
        sift_down(heap, m, 0);
    }
}

static float *centred_copy(int natoms, const float *mass, const float *xyz, long long n)
{
    size_t fs = (size_t)natoms * 3;
    float *c = (float *)malloc(sizeof(float) * fs * (size_t)(n > 0 ? n : 1));
    if (!c) return NULL;
    memcpy(c, xyz, sizeof(float) * fs * (size_t)n);
    for (long long f = 0; f < n; f++) oracle_reset_x(natoms, c + fs * (size_t)f, mass);
    return c;
}

static void rms_row(int mode, int natoms, const float *mass, const float *ref_c, const float *ref_raw,
                    long long n_ref, const float *fit_c, const float *fit_raw, int dofit, float *fit_buf,
                    float *ref_buf, double *row)
{
    size_t fs = (size_t)natoms * 3;
    if (mode == 0) {
        memcpy(fit_buf, fit_c, sizeof(float) * fs); /* once per fit row */
        for (long long r = 0; r < n_ref; r++) {
            memcpy(ref_buf, ref_c + fs * (size_t)r, sizeof(float) * fs);
            if (dofit) oracle_do_fit(natoms, mass, ref_buf, fit_buf);
            row[r] = (double)oracle_rmsdev(natoms, mass, ref_buf, fit_buf) * 10.0;
        }
    } else {
        for (long long r = 0; r < n_ref; r++)
            row[r] = oracle_rmsd_f64(natoms, mass, ref_raw + fs * (size_t)r, fit_raw, dofit) * 10.0;
    }
}

int oracle_rms_rows(int mode, int natoms, const float *mass, const float *ref_xyz, long long n_ref,
                    const float *fit_xyz, long long n_fit, int dofit, int nthreads, double *out)
{
    size_t fs = (size_t)natoms * 3;
    float *ref_c = NULL, *fit_c = NULL;
    if (mode == 0) {
        ref_c = centred_copy(natoms, mass, ref_xyz, n_ref);
        fit_c = centred_copy(natoms, mass, fit_xyz, n_fit);
        if (!ref_c || !fit_c) { free(ref_c); free(fit_c); return -1; }
    }
    if (nthreads <= 0) nthreads = oracle_max_threads();
#pragma omp parallel num_threads(nthreads)
    {
        float *fit_buf = (float *)malloc(sizeof(float) * fs);
        float *ref_buf = (float *)malloc(sizeof(float) * fs);
#pragma omp for schedule(static)
        for (long long f = 0; f < n_fit; f++)
            rms_row(mode, natoms, mass, ref_c, ref_xyz, n_ref, fit_c ? fit_c + fs * (size_t)f : NULL,
                    fit_xyz + fs * (size_t)f, dofit, fit_buf, ref_buf, out + (size_t)f * (size_t)n_ref);
        free(fit_buf);
        free(ref_buf);
    }
    free(ref_c);
    free(fit_c);
    return 0;
}

int oracle_knn_rms(int mode, int natoms, const float *mass, const float *ref_xyz, long long n_ref,
                   const float *fit_xyz, long long n_fit, int k, int dofit, int nthreads, double *out_dist,
                   int *out_idx)
{
    if (k < 0 || k > n_ref - 1) return -2;
    const int k1 = k + 1;
    size_t fs = (size_t)natoms * 3;
    float *ref_c = NULL, *fit_c = NULL;
    if (mode == 0) {
        ref_c = centred_copy(natoms, mass, ref_xyz, n_ref);
        fit_c = centred_copy(natoms, mass, fit_xyz, n_fit);
        if (!ref_c || !fit_c) { free(ref_c); free(fit_c); return -1; }
    }
    if (nthreads <= 0) nthreads = oracle_max_threads();
#pragma omp parallel num_threads(nthreads)
    {
        float *fit_buf = (float *)malloc(sizeof(float) * fs);
        float *ref_buf = (float *)malloc(sizeof(float) * fs);
        double *row = (double *)malloc(sizeof(double) * (size_t)n_ref);
        cand *heap = (cand *)malloc(sizeof(cand) * (size_t)k1);
#pragma omp for schedule(static)
        for (long long f = 0; f < n_fit; f++) {
            rms_row(mode, natoms, mass, ref_c, ref_xyz, n_ref, fit_c ? fit_c + fs * (size_t)f : NULL,
                    fit_xyz + fs * (size_t)f, dofit, fit_buf, ref_buf, row);
            select_smallest(row, n_ref, k1, heap);
            for (int j = 0; j < k; j++) { /* position 0 dropped: knn_rms.cpp:284-285 */
                out_dist[(size_t)f * k + j] = heap[j + 1].v;
                out_idx[(size_t)f * k + j] = heap[j + 1].i;
            }
        }
        free(fit_buf);
        free(ref_buf);
        free(row);
        free(heap);
    }
    free(ref_c);
    free(fit_c);
    return 0;
}

int oracle_knn_data(int metric, int dim, const double *ref, long long n_ref, const double *fit, long long n_fit,
                    int k, int nthreads, double *out_dist, int *out_idx)
{
    if (k < 0 || k > n_ref - 1) return -2;
    const int k1 = k + 1;
    if (nthreads <= 0) nthreads = oracle_max_threads();
#pragma omp parallel num_threads(nthreads)
    {
        double *row = (double *)malloc(sizeof(double) * (size_t)n_ref);
        cand *heap = (cand *)malloc(sizeof(cand) * (size_t)k1);
#pragma omp for schedule(static)
        for (long long f = 0; f < n_fit; f++) {
            const double *frow = fit + (size_t)f * dim;
            /* argument order as at knn_data.cpp:231-234: distance(size, fit_row, ref_row) */
            if (metric == 0)
                for (long long r = 0; r < n_ref; r++)
                    row[r] = oracle_euclidean_distance(dim, frow, ref + (size_t)r * dim);
            else
                for (long long r = 0; r < n_ref; r++)
                    row[r] = oracle_correlation_distance(dim, frow, ref + (size_t)r * dim);
            select_smallest(row, n_ref, k1, heap);
            for (int j = 0; j < k; j++) {
                out_dist[(size_t)f * k + j] = heap[j + 1].v;
                out_idx[(size_t)f * k + j] = heap[j + 1].i;
            }
        }
        free(row);
        free(heap);
    }
    return 0;
}

/* euclidean_distance_sparse, mdsctk.cpp:362-386: merge of two ascending index lists; a dimension present in
 * one vector only contributes its square, terms are added in ascending index order. */
double oracle_euclidean_distance_sparse(int ref_size, const int *ref_index, const double *ref_data, int fit_size,
                                        const int *fit_index, const double *fit_data)
{
    double value = 0.0;
    int ref = 0, fit = 0;
    for (ref = 0; ref < ref_size; ref++) {
        while (fit < fit_size && fit_index[fit] < ref_index[ref]) {
            value += (fit_data[fit] * fit_data[fit]);
            fit++;
        }
        if (fit < fit_size && ref_index[ref] == fit_index[fit]) {
            value += ((ref_data[ref] - fit_data[fit]) * (ref_data[ref] - fit_data[fit]));
            fit++;
        } else {
            value += (ref_data[ref] * ref_data[ref]);
        }
    }
    for (; fit < fit_size; fit++) value += (fit_data[fit] * fit_data[fit]);
    return sqrt(value);
}

/* knn_data_sparse.cpp:195-262 row loop on CSR-style inputs: vector v = entries [off[v], off[v+1]). */
int oracle_knn_data_sparse(const long long *ref_off, const int *ref_idx, const double *ref_val, long long n_ref,
                           const long long *fit_off, const int *fit_idx, const double *fit_val, long long n_fit, int k,
                           double *out_dist, int *out_idx)
{
    if (k < 0 || k > n_ref - 1) return -2;
    const int k1 = k + 1;
    double *row = (double *)malloc(sizeof(double) * (size_t)n_ref);
    cand *heap = (cand *)malloc(sizeof(cand) * (size_t)k1);
    for (long long f = 0; f < n_fit; f++) {
        for (long long r = 0; r < n_ref; r++)
            row[r] = oracle_euclidean_distance_sparse((int)(ref_off[r + 1] - ref_off[r]), ref_idx + ref_off[r], ref_val + ref_off[r],
                                                      (int)(fit_off[f + 1] - fit_off[f]), fit_idx + fit_off[f], fit_val + fit_off[f]);
        select_smallest(row, n_ref, k1, heap);
        for (int j = 0; j < k; j++) {
            out_dist[(size_t)f * k + j] = heap[j + 1].v;
            out_idx[(size_t)f * k + j] = heap[j + 1].i;
        }
    }
    free(row);
    free(heap);
    return 0;
}
