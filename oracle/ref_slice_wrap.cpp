// extern "C" doors into the slice of the reference that oracle/ref_slice.sh cuts out of /root/reference/mdsctk.{h,cpp}.
// Nothing here restates reference arithmetic: the prelude supplies what mdsctk.h would (macros, the `real` typedef, the unit
// constant GROMACS defines), the wrappers only marshal arrays.  The one loop below, ref_knn_data_rows, is the row loop of
// knn_data.cpp:195-250 reduced to its two calls -- distance function, then permutation<double>::sort(k+1) -- both of which
// ARE the reference's code.  Test infrastructure (see oracle.h); never linked into the product.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <omp.h>
#include <vector>
using namespace std;
typedef float real;                       // GROMACS single precision, as the reference is built
#define SQR(A) (A*A)                      /* mdsctk.h:84-85 */
#define AVG(A,B) ((A+B)/2.0)
#ifndef RAD2DEG
#define RAD2DEG (180.0/M_PI)              /* gromacs/math/units.h */
#endif
#include "slice_generated.inc"

extern "C" {

double ref_euclidean_distance(int size, double *reference, double *fitting) { return euclidean_distance(size, reference, fitting); }
double ref_correlation_distance(int size, double *reference, double *fitting) { return correlation_distance(size, reference, fitting); }
double ref_euclidean_distance_sparse(int rn, int *ri, double *rd, int fn, int *fi, double *fd)
{
    return euclidean_distance_sparse(rn, ri, rd, fn, fi, fd);
}

// permutation<double>::sort(k) on a copy of data[n]: the first k sorted values and their original positions
void ref_partial_sort(int n, int k, const double *data, double *sorted_k, int *index_k)
{
    permutation<double> p;
    p.data.assign(data, data + n);
    p.sort(k);
    const int m = k ? k : n;
    for (int i = 0; i < m; ++i) { sorted_k[i] = p.data[i]; index_k[i] = p.indices[i]; }
}

// rows of knn_data: for every fitting vector the distances to all reference vectors (metric 0 Euclidean, 1 correlation),
// permutation sort of the first k+1, position 0 dropped (knn_data.cpp:195-250).  dist/idx: [n_fit][k].
void ref_knn_data_rows(int n_fit, int n_ref, int dim, int k, int metric, double *fit, double *ref, double *dist, int *idx, int nthreads)
{
    double (*fn)(int, double *, double *) = metric ? correlation_distance : euclidean_distance;
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for num_threads(nthreads)      /* as at knn_data.cpp:196 / 229 */
    for (int f = 0; f < n_fit; ++f) {
        permutation<double> p;
        p.data.resize(n_ref);
        // argument order of the call sites knn_data.cpp:199-201 / 232-234: (size, fitting row, reference row)
        for (int r = 0; r < n_ref; ++r) p.data[r] = fn(dim, fit + (size_t)f * dim, ref + (size_t)r * dim);
        p.sort(k + 1);
        for (int j = 0; j < k; ++j) { dist[(size_t)f * k + j] = p.data[j + 1]; idx[(size_t)f * k + j] = p.indices[j + 1]; }
    }
}

// A: [n][k] ascending neighbour distances per frame (the reference's vector<double> per row), K: perplexity; s: [n] sigmas
void ref_entropic_affinity_sigmas(int n, int k, double K, const double *A, double *s)
{
    vector<vector<double> > rows(n);
    for (int i = 0; i < n; ++i) rows[i].assign(A + (size_t)i * k, A + (size_t)(i + 1) * k);
    entropic_affinity_sigmas(n, k, K, &rows[0], s);
}

void ref_sp_dsymv(int n, int *irow, int *pcol, double *A, double *v, double *w) { sp_dsymv(n, irow, pcol, A, v, w); }
void ref_sp_dgemv(int n, int *irow, int *pcol, double *A, double *v, double *w) { sp_dgemv(n, irow, pcol, A, v, w); }

float ref_torsion(float *p1, float *p2, float *p3, float *p4, int degrees) { return torsion(p1, p2, p3, p4, degrees != 0); }

}  // extern "C"
