/*
 * oracle.h -- CPU restatement of MDSCTK's all-pairs distance + kNN stage.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product
 * path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call this code, and there only as
 * the checker / timed CPU baseline.
 *
 * PARITY: PINNED FOR THE VECTOR PATH, UNPINNED FOR THE RMSD PATH.  The reference (jlphillipsphd/mdsctk) ships no golden
 * vectors and no known-answer tests and cannot be compiled as a whole in this image (it needs libgromacs,
 * Boost.program_options, Berkeley DB and ARPACK, none of which are present).
 *   Pinned against outputs of the reference itself: the plain C++ arithmetic of mdsctk.cpp / mdsctk.h -- euclidean_distance,
 *   correlation_distance, permutation<T>::sort, euclidean_distance_sparse, entropic_affinity_sigma(s), sp_dsymv / sp_dgemv,
 *   torsion -- compiles on its own; oracle/ref_slice.sh cuts those definitions out of /root/reference where they lie into
 *   oracle/_ref/ (git-ignored) and builds them unmodified.  tests/test_ref_slice.py holds this oracle to them bit for bit
 *   (knn_data in and out of sample, both metrics, D up to 512; the sparse metric; sigmas; torsions), and
 *   tests/golden/refslice_vectors.npz carries reference-made vectors to the GPU box.
 *   Unpinned: the RMSD arithmetic lives in GROMACS 5.0-5.1 (un-vendored, un-pinned: /root/reference/CMakeLists.txt:194-209),
 *   whose do_fit / rmsdev / reset_x algorithm is restated here and cross-checked only by an independent FP64 Kabsch and
 *   the invariance properties in tests/test_oracle.py.
 * The oracle is anchored on the reference's call sites:
 *   knn_rms.cpp:38-41    distance() = do_fit + rmsdev * 10
 *   knn_rms.cpp:181-206  mass weights, reset_x on every frame
 *   knn_rms.cpp:224-293  k clamp, row blocks, rank-0 drop, file layout
 *   knn_data.cpp:141-250 reader, blocks, writer (argument order of distance(): fitting row first)
 *   mdsctk.h:177-199     permutation<T>::sort(k) = partial_sort
 *   mdsctk.cpp:330-360   euclidean_distance / correlation_distance
 * and cross-checked by the invariance properties in tests/test_oracle.py.
 */
#ifndef MDSCTK_ORACLE_H
#define MDSCTK_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- XTC trajectory reader (GROMACS xtc / xdr3dfcoord format) -------------
 * Replaces open_xtc/read_first_xtc/read_next_xtc as used at
 * knn_rms.cpp:155-157,186-203.  Returns 0 on success. */
int oracle_xtc_scan(const char *path, int *natoms, long long *nframes);
/* xyz must hold nframes*natoms*3 floats (nm), AoS [frame][atom][3]. */
int oracle_xtc_read(const char *path, int natoms, long long nframes, float *xyz);

/* ---- topology masses (read_tps_conf(..., bMass=TRUE), knn_rms.cpp:150-153,181-182)
 * Reads a .pdb or .gro file; mass[] must hold max_atoms floats.
 * Returns the atom count, or a negative number on error. */
int oracle_top_masses(const char *path, float *mass, int max_atoms);

/* ---- per-frame / per-pair arithmetic -------------------------------------- */
/* reset_x(natoms,NULL,natoms,NULL,x,mass): float centre of mass removal. */
void oracle_reset_x(int natoms, float *x /*[natoms][3]*/, const float *mass);
/* do_fit: least-squares rotate x onto xp IN PLACE (float U, double 6x6 Jacobi). */
void oracle_do_fit(int natoms, const float *w, const float *xp, float *x);
/* rmsdev: sqrt(sum m |x-xp|^2 / sum m), float accumulation. */
float oracle_rmsdev(int natoms, const float *mass, const float *x, const float *xp);
/* FP64 optimal-superposition RMSD (nm) between two RAW (uncentred) float frames:
 * double centring, double cross-covariance, 4x4 key-matrix Jacobi.            */
double oracle_rmsd_f64(int natoms, const float *mass, const float *a, const float *b, int dofit);

/* mdsctk.cpp:330-335 and :337-360 */
double oracle_euclidean_distance(int size, const double *reference, const double *fitting);
double oracle_correlation_distance(int size, const double *reference, const double *fitting);

/* ---- whole-stage drivers ---------------------------------------------------
 * Both write out_dist[n_fit][k] (ascending, sorted position 0 dropped, exactly
 * like knn_rms.cpp:282-291) and out_idx[n_fit][k].  k must already be clamped
 * to n_ref-1.  Ties are ordered (distance, index) ascending.
 *
 * mode 0: reference-faithful float chain (frames centred in float with
 *         reset_x, then for each fit row: copy fit once, do_fit cumulatively
 *         in place against every ref, rmsdev*10) -- the timed CPU baseline.
 * mode 1: FP64 Kabsch on the raw frames (x10 to Angstrom).
 * ref_xyz / fit_xyz are RAW decoded frames (uncentred), nm.                     */
int oracle_knn_rms(int mode, int natoms, const float *mass,
                   const float *ref_xyz, long long n_ref,
                   const float *fit_xyz, long long n_fit,
                   int k, int dofit, int nthreads,
                   double *out_dist, int *out_idx);

/* metric 0 = euclidean, 1 = correlation.  rows are row-major double[n][dim]. */
int oracle_knn_data(int metric, int dim,
                    const double *ref, long long n_ref,
                    const double *fit, long long n_fit,
                    int k, int nthreads,
                    double *out_dist, int *out_idx);

/* Full distance row(s) without selection, for spot checks: out[n_fit][n_ref]. */
int oracle_rms_rows(int mode, int natoms, const float *mass,
                    const float *ref_xyz, long long n_ref,
                    const float *fit_xyz, long long n_fit,
                    int dofit, int nthreads, double *out);

int oracle_max_threads(void);

/* make_sysparse (make_sysparse.cpp:245-329): symmetric CSC from the kNN files; returns nnz (csc.c). */
/* auto_decomp_sparse (spectral.c): affinity stage, sp_dsymv, dense eigen-solve standing in for ARPACK */
double oracle_affinity(int n, const int *pcol, const int *irow, double *M, int k_a);
/* auto_decomp_sparse -K: entropic_affinity_sigmas (mdsctk.cpp:498-565); A = n rows of k sorted distances */
void oracle_entropic_affinity_sigmas(int n, int k, double K, const double *A, double *s);
void oracle_sp_dsymv(int n, const int *irow, const int *pcol, const double *A, const double *v, double *w);
int oracle_sym_eigs_largest(int n, const int *pcol, const int *irow, const double *M, int nev, double *evals, double *evecs,
                            double *residuals);

/* knn_data_sparse (mdsctk.cpp:362-386, knn_data_sparse.cpp:195-262); vectors in CSR form */
double oracle_euclidean_distance_sparse(int ref_size, const int *ref_index, const double *ref_data, int fit_size,
                                        const int *fit_index, const double *fit_data);
int oracle_knn_data_sparse(const long long *ref_off, const int *ref_idx, const double *ref_val, long long n_ref,
                           const long long *fit_off, const int *fit_idx, const double *fit_val, long long n_fit, int k,
                           double *out_dist, int *out_idx);

/* bb_xtc_to_phipsi / angles_to_sincos (featurize.c) */
float oracle_torsion(const float *pos1, const float *pos2, const float *pos3, const float *pos4);
void oracle_phipsi(const float *xyz, long long n, int natoms, double *phipsi);
void oracle_sincos(const double *angles, long long n, double *out);

/* make_gesparse [-s] (make_gesparse.cpp:246-333): general CSC; returns nnz (capacity n*k, or 2*n*k with symmetric) */
long long oracle_make_gesparse(const int *idx, const double *dist, long long n, int maxk, int k, int symmetric, int *pcol,
                               int *irow, double *val);
long long oracle_make_sysparse(const int *idx, const double *dist, long long n, int maxk, int k, int *pcol, int *irow,
                               double *val);

#ifdef __cplusplus
}
#endif
#endif
