/*
 * mdsctk_knn.h -- C ABI of the B200 all-pairs distance + k-nearest-neighbour stage.
 *
 * This is the drop-in boundary for MDSCTK's kNN hot path.  The reference has no
 * library/FFI seam for it (every tool is one main()): the seam
 * that exists is the per-pair distance function and the row loop around it.
 * Each entry point below names the reference code it replaces:
 *
 *   mdsctk_knn_rms_set_reference   knn_rms.cpp:181-194  weights[] = masses, reset_x on
 *                                                       every reference frame (load+centre)
 *   mdsctk_knn_rms_query           knn_rms.cpp:231-293  OpenMP row blocks:
 *                                    :38-41   distance() = do_fit + rmsdev * 10.0
 *                                    :277-278 fits[frame].sort(k1)  (mdsctk.h:177-199)
 *   mdsctk_knn_data_set_reference  knn_data.cpp:141-154 reference rows
 *   mdsctk_knn_data_query          knn_data.cpp:195-250 row blocks:
 *                                    :231-234 ::distance(vector_size, fit, ref)
 *                                             (euclidean_distance mdsctk.cpp:330-335,
 *                                              correlation_distance mdsctk.cpp:337-360)
 *                                    :235-236 fits[frame].sort(k1)
 *
 * Results are the k1 = k+1 smallest distances of every fit row, ascending, with
 * their reference indices -- i.e. permutation<double>::data[0..k1) / indices[0..k1)
 * after sort(k1).  The CALLER drops sorted position 0 and writes the files,
 * exactly like knn_rms.cpp:282-291 / knn_data.cpp:240-249.
 *
 * Conventions: plain pointers and sizes, no C++ or torch types.  Every function
 * returns 0 on success or a negative MDSCTK_KNN_E* code; the message is available
 * from mdsctk_knn_last_error().  Nothing throws or exits across the boundary.
 * A ctx is bound to ONE GPU and is not thread-safe; use one ctx per GPU (one
 * process per GPU under torchrun, or one host thread per GPU in the C++ tools).
 * Calls are synchronous on return.  There is no CPU fallback: without a usable
 * sm_100 device mdsctk_knn_create fails.
 */
#ifndef MDSCTK_KNN_H
#define MDSCTK_KNN_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDSCTK_KNN_ABI_VERSION 2

enum {
    MDSCTK_KNN_OK = 0,
    MDSCTK_KNN_EINVAL = -1,   /* bad argument                                    */
    MDSCTK_KNN_ECUDA = -2,    /* CUDA runtime / driver error                     */
    MDSCTK_KNN_ENOMEM = -3,   /* device or host allocation failed                */
    MDSCTK_KNN_ESTATE = -4,   /* call order (no reference set yet, ...)          */
    MDSCTK_KNN_ENODEV = -5,   /* no sm_100 device / device id out of range       */
    MDSCTK_KNN_EAUDIT = -6    /* a certified row disagrees with its exact recomputation (results were still
                                 written; the message names the rows) -- never expected, fails loudly */
};

enum { MDSCTK_KNN_EUCLIDEAN = 0, MDSCTK_KNN_CORRELATION = 1 };

/* rms_kernel option values */
enum {
    MDSCTK_KNN_RMS_SIMT_FP32 = 0,  /* FP32 CUDA-core contraction                                  */
    MDSCTK_KNN_RMS_TC_3XTF32 = 1,  /* tcgen05 kind::tf32, hi/lo operand split (3 MMAs)              */
    MDSCTK_KNN_RMS_TC_1XTF32 = 2,  /* tcgen05 kind::tf32, hi only (coarse filter)                   */
    MDSCTK_KNN_RMS_TC_3XBF16 = 3,  /* tcgen05 kind::f16 on bf16 hi/mid operand split (3 MMAs)       */
    MDSCTK_KNN_RMS_TC_3XFP16 = 4,  /* tcgen05 kind::f16 on fp16 hi/lo split of 64x (3 MMAs, 22 bits)          */
    MDSCTK_KNN_RMS_TC_2XFP16 = 5,  /* fit operand fp16 hi only, reference hi/lo (2 MMAs)            */
    MDSCTK_KNN_RMS_TC_1XFP16 = 6   /* DEFAULT: both operands fp16 hi only (1 MMA); exactness comes from the FP64
                                      re-score, whose certificate bounds the rounding rigorously
                                      (metric triangle inequality on the rounded structures)       */
};

typedef struct mdsctk_knn_ctx mdsctk_knn_ctx;

typedef struct mdsctk_knn_stats {
    double ms_upload;      /* host->device copies of inputs                         */
    double ms_pack;        /* centre + scale + SoA pack kernel                      */
    double ms_sweep;       /* all-pairs distance sweep + streaming selection        */
    double ms_rescore;     /* FP64 re-score, final sort, certification              */
    double ms_fallback;    /* exact FP64 rows for uncertified rows                  */
    double ms_download;    /* device->host copy of the k-lists                      */
    long long pairs;       /* n_fit * n_ref of the last query                       */
    long long launches;    /* kernels launched by the last query (sweep..fallback)  */
    long long fallback_rows;   /* rows that failed certification and were redone    */
    long long sweep_appends;   /* candidates appended by the sweep (diagnostic)     */
    double max_filter_err; /* max |approx d^2 - exact d^2| over all candidates      */
    double max_filter_spread; /* max over rows of max-min of (approx - exact) d^2        */
    double cert_eps;       /* absolute d^2 noise margin used by the certificate     */
    int rms_kernel;        /* kernel actually used by the last RMSD query           */
    int k_keep;            /* candidates kept per row (k1 + slack)                  */
    int lists_per_row;     /* candidate lists per row kept by the sweep             */
    int rescored_max;      /* most candidates any row needed before its certificate held */
    double cert_gres;      /* 2xFP16 / 1xFP16: largest operand-rounding residual norm of the reference set (nm) */
    long long audit_rows;       /* certified rows recomputed exactly (full FP64 row + exact selection) by the last query */
    long long audit_mismatches; /* of those, rows whose result differed: must be 0                                     */
    int sweep_version;          /* tensor-core RMSD sweep used: 2 = resident fit tile + pass director (rms_tc2.cu), 1 = rms_tc.cu, 0 = n/a */
} mdsctk_knn_stats;

int mdsctk_knn_abi_version(void);

/* One context per GPU.  device_id is a CUDA ordinal (honours CUDA_VISIBLE_DEVICES). */
int mdsctk_knn_create(mdsctk_knn_ctx **out, int device_id);
void mdsctk_knn_destroy(mdsctk_knn_ctx *ctx);
/* ctx may be NULL: returns the message of the last failed mdsctk_knn_create on this thread. */
const char *mdsctk_knn_last_error(const mdsctk_knn_ctx *ctx);

/* Tunables: "rms_kernel" (enum above; set it BEFORE loading the reference set: the pack kernel writes only the
 * operand planes that kernel reads -- a later change re-packs from the resident raw frames),
 * "slack" (extra candidates kept per row, -1 = auto),
 * "cert_scale_ppm" (certificate margin multiplier in parts-per-million of the default),
 * "chunk_rows" (fit rows per internal row block of a query; 0 = auto, the default: as many rows as 6 GB of candidate
 * lists allow for the RMSD path -- 265 216 at k = 64 -- and 131072 for knn_data; bounds the working memory the way the
 * reference's --block-size bounds its row buffers, knn_rms.cpp:213-221),
 * "data_kernel" (-1 auto, 0 exact FP64 sweep, 1 / 2 tensor-core filter with 3 / 1 fp16 parts + exact re-score),
 * "audit_rows" (RMSD path: certified rows per row block that are recomputed through the exact FP64 path and
 * compared, default 8; a mismatch makes the query return MDSCTK_KNN_EAUDIT; 0 = off),
 * "sweep_version" (1xFP16 sweep: 2 = resident fit tile + pass director where the tile fits, the default; 1 = the
 * streaming kernel of round 1),
 * "rms_wide_stages" (0/1, default 1: the version-2 sweep moves the reference operand in 64-atom stages of 128-byte rows
 * where three of them fit beside the resident fit tile; 0 = 32-atom stages; same output either way),
 * "force_exact" (0/1: every row goes through the exact FP64 path -- the certificate decides nothing; test hook),
 * "debug_tile" (0/1, see mdsctk_knn_debug_fetch_tile). */
int mdsctk_knn_set_option(mdsctk_knn_ctx *ctx, const char *key, long long value);
int mdsctk_knn_get_stats(const mdsctk_knn_ctx *ctx, mdsctk_knn_stats *out);

/* ------------------------------------------------------------------ RMSD path ---- */
/* xyz: HOST, AoS float[n_frames][n_atoms][3] in nm, UNcentred (decoded XTC frames).
 * mass: HOST float[n_atoms].  Uploads, centres (mass-weighted), scales by sqrt(m/M) and
 * packs on the GPU into the frames x 3 x atoms SoA layout. */
int mdsctk_knn_rms_set_reference(mdsctk_knn_ctx *ctx, const float *xyz, long long n_frames, int n_atoms,
                                 const float *mass);

/* fit_xyz == NULL: the fit set is the reference set (symmetric run, knn_rms.cpp:103-104).
 * k1 = neighbours INCLUDING sorted position 0 (k+1 of the tools), 1 <= k1 <= n_ref.
 * do_fit = 0 is the tools' --nofit.  out_dist[n_fit][k1] in Angstrom (nm * 10), ascending;
 * out_idx[n_fit][k1] reference frame numbers.  Ties: (distance, index) ascending.
 * out_dist/out_idx may both be NULL: results stay on the device for mdsctk_knn_fetch. */
int mdsctk_knn_rms_query(mdsctk_knn_ctx *ctx, const float *fit_xyz, long long n_fit, int k1, int do_fit,
                         double *out_dist, int *out_idx);

/* Row-sharded form (one process per GPU; every GPU holds the whole reference set):
 *   1. mdsctk_knn_rms_alloc_reference(n_total)          on every rank
 *   2. mdsctk_knn_rms_pack_shard(own frames, offset)    each rank packs ITS frames on ITS GPU
 *   3. all-gather the arrays listed by mdsctk_knn_rms_reference_arrays over NCCL
 *      (each is frame-major, bytes_per_frame[i] bytes per frame, n_total frames: the raw frames, the per-frame
 *      scalars -- G, centroid, singular values, rounded norms, rounding residuals -- and ONLY the operand planes
 *      of the sweep kernel selected when the set was allocated: 8 arrays, 5.5 KB per 300-atom frame with the
 *      default 1xFP16 kernel; at most 13 -- pass max_arrays >= 16)
 *   4. mdsctk_knn_rms_query_range(begin, n)             fit rows = reference rows [begin, begin+n) */
int mdsctk_knn_rms_alloc_reference(mdsctk_knn_ctx *ctx, long long n_total, int n_atoms, const float *mass);
int mdsctk_knn_rms_pack_shard(mdsctk_knn_ctx *ctx, const float *xyz, long long frame_offset, long long n_frames);
int mdsctk_knn_rms_reference_arrays(mdsctk_knn_ctx *ctx, int max_arrays, int *n_arrays, void **dev_ptrs,
                                    size_t *bytes_per_frame);
int mdsctk_knn_rms_query_range(mdsctk_knn_ctx *ctx, long long fit_begin, long long n_fit, int k1, int do_fit,
                               double *out_dist, int *out_idx);

/* ---------------------------------------------------------------- vector path ---- */
/* rows: HOST row-major double[n_rows][dim], exactly the bytes of a .pts file. */
int mdsctk_knn_data_set_reference(mdsctk_knn_ctx *ctx, const double *rows, long long n_rows, int dim);
/* metric: MDSCTK_KNN_EUCLIDEAN / MDSCTK_KNN_CORRELATION (knn_data -c). */
int mdsctk_knn_data_query(mdsctk_knn_ctx *ctx, const double *fit_rows, long long n_fit, int k1, int metric,
                          double *out_dist, int *out_idx);
int mdsctk_knn_data_alloc_reference(mdsctk_knn_ctx *ctx, long long n_total, int dim);
int mdsctk_knn_data_upload_shard(mdsctk_knn_ctx *ctx, const double *rows, long long row_offset, long long n_rows);
int mdsctk_knn_data_reference_arrays(mdsctk_knn_ctx *ctx, int max_arrays, int *n_arrays, void **dev_ptrs,
                                     size_t *bytes_per_row);
int mdsctk_knn_data_query_range(mdsctk_knn_ctx *ctx, long long fit_begin, long long n_fit, int k1, int metric,
                                double *out_dist, int *out_idx);

/* knn_data --sort false (knn_data.cpp:198-216): the full rows of distances, out[n_fit][n_reference] (host),
 * in reference order, same arithmetic as the queries.  fit_rows == NULL: the reference rows themselves. */
int mdsctk_knn_data_rows(mdsctk_knn_ctx *ctx, const double *fit_rows, long long n_fit, int metric, double *out);

/* ---- consumer of the kNN files: symmetric CSC matrix (make_sysparse.cpp:245-329) ----------------
 * idx / dist: host, row-major [n][maxk] exactly as indices.dat / distances.dat hold them; only the
 * first k entries of every row are used (make_sysparse's -n / --output-knn, make_sysparse.cpp:93-103).
 * Edge (min(i,j), max(i,j)) -> distance; when both endpoints list each other the value written from
 * the larger endpoint wins (the reference's Db::put overwrites in row order); self edges are dropped.
 * build: fills pcol[n+1] (host) and *nnz; fetch: irow[nnz], val[nnz] (host), rows ascending per column --
 * together the arrays make_sysparse writes after the leading int n (make_sysparse.cpp:310-329). */
int mdsctk_knn_csc_build_sym(mdsctk_knn_ctx *ctx, const int *idx, const double *dist, long long n, int maxk, int k,
                             int *pcol, long long *nnz);
/* General (non-symmetric) CSC matrix, make_gesparse.cpp:246-275: column i holds row i's entries (self entries
 * included, the last duplicate wins); with symmetric != 0 (make_gesparse -s) entry (j, i) <- d(i, j) is added
 * wherever row j does not list i itself (the first such entry sticks).  Same outputs as build_sym. */
int mdsctk_knn_csc_build_general(mdsctk_knn_ctx *ctx, const int *idx, const double *dist, long long n, int maxk, int k,
                                 int symmetric, int *pcol, long long *nnz);
int mdsctk_knn_csc_fetch(mdsctk_knn_ctx *ctx, int *irow, double *val);

/* ---- spectral stage: auto_decomp_sparse.cpp:150-236 / decomp_sparse ------------------------------------
 * pcol[n+1], irow[nnz], val[nnz]: host, the symmetric CSC matrix as make_sysparse writes it (strict upper
 * triangle).  k_sigma > 0: the values are distances; they are turned into Gaussian affinities with per-frame
 * sigmas (mean of the first k_sigma values of a frame in CSC traversal order) and normalised D^-1/2 W D^-1/2
 * (auto_decomp_sparse.cpp:150-198); k_sigma <= 0 and sigma > 0: one global sigma (decomp_sparse.cpp:150-175);
 * both <= 0: the matrix is decomposed as given.
 * Output: the nev algebraically largest eigenvalues, LARGEST FIRST (the order the tool writes them),
 * evecs[nev][n] (unit vectors; the sign of an eigenvector is arbitrary, as with ARPACK), residuals[nev] =
 * |A z - d z| / |d|, the average sigma, and the number of converged pairs.  ARPACK (runARPACK,
 * mdsctk.cpp:857-924) is replaced by a thick-restart Lanczos with the same basis size ncv = 10*nev+1. */
int mdsctk_knn_spectral_decomp(mdsctk_knn_ctx *ctx, int n, const int *pcol, const int *irow, const double *val, int k_sigma,
                               double sigma, int nev, double *evals, double *evecs, double *residuals, double *avg_sigma, int *n_converged);

/* Same with auto_decomp_sparse's -K / --k-perplexity (auto_decomp_sparse.cpp:153-173): k_perplexity > 0 replaces the mean
 * sigmas by ENTROPIC ones (entropic_affinity_sigmas, mdsctk.cpp:388-565): per frame the bandwidth whose Gaussian over its first
 * k_sigma sorted distances has perplexity k_perplexity (1 < k_perplexity < k_sigma <= 256).  sigmas: host double[n] or NULL,
 * receives the per-frame sigmas that were used.  Every frame is solved from the midpoint of its bracket (the reference chains
 * warm starts through the frames in order of their K-th distance), so sigmas agree with it to the solver's tolerance
 * (|perplexity error| < 1e-10), not to the last bit. */
int mdsctk_knn_spectral_decomp_ex(mdsctk_knn_ctx *ctx, int n, const int *pcol, const int *irow, const double *val, int k_sigma,
                                  double sigma, double k_perplexity, int nev, double *evals, double *evecs, double *residuals,
                                  double *avg_sigma, int *n_converged, double *sigmas);

/* ---- producers of knn_data's input: backbone phi/psi angles and their sin/cos embedding -------------
 * xyz: host, float[n_frames][n_atoms][3] (nm), backbone atoms N-CA-C only, one chain
 * (bb_xtc_to_phipsi.cpp:106-122).  T = 2*(n_atoms/3) - 2 angles per frame (radians, torsion() of
 * mdsctk.cpp:643-676 in float, widened to double).  phipsi: host double[n_frames][T] or NULL;
 * sincos: host double[n_frames][2T] (sin, cos interleaved, angles_to_sincos.cpp:107-118) or NULL. */
int mdsctk_knn_phipsi(mdsctk_knn_ctx *ctx, const float *xyz, long long n_frames, int n_atoms, double *phipsi, double *sincos);
/* angles: host double[n]; out: host double[2n] = sin, cos interleaved (angles_to_sincos.cpp:107-118). */
int mdsctk_knn_sincos(mdsctk_knn_ctx *ctx, const double *angles, long long n, double *out);

/* Diagnostic: after set_option("debug_tile", 1) a tensor-core RMSD query also captures the raw
 * TMEM accumulators of (fit tile 0, reference tile 0): out[128][9][48] floats, S_ab of fit row q
 * against reference j at out[q][3*a+b][j]. */
int mdsctk_knn_debug_fetch_tile(mdsctk_knn_ctx *ctx, float *out);

/* Diagnostic: copies one packed array of the RMSD reference set to the host (tests of the pack kernel).
 * which: 0 raw float[n][A][3], 1 G float[n], 2 cen double[n][4], 3 sig float[n][4], 4 Gh float[n], 5 G2 float[n],
 * 6 gres float[n][2], 7 planes float[n][3][A_pad], 8/9 TF32 hi/lo, 10/11 BF16 hi/mid, 12/13 FP16 hi/lo.
 * Returns MDSCTK_KNN_ESTATE when that array is not packed for the current kernel; *n_bytes = its size. */
int mdsctk_knn_debug_fetch_array(mdsctk_knn_ctx *ctx, int which, void *out, size_t out_capacity, size_t *n_bytes);

/* Diagnostic, host arithmetic only (no GPU, no context): the shared-memory / TMEM layout the version-2 RMSD sweep would use for
 * n_atoms atoms.  out12 = { supported (else the streaming kernel runs), wide stages, ring stages, bytes per stage, refine-queue
 * entries per warp, dynamic shared memory bytes, fit-tile chunks in shared memory, fit-tile k-steps in TMEM, stages per pass,
 * k-steps per pass, control-block bytes, shared-memory limit }. */
int mdsctk_knn_debug_rms_layout(int n_atoms, int wide_stages, int *out12);

/* CUDA-event stopwatch on the context's own stream (the stream every kernel of this library
 * is launched on): start records an event, stop records a second one, waits for it and returns
 * the elapsed device time in milliseconds. */
int mdsctk_knn_timer_start(mdsctk_knn_ctx *ctx);
int mdsctk_knn_timer_stop(mdsctk_knn_ctx *ctx, double *elapsed_ms);

/* Copies the k-lists of the last query ([n_fit][k1]) to host memory. */
int mdsctk_knn_fetch(mdsctk_knn_ctx *ctx, double *out_dist, int *out_idx);

/* Exact FP64 distance rows (no selection) computed on the GPU: out[n_fit][n_ref], Angstrom.
 * Diagnostic / spot-check entry point (the tools' --sort false, restricted to a row range). */
int mdsctk_knn_rms_rows(mdsctk_knn_ctx *ctx, long long fit_begin, long long n_fit, int do_fit, double *out);

#ifdef __cplusplus
}
#endif
#endif /* MDSCTK_KNN_H */
