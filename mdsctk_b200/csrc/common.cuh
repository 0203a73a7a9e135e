// common.cuh -- shared declarations of the B200 kNN stage (device layouts, launch entry points).
//
// HBM layout of a frame set (RMSD path).  "frames x 3 x atoms" SoA, frame-major so a
// contiguous block of frames is a contiguous block of bytes in every array (that is what
// makes the row-sharded all-gather a plain concatenation):
//   raw    float  [n][A][3]        decoded coordinates, nm, uncentred (FP64 re-score input)
//   planes float  [n][3][A_pad]    centred, scaled by sqrt(m_a / M); zero padded to A_pad
//   hi, lo float  [n][3][A_pad]    TF32 split of planes: hi = rn_tf32(x), lo = rn_tf32(x - hi)
//   bh, bm bf16   [n][3][A_pad]    BF16 split of planes: bh = rn_bf16(x), bm = rn_bf16(x - bh)
//   fh, fl fp16   [n][3][A_pad]    FP16 split of 64*planes: fh = rn_f16(64x), fl = rn_f16(64x - fh)
//   G      float  [n] (+64 pad)    sum_a (m_a/M) |x_a - c|^2          (fp32 copy for the sweep)
//   Gh, G2 float  [n] (+64 pad)    |fh/64|^2 and |(fh+fl)/64|^2: norms of the ROUNDED structures the
//                                  1xFP16 / 2xFP16 sweeps contract (their E0 uses these, not G)
//   gres   float  [n][2]           residual norms |x - fh/64|, |x - (fh+fl)/64| in nm, rounded up
//   sig    float  [n][4] (+64 pad) singular values of the weighted frame matrix, descending (+ one unused lane):
//                                  RMSD^2(x,y) >= sum_i (sig_i(x) - sig_i(y))^2, tested before the accumulators are read
//   cen    double [n][4]           mass-weighted centroid (x,y,z) and G in FP64
// With the weights normalised to sum 1, min-RMSD^2 = G_q + G_r - 2*lambda_max (nm^2).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef MDSCTK_TC_EXPERIMENTS
#define MDSCTK_TC_EXPERIMENTS 0   // 1: the tcgen05 sweep honours the MDSCTK_TC_DEBUG timing-experiment bits (never in production)
#endif

namespace mdsctk {

int tc_experiment_bits();         // MDSCTK_TC_DEBUG in experiment builds, 0 otherwise (rms_tc.cu)

constexpr float kRmsHalfScale = 64.0f;  // FP16 operand planes hold 64 * sqrt(w) (x - c): keeps the lo part normal
constexpr int kAtomPad = 16;  // A_pad = roundup(A, 16): SIMT k-chunk and 2 x UMMA_K(tf32)=8

inline int pad_atoms(int a) { return (a + kAtomPad - 1) / kAtomPad * kAtomPad; }

struct FrameSetView {
    const float *raw;     // [n][A][3]
    const float *planes;  // [n][3][A_pad]
    const float *G;       // [n]
    const float *Gh, *G2; // [n]     norms of the fp16-rounded structures (1 part / 2 parts)
    const float *gres;    // [n][2]  their distance from the true structure
    const float *sig;     // [n][4]  singular values of the weighted frame (descending) + 0: von Neumann pre-bound
    const void *fh, *fl;  // fp16 [n][3][A_pad]  the rounded operand planes themselves (64 * sqrt(w) (x - c), hi / lo part)
    const double *cen;    // [n][4]
    long long n;
    int A, A_pad;
};

// Per-row candidate lists kept in HBM/L2 while the sweep streams the reference set.
//   key [rows][cap]  approximate d^2 (float) or exact key (double path)
//   idx [rows][cap]  reference index
//   cnt [rows]       entries in use;  tau [rows] current admission threshold
template <typename KeyT>
struct CandLists {
    KeyT *key;
    int *idx;
    int *cnt;
    KeyT *tau;
    int cap;   // list capacity
    int keep;  // entries kept by a compaction (k1 + slack)
    int H;     // lists per fit row (list id = row * H + h): the tensor-core sweep keeps one list per
               // (reference segment, column half) so that no list is shared between threads
};

// ---- launch wrappers (defined in the .cu files) -----------------------------------
cudaError_t launch_pack_frames(const float *raw, const double *mass_norm, long long n, int A, int A_pad,
                               float *planes, float *hi, float *lo, void *bh, void *bm, void *fh, void *fl, float *G,
                               double *cen, float *Gh, float *G2, float *gres, float *sig, cudaStream_t st);

cudaError_t launch_rms_sweep_simt(const FrameSetView &fit, long long fit_begin, long long n_fit,
                                  const FrameSetView &ref, int do_fit, CandLists<float> cl, cudaStream_t st);

// tcgen05 sweep (rms_tc.cu).  *_hi / *_lo: TF32 split planes, same [n][3][A_pad] layout as planes.
// mode: 1 = 3xTF32 (hi/lo fp32 planes), 2 = 1xTF32 (hi only), 3 = 3xBF16 (bh/bm bf16 planes),
//       4 = 3xFP16 (fh/fl), 5 = 2xFP16 (fit fh only, reference fh/fl), 6 = 1xFP16 (fh only).
// cl.H must be rms_tc_lists_per_segment() * n_seg (reference segments x column groups).
// own_tile_scratch: NULL when the fit rows are reference frames fit_begin..; else device ints, one per 256 fit rows.
cudaError_t launch_rms_sweep_tc(int mode, const FrameSetView &fit, const void *fit_hi, const void *fit_lo,
                                long long fit_begin, long long n_fit, const FrameSetView &ref, const void *ref_hi,
                                const void *ref_lo, int do_fit, int n_seg, CandLists<float> cl, float *row_tau,
                                float g_ref_max, int *own_tile_scratch, float *debug_tile, int n_sms, cudaStream_t st);
// Second-generation 1xFP16 sweep (rms_tc2.cu): fit tile resident in shared memory + TMEM, reference-only ring, pass
// director.  Same lists / arguments as launch_rms_sweep_tc with mode 6; supported when the fit tile fits (A_pad <= 304).
bool rms_tc2_supported(int A_pad);
void rms_tc2_layout_info(int A_pad, int want_wide, int out[12]);
cudaError_t launch_rms_sweep_tc2(const FrameSetView &fit, long long fit_begin, long long n_fit, const FrameSetView &ref, int do_fit,
                                 int n_seg, CandLists<float> cl, float *row_tau, float g_ref_max, int *own_tile_scratch,
                                 float *debug_tile, int wide_stages, int n_sms, cudaStream_t st);
// reference fp16 planes re-ordered so that every ring stage of the sweep is one contiguous block (optional; NULL = frame-major planes)
cudaError_t launch_rms_guess_own_tile(const float4 *q_sig, long long q_begin, long long n_q, const float4 *r_sig, long long n_r,
                                      int *own_tile, cudaStream_t st);
int rms_tc_choose_segments(long long n_fit, long long n_ref, int n_sms);
int rms_tc_lists_per_segment();
int rms_tc_list_stride(int keep);

// FP64 re-score of the kept candidates, final (distance, index) sort, certificate.
//   out_dist [n_fit][k1] (Angstrom), out_idx [n_fit][k1], flags[n_fit] (1 = certified),
//   err_stats: device double[2] = {max |approx - exact| d^2, max per-row spread of (approx - exact)};
//   n_bad: device int counter of uncertified rows, listed in bad_rows.
//   fit_part / gres_ref_max: operand-rounding term of the certificate for the 2xFP16 / 1xFP16 sweeps
//   (fit_part = column of fit.gres the sweep's fit operand corresponds to, -1 = none; ref_parts = fp16 parts of
//   the reference operand, 1 or 2; gres_ref_max = largest residual norm of the reference operand), see rms_rescore.cu.
cudaError_t launch_rms_rescore(const FrameSetView &fit, long long fit_begin, long long n_fit,
                               const FrameSetView &ref, const double *mass_norm, int do_fit,
                               CandLists<float> cl, int k1, double eps_scale, float g_ref_max,
                               int fit_part, int ref_parts, float gres_ref_max, double *out_dist, int *out_idx, int *flags, double *err_stats, int *n_bad,
                               int *bad_rows, cudaStream_t st);

// Exact FP64 d^2 of fit row(s) against every reference frame: out[n_rows][n_ref] (nm^2).
cudaError_t launch_rms_exact_rows(const FrameSetView &fit, const int *row_ids, long long fit_begin,
                                  int n_rows, const FrameSetView &ref, const double *mass_norm, int do_fit,
                                  double *out_d2, cudaStream_t st);

// Exact top-k1 of full rows of doubles: rows_d2[n_rows][n_ref] -> out (written at row_ids[r]).
cudaError_t launch_select_rows_f64(const double *rows_d2, int n_rows, long long n_ref, int k1,
                                   const int *row_ids, double scale_sqrt, double *out_dist, int *out_idx,
                                   cudaStream_t st);

// Vector path (FP64 exact sweep).
cudaError_t launch_data_rowstats(const double *rows, long long n, int dim, double *stats, cudaStream_t st);
cudaError_t launch_data_sweep(const double *fit, const double *fit_stats, long long n_fit, const double *ref,
                              const double *ref_stats, long long n_ref, int dim, int metric,
                              CandLists<double> cl, cudaStream_t st);
cudaError_t launch_data_exact_rows(const double *fit, const double *fit_stats, long long n_fit, const double *ref,
                                   const double *ref_stats, long long n_ref, int dim, int metric, double *out, cudaStream_t st);
cudaError_t launch_data_finalize(CandLists<double> cl, long long n_fit, int k1, double *out_dist, int *out_idx,
                                 cudaStream_t st);

// Vector path, tensor-core filter + exact FP64 re-score (data_tc.cu).
cudaError_t launch_data_maxabs(const double *v, size_t n, double *out, cudaStream_t st);
// norm1 / gres (may be NULL): |hi|^2 (scaled units) and |x - hi/scale| (input units, rounded up) for the one-part filter
// stats != NULL: pack the STANDARDISED rows (v - mean) / (spread sqrt(dim - 1)) (correlation metric)
cudaError_t launch_data_pack(const double *rows, long long n, int dim, int D_pad, double scale, const double *stats, void *hi,
                             void *lo, float *norm, float *norm1, float *gres, cudaStream_t st);
cudaError_t launch_data_stats_check(const double *stats, long long n, int *bad, cudaStream_t st);   // rows with a zero / non-finite spread
int data_tc_pad_dim(int dim);
void data_tc_set_streaming(bool on);   // test hook: force the streaming mode of the one-part knn_data filter (default: resident fit tile)
int data_tc_list_stride(int keep);
int data_tc_choose_segments(long long n_fit, long long n_ref, int n_sms);
// fit_begin_in_ref: reference index of fit row 0 when the fit rows are reference rows, else -1
cudaError_t launch_data_sweep_tc(const void *fit_hi, const void *fit_lo, const float *fit_norm, long long n_fit,
                                 long long fit_begin_in_ref, const void *ref_hi, const void *ref_lo, const float *ref_norm,
                                 long long n_ref, int D_pad, double scale, int n_seg, CandLists<float> cl, float *row_tau,
                                 int one_part, int n_sms, cudaStream_t st);
cudaError_t launch_data_rescore(const double *fit, const double *ref, long long n_fit, int dim, int k1, CandLists<float> cl,
                                double eps_rel, const float *q_norm, double scale, float r_norm_max, const float *q_g,
                                float g_ref_max, const double *fit_stats, const double *ref_stats, double *out_dist, int *out_idx,
                                int *flags, double *err_stats, int *n_bad, int *bad_rows, cudaStream_t st);
cudaError_t launch_data_gather_rows(const double *src, const int *ids, int n_rows, int dim, double *dst, cudaStream_t st);
cudaError_t launch_data_scatter_out(const double *d_src, const int *i_src, const int *ids, int n_rows, int k1, double *d_dst,
                                    int *i_dst, cudaStream_t st);

// Symmetric CSC matrix from kNN lists (csc.cu; replaces make_sysparse.cpp:245-329).
// mode 0 make_sysparse, 1 make_gesparse, 2 make_gesparse -s
cudaError_t launch_csc_build(int mode, const int *d_idx, const double *d_dist, long long n, int maxk, int k, int *cnt, int *cur,
                             int *off, int *fin, int *pcol, int *scan_tmp, unsigned long long *seg_key, double *seg_val,
                             int *irow, double *val, cudaStream_t st);

// out[i] = sum in[0..i) (csc.cu); out may alias in; tmp: n/1024 + n/1024^2 + 8 ints
cudaError_t exclusive_scan(const int *in, int *out, long long n, int *tmp, cudaStream_t st);

// Spectral stage (spectral.cu; replaces auto_decomp_sparse.cpp:150-198 and runARPACK, mdsctk.cpp:857-924).
cudaError_t launch_spectral_adjacency(int n, int nnz, const int *pcol, const int *irow, int *deg_ptr, int *n_as_row, int *cur,
                                      int *scan_tmp, int *adj_other, int *adj_pos,
                                      cudaError_t (*scan)(const int *, int *, long long, int *, cudaStream_t), cudaStream_t st);
cudaError_t launch_spectral_affinity(int n, const int *pcol, const int *irow, const int *ptr, const int *adj_pos, int k_a, double sigma0,
                                     double perplexity, double *M, double *sigma, double *dinv, cudaStream_t st);
cudaError_t launch_spectral_spmv(int n, const int *ptr, const int *adj_other, const int *adj_pos, const double *M, const double *x,
                                 double *y, cudaStream_t st);
cudaError_t spectral_lanczos(int n, const int *ptr, const int *adj_other, const int *adj_pos, const double *M, int nev, int ncv,
                             int max_restarts, double tol, double *V, double *W, double *small, double *evals, double *d_evecs,
                             double *residuals, int *n_conv, int *n_restart, int *n_spmv, cudaStream_t st);

// Featurisers (featurize.cu): backbone torsions and the sin/cos embedding.
cudaError_t launch_phipsi(const float *xyz, long long n, int A, double *phipsi, double *sincos, cudaStream_t st);
cudaError_t launch_sincos(const double *angles, long long n, double *out, cudaStream_t st);

cudaError_t launch_max_float(const float *v, long long n, float *out, cudaStream_t st);
cudaError_t launch_max_float_strided(const float *v, long long n, int stride, float *out, cudaStream_t st);   // max v[i*stride]
cudaError_t launch_fill_u32(void *p, size_t n, uint32_t v, cudaStream_t st);
cudaError_t launch_iota_i32(int *p, int n, int start, int stride, cudaStream_t st);   // p[i] = start + i * stride
cudaError_t launch_audit_compare(const double *out_dist, const int *out_idx, const int *row_ids, const double *ex_dist,
                                 const int *ex_idx, int n_rows, int k1, int *mismatches, cudaStream_t st);

}  // namespace mdsctk
