// abi.cu -- the C ABI of include/mdsctk_knn.h: context, device memory, phase orchestration.
//
// Host-side counterpart of the reference's main() bodies between "load" and "write":
// knn_rms.cpp:181-293 and knn_data.cpp:141-250.  No CPU arithmetic happens here -- every
// distance, selection and sort runs in the kernels of this directory; this file only moves
// bytes, launches, and checks errors.  There is no CPU fallback.
#include "../../include/mdsctk_knn.h"
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace mdsctk;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    cudaError_t reserve(size_t want)
    {
        if (want <= bytes) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; bytes = 0; }
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) bytes = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// Operand families of a packed frame set.  raw, G, cen, sig (and the small Gh / G2 / gres scalars) always exist;
// the operand planes are written only for the sweep kernel in use (pack.cu), so the default 1xFP16 run holds
// raw + fh = 5.4 KB per 300-atom frame instead of 22 KB, and the all-gather of a sharded run moves a quarter.
enum { F_PLANES = 1, F_TF32 = 2, F_BF16 = 4, F_FP16 = 8, F_FP16LO = 16 };

int families_of_kernel(int rms_kernel)
{
    switch (rms_kernel) {
    case MDSCTK_KNN_RMS_SIMT_FP32: return F_PLANES;
    case MDSCTK_KNN_RMS_TC_3XTF32:
    case MDSCTK_KNN_RMS_TC_1XTF32: return F_TF32;
    case MDSCTK_KNN_RMS_TC_3XBF16: return F_BF16;
    case MDSCTK_KNN_RMS_TC_3XFP16:
    case MDSCTK_KNN_RMS_TC_2XFP16: return F_FP16 | F_FP16LO;
    default: return F_FP16;
    }
}

struct FrameSet {
    DevBuf raw, planes, hi, lo, bh, bm, fh, fl, G, cen, Gh, G2, gres, sig;
    long long n = 0;
    int A = 0, A_pad = 0;
    int have = 0;        // operand families packed for every frame of the set
    FrameSetView view() const
    {
        FrameSetView v;
        v.raw = raw.as<float>(); v.planes = planes.as<float>(); v.G = G.as<float>(); v.cen = cen.as<double>();
        v.Gh = Gh.as<float>(); v.G2 = G2.as<float>(); v.gres = gres.as<float>(); v.fh = fh.p; v.fl = fl.p; v.sig = sig.as<float>();
        v.n = n; v.A = A; v.A_pad = A_pad;
        return v;
    }
    cudaError_t alloc(long long n_, int A_)
    {
        if (n_ != n || A_ != A) have = 0;
        n = n_; A = A_; A_pad = pad_atoms(A_);
        cudaError_t e;
        if ((e = raw.reserve((size_t)n * A * 12)) != cudaSuccess) return e;
        if ((e = G.reserve((size_t)(n + 64) * 4)) != cudaSuccess) return e;  // tensor-core epilogue reads G in 48-wide tiles
        if ((e = Gh.reserve((size_t)(n + 64) * 4)) != cudaSuccess) return e;
        if ((e = G2.reserve((size_t)(n + 64) * 4)) != cudaSuccess) return e;
        if ((e = gres.reserve((size_t)n * 8)) != cudaSuccess) return e;
        if (sig.bytes < (size_t)(n + 64) * 16) {     // read in 48-frame tiles: the padding must be zero, not garbage
            if ((e = sig.reserve((size_t)(n + 64) * 16)) != cudaSuccess) return e;
            if ((e = cudaMemset(sig.p, 0, (size_t)(n + 64) * 16)) != cudaSuccess) return e;
        }
        return cen.reserve((size_t)n * 32);
    }
    cudaError_t reserve_families(int fam)
    {
        cudaError_t e;
        const size_t p4 = (size_t)n * 3 * A_pad * 4, p2 = (size_t)n * 3 * A_pad * 2;
        if ((fam & F_PLANES) && (e = planes.reserve(p4)) != cudaSuccess) return e;
        if ((fam & F_TF32) && ((e = hi.reserve(p4)) != cudaSuccess || (e = lo.reserve(p4)) != cudaSuccess)) return e;
        if ((fam & F_BF16) && ((e = bh.reserve(p2)) != cudaSuccess || (e = bm.reserve(p2)) != cudaSuccess)) return e;
        if ((fam & F_FP16) && (e = fh.reserve(p2)) != cudaSuccess) return e;
        if ((fam & F_FP16LO) && (e = fl.reserve(p2)) != cudaSuccess) return e;
        return cudaSuccess;
    }
    void release()
    {
        raw.release(); planes.release(); hi.release(); lo.release(); bh.release(); bm.release(); fh.release(); fl.release(); G.release();
        cen.release(); Gh.release(); G2.release(); gres.release(); sig.release(); n = 0; have = 0;
    }
};

struct PhaseTimer {
    cudaEvent_t a = nullptr, b = nullptr;
    void init() { cudaEventCreate(&a); cudaEventCreate(&b); }
    void destroy() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    void start(cudaStream_t s) { cudaEventRecord(a, s); }
    double stop(cudaStream_t s)
    {
        cudaEventRecord(b, s);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        return (double)ms;
    }
};

}  // namespace

struct mdsctk_knn_ctx {
    int dev = 0;
    cudaStream_t st = nullptr;
    std::string err;
    PhaseTimer tm, user_tm;
    // options
    // 1xFP16: one MMA per k-step on fp16-rounded operands; the sweep then measures the exact min-RMSD between the
    // ROUNDED structures, which the metric triangle inequality ties to the true distance within the rounding
    // residual norms recorded by pack.cu -- a rigorous term in the re-score certificate, so the result is still
    // index-exact.  1.9x the sweep rate of 3xFP16 (the previous default: three MMAs, 22 operand bits).
    int rms_kernel = MDSCTK_KNN_RMS_TC_1XFP16;
    bool rms_kernel_set = false;   // chosen by the caller: never substituted
    long long slack = -1;
    long long cert_scale_ppm = 1000000;
    // RMSD state
    FrameSet ref, fit;
    DevBuf wnorm;  // double[2A]: m_a / M, then sqrt(m_a / M)
    bool have_ref = false, gmax_dirty = true;
    int ref_pack_fam = F_FP16;   // operand families every shard of the current reference set is packed with (fixed at alloc)
    long long chunk_rows = 0;        // fit rows per internal row block of a query; 0 = as many as 6 GB of candidate lists allow
    long long audit_rows = 8;        // certified rows per row block recomputed exactly and compared (0 = off)
    bool force_exact = false;        // every row through the exact FP64 path (test hook)
    // ring-stage-ordered copy of the reference fp16 planes (rms_tc2.cu).  Off by default: measured neutral (C3 63.3 vs 63.6 ms,
    // C4 block 866 vs 873 ms) -- the ring's copy latency does not come from the 72 strided rows of a stage -- and it costs 1.9 GB
    bool rms_wide_stages = true;
    int data_segments = 0;           // same for the knn_data tensor filter
    int rms_segments = 0;            // reference segments per fit super-tile (0 = chosen by rms_tc_choose_segments)
    int sweep_version = 2;           // 1xFP16 sweep: 2 = rms_tc2.cu where the fit tile fits (default), 1 = rms_tc.cu
    DevBuf audit_ids, audit_seq, audit_dist, audit_idx;
    float g_ref_max = 0.f;
    float gres_ref_max[2] = {0.f, 0.f};   // largest rounding residual norm of the reference set (1 / 2 fp16 parts)
    // vector state
    DevBuf d_ref, d_fit, d_ref_stats, d_fit_stats;
    long long dn_ref = 0;
    int ddim = 0;
    bool have_dref = false, dstats_dirty = true;
    // tensor-core filter state of the vector path (data_tc.cu)
    DevBuf dt_ref_hi, dt_ref_lo, dt_ref_norm, dt_fit_hi, dt_fit_lo, dt_fit_norm;
    DevBuf dt_ref_norm1, dt_ref_g, dt_fit_norm1, dt_fit_g;   // one-part filter: norms of the hi parts, rounding residual norms
    float dt_g_ref_max = 0.f;
    DevBuf fb_rows, fb_key, fb_idx, fb_cnt, fb_tau, fb_dist, fb_oidx, fb_stats;
    int dt_metric = MDSCTK_KNN_EUCLIDEAN;   // metric the packed filter operands were built for
    bool dpack_dirty = true;
    double dt_scale = 1.0, dt_ref_maxabs = 0.0;
    float dt_rnorm_max = 0.f;
    int data_kernel = -1;      // -1 auto (= 2 for large Euclidean inputs), 0 exact FP64 sweep, 1 tensor-core filter (3xFP16) + exact re-score, 2 one-part filter (1xFP16)
    // CSC builder state (csc.cu)
    DevBuf c_idx, c_dist, c_ints, c_key, c_val, c_irow, c_oval;
    DevBuf f_in, f_ang, f_sc;   // featuriser buffers
    DevBuf s_int, s_val, s_vec, s_basis, s_rot, s_small, s_evec;   // spectral stage
    long long c_nnz = -1;
    // scratch + results
    DevBuf cand_key, cand_idx, cand_cnt, cand_tau, flags, bad_rows, scalars, rows_buf;
    DevBuf out_dist, out_idx, debug_tile, row_tau, own_tile;
    bool debug_tile_on = false;
    int n_sms = 148;
    long long out_rows = 0, out_off = 0;
    int out_k1 = 0;
    mdsctk_knn_stats stats;
};

namespace {

int fail(mdsctk_knn_ctx *c, int code, const std::string &msg)
{
    if (c) c->err = msg;
    return code;
}
int fail_cuda(mdsctk_knn_ctx *c, cudaError_t e, const char *what)
{
    cudaGetLastError();  // clear sticky-less error state
    return fail(c, e == cudaErrorMemoryAllocation ? MDSCTK_KNN_ENOMEM : MDSCTK_KNN_ECUDA,
                std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call, what)                                               \
    do {                                                             \
        cudaError_t e__ = (call);                                    \
        if (e__ != cudaSuccess) return fail_cuda(ctx, e__, what);    \
    } while (0)

struct Bind {
    explicit Bind(mdsctk_knn_ctx *c) { cudaSetDevice(c->dev); }
};

__global__ void sqrt_scale_kernel(double *v, size_t n, double scale)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        v[i] = sqrt(v[i]) * scale;
}

int upload_weights(mdsctk_knn_ctx *ctx, const float *mass, int A)
{
    std::vector<double> w(A);
    double M = 0.0;
    for (int i = 0; i < A; i++) M += (double)mass[i];
    if (!(M > 0.0)) return fail(ctx, MDSCTK_KNN_EINVAL, "masses must sum to a positive number");
    for (int i = 0; i < A; i++) w[i] = (double)mass[i] / M;
    // [A] weights m/M followed by [A] of their square roots (pack.cu scales every atom by sqrt(w): one FP64 sqrt per atom
    // and frame was a third of the pack kernel's time)
    w.resize(2 * (size_t)A);
    for (int i = 0; i < A; i++) w[(size_t)A + i] = std::sqrt(w[i]);
    CK(ctx->wnorm.reserve((size_t)A * 16), "cudaMalloc(weights)");
    CK(cudaMemcpyAsync(ctx->wnorm.p, w.data(), (size_t)A * 16, cudaMemcpyHostToDevice, ctx->st), "H2D weights");
    CK(cudaStreamSynchronize(ctx->st), "sync weights");
    return 0;
}

// Packs frames [off, off+n) of fs from its raw array for the operand families `fam` (F_FP16LO implies F_FP16:
// the lo part is the remainder of the hi part).
int pack_launch(mdsctk_knn_ctx *ctx, FrameSet &fs, long long off, long long n, int fam)
{
    if (fam & F_FP16LO) fam |= F_FP16;
    CK(fs.reserve_families(fam), "cudaMalloc(operand planes)");
    const size_t po = (size_t)off * 3 * fs.A_pad;
    CK(launch_pack_frames(fs.raw.as<float>() + (size_t)off * fs.A * 3, ctx->wnorm.as<double>(), n, fs.A, fs.A_pad,
                          (fam & F_PLANES) ? fs.planes.as<float>() + po : nullptr,
                          (fam & F_TF32) ? fs.hi.as<float>() + po : nullptr, (fam & F_TF32) ? fs.lo.as<float>() + po : nullptr,
                          (fam & F_BF16) ? fs.bh.as<uint16_t>() + po : nullptr, (fam & F_BF16) ? fs.bm.as<uint16_t>() + po : nullptr,
                          (fam & F_FP16) ? fs.fh.as<uint16_t>() + po : nullptr, (fam & F_FP16LO) ? fs.fl.as<uint16_t>() + po : nullptr,
                          fs.G.as<float>() + off, fs.cen.as<double>() + 4 * off, fs.Gh.as<float>() + off, fs.G2.as<float>() + off,
                          fs.gres.as<float>() + 2 * off, fs.sig.as<float>() + 4 * off, ctx->st), "pack_frames");
    return 0;
}

int pack_into(mdsctk_knn_ctx *ctx, FrameSet &fs, const float *xyz, long long off, long long n, int fam)
{
    const int A = fs.A;
    ctx->tm.start(ctx->st);
    CK(cudaMemcpyAsync(fs.raw.as<float>() + (size_t)off * A * 3, xyz, (size_t)n * A * 12, cudaMemcpyHostToDevice,
                       ctx->st), "H2D frames");
    ctx->stats.ms_upload += ctx->tm.stop(ctx->st);
    ctx->tm.start(ctx->st);
    const int rc = pack_launch(ctx, fs, off, n, fam);
    if (rc) return rc;
    ctx->stats.ms_pack += ctx->tm.stop(ctx->st);
    fs.have |= fam | ((fam & F_FP16LO) ? F_FP16 : 0);
    return 0;
}

// The sweep kernel of a query was chosen after the set was packed (set_option, or the fp16-overflow substitution):
// write the missing operand planes from the raw frames, which are always resident.
int ensure_families(mdsctk_knn_ctx *ctx, FrameSet &fs, int fam)
{
    const int missing = fam & ~fs.have;
    if (!missing) return 0;
    ctx->tm.start(ctx->st);
    const int rc = pack_launch(ctx, fs, 0, fs.n, missing);
    if (rc) return rc;
    ctx->stats.ms_pack += ctx->tm.stop(ctx->st);
    fs.have |= missing;
    return 0;
}

void choose_lists(const mdsctk_knn_ctx *ctx, int rms_kernel, int k1, int *keep, int *cap)
{
    // The admission threshold is the keep-th smallest approximate distance of the row, so `slack`
    // is what the adaptive re-score can fall back on when neighbours sit on a plateau of nearly
    // equal distances (thermal noise): measured on the 100k x 300 workload, slack >= 48 certifies
    // every row with the 1xFP16 certificate (>= 96 with the uniform noise bound of the 3x modes), C4's most
    // extended basin needs 2 k1 at k = 64, and the sweep costs ~4% more per extra 32 kept candidates.
    long long slack = ctx->slack >= 0 ? ctx->slack : std::max<long long>(rms_kernel >= MDSCTK_KNN_RMS_TC_2XFP16 ? 64 : 96, 2LL * k1);
    // the re-score works on at most 2048 candidates per row: large k (k = 1024 in examples/mld/figure-11.bash:41) keeps
    // what fits; rows that cannot be certified with that little slack are redone exactly, as always
    slack = std::min<long long>(slack, std::max<long long>(0, 2040 - (long long)k1));
    long long kp = ((long long)k1 + slack + 7) / 8 * 8;
    *keep = (int)kp;
    *cap = (int)((kp + 128 + 31) / 32 * 32);
}

// Bound on the NOISE of the filter, |(approx d^2 - exact d^2) - row-common bias|, as a fraction of
// E0 = (Gq+Gr)/2 (DESIGN.md "certificate").  The contraction error grows like sqrt(atoms) * 2^-24
// relative to E0; the constants are ~2x the largest half-spread observed over >1e6 candidates
// (stats.max_filter_spread reports it, and the certificate re-checks it on every row).
double default_eps_scale(int rms_kernel, int n_atoms)
{
    const double sa = std::sqrt((double)std::max(n_atoms, 128));
    switch (rms_kernel) {
    case MDSCTK_KNN_RMS_TC_1XTF32: return 4e-5 * sa;
    case MDSCTK_KNN_RMS_TC_3XBF16: return 1.5e-5;       // bf16 split residual 2^-18 per product dominates
    // the reduced FP16 modes bound the operand rounding separately and rigorously (gres, rms_rescore.cu); what is
    // left is the fp32 accumulation error of 19-38 MMAs per accumulator, which the re-score MEASURES on every
    // row (key - exact distance of the rounded structures, stats.max_filter_spread): largest half-spread seen
    // over 1.6e6 candidates 1.8e-5 nm^2 at E0 = 17.5 nm^2 (1.0e-6 E0); the bound is 2.5x that
    case MDSCTK_KNN_RMS_TC_2XFP16:
    case MDSCTK_KNN_RMS_TC_1XFP16: return 1.5e-7 * sa;
    case MDSCTK_KNN_RMS_TC_3XFP16:
    case MDSCTK_KNN_RMS_TC_3XTF32:
    default: return 5e-7 * sa;
    }
}

struct RmsPlan {
    int rms_kernel, keep, cap, n_seg, H, k1, do_fit;
    bool use_tc, oos;
};

// One row block of a query: sweep -> FP64 re-score + certificate -> exact rows for what could not be certified.
// Results land at out_dist / out_idx rows [row_off, row_off + n_fit) of the context's output buffers.
int rms_run_block(mdsctk_knn_ctx *ctx, FrameSet &fitset, const RmsPlan &P, long long fit_begin, long long n_fit, long long row_off)
{
    const FrameSetView ref = ctx->ref.view(), fit = fitset.view();
    mdsctk_knn_stats &S = ctx->stats;
    const int k1 = P.k1, do_fit = P.do_fit, rms_kernel = P.rms_kernel, H = P.H;
    CandLists<float> cl;
    cl.key = ctx->cand_key.as<float>(); cl.idx = ctx->cand_idx.as<int>(); cl.cnt = ctx->cand_cnt.as<int>();
    cl.tau = ctx->cand_tau.as<float>(); cl.cap = P.cap; cl.keep = P.keep; cl.H = H;
    double *d_err = ctx->scalars.as<double>() + 1;   // {max |err|, max spread}
    int *d_nbad = ctx->scalars.as<int>() + 8;
    double *o_dist = ctx->out_dist.as<double>() + (size_t)row_off * k1;
    int *o_idx = ctx->out_idx.as<int>() + (size_t)row_off * k1;
    CK(cudaMemsetAsync(ctx->scalars.p, 0, 64, ctx->st), "memset scalars");

    // ---- sweep: all pairs -> k1+slack candidates per row ------------------------------------
    ctx->tm.start(ctx->st);
    switch (rms_kernel) {
    case MDSCTK_KNN_RMS_SIMT_FP32:
        CK(launch_rms_sweep_simt(fit, fit_begin, n_fit, ref, do_fit, cl, ctx->st), "rms_sweep_simt");
        break;
    default:
        CK(launch_fill_u32(ctx->row_tau.p, (size_t)n_fit, 0x7f800000u, ctx->st), "fill row_tau");  // +inf
        {
            const void *q_hi = fitset.hi.p, *q_lo = fitset.lo.p, *r_hi = ctx->ref.hi.p, *r_lo = ctx->ref.lo.p;
            if (rms_kernel == MDSCTK_KNN_RMS_TC_3XBF16) {
                q_hi = fitset.bh.p; q_lo = fitset.bm.p; r_hi = ctx->ref.bh.p; r_lo = ctx->ref.bm.p;
            } else if (rms_kernel == MDSCTK_KNN_RMS_TC_3XFP16 || rms_kernel == MDSCTK_KNN_RMS_TC_2XFP16 ||
                       rms_kernel == MDSCTK_KNN_RMS_TC_1XFP16) {
                q_hi = fitset.fh.p; q_lo = fitset.fl.p; r_hi = ctx->ref.fh.p; r_lo = ctx->ref.fl.p;
                if (rms_kernel == MDSCTK_KNN_RMS_TC_1XFP16) { q_lo = q_hi; r_lo = r_hi; }    // no second part: never read
            }
            if (rms_kernel == MDSCTK_KNN_RMS_TC_1XFP16 && ctx->sweep_version != 1 && rms_tc2_supported(ref.A_pad)) {
                CK(launch_rms_sweep_tc2(fit, fit_begin, n_fit, ref, do_fit, P.n_seg, cl, ctx->row_tau.as<float>(), ctx->g_ref_max,
                                        P.oos ? ctx->own_tile.as<int>() : nullptr,
                                        ctx->debug_tile_on ? ctx->debug_tile.as<float>() : nullptr,
                                        ctx->rms_wide_stages ? 1 : 0, ctx->n_sms, ctx->st),
                   "rms_sweep_tc2");
                S.sweep_version = 2;
            } else {
                CK(launch_rms_sweep_tc(rms_kernel, fit, q_hi, q_lo, fit_begin, n_fit, ref, r_hi, r_lo, do_fit, P.n_seg, cl,
                                       ctx->row_tau.as<float>(), ctx->g_ref_max, P.oos ? ctx->own_tile.as<int>() : nullptr,
                                       ctx->debug_tile_on ? ctx->debug_tile.as<float>() : nullptr, ctx->n_sms, ctx->st),
                   "rms_sweep_tc");
                S.sweep_version = 1;
            }
        }
        break;
    }
    S.launches += 1;
    S.ms_sweep += ctx->tm.stop(ctx->st);
    CK(cudaGetLastError(), "sweep kernel");

    // ---- FP64 re-score + certificate -----------------------------------------------------------
    const double eps_scale = default_eps_scale(rms_kernel, ref.A) * (double)ctx->cert_scale_ppm * 1e-6;
    S.cert_eps = eps_scale * 0.5 * (double)ctx->g_ref_max * 2.0;
    ctx->tm.start(ctx->st);
    // operand-rounding term: 2xFP16 contracts fit fh with reference fh+fl, 1xFP16 fh with fh
    const int fit_part = (rms_kernel == MDSCTK_KNN_RMS_TC_2XFP16 || rms_kernel == MDSCTK_KNN_RMS_TC_1XFP16) ? 0 : -1;
    const float gres_ref = rms_kernel == MDSCTK_KNN_RMS_TC_1XFP16 ? ctx->gres_ref_max[0]
                           : (rms_kernel == MDSCTK_KNN_RMS_TC_2XFP16 ? ctx->gres_ref_max[1] : 0.0f);
    S.cert_gres = fit_part >= 0 ? (double)gres_ref : 0.0;
    CK(launch_rms_rescore(fit, fit_begin, n_fit, ref, ctx->wnorm.as<double>(), do_fit, cl, k1, eps_scale,
                          ctx->g_ref_max, fit_part, rms_kernel == MDSCTK_KNN_RMS_TC_2XFP16 ? 2 : 1, gres_ref, o_dist, o_idx, ctx->flags.as<int>(),
                          d_err, d_nbad, ctx->bad_rows.as<int>(), ctx->st), "rms_rescore");
    S.launches += 1;
    struct { double pad, err, spread, done_max; int nbad; } host_sc;
    CK(cudaMemcpyAsync(&host_sc, ctx->scalars.p, sizeof(host_sc), cudaMemcpyDeviceToHost, ctx->st), "D2H scalars");
    S.ms_rescore += ctx->tm.stop(ctx->st);
    CK(cudaGetLastError(), "rescore kernel");
    S.max_filter_err = std::max(S.max_filter_err, host_sc.err);
    S.max_filter_spread = std::max(S.max_filter_spread, host_sc.spread);
    S.rescored_max = std::max(S.rescored_max, (int)host_sc.done_max);
    S.fallback_rows += host_sc.nbad;

    // ---- rows whose certificate failed: exact FP64 rows + exact selection -------------------
#if MDSCTK_TC_EXPERIMENTS
    if (tc_experiment_bits() & (1 | 64)) host_sc.nbad = 0;   // timing experiments that skip the QCP leave every row uncertified
#endif
    if (ctx->force_exact) {
        S.fallback_rows += n_fit - host_sc.nbad;
        host_sc.nbad = (int)n_fit;
        CK(launch_iota_i32(ctx->bad_rows.as<int>(), (int)n_fit, 0, 1, ctx->st), "iota(bad_rows)");
    }
    if (host_sc.nbad > 0) {
        ctx->tm.start(ctx->st);
        const size_t row_bytes = (size_t)ref.n * 8;
        int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)host_sc.nbad, ((size_t)256 << 20) / row_bytes));
        CK(ctx->rows_buf.reserve((size_t)chunk * row_bytes), "cudaMalloc(rows_buf)");
        for (int off = 0; off < host_sc.nbad; off += chunk) {
            const int nr = std::min(chunk, host_sc.nbad - off);
            CK(launch_rms_exact_rows(fit, ctx->bad_rows.as<int>() + off, fit_begin, nr, ref, ctx->wnorm.as<double>(),
                                     do_fit, ctx->rows_buf.as<double>(), ctx->st), "rms_exact_rows");
            CK(launch_select_rows_f64(ctx->rows_buf.as<double>(), nr, ref.n, k1, ctx->bad_rows.as<int>() + off, 10.0,
                                      o_dist, o_idx, ctx->st), "select_rows");
            S.launches += 2;
        }
        S.ms_fallback += ctx->tm.stop(ctx->st);
        CK(cudaGetLastError(), "fallback kernels");
    }

    // ---- audit: a sample of rows recomputed through the exact path must agree -------------------
    // The certificate's accumulation-noise term is a measured bound, not a derived one; this
    // check does not depend on it: full FP64 rows + exact selection for rows spread over the block, compared with
    // what the filter + certificate produced.  Costs ~1 % of a block; a mismatch is an error, not a statistic.
    const int na = ctx->force_exact ? 0 : (int)std::min<long long>(ctx->audit_rows, n_fit);
    if (na > 0) {
        ctx->tm.start(ctx->st);
        CK(ctx->audit_ids.reserve((size_t)na * 4), "cudaMalloc(audit_ids)");
        CK(ctx->audit_seq.reserve((size_t)na * 4), "cudaMalloc(audit_seq)");
        CK(ctx->audit_dist.reserve((size_t)na * k1 * 8), "cudaMalloc(audit_dist)");
        CK(ctx->audit_idx.reserve((size_t)na * k1 * 4), "cudaMalloc(audit_idx)");
        const int stride = (int)std::max<long long>(1, n_fit / na);
        const int start = (int)((fit_begin / 7 + row_off / 3) % stride);      // a different residue in every block
        CK(launch_iota_i32(ctx->audit_ids.as<int>(), na, start, stride, ctx->st), "iota(audit_ids)");
        CK(launch_iota_i32(ctx->audit_seq.as<int>(), na, 0, 1, ctx->st), "iota(audit_seq)");
        int *d_mis = ctx->scalars.as<int>() + 12;
        const size_t row_bytes = (size_t)ref.n * 8;
        const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)na, ((size_t)256 << 20) / row_bytes));
        CK(ctx->rows_buf.reserve((size_t)chunk * row_bytes), "cudaMalloc(rows_buf)");
        for (int off = 0; off < na; off += chunk) {
            const int nr = std::min(chunk, na - off);
            CK(launch_rms_exact_rows(fit, ctx->audit_ids.as<int>() + off, fit_begin, nr, ref, ctx->wnorm.as<double>(),
                                     do_fit, ctx->rows_buf.as<double>(), ctx->st), "rms_exact_rows(audit)");
            CK(launch_select_rows_f64(ctx->rows_buf.as<double>(), nr, ref.n, k1, ctx->audit_seq.as<int>() + off, 10.0,
                                      ctx->audit_dist.as<double>(), ctx->audit_idx.as<int>(), ctx->st), "select_rows(audit)");
            S.launches += 2;
        }
        CK(launch_audit_compare(o_dist, o_idx, ctx->audit_ids.as<int>(), ctx->audit_dist.as<double>(), ctx->audit_idx.as<int>(),
                                na, k1, d_mis, ctx->st), "audit_compare");
        int mis = 0;
        CK(cudaMemcpyAsync(&mis, d_mis, 4, cudaMemcpyDeviceToHost, ctx->st), "D2H audit");
        S.ms_fallback += ctx->tm.stop(ctx->st);
        S.launches += 1;
        S.audit_rows += na;
        S.audit_mismatches += mis;
    }
    return 0;
}

// The row loop of knn_rms.cpp:231-293 on the GPU: the fit rows go through in row blocks of ctx->chunk_rows (the
// reference's --block-size exists for the same reason: it bounds the working memory of a block of rows).
int rms_run(mdsctk_knn_ctx *ctx, FrameSet &fitset, long long fit_begin, long long n_fit, int k1, int do_fit,
            double *out_dist, int *out_idx)
{
    const FrameSetView ref = ctx->ref.view();
    if (n_fit <= 0) return fail(ctx, MDSCTK_KNN_EINVAL, "n_fit must be positive");
    if (k1 < 1 || k1 > ref.n) return fail(ctx, MDSCTK_KNN_EINVAL, "k1 must satisfy 1 <= k1 <= n_reference");
    if (k1 > 2040) return fail(ctx, MDSCTK_KNN_EINVAL, "k1 > 2040 is not supported");
    if ((size_t)ref.A * 24 + 2048 * 12 > 200 * 1024)
        return fail(ctx, MDSCTK_KNN_EINVAL, "n_atoms too large for the FP64 re-score kernel (max ~7500)");
    mdsctk_knn_stats &S = ctx->stats;
    S.ms_sweep = S.ms_rescore = S.ms_fallback = S.ms_download = 0;
    S.pairs = n_fit * ref.n; S.launches = 0; S.fallback_rows = 0; S.sweep_appends = 0; S.max_filter_err = 0;
    S.max_filter_spread = 0; S.rescored_max = 0; S.audit_rows = 0; S.audit_mismatches = 0;
    CK(ctx->scalars.reserve(64), "cudaMalloc(scalars)");
    if (ctx->gmax_dirty) {
        CK(launch_max_float(ref.G, ref.n, ctx->scalars.as<float>(), ctx->st), "max(G)");
        CK(cudaMemcpyAsync(&ctx->g_ref_max, ctx->scalars.p, 4, cudaMemcpyDeviceToHost, ctx->st), "D2H max(G)");
        CK(cudaStreamSynchronize(ctx->st), "sync max(G)");
    }
    // 64 * sqrt(G) bounds every fp16 operand element (overflow at 65504): the DEFAULT kernel gives way to 3xTF32
    // for such coordinates; an explicitly chosen fp16 kernel is refused instead
    RmsPlan P;
    P.rms_kernel = ctx->rms_kernel; P.k1 = k1; P.do_fit = do_fit;
    const bool fp16_overflow = 64.0 * std::sqrt((double)ctx->g_ref_max) > 3.0e4;
    if (P.rms_kernel >= MDSCTK_KNN_RMS_TC_3XFP16 && fp16_overflow) {
        if (ctx->rms_kernel_set)
            return fail(ctx, MDSCTK_KNN_EINVAL, "coordinates too large for the fp16 kernels; use rms_kernel=1 (3xTF32)");
        P.rms_kernel = MDSCTK_KNN_RMS_TC_3XTF32;
    }
    S.rms_kernel = P.rms_kernel; S.sweep_version = 0;
    // operand planes of this kernel: packed at load time for the kernel selected then, otherwise written now
    const int fam = families_of_kernel(P.rms_kernel);
    int rc = ensure_families(ctx, ctx->ref, fam);
    if (rc) return rc;
    if (&fitset != &ctx->ref && (rc = ensure_families(ctx, fitset, fam)) != 0) return rc;
    if (ctx->gmax_dirty) {
        for (int part = 0; part < 2; ++part) {
            CK(launch_max_float_strided(ref.gres + part, ref.n, 2, ctx->scalars.as<float>() + 1 + part, ctx->st), "max(gres)");
            CK(cudaMemcpyAsync(&ctx->gres_ref_max[part], ctx->scalars.as<float>() + 1 + part, 4, cudaMemcpyDeviceToHost, ctx->st),
               "D2H max(gres)");
        }
        CK(cudaStreamSynchronize(ctx->st), "sync max(gres)");
        ctx->gmax_dirty = false;
    }
    P.use_tc = P.rms_kernel != MDSCTK_KNN_RMS_SIMT_FP32;
    // Row blocks: at most chunk_rows rows, balanced, and -- for the persistent tensor-core sweep, whose work items are 256-row
    // super-tiles taken by the SM pairs in waves -- a whole number of waves each, so that only the last block of a query has a
    // partial wave (the segment count of every block is chosen for ITS item count: rms_tc_choose_segments).
    choose_lists(ctx, P.rms_kernel, k1, &P.keep, &P.cap);
    // tensor-core sweep: four private sub-lists per (row, segment), merged into one at the end of the segment
    if (P.use_tc) P.cap = rms_tc_list_stride(P.keep);
    // few, large blocks: a launch ends with the pairs finishing up to one work item apart (an item is ~25 ms at 1M frames with
    // four segments), so every extra launch costs a few per cent; the candidate lists (rows x segments x cap x 8 B) are what limits a block
    long long auto_rows = (long long)(6.0e9 / ((double)(P.use_tc ? 4 : 1) * P.cap * 8.0));
    auto_rows = std::max<long long>(16384, std::min<long long>(auto_rows, 1LL << 20));
    long long block = std::min<long long>(n_fit, ctx->chunk_rows > 0 ? std::max<long long>(256, ctx->chunk_rows) : auto_rows);
    if (n_fit > block) {
        const long long n_blocks = (n_fit + block - 1) / block;
        block = (n_fit + n_blocks - 1) / n_blocks;
        if (P.use_tc) {
            const long long wave = (long long)std::max(1, ctx->n_sms / 2) * 256;
            if (block >= wave) block = (block + wave - 1) / wave * wave;
        }
    }
    P.oos = &fitset != &ctx->ref;      // out-of-sample: the fit rows are not reference frames
    auto plan_block = [&](long long nb, RmsPlan *Pb) {
        *Pb = P;
        Pb->n_seg = P.use_tc ? (ctx->rms_segments > 0 ? ctx->rms_segments : rms_tc_choose_segments(nb, ref.n, ctx->n_sms)) : 1;
        Pb->H = P.use_tc ? rms_tc_lists_per_segment() * Pb->n_seg : 1;
    };
    size_t max_lists = 0, max_rows = 0;
    int max_H = 1;
    for (long long off = 0; off < n_fit; off += block) {
        const long long nb = std::min(block, n_fit - off);
        RmsPlan Pb;
        plan_block(nb, &Pb);
        max_lists = std::max(max_lists, (size_t)nb * Pb.H);
        max_rows = std::max(max_rows, (size_t)nb);
        max_H = std::max(max_H, Pb.H);
    }
    S.k_keep = P.keep;
    S.lists_per_row = max_H;
    if ((size_t)ref.A * 48 + (size_t)std::min(P.keep * max_H, 2048) * 2 * 28 + 64 * 80 > 220 * 1024 || P.keep > 2048)
        return fail(ctx, MDSCTK_KNN_EINVAL, "k too large for the FP64 re-score kernel's shared memory");
    CK(ctx->cand_key.reserve(max_lists * P.cap * 4), "cudaMalloc(cand_key)");
    CK(ctx->cand_idx.reserve(max_lists * P.cap * 4), "cudaMalloc(cand_idx)");
    CK(ctx->cand_cnt.reserve(max_lists * 4), "cudaMalloc(cand_cnt)");
    CK(ctx->cand_tau.reserve(max_lists * 8), "cudaMalloc(cand_tau)");
    CK(ctx->flags.reserve(max_rows * 4), "cudaMalloc(flags)");
    CK(ctx->bad_rows.reserve(max_rows * 4), "cudaMalloc(bad_rows)");
    CK(ctx->out_dist.reserve((size_t)n_fit * k1 * 8), "cudaMalloc(out_dist)");
    CK(ctx->out_idx.reserve((size_t)n_fit * k1 * 4), "cudaMalloc(out_idx)");
    if (P.use_tc) {
        if (ctx->debug_tile_on) CK(ctx->debug_tile.reserve(128 * 432 * 4), "cudaMalloc(debug_tile)");
        CK(ctx->row_tau.reserve(max_rows * 4), "cudaMalloc(row_tau)");
        if (P.oos) CK(ctx->own_tile.reserve((size_t)((max_rows + 255) / 256) * 4), "cudaMalloc(own_tile)");
    }
    ctx->out_rows = n_fit; ctx->out_k1 = k1;
    for (long long off = 0; off < n_fit; off += block) {
        const long long nb = std::min(block, n_fit - off);
        RmsPlan Pb;
        plan_block(nb, &Pb);
        if ((rc = rms_run_block(ctx, fitset, Pb, fit_begin + off, nb, off)) != 0) return rc;
    }
    if (out_dist || out_idx) {
        rc = mdsctk_knn_fetch(ctx, out_dist, out_idx);
        if (rc) return rc;
    }
    if (S.audit_mismatches > 0) {
        char b[200];
        snprintf(b, sizeof b, "audit: %lld of %lld certified rows differ from their exact FP64 recomputation (set force_exact=1 "
                 "and report this input)", S.audit_mismatches, S.audit_rows);
        return fail(ctx, MDSCTK_KNN_EAUDIT, b);
    }
    return 0;
}

// Euclidean knn_data through the tensor-core filter (data_tc.cu): pack -> sweep -> exact FP64 re-score
// with certificate -> exact FP64 sweep of the rows that could not be certified.
int data_run_tc(mdsctk_knn_ctx *ctx, const double *d_fit, bool fit_is_ref, long long fit_begin, long long n_fit, int k1,
                int metric, const double *fit_stats, const double *ref_stats)
{
    // correlation metric (knn_data -c): the same filter on the standardised rows (unit vectors: |z_x - z_y|^2 is four times
    // the squared correlation distance), exact re-score with correlation_distance's own arithmetic
    const bool corr = metric == MDSCTK_KNN_CORRELATION;
    if (ctx->dt_metric != metric) { ctx->dpack_dirty = true; ctx->dt_metric = metric; }
    mdsctk_knn_stats &S = ctx->stats;
    const int dim = ctx->ddim, D_pad = data_tc_pad_dim(dim);
    const long long n_ref = ctx->dn_ref;
    CK(ctx->scalars.reserve(64), "cudaMalloc(scalars)");
    ctx->tm.start(ctx->st);
    if (ctx->dpack_dirty) {
        if (corr) {
            int bad = 0;
            CK(launch_data_stats_check(ref_stats, n_ref, ctx->scalars.as<int>() + 12, ctx->st), "stats check");
            CK(cudaMemcpyAsync(&bad, ctx->scalars.as<int>() + 12, 4, cudaMemcpyDeviceToHost, ctx->st), "D2H stats check");
            CK(cudaStreamSynchronize(ctx->st), "sync stats check");
            if (bad) return 1;                       // constant / non-finite rows: the exact sweep handles them
            ctx->dt_ref_maxabs = 1.0;                // standardised rows are unit vectors
        } else {
            CK(launch_data_maxabs(ctx->d_ref.as<double>(), (size_t)n_ref * dim, ctx->scalars.as<double>(), ctx->st), "maxabs");
            CK(cudaMemcpyAsync(&ctx->dt_ref_maxabs, ctx->scalars.p, 8, cudaMemcpyDeviceToHost, ctx->st), "D2H maxabs");
            CK(cudaStreamSynchronize(ctx->st), "sync maxabs");
        }
        // power-of-two scale that puts the largest reference value near 1024 (fp16 keeps 11 bits below it)
        int e = 0;
        if (ctx->dt_ref_maxabs > 0.0 && std::isfinite(ctx->dt_ref_maxabs)) e = 10 - (int)std::ceil(std::log2(ctx->dt_ref_maxabs));
        e = std::max(-100, std::min(100, e));
        ctx->dt_scale = std::ldexp(1.0, e);
        CK(ctx->dt_ref_hi.reserve((size_t)n_ref * D_pad * 2), "cudaMalloc(ref hi)");
        CK(ctx->dt_ref_lo.reserve((size_t)n_ref * D_pad * 2), "cudaMalloc(ref lo)");
        CK(ctx->dt_ref_norm.reserve((size_t)(n_ref + 512) * 4), "cudaMalloc(ref norm)");
        CK(cudaMemsetAsync(ctx->dt_ref_norm.p, 0, (size_t)(n_ref + 512) * 4, ctx->st), "memset ref norm");
        CK(ctx->dt_ref_norm1.reserve((size_t)(n_ref + 512) * 4), "cudaMalloc(ref norm1)");
        CK(cudaMemsetAsync(ctx->dt_ref_norm1.p, 0, (size_t)(n_ref + 512) * 4, ctx->st), "memset ref norm1");
        CK(ctx->dt_ref_g.reserve((size_t)n_ref * 4), "cudaMalloc(ref g)");
        CK(launch_data_pack(ctx->d_ref.as<double>(), n_ref, dim, D_pad, ctx->dt_scale, corr ? ref_stats : nullptr, ctx->dt_ref_hi.p, ctx->dt_ref_lo.p,
                            ctx->dt_ref_norm.as<float>(), ctx->dt_ref_norm1.as<float>(), ctx->dt_ref_g.as<float>(), ctx->st),
           "data_pack(ref)");
        CK(launch_max_float(ctx->dt_ref_norm.as<float>(), n_ref, ctx->scalars.as<float>() + 4, ctx->st), "max(norm)");
        CK(cudaMemcpyAsync(&ctx->dt_rnorm_max, ctx->scalars.as<float>() + 4, 4, cudaMemcpyDeviceToHost, ctx->st), "D2H max(norm)");
        CK(launch_max_float(ctx->dt_ref_g.as<float>(), n_ref, ctx->scalars.as<float>() + 5, ctx->st), "max(g)");
        CK(cudaMemcpyAsync(&ctx->dt_g_ref_max, ctx->scalars.as<float>() + 5, 4, cudaMemcpyDeviceToHost, ctx->st), "D2H max(g)");
        CK(cudaStreamSynchronize(ctx->st), "sync max(norm)");
        ctx->dpack_dirty = false;
        S.launches += 3;
    }
    const bool one = ctx->data_kernel == 2 || ctx->data_kernel < 0;      // one-part (1xFP16) filter: what auto selects
    const void *fit_hi, *fit_lo;
    const float *fit_norm, *fit_norm1, *fit_g;
    if (fit_is_ref) {
        fit_hi = ctx->dt_ref_hi.as<uint16_t>() + (size_t)fit_begin * D_pad;
        fit_lo = ctx->dt_ref_lo.as<uint16_t>() + (size_t)fit_begin * D_pad;
        fit_norm = ctx->dt_ref_norm.as<float>() + fit_begin;
        fit_norm1 = ctx->dt_ref_norm1.as<float>() + fit_begin;
        fit_g = ctx->dt_ref_g.as<float>() + fit_begin;
    } else {
        double fit_max = 0.0;
        if (corr) {
            int bad = 0;
            CK(launch_data_stats_check(fit_stats, n_fit, ctx->scalars.as<int>() + 12, ctx->st), "stats check(fit)");
            CK(cudaMemcpyAsync(&bad, ctx->scalars.as<int>() + 12, 4, cudaMemcpyDeviceToHost, ctx->st), "D2H stats check");
            CK(cudaStreamSynchronize(ctx->st), "sync stats check");
            if (bad) return 1;
            fit_max = 1.0;
        } else {
            CK(launch_data_maxabs(d_fit, (size_t)n_fit * dim, ctx->scalars.as<double>(), ctx->st), "maxabs(fit)");
            CK(cudaMemcpyAsync(&fit_max, ctx->scalars.p, 8, cudaMemcpyDeviceToHost, ctx->st), "D2H maxabs");
            CK(cudaStreamSynchronize(ctx->st), "sync maxabs");
        }
        if (!(fit_max * ctx->dt_scale < 3.0e4)) return 1;   // would overflow fp16: the caller takes the exact path
        CK(ctx->dt_fit_hi.reserve((size_t)n_fit * D_pad * 2), "cudaMalloc(fit hi)");
        CK(ctx->dt_fit_lo.reserve((size_t)n_fit * D_pad * 2), "cudaMalloc(fit lo)");
        CK(ctx->dt_fit_norm.reserve((size_t)(n_fit + 512) * 4), "cudaMalloc(fit norm)");
        CK(ctx->dt_fit_norm1.reserve((size_t)(n_fit + 512) * 4), "cudaMalloc(fit norm1)");
        CK(ctx->dt_fit_g.reserve((size_t)n_fit * 4), "cudaMalloc(fit g)");
        CK(launch_data_pack(d_fit, n_fit, dim, D_pad, ctx->dt_scale, corr ? fit_stats : nullptr, ctx->dt_fit_hi.p, ctx->dt_fit_lo.p,
                            ctx->dt_fit_norm.as<float>(), ctx->dt_fit_norm1.as<float>(), ctx->dt_fit_g.as<float>(), ctx->st),
           "data_pack(fit)");
        fit_hi = ctx->dt_fit_hi.p; fit_lo = ctx->dt_fit_lo.p; fit_norm = ctx->dt_fit_norm.as<float>();
        fit_norm1 = ctx->dt_fit_norm1.as<float>(); fit_g = ctx->dt_fit_g.as<float>();
        S.launches += 2;
    }
    S.ms_pack += ctx->tm.stop(ctx->st);

    const long long slack = ctx->slack >= 0 ? ctx->slack : std::max<long long>(32, k1 / 2);
    const int keep = (int)(((long long)k1 + slack + 7) / 8 * 8);
    const int cap = data_tc_list_stride(keep);
    const int n_seg = ctx->data_segments > 0 ? ctx->data_segments : data_tc_choose_segments(n_fit, n_ref, ctx->n_sms);
    S.k_keep = keep; S.lists_per_row = n_seg;
    if ((size_t)dim * 8 + (size_t)keep * n_seg * 2 * 28 > 170 * 1024) return 1;   // re-score working set: exact path instead
    CandLists<float> cl;
    CK(ctx->cand_key.reserve((size_t)n_fit * n_seg * cap * 4), "cudaMalloc(cand_key)");
    CK(ctx->cand_idx.reserve((size_t)n_fit * n_seg * cap * 4), "cudaMalloc(cand_idx)");
    CK(ctx->cand_cnt.reserve((size_t)n_fit * n_seg * 4), "cudaMalloc(cand_cnt)");
    CK(ctx->cand_tau.reserve((size_t)n_fit * n_seg * 8), "cudaMalloc(cand_tau)");
    CK(ctx->flags.reserve((size_t)n_fit * 4), "cudaMalloc(flags)");
    CK(ctx->bad_rows.reserve((size_t)n_fit * 4), "cudaMalloc(bad_rows)");
    CK(ctx->row_tau.reserve((size_t)n_fit * 4), "cudaMalloc(row_tau)");
    CK(ctx->out_dist.reserve((size_t)n_fit * k1 * 8), "cudaMalloc(out_dist)");
    CK(ctx->out_idx.reserve((size_t)n_fit * k1 * 4), "cudaMalloc(out_idx)");
    cl.key = ctx->cand_key.as<float>(); cl.idx = ctx->cand_idx.as<int>(); cl.cnt = ctx->cand_cnt.as<int>();
    cl.tau = ctx->cand_tau.as<float>(); cl.cap = cap; cl.keep = keep; cl.H = n_seg;
    CK(cudaMemsetAsync(ctx->scalars.p, 0, 64, ctx->st), "memset scalars");
    double *d_err = ctx->scalars.as<double>() + 1;
    int *d_nbad = ctx->scalars.as<int>() + 8;

    ctx->tm.start(ctx->st);
    CK(launch_fill_u32(ctx->row_tau.p, (size_t)n_fit, 0x7f800000u, ctx->st), "fill row_tau");
    CK(launch_data_sweep_tc(fit_hi, fit_lo, one ? fit_norm1 : fit_norm, n_fit, fit_is_ref ? fit_begin : -1, ctx->dt_ref_hi.p,
                            ctx->dt_ref_lo.p, one ? ctx->dt_ref_norm1.as<float>() : ctx->dt_ref_norm.as<float>(), n_ref, D_pad,
                            ctx->dt_scale, n_seg, cl, ctx->row_tau.as<float>(), one ? 1 : 0, ctx->n_sms, ctx->st), "data_sweep_tc");
    S.ms_sweep = ctx->tm.stop(ctx->st);
    CK(cudaGetLastError(), "data tensor sweep kernel");
    S.launches += 2;

    // filter noise: fp32 accumulation of 3 * dim products and the three-term key, relative to |x|^2 + |y|^2
    // (measured half-spread at dim 512: 2.1e-6; the certificate re-checks the observed spread on every row)
    const double eps_rel = 2.0e-6 * std::sqrt((double)std::max(dim, 64) / 64.0) * (double)ctx->cert_scale_ppm * 1e-6;
    S.cert_eps = eps_rel * 2.0 * (double)ctx->dt_rnorm_max / (ctx->dt_scale * ctx->dt_scale);
    ctx->tm.start(ctx->st);
    S.cert_gres = one ? (double)ctx->dt_g_ref_max : 0.0;
    CK(launch_data_rescore(d_fit, ctx->d_ref.as<double>(), n_fit, dim, k1, cl, eps_rel, fit_norm, ctx->dt_scale,
                           ctx->dt_rnorm_max, one ? fit_g : nullptr, ctx->dt_g_ref_max, corr ? fit_stats : nullptr, corr ? ref_stats : nullptr,
                           (ctx->out_dist.as<double>() + (size_t)ctx->out_off * k1), (ctx->out_idx.as<int>() + (size_t)ctx->out_off * k1), ctx->flags.as<int>(), d_err,
                           d_nbad, ctx->bad_rows.as<int>(), ctx->st), "data_rescore");
    struct { double pad, err, spread, done_max; int nbad; } host_sc;
    CK(cudaMemcpyAsync(&host_sc, ctx->scalars.p, sizeof(host_sc), cudaMemcpyDeviceToHost, ctx->st), "D2H scalars");
    S.ms_rescore = ctx->tm.stop(ctx->st);
    CK(cudaGetLastError(), "data rescore kernel");
    S.launches += 1;
    S.max_filter_err = host_sc.err; S.max_filter_spread = host_sc.spread; S.rescored_max = (int)host_sc.done_max;
    S.fallback_rows = host_sc.nbad;

    if (host_sc.nbad > 0) {     // exact FP64 sweep of the uncertified rows (gathered), results scattered back
        ctx->tm.start(ctx->st);
        const int nb = host_sc.nbad;
        const long long fslack = 8;
        const int fkeep = (int)(((long long)k1 + fslack + 7) / 8 * 8);
        const int fcap = (fkeep + 160 + 31) / 32 * 32;
        CK(ctx->fb_rows.reserve((size_t)nb * dim * 8), "cudaMalloc(fb_rows)");
        CK(ctx->fb_key.reserve((size_t)nb * fcap * 8), "cudaMalloc(fb_key)");
        CK(ctx->fb_idx.reserve((size_t)nb * fcap * 4), "cudaMalloc(fb_idx)");
        CK(ctx->fb_cnt.reserve((size_t)nb * 4), "cudaMalloc(fb_cnt)");
        CK(ctx->fb_tau.reserve((size_t)nb * 8), "cudaMalloc(fb_tau)");
        CK(ctx->fb_dist.reserve((size_t)nb * k1 * 8), "cudaMalloc(fb_dist)");
        CK(ctx->fb_oidx.reserve((size_t)nb * k1 * 4), "cudaMalloc(fb_oidx)");
        CK(launch_data_gather_rows(d_fit, ctx->bad_rows.as<int>(), nb, dim, ctx->fb_rows.as<double>(), ctx->st), "gather rows");
        CandLists<double> fl;
        fl.key = ctx->fb_key.as<double>(); fl.idx = ctx->fb_idx.as<int>(); fl.cnt = ctx->fb_cnt.as<int>();
        fl.tau = ctx->fb_tau.as<double>(); fl.cap = fcap; fl.keep = fkeep; fl.H = 1;
        const double *fb_stats = nullptr;
        if (corr) {
            CK(ctx->fb_stats.reserve((size_t)nb * 16), "cudaMalloc(fb_stats)");
            CK(launch_data_rowstats(ctx->fb_rows.as<double>(), nb, dim, ctx->fb_stats.as<double>(), ctx->st), "data_rowstats(fallback)");
            fb_stats = ctx->fb_stats.as<double>();
        }
        CK(launch_data_sweep(ctx->fb_rows.as<double>(), fb_stats, nb, ctx->d_ref.as<double>(), corr ? ref_stats : nullptr, n_ref, dim,
                             metric, fl, ctx->st), "data_sweep(fallback)");
        CK(launch_data_finalize(fl, nb, k1, ctx->fb_dist.as<double>(), ctx->fb_oidx.as<int>(), ctx->st), "data_finalize(fallback)");
        CK(launch_data_scatter_out(ctx->fb_dist.as<double>(), ctx->fb_oidx.as<int>(), ctx->bad_rows.as<int>(), nb, k1,
                                   (ctx->out_dist.as<double>() + (size_t)ctx->out_off * k1), (ctx->out_idx.as<int>() + (size_t)ctx->out_off * k1), ctx->st), "scatter rows");
        S.ms_fallback = ctx->tm.stop(ctx->st);
        CK(cudaGetLastError(), "data fallback kernels");
        S.launches += 4;
    }
    return 0;
}

int data_run_block(mdsctk_knn_ctx *ctx, const double *d_fit, bool fit_is_ref, long long fit_begin, long long n_fit, int k1,
                   int metric)
{
    mdsctk_knn_stats &S = ctx->stats;
    S.ms_sweep = S.ms_rescore = S.ms_fallback = S.ms_download = 0;
    S.pairs = n_fit * ctx->dn_ref; S.launches = 0; S.fallback_rows = 0; S.max_filter_err = 0; S.cert_eps = 0;
    const int dim = ctx->ddim;
    S.max_filter_spread = 0; S.rescored_max = 0; S.lists_per_row = 1;
    const double *fit_stats = nullptr, *ref_stats = nullptr;
    if (metric == MDSCTK_KNN_CORRELATION) {
        if (ctx->dstats_dirty) {
            CK(ctx->d_ref_stats.reserve((size_t)ctx->dn_ref * 16), "cudaMalloc(ref_stats)");
            CK(launch_data_rowstats(ctx->d_ref.as<double>(), ctx->dn_ref, dim, ctx->d_ref_stats.as<double>(), ctx->st),
               "data_rowstats");
            ctx->dstats_dirty = false;
            S.launches++;
        }
        ref_stats = ctx->d_ref_stats.as<double>();
        if (fit_is_ref) {
            fit_stats = ref_stats + 2 * fit_begin;
        } else {
            CK(ctx->d_fit_stats.reserve((size_t)n_fit * 16), "cudaMalloc(fit_stats)");
            CK(launch_data_rowstats(d_fit, n_fit, dim, ctx->d_fit_stats.as<double>(), ctx->st), "data_rowstats");
            fit_stats = ctx->d_fit_stats.as<double>();
            S.launches++;
        }
    }
    // tensor-core filter (both metrics: correlation = Euclidean on the standardised rows); by default only where the
    // contraction is worth staging
    const bool want_tc = ctx->data_kernel >= 1 || (ctx->data_kernel < 0 && dim >= 8 && (double)n_fit * (double)ctx->dn_ref >= 2.5e7);
    if (want_tc) {
        const int rc = data_run_tc(ctx, d_fit, fit_is_ref, fit_begin, n_fit, k1, metric, fit_stats, ref_stats);
        if (rc != 1) return rc;        // 1: not applicable to this input -> exact sweep below
    }
    const long long slack = ctx->slack >= 0 ? ctx->slack : 8;
    const int keep = (int)(((long long)k1 + slack + 7) / 8 * 8);
    const int cap = (keep + 160 + 31) / 32 * 32;
    S.k_keep = keep;
    CandLists<double> cl;
    CK(ctx->cand_key.reserve((size_t)n_fit * cap * 8), "cudaMalloc(cand_key)");
    CK(ctx->cand_idx.reserve((size_t)n_fit * cap * 4), "cudaMalloc(cand_idx)");
    CK(ctx->cand_cnt.reserve((size_t)n_fit * 4), "cudaMalloc(cand_cnt)");
    CK(ctx->cand_tau.reserve((size_t)n_fit * 8), "cudaMalloc(cand_tau)");
    CK(ctx->out_dist.reserve((size_t)n_fit * k1 * 8), "cudaMalloc(out_dist)");
    CK(ctx->out_idx.reserve((size_t)n_fit * k1 * 4), "cudaMalloc(out_idx)");
    cl.key = ctx->cand_key.as<double>(); cl.idx = ctx->cand_idx.as<int>(); cl.cnt = ctx->cand_cnt.as<int>();
    cl.tau = ctx->cand_tau.as<double>(); cl.cap = cap; cl.keep = keep; cl.H = 1;
    ctx->tm.start(ctx->st);
    CK(launch_data_sweep(d_fit, fit_stats, n_fit, ctx->d_ref.as<double>(), ref_stats, ctx->dn_ref, dim, metric, cl,
                         ctx->st), "data_sweep");
    S.ms_sweep = ctx->tm.stop(ctx->st);
    CK(cudaGetLastError(), "data sweep kernel");
    ctx->tm.start(ctx->st);
    CK(launch_data_finalize(cl, n_fit, k1, (ctx->out_dist.as<double>() + (size_t)ctx->out_off * k1), (ctx->out_idx.as<int>() + (size_t)ctx->out_off * k1), ctx->st),
       "data_finalize");
    S.ms_rescore = ctx->tm.stop(ctx->st);
    CK(cudaGetLastError(), "data finalize kernel");
    S.launches += 2;
    return 0;
}

// The row loop of knn_data.cpp:195-250: row blocks of ctx->chunk_rows fit rows, results in row order.
int data_run(mdsctk_knn_ctx *ctx, const double *d_fit, bool fit_is_ref, long long fit_begin, long long n_fit, int k1,
             int metric, double *out_dist, int *out_idx)
{
    if (n_fit <= 0) return fail(ctx, MDSCTK_KNN_EINVAL, "n_fit must be positive");
    if (k1 < 1 || k1 > ctx->dn_ref) return fail(ctx, MDSCTK_KNN_EINVAL, "k1 must satisfy 1 <= k1 <= n_reference");
    if (k1 > 2040) return fail(ctx, MDSCTK_KNN_EINVAL, "k1 > 2040 is not supported");
    if (metric != MDSCTK_KNN_EUCLIDEAN && metric != MDSCTK_KNN_CORRELATION)
        return fail(ctx, MDSCTK_KNN_EINVAL, "unknown metric");
    CK(ctx->out_dist.reserve((size_t)n_fit * k1 * 8), "cudaMalloc(out_dist)");
    CK(ctx->out_idx.reserve((size_t)n_fit * k1 * 4), "cudaMalloc(out_idx)");
    ctx->out_rows = n_fit; ctx->out_k1 = k1;
    const long long block = std::min<long long>(n_fit, ctx->chunk_rows > 0 ? std::max<long long>(256, ctx->chunk_rows) : 131072);
    mdsctk_knn_stats tot = ctx->stats;
    tot.ms_sweep = tot.ms_rescore = tot.ms_fallback = tot.ms_download = 0;
    tot.launches = 0; tot.fallback_rows = 0; tot.max_filter_err = 0; tot.max_filter_spread = 0; tot.rescored_max = 0;
    const double pack0 = ctx->stats.ms_pack;
    for (long long off = 0; off < n_fit; off += block) {
        const long long nb = std::min(block, n_fit - off);
        ctx->out_off = off;
        const int rc = data_run_block(ctx, d_fit + (size_t)off * ctx->ddim, fit_is_ref, fit_begin + off, nb, k1, metric);
        ctx->out_off = 0;
        if (rc) return rc;
        const mdsctk_knn_stats &B = ctx->stats;
        tot.ms_sweep += B.ms_sweep; tot.ms_rescore += B.ms_rescore; tot.ms_fallback += B.ms_fallback;
        tot.launches += B.launches; tot.fallback_rows += B.fallback_rows;
        tot.max_filter_err = std::max(tot.max_filter_err, B.max_filter_err);
        tot.max_filter_spread = std::max(tot.max_filter_spread, B.max_filter_spread);
        tot.rescored_max = std::max(tot.rescored_max, B.rescored_max);
        tot.cert_eps = B.cert_eps; tot.cert_gres = B.cert_gres; tot.k_keep = B.k_keep; tot.lists_per_row = B.lists_per_row;
        tot.ms_pack = B.ms_pack;
    }
    (void)pack0;
    tot.pairs = n_fit * ctx->dn_ref;
    ctx->stats = tot;
    if (out_dist || out_idx) return mdsctk_knn_fetch(ctx, out_dist, out_idx);
    return 0;
}

}  // namespace

extern "C" {

int mdsctk_knn_abi_version(void) { return MDSCTK_KNN_ABI_VERSION; }

int mdsctk_knn_create(mdsctk_knn_ctx **out, int device_id)
{
    if (!out) { g_create_error = "out is NULL"; return MDSCTK_KNN_EINVAL; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (this library has no CPU fallback)";
        return MDSCTK_KNN_ENODEV;
    }
    if (device_id < 0 || device_id >= n) { g_create_error = "device_id out of range"; return MDSCTK_KNN_ENODEV; }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device_id)) != cudaSuccess) {
        g_create_error = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e);
        return MDSCTK_KNN_ECUDA;
    }
    if (prop.major != 10) {
        char b[160];
        snprintf(b, sizeof b, "device %d (%s, sm_%d%d) is not a Blackwell sm_100 GPU; kernels are built for sm_100a only",
                 device_id, prop.name, prop.major, prop.minor);
        g_create_error = b;
        return MDSCTK_KNN_ENODEV;
    }
    mdsctk_knn_ctx *c = new (std::nothrow) mdsctk_knn_ctx();
    if (!c) { g_create_error = "out of host memory"; return MDSCTK_KNN_ENOMEM; }
    c->dev = device_id;
    c->n_sms = prop.multiProcessorCount;
    memset(&c->stats, 0, sizeof c->stats);
    cudaSetDevice(device_id);
    if ((e = cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking)) != cudaSuccess) {
        g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
        delete c;
        return MDSCTK_KNN_ECUDA;
    }
    c->tm.init();
    c->user_tm.init();
    *out = c;
    return 0;
}

void mdsctk_knn_destroy(mdsctk_knn_ctx *ctx)
{
    if (!ctx) return;
    Bind b(ctx);
    cudaStreamSynchronize(ctx->st);
    ctx->ref.release(); ctx->fit.release(); ctx->wnorm.release();
    ctx->d_ref.release(); ctx->d_fit.release(); ctx->d_ref_stats.release(); ctx->d_fit_stats.release();
    ctx->cand_key.release(); ctx->cand_idx.release(); ctx->cand_cnt.release(); ctx->cand_tau.release();
    ctx->flags.release(); ctx->bad_rows.release(); ctx->scalars.release(); ctx->rows_buf.release();
    ctx->out_dist.release(); ctx->out_idx.release(); ctx->debug_tile.release(); ctx->row_tau.release(); ctx->own_tile.release();
    for (DevBuf *b2 : {&ctx->dt_ref_hi, &ctx->dt_ref_lo, &ctx->dt_ref_norm, &ctx->dt_fit_hi, &ctx->dt_fit_lo, &ctx->dt_fit_norm,
                       &ctx->dt_ref_norm1, &ctx->dt_ref_g, &ctx->dt_fit_norm1, &ctx->dt_fit_g, &ctx->fb_rows, &ctx->fb_key,
                       &ctx->fb_idx, &ctx->fb_cnt, &ctx->fb_tau, &ctx->fb_dist, &ctx->fb_oidx, &ctx->fb_stats, &ctx->c_idx, &ctx->c_dist,
                       &ctx->c_ints, &ctx->c_key, &ctx->c_val, &ctx->c_irow, &ctx->c_oval, &ctx->f_in, &ctx->f_ang, &ctx->f_sc,
                       &ctx->audit_ids, &ctx->audit_seq, &ctx->audit_dist, &ctx->audit_idx, &ctx->s_int, &ctx->s_val, &ctx->s_vec, &ctx->s_basis, &ctx->s_rot, &ctx->s_small, &ctx->s_evec})
        b2->release();
    ctx->tm.destroy();
    ctx->user_tm.destroy();
    cudaStreamDestroy(ctx->st);
    delete ctx;
}

const char *mdsctk_knn_last_error(const mdsctk_knn_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int mdsctk_knn_set_option(mdsctk_knn_ctx *ctx, const char *key, long long value)
{
    if (!ctx || !key) return MDSCTK_KNN_EINVAL;
    if (!strcmp(key, "rms_kernel")) {
        if (value < 0 || value > 6) return fail(ctx, MDSCTK_KNN_EINVAL, "rms_kernel must be 0..6");
        ctx->rms_kernel = (int)value;
        ctx->rms_kernel_set = true;
    } else if (!strcmp(key, "slack")) {
        if (value < -1 || value > 1024) return fail(ctx, MDSCTK_KNN_EINVAL, "slack out of range");
        ctx->slack = value;
    } else if (!strcmp(key, "cert_scale_ppm")) {
        if (value < 0) return fail(ctx, MDSCTK_KNN_EINVAL, "cert_scale_ppm must be >= 0");
        ctx->cert_scale_ppm = value;
    } else if (!strcmp(key, "data_kernel")) {
        if (value < -1 || value > 2)
            return fail(ctx, MDSCTK_KNN_EINVAL, "data_kernel must be -1 (auto), 0 (exact), 1 (tensor, 3xFP16) or 2 (tensor, 1xFP16)");
        ctx->data_kernel = (int)value;
    } else if (!strcmp(key, "chunk_rows")) {
        if (value != 0 && value < 256) return fail(ctx, MDSCTK_KNN_EINVAL, "chunk_rows must be 0 (auto) or >= 256");
        ctx->chunk_rows = value;
    } else if (!strcmp(key, "audit_rows")) {
        if (value < 0 || value > 65536) return fail(ctx, MDSCTK_KNN_EINVAL, "audit_rows out of range");
        ctx->audit_rows = value;
    } else if (!strcmp(key, "sweep_version")) {
        if (value != 1 && value != 2) return fail(ctx, MDSCTK_KNN_EINVAL, "sweep_version must be 1 or 2");
        ctx->sweep_version = (int)value;
    } else if (!strcmp(key, "rms_segments")) {
        if (value < 0 || value > 32) return fail(ctx, MDSCTK_KNN_EINVAL, "rms_segments must be 0 (auto) .. 32");
        ctx->rms_segments = (int)value;
    } else if (!strcmp(key, "data_streaming")) {
        data_tc_set_streaming(value != 0);
    } else if (!strcmp(key, "data_segments")) {
        if (value < 0 || value > 8) return fail(ctx, MDSCTK_KNN_EINVAL, "data_segments must be 0 (auto) .. 8");
        ctx->data_segments = (int)value;
    } else if (!strcmp(key, "rms_wide_stages")) {
        ctx->rms_wide_stages = value != 0;
    } else if (!strcmp(key, "force_exact")) {
        ctx->force_exact = value != 0;
    } else if (!strcmp(key, "debug_tile")) {
        ctx->debug_tile_on = value != 0;
    } else {
        return fail(ctx, MDSCTK_KNN_EINVAL, std::string("unknown option: ") + key);
    }
    return 0;
}

int mdsctk_knn_get_stats(const mdsctk_knn_ctx *ctx, mdsctk_knn_stats *out)
{
    if (!ctx || !out) return MDSCTK_KNN_EINVAL;
    *out = ctx->stats;
    return 0;
}

/* ------------------------------------------------------------------ RMSD path ---- */
int mdsctk_knn_rms_alloc_reference(mdsctk_knn_ctx *ctx, long long n_total, int n_atoms, const float *mass)
{
    if (!ctx) return MDSCTK_KNN_EINVAL;
    if (n_total <= 0 || n_atoms <= 0 || !mass) return fail(ctx, MDSCTK_KNN_EINVAL, "bad reference shape or NULL mass");
    if (n_total > 0x7fffffffLL) return fail(ctx, MDSCTK_KNN_EINVAL, "more than 2^31-1 reference frames");
    Bind b(ctx);
    ctx->have_ref = false;
    int rc = upload_weights(ctx, mass, n_atoms);
    if (rc) return rc;
    CK(ctx->ref.alloc(n_total, n_atoms), "cudaMalloc(reference set)");
    ctx->ref.have = 0;                                    // new frames are coming: nothing is packed yet
    ctx->ref_pack_fam = families_of_kernel(ctx->rms_kernel);
    CK(ctx->ref.reserve_families(ctx->ref_pack_fam), "cudaMalloc(reference operand planes)");
    ctx->stats.ms_upload = ctx->stats.ms_pack = 0;
    ctx->have_ref = true;
    ctx->gmax_dirty = true;
    return 0;
}

int mdsctk_knn_rms_pack_shard(mdsctk_knn_ctx *ctx, const float *xyz, long long frame_offset, long long n_frames)
{
    if (!ctx) return MDSCTK_KNN_EINVAL;
    if (!ctx->have_ref) return fail(ctx, MDSCTK_KNN_ESTATE, "call mdsctk_knn_rms_alloc_reference first");
    if (!xyz || frame_offset < 0 || n_frames < 0 || frame_offset + n_frames > ctx->ref.n)
        return fail(ctx, MDSCTK_KNN_EINVAL, "shard outside the allocated reference set");
    Bind b(ctx);
    ctx->gmax_dirty = true;
    if (n_frames == 0) return 0;
    return pack_into(ctx, ctx->ref, xyz, frame_offset, n_frames, ctx->ref_pack_fam);
}

int mdsctk_knn_rms_reference_arrays(mdsctk_knn_ctx *ctx, int max_arrays, int *n_arrays, void **dev_ptrs,
                                    size_t *bytes_per_frame)
{
    if (!ctx || !n_arrays) return MDSCTK_KNN_EINVAL;
    if (!ctx->have_ref) return fail(ctx, MDSCTK_KNN_ESTATE, "no reference set");
    // always: raw, G, cen, sig, Gh, G2, gres; then the operand planes of the families this set is packed with
    struct { void *p; size_t b; } arr[16];
    int na = 0;
    const FrameSet &R = ctx->ref;
    const size_t p4 = (size_t)R.A_pad * 12, p2 = (size_t)R.A_pad * 6;
    arr[na++] = {R.raw.p, (size_t)R.A * 12};
    arr[na++] = {R.G.p, 4}; arr[na++] = {R.cen.p, 32}; arr[na++] = {R.sig.p, 16};
    arr[na++] = {R.Gh.p, 4}; arr[na++] = {R.G2.p, 4}; arr[na++] = {R.gres.p, 8};
    const int fam = ctx->ref_pack_fam | ((ctx->ref_pack_fam & F_FP16LO) ? F_FP16 : 0);
    if (fam & F_PLANES) arr[na++] = {R.planes.p, p4};
    if (fam & F_TF32) { arr[na++] = {R.hi.p, p4}; arr[na++] = {R.lo.p, p4}; }
    if (fam & F_BF16) { arr[na++] = {R.bh.p, p2}; arr[na++] = {R.bm.p, p2}; }
    if (fam & F_FP16) arr[na++] = {R.fh.p, p2};
    if (fam & F_FP16LO) arr[na++] = {R.fl.p, p2};
    *n_arrays = na;
    if (max_arrays < na || !dev_ptrs || !bytes_per_frame) return fail(ctx, MDSCTK_KNN_EINVAL, "need room for 16 arrays");
    for (int i = 0; i < na; ++i) { dev_ptrs[i] = arr[i].p; bytes_per_frame[i] = arr[i].b; }
    ctx->ref.have = fam;     // the caller is about to fill the other ranks' frames of exactly these arrays
    ctx->gmax_dirty = true;  // the caller is about to overwrite them (all-gather)
    return 0;
}

int mdsctk_knn_rms_set_reference(mdsctk_knn_ctx *ctx, const float *xyz, long long n_frames, int n_atoms,
                                 const float *mass)
{
    int rc = mdsctk_knn_rms_alloc_reference(ctx, n_frames, n_atoms, mass);
    if (rc) return rc;
    return mdsctk_knn_rms_pack_shard(ctx, xyz, 0, n_frames);
}

int mdsctk_knn_rms_query_range(mdsctk_knn_ctx *ctx, long long fit_begin, long long n_fit, int k1, int do_fit,
                               double *out_dist, int *out_idx)
{
    if (!ctx) return MDSCTK_KNN_EINVAL;
    if (!ctx->have_ref) return fail(ctx, MDSCTK_KNN_ESTATE, "no reference set");
    if (fit_begin < 0 || n_fit <= 0 || fit_begin + n_fit > ctx->ref.n)
        return fail(ctx, MDSCTK_KNN_EINVAL, "fit range outside the reference set");
    Bind b(ctx);
    return rms_run(ctx, ctx->ref, fit_begin, n_fit, k1, do_fit, out_dist, out_idx);
}

int mdsctk_knn_rms_query(mdsctk_knn_ctx *ctx, const float *fit_xyz, long long n_fit, int k1, int do_fit,
                         double *out_dist, int *out_idx)
{
    if (!ctx) return MDSCTK_KNN_EINVAL;
    if (!ctx->have_ref) return fail(ctx, MDSCTK_KNN_ESTATE, "no reference set");
    if (!fit_xyz) {
        if (n_fit != ctx->ref.n) return fail(ctx, MDSCTK_KNN_EINVAL, "fit_xyz == NULL requires n_fit == n_reference");
        return mdsctk_knn_rms_query_range(ctx, 0, n_fit, k1, do_fit, out_dist, out_idx);
    }
    if (n_fit <= 0) return fail(ctx, MDSCTK_KNN_EINVAL, "n_fit must be positive");
    Bind b(ctx);
    CK(ctx->fit.alloc(n_fit, ctx->ref.A), "cudaMalloc(fit set)");
    ctx->fit.have = 0;
    int rc = pack_into(ctx, ctx->fit, fit_xyz, 0, n_fit, families_of_kernel(ctx->rms_kernel));
    if (rc) return rc;
    return rms_run(ctx, ctx->fit, 0, n_fit, k1, do_fit, out_dist, out_idx);
}

int mdsctk_knn_rms_rows(mdsctk_knn_ctx *ctx, long long fit_begin, long long n_fit, int do_fit, double *out)
{
    if (!ctx || !out) return MDSCTK_KNN_EINVAL;
    if (!ctx->have_ref) return fail(ctx, MDSCTK_KNN_ESTATE, "no reference set");
    if (fit_begin < 0 || n_fit <= 0 || fit_begin + n_fit > ctx->ref.n)
        return fail(ctx, MDSCTK_KNN_EINVAL, "fit range outside the reference set");
    Bind b(ctx);
    const FrameSetView ref = ctx->ref.view();
    const size_t row_bytes = (size_t)ref.n * 8;
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_fit, ((size_t)256 << 20) / row_bytes));
    CK(ctx->rows_buf.reserve((size_t)chunk * row_bytes), "cudaMalloc(rows_buf)");
    for (long long off = 0; off < n_fit; off += chunk) {
        const int nr = (int)std::min<long long>(chunk, n_fit - off);
        CK(launch_rms_exact_rows(ref, nullptr, fit_begin + off, nr, ref, ctx->wnorm.as<double>(), do_fit,
                                 ctx->rows_buf.as<double>(), ctx->st), "rms_exact_rows");
        sqrt_scale_kernel<<<592, 256, 0, ctx->st>>>(ctx->rows_buf.as<double>(), (size_t)nr * ref.n, 10.0);
        CK(cudaGetLastError(), "sqrt_scale");
        CK(cudaMemcpyAsync(out + (size_t)off * ref.n, ctx->rows_buf.p, (size_t)nr * row_bytes, cudaMemcpyDeviceToHost,
                           ctx->st), "D2H rows");
        CK(cudaStreamSynchronize(ctx->st), "sync rows");
    }
    return 0;
}

/* ---------------------------------------------------------------- vector path ---- */
int mdsctk_knn_data_alloc_reference(mdsctk_knn_ctx *ctx, long long n_total, int dim)
{
    if (!ctx) return MDSCTK_KNN_EINVAL;
    if (n_total <= 0 || dim <= 0) return fail(ctx, MDSCTK_KNN_EINVAL, "bad reference shape");
    if (n_total > 0x7fffffffLL) return fail(ctx, MDSCTK_KNN_EINVAL, "more than 2^31-1 reference rows");
    Bind b(ctx);
    ctx->have_dref = false;
    CK(ctx->d_ref.reserve((size_t)n_total * dim * 8), "cudaMalloc(reference rows)");
    ctx->dn_ref = n_total; ctx->ddim = dim; ctx->have_dref = true; ctx->dstats_dirty = true; ctx->dpack_dirty = true;
    ctx->stats.ms_upload = ctx->stats.ms_pack = 0;
    return 0;
}

int mdsctk_knn_data_upload_shard(mdsctk_knn_ctx *ctx, const double *rows, long long row_offset, long long n_rows)
{
    if (!ctx) return MDSCTK_KNN_EINVAL;
    if (!ctx->have_dref) return fail(ctx, MDSCTK_KNN_ESTATE, "call mdsctk_knn_data_alloc_reference first");
    if (!rows || row_offset < 0 || n_rows < 0 || row_offset + n_rows > ctx->dn_ref)
        return fail(ctx, MDSCTK_KNN_EINVAL, "shard outside the allocated reference set");
    Bind b(ctx);
    ctx->dstats_dirty = true;
    ctx->dpack_dirty = true;
    if (n_rows == 0) return 0;
    ctx->tm.start(ctx->st);
    CK(cudaMemcpyAsync(ctx->d_ref.as<double>() + (size_t)row_offset * ctx->ddim, rows, (size_t)n_rows * ctx->ddim * 8,
                       cudaMemcpyHostToDevice, ctx->st), "H2D rows");
    ctx->stats.ms_upload += ctx->tm.stop(ctx->st);
    return 0;
}

int mdsctk_knn_data_reference_arrays(mdsctk_knn_ctx *ctx, int max_arrays, int *n_arrays, void **dev_ptrs,
                                     size_t *bytes_per_row)
{
    if (!ctx || !n_arrays) return MDSCTK_KNN_EINVAL;
    if (!ctx->have_dref) return fail(ctx, MDSCTK_KNN_ESTATE, "no reference set");
    *n_arrays = 1;
    if (max_arrays < 1 || !dev_ptrs || !bytes_per_row) return fail(ctx, MDSCTK_KNN_EINVAL, "need room for 1 array");
    dev_ptrs[0] = ctx->d_ref.p;
    bytes_per_row[0] = (size_t)ctx->ddim * 8;
    ctx->dstats_dirty = true;
    ctx->dpack_dirty = true;
    return 0;
}

int mdsctk_knn_data_set_reference(mdsctk_knn_ctx *ctx, const double *rows, long long n_rows, int dim)
{
    int rc = mdsctk_knn_data_alloc_reference(ctx, n_rows, dim);
    if (rc) return rc;
    return mdsctk_knn_data_upload_shard(ctx, rows, 0, n_rows);
}

int mdsctk_knn_data_query_range(mdsctk_knn_ctx *ctx, long long fit_begin, long long n_fit, int k1, int metric,
                                double *out_dist, int *out_idx)
{
    if (!ctx) return MDSCTK_KNN_EINVAL;
    if (!ctx->have_dref) return fail(ctx, MDSCTK_KNN_ESTATE, "no reference set");
    if (fit_begin < 0 || n_fit <= 0 || fit_begin + n_fit > ctx->dn_ref)
        return fail(ctx, MDSCTK_KNN_EINVAL, "fit range outside the reference set");
    Bind b(ctx);
    return data_run(ctx, ctx->d_ref.as<double>() + (size_t)fit_begin * ctx->ddim, true, fit_begin, n_fit, k1, metric,
                    out_dist, out_idx);
}

int mdsctk_knn_data_query(mdsctk_knn_ctx *ctx, const double *fit_rows, long long n_fit, int k1, int metric,
                          double *out_dist, int *out_idx)
{
    if (!ctx) return MDSCTK_KNN_EINVAL;
    if (!ctx->have_dref) return fail(ctx, MDSCTK_KNN_ESTATE, "no reference set");
    if (!fit_rows) {
        if (n_fit != ctx->dn_ref) return fail(ctx, MDSCTK_KNN_EINVAL, "fit_rows == NULL requires n_fit == n_reference");
        return mdsctk_knn_data_query_range(ctx, 0, n_fit, k1, metric, out_dist, out_idx);
    }
    if (n_fit <= 0) return fail(ctx, MDSCTK_KNN_EINVAL, "n_fit must be positive");
    Bind b(ctx);
    CK(ctx->d_fit.reserve((size_t)n_fit * ctx->ddim * 8), "cudaMalloc(fit rows)");
    ctx->tm.start(ctx->st);
    CK(cudaMemcpyAsync(ctx->d_fit.p, fit_rows, (size_t)n_fit * ctx->ddim * 8, cudaMemcpyHostToDevice, ctx->st),
       "H2D fit rows");
    ctx->stats.ms_upload += ctx->tm.stop(ctx->st);
    return data_run(ctx, ctx->d_fit.as<double>(), false, 0, n_fit, k1, metric, out_dist, out_idx);
}

int mdsctk_knn_data_rows(mdsctk_knn_ctx *ctx, const double *fit_rows, long long n_fit, int metric, double *out)
{
    if (!ctx || !out) return MDSCTK_KNN_EINVAL;
    if (!ctx->have_dref) return fail(ctx, MDSCTK_KNN_ESTATE, "no reference set");
    if (metric != MDSCTK_KNN_EUCLIDEAN && metric != MDSCTK_KNN_CORRELATION) return fail(ctx, MDSCTK_KNN_EINVAL, "unknown metric");
    if (!fit_rows && n_fit != ctx->dn_ref) return fail(ctx, MDSCTK_KNN_EINVAL, "fit_rows == NULL requires n_fit == n_reference");
    if (n_fit <= 0) return fail(ctx, MDSCTK_KNN_EINVAL, "n_fit must be positive");
    Bind b(ctx);
    const int dim = ctx->ddim;
    const long long n_ref = ctx->dn_ref;
    const double *d_fit = ctx->d_ref.as<double>();
    if (fit_rows) {
        CK(ctx->d_fit.reserve((size_t)n_fit * dim * 8), "cudaMalloc(fit rows)");
        CK(cudaMemcpyAsync(ctx->d_fit.p, fit_rows, (size_t)n_fit * dim * 8, cudaMemcpyHostToDevice, ctx->st), "H2D fit rows");
        d_fit = ctx->d_fit.as<double>();
    }
    const double *fit_stats = nullptr, *ref_stats = nullptr;
    if (metric == MDSCTK_KNN_CORRELATION) {
        CK(ctx->d_ref_stats.reserve((size_t)n_ref * 16), "cudaMalloc(ref_stats)");
        CK(launch_data_rowstats(ctx->d_ref.as<double>(), n_ref, dim, ctx->d_ref_stats.as<double>(), ctx->st), "data_rowstats");
        ctx->dstats_dirty = false;
        ref_stats = fit_stats = ctx->d_ref_stats.as<double>();
        if (fit_rows) {
            CK(ctx->d_fit_stats.reserve((size_t)n_fit * 16), "cudaMalloc(fit_stats)");
            CK(launch_data_rowstats(d_fit, n_fit, dim, ctx->d_fit_stats.as<double>(), ctx->st), "data_rowstats");
            fit_stats = ctx->d_fit_stats.as<double>();
        }
    }
    const size_t row_bytes = (size_t)n_ref * 8;
    const long long chunk = (long long)std::max<size_t>(1, std::min<size_t>((size_t)n_fit, ((size_t)256 << 20) / row_bytes));
    CK(ctx->rows_buf.reserve((size_t)chunk * row_bytes), "cudaMalloc(rows_buf)");
    for (long long off = 0; off < n_fit; off += chunk) {
        const long long nr = std::min<long long>(chunk, n_fit - off);
        CK(launch_data_exact_rows(d_fit + (size_t)off * dim, fit_stats ? fit_stats + 2 * off : nullptr, nr, ctx->d_ref.as<double>(),
                                  ref_stats, n_ref, dim, metric, ctx->rows_buf.as<double>(), ctx->st), "data_exact_rows");
        CK(cudaMemcpyAsync(out + (size_t)off * n_ref, ctx->rows_buf.p, (size_t)nr * row_bytes, cudaMemcpyDeviceToHost, ctx->st),
           "D2H rows");
        CK(cudaStreamSynchronize(ctx->st), "sync rows");
    }
    return 0;
}

/* ---------------------------------------------------------------- CSC builder ---- */
static int csc_build(mdsctk_knn_ctx *ctx, int mode, const int *idx, const double *dist, long long n, int maxk, int k, int *pcol,
                     long long *nnz);

int mdsctk_knn_csc_build_sym(mdsctk_knn_ctx *ctx, const int *idx, const double *dist, long long n, int maxk, int k,
                             int *pcol, long long *nnz)
{
    return csc_build(ctx, 0, idx, dist, n, maxk, k, pcol, nnz);
}

int mdsctk_knn_csc_build_general(mdsctk_knn_ctx *ctx, const int *idx, const double *dist, long long n, int maxk, int k,
                                 int symmetric, int *pcol, long long *nnz)
{
    return csc_build(ctx, symmetric ? 2 : 1, idx, dist, n, maxk, k, pcol, nnz);
}

static int csc_build(mdsctk_knn_ctx *ctx, int mode, const int *idx, const double *dist, long long n, int maxk, int k, int *pcol,
                     long long *nnz)
{
    if (!ctx) return MDSCTK_KNN_EINVAL;
    if (!idx || !dist || !pcol || !nnz || n <= 0 || maxk <= 0 || k < 0 || k > maxk)
        return fail(ctx, MDSCTK_KNN_EINVAL, "csc_build_sym: bad arguments (need 0 <= k <= maxk, n > 0)");
    const int fan = mode == 2 ? 2 : 1;          // entries a kNN item can produce
    if ((double)n * (double)std::max(k, 1) * fan >= 2147483647.0)
        return fail(ctx, MDSCTK_KNN_EINVAL, "csc_build: the entry count must stay below 2^31 (int offsets, as in the reference)");
    Bind b(ctx);
    mdsctk_knn_stats &S = ctx->stats;
    S.ms_upload = S.ms_sweep = S.ms_download = 0; S.launches = 0;
    ctx->c_nnz = -1;
    const size_t ne = (size_t)n * maxk, nk = (size_t)n * std::max(k, 1) * fan;
    const size_t nscan = (size_t)n / 1024 + (size_t)n / (1024 * 1024) + 16;
    CK(ctx->c_idx.reserve(ne * 4), "cudaMalloc(csc idx)");
    CK(ctx->c_dist.reserve(ne * 8), "cudaMalloc(csc dist)");
    CK(ctx->c_ints.reserve(((size_t)(n + 1) * 5 + nscan) * 4), "cudaMalloc(csc counters)");
    CK(ctx->c_key.reserve(nk * 8), "cudaMalloc(csc keys)");
    CK(ctx->c_val.reserve(nk * 8), "cudaMalloc(csc vals)");
    CK(ctx->c_irow.reserve(nk * 4), "cudaMalloc(csc irow)");
    CK(ctx->c_oval.reserve(nk * 8), "cudaMalloc(csc val)");
    ctx->tm.start(ctx->st);
    CK(cudaMemcpyAsync(ctx->c_idx.p, idx, ne * 4, cudaMemcpyHostToDevice, ctx->st), "H2D indices");
    CK(cudaMemcpyAsync(ctx->c_dist.p, dist, ne * 8, cudaMemcpyHostToDevice, ctx->st), "H2D distances");
    S.ms_upload = ctx->tm.stop(ctx->st);
    int *ints = ctx->c_ints.as<int>();
    int *cnt = ints, *cur = ints + (n + 1), *off = ints + 2 * (n + 1), *fin = ints + 3 * (n + 1), *d_pcol = ints + 4 * (n + 1);
    int *scan_tmp = ints + 5 * (n + 1);
    ctx->tm.start(ctx->st);
    if (k > 0) {
        CK(launch_csc_build(mode, ctx->c_idx.as<int>(), ctx->c_dist.as<double>(), n, maxk, k, cnt, cur, off, fin, d_pcol, scan_tmp,
                                ctx->c_key.as<unsigned long long>(), ctx->c_val.as<double>(), ctx->c_irow.as<int>(),
                                ctx->c_oval.as<double>(), ctx->st), "csc_build");
        S.launches = 11;
    } else {
        CK(cudaMemsetAsync(d_pcol, 0, (size_t)(n + 1) * 4, ctx->st), "memset pcol");
    }
    S.ms_sweep = ctx->tm.stop(ctx->st);
    CK(cudaGetLastError(), "csc kernels");
    ctx->tm.start(ctx->st);
    CK(cudaMemcpyAsync(pcol, d_pcol, (size_t)(n + 1) * 4, cudaMemcpyDeviceToHost, ctx->st), "D2H pcol");
    S.ms_download = ctx->tm.stop(ctx->st);
    ctx->c_nnz = pcol[n];
    *nnz = ctx->c_nnz;
    S.pairs = (long long)n * k;
    return 0;
}

int mdsctk_knn_csc_fetch(mdsctk_knn_ctx *ctx, int *irow, double *val)
{
    if (!ctx) return MDSCTK_KNN_EINVAL;
    if (ctx->c_nnz < 0) return fail(ctx, MDSCTK_KNN_ESTATE, "csc_fetch: no matrix built");
    if (ctx->c_nnz == 0) return 0;
    if (!irow || !val) return fail(ctx, MDSCTK_KNN_EINVAL, "csc_fetch: NULL output");
    Bind b(ctx);
    ctx->tm.start(ctx->st);
    CK(cudaMemcpyAsync(irow, ctx->c_irow.p, (size_t)ctx->c_nnz * 4, cudaMemcpyDeviceToHost, ctx->st), "D2H irow");
    CK(cudaMemcpyAsync(val, ctx->c_oval.p, (size_t)ctx->c_nnz * 8, cudaMemcpyDeviceToHost, ctx->st), "D2H val");
    ctx->stats.ms_download += ctx->tm.stop(ctx->st);
    return 0;
}

/* ---------------------------------------------------------------- spectral stage ---- */
int mdsctk_knn_spectral_decomp(mdsctk_knn_ctx *ctx, int n, const int *pcol, const int *irow, const double *val, int k_sigma,
                               double sigma, int nev, double *evals, double *evecs, double *residuals, double *avg_sigma, int *n_converged)
{
    return mdsctk_knn_spectral_decomp_ex(ctx, n, pcol, irow, val, k_sigma, sigma, 0.0, nev, evals, evecs, residuals, avg_sigma, n_converged,
                                         nullptr);
}

int mdsctk_knn_spectral_decomp_ex(mdsctk_knn_ctx *ctx, int n, const int *pcol, const int *irow, const double *val, int k_sigma,
                                  double sigma, double k_perplexity, int nev, double *evals, double *evecs, double *residuals,
                                  double *avg_sigma, int *n_converged, double *sigmas)
{
    if (!ctx) return MDSCTK_KNN_EINVAL;
    if (!pcol || !irow || !val || !evals || !evecs || !residuals || n < 2 || nev < 1 || nev >= n)
        return fail(ctx, MDSCTK_KNN_EINVAL, "spectral_decomp: need n >= 2, 1 <= nev < n and non-NULL arrays");
    if (k_perplexity != 0.0 && (!(k_perplexity > 1.0) || k_sigma < 2 || k_sigma > 256 || !(k_perplexity < (double)k_sigma)))
        return fail(ctx, MDSCTK_KNN_EINVAL, "spectral_decomp: entropic affinities need 1 < k_perplexity < k_sigma <= 256");
    const int nnz = pcol[n];
    if (nnz < 0) return fail(ctx, MDSCTK_KNN_EINVAL, "spectral_decomp: bad pcol");
    Bind b(ctx);
    mdsctk_knn_stats &S = ctx->stats;
    S.ms_upload = S.ms_pack = S.ms_sweep = S.ms_download = 0; S.launches = 0;
    const int ncv = std::min(n, 10 * nev + 1);                       // runARPACK, mdsctk.cpp:869
    const size_t nscan = (size_t)n / 1024 + (size_t)n / (1024 * 1024) + 16;
    // ints: pcol (n+1) | irow (nnz) | ptr (n+1) | n_as_row (n+1) | cur (n+1) | scan_tmp | adj_other (2nnz) | adj_pos (2nnz)
    CK(ctx->s_int.reserve(((size_t)4 * (n + 1) + nscan + (size_t)5 * std::max(nnz, 1)) * 4), "cudaMalloc(spectral ints)");
    CK(ctx->s_val.reserve((size_t)std::max(nnz, 1) * 8), "cudaMalloc(spectral values)");
    CK(ctx->s_vec.reserve((size_t)2 * n * 8), "cudaMalloc(sigma, dinv)");
    CK(ctx->s_basis.reserve((size_t)(ncv + 1) * n * 8), "cudaMalloc(Lanczos basis)");
    CK(ctx->s_rot.reserve((size_t)ncv * n * 8), "cudaMalloc(rotation scratch)");
    CK(ctx->s_small.reserve(((size_t)(296 + 2) * (ncv + 1) + (size_t)ncv * ncv) * 8), "cudaMalloc(small)");
    CK(ctx->s_evec.reserve((size_t)nev * n * 8), "cudaMalloc(eigenvectors)");
    int *d_pcol = ctx->s_int.as<int>(), *d_irow = d_pcol + (n + 1), *d_ptr = d_irow + std::max(nnz, 1), *d_asrow = d_ptr + (n + 1);
    int *d_cur = d_asrow + (n + 1), *d_scan = d_cur + (n + 1), *d_other = d_scan + nscan, *d_pos = d_other + 2 * (size_t)std::max(nnz, 1);
    double *d_M = ctx->s_val.as<double>(), *d_sigma = ctx->s_vec.as<double>(), *d_dinv = d_sigma + n;
    ctx->tm.start(ctx->st);
    CK(cudaMemcpyAsync(d_pcol, pcol, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, ctx->st), "H2D pcol");
    CK(cudaMemcpyAsync(d_irow, irow, (size_t)nnz * 4, cudaMemcpyHostToDevice, ctx->st), "H2D irow");
    CK(cudaMemcpyAsync(d_M, val, (size_t)nnz * 8, cudaMemcpyHostToDevice, ctx->st), "H2D values");
    S.ms_upload = ctx->tm.stop(ctx->st);
    ctx->tm.start(ctx->st);
    CK(launch_spectral_adjacency(n, nnz, d_pcol, d_irow, d_ptr, d_asrow, d_cur, d_scan, d_other, d_pos, exclusive_scan, ctx->st),
       "spectral adjacency");
    double avg = 0.0;
    if (k_sigma > 0 || sigma > 0.0) {
        CK(launch_spectral_affinity(n, d_pcol, d_irow, d_ptr, d_pos, k_sigma, sigma, k_perplexity, d_M, d_sigma, d_dinv, ctx->st),
           "spectral affinity");
        std::vector<double> hs((size_t)n);
        CK(cudaMemcpyAsync(hs.data(), d_sigma, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->st), "D2H sigma");
        CK(cudaStreamSynchronize(ctx->st), "sync sigma");
        if (sigmas) memcpy(sigmas, hs.data(), (size_t)n * 8);
        for (int i = 0; i < n; ++i) avg += hs[(size_t)i];       // auto_decomp_sparse.cpp:199-201
        avg /= (double)n;
    }
    if (avg_sigma) *avg_sigma = avg;
    S.ms_pack = ctx->tm.stop(ctx->st);
    CK(cudaGetLastError(), "spectral affinity kernels");
    ctx->tm.start(ctx->st);
    int nconv = 0, nrestart = 0, nspmv = 0;
    CK(spectral_lanczos(n, d_ptr, d_other, d_pos, d_M, nev, ncv, 100 * nev, 1e-13, ctx->s_basis.as<double>(), ctx->s_rot.as<double>(),
                        ctx->s_small.as<double>(), evals, ctx->s_evec.as<double>(), residuals, &nconv, &nrestart, &nspmv, ctx->st),
       "spectral_lanczos");
    S.ms_sweep = ctx->tm.stop(ctx->st);
    if (n_converged) *n_converged = nconv;
    if (nconv < 0) return fail(ctx, MDSCTK_KNN_ESTATE, "spectral_decomp: the operator has fewer than nev independent directions");
    S.launches = nspmv; S.rescored_max = nrestart;
    ctx->tm.start(ctx->st);
    CK(cudaMemcpyAsync(evecs, ctx->s_evec.p, (size_t)nev * n * 8, cudaMemcpyDeviceToHost, ctx->st), "D2H eigenvectors");
    S.ms_download = ctx->tm.stop(ctx->st);
    S.pairs = (long long)nnz;
    return 0;
}

/* ---------------------------------------------------------------- featurisers ---- */
int mdsctk_knn_phipsi(mdsctk_knn_ctx *ctx, const float *xyz, long long n_frames, int n_atoms, double *phipsi, double *sincos)
{
    if (!ctx) return MDSCTK_KNN_EINVAL;
    const int T = 2 * (n_atoms / 3) - 2;
    if (!xyz || n_frames <= 0 || n_atoms < 6 || T <= 0 || (!phipsi && !sincos))
        return fail(ctx, MDSCTK_KNN_EINVAL, "phipsi: need frames of at least 6 backbone atoms and an output");
    Bind b(ctx);
    mdsctk_knn_stats &S = ctx->stats;
    S.ms_upload = S.ms_sweep = S.ms_download = 0; S.launches = 1;
    const size_t na = (size_t)n_frames * T;
    CK(ctx->f_in.reserve((size_t)n_frames * n_atoms * 12), "cudaMalloc(frames)");
    if (phipsi) CK(ctx->f_ang.reserve(na * 8), "cudaMalloc(angles)");
    if (sincos) CK(ctx->f_sc.reserve(na * 16), "cudaMalloc(sincos)");
    ctx->tm.start(ctx->st);
    CK(cudaMemcpyAsync(ctx->f_in.p, xyz, (size_t)n_frames * n_atoms * 12, cudaMemcpyHostToDevice, ctx->st), "H2D frames");
    S.ms_upload = ctx->tm.stop(ctx->st);
    ctx->tm.start(ctx->st);
    CK(launch_phipsi(ctx->f_in.as<float>(), n_frames, n_atoms, phipsi ? ctx->f_ang.as<double>() : nullptr,
                     sincos ? ctx->f_sc.as<double>() : nullptr, ctx->st), "phipsi");
    S.ms_sweep = ctx->tm.stop(ctx->st);
    CK(cudaGetLastError(), "phipsi kernel");
    ctx->tm.start(ctx->st);
    if (phipsi) CK(cudaMemcpyAsync(phipsi, ctx->f_ang.p, na * 8, cudaMemcpyDeviceToHost, ctx->st), "D2H angles");
    if (sincos) CK(cudaMemcpyAsync(sincos, ctx->f_sc.p, na * 16, cudaMemcpyDeviceToHost, ctx->st), "D2H sincos");
    S.ms_download = ctx->tm.stop(ctx->st);
    S.pairs = (long long)na;
    return 0;
}

int mdsctk_knn_sincos(mdsctk_knn_ctx *ctx, const double *angles, long long n, double *out)
{
    if (!ctx) return MDSCTK_KNN_EINVAL;
    if (!angles || !out || n <= 0) return fail(ctx, MDSCTK_KNN_EINVAL, "sincos: bad arguments");
    Bind b(ctx);
    CK(ctx->f_ang.reserve((size_t)n * 8), "cudaMalloc(angles)");
    CK(ctx->f_sc.reserve((size_t)n * 16), "cudaMalloc(sincos)");
    CK(cudaMemcpyAsync(ctx->f_ang.p, angles, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->st), "H2D angles");
    CK(launch_sincos(ctx->f_ang.as<double>(), n, ctx->f_sc.as<double>(), ctx->st), "sincos");
    CK(cudaMemcpyAsync(out, ctx->f_sc.p, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->st), "D2H sincos");
    CK(cudaStreamSynchronize(ctx->st), "sync sincos");
    return 0;
}

int mdsctk_knn_debug_fetch_tile(mdsctk_knn_ctx *ctx, float *out)
{
    if (!ctx || !out) return MDSCTK_KNN_EINVAL;
    if (!ctx->debug_tile.p) return fail(ctx, MDSCTK_KNN_ESTATE, "no debug tile captured");
    Bind b(ctx);
    CK(cudaMemcpy(out, ctx->debug_tile.p, 128 * 432 * 4, cudaMemcpyDeviceToHost), "D2H debug tile");
    return 0;
}

int mdsctk_knn_debug_fetch_array(mdsctk_knn_ctx *ctx, int which, void *out, size_t out_capacity, size_t *n_bytes)
{
    if (!ctx || !n_bytes) return MDSCTK_KNN_EINVAL;
    if (!ctx->have_ref) return fail(ctx, MDSCTK_KNN_ESTATE, "no reference set");
    const FrameSet &R = ctx->ref;
    const size_t n = (size_t)R.n, p4 = n * 3 * R.A_pad * 4, p2 = n * 3 * R.A_pad * 2;
    struct { const DevBuf *b; size_t bytes; int fam; } tab[14] = {
        {&R.raw, n * R.A * 12, 0}, {&R.G, n * 4, 0}, {&R.cen, n * 32, 0}, {&R.sig, n * 16, 0}, {&R.Gh, n * 4, F_FP16},
        {&R.G2, n * 4, F_FP16LO}, {&R.gres, n * 8, F_FP16}, {&R.planes, p4, F_PLANES}, {&R.hi, p4, F_TF32}, {&R.lo, p4, F_TF32},
        {&R.bh, p2, F_BF16}, {&R.bm, p2, F_BF16}, {&R.fh, p2, F_FP16}, {&R.fl, p2, F_FP16LO}};
    if (which < 0 || which >= 14) return fail(ctx, MDSCTK_KNN_EINVAL, "debug_fetch_array: which must be 0..13");
    if ((tab[which].fam & ~R.have) != 0 || !tab[which].b->p) return fail(ctx, MDSCTK_KNN_ESTATE, "that array is not packed for the current kernel");
    *n_bytes = tab[which].bytes;
    if (!out) return 0;
    if (out_capacity < tab[which].bytes) return fail(ctx, MDSCTK_KNN_EINVAL, "debug_fetch_array: buffer too small");
    Bind b(ctx);
    CK(cudaStreamSynchronize(ctx->st), "sync");
    CK(cudaMemcpy(out, tab[which].b->p, tab[which].bytes, cudaMemcpyDeviceToHost), "D2H array");
    return 0;
}

int mdsctk_knn_debug_rms_layout(int n_atoms, int wide_stages, int *out12)
{
    if (n_atoms <= 0 || !out12) return MDSCTK_KNN_EINVAL;
    rms_tc2_layout_info((n_atoms + 15) / 16 * 16, wide_stages, out12);
    return 0;
}

int mdsctk_knn_timer_start(mdsctk_knn_ctx *ctx)
{
    if (!ctx) return MDSCTK_KNN_EINVAL;
    Bind b(ctx);
    ctx->user_tm.start(ctx->st);
    CK(cudaGetLastError(), "timer_start");
    return 0;
}

int mdsctk_knn_timer_stop(mdsctk_knn_ctx *ctx, double *elapsed_ms)
{
    if (!ctx || !elapsed_ms) return MDSCTK_KNN_EINVAL;
    Bind b(ctx);
    *elapsed_ms = ctx->user_tm.stop(ctx->st);
    CK(cudaGetLastError(), "timer_stop");
    return 0;
}

int mdsctk_knn_fetch(mdsctk_knn_ctx *ctx, double *out_dist, int *out_idx)
{
    if (!ctx) return MDSCTK_KNN_EINVAL;
    if (ctx->out_rows <= 0) return fail(ctx, MDSCTK_KNN_ESTATE, "no query results to fetch");
    Bind b(ctx);
    const size_t n = (size_t)ctx->out_rows * ctx->out_k1;
    ctx->tm.start(ctx->st);
    if (out_dist) CK(cudaMemcpyAsync(out_dist, ctx->out_dist.p, n * 8, cudaMemcpyDeviceToHost, ctx->st), "D2H distances");
    if (out_idx) CK(cudaMemcpyAsync(out_idx, ctx->out_idx.p, n * 4, cudaMemcpyDeviceToHost, ctx->st), "D2H indices");
    ctx->stats.ms_download = ctx->tm.stop(ctx->st);
    CK(cudaGetLastError(), "fetch");
    return 0;
}

}  // extern "C"
