// spectral.cu -- affinity / normalised-Laplacian stage and eigen-solve of auto_decomp_sparse on the GPU.
//
// Replaces auto_decomp_sparse.cpp:150-198 (distances of the symmetric CSC matrix -> Gaussian affinities with
// per-frame sigmas -> D^-1/2 W D^-1/2) and runARPACK (mdsctk.cpp:857-924: the nev algebraically largest
// eigenpairs, ARPACK dsaupd/dseupd with ncv = 10*nev+1 and sp_dsymv, mdsctk.cpp:293-313, as the operator).
//   adjacency   the matrix is stored as its strict upper triangle in CSC; every frame needs both its column
//               entries and the entries where it is the row, so a full adjacency (CSR over both directions,
//               pointing back into the value array) is built once: count, scan, scatter, per-frame sort by
//               the other endpoint.
//   sigma       the reference collects a frame's values in CSC traversal order -- entries where it is the ROW
//               (ascending column) first, then its own column -- keeps the first k_a, and divides their sum
//               by k_a (auto_decomp_sparse.cpp:156-170); same rule here, summed in ascending value order.
//   SpMV        y = A x by gathering over the adjacency (no atomics, deterministic); HBM-bound: 12 B per
//               stored entry per direction.
//   eigen-solve thick-restart Lanczos with full re-orthogonalisation in FP64, basis of ncv = 10*nev+1 vectors
//               like ARPACK's, Ritz problem of the small projected matrix solved on the host (Jacobi);
//               converged when every wanted Ritz pair has residual <= tol * |theta|.
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace mdsctk {

namespace spk {
constexpr int BLOCK = 256;
constexpr int DOT_CHUNKS = 296;      // partial sums per basis vector (2 per SM)
}  // namespace spk

// ------------------------------------------------------------------ adjacency ----
__global__ void adj_count_kernel(const int *__restrict__ pcol, const int *__restrict__ irow, int n, int *__restrict__ deg)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int b = pcol[c], e = pcol[c + 1];
    atomicAdd(&deg[c], e - b);
    for (int y = b; y < e; ++y) atomicAdd(&deg[irow[y]], 1);
}

// adj_other[p] = the other endpoint, adj_pos[p] = index into the value array.  The entries where the frame is
// the ROW come first (they are sorted by column afterwards), then the frame's own column in row order.
__global__ void adj_fill_kernel(const int *__restrict__ pcol, const int *__restrict__ irow, int n, const int *__restrict__ ptr,
                                const int *__restrict__ n_as_row, int *__restrict__ cur, int *__restrict__ adj_other,
                                int *__restrict__ adj_pos)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int b = pcol[c], e = pcol[c + 1];
    const int own0 = ptr[c] + n_as_row[c];
    for (int y = b; y < e; ++y) {
        adj_other[own0 + (y - b)] = irow[y];
        adj_pos[own0 + (y - b)] = y;
        const int r = irow[y];
        const int p = ptr[r] + atomicAdd(&cur[r], 1);
        adj_other[p] = c;
        adj_pos[p] = y;
    }
}

__global__ void as_row_count_kernel(const int *__restrict__ irow, int nnz, int *__restrict__ n_as_row)
{
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y < nnz) atomicAdd(&n_as_row[irow[y]], 1);
}

// one thread per frame: insertion sort of its "as row" entries by column (short lists; hubs are rare)
__global__ void adj_sort_rows_kernel(int n, const int *__restrict__ ptr, const int *__restrict__ n_as_row,
                                     int *__restrict__ adj_other, int *__restrict__ adj_pos)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int b = ptr[v], m = n_as_row[v];
    for (int a = 1; a < m; ++a) {
        const int o = adj_other[b + a], p = adj_pos[b + a];
        int i = a - 1;
        while (i >= 0 && adj_other[b + i] > o) { adj_other[b + i + 1] = adj_other[b + i]; adj_pos[b + i + 1] = adj_pos[b + i]; --i; }
        adj_other[b + i + 1] = o; adj_pos[b + i + 1] = p;
    }
}

// ------------------------------------------------------------------ affinity ----
__global__ void sigma_kernel(int n, const int *__restrict__ ptr, const int *__restrict__ adj_pos, const double *__restrict__ M,
                             int k_a, double *__restrict__ sigma)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int b = ptr[v], m = min(ptr[v + 1] - b, k_a);
    // ascending-order sum of the first m values (the reference sorts them before adding): m selections of the next
    // (value, position) pair in lexicographic order -- always m iterations, whatever the values (duplicates, infinities);
    // NaNs compare false everywhere, are never selected, and poison the sum as they do in the reference
    double s = 0.0, lastv = -__longlong_as_double(0x7ff0000000000000LL);
    int lasti = -1;
    for (int t = 0; t < m; ++t) {
        double cur = 0.0;
        int ci = -1;
        for (int i = 0; i < m; ++i) {
            const double x = M[adj_pos[b + i]];
            if ((x > lastv || (x == lastv && i > lasti)) && (ci < 0 || x < cur)) { cur = x; ci = i; }
        }
        if (ci < 0) { s += __longlong_as_double(0x7ff8000000000000LL); break; }
        s += cur;
        lastv = cur; lasti = ci;
    }
    sigma[v] = s / (double)k_a;
}

// Entropic affinities (auto_decomp_sparse -K, entropic_affinity_sigmas / entropic_affinity_sigma, mdsctk.cpp:388-565): for
// every frame the bandwidth whose Gaussian over the frame's first k_a (sorted) distances has perplexity K, i.e. the
// root in b = log beta of  e(b) = beta m1(b) + log m0(b) - log K,  m0 = sum exp(-d^2 beta), m1 = sum d^2 exp(-d^2 beta) / m0;
// sigma = 1 / sqrt(2 beta).  Same bracket [B_lower, B_upper], same safeguarded Newton iteration (bisection whenever the
// function or its gradient misbehaves, or the step leaves the bracket; a forced bisection every 20 steps), same stopping
// rules (|e| < 1e-10, bracket narrower than 10 sqrt(eps)).  One thread per frame.  The reference walks the frames in the
// order of their K-th distance and starts each from the previous frame's solution, falling back to the bracket's midpoint
// when that lies outside -- a warm start only; here every frame starts from its midpoint (frames are independent), so the
// two agree to the iteration's tolerance, not to the last bit.  Frames with fewer than two entries keep the mean sigma.
constexpr int ENTROPIC_MAX_K = 256;

__global__ void entropic_sigma_kernel(int n, const int *__restrict__ ptr, const int *__restrict__ adj_pos, const double *__restrict__ M,
                                      int k_a, double logK, double logN, double p1, double *__restrict__ sigma)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int b0 = ptr[v], k = min(min(ptr[v + 1] - b0, k_a), ENTROPIC_MAX_K);
    if (k < 2) return;
    double a2[ENTROPIC_MAX_K];                        // squared distances, ascending (the reference sorts sorted_A[x])
    for (int i = 0; i < k; ++i) {
        const double d = M[adj_pos[b0 + i]];
        int j = i - 1;
        while (j >= 0 && a2[j] > d * d) { a2[j + 1] = a2[j]; --j; }
        a2[j + 1] = d * d;
    }
    const double N = (double)k_a, logNK = logN - logK;
    double BU = log((2.0 * log(p1 * (N - 1.0) / (1.0 - p1))) / (a2[1] - a2[0]));
    const double bL1 = log((2.0 * logNK / (1.0 - (1.0 / N))) / (a2[k - 1] - a2[0]));
    const double bL2 = log((2.0 * sqrt(logNK)) / sqrt(a2[k - 1] * a2[k - 1] - a2[0] * a2[0]));
    double BL = bL1 > bL2 ? bL1 : bL2;
    const double tol = 1e-10, realmin = 2.225074e-308, eps = 1.4901161193847656e-08;   // sqrt(2^-52), getEPS()
    double b = 0.5 * (BL + BU);
    int it = 1;
    for (int guard = 0; guard < 100000; ++guard) {
        const double bE = exp(b);
        double m0 = 0.0;
        for (int x = 0; x < k; ++x) m0 += exp(-a2[x] * bE);
        bool pbm = false;
        double e, m1 = 0.0;
        if (m0 < realmin) {
            e = -logK;
            pbm = true;
        } else {
            for (int x = 0; x < k; ++x) m1 += exp(-a2[x] * bE) * (a2[x] / m0);
            e = bE * m1 + log(m0) - logK;
        }
        if (fabs(e) < tol) break;
        if (BU - BL < 10.0 * eps) break;
        if (e < 0.0 && b <= BU) BU = b;
        else if (e > 0.0 && b >= BL) BL = b;
        pbm = pbm || e < -logK || e > logN - logK;
        double g = 0.0;
        if (!pbm) {
            if (it == 20) { b = 0.5 * (BL + BU); it = 1; continue; }
            double m2 = 0.0;
            for (int x = 0; x < k; ++x) m2 += exp(-a2[x] * bE) * (a2[x] / m0) * a2[x];
            g = bE * bE * (m1 * m1 - m2);
            if (g == 0.0) pbm = true;
        }
        if (pbm) {
            double esum = 0.0;
            for (int x = 0; x < k; ++x) esum += exp(-a2[x] * exp(BL)) + exp(-a2[x] * exp(BU));
            if (esum < 2.0 * sqrt(realmin)) break;
            b = 0.5 * (BL + BU);
            it = 1;
            continue;
        }
        b += -e / g;
        if (b < BL || b > BU) { b = 0.5 * (BL + BU); it = 0; }
        ++it;
    }
    sigma[v] = 1.0 / sqrt(2.0 * exp(b));
}

__global__ void affinity_kernel(int n, const int *__restrict__ pcol, const int *__restrict__ irow, const double *__restrict__ sigma,
                                double *__restrict__ M)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    for (int y = pcol[c]; y < pcol[c + 1]; ++y) M[y] = exp(-(M[y] * M[y]) / (2.0 * sigma[c] * sigma[irow[y]]));
}

__global__ void degree_kernel(int n, const int *__restrict__ ptr, const int *__restrict__ adj_pos, const double *__restrict__ M,
                              double *__restrict__ dinv)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    double d = 0.0;
    for (int p = ptr[v]; p < ptr[v + 1]; ++p) d += M[adj_pos[p]];
    dinv[v] = 1.0 / sqrt(d);
}

__global__ void normalise_kernel(int n, const int *__restrict__ pcol, const int *__restrict__ irow, const double *__restrict__ dinv,
                                 double *__restrict__ M)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    for (int y = pcol[c]; y < pcol[c + 1]; ++y) M[y] *= dinv[irow[y]] * dinv[c];
}

// ------------------------------------------------------------------ vectors ----
__global__ void spmv_adj_kernel(int n, const int *__restrict__ ptr, const int *__restrict__ adj_other, const int *__restrict__ adj_pos,
                                const double *__restrict__ M, const double *__restrict__ x, double *__restrict__ y)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    double s = 0.0;
    for (int p = ptr[v]; p < ptr[v + 1]; ++p) s += M[adj_pos[p]] * x[adj_other[p]];
    y[v] = s;
}

// partial[chunk][j] = sum over the chunk's rows of V[:, j] * w      (V column-major, ld = n)
__global__ void __launch_bounds__(spk::BLOCK) dots_partial_kernel(const double *__restrict__ V, int n, int nv, const double *__restrict__ w,
                                                                  double *__restrict__ partial)
{
    __shared__ double s_red[spk::BLOCK / 32];
    const int chunk = blockIdx.x, j = blockIdx.y;
    const long long per = ((long long)n + gridDim.x - 1) / gridDim.x;
    const long long b = chunk * per, e = min((long long)n, b + per);
    const double *col = V + (size_t)j * n;
    double s = 0.0;
    for (long long i = b + threadIdx.x; i < e; i += blockDim.x) s += col[i] * w[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < spk::BLOCK / 32; ++k) t += s_red[k];
        partial[(size_t)chunk * nv + j] = t;
    }
}

__global__ void dots_final_kernel(const double *__restrict__ partial, int chunks, int nv, double *__restrict__ out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nv) return;
    double t = 0.0;
    for (int c = 0; c < chunks; ++c) t += partial[(size_t)c * nv + j];
    out[j] = t;
}

// w -= V[:, 0..nv) h
__global__ void subtract_basis_kernel(const double *__restrict__ V, int n, int nv, const double *__restrict__ h, double *__restrict__ w)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = w[i];
    for (int j = 0; j < nv; ++j) s -= V[(size_t)j * n + i] * h[j];
    w[i] = s;
}

__global__ void scale_copy_kernel(const double *__restrict__ src, int n, double alpha, double *__restrict__ dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i] * alpha;
}

// out[:, c] = sum_j V[:, j] S[j][c]   (S: nv x nc, row-major, in global memory; 8 output columns per block.y)
__global__ void rotate_basis_kernel(const double *__restrict__ V, int n, int nv, const double *__restrict__ S, int nc,
                                    double *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int c0 = blockIdx.y * 8;
    if (i >= n) return;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < nv; ++j) {
        const double v = V[(size_t)j * n + i];
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (c0 + c < nc) acc[c] += v * S[(size_t)j * nc + c0 + c];
    }
#pragma unroll
    for (int c = 0; c < 8; ++c)
        if (c0 + c < nc) out[(size_t)(c0 + c) * n + i] = acc[c];
}

__global__ void init_vector_kernel(double *v, int n, unsigned long long seed)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);     // splitmix64
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    v[i] = (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}

// ------------------------------------------------------------------ host side ----
// cyclic Jacobi on a small dense symmetric matrix (row-major m x m); eigenvalues ascending, eigenvectors in the COLUMNS of z
static void small_sym_eig(int m, std::vector<double> &a, std::vector<double> &w, std::vector<double> &z)
{
    z.assign((size_t)m * m, 0.0);
    for (int i = 0; i < m; ++i) z[(size_t)i * m + i] = 1.0;
    for (int sweep = 0; sweep < 80; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int p = 0; p < m; ++p) { diag += a[(size_t)p * m + p] * a[(size_t)p * m + p]; for (int q = p + 1; q < m; ++q) off += a[(size_t)p * m + q] * a[(size_t)p * m + q]; }
        if (off <= 1e-32 * diag || off < 1e-300) break;
        for (int p = 0; p < m - 1; ++p)
            for (int q = p + 1; q < m; ++q) {
                const double apq = a[(size_t)p * m + q];
                if (std::fabs(apq) < 1e-300) continue;
                const double theta = (a[(size_t)q * m + q] - a[(size_t)p * m + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < m; ++k) {
                    const double akp = a[(size_t)k * m + p], akq = a[(size_t)k * m + q];
                    a[(size_t)k * m + p] = c * akp - s * akq; a[(size_t)k * m + q] = s * akp + c * akq;
                }
                for (int k = 0; k < m; ++k) {
                    const double apk = a[(size_t)p * m + k], aqk = a[(size_t)q * m + k];
                    a[(size_t)p * m + k] = c * apk - s * aqk; a[(size_t)q * m + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < m; ++k) {
                    const double zkp = z[(size_t)k * m + p], zkq = z[(size_t)k * m + q];
                    z[(size_t)k * m + p] = c * zkp - s * zkq; z[(size_t)k * m + q] = s * zkp + c * zkq;
                }
            }
    }
    w.resize(m);
    std::vector<int> order(m);
    for (int i = 0; i < m; ++i) { w[i] = a[(size_t)i * m + i]; order[i] = i; }
    std::sort(order.begin(), order.end(), [&](int x, int y) { return w[x] < w[y]; });
    std::vector<double> w2(m), z2((size_t)m * m);
    for (int c = 0; c < m; ++c) { w2[c] = w[order[c]]; for (int r = 0; r < m; ++r) z2[(size_t)r * m + c] = z[(size_t)r * m + order[c]]; }
    w.swap(w2); z.swap(z2);
}

#define SP_CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return e_; } while (0)

static inline unsigned grid_for(long long n) { return (unsigned)((n + spk::BLOCK - 1) / spk::BLOCK); }

// Builds the adjacency of the strict-upper-CSC matrix.  ints: workspace of 4*(n+1) + 2*nnz*... see sizes below.
cudaError_t launch_spectral_adjacency(int n, int nnz, const int *pcol, const int *irow, int *deg_ptr /*n+1*/, int *n_as_row /*n+1*/,
                                      int *cur /*n+1*/, int *scan_tmp, int *adj_other /*2nnz*/, int *adj_pos /*2nnz*/,
                                      cudaError_t (*scan)(const int *, int *, long long, int *, cudaStream_t), cudaStream_t st)
{
    SP_CK(cudaMemsetAsync(deg_ptr, 0, (size_t)(n + 1) * 4, st));
    SP_CK(cudaMemsetAsync(n_as_row, 0, (size_t)(n + 1) * 4, st));
    SP_CK(cudaMemsetAsync(cur, 0, (size_t)(n + 1) * 4, st));
    adj_count_kernel<<<grid_for(n), spk::BLOCK, 0, st>>>(pcol, irow, n, deg_ptr);
    if (nnz > 0) as_row_count_kernel<<<grid_for(nnz), spk::BLOCK, 0, st>>>(irow, nnz, n_as_row);
    SP_CK(scan(deg_ptr, deg_ptr, n + 1, scan_tmp, st));
    adj_fill_kernel<<<grid_for(n), spk::BLOCK, 0, st>>>(pcol, irow, n, deg_ptr, n_as_row, cur, adj_other, adj_pos);
    adj_sort_rows_kernel<<<grid_for(n), spk::BLOCK, 0, st>>>(n, deg_ptr, n_as_row, adj_other, adj_pos);
    return cudaGetLastError();
}

__global__ void fill_double_kernel(double *v, int n, double x)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = x;
}

// k_a > 0: per-frame sigmas (auto_decomp_sparse); else the global sigma0 (decomp_sparse.cpp:157-158)
cudaError_t launch_spectral_affinity(int n, const int *pcol, const int *irow, const int *ptr, const int *adj_pos, int k_a, double sigma0,
                                     double perplexity, double *M, double *sigma, double *dinv, cudaStream_t st)
{
    if (k_a > 0) sigma_kernel<<<grid_for(n), spk::BLOCK, 0, st>>>(n, ptr, adj_pos, M, k_a, sigma);
    else fill_double_kernel<<<grid_for(n), spk::BLOCK, 0, st>>>(sigma, n, sigma0);
    if (k_a > 0 && perplexity > 0.0) {
        // p1(N, log K), entropic_affinity_sigmas (mdsctk.cpp:515-524)
        const double N = (double)k_a, logK = std::log(perplexity), logN = std::log(N);
        double p1;
        if (logK > std::log(std::sqrt(2.0 * N))) {
            p1 = 3.0 / 4.0;
        } else {
            p1 = 1.0 / 4.0;
            for (int x = 0; x < 100; x++) p1 -= (-p1 * std::log(p1 / N) - logK) / (-std::log(p1 / N) + 1.0);
            p1 = 1.0 - (p1 / 2.0);
        }
        entropic_sigma_kernel<<<(n + 63) / 64, 64, 0, st>>>(n, ptr, adj_pos, M, k_a, logK, logN, p1, sigma);
    }
    affinity_kernel<<<grid_for(n), spk::BLOCK, 0, st>>>(n, pcol, irow, sigma, M);
    degree_kernel<<<grid_for(n), spk::BLOCK, 0, st>>>(n, ptr, adj_pos, M, dinv);
    normalise_kernel<<<grid_for(n), spk::BLOCK, 0, st>>>(n, pcol, irow, dinv, M);
    return cudaGetLastError();
}

cudaError_t launch_spectral_spmv(int n, const int *ptr, const int *adj_other, const int *adj_pos, const double *M, const double *x,
                                 double *y, cudaStream_t st)
{
    spmv_adj_kernel<<<grid_for(n), spk::BLOCK, 0, st>>>(n, ptr, adj_other, adj_pos, M, x, y);
    return cudaGetLastError();
}

// Thick-restart Lanczos for the nev algebraically largest eigenpairs.
//   V: device, (ncv + 1) * n doubles; W: device, ncv * n doubles (rotation scratch); small: device, (DOT_CHUNKS + 2) * (ncv + 1) + ncv*ncv doubles
//   evals[nev] descending, d_evecs: device nev * n (row e = eigenvector e), residuals[nev] = |A z - d z| / |d|
//   returns the number of converged pairs in *n_conv, restarts in *n_restart
cudaError_t spectral_lanczos(int n, const int *ptr, const int *adj_other, const int *adj_pos, const double *M, int nev, int ncv,
                             int max_restarts, double tol, double *V, double *W, double *small, double *evals, double *d_evecs,
                             double *residuals, int *n_conv, int *n_restart, int *n_spmv, cudaStream_t st)
{
    const int m = ncv;
    double *d_partial = small, *d_h = small + (size_t)spk::DOT_CHUNKS * (m + 1), *d_S = d_h + 2 * (m + 1);
    std::vector<double> T((size_t)m * m, 0.0), h(m + 1), h2(m + 1), theta, S;
    const int chunks = std::max(1, std::min(spk::DOT_CHUNKS, n / 4096 + 1));
    auto dots = [&](const double *basis, int nv, const double *w, double *out_host) -> cudaError_t {
        dim3 g(chunks, nv);
        dots_partial_kernel<<<g, spk::BLOCK, 0, st>>>(basis, n, nv, w, d_partial);
        dots_final_kernel<<<(nv + 127) / 128, 128, 0, st>>>(d_partial, chunks, nv, d_h);
        SP_CK(cudaMemcpyAsync(out_host, d_h, (size_t)nv * 8, cudaMemcpyDeviceToHost, st));
        return cudaStreamSynchronize(st);
    };
    auto subtract = [&](const double *basis, int nv, const double *coef_host, double *w) -> cudaError_t {
        SP_CK(cudaMemcpyAsync(d_h + (m + 1), coef_host, (size_t)nv * 8, cudaMemcpyHostToDevice, st));
        subtract_basis_kernel<<<grid_for(n), spk::BLOCK, 0, st>>>(basis, n, nv, d_h + (m + 1), w);
        return cudaGetLastError();
    };
    // v0
    init_vector_kernel<<<grid_for(n), spk::BLOCK, 0, st>>>(V, n, 0x5eedULL);
    double nrm2 = 0.0;
    SP_CK(dots(V, 1, V, &nrm2));
    scale_copy_kernel<<<grid_for(n), spk::BLOCK, 0, st>>>(V, n, 1.0 / std::sqrt(nrm2), V);

    int l = 0;                       // basis vectors kept from the previous restart (V[:, 0..l) are Ritz vectors, V[:, l] the next Lanczos vector)
    double beta_m = 0.0;
    int restarts = 0, spmvs = 0, mm = m;
    *n_conv = 0;
    for (;; ++restarts) {
        mm = m;
        for (int j = l; j < m; ++j) {
            double *w = V + (size_t)(j + 1) * n;
            SP_CK(launch_spectral_spmv(n, ptr, adj_other, adj_pos, M, V + (size_t)j * n, w, st));
            ++spmvs;
            // full re-orthogonalisation, twice (classical Gram-Schmidt with refinement)
            SP_CK(dots(V, j + 1, w, h.data()));
            SP_CK(subtract(V, j + 1, h.data(), w));
            SP_CK(dots(V, j + 1, w, h2.data()));
            SP_CK(subtract(V, j + 1, h2.data(), w));
            for (int i = 0; i <= j; ++i) { const double t = h[i] + h2[i]; T[(size_t)i * m + j] = t; T[(size_t)j * m + i] = t; }
            double b2 = 0.0;
            SP_CK(dots(w, 1, w, &b2));
            const double beta = std::sqrt(std::max(b2, 0.0));
            beta_m = beta;
            if (!(beta > 1e-14)) {
                // invariant subspace (a disconnected component, or fewer distinct directions than basis vectors): the
                // recurrence has nothing to continue with.  Carry on from a fresh random vector orthogonal to the basis
                // (its coupling to the previous vector is exactly the zero beta), so that the basis still reaches ncv
                // vectors and all nev pairs come out; only when no such direction is left is the basis truly complete.
                beta_m = 0.0;
                if (j + 1 >= m) break;
                init_vector_kernel<<<grid_for(n), spk::BLOCK, 0, st>>>(w, n, 0x5eedULL + 977ULL * (unsigned long long)(j + 1) + 131071ULL * (unsigned long long)restarts);
                double r0 = 0.0, r1 = 0.0;
                SP_CK(dots(w, 1, w, &r0));
                for (int pass = 0; pass < 2; ++pass) {
                    SP_CK(dots(V, j + 1, w, h2.data()));
                    SP_CK(subtract(V, j + 1, h2.data(), w));
                }
                SP_CK(dots(w, 1, w, &r1));
                if (!(r1 > 1e-16 * r0)) { mm = j + 1; break; }
                scale_copy_kernel<<<grid_for(n), spk::BLOCK, 0, st>>>(w, n, 1.0 / std::sqrt(r1), w);
                continue;
            }
            scale_copy_kernel<<<grid_for(n), spk::BLOCK, 0, st>>>(w, n, 1.0 / beta, w);
        }
        // Ritz problem of the leading mm x mm block
        std::vector<double> Tm((size_t)mm * mm);
        for (int i = 0; i < mm; ++i) for (int j = 0; j < mm; ++j) Tm[(size_t)i * mm + j] = T[(size_t)i * m + j];
        small_sym_eig(mm, Tm, theta, S);      // ascending; eigenvectors in columns of S (mm x mm row-major)
        const int want = std::min(nev, mm);
        int conv = 0;
        // ARPACK's test with tol = machine epsilon (dsaupd, mdsctk.cpp:869-893): |r| <= tol * max(eps^(2/3), |theta|), plus
        // the rounding floor eps * |A| below which no residual estimate means anything (small eigenvalues converge too)
        double anorm = 0.0;
        for (int i = 0; i < mm; ++i) anorm = std::max(anorm, std::fabs(theta[i]));
        for (int e = 0; e < want; ++e) {
            const int c = mm - 1 - e;
            const double r = std::fabs(beta_m * S[(size_t)(mm - 1) * mm + c]);
            if (r <= tol * std::max(std::fabs(theta[c]), 3.7e-11) + 4.5e-16 * anorm) ++conv;
        }
        *n_conv = conv;
        if (conv >= want || restarts >= max_restarts || mm < m) {
            // final Ritz vectors: columns mm-1-e of S
            std::vector<double> Ssel((size_t)mm * want);
            for (int r = 0; r < mm; ++r) for (int e = 0; e < want; ++e) Ssel[(size_t)r * want + e] = S[(size_t)r * mm + (mm - 1 - e)];
            SP_CK(cudaMemcpyAsync(d_S, Ssel.data(), Ssel.size() * 8, cudaMemcpyHostToDevice, st));
            dim3 g(grid_for(n), (want + 7) / 8);
            rotate_basis_kernel<<<g, spk::BLOCK, 0, st>>>(V, n, mm, d_S, want, d_evecs);
            for (int e = 0; e < want; ++e) evals[e] = theta[mm - 1 - e];
            // true residuals |A z - d z| / |d| (auto_decomp_sparse.cpp:221-229)
            for (int e = 0; e < want; ++e) {
                double *z = d_evecs + (size_t)e * n;
                SP_CK(launch_spectral_spmv(n, ptr, adj_other, adj_pos, M, z, W, st));
                const double coef = evals[e];
                SP_CK(subtract(z, 1, &coef, W));
                double r2 = 0.0;
                SP_CK(dots(W, 1, W, &r2));
                residuals[e] = std::sqrt(std::max(r2, 0.0)) / std::fabs(evals[e]);
            }
            if (want < nev) *n_conv = -1;      // fewer independent directions than requested pairs: the caller reports it
            break;
        }
        // thick restart: keep the k largest Ritz pairs and the residual vector
        const int k = std::min(mm - 1, nev + std::max(1, (mm - nev) / 2));
        std::vector<double> Ssel((size_t)mm * k);
        for (int r = 0; r < mm; ++r) for (int e = 0; e < k; ++e) Ssel[(size_t)r * k + e] = S[(size_t)r * mm + (mm - 1 - e)];
        SP_CK(cudaMemcpyAsync(d_S, Ssel.data(), Ssel.size() * 8, cudaMemcpyHostToDevice, st));
        dim3 g(grid_for(n), (k + 7) / 8);
        rotate_basis_kernel<<<g, spk::BLOCK, 0, st>>>(V, n, mm, d_S, k, W);
        SP_CK(cudaMemcpyAsync(V + (size_t)k * n, V + (size_t)mm * n, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));   // residual vector
        SP_CK(cudaMemcpyAsync(V, W, (size_t)k * n * 8, cudaMemcpyDeviceToDevice, st));
        std::fill(T.begin(), T.end(), 0.0);
        for (int e = 0; e < k; ++e) T[(size_t)e * m + e] = theta[mm - 1 - e];
        l = k;
    }
    *n_restart = restarts;
    *n_spmv = spmvs;
    return cudaStreamSynchronize(st);
}

}  // namespace mdsctk
