// csc.cu -- CSC matrices from the kNN lists on the GPU (the consumers make_sysparse / make_gesparse).
//
// Replaces make_sysparse.cpp:245-329: the reference inserts, row by row, the edge
// (min(i,j), max(i,j)) -> distance of the first k entries of every kNN row into an on-disk
// Berkeley-DB B-tree (Db::put overwrites: the LAST insertion of an edge wins), then walks the
// tree in (from, to) order and writes  int n; int pcol[n+1]; int irow[nnz]; double val[nnz].
// Here: HBM-bound integer/byte work, no tree --
//   1. count    one thread per (row, entry): candidates per column c = min(i,j)      (atomicAdd)
//   2. scan     exclusive prefix sum of the counts                                    (3-kernel scan)
//   3. scatter  (to, priority, distance) into the column's segment; priority encodes the
//               reference's insertion order: an edge written from its larger endpoint (the later
//               row) beats the one written from the smaller, later entries of a row beat earlier
//   4. resolve  one warp per column: sort the segment by (to, priority) (bitonic in shared
//               memory; global-memory fallback for hub columns), keep the last of every `to`
//   5. scan + compact into pcol / irow / val
// Algorithmic bytes per input entry: 12 read + 20 written + 20 read + <=12 written.
#include "common.cuh"

namespace mdsctk {

namespace csc {
constexpr int SCAN_BLOCK = 1024;
constexpr int SEG_CAP = 512;          // entries a warp sorts in shared memory
constexpr int WARPS = 4;
}  // namespace csc

// mode 0: make_sysparse (edge (min,max), larger endpoint's value wins)
// mode 1: make_gesparse (column = the listing row, every entry kept, last duplicate wins)
// mode 2: make_gesparse -s (additionally (j, i) <- d(i, j) where row j does not list i; first such entry wins)
__global__ void csc_count_kernel(const int *__restrict__ idx, long long n, int maxk, int k, int mode, int *__restrict__ cnt)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * k) return;
    const long long i = t / k;
    const int x = (int)(t - i * k);
    const int j = idx[i * maxk + x];
    if (mode == 0) {
        if (j == (int)i) return;
        const int c = j < (int)i ? j : (int)i;
        if (c >= 0 && c < n) atomicAdd(&cnt[c], 1);
    } else {
        atomicAdd(&cnt[i], 1);
        if (mode == 2 && j != (int)i && j >= 0 && j < n) atomicAdd(&cnt[j], 1);
    }
}

// ---- exclusive scan of int[n] (values and total fit in int: nnz < 2^31 as in the reference's int pcol) ----
__global__ void __launch_bounds__(csc::SCAN_BLOCK) scan_block_kernel(const int *__restrict__ in, int *__restrict__ out,
                                                                     int *__restrict__ block_sums, long long n)
{
    __shared__ int s_warp[32];
    const long long i = (long long)blockIdx.x * csc::SCAN_BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int v = i < n ? in[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += u;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    const int base = warp ? s_warp[warp - 1] : 0;
    if (i < n) out[i] = base + incl - v;
    if (threadIdx.x == csc::SCAN_BLOCK - 1 && block_sums) block_sums[blockIdx.x] = base + incl;
}

__global__ void scan_add_kernel(int *__restrict__ out, const int *__restrict__ block_offsets, long long n)
{
    const long long i = (long long)blockIdx.x * csc::SCAN_BLOCK + threadIdx.x;
    if (i < n) out[i] += block_offsets[blockIdx.x];
}

// out[i] = sum of in[0..i), out may alias in; tmp: two levels of block sums (ceil(n/1024) + ceil(n/1024^2) + 2 ints)
cudaError_t exclusive_scan(const int *in, int *out, long long n, int *tmp, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    const long long nb = (n + csc::SCAN_BLOCK - 1) / csc::SCAN_BLOCK;
    scan_block_kernel<<<(unsigned)nb, csc::SCAN_BLOCK, 0, st>>>(in, out, nb > 1 ? tmp : nullptr, n);
    if (nb > 1) {
        cudaError_t e = exclusive_scan(tmp, tmp, nb, tmp + nb, st);
        if (e != cudaSuccess) return e;
        scan_add_kernel<<<(unsigned)nb, csc::SCAN_BLOCK, 0, st>>>(out, tmp, n);
    }
    return cudaGetLastError();
}

// Segment entry: key = (row index `to`) << 32 | priority; per `to` the entry with the LARGEST priority is the one
// the reference's B-tree holds at the end (make_sysparse.cpp:245-277, make_gesparse.cpp:246-275).
__global__ void csc_scatter_kernel(const int *__restrict__ idx, const double *__restrict__ dist, long long n, int maxk,
                                   int k, int mode, const int *__restrict__ off, int *__restrict__ cur,
                                   unsigned long long *__restrict__ seg_key, double *__restrict__ seg_val)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * k) return;
    const long long i = t / k;
    const int x = (int)(t - i * k);
    const int j = idx[i * maxk + x];
    const double d = dist[i * maxk + x];
    auto emit = [&](int c, int r, unsigned prio) {
        const int pos = off[c] + atomicAdd(&cur[c], 1);
        seg_key[pos] = ((unsigned long long)(unsigned)r << 32) | prio;
        seg_val[pos] = d;
    };
    if (mode == 0) {
        if (j == (int)i) return;
        const bool rev = j < (int)i;             // written from the larger endpoint: processed later, overwrites
        const int c = rev ? j : (int)i, r = rev ? (int)i : j;
        if (c < 0 || c >= n) return;
        emit(c, r, (rev ? 0x80000000u : 0u) | (unsigned)x);          // later entries of a row overwrite earlier ones
    } else {
        emit((int)i, j, 0x80000000u | (unsigned)x);                  // direct put: always overwrites, last duplicate wins
        // symmetric fill: written only while (j, i) is absent, so the first attempt sticks unless row j lists i itself
        if (mode == 2 && j != (int)i && j >= 0 && j < n) emit(j, (int)i, (unsigned)(maxk - 1 - x));
    }
}

// Bitonic sort of n (any length) 64-bit keys + payload, ascending.  "Flip" formulation: every
// compare-exchange is ascending, so padding up to the next power of two can stay virtual (a partner
// index >= n stands for a maximal key that never moves).
template <typename Sync>
__device__ inline void bitonic_any(unsigned long long *key, double *val, int n, int tid, int nthr, Sync sync)
{
    int P = 1;
    while (P < n) P <<= 1;
    auto cmpswap = [&](int i, int p) {
        const unsigned long long a = key[i], b = key[p];
        if (b < a) {
            key[i] = b; key[p] = a;
            const double t = val[i]; val[i] = val[p]; val[p] = t;
        }
    };
    for (int kk = 2; kk <= P; kk <<= 1) {
        for (int i = tid; i < P; i += nthr) {
            const int p = i ^ (kk - 1);          // mirror within the block of kk
            if (p > i && p < n) cmpswap(i, p);
        }
        sync();
        for (int jj = kk >> 2; jj > 0; jj >>= 1) {
            for (int i = tid; i < P; i += nthr) {
                const int p = i ^ jj;
                if (p > i && p < n) cmpswap(i, p);
            }
            sync();
        }
    }
}

// One warp per column: order the segment by (to, priority); entry e survives if the next entry has a
// different `to`.  The survivors are compacted to the front of the segment; final[c] = their count.
__global__ void __launch_bounds__(csc::WARPS * 32) csc_resolve_kernel(const int *__restrict__ off, long long n,
                                                                     unsigned long long *__restrict__ seg_key,
                                                                     double *__restrict__ seg_val, int *__restrict__ fin)
{
    __shared__ unsigned long long s_key[csc::WARPS][csc::SEG_CAP];
    __shared__ double s_val[csc::WARPS][csc::SEG_CAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long c = (long long)blockIdx.x * csc::WARPS + warp;
    if (c >= n) return;
    const int b = off[c], len = off[c + 1] - b;
    if (len == 0) { if (lane == 0) fin[c] = 0; return; }
    unsigned long long *key = seg_key + b;
    double *val = seg_val + b;
    const bool in_smem = len <= csc::SEG_CAP;
    if (in_smem) {
        for (int i = lane; i < len; i += 32) { s_key[warp][i] = key[i]; s_val[warp][i] = val[i]; }
        __syncwarp();
        bitonic_any(s_key[warp], s_val[warp], len, lane, 32, [] { __syncwarp(); });
    } else {
        // hub column: same network on the global-memory segment (rare; __syncwarp orders the warp's accesses)
        bitonic_any(key, val, len, lane, 32, [] { __threadfence_block(); __syncwarp(); });
    }
    const unsigned long long *sk = in_smem ? s_key[warp] : key;
    const double *sv = in_smem ? s_val[warp] : val;
    int out = 0;
    for (int base = 0; base < len; base += 32) {
        const int i = base + lane;
        bool keep = false;
        unsigned long long kv = 0; double dv = 0.0;
        if (i < len) {
            kv = sk[i]; dv = sv[i];
            keep = (i + 1 == len) || ((sk[i + 1] >> 32) != (kv >> 32));
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);   // also orders the reads before the writes below
        if (keep) {
            const int pos = out + __popc(m & ((1u << lane) - 1u));
            key[pos] = kv;           // pos <= i: never overtakes an unread element of a later chunk
            val[pos] = dv;
        }
        out += __popc(m);
        __syncwarp();
    }
    if (lane == 0) fin[c] = out;
}

__global__ void csc_compact_kernel(const int *__restrict__ off, const int *__restrict__ pcol, long long n,
                                   const unsigned long long *__restrict__ seg_key, const double *__restrict__ seg_val,
                                   int *__restrict__ irow, double *__restrict__ val)
{
    const int lane = threadIdx.x & 31;
    const long long c = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= n) return;
    const int src = off[c], dst = pcol[c], len = pcol[c + 1] - dst;
    for (int i = lane; i < len; i += 32) {
        irow[dst + i] = (int)(seg_key[src + i] >> 32);
        val[dst + i] = seg_val[src + i];
    }
}

// Device-side driver.  d_idx [n][maxk], d_dist [n][maxk]; work buffers sized by the caller:
//   cnt, cur, off, fin, pcol: n + 1 ints each; scan_tmp: n/1024 + n/1024^2 + 8 ints;
//   seg_key / seg_val, irow / val: n * k entries (2 n k for mode 2), the upper bound of nnz.
// *nnz_host is valid after the stream has been synchronised by the caller's copy of pcol[n].
cudaError_t launch_csc_build(int mode, const int *d_idx, const double *d_dist, long long n, int maxk, int k, int *cnt, int *cur,
                             int *off, int *fin, int *pcol, int *scan_tmp, unsigned long long *seg_key, double *seg_val,
                             int *irow, double *val, cudaStream_t st)
{
    if (n <= 0 || k <= 0) return cudaSuccess;
    cudaError_t e;
    if ((e = cudaMemsetAsync(cnt, 0, (size_t)(n + 1) * 4, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(cur, 0, (size_t)(n + 1) * 4, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(fin, 0, (size_t)(n + 1) * 4, st)) != cudaSuccess) return e;
    const long long total = n * k;
    const unsigned grid = (unsigned)((total + 255) / 256);
    csc_count_kernel<<<grid, 256, 0, st>>>(d_idx, n, maxk, k, mode, cnt);
    if ((e = exclusive_scan(cnt, off, n + 1, scan_tmp, st)) != cudaSuccess) return e;
    csc_scatter_kernel<<<grid, 256, 0, st>>>(d_idx, d_dist, n, maxk, k, mode, off, cur, seg_key, seg_val);
    csc_resolve_kernel<<<(unsigned)((n + csc::WARPS - 1) / csc::WARPS), csc::WARPS * 32, 0, st>>>(off, n, seg_key, seg_val, fin);
    if ((e = exclusive_scan(fin, pcol, n + 1, scan_tmp, st)) != cudaSuccess) return e;
    csc_compact_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(off, pcol, n, seg_key, seg_val, irow, val);
    return cudaGetLastError();
}

}  // namespace mdsctk
