// sort.cuh -- block-level bitonic sort of (distance, index) pairs in shared memory, and a
// block-level radix select on 64-bit keys.  Final ordering of every k-list is
// (distance, index) ascending -- the documented tie order replacing libstdc++'s unspecified
// heap-select order in permutation<T>::sort (mdsctk.h:188-196).
#pragma once
#include <cuda_runtime.h>

namespace mdsctk {

__device__ __forceinline__ bool pair_less(double da, int ia, double db, int ib)
{
    return da < db || (da == db && ia < ib);
}

// n must be a power of two; all threads of the block participate.
__device__ inline void block_bitonic_sort(double *d, int *ix, int n)
{
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const bool up = (i & k) == 0;
                    const double di = d[i], dp = d[p];
                    const int ii = ix[i], ip = ix[p];
                    const bool swap = up ? pair_less(dp, ip, di, ii) : pair_less(di, ii, dp, ip);
                    if (swap) { d[i] = dp; d[p] = di; ix[i] = ip; ix[p] = ii; }
                }
            }
            __syncthreads();
        }
    }
}

// atomicMax on a non-negative double (its bit pattern orders like an unsigned integer)
__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v)
{
    atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}

__device__ __forceinline__ int next_pow2(int v)
{
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace mdsctk
