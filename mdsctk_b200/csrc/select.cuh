// select.cuh -- streaming top-k: threshold-gated appends to per-row candidate lists and a
// warp-cooperative radix-select compaction.
//
// Replaces permutation<double>::sort(k1) (mdsctk.h:177-199: std::partial_sort of a fully
// materialised row of N_ref doubles).  Here a row is never materialised: a pair is appended
// to its row's list only if its key is below the row's admission threshold tau, and when a
// list is about to fill up one warp reduces it to the `keep` smallest entries and lowers
// tau to the largest key kept.  Invariant: every pair NOT in the list has key >= tau.
#pragma once
#include "common.cuh"

namespace mdsctk {

template <typename KeyT> struct KeyBits;
template <> struct KeyBits<float> {
    using U = uint32_t;
    static constexpr int kPasses = 4;
    __device__ static U to_bits(float k) { return __float_as_uint(k); }   // keys are >= +0
    __device__ static float from_bits(U u) { return __uint_as_float(u); }
    __device__ static float inf() { return __uint_as_float(0x7f800000u); }
};
template <> struct KeyBits<double> {
    using U = unsigned long long;
    static constexpr int kPasses = 8;
    __device__ static U to_bits(double k) { return (U)__double_as_longlong(k); }
    __device__ static double from_bits(U u) { return __longlong_as_double((long long)u); }
    __device__ static double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
};

// One warp.  keys/idxs: the row's list (global memory), cnt entries in use.  Keeps the
// `keep` smallest (ties on the boundary key: first come) compacted to [0, keep) and returns
// the largest kept key.  hist: 256 words of shared memory private to this warp.
template <typename KeyT>
__device__ KeyT warp_compact_list(KeyT *keys, int *idxs, int cnt, int keep, unsigned *hist)
{
    using KB = KeyBits<KeyT>;
    using U = typename KB::U;
    const int lane = threadIdx.x & 31;
    U prefix = 0, mask = 0;
    int remaining = keep;  // rank (1-based) of the wanted key among entries matching prefix
    for (int pass = 0; pass < KB::kPasses; ++pass) {
        const int shift = (KB::kPasses - 1 - pass) * 8;
        for (int b = lane; b < 256; b += 32) hist[b] = 0;
        __syncwarp();
        for (int i = lane; i < cnt; i += 32) {
            U k = KB::to_bits(keys[i]);
            if ((k & mask) == prefix) atomicAdd(&hist[(unsigned)(k >> shift) & 255u], 1u);
        }
        __syncwarp();
        // lane owns bins [8*lane, 8*lane+8)
        unsigned local[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { local[j] = hist[8 * lane + j]; sum += local[j]; }
        unsigned incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        unsigned excl = incl - sum;
        unsigned hit = __ballot_sync(0xffffffffu, incl >= (unsigned)remaining);
        int owner = __ffs(hit) - 1;  // first lane whose cumulative count reaches the rank
        int bin = 0, before = 0;
        if (lane == owner) {
            unsigned run = excl;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (run + local[j] >= (unsigned)remaining) { bin = 8 * lane + j; before = (int)run; break; }
                run += local[j];
            }
        }
        bin = __shfl_sync(0xffffffffu, bin, owner);
        before = __shfl_sync(0xffffffffu, before, owner);
        prefix |= (U)bin << shift;
        mask |= (U)255 << shift;
        remaining -= before;
        __syncwarp();
    }
    // prefix = keep-th smallest key; `remaining` of the entries equal to it are kept
    const U kth = prefix;
    int out = 0, eq_taken = 0;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int base = 0; base < cnt; base += 32) {
        const int i = base + lane;
        const bool valid = i < cnt;
        KeyT kv = valid ? keys[i] : KB::inf();
        int iv = valid ? idxs[i] : 0;
        U k = KB::to_bits(kv);
        const bool eq = valid && k == kth;
        const unsigned eqm = __ballot_sync(0xffffffffu, eq);
        const bool take = valid && (k < kth || (eq && eq_taken + __popc(eqm & lt_mask) < remaining));
        const unsigned tm = __ballot_sync(0xffffffffu, take);  // also orders the loads before the stores
        if (take) {
            const int pos = out + __popc(tm & lt_mask);
            keys[pos] = kv;
            idxs[pos] = iv;
        }
        out += __popc(tm);
        eq_taken += __popc(eqm);
        __syncwarp();
    }
    return KB::from_bits(kth);
}

// warp_compact_list with a 64-bin histogram (6-bit digits, 6 passes over the 32 key bits): a quarter of the shared memory
// of the 256-bin version in select.cuh -- every kilobyte of control state is a kilobyte less of operand ring.  Same contract.
__device__ inline float warp_compact_list6(float *keys, int *idxs, int cnt, int keep, unsigned *hist)
{
    const int lane = threadIdx.x & 31;
    uint32_t prefix = 0, mask = 0;
    int remaining = keep;
    for (int pass = 0; pass < 6; ++pass) {
        const int shift = pass < 5 ? 26 - 6 * pass : 0;
        const uint32_t dmask = pass < 5 ? 63u : 3u;
        hist[lane] = 0; hist[lane + 32] = 0;
        __syncwarp();
        for (int i = lane; i < cnt; i += 32) {
            const uint32_t k = __float_as_uint(keys[i]);
            if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & dmask], 1u);
        }
        __syncwarp();
        unsigned local[2], sum = 0;               // lane owns bins 2 lane, 2 lane + 1
#pragma unroll
        for (int j = 0; j < 2; ++j) { local[j] = hist[2 * lane + j]; sum += local[j]; }
        unsigned incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const unsigned excl = incl - sum;
        const unsigned hit = __ballot_sync(0xffffffffu, incl >= (unsigned)remaining);
        const int owner = __ffs(hit) - 1;
        int bin = 0, before = 0;
        if (lane == owner) {
            unsigned run = excl;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (run + local[j] >= (unsigned)remaining) { bin = 2 * lane + j; before = (int)run; break; }
                run += local[j];
            }
        }
        bin = __shfl_sync(0xffffffffu, bin, owner);
        before = __shfl_sync(0xffffffffu, before, owner);
        prefix |= (uint32_t)bin << shift;
        mask |= dmask << shift;
        remaining -= before;
        __syncwarp();
    }
    const uint32_t kth = prefix;
    int out = 0, eq_taken = 0;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int base = 0; base < cnt; base += 32) {
        const int i = base + lane;
        const bool valid = i < cnt;
        const float kv = valid ? keys[i] : __uint_as_float(0x7f800000u);
        const int iv = valid ? idxs[i] : 0;
        const uint32_t k = __float_as_uint(kv);
        const bool eq = valid && k == kth;
        const unsigned eqm = __ballot_sync(0xffffffffu, eq);
        const bool take = valid && (k < kth || (eq && eq_taken + __popc(eqm & lt_mask) < remaining));
        const unsigned tm = __ballot_sync(0xffffffffu, take);
        if (take) {
            const int pos = out + __popc(tm & lt_mask);
            keys[pos] = kv;
            idxs[pos] = iv;
        }
        out += __popc(tm);
        eq_taken += __popc(eqm);
        __syncwarp();
    }
    return __uint_as_float(kth);
}


}  // namespace mdsctk
