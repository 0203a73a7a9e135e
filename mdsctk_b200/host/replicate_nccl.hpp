// replicate_nccl.hpp -- north_star's multi-GPU load phase for the C++ tools: every GPU uploads and packs ITS shard of
// the reference set, then the packed arrays are replicated over NVLink / NVSwitch with NCCL; fit rows are sharded, the
// row loop itself needs no communication (each OpenMP iteration of knn_rms.cpp:268-279 owns its row).
//
// One process, one NCCL communicator per GPU (ncclCommInitAll), everything issued from the calling thread inside one
// ncclGroup: for every frame-major array the C ABI lists (mdsctk_knn_*_reference_arrays) and every rank, one
// ncclBroadcast of that rank's shard in place -- the all-gather of unequal shards.  Only the arrays the selected sweep
// kernel reads exist, so the default 1xFP16 run moves 5.5 KB per 300-atom frame.
#pragma once
#include "../../include/mdsctk_knn.h"

#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <functional>
#include <string>
#include <thread>
#include <vector>

namespace mdsctk_cli {

struct ShardRange { long long begin, count; };

inline ShardRange shard_of(long long n_total, int world, int rank)
{
    const long long per = (n_total + world - 1) / world;
    const long long b = std::min<long long>((long long)rank * per, n_total);
    return {b, std::min(per, n_total - b)};
}

// load_shard(g, range): uploads (and packs) rank g's rows on GPU g through the C ABI; list_arrays(g, ptrs, bytes_per_row):
// the device arrays to replicate.  Returns false with *err set on failure.
inline bool replicate_reference_nccl(int ngpus, long long n_total, const std::function<int(int, ShardRange)> &load_shard,
                                     const std::function<int(int, void **, size_t *, int *)> &list_arrays,
                                     const std::function<const char *(int)> &last_error, std::string *err, double *nccl_ms)
{
    std::vector<std::thread> pool;
    std::vector<int> status(ngpus, 0);
    for (int g = 0; g < ngpus; ++g)
        pool.emplace_back([&, g]() { status[g] = load_shard(g, shard_of(n_total, ngpus, g)); });
    for (auto &t : pool) t.join();
    for (int g = 0; g < ngpus; ++g)
        if (status[g] != 0) { *err = last_error(g); return false; }
    if (nccl_ms) *nccl_ms = 0.0;
    if (ngpus == 1) return true;

    std::vector<void *> ptrs((size_t)ngpus * 16, nullptr);
    std::vector<size_t> bpr((size_t)ngpus * 16, 0);
    int n_arrays = -1;
    for (int g = 0; g < ngpus; ++g) {
        int na = 0;
        if (list_arrays(g, &ptrs[(size_t)g * 16], &bpr[(size_t)g * 16], &na) != 0) { *err = last_error(g); return false; }
        if (n_arrays >= 0 && na != n_arrays) { *err = "GPUs disagree on the arrays to replicate"; return false; }
        n_arrays = na;
    }
    std::vector<int> devs(ngpus);
    for (int g = 0; g < ngpus; ++g) devs[g] = g;
    std::vector<ncclComm_t> comms(ngpus);
    ncclResult_t r = ncclCommInitAll(comms.data(), ngpus, devs.data());
    if (r != ncclSuccess) { *err = std::string("ncclCommInitAll: ") + ncclGetErrorString(r); return false; }
    std::vector<cudaStream_t> streams(ngpus);
    std::vector<cudaEvent_t> e0(ngpus), e1(ngpus);
    for (int g = 0; g < ngpus; ++g) {
        cudaSetDevice(g);
        cudaStreamCreateWithFlags(&streams[g], cudaStreamNonBlocking);
        cudaEventCreate(&e0[g]); cudaEventCreate(&e1[g]);
        cudaEventRecord(e0[g], streams[g]);
    }
    bool ok = true;
    ncclGroupStart();
    for (int g = 0; g < ngpus && ok; ++g)
        for (int a = 0; a < n_arrays && ok; ++a)
            for (int root = 0; root < ngpus && ok; ++root) {
                const ShardRange s = shard_of(n_total, ngpus, root);
                if (s.count <= 0) continue;
                char *p = static_cast<char *>(ptrs[(size_t)g * 16 + a]) + (size_t)s.begin * bpr[(size_t)g * 16 + a];
                r = ncclBroadcast(p, p, (size_t)s.count * bpr[(size_t)g * 16 + a], ncclChar, root, comms[g], streams[g]);
                if (r != ncclSuccess) { *err = std::string("ncclBroadcast: ") + ncclGetErrorString(r); ok = false; }
            }
    r = ncclGroupEnd();
    if (ok && r != ncclSuccess) { *err = std::string("ncclGroupEnd: ") + ncclGetErrorString(r); ok = false; }
    for (int g = 0; g < ngpus; ++g) {
        cudaSetDevice(g);
        cudaEventRecord(e1[g], streams[g]);
        if (cudaStreamSynchronize(streams[g]) != cudaSuccess && ok) { *err = "NCCL replication failed on a stream"; ok = false; }
        float ms = 0.f;
        if (ok && cudaEventElapsedTime(&ms, e0[g], e1[g]) == cudaSuccess && nccl_ms) *nccl_ms = std::max(*nccl_ms, (double)ms);
        cudaEventDestroy(e0[g]); cudaEventDestroy(e1[g]);
        cudaStreamDestroy(streams[g]);
    }
    for (int g = 0; g < ngpus; ++g) ncclCommDestroy(comms[g]);
    return ok;
}

}  // namespace mdsctk_cli
