// knn_data -- B200 build of MDSCTK's knn_data tool.
//
// Same command line, stdout and output files as the reference tool (knn_data.cpp:43-263): the k
// nearest reference vectors of every fitting vector by Euclidean (or, with -c, correlation)
// distance; input files are headerless row-major doubles, vector_size per row, a trailing partial
// row is dropped (knn_data.cpp:141-168).  Distances are bit-identical to the CPU tool's
// (mdsctk.cpp:330-360 arithmetic kept operation for operation in FP64 on the GPU).
#include "../../include/mdsctk_knn.h"
#include "options.hpp"
#include "replicate_nccl.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <fstream>
#include <iostream>
#include <thread>
#include <vector>

using namespace mdsctk_cli;

namespace {

bool read_rows(const std::string &path, int dim, std::vector<double> *rows, long long *n)
{
    std::ifstream f(path.c_str(), std::ios::binary | std::ios::ate);
    if (!f) return false;
    const std::streamsize bytes = f.tellg();
    f.seekg(0);
    *n = (long long)(bytes / (std::streamsize)(sizeof(double) * dim));   // partial last row dropped
    rows->resize((size_t)*n * dim);
    if (*n > 0) f.read(reinterpret_cast<char *>(rows->data()), (std::streamsize)(*n * dim * sizeof(double)));
    return true;
}

}  // namespace

int main(int argc, char *argv[])
{
    const char *program_name = "knn_data";
    banner(program_name);
    std::cout << "   Computes the k nearest neighbors of all pairs of" << std::endl;
    std::cout << "   vectors in the given binary data files." << std::endl << std::endl;
    std::cout << "   Use -h or --help to see the complete list of options." << std::endl << std::endl;

    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    Options po;
    po.add("help", 'h', Options::SWITCH, "show this help message and exit");
    po.add("threads", 't', Options::VALUE, "Input:  Number of threads to start (int, ignored on GPU)", std::to_string(hw), true);
    po.add("knn", 'k', Options::VALUE, "Input:  K-nearest neighbors (int)");
    po.add("sort", 's', Options::VALUE, "Input:  Find K-nn,false=full distance matix (bool)", "1", true);
    po.add("vector-size", 'v', Options::VALUE, "Input:  Data vector length (int)");
    po.add("block-size", 'b', Options::VALUE, "Input:  Workgroup block size in # frames (int, ignored on GPU)", "128", true);
    po.add("correlation", 'c', Options::SWITCH, "Input:  Use correlation distance (bool)");
    po.add("reference-file", 'r', Options::VALUE, "Input:  Reference data file (string:filename)", "reference.pts", true);
    po.add("fit-file", 'f', Options::VALUE, "Input:  Fitting data file (string:filename)");
    po.add("distance-file", 'd', Options::VALUE, "Output: K-nn distances file (string:filename)", "distances.dat", true);
    po.add("index-file", 'i', Options::VALUE, "Output: K-nn indices file (string:filename)", "indices.dat", true);
    po.add("gpus", 'g', Options::VALUE, "Input:  Number of GPUs; fit rows are sharded across them (int)", "1", true);

    int nthreads, k = 0, vector_size = 0, blksize, ngpus;
    bool sort, corr;
    try {
        po.parse(argc, argv);
        if (po.count("help")) {
            std::cout << "usage: " << program_name << " [options]" << std::endl;
            po.print(std::cout, "Program options");
            return 1;
        }
        nthreads = po.integer("threads");
        if (po.count("knn")) k = po.integer("knn");
        if (po.count("vector-size")) vector_size = po.integer("vector-size");
        sort = po.boolean("sort");
        corr = po.count("correlation");
        blksize = po.integer("block-size");
        ngpus = po.integer("gpus");
    } catch (const std::exception &e) {
        std::cout << "ERROR: " << e.what() << std::endl;
        return 2;
    }
    bool optsOK = true;
    if (!po.count("knn") && sort) { std::cout << "ERROR: --knn not supplied." << std::endl << std::endl; optsOK = false; }
    if (!po.count("vector-size")) { std::cout << "ERROR: --vector-size not supplied." << std::endl << std::endl; optsOK = false; }
    if (!optsOK) return -1;
    if (vector_size <= 0) { std::cout << "ERROR: --vector-size must be positive." << std::endl; return -1; }
    const std::string ref_filename = po.str("reference-file");
    const std::string fit_filename = po.count("fit-file") ? po.str("fit-file") : ref_filename;
    const std::string d_filename = po.str("distance-file"), i_filename = po.str("index-file");

    std::cout << "Running with the following options:" << std::endl;
    std::cout << "threads =        " << nthreads << std::endl;
    std::cout << "knn =            " << k << std::endl;
    std::cout << "sort =           " << sort << std::endl;
    std::cout << "vector-size =    " << vector_size << std::endl;
    std::cout << "reference-file = " << ref_filename << std::endl;
    std::cout << "fit-file =       " << fit_filename << std::endl;
    std::cout << "distance-file =  " << d_filename << std::endl;
    std::cout << "index-file =     " << i_filename << std::endl;
    std::cout << std::endl;

    std::vector<double> ref, fit;
    long long n_ref = 0, n_fit = 0;
    std::cout << "Reading reference coordinates from file: " << ref_filename << " ... ";
    if (!read_rows(ref_filename, vector_size, &ref, &n_ref)) { std::cout << std::endl << "ERROR: cannot read " << ref_filename << std::endl; return 3; }
    std::cout << "done." << std::endl;
    std::cout << "Number of reference coordinates: " << n_ref << std::endl;
    const bool same = fit_filename == ref_filename;
    std::cout << "Reading fitting coordinates from file: " << fit_filename << " ... ";
    if (same) n_fit = n_ref;
    else if (!read_rows(fit_filename, vector_size, &fit, &n_fit)) { std::cout << std::endl << "ERROR: cannot read " << fit_filename << std::endl; return 3; }
    std::cout << "done." << std::endl;
    std::cout << "Number of fitting coordinates: " << n_fit << std::endl;
    if (n_ref <= 0 || n_fit <= 0) { std::cout << "ERROR: empty input" << std::endl; return 3; }

    std::ofstream distances(d_filename.c_str(), std::ios::binary | std::ios::trunc);
    std::ofstream indices(i_filename.c_str(), std::ios::binary | std::ios::trunc);
    if (!distances || !indices) { std::cout << "ERROR: cannot open the output files" << std::endl; return 3; }
    if ((long long)blksize > n_fit) blksize = (int)n_fit;
    std::cout << "Block size: " << blksize << std::endl;

    if (n_ref - 1 < k) k = (int)(n_ref - 1);   // knn_data.cpp:186-188
    const int k1 = k + 1;

    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    ngpus = std::max(1, std::min(ngpus, std::max(ndev, 1)));
    std::vector<mdsctk_knn_ctx *> ctx(ngpus, nullptr);
    for (int g = 0; g < ngpus; ++g)
        if (mdsctk_knn_create(&ctx[g], g) != 0) { std::cout << "ERROR: " << mdsctk_knn_last_error(nullptr) << std::endl; return 5; }
    {
        // every GPU uploads its own shard of the reference rows, NCCL replicates them (replicate_nccl.hpp)
        std::string rerr;
        double nccl_ms = 0.0;
        const bool ok = replicate_reference_nccl(
            ngpus, n_ref,
            [&](int g, ShardRange s) {
                int rc = mdsctk_knn_data_alloc_reference(ctx[g], n_ref, vector_size);
                if (rc == 0 && s.count > 0) rc = mdsctk_knn_data_upload_shard(ctx[g], ref.data() + (size_t)s.begin * vector_size, s.begin, s.count);
                return rc;
            },
            [&](int g, void **ptrs, size_t *bpr, int *na) { return mdsctk_knn_data_reference_arrays(ctx[g], 16, na, ptrs, bpr); },
            [&](int g) { return mdsctk_knn_last_error(ctx[g]); }, &rerr, &nccl_ms);
        if (!ok) { std::cout << "ERROR: " << rerr << std::endl; return 5; }
        if (ngpus > 1) std::cout << "Reference rows uploaded on " << ngpus << " GPUs and replicated with NCCL in " << nccl_ms << " ms." << std::endl;
    }
    const int metric = corr ? MDSCTK_KNN_CORRELATION : MDSCTK_KNN_EUCLIDEAN;
    if (!sort) {
        // full rows in reference order (knn_data.cpp:198-216).  The reference writes its never-filled index
        // vector here (undefined behaviour); this build writes the identity permutation, like knn_rms.
        const long long rows = std::max<long long>(1, std::min<long long>(n_fit, (64LL << 20) / (n_ref * 8)));
        std::vector<double> buf((size_t)rows * n_ref);
        std::vector<int> ident((size_t)n_ref);
        for (long long j = 0; j < n_ref; ++j) ident[(size_t)j] = (int)j;
        int rc = 0;
        for (long long f = 0; f < n_fit && rc == 0; f += rows) {
            const long long n = std::min(rows, n_fit - f);
            const double *fr = (same ? ref.data() : fit.data()) + (size_t)f * vector_size;
            if (mdsctk_knn_data_rows(ctx[0], fr, n, metric, buf.data()) != 0) {
                std::cout << "ERROR: " << mdsctk_knn_last_error(ctx[0]) << std::endl;
                rc = 5;
                break;
            }
            for (long long r = 0; r < n; ++r) {
                distances.write(reinterpret_cast<const char *>(&buf[(size_t)r * n_ref]), (std::streamsize)(sizeof(double) * n_ref));
                indices.write(reinterpret_cast<const char *>(ident.data()), (std::streamsize)(sizeof(int) * n_ref));
            }
        }
        std::cout << std::endl << std::endl;
        for (auto *c : ctx) mdsctk_knn_destroy(c);
        return rc;
    }
    std::vector<double> dist((size_t)n_fit * k1);
    std::vector<int> idx((size_t)n_fit * k1);
    std::vector<std::thread> pool;
    std::vector<int> status(ngpus, 0);
    const long long shard = (n_fit + ngpus - 1) / ngpus;
    for (int g = 0; g < ngpus; ++g) {
        pool.emplace_back([&, g]() {
            const long long b = std::min(n_fit, g * shard), n = std::min(shard, n_fit - b);
            if (n <= 0) return;
            double *od = dist.data() + (size_t)b * k1;
            int *oi = idx.data() + (size_t)b * k1;
            status[g] = same ? mdsctk_knn_data_query_range(ctx[g], b, n, k1, metric, od, oi)
                             : mdsctk_knn_data_query(ctx[g], fit.data() + (size_t)b * vector_size, n, k1, metric, od, oi);
        });
    }
    for (auto &t : pool) t.join();
    int rc = 0;
    for (int g = 0; g < ngpus; ++g)
        if (status[g] != 0) { std::cout << "ERROR: " << mdsctk_knn_last_error(ctx[g]) << std::endl; rc = 5; }
    if (rc == 0) {
        for (long long f = 0; f < n_fit; ++f) {   // sorted position 0 is dropped (knn_data.cpp:240-249)
            distances.write(reinterpret_cast<const char *>(&dist[(size_t)f * k1 + 1]), sizeof(double) * k);
            indices.write(reinterpret_cast<const char *>(&idx[(size_t)f * k1 + 1]), sizeof(int) * k);
        }
    }
    std::cout << std::endl << std::endl;
    for (auto *c : ctx) mdsctk_knn_destroy(c);
    return rc;
}
