// knn_data_sparse -- B200 build of MDSCTK's knn_data_sparse tool.
//
// Same command line, stdout and output files as the reference tool (knn_data_sparse.cpp:38-275): the k
// nearest reference vectors of every fitting vector, the vectors given in sparse form (index file: per
// vector `int n; int index[n]`, data file: `double value[n]`; e.g. contact profiles).  The reference
// merges two ascending index lists per pair (euclidean_distance_sparse, mdsctk.cpp:362-386) and adds the
// terms in ascending index order; a dimension present in one vector only contributes its square.  The
// same sum, bit for bit, is the dense Euclidean distance over the UNION of the indices in ascending order
// (absent entries are 0: (r-0)^2 = r*r, adding +0.0 changes nothing), so the vectors are densified over
// the compacted union of indices and handed to the knn_data path of the library (tensor-core filter +
// exact FP64 re-score, or the exact FP64 sweep).
#include "../../include/mdsctk_knn.h"
#include "options.hpp"

#include <algorithm>
#include <fstream>
#include <iostream>
#include <thread>
#include <vector>

using namespace mdsctk_cli;

namespace {

struct SparseSet {
    std::vector<long long> off{0};
    std::vector<int> idx;
    std::vector<double> val;
    long long n() const { return (long long)off.size() - 1; }
};

// knn_data_sparse.cpp:147-160: records until the index file ends
bool read_sparse(const std::string &index_path, const std::string &data_path, SparseSet *s)
{
    std::ifstream index(index_path.c_str(), std::ios::binary), data(data_path.c_str(), std::ios::binary);
    if (!index || !data) return false;
    int n = 0;
    while (index.read(reinterpret_cast<char *>(&n), sizeof(int))) {
        if (n < 0) return false;
        const size_t at = s->idx.size();
        s->idx.resize(at + (size_t)n);
        s->val.resize(at + (size_t)n);
        if (n > 0) {
            index.read(reinterpret_cast<char *>(&s->idx[at]), (std::streamsize)(sizeof(int) * n));
            data.read(reinterpret_cast<char *>(&s->val[at]), (std::streamsize)(sizeof(double) * n));
            if (!index || !data) return false;
        }
        s->off.push_back((long long)s->idx.size());
    }
    return true;
}

void densify(const SparseSet &s, const std::vector<int> &dims, std::vector<double> *rows)
{
    const size_t D = dims.size();
    rows->assign((size_t)s.n() * D, 0.0);
    for (long long v = 0; v < s.n(); ++v)
        for (long long e = s.off[(size_t)v]; e < s.off[(size_t)v + 1]; ++e) {
            const size_t col = (size_t)(std::lower_bound(dims.begin(), dims.end(), s.idx[(size_t)e]) - dims.begin());
            (*rows)[(size_t)v * D + col] = s.val[(size_t)e];
        }
}

}  // namespace

int main(int argc, char *argv[])
{
    const char *program_name = "knn_data_sparse";
    banner(program_name);
    std::cout << "   Computes the k nearest neighbors of all pairs of" << std::endl;
    std::cout << "   vectors in the given sparse binary data files." << std::endl << std::endl;
    std::cout << "   Use -h or --help to see the complete list of options." << std::endl << std::endl;

    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    Options po;
    po.add("help", 'h', Options::SWITCH, "show this help message and exit");
    po.add("threads", 't', Options::VALUE, "Input:  Number of threads to start (int, ignored on GPU)", std::to_string(hw), true);
    po.add("knn", 'k', Options::VALUE, "Input:  K-nearest neighbors (int)");
    po.add("sort", 's', Options::VALUE, "Input:  Find K-nn,false=full distance matix (bool)", "1", true);
    po.add("block-size", 'b', Options::VALUE, "Input:  Workgroup block size in # frames (int, ignored on GPU)", "128", true);
    po.add("reference-index-file", 'R', Options::VALUE, "Input:  Reference index file (string:filename)", "reference.svi", true);
    po.add("reference-data-file", 'r', Options::VALUE, "Input:  Reference data file (string:filename)", "reference.svd", true);
    po.add("fit-index-file", 'F', Options::VALUE, "Input:  Fitting index file (string:filename)");
    po.add("fit-data-file", 'f', Options::VALUE, "Input:  Fitting data file (string:filename)");
    po.add("distance-file", 'd', Options::VALUE, "Output: K-nn distances file (string:filename)", "distances.dat", true);
    po.add("index-file", 'i', Options::VALUE, "Output: K-nn indices file (string:filename)", "indices.dat", true);

    int nthreads, k = 0, blksize;
    bool sort;
    try {
        po.parse(argc, argv);
        if (po.count("help")) {
            std::cout << "usage: " << program_name << " [options]" << std::endl;
            po.print(std::cout, "Program options");
            return 1;
        }
        nthreads = po.integer("threads");
        if (po.count("knn")) k = po.integer("knn");
        sort = po.boolean("sort");
        blksize = po.integer("block-size");
    } catch (const std::exception &e) {
        std::cout << "ERROR: " << e.what() << std::endl;
        return 2;
    }
    if (!po.count("knn") && sort) { std::cout << "ERROR: --knn not supplied." << std::endl << std::endl; return -1; }
    const std::string ref_index_filename = po.str("reference-index-file"), ref_data_filename = po.str("reference-data-file");
    const std::string fit_index_filename = po.count("fit-index-file") ? po.str("fit-index-file") : ref_index_filename;
    const std::string fit_data_filename = po.count("fit-data-file") ? po.str("fit-data-file") : ref_data_filename;
    const std::string d_filename = po.str("distance-file"), i_filename = po.str("index-file");

    std::cout << "Running with the following options:" << std::endl;
    std::cout << "threads =              " << nthreads << std::endl;
    std::cout << "knn =                  " << k << std::endl;
    std::cout << "sort =                 " << sort << std::endl;
    std::cout << "reference-index-file = " << ref_index_filename << std::endl;
    std::cout << "reference-file =       " << ref_data_filename << std::endl;
    std::cout << "fit-index-file =       " << fit_index_filename << std::endl;
    std::cout << "fit-file =             " << fit_data_filename << std::endl;
    std::cout << "distance-file =        " << d_filename << std::endl;
    std::cout << "index-file =           " << i_filename << std::endl;
    std::cout << std::endl;

    SparseSet ref, fit_own;
    std::cout << "Reading reference coordinates from files...";
    if (!read_sparse(ref_index_filename, ref_data_filename, &ref)) {
        std::cout << std::endl << "ERROR: cannot read " << ref_index_filename << " / " << ref_data_filename << std::endl;
        return 3;
    }
    std::cout << "done." << std::endl;
    const bool same = fit_index_filename == ref_index_filename && fit_data_filename == ref_data_filename;
    std::cout << "Reading fitting coordinates from files...";
    if (!same && !read_sparse(fit_index_filename, fit_data_filename, &fit_own)) {
        std::cout << std::endl << "ERROR: cannot read " << fit_index_filename << " / " << fit_data_filename << std::endl;
        return 3;
    }
    std::cout << "done." << std::endl;
    const SparseSet &fit = same ? ref : fit_own;
    const long long n_ref = ref.n(), n_fit = fit.n();
    if (n_ref <= 0 || n_fit <= 0) { std::cout << "ERROR: empty input" << std::endl; return 3; }

    std::ofstream distances(d_filename.c_str(), std::ios::binary | std::ios::trunc);
    std::ofstream indices(i_filename.c_str(), std::ios::binary | std::ios::trunc);
    if (!distances || !indices) { std::cout << "ERROR: cannot open the output files" << std::endl; return 3; }
    if ((long long)blksize > n_fit) blksize = (int)n_fit;
    std::cout << "Block size: " << blksize << std::endl;
    if (n_ref - 1 < k) k = (int)(n_ref - 1);      // knn_data_sparse.cpp:186-188
    const int k1 = k + 1;

    // compacted union of the indices, ascending: the dense dimension order = the merge order of the reference
    std::vector<int> dims(ref.idx);
    if (!same) dims.insert(dims.end(), fit.idx.begin(), fit.idx.end());
    std::sort(dims.begin(), dims.end());
    dims.erase(std::unique(dims.begin(), dims.end()), dims.end());
    if (dims.empty()) dims.push_back(0);
    const int D = (int)dims.size();
    std::vector<double> ref_rows, fit_rows;
    densify(ref, dims, &ref_rows);
    if (!same) densify(fit, dims, &fit_rows);

    mdsctk_knn_ctx *ctx = nullptr;
    if (mdsctk_knn_create(&ctx, 0) != 0) { std::cout << "ERROR: " << mdsctk_knn_last_error(nullptr) << std::endl; return 5; }
    int rc = 0;
    if (mdsctk_knn_data_set_reference(ctx, ref_rows.data(), n_ref, D) != 0) rc = 5;
    if (rc == 0 && sort) {
        std::vector<double> dist((size_t)n_fit * k1);
        std::vector<int> idx((size_t)n_fit * k1);
        if (mdsctk_knn_data_query(ctx, same ? nullptr : fit_rows.data(), n_fit, k1, MDSCTK_KNN_EUCLIDEAN, dist.data(), idx.data()) != 0) rc = 5;
        for (long long f = 0; f < n_fit && rc == 0; ++f) {   // sorted position 0 is dropped (knn_data_sparse.cpp:207-216)
            distances.write(reinterpret_cast<const char *>(&dist[(size_t)f * k1 + 1]), sizeof(double) * k);
            indices.write(reinterpret_cast<const char *>(&idx[(size_t)f * k1 + 1]), sizeof(int) * k);
        }
    } else if (rc == 0) {
        const long long rows = std::max<long long>(1, std::min<long long>(n_fit, (64LL << 20) / (n_ref * 8)));
        std::vector<double> buf((size_t)rows * n_ref);
        std::vector<int> ident((size_t)n_ref);
        for (long long j = 0; j < n_ref; ++j) ident[(size_t)j] = (int)j;
        for (long long f = 0; f < n_fit && rc == 0; f += rows) {
            const long long n = std::min(rows, n_fit - f);
            const double *fr = (same ? ref_rows.data() : fit_rows.data()) + (size_t)f * D;
            if (mdsctk_knn_data_rows(ctx, fr, n, MDSCTK_KNN_EUCLIDEAN, buf.data()) != 0) { rc = 5; break; }
            for (long long r = 0; r < n; ++r) {
                distances.write(reinterpret_cast<const char *>(&buf[(size_t)r * n_ref]), (std::streamsize)(sizeof(double) * n_ref));
                indices.write(reinterpret_cast<const char *>(ident.data()), (std::streamsize)(sizeof(int) * n_ref));
            }
        }
    }
    if (rc != 0) std::cout << "ERROR: " << mdsctk_knn_last_error(ctx) << std::endl;
    std::cout << std::endl << std::endl;
    mdsctk_knn_destroy(ctx);
    return rc;
}
