// make_gesparse -- B200 build of MDSCTK's make_gesparse tool.
//
// Same command line, stdout and output file as the reference tool (make_gesparse.cpp:36-367):
// reads indices.dat / distances.dat as written by knn_rms / knn_data (-k entries per row), keeps the
// first -n of them and writes the general (non-symmetric) CSC matrix
//     int n; int pcol[n+1]; int irow[nnz]; double val[nnz]          (make_gesparse.cpp:314-333)
// column i = the entries of kNN row i; with -s the transposed entry is added wherever it is missing
// (make_gesparse.cpp:259-275).  The reference goes through an on-disk Berkeley-DB B-tree plus `cat`
// of three temporary files; here the entries are resolved on the GPU (csc.cu).
#include "../../include/mdsctk_knn.h"
#include "options.hpp"

#include <fstream>
#include <iostream>
#include <vector>

using namespace mdsctk_cli;

int main(int argc, char *argv[])
{
    const char *program_name = "make_gesparse";
    banner(program_name);
    std::cout << "   Converts the results from knn_* into general CSC format." << std::endl << std::endl;
    std::cout << "   Normally, the number of nearest neighbors in the input" << std::endl;
    std::cout << "   distances is used for constructing the CSC matrix." << std::endl;
    std::cout << "   However, you can set output-knn <= knn in order to" << std::endl;
    std::cout << "   subselect the number of neighbors to consider in the" << std::endl;
    std::cout << "   CSC representation. This makes it easy to store a" << std::endl;
    std::cout << "   large number of neighbors using knn_* but then use" << std::endl;
    std::cout << "   a subset for, say, computing approximate geodesic" << std::endl;
    std::cout << "   distances." << std::endl << std::endl;
    std::cout << "   Use -h or --help to see the complete list of options." << std::endl << std::endl;

    Options po;
    po.add("help", 'h', Options::SWITCH, "show this help message and exit");
    po.add("knn", 'k', Options::VALUE, "Input:  K-nearest neighbors (int)");
    po.add("output-knn", 'n', Options::VALUE, "Input:  K-nn to keep in output (int)");
    po.add("symmetric", 's', Options::SWITCH, "Input:  Enforce symmetry (bool)");
    po.add("index-file", 'i', Options::VALUE, "Input:  Index file (string:filename)", "indices.dat", true);
    po.add("distance-file", 'd', Options::VALUE, "Input:  Distances file (string:filename)", "distances.dat", true);
    po.add("output-file", 'o', Options::VALUE, "Output: General sparse matrix file (string:filename)", "distances.gsm", true);

    int maxk = 0, k = 0;
    bool symmetric = false;
    try {
        po.parse(argc, argv);
        if (po.count("help")) {
            std::cout << "usage: " << program_name << " [options]" << std::endl;
            po.print(std::cout, "Program options");
            return 1;
        }
        if (po.count("knn")) maxk = po.integer("knn");
        symmetric = po.count("symmetric");
        k = po.count("output-knn") ? po.integer("output-knn") : maxk;       // make_gesparse.cpp:95-96
    } catch (const std::exception &e) {
        std::cout << "ERROR: " << e.what() << std::endl;
        return 2;
    }
    bool optsOK = true;
    if (!po.count("knn")) { std::cout << "ERROR: --knn not supplied." << std::endl << std::endl; optsOK = false; }
    if (maxk < k) {                                                          // make_gesparse.cpp:99-103
        std::cout << "ERROR: Output k (" << k << ") is not less than the input k (" << maxk << ")." << std::endl << std::endl;
        optsOK = false;
    }
    if (!optsOK) return -1;
    if (maxk <= 0 || k < 0) { std::cout << "ERROR: --knn must be positive." << std::endl; return -1; }
    const std::string i_filename = po.str("index-file"), d_filename = po.str("distance-file"), o_filename = po.str("output-file");

    std::cout << "Running with the following options:" << std::endl;
    std::cout << "knn =            " << maxk << std::endl;
    std::cout << "output-knn =     " << k << std::endl;
    std::cout << "index-file =     " << i_filename << std::endl;
    std::cout << "distances-file = " << d_filename << std::endl;
    std::cout << "output-file =    " << o_filename << std::endl;
    std::cout << std::endl;

    std::ifstream distances(d_filename.c_str(), std::ios::in | std::ios::binary | std::ios::ate);
    if (!distances.good()) {
        std::cout << "***ERROR***" << std::endl << "Could not open file: " << d_filename << std::endl << std::endl;
        return -1;
    }
    std::ifstream indices(i_filename.c_str(), std::ios::in | std::ios::binary | std::ios::ate);
    if (!indices.good()) {
        std::cout << "***ERROR***" << std::endl << "Could not open file: " << i_filename << std::endl << std::endl;
        return -1;
    }
    std::ofstream o_file(o_filename.c_str(), std::ios::out | std::ios::binary | std::ios::trunc);
    if (!o_file.good()) {
        std::cout << "***ERROR***" << std::endl << "Could not open file: " << o_filename << std::endl << std::endl;
        return -1;
    }
    // frames = complete rows present in BOTH files (the reference reads row by row until either ends)
    const long long n_i = (long long)indices.tellg() / (long long)(sizeof(int) * maxk);
    const long long n_d = (long long)distances.tellg() / (long long)(sizeof(double) * maxk);
    const long long n = n_i < n_d ? n_i : n_d;
    std::vector<int> idx((size_t)n * maxk);
    std::vector<double> dist((size_t)n * maxk);
    indices.seekg(0); distances.seekg(0);
    if (n > 0) {
        indices.read(reinterpret_cast<char *>(idx.data()), (std::streamsize)(sizeof(int) * idx.size()));
        distances.read(reinterpret_cast<char *>(dist.data()), (std::streamsize)(sizeof(double) * dist.size()));
    }

    std::cout << "Creating sparse matrix database..." << std::endl;
    std::vector<int> pcol((size_t)n + 1, 0), irow;
    std::vector<double> val;
    long long nnz = 0;
    if (n > 0) {
        mdsctk_knn_ctx *ctx = nullptr;
        if (mdsctk_knn_create(&ctx, 0) != 0) { std::cout << "ERROR: " << mdsctk_knn_last_error(nullptr) << std::endl; return 5; }
        if (mdsctk_knn_csc_build_general(ctx, idx.data(), dist.data(), n, maxk, k, symmetric ? 1 : 0, pcol.data(), &nnz) != 0) {
            std::cout << "ERROR: " << mdsctk_knn_last_error(ctx) << std::endl;
            return 5;
        }
        irow.resize((size_t)nnz); val.resize((size_t)nnz);
        if (mdsctk_knn_csc_fetch(ctx, irow.data(), val.data()) != 0) {
            std::cout << "ERROR: " << mdsctk_knn_last_error(ctx) << std::endl;
            return 5;
        }
        mdsctk_knn_destroy(ctx);
    }
    std::cout << std::endl << "Converting database to sparse matrix..." << std::endl;
    const int n32 = (int)n;
    o_file.write(reinterpret_cast<const char *>(&n32), sizeof(int));
    o_file.write(reinterpret_cast<const char *>(pcol.data()), (std::streamsize)(sizeof(int) * pcol.size()));
    o_file.write(reinterpret_cast<const char *>(irow.data()), (std::streamsize)(sizeof(int) * irow.size()));
    o_file.write(reinterpret_cast<const char *>(val.data()), (std::streamsize)(sizeof(double) * val.size()));
    o_file.close();
    if (!o_file.good()) std::cout << "Could not create general CSC matrix file: " << o_filename << std::endl;
    std::cout << std::endl << std::endl;
    return 0;
}
