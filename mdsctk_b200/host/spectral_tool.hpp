// spectral_tool.hpp -- shared body of the auto_decomp_sparse / decomp_sparse tools: read the symmetric CSC
// matrix (CSC_matrix, mdsctk.cpp:44-59), run the spectral stage of the library, write the three text files
// exactly as the reference's default (non-DECOMP_WRITE_DOUBLE) build does: operator<< of a double, i.e. six
// significant digits; eigenvalues largest first; one eigenvector per line, every value followed by a blank
// (auto_decomp_sparse.cpp:207-236, decomp_sparse.cpp:228-262).
#pragma once
#include "../../include/mdsctk_knn.h"

#include <cmath>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

namespace mdsctk_cli {

inline double sqrt_machine_eps()          // getEPS(), mdsctk.cpp:285-290
{
    double eps = 1.0;
    do { eps /= 2.0; } while (1.0 + (eps / 2.0) != 1.0);
    return std::sqrt(eps);
}

inline int run_spectral_tool(const std::string &ssm_filename, const std::string &evals_filename, const std::string &evecs_filename,
                             const std::string &residuals_filename, int k_sigma, double sigma, int nev, double k_perplexity = 0.0)
{
    std::ifstream ssm(ssm_filename.c_str(), std::ios::binary);
    if (!ssm) { std::cout << "ERROR: cannot read " << ssm_filename << std::endl; return 3; }
    int n = 0;
    ssm.read(reinterpret_cast<char *>(&n), sizeof(int));
    if (!ssm || n < 2) { std::cout << "ERROR: " << ssm_filename << " is not a symmetric sparse matrix file" << std::endl; return 3; }
    std::vector<int> pcol((size_t)n + 1);
    ssm.read(reinterpret_cast<char *>(pcol.data()), (std::streamsize)(sizeof(int) * pcol.size()));
    const int nnz = ssm ? pcol[(size_t)n] : -1;
    if (nnz < 0) { std::cout << "ERROR: " << ssm_filename << " is truncated" << std::endl; return 3; }
    std::vector<int> irow((size_t)nnz);
    std::vector<double> val((size_t)nnz);
    ssm.read(reinterpret_cast<char *>(irow.data()), (std::streamsize)(sizeof(int) * irow.size()));
    ssm.read(reinterpret_cast<char *>(val.data()), (std::streamsize)(sizeof(double) * val.size()));
    if (!ssm) { std::cout << "ERROR: " << ssm_filename << " is truncated" << std::endl; return 3; }
    if (nev < 1 || nev >= n) { std::cout << "ERROR: --nevals must be between 1 and n-1" << std::endl; return -1; }

    std::ofstream eigenvalues(evals_filename.c_str()), eigenvectors(evecs_filename.c_str()), residuals(residuals_filename.c_str());
    if (!eigenvalues || !eigenvectors || !residuals) { std::cout << "ERROR: cannot open the output files" << std::endl; return 3; }

    mdsctk_knn_ctx *ctx = nullptr;
    if (mdsctk_knn_create(&ctx, 0) != 0) { std::cout << "ERROR: " << mdsctk_knn_last_error(nullptr) << std::endl; return 5; }
    std::vector<double> d((size_t)nev), Z((size_t)nev * n), res((size_t)nev);
    double avg_sigma = 0.0;
    int nconv = 0;
    if (mdsctk_knn_spectral_decomp_ex(ctx, n, pcol.data(), irow.data(), val.data(), k_sigma, sigma, k_perplexity, nev, d.data(), Z.data(),
                                      res.data(), &avg_sigma, &nconv, nullptr) != 0) {
        std::cout << "ERROR: " << mdsctk_knn_last_error(ctx) << std::endl;
        return 5;
    }
    mdsctk_knn_destroy(ctx);
    if (k_sigma > 0) std::cout << "Average sigma: " << avg_sigma << std::endl << std::endl;      // auto_decomp_sparse.cpp:199-203
    std::cout << "Number of converged eigenvalues/vectors found: " << nconv << std::endl;
    double max_residual = 0.0;
    for (int x = 0; x < nev; ++x) {                    // already largest first (the reference walks ARPACK's ascending order backwards)
        eigenvalues << d[(size_t)x] << std::endl;
        for (int y = 0; y < n; ++y) eigenvectors << Z[(size_t)x * n + y] << " ";
        eigenvectors << std::endl;
        residuals << res[(size_t)x] << std::endl;
        if (res[(size_t)x] > max_residual) max_residual = res[(size_t)x];
    }
    const double eps = sqrt_machine_eps();
    std::cout << "Maximum residual: " << max_residual << " (eps: " << eps << ")" << std::endl;
    if (max_residual > eps) {
        std::cout << "*** Max residual too high (max_r > eps)!" << std::endl;
        std::cout << "*** Please, check results manually..." << std::endl;
    }
    std::cout << std::endl;
    return 0;
}

}  // namespace mdsctk_cli
