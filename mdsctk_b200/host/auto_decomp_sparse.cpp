// auto_decomp_sparse -- B200 build of MDSCTK's auto_decomp_sparse tool.
//
// Same command line, stdout and output files as the reference tool (auto_decomp_sparse.cpp:36-263): distances of
// the symmetric CSC matrix -> Gaussian affinities with per-frame sigmas (-k) -> D^-1/2 W D^-1/2 -> the -n largest
// eigenpairs.  The affinity stage and the eigen-solve run on the GPU (spectral.cu; ARPACK is replaced by a
// thick-restart Lanczos, so eigenvector signs may differ from a given ARPACK build's).  The entropic-affinity
// option -K / --k-perplexity (mdsctk.cpp:388-520) is not available in this build.
#include "options.hpp"
#include "spectral_tool.hpp"

#include <cmath>
#include <cstdlib>

using namespace mdsctk_cli;

int main(int argc, char *argv[])
{
    const char *program_name = "auto_decomp_sparse";
    banner(program_name);
    std::cout << std::endl << std::endl;
    Options po;
    po.add("help", 'h', Options::SWITCH, "show this help message and exit");
    po.add("k-sigma", 'k', Options::VALUE, "Input:  K-nn to average for sigmas (int)");
    po.add("k-perplexity", 'K', Options::VALUE, "Input:  Desired perplexity within knn (real)");
    po.add("nevals", 'n', Options::VALUE, "Input:  Number of eigenvalues/vectors (int)");
    po.add("ssm-file", 's', Options::VALUE, "Input:  Symmetric sparse matrix file (string:filename)", "distances.ssm", true);
    po.add("evals-file", 'v', Options::VALUE, "Output:  Eigenvalues file (string:filename)", "eigenvalues.dat", true);
    po.add("evecs-file", 'e', Options::VALUE, "Output: Eigenvectors file (string:filename)", "eigenvectors.dat", true);
    po.add("residuals-file", 'r', Options::VALUE, "Output: Residuals file (string:filename)", "residuals.dat", true);
    int k_a = 0, nev = 0;
    try {
        po.parse(argc, argv);
        if (po.count("help")) {
            std::cout << "usage: " << program_name << " [options]" << std::endl;
            po.print(std::cout, "Program options");
            return 1;
        }
        if (po.count("k-sigma")) k_a = po.integer("k-sigma");
        if (po.count("nevals")) nev = po.integer("nevals");
    } catch (const std::exception &e) {
        std::cout << "ERROR: " << e.what() << std::endl;
        return 2;
    }
    bool optsOK = true;
    if (!po.count("k-sigma")) { std::cout << "ERROR: --k-sigma not supplied." << std::endl << std::endl; optsOK = false; }
    if (!po.count("nevals")) { std::cout << "ERROR: --nevals not supplied." << std::endl << std::endl; optsOK = false; }
    if (!optsOK) return -1;
    double K = 0.0;
    const bool pSet = po.count("k-perplexity") != 0;
    if (pSet) {
        K = std::atof(po.str("k-perplexity").c_str());
        if (std::ceil(K) > (double)k_a)                                   // auto_decomp_sparse.cpp:104-110
            std::cout << "WARNING: --k-perplexity (" << K << ") is higher than --k-sigma (" << k_a << ")" << std::endl;
    }
    if (k_a < 1) { std::cout << "ERROR: --k-sigma must be positive." << std::endl; return -1; }
    std::cout << "Running with the following options:" << std::endl;
    std::cout << "k-sigma =        " << k_a << std::endl;
    if (pSet) std::cout << "k-perplexity =   " << K << std::endl;
    std::cout << "nevals =         " << nev << std::endl;
    std::cout << "ssm-file =       " << po.str("ssm-file") << std::endl;
    std::cout << "residuals-file = " << po.str("residuals-file") << std::endl;
    std::cout << "evals-file =     " << po.str("evals-file") << std::endl;
    std::cout << "evecs-file =     " << po.str("evecs-file") << std::endl;
    std::cout << std::endl;
    return run_spectral_tool(po.str("ssm-file"), po.str("evals-file"), po.str("evecs-file"), po.str("residuals-file"), k_a, 0.0, nev, K);
}
