// decomp_sparse -- B200 build of MDSCTK's decomp_sparse tool.
//
// Same command line, stdout and output files as the reference tool (decomp_sparse.cpp:36-300): like
// auto_decomp_sparse with one global kernel width -q / --sigma (decomp_sparse.cpp:157-158).
#include "options.hpp"
#include "spectral_tool.hpp"

#include <cstdlib>

using namespace mdsctk_cli;

int main(int argc, char *argv[])
{
    const char *program_name = "decomp_sparse";
    banner(program_name);
    std::cout << std::endl << std::endl;
    Options po;
    po.add("help", 'h', Options::SWITCH, "show this help message and exit");
    po.add("sigma", 'q', Options::VALUE, "Input:  Standard deviation of gaussian kernel (real)");
    po.add("nevals", 'n', Options::VALUE, "Input:  Number of eigenvalues/vectors (int)");
    po.add("ssm-file", 's', Options::VALUE, "Input:  Symmetric sparse matrix file (string:filename)", "distances.ssm", true);
    po.add("evals-file", 'v', Options::VALUE, "Output:  Eigenvalues file (string:filename)", "eigenvalues.dat", true);
    po.add("evecs-file", 'e', Options::VALUE, "Output: Eigenvectors file (string:filename)", "eigenvectors.dat", true);
    po.add("residuals-file", 'r', Options::VALUE, "Output: Residuals file (string:filename)", "residuals.dat", true);
    double sigma = 0.0;
    int nev = 0;
    try {
        po.parse(argc, argv);
        if (po.count("help")) {
            std::cout << "usage: " << program_name << " [options]" << std::endl;
            po.print(std::cout, "Program options");
            return 1;
        }
        if (po.count("sigma")) {
            char *end = nullptr;
            sigma = std::strtod(po.str("sigma").c_str(), &end);
            if (po.str("sigma").empty() || *end) throw std::runtime_error("the argument ('" + po.str("sigma") + "') for option '--sigma' is invalid");
        }
        if (po.count("nevals")) nev = po.integer("nevals");
    } catch (const std::exception &e) {
        std::cout << "ERROR: " << e.what() << std::endl;
        return 2;
    }
    bool optsOK = true;
    if (!po.count("sigma")) { std::cout << "ERROR: --sigma not supplied." << std::endl << std::endl; optsOK = false; }
    if (!po.count("nevals")) { std::cout << "ERROR: --nevals not supplied." << std::endl << std::endl; optsOK = false; }
    if (!optsOK) return -1;
    if (!(sigma > 0.0)) { std::cout << "ERROR: --sigma must be positive." << std::endl; return -1; }
    std::cout << "Running with the following options:" << std::endl;
    std::cout << "sigma =          " << sigma << std::endl;
    std::cout << "nevals =         " << nev << std::endl;
    std::cout << "ssm-file =       " << po.str("ssm-file") << std::endl;
    std::cout << "evals-file =     " << po.str("evals-file") << std::endl;
    std::cout << "evecs-file =     " << po.str("evecs-file") << std::endl;
    std::cout << "residuals-file = " << po.str("residuals-file") << std::endl;
    std::cout << std::endl;
    return run_spectral_tool(po.str("ssm-file"), po.str("evals-file"), po.str("evecs-file"), po.str("residuals-file"), 0, sigma, nev);
}
