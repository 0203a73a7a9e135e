// angles_to_sincos -- B200 build of MDSCTK's angles_to_sincos tool.
//
// Same command line, stdout and output file as the reference tool (angles_to_sincos.cpp:38-128): every
// (binary double) angle a of the input file becomes the pair sin a, cos a in the output file.
#include "../../include/mdsctk_knn.h"
#include "options.hpp"

#include <fstream>
#include <iostream>
#include <vector>

using namespace mdsctk_cli;

int main(int argc, char *argv[])
{
    const char *program_name = "angles_to_sincos";
    banner(program_name);
    std::cout << "   Convert the (binary-double) angles from the input file" << std::endl;
    std::cout << "   to sin-cos euclidean coordinates and write the results" << std::endl;
    std::cout << "   to the provided output file." << std::endl << std::endl;
    std::cout << "   Use -h or --help to see the complete list of options." << std::endl << std::endl;

    Options po;
    po.add("help", 'h', Options::SWITCH, "show this help message and exit");
    po.add("input-file", 'i', Options::VALUE, "Input:  Phi-psi angle data file (string:filename)", "phipsi.dat", true);
    po.add("output-file", 'o', Options::VALUE, "Output: Projected angle data file (string:filename)", "sincos.dat", true);
    try {
        po.parse(argc, argv);
        if (po.count("help")) {
            std::cout << "usage: " << program_name << " [options]" << std::endl;
            po.print(std::cout, "Program options");
            return 1;
        }
    } catch (const std::exception &e) {
        std::cout << "ERROR: " << e.what() << std::endl;
        return 2;
    }
    const std::string input_filename = po.str("input-file"), output_filename = po.str("output-file");
    std::cout << "Running with the following options:" << std::endl;
    std::cout << "input-file  = " << input_filename << std::endl;
    std::cout << "output-file = " << output_filename << std::endl << std::endl;

    std::ifstream myin(input_filename.c_str(), std::ios::binary | std::ios::ate);
    if (!myin) { std::cout << "ERROR: cannot read " << input_filename << std::endl; return 3; }
    const long long input_length = (long long)myin.tellg();
    const long long n = input_length / 8;                       // trailing partial double dropped (:101-103,113-117)
    std::vector<double> data((size_t)n), result((size_t)n * 2);
    myin.seekg(0);
    if (n > 0) myin.read(reinterpret_cast<char *>(data.data()), (std::streamsize)(n * 8));
    if (n > 0) {
        mdsctk_knn_ctx *ctx = nullptr;
        if (mdsctk_knn_create(&ctx, 0) != 0) { std::cout << "ERROR: " << mdsctk_knn_last_error(nullptr) << std::endl; return 5; }
        if (mdsctk_knn_sincos(ctx, data.data(), n, result.data()) != 0) {
            std::cout << "ERROR: " << mdsctk_knn_last_error(ctx) << std::endl;
            return 5;
        }
        mdsctk_knn_destroy(ctx);
    }
    std::ofstream myout(output_filename.c_str(), std::ios::binary | std::ios::trunc);
    myout.write(reinterpret_cast<const char *>(result.data()), (std::streamsize)(result.size() * sizeof(double)));
    if (!myout) { std::cout << "ERROR: cannot write " << output_filename << std::endl; return 3; }
    std::cout << "Wrote " << (input_length / 8) << " sin-cos pairs (" << (input_length / 4) << " total values)." << std::endl;
    std::cout << std::endl;
    return 0;
}
