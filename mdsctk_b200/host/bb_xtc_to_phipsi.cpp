// bb_xtc_to_phipsi -- B200 build of MDSCTK's bb_xtc_to_phipsi tool.
//
// Same command line, stdout and output file as the reference tool (bb_xtc_to_phipsi.cpp:38-131): the
// XTC file holds only the N-CA-C backbone atoms of one chain; every frame becomes 2*(natoms/3)-2
// torsion angles (radians) written as headerless doubles.  Frames are decoded by all host threads
// (trajectory.cpp), the torsions are computed on the GPU (featurize.cu, torsion() of mdsctk.cpp:643-676).
// New: -s/--sincos-file also writes the sin/cos embedding (what angles_to_sincos would produce from the
// angle file) from the same kernel launch.
#include "../../include/mdsctk_knn.h"
#include "options.hpp"
#include "trajectory.hpp"

#include <algorithm>
#include <fstream>
#include <iostream>
#include <thread>
#include <vector>

using namespace mdsctk_cli;

int main(int argc, char *argv[])
{
    const char *program_name = "bb_xtc_to_phipsi";
    banner(program_name);
    std::cout << "   Convert the provided XTC file to phipsi angles and" << std::endl;
    std::cout << "   write the results to the selected output file." << std::endl;
    std::cout << "   Note that this code anticipates a single protein" << std::endl;
    std::cout << "   chain, and only the N-CA-C atoms to be present in" << std::endl;
    std::cout << "   the XTC file." << std::endl << std::endl;
    std::cout << "   Use -h or --help to see the complete list of options." << std::endl << std::endl;

    Options po;
    po.add("help", 'h', Options::SWITCH, "show this help message and exit");
    po.add("xtc-file", 'x', Options::VALUE, "Input:  Trajectory file (string:filename)", "traj.xtc", true);
    po.add("output-file", 'o', Options::VALUE, "Output: Phi-phi angle data file (string:filename)", "phipsi.dat", true);
    po.add("sincos-file", 's', Options::VALUE, "Output: sin/cos embedding of the angles (string:filename, optional)");
    po.add("threads", 't', Options::VALUE, "Input:  Number of decode threads (int)",
           std::to_string(std::max(1u, std::thread::hardware_concurrency())), true);
    int nthreads;
    try {
        po.parse(argc, argv);
        if (po.count("help")) {
            std::cout << "usage: " << program_name << " [options]" << std::endl;
            po.print(std::cout, "Program options");
            return 1;
        }
        nthreads = po.integer("threads");
    } catch (const std::exception &e) {
        std::cout << "ERROR: " << e.what() << std::endl;
        return 2;
    }
    const std::string xtc_filename = po.str("xtc-file"), output_filename = po.str("output-file");
    std::cout << "Running with the following options:" << std::endl;
    std::cout << "xtc-file =    " << xtc_filename << std::endl;
    std::cout << "output-file = " << output_filename << std::endl << std::endl;

    XtcFile xtc;
    std::string err;
    if (!xtc.open(xtc_filename, &err)) { std::cout << "ERROR: " << err << std::endl; return 3; }
    const long long frames = xtc.frames();
    const int natoms = xtc.natoms;
    const long long T = 2 * (natoms / 3) - 2;
    if (frames <= 0 || T <= 0) { std::cout << "ERROR: need at least 6 backbone atoms per frame" << std::endl; return 3; }
    std::vector<float> xyz((size_t)frames * natoms * 3);
    if (!xtc.decode_all(xyz.data(), nthreads, &err)) { std::cout << "ERROR: " << err << std::endl; return 3; }

    const bool want_sc = po.count("sincos-file");
    std::vector<double> ang((size_t)frames * T), sc(want_sc ? (size_t)frames * T * 2 : 0);
    mdsctk_knn_ctx *ctx = nullptr;
    if (mdsctk_knn_create(&ctx, 0) != 0) { std::cout << "ERROR: " << mdsctk_knn_last_error(nullptr) << std::endl; return 5; }
    if (mdsctk_knn_phipsi(ctx, xyz.data(), frames, natoms, ang.data(), want_sc ? sc.data() : nullptr) != 0) {
        std::cout << "ERROR: " << mdsctk_knn_last_error(ctx) << std::endl;
        return 5;
    }
    mdsctk_knn_destroy(ctx);
    std::ofstream output(output_filename.c_str(), std::ios::binary | std::ios::trunc);
    output.write(reinterpret_cast<const char *>(ang.data()), (std::streamsize)(ang.size() * sizeof(double)));
    if (!output) { std::cout << "ERROR: cannot write " << output_filename << std::endl; return 3; }
    if (want_sc) {
        std::ofstream so(po.str("sincos-file").c_str(), std::ios::binary | std::ios::trunc);
        so.write(reinterpret_cast<const char *>(sc.data()), (std::streamsize)(sc.size() * sizeof(double)));
        if (!so) { std::cout << "ERROR: cannot write " << po.str("sincos-file") << std::endl; return 3; }
    }
    std::cout << "Wrote " << frames << " vectors of length " << T << " (" << (frames * T) << " total values)." << std::endl;
    std::cout << std::endl;
    return 0;
}
