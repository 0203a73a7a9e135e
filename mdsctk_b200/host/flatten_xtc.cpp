// flatten_xtc -- exports an xtc trajectory as a flat file of native float32 coordinates,
// float[frame][atom][3] in nm (the .crd layout of the reference tool, flatten_xtc.cpp:110).
// Same options as the reference (flatten_xtc.cpp:55-60) plus -t/--threads: frames are indexed by a
// header walk and decoded in parallel.  knn_rms accepts the .crd file in place of an .xtc.
#include "options.hpp"
#include "trajectory.hpp"

#include <algorithm>
#include <cstring>
#include <fstream>
#include <iostream>
#include <thread>
#include <vector>

using namespace mdsctk_cli;

int main(int argc, char *argv[])
{
    const char *program_name = "flatten_xtc";
    banner(program_name);
    std::cout << "   Exports the provided xtc file into a simple 32-bit." << std::endl;
    std::cout << "   flat file." << std::endl << std::endl;
    std::cout << "   Use -h or --help to see the complete list of options." << std::endl << std::endl;
    Options po;
    po.add("help", 'h', Options::SWITCH, "show this help message and exit");
    po.add("xtc-file", 'x', Options::VALUE, "Input:  Trajectory file (string:filename)", "traj.xtc", true);
    po.add("output-file", 'o', Options::VALUE, "Output: Flattened trajectory file (string:filename)", "traj.crd", true);
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
    std::vector<float> xyz((size_t)xtc.frames() * xtc.natoms * 3);
    if (!xtc.decode_all(xyz.data(), nthreads, &err)) { std::cout << "ERROR: " << err << std::endl; return 3; }
    std::ofstream out(output_filename.c_str(), std::ios::binary | std::ios::trunc);
    out.write(reinterpret_cast<const char *>(xyz.data()), (std::streamsize)(xyz.size() * sizeof(float)));
    if (!out) { std::cout << "ERROR: cannot write " << output_filename << std::endl; return 3; }

    auto be_float = [&](size_t off) {
        const unsigned char *p = xtc.bytes.data() + off;
        unsigned u = (unsigned)p[0] << 24 | (unsigned)p[1] << 16 | (unsigned)p[2] << 8 | p[3];
        float f;
        std::memcpy(&f, &u, 4);
        return f;
    };
    std::cout << "XTC Statistics - " << xtc_filename << std::endl;
    std::cout << "Number of frames: " << xtc.frames() << std::endl;
    std::cout << "Nunber of atoms:  " << xtc.natoms << std::endl;
    std::cout << "Start time:       " << be_float(xtc.frame_offset.front() + 12) << std::endl;
    std::cout << "End time:         " << be_float(xtc.frame_offset.back() + 12) << std::endl;
    if (xtc.natoms > 9) std::cout << "Precision:        " << be_float(xtc.frame_offset.back() + 56) << std::endl;
    std::cout << std::endl << "All frames flattened..." << std::endl << std::endl;
    return 0;
}
