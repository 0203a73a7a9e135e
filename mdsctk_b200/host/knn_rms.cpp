// knn_rms -- B200 build of MDSCTK's knn_rms tool.
//
// Same command line, stdout and output files as the reference tool (knn_rms.cpp:43-307): for each
// frame of the fitting trajectory the k nearest reference frames by optimal-superposition,
// mass-weighted RMSD, written as headerless row-major double[n_fit][k] (Angstrom) and
// int32[n_fit][k], sorted position 0 dropped (knn_rms.cpp:282-291) -- make_sysparse /
// make_gesparse read them unchanged.  The per-pair work runs on the GPU(s) through the C ABI
// of include/mdsctk_knn.h; there is no CPU fallback.
//
// Differences from the reference, all documented in INTEGRATION.md: --threads sets the host
// threads used to decode xtc frames (the GPU does the distance work), --block-size is accepted
// and only echoed, k AND k+1 are clamped to the frame count (the reference clamps k only,
// knn_rms.cpp:131,224-225), --sort false writes exact FP64 rows with identity indices (the
// reference writes an uninitialised index vector), and two options are new: --gpus, --mass-file.
#include "../../include/mdsctk_knn.h"
#include "options.hpp"
#include "replicate_nccl.hpp"
#include "trajectory.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <thread>
#include <vector>

using namespace mdsctk_cli;

namespace {

struct Gpu {
    mdsctk_knn_ctx *ctx = nullptr;
    int dev = 0;
};

bool replicate_reference(std::vector<Gpu> &gpus, const float *xyz, long long n, int natoms, const float *mass)
{
    // every GPU uploads + packs its own shard of the reference frames, NCCL replicates the packed arrays (replicate_nccl.hpp)
    std::string err;
    double nccl_ms = 0.0;
    const bool ok = replicate_reference_nccl(
        (int)gpus.size(), n,
        [&](int g, ShardRange s) {
            int rc = mdsctk_knn_rms_alloc_reference(gpus[g].ctx, n, natoms, mass);
            if (rc == 0 && s.count > 0) rc = mdsctk_knn_rms_pack_shard(gpus[g].ctx, xyz + (size_t)s.begin * natoms * 3, s.begin, s.count);
            return rc;
        },
        [&](int g, void **ptrs, size_t *bpf, int *na) { return mdsctk_knn_rms_reference_arrays(gpus[g].ctx, 16, na, ptrs, bpf); },
        [&](int g) { return mdsctk_knn_last_error(gpus[g].ctx); }, &err, &nccl_ms);
    if (!ok) std::cout << "ERROR: " << err << std::endl;
    else if (gpus.size() > 1) std::cout << "Reference set packed on " << gpus.size() << " GPUs and replicated with NCCL in " << nccl_ms << " ms." << std::endl;
    return ok;
}

}  // namespace

int main(int argc, char *argv[])
{
    const char *program_name = "knn_rms";
    banner(program_name);
    std::cout << "   Computes the k nearest neighbors of all reference" << std::endl;
    std::cout << "   structures in the given xtc file for each structure" << std::endl;
    std::cout << "   in the given fitting xtc file. (Uses the same file for" << std::endl;
    std::cout << "   as reference by default to make a symmetric comparison.)" << std::endl;
    std::cout << "   A topology PDB file should be provided for determining" << std::endl;
    std::cout << "   the mass of each atom." << std::endl << std::endl;
    std::cout << "   Use -h or --help to see the complete list of options." << std::endl << std::endl;

    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    Options po;
    po.add("help", 'h', Options::SWITCH, "show this help message and exit");
    po.add("threads", 't', Options::VALUE, "Input:  Number of host threads for xtc decoding (int)", std::to_string(hw), true);
    po.add("knn", 'k', Options::VALUE, "Input:  K-nearest neighbors (int)");
    po.add("sort", 's', Options::VALUE, "Input:  Find K-nn,false=full distance matix (bool)", "1", true);
    po.add("nofit", 'n', Options::VALUE, "Input:  DO NOT rotate+fit before RMSD calculation (bool)", "0", true);
    po.add("block-size", 'b', Options::VALUE, "Input:  Workgroup block size in # frames (int, ignored on GPU)", "128", true);
    po.add("topology-file", 'p', Options::VALUE, "Input:  Topology file [.pdb,.gro] (string:filename)", "topology.pdb", true);
    po.add("reference-file", 'r', Options::VALUE, "Input:  Reference [.xtc] file (string:filename)", "reference.xtc", true);
    po.add("fit-file", 'f', Options::VALUE, "Input:  Fitting [.xtc] file (string:filename)");
    po.add("distance-file", 'd', Options::VALUE, "Output: K-nn distances file (string:filename)", "distances.dat", true);
    po.add("index-file", 'i', Options::VALUE, "Output: K-nn indices file (string:filename)", "indices.dat", true);
    po.add("gpus", 'g', Options::VALUE, "Input:  Number of GPUs; fit rows are sharded across them (int)", "1", true);
    po.add("mass-file", 'm', Options::VALUE, "Input:  One mass per atom, overrides the topology lookup (string:filename)");

    int nthreads, k = 0, blksize, ngpus;
    bool sort, nofit;
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
        nofit = po.boolean("nofit");
        blksize = po.integer("block-size");
        ngpus = po.integer("gpus");
    } catch (const std::exception &e) {
        std::cout << "ERROR: " << e.what() << std::endl;
        return 2;
    }
    if (!po.count("knn") && sort) {
        std::cout << "ERROR: --knn not supplied." << std::endl << std::endl;
        return -1;
    }
    const std::string top_filename = po.str("topology-file"), ref_filename = po.str("reference-file");
    const std::string fit_filename = po.count("fit-file") ? po.str("fit-file") : ref_filename;
    const std::string d_filename = po.str("distance-file"), i_filename = po.str("index-file");

    std::cout << "Running with the following options:" << std::endl;
    std::cout << "threads =        " << nthreads << std::endl;
    std::cout << "knn =            " << k << std::endl;
    std::cout << "sort =           " << sort << std::endl;
    std::cout << "nofit =          " << nofit << std::endl;
    std::cout << "topology-file =  " << top_filename << std::endl;
    std::cout << "reference-file = " << ref_filename << std::endl;
    std::cout << "fit-file =       " << fit_filename << std::endl;
    std::cout << "distance-file =  " << d_filename << std::endl;
    std::cout << "index-file =     " << i_filename << std::endl;
    std::cout << std::endl;

    std::string err;
    std::vector<float> mass;
    std::cout << "Reading topology information from " << top_filename << " ... ";
    if (po.count("mass-file") ? !read_mass_file(po.str("mass-file"), &mass, &err)
                              : !read_topology_masses(top_filename, &mass, &err)) {
        std::cout << std::endl << "ERROR: " << err << std::endl;
        return 3;
    }
    std::cout << "done." << std::endl;

    // a flat float32 .crd file (flatten_xtc output) is accepted in place of an .xtc
    auto open_any = [&](XtcFile &f, const std::string &path) {
        if (path.size() > 4 && path.compare(path.size() - 4, 4, ".crd") == 0) {
            std::ifstream in(path.c_str(), std::ios::binary | std::ios::ate);
            if (!in) { err = "cannot open " + path; return false; }
            const size_t bytes = (size_t)in.tellg(), per = mass.size() * 12;
            in.seekg(0);
            f.natoms = (int)mass.size();
            f.bytes.resize(bytes / per * per);
            in.read(reinterpret_cast<char *>(f.bytes.data()), (std::streamsize)f.bytes.size());
            f.frame_offset.clear();
            for (size_t o = 0; o + per <= f.bytes.size(); o += per) f.frame_offset.push_back(o);
            f.is_flat = true;
            if (f.frame_offset.empty()) { err = path + " holds no frame"; return false; }
            return true;
        }
        return f.open(path, &err);
    };
    XtcFile ref_file, fit_file_storage;
    if (!open_any(ref_file, ref_filename)) { std::cout << "ERROR: " << err << std::endl; return 3; }
    const bool same = fit_filename == ref_filename;
    if (!same && !open_any(fit_file_storage, fit_filename)) { std::cout << "ERROR: " << err << std::endl; return 3; }
    const XtcFile &fit_file = same ? ref_file : fit_file_storage;
    const XtcFile *both[2] = {&ref_file, &fit_file};
    for (const XtcFile *f : both) {
        if (f->natoms != (int)mass.size()) {  // knn_rms.cpp:158-178
            std::cout << "*** ERROR ***" << std::endl;
            std::cout << "Number of atoms in topology file (" << mass.size() << ") "
                      << "does not match the number of atoms "
                      << "in the XTC file (" << (f == &ref_file ? ref_filename : fit_filename) << " : " << f->natoms << ")."
                      << std::endl;
            return 4;
        }
    }
    const int natoms = ref_file.natoms;
    const long long n_ref = ref_file.frames(), n_fit = fit_file.frames();

    // pinned host staging so the H2D copies run at full PCIe rate
    float *ref_xyz = nullptr, *fit_xyz = nullptr;
    if (cudaMallocHost(&ref_xyz, (size_t)n_ref * natoms * 12) != cudaSuccess ||
        (!same && cudaMallocHost(&fit_xyz, (size_t)n_fit * natoms * 12) != cudaSuccess)) {
        std::cout << "ERROR: cannot allocate pinned host memory (is a CUDA device present?)" << std::endl;
        return 5;
    }
    std::cout << "Reading reference coordinates from file: " << ref_filename << " ... ";
    std::cout.flush();
    if (!ref_file.decode_all(ref_xyz, nthreads, &err)) { std::cout << "ERROR: " << err << std::endl; return 3; }
    std::cout << "done." << std::endl;
    std::cout << "Reading fitting coordinates from file: " << fit_filename << " ... ";
    std::cout.flush();
    if (!same && !fit_file.decode_all(fit_xyz, nthreads, &err)) { std::cout << "ERROR: " << err << std::endl; return 3; }
    std::cout << "done." << std::endl;

    std::ofstream distances(d_filename.c_str(), std::ios::binary | std::ios::trunc);
    std::ofstream indices(i_filename.c_str(), std::ios::binary | std::ios::trunc);
    if (!distances || !indices) { std::cout << "ERROR: cannot open the output files" << std::endl; return 3; }

    if ((long long)blksize > n_fit) blksize = (int)n_fit;
    std::cout << "Block size: " << blksize << std::endl;

    // Fix k if number of frames is too small (both k and k1, unlike knn_rms.cpp:224-225)
    if (n_ref - 1 < k) k = (int)(n_ref - 1);
    const int k1 = k + 1;

    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    ngpus = std::max(1, std::min(ngpus, std::max(ndev, 1)));
    std::vector<Gpu> gpus(ngpus);
    for (int g = 0; g < ngpus; ++g) {
        gpus[g].dev = g;
        if (mdsctk_knn_create(&gpus[g].ctx, g) != 0) {
            std::cout << "ERROR: " << mdsctk_knn_last_error(nullptr) << std::endl;
            return 5;
        }
    }
    if (!replicate_reference(gpus, ref_xyz, n_ref, natoms, mass.data())) return 5;

    int rc = 0;
    if (sort) {
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
                status[g] = same ? mdsctk_knn_rms_query_range(gpus[g].ctx, b, n, k1, !nofit, od, oi)
                                 : mdsctk_knn_rms_query(gpus[g].ctx, fit_xyz + (size_t)b * natoms * 3, n, k1, !nofit, od, oi);
            });
        }
        for (auto &t : pool) t.join();
        for (int g = 0; g < ngpus; ++g)
            if (status[g] != 0) { std::cout << "ERROR: " << mdsctk_knn_last_error(gpus[g].ctx) << std::endl; rc = 5; }
        if (rc == 0) {
            // closest k RMSD alignment scores and indices; sorted position 0 is dropped (knn_rms.cpp:282-291)
            for (long long f = 0; f < n_fit; ++f) {
                distances.write(reinterpret_cast<const char *>(&dist[(size_t)f * k1 + 1]), sizeof(double) * k);
                indices.write(reinterpret_cast<const char *>(&idx[(size_t)f * k1 + 1]), sizeof(int) * k);
            }
        }
    } else {
        if (!same) {
            std::cout << "ERROR: --sort false with a separate --fit-file is not supported by this build" << std::endl;
            rc = 6;
        } else {
            const long long rows = std::max<long long>(1, std::min<long long>(n_fit, (64LL << 20) / (n_ref * 8)));
            std::vector<double> buf((size_t)rows * n_ref);
            std::vector<int> ident((size_t)n_ref);
            for (long long j = 0; j < n_ref; ++j) ident[(size_t)j] = (int)j;
            for (long long f = 0; f < n_fit && rc == 0; f += rows) {
                const long long n = std::min(rows, n_fit - f);
                if (mdsctk_knn_rms_rows(gpus[0].ctx, f, n, !nofit, buf.data()) != 0) {
                    std::cout << "ERROR: " << mdsctk_knn_last_error(gpus[0].ctx) << std::endl;
                    rc = 5;
                    break;
                }
                for (long long r = 0; r < n; ++r) {
                    distances.write(reinterpret_cast<const char *>(&buf[(size_t)r * n_ref]), sizeof(double) * n_ref);
                    indices.write(reinterpret_cast<const char *>(ident.data()), sizeof(int) * n_ref);
                }
            }
        }
    }
    std::cout << std::endl << std::endl;

    for (auto &g : gpus) mdsctk_knn_destroy(g.ctx);
    cudaFreeHost(ref_xyz);
    if (fit_xyz) cudaFreeHost(fit_xyz);
    return rc;
}
