// trajectory.hpp -- frame loading for knn_rms: XTC trajectories and topology masses.
//
// Replaces the GROMACS calls of the reference's load phase (knn_rms.cpp:150-157,186-203:
// read_tps_conf(..., bMass=TRUE), open_xtc / read_first_xtc / read_next_xtc).  Frames are decoded
// on the host -- the xtc bit stream is serial within a frame -- but the file is first indexed by
// walking the frame headers, so frames are decoded by all host threads in parallel (the
// reference decodes the file serially, twice).  Centring and packing happen on the GPU.
#pragma once
#include <string>
#include <vector>

namespace mdsctk_cli {

struct XtcFile {
    std::vector<unsigned char> bytes;
    std::vector<size_t> frame_offset;   // start of every frame header
    int natoms = 0;
    bool is_flat = false;               // bytes already hold float[frames][natoms][3] (.crd)
    bool open(const std::string &path, std::string *err);
    long long frames() const { return (long long)frame_offset.size(); }
    // xyz: float[frames][natoms][3] (nm), AoS, uncentred.  Returns false on a corrupt frame.
    bool decode_all(float *xyz, int nthreads, std::string *err) const;
};

// Per-atom masses from the atom names of a .pdb / .gro topology (GROMACS atommass.dat values).
// An optional mass file (one number per line) overrides the lookup.
bool read_topology_masses(const std::string &path, std::vector<float> *mass, std::string *err);
bool read_mass_file(const std::string &path, std::vector<float> *mass, std::string *err);

}  // namespace mdsctk_cli
