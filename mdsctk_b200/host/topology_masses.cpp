// topology_masses -- prints the per-atom masses knn_rms derives from a .pdb / .gro topology (one per line).
// The reference takes them from GROMACS' read_tps_conf(..., bMass=TRUE) (knn_rms.cpp:150-153, 181-182); this is the
// same lookup the B200 knn_rms uses (trajectory.cpp), exposed so that it can be checked and so that a mass file for
// --mass-file can be produced and edited.  CPU only.
#include "trajectory.hpp"

#include <cstdio>
#include <iostream>
#include <vector>

int main(int argc, char *argv[])
{
    if (argc != 2) {
        std::cout << "usage: topology_masses topology.{pdb,gro}" << std::endl;
        return 1;
    }
    std::vector<float> mass;
    std::string err;
    if (!mdsctk_cli::read_topology_masses(argv[1], &mass, &err)) {
        std::cout << "ERROR: " << err << std::endl;
        return 3;
    }
    for (float m : mass) std::printf("%.5f\n", m);
    return 0;
}
