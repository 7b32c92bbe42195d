// tests/host/drive_lbmethod.cpp -- what reference src/main_plasma.cpp does (construct LBmethod from
// SI parameters, Run_simulation), with the lattice and solver choice on the command line.
#include "plasma.hpp"

#include <cstdio>
#include <cstdlib>
#include <exception>

int main(int argc, char** argv)
{
    if (argc < 6) { std::fprintf(stderr, "usage: %s NX NY NSTEPS poisson bc [frames]\n", argv[0]); return 2; }
    const bool frames = argc > 6 && std::atoi(argv[6]) != 0;      // alternate output path (visualize_frames.hpp)
    const int NX = std::atoi(argv[1]), NY = std::atoi(argv[2]), NSTEPS = std::atoi(argv[3]);
    const auto ptype = static_cast<poisson::PoissonType>(std::atoi(argv[4]));
    const auto btype = static_cast<streaming::BCType>(std::atoi(argv[5]));
    try {
        // physical parameters of reference src/main_plasma.cpp:22-51
        LBmethod lb(NSTEPS, NX, NY, 1, 1, 1, 1e-2, 0.0, 1e4, 300, 300, 1e11, 1e18, ptype, btype, 1.8);
        if (frames) lb.Run_simulation_frames();
        else lb.Run_simulation();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
