// plasma.hpp -- class LBmethod with the reference's public surface (reference
// include/plasma.hpp:18-52): the 16-argument constructor taking SI parameters and
// Run_simulation().  src/main_plasma.cpp of the reference compiles against this header unchanged.
//
// Unlike the reference, the object owns no lattice-sized host state besides the 15 fields it
// hands to the visualiser: populations, moments, potential and field live in B200 HBM inside a
// plbm_ctx (include/plbm.h), and one time step is the fused kernel + Poisson solve of libplbm.so.
#pragma once

#include "collisions.hpp"
#include "poisson.hpp"
#include "streaming.hpp"
#include "utils.hpp"
#include "visualize.hpp"

#include <array>
#include <cstddef>
#include <vector>

struct plbm_ctx;
struct plbm_group;

class LBmethod {
public:
    // Same parameter list as the reference (include/plasma.hpp:34-49).  n_cores is accepted and
    // ignored: there is no OpenMP path.  With PLBM_DEVICES=N (2..16) in the environment the lattice is cut into N y-slabs on
    // the GPUs 0..N-1 of the box (periodic boundaries + spectral Poisson only; anything else keeps one GPU) -- main_plasma.cpp
    // does not change.  Throws std::runtime_error when no CUDA device is present
    // or the configuration is not available on the device.
    LBmethod(const int NSTEPS, const int NX, const int NY, const size_t n_cores,
             const int Z_ion, const int A_ion,
             const double Ex_SI, const double Ey_SI,
             const double T_e_SI_init, const double T_i_SI_init, const double T_n_SI_init,
             const double n_e_SI_init, const double n_n_SI_init,
             const poisson::PoissonType poisson_type, const streaming::BCType bc_type,
             const double omega_sor);
    ~LBmethod();
    LBmethod(const LBmethod&) = delete;
    LBmethod& operator=(const LBmethod&) = delete;

    // The complete time loop of the reference (src/plasma.cpp:459-529): per step advance the
    // lattice on the device, copy the 15 visualised fields to the host, call the visualiser.
    void Run_simulation();
    // same loop through the alternate output path: the visualiser's CV_32F matrices and sample-point series are
    // formed on the device and handed to visualize::UpdateVisualizationFrames (include/visualize_frames.hpp)
    void Run_simulation_frames();

    // ---- additions (not in the reference) used by benchmarks and tests -----------------------
    void Step(int nsteps, bool fetch_fields);                 // nsteps of the loop body without visualisation
    const std::vector<double>& Field(int plbm_field_id) const { return fields_[plbm_field_id]; }
    void FetchPotential(std::vector<double>& phi) const;

private:
    const int NSTEPS, NX, NY;
    const size_t n_cores;
    plbm_ctx* ctx_ = nullptr;                                 // one GPU ...
    plbm_group* group_ = nullptr;                             // ... or, with PLBM_DEVICES=N in the environment, N GPUs (y-slabs, include/plbm.h)
    int step_(int nsteps, int want_fields);
    std::vector<double> fields_[15];                          // order of visualize::UpdateVisualization
    std::vector<double> inflight_[15];                        // Run_simulation: the step being copied while fields_ is drawn
    bool pinned_ = false;                                     // Run_simulation has tried to page-lock the field sets
    std::vector<void*> pinned_ptrs_;                          // the buffers that were actually registered
    void begin_fetch();
    void finish_fetch();
    void fetch_fields();
};
