// plasma.cpp -- LBmethod over the C ABI of include/plbm.h.
// Mirrors the reference's constructor + Run_simulation (src/plasma.cpp:22-124, 459-529); the loop
// body itself (UpdateMacro, ComputeEquilibrium, Collide, Stream, SolvePoisson) runs on the device.
#include "plasma.hpp"
#include "plbm.h"
#include "visualize_frames.hpp"

#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>

// optional: only a visualiser that implements the alternate entry defines it (the reference's visualize.cpp does not)
namespace visualize {
void UpdateVisualizationFrames(int, int, int, const float* const[VF_COUNT], const double[VS_COUNT][9]) __attribute__((weak));
}

namespace {
[[noreturn]] void raise(const char* what)
{
    throw std::runtime_error(std::string(what) + ": " + plbm_last_error());
}
}

LBmethod::LBmethod(const int _NSTEPS, const int _NX, const int _NY, const size_t _n_cores,
                   const int Z_ion, const int A_ion, const double Ex_SI, const double Ey_SI,
                   const double T_e_SI_init, const double T_i_SI_init, const double T_n_SI_init,
                   const double n_e_SI_init, const double n_n_SI_init,
                   const poisson::PoissonType poisson_type, const streaming::BCType bc_type, const double omega_sor)
    : NSTEPS(_NSTEPS), NX(_NX), NY(_NY), n_cores(_n_cores)
{
    plbm_config cfg = {};
    cfg.NX = NX; cfg.NY = NY;
    cfg.poisson_type = static_cast<int>(poisson_type);     // same enumerator order as the reference
    cfg.bc_type = static_cast<int>(bc_type);
    cfg.omega_sor = omega_sor;
    cfg.rank = 0; cfg.nranks = 1; cfg.y0 = 0; cfg.NY_local = NY; cfg.device = -1;
    if (plbm_units_from_si(Z_ion, A_ion, Ex_SI, Ey_SI, T_e_SI_init, T_i_SI_init, T_n_SI_init, n_e_SI_init, n_n_SI_init, &cfg))
        raise("LBmethod: unit conversion");
    // PLBM_DEVICES=N: the same object on N GPUs (one context per device, slabs wired through peer memory, one host thread per device)
    int ndev = 1;
    if (const char* e = std::getenv("PLBM_DEVICES")) ndev = std::atoi(e);
    const bool sliceable = poisson_type == poisson::PoissonType::FFT && bc_type == streaming::BCType::Periodic;
    if (ndev > 1 && sliceable) {
        if (plbm_group_create(&cfg, ndev, &group_)) raise("LBmethod: device state (PLBM_DEVICES)");
        if (plbm_group_initialize(group_)) { plbm_group_destroy(group_); group_ = nullptr; raise("LBmethod: Initialize"); }
    } else {
        if (ndev > 1) std::cerr << "LBmethod: PLBM_DEVICES=" << ndev << " ignored (several GPUs need the periodic spectral configuration)" << std::endl;
        if (plbm_create(&cfg, &ctx_)) raise("LBmethod: device state");
        if (plbm_initialize(ctx_)) { plbm_destroy(ctx_); ctx_ = nullptr; raise("LBmethod: Initialize"); }
    }
    const size_t n = static_cast<size_t>(NX) * NY;
    for (auto& f : fields_) f.assign(n, 0.0);
}

LBmethod::~LBmethod()
{
    if (ctx_) plbm_fetch_wait(ctx_);
    if (group_) plbm_group_fetch_wait(group_);
    for (void* p : pinned_ptrs_) plbm_unpin_host(p);           // exactly what was registered
    plbm_destroy(ctx_);
    plbm_group_destroy(group_);
}

int LBmethod::step_(int nsteps, int want_fields)
{
    return group_ ? plbm_group_step(group_, nsteps, want_fields) : plbm_step(ctx_, nsteps, want_fields);
}

// Run_simulation hands every step's 15 fields to the visualiser: the copy of step t+1 runs on the library's copy
// stream into inflight_ while the device already computes and the host draws step t from fields_.
void LBmethod::begin_fetch()
{
    double* out[PLBM_NUM_FIELDS] = {};
    for (int k = 0; k < 15; ++k) out[k] = inflight_[k].data();
    if (group_ ? plbm_group_fetch_begin(group_, out) : plbm_fetch_begin(ctx_, out)) raise("LBmethod: field fetch");
}
void LBmethod::finish_fetch()
{
    if (group_ ? plbm_group_fetch_wait(group_) : plbm_fetch_wait(ctx_)) raise("LBmethod: field fetch");
    for (int k = 0; k < 15; ++k) fields_[k].swap(inflight_[k]);
}

void LBmethod::fetch_fields()
{
    double* out[PLBM_NUM_FIELDS] = {};
    for (int k = 0; k < 15; ++k) out[k] = fields_[k].data();
    if (group_ ? plbm_group_download_fields(group_, out) : plbm_download_fields(ctx_, out)) raise("LBmethod: field download");
}

void LBmethod::FetchPotential(std::vector<double>& phi) const
{
    phi.resize(static_cast<size_t>(NX) * NY);
    double* out[PLBM_NUM_FIELDS] = {};
    out[PLBM_F_PHI] = phi.data();
    if (group_ ? plbm_group_download_fields(group_, out) : plbm_download_fields(ctx_, out)) raise("LBmethod: potential download");
}

void LBmethod::Step(int nsteps, bool want_fields)
{
    if (step_(nsteps, want_fields ? 1 : 0)) raise("LBmethod: time step");
    if (want_fields) fetch_fields();
    else if (group_ ? plbm_group_sync(group_) : plbm_sync(ctx_)) raise("LBmethod: sync");
}

void LBmethod::Run_simulation()
{
    visualize::InitVisualization(NX, NY, NSTEPS);
    if (!pinned_) {                                           // page-lock both field sets once (best effort, tracked per buffer)
        const size_t bytes = sizeof(double) * static_cast<size_t>(NX) * NY;
        for (auto& f : inflight_) f.assign(static_cast<size_t>(NX) * NY, 0.0);
        pinned_ = true;
        for (auto* set : { fields_, inflight_ })
            for (int k = 0; k < 15; ++k)
                if (plbm_pin_host(set[k].data(), bytes) == 0) pinned_ptrs_.push_back(set[k].data());
    }
    if (NSTEPS > 0) {
        if (step_(1, 1)) raise("LBmethod: time step");
        begin_fetch();
    }
    for (int t = 0; t < NSTEPS; ++t) {
        finish_fetch();                                       // fields_ = step t
        if (t + 1 < NSTEPS) {
            if (step_(1, 1)) raise("LBmethod: time step");
            begin_fetch();
        }
        visualize::UpdateVisualization(t, NX, NY,
                                       fields_[PLBM_F_UX_E], fields_[PLBM_F_UY_E],
                                       fields_[PLBM_F_UX_I], fields_[PLBM_F_UY_I],
                                       fields_[PLBM_F_UX_N], fields_[PLBM_F_UY_N],
                                       fields_[PLBM_F_T_E], fields_[PLBM_F_T_I], fields_[PLBM_F_T_N],
                                       fields_[PLBM_F_RHO_E], fields_[PLBM_F_RHO_I], fields_[PLBM_F_RHO_N],
                                       fields_[PLBM_F_RHO_Q], fields_[PLBM_F_EX], fields_[PLBM_F_EY]);
    }
    visualize::CloseVisualization();
    std::cout << "Simulation ended " << std::endl;
}

void LBmethod::Run_simulation_frames()
{
    if (group_) throw std::runtime_error("LBmethod::Run_simulation_frames: not available with PLBM_DEVICES (use Run_simulation)");
    if (!visualize::UpdateVisualizationFrames)
        throw std::runtime_error("LBmethod::Run_simulation_frames: the linked visualiser does not implement visualize::UpdateVisualizationFrames");
    visualize::InitVisualization(NX, NY, NSTEPS);
    const size_t n = static_cast<size_t>(NX) * NY;
    // two host sets: the one being drawn and the one being filled by the copy stream
    std::vector<float> mats[2][PLBM_NUM_FRAMES];
    std::vector<double> series[2];
    for (int b = 0; b < 2; ++b) {
        for (auto& m : mats[b]) { m.assign(n, 0.0f); plbm_pin_host(m.data(), sizeof(float) * n); }
        series[b].assign(static_cast<size_t>(PLBM_NUM_SERIES) * PLBM_NUM_POINTS, 0.0);
    }
    auto begin = [&](int b) {
        float* out[PLBM_NUM_FRAMES];
        for (int k = 0; k < PLBM_NUM_FRAMES; ++k) out[k] = mats[b][k].data();
        if (plbm_frames_begin(ctx_, out, series[b].data())) raise("LBmethod: frame fetch");
    };
    if (NSTEPS > 0) {
        if (plbm_step(ctx_, 1, 1)) raise("LBmethod: time step");
        begin(0);
    }
    for (int t = 0; t < NSTEPS; ++t) {
        const int cur = t & 1;
        if (plbm_fetch_wait(ctx_)) raise("LBmethod: frame fetch");
        if (t + 1 < NSTEPS) {
            if (plbm_step(ctx_, 1, 1)) raise("LBmethod: time step");
            begin(cur ^ 1);
        }
        const float* m[PLBM_NUM_FRAMES];
        for (int k = 0; k < PLBM_NUM_FRAMES; ++k) m[k] = mats[cur][k].data();
        visualize::UpdateVisualizationFrames(t, NX, NY, m, reinterpret_cast<const double (*)[PLBM_NUM_POINTS]>(series[cur].data()));
    }
    for (int b = 0; b < 2; ++b) for (auto& m : mats[b]) plbm_unpin_host(m.data());
    visualize::CloseVisualization();
    std::cout << "Simulation ended " << std::endl;
}
