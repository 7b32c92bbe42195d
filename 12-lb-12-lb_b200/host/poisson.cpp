// poisson.cpp -- poisson::SolvePoisson and friends with the reference's signatures
// (include/poisson.hpp) over a process-global, fields-only plbm_ctx.  The reference keeps phi and
// its FFTW plans in file statics initialised by std::call_once (src/poisson.cpp:9-23, 34-41,
// 373-376): the lattice size of the first call is the one that sticks, and so it is here for the
// device context.  Solver type, boundary type and omega are arguments of EVERY call in the
// reference (src/poisson.cpp:25-82) and are handed to the context per call (plbm_host_poisson_config).
#include "poisson.hpp"
#include "plbm.h"

#include <cstdlib>
#include <mutex>
#include <stdexcept>
#include <string>

namespace poisson {

namespace {

plbm_ctx* g_ctx = nullptr;
std::once_flag g_once;

void destroy_ctx() { plbm_destroy(g_ctx); g_ctx = nullptr; }

[[noreturn]] void raise(const char* what) { throw std::runtime_error(std::string(what) + ": " + plbm_last_error()); }

plbm_ctx* context(int NX, int NY, PoissonType type, streaming::BCType bc, double omega)
{
    std::call_once(g_once, [&]() {
        plbm_config cfg = {};
        cfg.NX = NX; cfg.NY = NY;
        cfg.poisson_type = static_cast<int>(type);
        cfg.bc_type = static_cast<int>(bc);
        cfg.omega_sor = omega;
        cfg.cs2 = 1.0 / 3.0; cfg.Kb = 1.0 / 3.0;               // unused by the Poisson path, must be valid divisors
        cfg.m[0] = cfg.m[1] = cfg.m[2] = 1.0;
        cfg.nranks = 1; cfg.NY_local = NY; cfg.device = -1;
        cfg.fields_only = 1;
        if (plbm_create(&cfg, &g_ctx)) raise("poisson: device state");
        std::atexit(destroy_ctx);                               // reference src/poisson.cpp:375
    });
    if (!g_ctx) throw std::runtime_error("poisson: device state unavailable");
    if (plbm_host_poisson_config(g_ctx, static_cast<int>(type), static_cast<int>(bc), omega)) raise("poisson: solver selection");
    return g_ctx;
}

void solver(const std::vector<double>& rho_q, int NX, int NY, PoissonType type, double omega, const char* what, int code = -1)
{
    plbm_ctx* c = context(NX, NY, type, streaming::BCType::Periodic, omega);
    if (plbm_host_poisson_solver(c, code >= 0 ? code : static_cast<int>(type), rho_q.data())) raise(what);
}

} // namespace

void SolvePoisson(std::vector<double>& Ex, std::vector<double>& Ey, const std::vector<double>& rho_q,
                  const int NX, const int NY, const double omega, const PoissonType type, const streaming::BCType bc_type)
{
    plbm_ctx* c = context(NX, NY, type, bc_type, omega);
    if (plbm_host_solve_poisson(c, rho_q.data(), Ex.data(), Ey.data())) raise("poisson::SolvePoisson");
}

void SolvePoisson_GS(const std::vector<double>& rho_q, const int NX, const int NY) { solver(rho_q, NX, NY, PoissonType::GS, 0.0, "poisson::SolvePoisson_GS"); }
void SolvePoisson_SOR(const std::vector<double>& rho_q, const int NX, const int NY, const double omega) { solver(rho_q, NX, NY, PoissonType::SOR, omega, "poisson::SolvePoisson_SOR"); }
void SolvePoisson_FFT(const std::vector<double>& rho_q, const int NX, const int NY) { solver(rho_q, NX, NY, PoissonType::FFT, 0.0, "poisson::SolvePoisson_FFT"); }
void SolvePoisson_9point(const std::vector<double>& rho_q, const int NX, const int NY) { solver(rho_q, NX, NY, PoissonType::NPS, 0.0, "poisson::SolvePoisson_9point"); }

// The *_Periodic solvers (reference src/poisson.cpp:146-211, 283-354, 487-546) are never called by the reference's own loop
// (its Periodic branch runs the Dirichlet solvers, :48-56) but they are public: same sweeps over all cells with wrapped neighbours.
void SolvePoisson_GS_Periodic(const std::vector<double>& rho_q, const int NX, const int NY) { solver(rho_q, NX, NY, PoissonType::GS, 0.0, "poisson::SolvePoisson_GS_Periodic", PLBM_POISSON_GS_PERIODIC); }
void SolvePoisson_SOR_Periodic(const std::vector<double>& rho_q, const int NX, const int NY, const double omega) { solver(rho_q, NX, NY, PoissonType::SOR, omega, "poisson::SolvePoisson_SOR_Periodic", PLBM_POISSON_SOR_PERIODIC); }
void SolvePoisson_9point_Periodic(const std::vector<double>& rho_q, const int NX, const int NY) { solver(rho_q, NX, NY, PoissonType::NPS, 0.0, "poisson::SolvePoisson_9point_Periodic", PLBM_POISSON_NPS_PERIODIC); }

void ComputeElectricField(std::vector<double>& Ex, std::vector<double>& Ey, const int NX, const int NY)
{
    plbm_ctx* c = context(NX, NY, PoissonType::GS, streaming::BCType::BounceBack, 0.0);
    if (plbm_host_efield(c, PLBM_BC_BOUNCEBACK, Ex.data(), Ey.data())) raise("poisson::ComputeElectricField");
}

void ComputeElectricField_Periodic(std::vector<double>& Ex, std::vector<double>& Ey, const int NX, const int NY)
{
    plbm_ctx* c = context(NX, NY, PoissonType::FFT, streaming::BCType::Periodic, 0.0);
    if (plbm_host_efield(c, PLBM_BC_PERIODIC, Ex.data(), Ey.data())) raise("poisson::ComputeElectricField_Periodic");
}

// plan management is internal to the device context (reference src/poisson.cpp:611-635)
void InitPoissonFFT(const int NX, const int NY, const int, const int) { context(NX, NY, PoissonType::FFT, streaming::BCType::Periodic, 0.0); }
void FinalizePoissonFFT() {}

} // namespace poisson
