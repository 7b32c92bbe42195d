// poisson.hpp -- public surface of the reference's Poisson module (reference
// include/poisson.hpp:15-112), implemented on the B200 through include/plbm.h.
// Like the reference (static phi and plans, src/poisson.cpp:9-23) the module keeps hidden
// process-global state: the potential lives on the device between calls (warm start for the
// iterative solvers) and the first call fixes the lattice size.
// The reference header pulls in <fftw3.h>; nothing here needs FFTW -- the spectral solve is the
// library's own FFT.
#pragma once

#include "streaming.hpp"   // BCType
#include "utils.hpp"

#include <cmath>
#include <cstddef>
#include <vector>

namespace poisson {

enum class PoissonType {
    NONE,
    GS,
    SOR,
    FFT,
    NPS
};

// dispatcher, reference src/poisson.cpp:25-82 (including its quirks: NONE zeroes E once; FFT with
// walls leaves E untouched; periodic GS/SOR/NPS use the wall solvers)
void SolvePoisson(std::vector<double>& Ex, std::vector<double>& Ey, const std::vector<double>& rho_q,
                  const int NX, const int NY, const double omega,
                  const PoissonType type, const streaming::BCType bc_type);

void SolvePoisson_GS(const std::vector<double>& rho_q, const int NX, const int NY);
void SolvePoisson_GS_Periodic(const std::vector<double>& rho_q, const int NX, const int NY);
void SolvePoisson_SOR(const std::vector<double>& rho_q, const int NX, const int NY, const double omega);
void SolvePoisson_SOR_Periodic(const std::vector<double>& rho_q, const int NX, const int NY, const double omega);
void SolvePoisson_FFT(const std::vector<double>& rho_q, const int NXf, const int NYf);
void SolvePoisson_9point(const std::vector<double>& rho_q, const int NX, const int NY);
void SolvePoisson_9point_Periodic(const std::vector<double>& rho_q, const int NX, const int NY);

void ComputeElectricField(std::vector<double>& Ex, std::vector<double>& Ey, const int NX, const int NY);
void ComputeElectricField_Periodic(std::vector<double>& Ex, std::vector<double>& Ey, const int NX, const int NY);

void InitPoissonFFT(const int NX, const int NY, const int NY_half, const int real_size);
void FinalizePoissonFFT();

} // namespace poisson
