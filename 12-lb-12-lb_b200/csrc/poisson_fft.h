// poisson_fft.h -- device-side plan of the spectral Poisson solve (see poisson_fft.cu).
#pragma once
#include <cuda_runtime.h>
#include "fft.cuh"

namespace plbm {

struct PoissonFftDev {
    int n0, n1;          // n0 = NX "rows" of n1 = NY contiguous values (poisson.cpp:621-622)
    FftPlan row, col;    // length n1 / length n0
    cpx* T;              // (n1/2+1) x n0 transposed half spectrum
    const double* sx2;   // sin^2(pi*kx/NX) per row index i        (poisson.cpp:391-395)
    const double* sy2;   // sin^2(pi*ky/NY) per column index j
    double norm;         // 1.0 / (NX*NY)                           (poisson.cpp:415)
};

cudaError_t poisson_fft_configure();
cudaError_t launch_poisson_fft(const PoissonFftDev& p, const double* rho_q, double* phi, cudaStream_t stream);
cudaError_t launch_efield_periodic(const double* phi, double* Ex, double* Ey, int NX, int NY, cudaStream_t stream);

} // namespace plbm
