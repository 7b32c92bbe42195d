// poisson_iter.h -- iterative Poisson solvers and wall field reconstruction (poisson_iter.cu).
#pragma once
#include <cuda_runtime.h>

namespace plbm {

// kind 0 = Gauss-Seidel, 1 = SOR, 2 = 9-point; periodic: the *_Periodic variants (all cells, wrapped neighbours) instead of the
// Dirichlet ones.  err_bits: 2 x u64 device scratch; iters_out: device int (may be null).
cudaError_t launch_poisson_iterative(int kind, bool periodic, double* phi, const double* rho_q, int NX, int NY, double omega,
                                     unsigned long long* err_bits, int* iters_out, cudaStream_t stream);
cudaError_t launch_efield_walls(const double* phi, double* Ex, double* Ey, int NX, int NY, cudaStream_t stream);

} // namespace plbm
