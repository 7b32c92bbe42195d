// phases.h -- launchers of the per-phase kernels (phases.cu) on AoS device arrays.
#pragma once
#include <cuda_runtime.h>
#include "lbm_consts.h"

namespace plbm {

// Device pointers in the reference's host layout.  feq/geq: [species][0 = self, 1/2 = the two
// cross terms in the reference's order e:(e_i,e_n) i:(i_e,i_n) n:(n_e,n_i)].
struct PhaseArrays {
    double* f[3]; double* g[3]; double* tmp[3];
    double* feq[3][3]; double* geq[3][3];
    double* rho[3]; double* ux[3]; double* uy[3]; double* T[3];
    double* upx[3]; double* upy[3];          // pair velocities e-i, e-n, i-n
    double* Ex; double* Ey; double* rho_q;
};
struct PhaseUnits {
    double cs2, Kb;
    double q[3], m[3];
};

cudaError_t launch_update_macro(const PhaseArrays& a, const PhaseUnits& u, int N, cudaStream_t s);
cudaError_t launch_equilibrium(const PhaseArrays& a, const PhaseUnits& u, int N, cudaStream_t s);
cudaError_t launch_thermal_collisions(const PhaseArrays& a, const PhaseUnits& u, int N, cudaStream_t s);   // writes a.tmp
cudaError_t launch_collisions(const PhaseArrays& a, const PhaseUnits& u, int N, cudaStream_t s);           // writes a.tmp
cudaError_t launch_stream_periodic(const double* const src[3], double* const dst[3], int NX, int NY, cudaStream_t s);
// bounce-back walls: dst must hold the stale contents the reference's temp_* arrays would hold
cudaError_t launch_stream_bounceback(const double* const src[3], double* const dst[3], int NX, int NY, cudaStream_t s);
cudaError_t launch_init_aos(double* const f[3], double* const g[3], int NX, int NY, const double rho_init[3], const double T_init[3], cudaStream_t s);

} // namespace plbm
