// layout.h -- K4: host-API layout (AoS, reference include/utils.hpp:6-11) <-> device SoA planes.
#pragma once
#include <cuda_runtime.h>
#include "lbm_consts.h"

namespace plbm {

// AoS populations at the top of the time loop -> "post-collision, pre-stream" SoA planes
// (the inverse of the pull K1 performs).  kind 0 = f, 1 = g.
cudaError_t launch_aos_to_soa(const double* aos, double* planes, int species, int kind, const LbmGeom& g, cudaStream_t s);
// SoA planes -> AoS populations at the top of the time loop (applies the pull).
cudaError_t launch_soa_to_aos(const double* planes, double* aos, int species, int kind, const LbmGeom& g, cudaStream_t s);
// Bounce-back walls (walls.cuh): the planes hold either the state at the top of the loop itself (identity) or
// post-collision values that the walls pull turns into that state.
cudaError_t launch_initialize_identity(double* planes, const LbmGeom& g, int NY,
                                       const double rho_init[3], const double T_init[3], const double w[3], cudaStream_t s);
cudaError_t launch_aos_to_soa_identity(const double* aos, double* planes, int species, int kind, const LbmGeom& g, cudaStream_t s);
cudaError_t launch_soa_to_aos_walls(const double* planes, const double* rim, int nrim, double* aos, int species, int kind,
                                    const LbmGeom& g, int identity, cudaStream_t s);
// LBmethod::Initialize (reference src/plasma.cpp:131-158) written directly in plane layout,
// including halo rows.  NY = global rows, y0 = first global row of the slab.
cudaError_t launch_initialize(double* planes, const LbmGeom& g, int NY, int y0,
                              const double rho_init[3], const double T_init[3], const double w[3], cudaStream_t s);
// K5: 18 boundary rows per side (see layout.cu); send_* / recv_* are [18][NX]
cudaError_t launch_halo_pack(const double* planes, double* send_lo, double* send_hi, const LbmGeom& g, cudaStream_t s);
cudaError_t launch_halo_unpack(double* planes, const double* recv_lo, const double* recv_hi, const LbmGeom& g, cudaStream_t s);
cudaError_t launch_fill(double* p, double v, size_t n, cudaStream_t s);

} // namespace plbm
