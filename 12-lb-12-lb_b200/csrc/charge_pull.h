// charge_pull.h -- the charge density of the next step from the planes a step has just written (charge_pull.cu).
#pragma once
#include <cuda_runtime.h>
#include "lbm_consts.h"

namespace plbm {

// planes: post-collision populations (halo rows current); q, m: charge and mass of electrons [0] and ions [1]
cudaError_t launch_charge_pull(const double* planes, double* rho_q, const LbmGeom& g, const double q[2], const double m[2], cudaStream_t stream);

} // namespace plbm
