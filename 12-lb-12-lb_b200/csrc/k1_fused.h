// k1_fused.h -- launcher of the fused pull-stream + moments + equilibrium + collision kernel.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "lbm_consts.h"
#include "walls.cuh"

namespace plbm {

// The 12 moment fields LBmethod hands to visualize::UpdateVisualization
// (/root/reference/src/plasma.cpp:516-522); rho_q, Ex, Ey live in their own buffers.
struct MacroOut {
    double* ux[3];
    double* uy[3];
    double* T[3];
    double* rho[3];
};

cudaError_t launch_k1_fused(const double* src, double* dst, const double* Ex, const double* Ey, double* rho_q,
                            const MacroOut* mo, const LbmConsts& c, const LbmGeom& g, cudaStream_t stream);

// same step with E = -grad(phi) formed inside the kernel (periodic central differences, bit-identical to K3);
// below/above: boundary rows of phi owned by the neighbouring slabs, nullptr on a single slab
cudaError_t launch_k1_fused_phi(const double* src, double* dst, const double* phi, const double* below, const double* above, double* rho_q,
                                const MacroOut* mo, const LbmConsts& c, const LbmGeom& g, cudaStream_t stream);

// Periodic lattices, persistent warps with a TMA-filled pool of stash buffers (k1_kernel.cuh: k1_pool_kernel).  `tmap` describes
// the planes of `src` as the 4-D tensor [6][9][NYl + 2][NX]; make_k1_tensor_map builds it (once per population buffer).  E comes
// from phi when `phi` is given (below/above as in launch_k1_fused_phi), else from the Ex/Ey arrays.
cudaError_t make_k1_tensor_map(CUtensorMap* tmap, const double* planes, const LbmGeom& g, bool pool);   // pool: boxes of k1_pool_kernel, else of k1_tma_kernel
// one CTA per 64-cell tile as launch_k1_fused_phi, the pull done by the TMA engine (k1_tma_kernel)
cudaError_t launch_k1_tma(const CUtensorMap& tmap, const double* src, double* dst, const double* Ex, const double* Ey,
                          const double* phi, const double* below, const double* above, double* rho_q,
                          const MacroOut* mo, const LbmConsts& c, const LbmGeom& g, cudaStream_t stream);
cudaError_t launch_k1_pool(const CUtensorMap& tmap, const double* src, double* dst, const double* Ex, const double* Ey,
                           const double* phi, const double* below, const double* above, double* rho_q,
                           const MacroOut* mo, const LbmConsts& c, const LbmGeom& g, cudaStream_t stream);
int k1_prefetch_rows(int NX, int NYl);   // LbmGeom::prefetch_rows for a slab of this shape on the current device

// bounce-back walls (single slab): the pull follows walls.cuh, E comes from the arrays
cudaError_t launch_k1_fused_walls(const double* src, double* dst, const double* Ex, const double* Ey, double* rho_q,
                                  const MacroOut* mo, const LbmConsts& c, const LbmGeom& g, const WallArgs& wa, cudaStream_t stream);

} // namespace plbm
