// k1_walls.cu -- K1 with bounce-back walls (walls.cuh): instantiations and launcher of the kernel in k1_kernel.cuh.
// A separate translation unit so that it compiles beside k1_fused.cu.
#include "k1_kernel.cuh"

namespace plbm {

template <bool M>
static cudaError_t k1_walls_launch(const double* src, double* dst, const double* Ex, const double* Ey, double* rho_q,
                                   const MacroOut& mo, const LbmConsts& c, const LbmGeom& g, const WallArgs& wa, cudaStream_t stream)
{
    static bool configured = false;
    if (!configured) {
        cudaError_t err = cudaFuncSetAttribute(k1_walls_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1_smem_bytes());
        if (err != cudaSuccess) return err;
        // the parked populations are the only on-chip data: give shared memory the carve-out the CTAs per SM need
        if (PLBM_K1_CARVEOUT >= 0) err = cudaFuncSetAttribute(k1_walls_kernel<M>, cudaFuncAttributePreferredSharedMemoryCarveout, PLBM_K1_CARVEOUT);
        if (err != cudaSuccess) return err;
        configured = true;
    }
    dim3 grid((g.NX + K1_THREADS - 1) / K1_THREADS, g.NYl);
    k1_walls_kernel<M><<<grid, K1_THREADS, k1_smem_bytes(), stream>>>(src, dst, Ex, Ey, rho_q, mo, c, g, wa);
    return cudaGetLastError();
}

cudaError_t launch_k1_fused_walls(const double* src, double* dst, const double* Ex, const double* Ey, double* rho_q,
                                  const MacroOut* mo, const LbmConsts& c, const LbmGeom& g, const WallArgs& wa, cudaStream_t stream)
{
    if (mo) return k1_walls_launch<true>(src, dst, Ex, Ey, rho_q, *mo, c, g, wa, stream);
    return k1_walls_launch<false>(src, dst, Ex, Ey, rho_q, MacroOut{}, c, g, wa, stream);
}

} // namespace plbm
