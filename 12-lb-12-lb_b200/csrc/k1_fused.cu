// k1_fused.cu -- K1 for periodic lattices: instantiations and launchers of the kernel in k1_kernel.cuh.
#include "k1_kernel.cuh"

namespace plbm {

template <bool M, bool P>
static cudaError_t k1_launch(const double* src, double* dst, const double* a, const double* b, const double* e, double* rho_q,
                             const MacroOut& mo, const LbmConsts& c, const LbmGeom& g, cudaStream_t stream)
{
    static bool configured = false;
    if (!configured) {
        cudaError_t err = cudaFuncSetAttribute(k1_fused_kernel<M, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1_smem_bytes());
        if (err != cudaSuccess) return err;
        configured = true;
    }
    dim3 grid((g.NX + K1_THREADS - 1) / K1_THREADS, g.NYl);
    k1_fused_kernel<M, P><<<grid, K1_THREADS, k1_smem_bytes(), stream>>>(src, dst, a, b, e, rho_q, mo, c, g);
    return cudaGetLastError();
}

cudaError_t launch_k1_fused(const double* src, double* dst, const double* Ex, const double* Ey, double* rho_q,
                            const MacroOut* mo, const LbmConsts& c, const LbmGeom& g, cudaStream_t stream)
{
    if (mo) return k1_launch<true, false>(src, dst, Ex, Ey, nullptr, rho_q, *mo, c, g, stream);
    return k1_launch<false, false>(src, dst, Ex, Ey, nullptr, rho_q, MacroOut{}, c, g, stream);
}

cudaError_t launch_k1_fused_phi(const double* src, double* dst, const double* phi, const double* below, const double* above, double* rho_q,
                                const MacroOut* mo, const LbmConsts& c, const LbmGeom& g, cudaStream_t stream)
{
    if (mo) return k1_launch<true, true>(src, dst, phi, below, above, rho_q, *mo, c, g, stream);
    return k1_launch<false, true>(src, dst, phi, below, above, rho_q, MacroOut{}, c, g, stream);
}

} // namespace plbm
