// k1_fused.cu -- K1 for periodic lattices: instantiations and launchers of the kernel in k1_kernel.cuh.
#include <cstdlib>
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
        // the parked populations are the only on-chip data: give shared memory the carve-out the CTAs per SM need
        if (PLBM_K1_CARVEOUT >= 0) err = cudaFuncSetAttribute(k1_fused_kernel<M, P>, cudaFuncAttributePreferredSharedMemoryCarveout, PLBM_K1_CARVEOUT);
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

template <bool M, bool P>
static cudaError_t k1_pool_launch(const CUtensorMap& tmap, const double* src, double* dst, const double* a, const double* b, const double* e,
                                  double* rho_q, const MacroOut& mo, const LbmConsts& c, const LbmGeom& g, cudaStream_t stream)
{
    static int sms = 0;
    if (!sms) {
        cudaError_t err = cudaFuncSetAttribute(k1_pool_kernel<M, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1_pool_smem_bytes());
        if (err != cudaSuccess) return err;
        int dev = 0;
        if ((err = cudaGetDevice(&dev)) != cudaSuccess) return err;
        if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return err;
    }
    // persistent warps: one CTA of POOL_WARPS warps per SM, or fewer when the slab has fewer tiles
    const long long tiles = (long long)((g.NX + POOL_TILE - 1) / POOL_TILE) * g.NYl;
    const int ctas = (int)((tiles + POOL_WARPS - 1) / POOL_WARPS < sms ? (tiles + POOL_WARPS - 1) / POOL_WARPS : sms);
    k1_pool_kernel<M, P><<<ctas, POOL_THREADS, k1_pool_smem_bytes(), stream>>>(tmap, src, dst, a, b, e, rho_q, mo, c, g);
    return cudaGetLastError();
}

template <bool M, bool P>
static cudaError_t k1_tma_launch(const CUtensorMap& tmap, const double* src, double* dst, const double* a, const double* b, const double* e,
                                 double* rho_q, const MacroOut& mo, const LbmConsts& c, const LbmGeom& g, cudaStream_t stream)
{
    static bool configured = false;
    if (!configured) {
        cudaError_t err = cudaFuncSetAttribute(k1_tma_kernel<M, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1_tma_smem_bytes());
        if (err != cudaSuccess) return err;
        configured = true;
    }
    dim3 grid((g.NX + K1_THREADS - 1) / K1_THREADS, g.NYl);
    k1_tma_kernel<M, P><<<grid, K1_THREADS, k1_tma_smem_bytes(), stream>>>(tmap, src, dst, a, b, e, rho_q, mo, c, g);
    return cudaGetLastError();
}

cudaError_t launch_k1_tma(const CUtensorMap& tmap, const double* src, double* dst, const double* Ex, const double* Ey,
                          const double* phi, const double* below, const double* above, double* rho_q,
                          const MacroOut* mo, const LbmConsts& c, const LbmGeom& g, cudaStream_t stream)
{
    if (phi) {
        if (mo) return k1_tma_launch<true, true>(tmap, src, dst, phi, below, above, rho_q, *mo, c, g, stream);
        return k1_tma_launch<false, true>(tmap, src, dst, phi, below, above, rho_q, MacroOut{}, c, g, stream);
    }
    if (mo) return k1_tma_launch<true, false>(tmap, src, dst, Ex, Ey, nullptr, rho_q, *mo, c, g, stream);
    return k1_tma_launch<false, false>(tmap, src, dst, Ex, Ey, nullptr, rho_q, MacroOut{}, c, g, stream);
}

cudaError_t launch_k1_pool(const CUtensorMap& tmap, const double* src, double* dst, const double* Ex, const double* Ey,
                           const double* phi, const double* below, const double* above, double* rho_q,
                           const MacroOut* mo, const LbmConsts& c, const LbmGeom& g, cudaStream_t stream)
{
    if (phi) {
        if (mo) return k1_pool_launch<true, true>(tmap, src, dst, phi, below, above, rho_q, *mo, c, g, stream);
        return k1_pool_launch<false, true>(tmap, src, dst, phi, below, above, rho_q, MacroOut{}, c, g, stream);
    }
    if (mo) return k1_pool_launch<true, false>(tmap, src, dst, Ex, Ey, nullptr, rho_q, *mo, c, g, stream);
    return k1_pool_launch<false, false>(tmap, src, dst, Ex, Ey, nullptr, rho_q, MacroOut{}, c, g, stream);
}

// Rows between a CTA and the row it asks L2 for (k1_kernel.cuh, PLBM_K1_PREFETCH).  Measured at 2048^2 (profiles/r2_k1_sweeps.md):
// about 1/7 of a wave of resident CTAs ahead in launch order (blockIdx.x fastest) is best (-1.5 %); a whole wave ahead is slower
// than no prefetch (+3.5 %).  0 when the slab has fewer CTAs than one wave.  PLBM_K1_PREFETCH_ROWS overrides (tuning).
int k1_prefetch_rows(int NX, int NYl)
{
    if (const char* e = std::getenv("PLBM_K1_PREFETCH_ROWS")) return std::atoi(e);
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k1_fused_kernel<false, true>, K1_THREADS, k1_smem_bytes()) != cudaSuccess) return 0;
    const int per_row = (NX + K1_THREADS - 1) / K1_THREADS;
    const int wave = sms * per_sm;
    if ((long long)per_row * NYl <= wave) return 0;
    const int rows = (wave / 7 + per_row - 1) / per_row;
    return rows < 1 ? 1 : rows;
}

// The population planes as the 4-D tensor [sk 6][direction 9][storage row NYl+2][x NX] of doubles; box = one row segment of a
// 32-cell tile plus the two cells its +-1 pull offsets need, for the six distributions of one direction (k1_pool_kernel).  cuTensorMapEncodeTiled is taken from the driver at run time (no link-time libcuda).
cudaError_t make_k1_tensor_map(CUtensorMap* tmap, const double* planes, const LbmGeom& g, bool pool)
{
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (err != cudaSuccess) return err;
        if (!fn || q != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
        encode = (encode_fn)fn;
    }
    const cuuint64_t dims[4] = { (cuuint64_t)g.NX, (cuuint64_t)g.NYl + 2, (cuuint64_t)NQ, (cuuint64_t)(2 * NSPEC) };
    const cuuint64_t strides[3] = { (cuuint64_t)g.pitch * sizeof(double), (cuuint64_t)g.plane * sizeof(double),
                                    (cuuint64_t)g.plane * NQ * sizeof(double) };
    const cuuint32_t box[4] = { (cuuint32_t)(pool ? POOL_BOX_W : TMA_BOX_W), 1u, 1u, (cuuint32_t)(2 * NSPEC) };
    const cuuint32_t estr[4] = { 1u, 1u, 1u, 1u };
    int promo = (int)CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    if (const char* e = std::getenv("PLBM_TMA_L2PROMO")) promo = std::atoi(e);     // tuning: 0 none, 1 64 B, 2 128 B, 3 256 B
    const CUresult r = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(planes), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

} // namespace plbm
