// poisson_rows_fwd.cu -- P1 of the spectral Poisson solve (see poisson_fft.cu): rows r2c.
#include "poisson_fft_kernels.cuh"

namespace plbm {

// two real rows as one complex sequence, read straight from global memory by the first pass
struct RowPairIn {
    const double* rowa; const double* rowb; bool paired;
    static constexpr bool is_smem = false;
    __device__ __forceinline__ cpx load(int j) const
    {
        cpx z;
        z.re = __ldg(rowa + j);
        z.im = paired ? __ldg(rowb + j) : 0.0;
        return z;
    }
};

// `in` holds the nyl local rows; T1 is [nh][nyl] (local rows), so the block of spectral columns a
// peer owns is contiguous and can be sent as is.
template <int FFT_CAP, int TAIL, int ODD>
__global__ void __launch_bounds__(FFT_CAP, 1)
poisson_rows_fwd_kernel(const double* __restrict__ in, cpx* __restrict__ T, const __grid_constant__ FftPlan plan,
                        int n0, int n1, int nh)
{
    extern __shared__ cpx fbuf[];
    const int ra = 2 * blockIdx.x, rb = ra + 1;        // local row pair; n0 = number of LOCAL rows here
    const bool paired = rb < n0;
    const RowPairIn src{ in + (size_t)ra * n1, in + (size_t)rb * n1, paired };
    const FftSmem sm{ fbuf };
    fft_run<-1, TAIL, ODD>(plan, fbuf, src, sm);
    if (!paired) {
        for (int k = threadIdx.x; k < nh; k += blockDim.x) T[(size_t)k * n0 + ra] = sm.load(k);
        return;
    }
    const bool wide = (n0 & 1) == 0;                   // (k*n0 + ra) even: the pair is 32-byte aligned
    #pragma unroll 4
    for (int k = threadIdx.x; k < nh; k += blockDim.x) {
        const cpx Z = sm.load(k);
        const cpx Zm = sm.load(k == 0 ? 0 : n1 - k);
        cpx A, B;
        A.re = __dmul_rn(0.5, __dadd_rn(Z.re, Zm.re));
        A.im = __dmul_rn(0.5, __dsub_rn(Z.im, Zm.im));
        B.re = __dmul_rn(0.5, __dadd_rn(Z.im, Zm.im));
        B.im = __dmul_rn(0.5, __dsub_rn(Zm.re, Z.re));
        cpx* dst = T + (size_t)k * n0 + ra;
        if (wide) store_pair(dst, A, B);
        else { dst[0] = A; dst[1] = B; }
    }
}

cudaError_t configure_poisson_rows_fwd(const PoissonFftDev& p)
{
    return with_shape(p.row, [&](auto CAP, auto TAIL, auto ODD) {
        return allow_smem(poisson_rows_fwd_kernel<decltype(CAP)::value, decltype(TAIL)::value, decltype(ODD)::value>);
    });
}

cudaError_t launch_poisson_rows_fwd(const PoissonFftDev& p, const double* rho_q, cudaStream_t stream)
{
    const int nh = p.n1 / 2 + 1, t = p.row.threads, grid = (p.nyl + 1) / 2;
    const size_t sm = fft_smem_bytes(p.n1);
    return with_shape(p.row, [&](auto CAP, auto TAIL, auto ODD) {
        poisson_rows_fwd_kernel<decltype(CAP)::value, decltype(TAIL)::value, decltype(ODD)::value>
            <<<grid, t, sm, stream>>>(rho_q, p.T1, p.row, p.nyl, p.n1, nh);
        return cudaGetLastError();
    });
}

} // namespace plbm
