// poisson_rows_inv.cu -- P3 of the spectral Poisson solve (see poisson_fft.cu): rows c2r.
#include "poisson_fft_kernels.cuh"

namespace plbm {

// copya / copyb: second destination of the row (a neighbouring slab's copy of my boundary row, in peer memory), or nullptr
struct RowPairOut {
    double* rowa; double* rowb; bool paired; double norm; double* copya; double* copyb;
    static constexpr bool is_smem = false;
    __device__ __forceinline__ void store(int j, cpx z) const
    {
        const double a = __dmul_rn(z.re, norm);                 // poisson.cpp:415-419
        rowa[j] = a;
        if (copya) copya[j] = a;
        if (paired) {
            const double b = __dmul_rn(z.im, norm);
            rowb[j] = b;
            if (copyb) copyb[j] = b;
        }
    }
};

template <int FFT_CAP, int TAIL, int ODD>
__global__ void __launch_bounds__(FFT_CAP, 1)
poisson_rows_inv_kernel(const cpx* __restrict__ T, double* __restrict__ phi, const __grid_constant__ FftPlan plan,
                        int n0, int n1, int nh, double norm, double* first_row_copy, double* last_row_copy)
{
    extern __shared__ cpx fbuf[];
    const int ra = 2 * blockIdx.x, rb = ra + 1;
    const bool paired = rb < n0;
    // packed spectrum of the row pair rebuilt from the two half spectra (c2r contract: the imaginary parts of the
    // DC and Nyquist terms are ignored); every (Ha, Hb) is fetched once and feeds elements k and n1-k
    const FftSmem sm{ fbuf };
    const bool wide = paired && (n0 & 1) == 0;
    #pragma unroll 4
    for (int k = threadIdx.x; k < nh; k += blockDim.x) {
        const cpx* h = T + (size_t)k * n0 + ra;
        cpx Ha, Hb = { 0.0, 0.0 };
        if (wide) load_pair(h, Ha, Hb);
        else { Ha = h[0]; if (paired) Hb = h[1]; }
        const bool self_conj = (k == 0) || (2 * k == n1);
        const double ar = Ha.re, ai = self_conj ? 0.0 : Ha.im;
        const double br = Hb.re, bi = self_conj ? 0.0 : Hb.im;
        sm.store(k, { __dsub_rn(ar, bi), __dadd_rn(ai, br) });
        if (!self_conj) sm.store(n1 - k, { __dadd_rn(ar, bi), __dsub_rn(br, ai) });
    }
    __syncthreads();
    // the slab's first row goes to the lower neighbour as "the row above it", the last row to the upper neighbour
    const int last = n0 - 1;
    const RowPairOut dst{ phi + (size_t)ra * n1, phi + (size_t)rb * n1, paired, norm,
                          ra == 0 ? first_row_copy : (ra == last ? last_row_copy : nullptr),
                          (paired && rb == last) ? last_row_copy : nullptr };
    fft_run<+1, TAIL, ODD>(plan, fbuf, sm, dst);
}

cudaError_t configure_poisson_rows_inv(const PoissonFftDev& p)
{
    return with_shape(p.row, [&](auto CAP, auto TAIL, auto ODD) {
        return allow_smem(poisson_rows_inv_kernel<decltype(CAP)::value, decltype(TAIL)::value, decltype(ODD)::value>);
    });
}

cudaError_t launch_poisson_rows_inv(const PoissonFftDev& p, double* phi, cudaStream_t stream, double* first_row_copy, double* last_row_copy)
{
    const int nh = p.n1 / 2 + 1, t = p.row.threads, grid = (p.nyl + 1) / 2;
    const size_t sm = fft_smem_bytes(p.n1);
    return with_shape(p.row, [&](auto CAP, auto TAIL, auto ODD) {
        poisson_rows_inv_kernel<decltype(CAP)::value, decltype(TAIL)::value, decltype(ODD)::value>
            <<<grid, t, sm, stream>>>(p.T1, phi, p.row, p.nyl, p.n1, nh, p.norm, first_row_copy, last_row_copy);
        return cudaGetLastError();
    });
}

} // namespace plbm
