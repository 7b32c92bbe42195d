// poisson_cols.cu -- P2 of the spectral Poisson solve (see poisson_fft.cu): columns forward, symbol, inverse.
#include "poisson_fft_kernels.cuh"

namespace plbm {

// element i (row index of the whole lattice) of spectral column kl: in T2 = [rank s][k_local][rows of s] after
// an all-to-all, or (peer != nullptr) in place in the owning slab's T1 = [k][rows of s], reached through peer memory
struct ColumnIO {
    cpx* T; const SlabTable* tab; const PeerTable* peer; int kl; int kg; int n0;
    static constexpr bool is_smem = false;
    __device__ __forceinline__ cpx* at(int i) const
    {
        if (tab->nranks == 1) return T + (size_t)kl * n0 + i;
        int sr = 0;
        while (i >= tab->y0[sr + 1]) ++sr;
        const int rows = tab->y0[sr + 1] - tab->y0[sr];
        if (peer) return peer->t1[sr] + (size_t)kg * rows + (i - tab->y0[sr]);
        return T + (size_t)tab->nkl * tab->y0[sr] + (size_t)kl * rows + (i - tab->y0[sr]);
    }
    __device__ __forceinline__ cpx load(int i) const
    {
        const double2 t = *reinterpret_cast<const double2*>(at(i));
        return { t.x, t.y };
    }
    __device__ __forceinline__ void store(int i, cpx v) const { *reinterpret_cast<double2*>(at(i)) = make_double2(v.re, v.im); }
};

// last pass of the forward column transform: phi_hat = rho_hat / denom (poisson.cpp:388-409), left in shared memory
struct SymbolOut {
    cpx* buf; const double* sx2; double syk;
    static constexpr bool is_smem = true;
    __device__ __forceinline__ void store(int i, cpx v) const
    {
        const double denom = __dmul_rn(4.0, __dadd_rn(__ldg(sx2 + i), syk));
        if (denom > 1e-15) {
            v.re = __ddiv_rn(v.re, denom);
            v.im = __ddiv_rn(v.im, denom);
        } else {
            v.re = 0.0; v.im = 0.0;
        }
        buf[fft_slot(i)] = v;
    }
};

// One spectral column (all kx): forward, division by the symbol, inverse -- the column never leaves the SM.
template <int FFT_CAP, int TAIL, int ODD>
__global__ void __launch_bounds__(FFT_CAP, 1)
poisson_cols_kernel(cpx* T, const __grid_constant__ FftPlan plan,
                    const double* __restrict__ sx2, const double* __restrict__ sy2, int n0,
                    const __grid_constant__ SlabTable tab, int k0, const __grid_constant__ PeerTable peer, int use_peer)
{
    extern __shared__ cpx fbuf[];
    const int kl = blockIdx.x;
    const ColumnIO col{ T, &tab, use_peer ? &peer : nullptr, kl, k0 + kl, n0 };
    const SymbolOut div{ fbuf, sx2, __ldg(sy2 + k0 + kl) };
    const FftSmem sm{ fbuf };
    fft_run<-1, TAIL, ODD>(plan, fbuf, col, div);
    fft_run<+1, TAIL, ODD>(plan, fbuf, sm, col);
}

cudaError_t configure_poisson_cols(const PoissonFftDev& p)
{
    return with_shape(p.col, [&](auto CAP, auto TAIL, auto ODD) {
        return allow_smem(poisson_cols_kernel<decltype(CAP)::value, decltype(TAIL)::value, decltype(ODD)::value>);
    });
}

cudaError_t launch_poisson_cols(const PoissonFftDev& p, cudaStream_t stream, const PeerTable* peer)
{
    if (p.tab.nkl <= 0) return cudaSuccess;
    const PeerTable none = {};
    const PeerTable& pt = peer ? *peer : none;
    const int use_peer = peer != nullptr;
    const int t = p.col.threads;
    const size_t sm = fft_smem_bytes(p.n0);
    return with_shape(p.col, [&](auto CAP, auto TAIL, auto ODD) {
        poisson_cols_kernel<decltype(CAP)::value, decltype(TAIL)::value, decltype(ODD)::value>
            <<<p.tab.nkl, t, sm, stream>>>(p.T2, p.col, p.sx2, p.sy2, p.n0, p.tab, p.k0, pt, use_peer);
        return cudaGetLastError();
    });
}

} // namespace plbm
