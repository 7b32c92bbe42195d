// poisson_cols.cu -- P2 of the spectral Poisson solve (see poisson_fft.cu): columns forward, symbol, inverse.
#include <cstdlib>
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


// ---- P2 over peer memory with the gather off the FFT's critical path ------------------------------------------------------------
// poisson_cols_kernel with peers is one CTA per SM at the long lengths (a column is 100-140 KB of shared memory), and all CTAs of
// a wave wait for their remote loads together, compute together and store together: at 8 GPUs the kernel takes the SUM of its
// NVLink time and its arithmetic (286 us at 8192^2 against ~110 us of arithmetic).  Here a few COPIER CTAs (the lowest block
// indices, so they are scheduled first and never wait for anybody) pull the slabs' shares of this rank's columns into the
// contiguous local buffer T2 = [k_local][n0], group after group, and publish a flag per group; the other CTAs claim columns in the
// same order, wait for the column's flag, transform it out of T2 and write the result straight into the owning slabs' T1 (posted
// stores).  The transfers then run beside the arithmetic instead of between it.  Same arithmetic per column, bit-identical.
struct GatherArgs {
    cpx* T2;               // [nkl][n0]
    unsigned* flags;       // [0],[1]: column claim counters (epoch parity), [2 + g]: epoch at which group g was gathered
    unsigned epoch;        // 1, 2, 3, ... per launch
    int ncopy;             // copier CTAs
    int group;             // columns per group
    int vec32;             // every slab boundary and n0 are even: 256-bit accesses
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// source of element i of spectral column kg in the owning slab's T1
__device__ __forceinline__ const cpx* peer_elem(const SlabTable& tab, const PeerTable& peer, int kg, int i)
{
    int sr = 0;
    while (i >= tab.y0[sr + 1]) ++sr;
    const int rows = tab.y0[sr + 1] - tab.y0[sr];
    return peer.t1[sr] + (size_t)kg * rows + (i - tab.y0[sr]);
}

template <int W>   // W = 1: one complex per access, 2: two (256-bit)
__device__ __forceinline__ void gather_group(const SlabTable& tab, const PeerTable& peer, cpx* T2, int n0, int k0, int kl0, int ncols)
{
    constexpr int U = 8;                                          // accesses in flight per thread
    const int per_col = n0 / W;
    const int total = ncols * per_col;
    for (int base = threadIdx.x; base < total; base += U * blockDim.x) {
        cpx a[U][W];
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = base + u * blockDim.x;
            if (e < total) {
                const int c = e / per_col, i = (e - c * per_col) * W;
                const cpx* src = peer_elem(tab, peer, k0 + kl0 + c, i);
                if constexpr (W == 2) load_pair(src, a[u][0], a[u][1]);
                else { const double2 t = *reinterpret_cast<const double2*>(src); a[u][0] = { t.x, t.y }; }
            }
        }
        #pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = base + u * blockDim.x;
            if (e < total) {
                const int c = e / per_col, i = (e - c * per_col) * W;
                cpx* dst = T2 + (size_t)(kl0 + c) * n0 + i;
                if constexpr (W == 2) store_pair(dst, a[u][0], a[u][1]);
                else *reinterpret_cast<double2*>(dst) = make_double2(a[u][0].re, a[u][0].im);
            }
        }
    }
}

// a gathered column in T2 (read past L1: other SMs wrote it during this kernel); results go to the owning slabs' T1
struct GatheredColumnIO {
    const cpx* local; ColumnIO remote;
    static constexpr bool is_smem = false;
    __device__ __forceinline__ cpx load(int i) const
    {
        const double2 t = __ldcg(reinterpret_cast<const double2*>(local + i));
        return { t.x, t.y };
    }
    __device__ __forceinline__ void store(int i, cpx v) const { remote.store(i, v); }
};

template <int FFT_CAP, int TAIL, int ODD>
__global__ void __launch_bounds__(FFT_CAP, 1)
poisson_cols_gather_kernel(const __grid_constant__ FftPlan plan, const double* __restrict__ sx2, const double* __restrict__ sy2, int n0,
                           const __grid_constant__ SlabTable tab, int k0, const __grid_constant__ PeerTable peer,
                           const __grid_constant__ GatherArgs ga)
{
    extern __shared__ cpx fbuf[];
    __shared__ int s_col;
    const int nkl = tab.nkl;
    unsigned* claim = ga.flags + (ga.epoch & 1u);
    unsigned* ready = ga.flags + 2;
    if ((int)blockIdx.x < ga.ncopy) {
        // ---- copier ----
        if (blockIdx.x == 0 && threadIdx.x == 0) ga.flags[(ga.epoch + 1u) & 1u] = 0u;      // the NEXT launch's claim counter
        const int ngroups = (nkl + ga.group - 1) / ga.group;
        for (int g = blockIdx.x; g < ngroups; g += ga.ncopy) {
            const int kl0 = g * ga.group, ncols = min(ga.group, nkl - kl0);
            if (ga.vec32) gather_group<2>(tab, peer, ga.T2, n0, k0, kl0, ncols);
            else gather_group<1>(tab, peer, ga.T2, n0, k0, kl0, ncols);
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) st_release_gpu(ready + g, ga.epoch);
        }
        return;
    }
    // ---- transform ----
    const FftSmem sm{ fbuf };
    for (;;) {
        __syncthreads();                                  // the previous column's last pass has read the shared buffer; s_col consumed
        if (threadIdx.x == 0) {
            const unsigned t = atomicAdd(claim, 1u);
            int kl = (t < (unsigned)nkl) ? (int)t : -1;
            if (kl >= 0) {
                const unsigned* f = ready + kl / ga.group;
                unsigned spins = 0;
                while (ld_acquire_gpu(f) != ga.epoch) {
                    __nanosleep(64);
                    if (++spins > (1u << 25)) { kl = -2; break; }   // ~seconds: give up rather than hang the device (results are then wrong, the parity gates see it)
                }
            }
            s_col = kl;
        }
        __syncthreads();
        const int kl = s_col;
        if (kl < 0) break;
        const GatheredColumnIO col{ ga.T2 + (size_t)kl * n0, ColumnIO{ nullptr, &tab, &peer, kl, k0 + kl, n0 } };
        const SymbolOut div{ fbuf, sx2, __ldg(sy2 + k0 + kl) };
        fft_run<-1, TAIL, ODD>(plan, fbuf, col, div);
        fft_run<+1, TAIL, ODD>(plan, fbuf, sm, col);
    }
}

cudaError_t configure_poisson_cols(const PoissonFftDev& p)
{
    return with_shape(p.col, [&](auto CAP, auto TAIL, auto ODD) {
        cudaError_t e = allow_smem(poisson_cols_kernel<decltype(CAP)::value, decltype(TAIL)::value, decltype(ODD)::value>);
        if (e != cudaSuccess || p.tab.nranks == 1) return e;
        return allow_smem(poisson_cols_gather_kernel<decltype(CAP)::value, decltype(TAIL)::value, decltype(ODD)::value>);
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


// P2 through peer memory with copier CTAs (poisson_cols_gather_kernel).  p.p2_flags: nkl + 2 zero-initialised words.
cudaError_t launch_poisson_cols_gather(PoissonFftDev& p, cudaStream_t stream, const PeerTable& peer)
{
    if (p.tab.nkl <= 0) return cudaSuccess;
    static int sms = 0, want_copiers = 0;
    if (!sms) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        const char* env = std::getenv("PLBM_P2_COPIERS");
        want_copiers = env ? std::atoi(env) : 16;
        if (want_copiers < 1) want_copiers = 1;
    }
    GatherArgs ga;
    ga.T2 = p.T2;
    ga.flags = p.p2_flags;
    ga.epoch = ++p.p2_epoch;
    // ~256 KB per group: long enough for the copier's loads in flight, short enough that the first columns are ready early
    int group = (int)((256 * 1024) / (sizeof(cpx) * (size_t)p.n0));
    ga.group = group < 1 ? 1 : (group > 16 ? 16 : group);
    const int ngroups = (p.tab.nkl + ga.group - 1) / ga.group;
    ga.ncopy = want_copiers < ngroups ? want_copiers : ngroups;
    if (ga.ncopy > sms - 1) ga.ncopy = sms - 1;
    int nfft = sms - ga.ncopy;
    if (nfft > p.tab.nkl) nfft = p.tab.nkl;
    ga.vec32 = (p.n0 % 2 == 0);
    for (int r = 0; r <= p.tab.nranks; ++r) if (p.tab.y0[r] % 2) ga.vec32 = 0;
    const int t = p.col.threads;
    const size_t sm = fft_smem_bytes(p.n0);
    return with_shape(p.col, [&](auto CAP, auto TAIL, auto ODD) {
        poisson_cols_gather_kernel<decltype(CAP)::value, decltype(TAIL)::value, decltype(ODD)::value>
            <<<ga.ncopy + nfft, t, sm, stream>>>(p.col, p.sx2, p.sy2, p.n0, p.tab, p.k0, peer, ga);
        return cudaGetLastError();
    });
}

} // namespace plbm
