// poisson_fft.cu -- spectral Poisson solve and field reconstruction on the device.
//
//   poisson::SolvePoisson_FFT               /root/reference/src/poisson.cpp:365-420
//   poisson::InitPoissonFFT                 /root/reference/src/poisson.cpp:611-623
//   poisson::ComputeElectricField_Periodic  /root/reference/src/poisson.cpp:589-607
//
// The reference hands the flat x-fastest rho_q array to fftw_plan_dft_r2c_2d(NX, NY): n0 = NX rows
// of n1 = NY contiguous reals (only a true 2-D transform when NX == NY; the quirk is kept).  Three
// passes, each one CTA per sequence with the whole sequence in shared memory:
//   P1  rows r2c   two real rows per complex transform, Hermitian split, written TRANSPOSED
//                  T[k][r] so that P2 is contiguous (and so that slabs can be exchanged by an
//                  all-to-all in the multi-GPU path)
//   P2  columns    forward transform, division by the discrete-Laplacian symbol
//                  4(sin^2(pi kx/NX) + sin^2(pi ky/NY)) (K2), inverse transform -- one kernel,
//                  the column never leaves the SM
//   P3  rows c2r   rebuild the packed spectrum of a row pair, inverse transform, scale by 1/(NX*NY)
// then K3: E = -grad phi by wrapped central differences.
// Algorithmic traffic: P1 8+8, P2 8+8, P3 8+8, K3 8+16 = 72 B per cell.
#include "poisson_fft.h"
#include "fft.cuh"
#include "host_tables.h"

#include <type_traits>


namespace plbm {

// FFT_CAP = the largest CTA the instantiation may be launched with (sets the register budget:
// 512 threads -> 128 registers, 768 -> 85 with spills; only sequences longer than 8192 need the latter).

// (A, B) of a row pair are neighbours in T[k][row]: one 256-bit access when the pair is 32-byte aligned
__device__ __forceinline__ void store_pair(cpx* p, cpx a, cpx b)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(a.re), "d"(a.im), "d"(b.re), "d"(b.im) : "memory");
}
__device__ __forceinline__ void load_pair(const cpx* p, cpx& a, cpx& b)
{
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a.re), "=d"(a.im), "=d"(b.re), "=d"(b.im) : "l"(p));
}

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

// K3, poisson.cpp:589-607.  NY = local rows; below/above = the neighbouring slabs' boundary rows of
// phi (nullptr: single slab, wrap inside the array).
__global__ void efield_periodic_kernel(const double* __restrict__ phi, const double* __restrict__ below, const double* __restrict__ above,
                                       double* __restrict__ Ex, double* __restrict__ Ey, int NX, int NY)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= NX) return;
    const int im1 = (i == 0) ? NX - 1 : i - 1, ip1 = (i == NX - 1) ? 0 : i + 1;
    const size_t row = (size_t)j * NX;
    const double* rm = (j > 0) ? phi + row - NX : (below ? below : phi + (size_t)(NY - 1) * NX);
    const double* rp = (j < NY - 1) ? phi + row + NX : (above ? above : phi);
    Ex[row + i] = __dmul_rn(-0.5, __dsub_rn(__ldg(phi + row + ip1), __ldg(phi + row + im1)));
    Ey[row + i] = __dmul_rn(-0.5, __dsub_rn(__ldg(rp + i), __ldg(rm + i)));
}

FftPlan make_fft_plan(int n, const cpx* tw)
{
    FftPlan P = {};
    P.n = n; P.tw = tw;
    int radix[64];
    const int ns = host_factorize(n, radix);
    int s = 1, nsub = n, log2s = 0;
    for (int i = 0; i < ns;) {
        FftPass& ps = P.pass[P.npass++];
        ps.s = s; ps.log2s = log2s;
        int R;
        if (radix[i] == 4 && i + 1 < ns && radix[i + 1] == 4) { ps.kind = FFT_PASS_44; R = 16; i += 2; P.n44++; }
        else if (radix[i] == 4 && i + 1 < ns && radix[i + 1] == 2) { ps.kind = FFT_PASS_42; R = 8; i += 2; P.tail = FFT_TAIL_42; }
        else if (radix[i] == 4) { ps.kind = FFT_PASS_4; R = 4; i += 1; P.tail = FFT_TAIL_4; }
        else if (radix[i] == 2) { ps.kind = FFT_PASS_2; R = 2; i += 1; P.tail = FFT_TAIL_2; }
        else {
            ps.kind = FFT_PASS_ODD; R = radix[i]; ps.r = (unsigned short)R; i += 1;
            const int small = (R == 3 || R == 5) ? R : FFT_ODD_GENERIC;
            P.odd = (P.odd == FFT_ODD_NONE || P.odd == small) ? small : FFT_ODD_GENERIC;
        }
        nsub /= R;
        ps.m = nsub;
        s *= R;
        while ((1 << (log2s + 1)) <= s) ++log2s;
    }
    const int per_thread = (P.odd == FFT_ODD_3 || P.odd == FFT_ODD_5) ? 15 : FFT_EPT;
    P.threads = ((n + per_thread - 1) / per_thread + 31) & ~31;
    if (P.threads < 64) P.threads = 64;
    if (P.threads > 512 && P.odd != FFT_ODD_NONE) {               // the 768-thread kernels carry the generic odd pass only
        P.odd = FFT_ODD_GENERIC;
        P.threads = ((n + FFT_EPT - 1) / FFT_EPT + 31) & ~31;
    }
    return P;
}

static size_t fft_smem_bytes(int n) { return sizeof(cpx) * (size_t)fft_smem_elems(n); }

// kernels are specialised on the plan's shape (tail pass kind, odd passes present) and on the CTA cap
template <class F>
static cudaError_t with_shape(const FftPlan& P, F&& f)
{
    auto odd = [&](auto TAIL) {
        if (P.threads > 512) {
            if (P.odd != FFT_ODD_NONE) return f(std::integral_constant<int, 768>{}, TAIL, std::integral_constant<int, FFT_ODD_GENERIC>{});
            return f(std::integral_constant<int, 768>{}, TAIL, std::integral_constant<int, FFT_ODD_NONE>{});
        }
        switch (P.odd) {
        case FFT_ODD_3: return f(std::integral_constant<int, 512>{}, TAIL, std::integral_constant<int, FFT_ODD_3>{});
        case FFT_ODD_5: return f(std::integral_constant<int, 512>{}, TAIL, std::integral_constant<int, FFT_ODD_5>{});
        case FFT_ODD_GENERIC: return f(std::integral_constant<int, 512>{}, TAIL, std::integral_constant<int, FFT_ODD_GENERIC>{});
        default: return f(std::integral_constant<int, 512>{}, TAIL, std::integral_constant<int, FFT_ODD_NONE>{});
        }
    };
    switch (P.tail) {
    case FFT_TAIL_42: return odd(std::integral_constant<int, FFT_TAIL_42>{});
    case FFT_TAIL_4: return odd(std::integral_constant<int, FFT_TAIL_4>{});
    case FFT_TAIL_2: return odd(std::integral_constant<int, FFT_TAIL_2>{});
    default: return odd(std::integral_constant<int, FFT_TAIL_NONE>{});
    }
}

template <class K>
static cudaError_t allow_smem(K kernel)
{
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fft_smem_bytes(FFT_MAX_N));
}

cudaError_t poisson_fft_configure(const PoissonFftDev& p)
{
    cudaError_t e = with_shape(p.row, [&](auto CAP, auto TAIL, auto ODD) {
        cudaError_t r = allow_smem(poisson_rows_fwd_kernel<decltype(CAP)::value, decltype(TAIL)::value, decltype(ODD)::value>);
        if (r == cudaSuccess) r = allow_smem(poisson_rows_inv_kernel<decltype(CAP)::value, decltype(TAIL)::value, decltype(ODD)::value>);
        return r;
    });
    if (e != cudaSuccess) return e;
    return with_shape(p.col, [&](auto CAP, auto TAIL, auto ODD) {
        return allow_smem(poisson_cols_kernel<decltype(CAP)::value, decltype(TAIL)::value, decltype(ODD)::value>);
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

cudaError_t launch_efield_periodic(const double* phi, const double* below, const double* above, double* Ex, double* Ey,
                                   int NX, int NYl, cudaStream_t stream)
{
    dim3 grid((NX + 255) / 256, NYl);
    efield_periodic_kernel<<<grid, 256, 0, stream>>>(phi, below, above, Ex, Ey, NX, NYl);
    return cudaGetLastError();
}

} // namespace plbm
