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

#include <type_traits>
#include <utility>

namespace plbm {

// `in` holds the nyl local rows; T1 is [nh][nyl] (local rows), so the block of spectral columns a
// peer owns is contiguous and can be sent as is.
template <int THREADS, int EPT, bool TWS>
__global__ void __launch_bounds__(THREADS)
poisson_rows_fwd_kernel(const double* __restrict__ in, cpx* __restrict__ T, const __grid_constant__ FftPlan plan,
                        int n0, int n1, int nh)
{
    extern __shared__ cpx fbuf[];
    const cpx* tw = plan.tw;
    if (TWS) {                                     // twiddle table next to the sequence: no global latency inside the stages
        cpx* tws = fbuf + plan.n;
        for (int t = threadIdx.x; t < plan.n; t += blockDim.x) tws[t] = plan.tw[t];
        tw = tws;
    }
    const int ra = 2 * blockIdx.x, rb = ra + 1;        // local row pair; n0 = number of LOCAL rows here
    const bool paired = rb < n0;
    const double* rowa = in + (size_t)ra * n1;
    const double* rowb = in + (size_t)rb * n1;
    for (int j = threadIdx.x; j < n1; j += blockDim.x) {
        cpx z;
        z.re = __ldg(rowa + j);
        z.im = paired ? __ldg(rowb + j) : 0.0;
        fbuf[j] = z;
    }
    fft_smem<-1, EPT, TWS>(fbuf, plan, tw);
    for (int k = threadIdx.x; k < nh; k += blockDim.x) {
        const cpx Z = fbuf[k];
        if (!paired) {
            T[(size_t)k * n0 + ra] = Z;
        } else {
            const cpx Zm = fbuf[k == 0 ? 0 : n1 - k];
            cpx A, B;
            A.re = __dmul_rn(0.5, __dadd_rn(Z.re, Zm.re));
            A.im = __dmul_rn(0.5, __dsub_rn(Z.im, Zm.im));
            B.re = __dmul_rn(0.5, __dadd_rn(Z.im, Zm.im));
            B.im = __dmul_rn(0.5, __dsub_rn(Zm.re, Z.re));
            T[(size_t)k * n0 + ra] = A;
            T[(size_t)k * n0 + rb] = B;
        }
    }
}

// One spectral column (all kx): forward, phi_hat = rho_hat / denom (poisson.cpp:388-409), inverse.
// T2 holds this rank's columns as received from every rank s: [s][k_local][rows of s].
template <int THREADS, int EPT, bool TWS>
__global__ void __launch_bounds__(THREADS)
poisson_cols_kernel(cpx* __restrict__ T, const __grid_constant__ FftPlan plan,
                    const double* __restrict__ sx2, const double* __restrict__ sy2, int n0,
                    const __grid_constant__ SlabTable tab, int k0)
{
    extern __shared__ cpx fbuf[];
    const cpx* tw = plan.tw;
    if (TWS) {                                     // twiddle table next to the sequence: no global latency inside the stages
        cpx* tws = fbuf + plan.n;
        for (int t = threadIdx.x; t < plan.n; t += blockDim.x) tws[t] = plan.tw[t];
        tw = tws;
    }
    const int kl = blockIdx.x;
    for (int sr = 0; sr < tab.nranks; ++sr) {
        const int rows = tab.y0[sr + 1] - tab.y0[sr];
        const cpx* seg = T + (size_t)tab.nkl * tab.y0[sr] + (size_t)kl * rows;
        for (int r = threadIdx.x; r < rows; r += blockDim.x) fbuf[tab.y0[sr] + r] = seg[r];
    }
    fft_smem<-1, EPT, TWS>(fbuf, plan, tw);
    const double syk = __ldg(sy2 + k0 + kl);
    for (int i = threadIdx.x; i < n0; i += blockDim.x) {
        const double denom = __dmul_rn(4.0, __dadd_rn(__ldg(sx2 + i), syk));
        cpx v = fbuf[i];
        if (denom > 1e-15) {
            v.re = __ddiv_rn(v.re, denom);
            v.im = __ddiv_rn(v.im, denom);
        } else {
            v.re = 0.0; v.im = 0.0;
        }
        fbuf[i] = v;
    }
    fft_smem<+1, EPT, TWS>(fbuf, plan, tw);
    for (int sr = 0; sr < tab.nranks; ++sr) {
        const int rows = tab.y0[sr + 1] - tab.y0[sr];
        cpx* seg = T + (size_t)tab.nkl * tab.y0[sr] + (size_t)kl * rows;
        for (int r = threadIdx.x; r < rows; r += blockDim.x) seg[r] = fbuf[tab.y0[sr] + r];
    }
}

template <int THREADS, int EPT, bool TWS>
__global__ void __launch_bounds__(THREADS)
poisson_rows_inv_kernel(const cpx* __restrict__ T, double* __restrict__ phi, const __grid_constant__ FftPlan plan,
                        int n0, int n1, int nh, double norm)
{
    extern __shared__ cpx fbuf[];
    const cpx* tw = plan.tw;
    if (TWS) {                                     // twiddle table next to the sequence: no global latency inside the stages
        cpx* tws = fbuf + plan.n;
        for (int t = threadIdx.x; t < plan.n; t += blockDim.x) tws[t] = plan.tw[t];
        tw = tws;
    }
    const int ra = 2 * blockIdx.x, rb = ra + 1;
    const bool paired = rb < n0;
    for (int k = threadIdx.x; k < nh; k += blockDim.x) {
        const cpx Ha = T[(size_t)k * n0 + ra];
        cpx Hb = { 0.0, 0.0 };
        if (paired) Hb = T[(size_t)k * n0 + rb];
        const bool self_conj = (k == 0) || (2 * k == n1);      // DC / Nyquist: imaginary part ignored (c2r contract)
        const double ar = Ha.re, ai = self_conj ? 0.0 : Ha.im;
        const double br = Hb.re, bi = self_conj ? 0.0 : Hb.im;
        fbuf[k] = { __dsub_rn(ar, bi), __dadd_rn(ai, br) };
        if (!self_conj) fbuf[n1 - k] = { __dadd_rn(ar, bi), __dsub_rn(br, ai) };
    }
    fft_smem<+1, EPT, TWS>(fbuf, plan, tw);
    double* rowa = phi + (size_t)ra * n1;
    double* rowb = phi + (size_t)rb * n1;
    for (int j = threadIdx.x; j < n1; j += blockDim.x) {
        const cpx z = fbuf[j];
        rowa[j] = __dmul_rn(z.re, norm);                        // poisson.cpp:415-419
        if (paired) rowb[j] = __dmul_rn(z.im, norm);
    }
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

// CTA shape per sequence length: THREADS threads, each holding at most EPT elements per stage
// (THREADS * EPT >= n).  Small EPT keeps the butterflies in registers; 64 threads minimum.
struct FftShape { int threads, ept; };
static FftShape fft_shape(int n)
{
    if (n <= 64 * 4) return { 64, 4 };
    if (n <= 128 * 4) return { 128, 4 };
    if (n <= 256 * 4) return { 256, 4 };
    if (n <= 256 * 8) return { 256, 8 };
    if (n <= 512 * 8) return { 512, 8 };
    if (n <= 1024 * 8) return { 1024, 8 };
    return { 1024, 12 };
}

template <class F>
static cudaError_t with_shape(int n, F&& f)
{
    const FftShape s = fft_shape(n);
    if (s.threads == 64) return f(std::integral_constant<int, 64>{}, std::integral_constant<int, 4>{});
    if (s.threads == 128) return f(std::integral_constant<int, 128>{}, std::integral_constant<int, 4>{});
    if (s.threads == 256 && s.ept == 4) return f(std::integral_constant<int, 256>{}, std::integral_constant<int, 4>{});
    if (s.threads == 256) return f(std::integral_constant<int, 256>{}, std::integral_constant<int, 8>{});
    if (s.threads == 512) return f(std::integral_constant<int, 512>{}, std::integral_constant<int, 8>{});
    if (s.ept == 8) return f(std::integral_constant<int, 1024>{}, std::integral_constant<int, 8>{});
    return f(std::integral_constant<int, 1024>{}, std::integral_constant<int, 12>{});
}

// up to this length the twiddle table travels with the sequence in shared memory (2 x 16 B x n <= 128 KB)
static bool tw_in_smem(int n) { (void)n; return false; }   // measured: no gain on B200 (the stages are bound by shared-memory round trips, not twiddle latency)

template <class K>
static cudaError_t allow_smem(K kernel)
{
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(cpx) * FFT_MAX_N));
}

cudaError_t poisson_fft_configure()
{
    cudaError_t e = cudaSuccess;
    auto all = [&](auto T, auto E) {
        constexpr int t = decltype(T)::value, p = decltype(E)::value;
        if (e == cudaSuccess) e = allow_smem(poisson_rows_fwd_kernel<t, p, false>);
        if (e == cudaSuccess) e = allow_smem(poisson_cols_kernel<t, p, false>);
        if (e == cudaSuccess) e = allow_smem(poisson_rows_inv_kernel<t, p, false>);
        if (e == cudaSuccess) e = allow_smem(poisson_rows_fwd_kernel<t, p, true>);
        if (e == cudaSuccess) e = allow_smem(poisson_cols_kernel<t, p, true>);
        if (e == cudaSuccess) e = allow_smem(poisson_rows_inv_kernel<t, p, true>);
        return e;
    };
    for (int n : { 256, 512, 1024, 2048, 4096, 8192, 12288 }) with_shape(n, all);
    return e;
}

cudaError_t launch_poisson_rows_fwd(const PoissonFftDev& p, const double* rho_q, cudaStream_t stream)
{
    const int nh = p.n1 / 2 + 1;
    return with_shape(p.n1, [&](auto T, auto E) {
        constexpr int t = decltype(T)::value, e = decltype(E)::value;
        if (tw_in_smem(p.n1)) poisson_rows_fwd_kernel<t, e, true><<<(p.nyl + 1) / 2, t, 2 * sizeof(cpx) * p.n1, stream>>>(rho_q, p.T1, p.row, p.nyl, p.n1, nh);
        else poisson_rows_fwd_kernel<t, e, false><<<(p.nyl + 1) / 2, t, sizeof(cpx) * p.n1, stream>>>(rho_q, p.T1, p.row, p.nyl, p.n1, nh);
        return cudaGetLastError();
    });
}
cudaError_t launch_poisson_cols(const PoissonFftDev& p, cudaStream_t stream)
{
    if (p.tab.nkl <= 0) return cudaSuccess;
    return with_shape(p.n0, [&](auto T, auto E) {
        constexpr int t = decltype(T)::value, e = decltype(E)::value;
        if (tw_in_smem(p.n0)) poisson_cols_kernel<t, e, true><<<p.tab.nkl, t, 2 * sizeof(cpx) * p.n0, stream>>>(p.T2, p.col, p.sx2, p.sy2, p.n0, p.tab, p.k0);
        else poisson_cols_kernel<t, e, false><<<p.tab.nkl, t, sizeof(cpx) * p.n0, stream>>>(p.T2, p.col, p.sx2, p.sy2, p.n0, p.tab, p.k0);
        return cudaGetLastError();
    });
}
cudaError_t launch_poisson_rows_inv(const PoissonFftDev& p, double* phi, cudaStream_t stream)
{
    const int nh = p.n1 / 2 + 1;
    return with_shape(p.n1, [&](auto T, auto E) {
        constexpr int t = decltype(T)::value, e = decltype(E)::value;
        if (tw_in_smem(p.n1)) poisson_rows_inv_kernel<t, e, true><<<(p.nyl + 1) / 2, t, 2 * sizeof(cpx) * p.n1, stream>>>(p.T1, phi, p.row, p.nyl, p.n1, nh, p.norm);
        else poisson_rows_inv_kernel<t, e, false><<<(p.nyl + 1) / 2, t, sizeof(cpx) * p.n1, stream>>>(p.T1, phi, p.row, p.nyl, p.n1, nh, p.norm);
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
