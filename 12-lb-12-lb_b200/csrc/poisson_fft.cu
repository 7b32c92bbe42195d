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
// The three passes are separate translation units (poisson_rows_fwd.cu, poisson_cols.cu, poisson_rows_inv.cu);
// this file holds the plan builder, the configuration entry point and K3.
#include "poisson_fft_kernels.cuh"
#include "host_tables.h"

namespace plbm {

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
            if (R == 3) P.n3++;
            if (P.odd == FFT_ODD_NONE || P.odd == small) P.odd = small;
            else if (small != FFT_ODD_GENERIC && (P.odd == FFT_ODD_3 || P.odd == FFT_ODD_5 || P.odd == FFT_ODD_35)) P.odd = FFT_ODD_35;
            else P.odd = FFT_ODD_GENERIC;
        }
        nsub /= R;
        ps.m = nsub;
        s *= R;
        while ((1 << (log2s + 1)) <= s) ++log2s;
    }
    const int per_thread = (P.odd == FFT_ODD_3 || P.odd == FFT_ODD_5 || P.odd == FFT_ODD_35) ? 15 : FFT_EPT;
    P.threads = ((n + per_thread - 1) / per_thread + 31) & ~31;
    if (P.threads < 64) P.threads = 64;
    if (P.threads > 512 && P.odd != FFT_ODD_NONE) {               // the 768-thread kernels carry the generic odd pass only
        P.odd = FFT_ODD_GENERIC;
        P.threads = ((n + FFT_EPT - 1) / FFT_EPT + 31) & ~31;
    }
    return P;
}

cudaError_t poisson_fft_configure(const PoissonFftDev& p)
{
    cudaError_t e = configure_poisson_rows_fwd(p);
    if (e == cudaSuccess) e = configure_poisson_cols(p);
    if (e == cudaSuccess) e = configure_poisson_rows_inv(p);
    return e;
}

cudaError_t launch_efield_periodic(const double* phi, const double* below, const double* above, double* Ex, double* Ey,
                                   int NX, int NYl, cudaStream_t stream)
{
    dim3 grid((NX + 255) / 256, NYl);
    efield_periodic_kernel<<<grid, 256, 0, stream>>>(phi, below, above, Ex, Ey, NX, NYl);
    return cudaGetLastError();
}

} // namespace plbm
