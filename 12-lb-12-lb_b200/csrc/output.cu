// output.cu -- alternate output path: the CV_32F matrices and the sample-point series of the visualiser.
//
//   visualize::VisualizationDensity / Velocity / Temperature   /root/reference/src/visualize.cpp:226-315
//       mat.at<float>(y, x) = static_cast<float>(field[INDEX(x, y, NX)]);   |u| = float(sqrt(ux*ux + uy*uy)) in double
//   visualize::UpdateVisualization, "Record into buffers"      /root/reference/src/visualize.cpp:168-219
//       sample points of InitVisualization (:72-80): centre +- NX/4, NY/4
//
// The reference copies 120 B per cell and step to the host only to narrow 12 of the values to float; doing
// the narrowing here leaves 48 B per cell to cross PCIe.  Same operations, same rounding (cvt.rn.f32.f64,
// sqrt.rn.f64, no FMA contraction), so the matrices are bit-identical to the ones visualize.cpp builds.
#include "output.h"

namespace plbm {

__device__ __forceinline__ float magnitude(double a, double b)
{
    return (float)__dsqrt_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)));
}

__global__ void __launch_bounds__(256) frames_kernel(const FrameFields f, float* __restrict__ out, size_t n)
{
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double uxe = f.ux_e[i], uye = f.uy_e[i], uxi = f.ux_i[i], uyi = f.uy_i[i];
    out[0 * n + i] = (float)f.rho_e[i];
    out[1 * n + i] = (float)f.rho_i[i];
    out[2 * n + i] = (float)f.rho_q[i];
    out[3 * n + i] = (float)uxe;
    out[4 * n + i] = (float)uye;
    out[5 * n + i] = magnitude(uxe, uye);
    out[6 * n + i] = (float)uxi;
    out[7 * n + i] = (float)uyi;
    out[8 * n + i] = magnitude(uxi, uyi);
    out[9 * n + i] = (float)f.T_e[i];
    out[10 * n + i] = (float)f.T_i[i];
    out[11 * n + i] = (float)f.T_n[i];
}

__global__ void series_kernel(const SeriesFields f, double* __restrict__ out, int NX, int NY, int y0, int NYl)
{
    const int p = threadIdx.x;
    if (p >= NUM_POINTS) return;
    const int cx = NX / 2, cy = NY / 2, dx = NX / 4, dy = NY / 4;          // visualize.cpp:72-80
    const int sx[NUM_POINTS] = { 0, 1, -1, 0, 0, 1, 1, -1, -1 }, sy[NUM_POINTS] = { 0, 0, 0, 1, -1, 1, -1, 1, -1 };
    const int i = cx + sx[p] * dx, j = cy + sy[p] * dy;
    if (j < y0 || j >= y0 + NYl || i < 0 || i >= NX) return;
    const size_t idx = (size_t)i + (size_t)NX * (j - y0);
    auto put = [&](int q, double v) { out[q * NUM_POINTS + p] = v; };
    auto mag = [](double a, double b) { return __dsqrt_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b))); };
    for (int s = 0; s < 3; ++s) {
        const double ux = f.ux[s][idx], uy = f.uy[s][idx];
        put(3 * s, ux); put(3 * s + 1, uy); put(3 * s + 2, mag(ux, uy));
        put(9 + s, f.T[s][idx]);
        put(12 + s, f.rho[s][idx]);
    }
    const double ex = f.Ex[idx], ey = f.Ey[idx];
    put(15, f.rho_q[idx]); put(16, ex); put(17, ey); put(18, mag(ex, ey));
}

cudaError_t launch_frames(const FrameFields& f, float* frames, size_t n, cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    frames_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(f, frames, n);
    return cudaGetLastError();
}

cudaError_t launch_series(const SeriesFields& f, double* series, int NX, int NY, int y0, int NYl, cudaStream_t stream)
{
    series_kernel<<<1, 32, 0, stream>>>(f, series, NX, NY, y0, NYl);
    return cudaGetLastError();
}

} // namespace plbm
