// output.h -- device side of the alternate output path: what visualize::UpdateVisualization
// derives from the 15 FP64 fields before its first OpenCV call, formed on the GPU.
#pragma once
#include <cuda_runtime.h>

namespace plbm {

constexpr int NUM_FRAMES = 12;   // rho_e, rho_i, rho_q, ux_e, uy_e, |u_e|, ux_i, uy_i, |u_i|, T_e, T_i, T_n
constexpr int NUM_SERIES = 19;   // ux_e, uy_e, |u_e|, ux_i, uy_i, |u_i|, ux_n, uy_n, |u_n|, T_e, T_i, T_n, rho_e, rho_i, rho_n, rho_q, Ex, Ey, |E|
constexpr int NUM_POINTS = 9;

struct FrameFields {             // device, NX*NYl doubles each
    const double *rho_e, *rho_i, *rho_q, *ux_e, *uy_e, *ux_i, *uy_i, *T_e, *T_i, *T_n;
};
struct SeriesFields {
    const double *ux[3], *uy[3], *T[3], *rho[3], *rho_q, *Ex, *Ey;
};

// frames: NUM_FRAMES consecutive planes of n floats
cudaError_t launch_frames(const FrameFields& f, float* frames, size_t n, cudaStream_t stream);
// series: [NUM_SERIES][NUM_POINTS] doubles; points outside rows [y0, y0 + NYl) are left untouched
cudaError_t launch_series(const SeriesFields& f, double* series, int NX, int NY, int y0, int NYl, cudaStream_t stream);

} // namespace plbm
