// layout.cu -- K4 (AoS <-> SoA) and the initial condition.  See layout.h.
#include "layout.h"
#include "walls.cuh"

namespace plbm {

__constant__ int c_cx[NQ] = { 0, 1, 0, -1, 0, 1, -1, -1, 1 };   // reference src/plasma.cpp:10-11
__constant__ int c_cy[NQ] = { 0, 0, 1, 0, -1, 1, 1, -1, -1 };

// storage offset inside one plane of the cell that direction i of cell (x, y) is pulled from
__device__ __forceinline__ long long pull_offset(int x, int y, int i, const LbmGeom& g)
{
    int xs = x - c_cx[i];
    if (xs < 0) xs += g.NX; else if (xs >= g.NX) xs -= g.NX;
    int ys = y - c_cy[i];
    if (g.wrap_y) { if (ys < 0) ys += g.NYl; else if (ys >= g.NYl) ys -= g.NYl; }
    return (long long)(ys + 1) * g.pitch + xs;
}

__global__ void aos_to_soa_kernel(const double* __restrict__ aos, double* __restrict__ planes, int sk, LbmGeom g)
{
    const long long n = (long long)g.NX * g.NYl * NQ;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % NQ);
        const long long c = t / NQ;
        const int x = (int)(c % g.NX), y = (int)(c / g.NX);
        planes[(long long)(sk * NQ + i) * g.plane + pull_offset(x, y, i, g)] = aos[t];
    }
}

__global__ void soa_to_aos_kernel(const double* __restrict__ planes, double* __restrict__ aos, int sk, LbmGeom g)
{
    const long long n = (long long)g.NX * g.NYl * NQ;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % NQ);
        const long long c = t / NQ;
        const int x = (int)(c % g.NX), y = (int)(c / g.NX);
        aos[t] = planes[(long long)(sk * NQ + i) * g.plane + pull_offset(x, y, i, g)];
    }
}

// walls: the planes hold the state at the top of the loop itself (identity) ...
__global__ void aos_to_soa_identity_kernel(const double* __restrict__ aos, double* __restrict__ planes, int sk, LbmGeom g)
{
    const long long n = (long long)g.NX * g.NYl * NQ;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t % NQ);
        const long long c = t / NQ;
        const int x = (int)(c % g.NX), y = (int)(c / g.NX);
        planes[(long long)(sk * NQ + i) * g.plane + (long long)(y + 1) * g.pitch + x] = aos[t];
    }
}
// ... or post-collision values that still have to be streamed against the walls (walls.cuh): sk = species*2 + kind
__global__ void soa_to_aos_walls_kernel(const double* __restrict__ planes, const double* __restrict__ rim, int nrim,
                                        double* __restrict__ aos, int sk, LbmGeom g, int identity)
{
    const long long n = (long long)g.NX * g.NYl * NQ;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(t % NQ);
        const long long c = t / NQ;
        const int x = (int)(c % g.NX), y = (int)(c / g.NX);
        const long long own = (long long)(y + 1) * g.pitch + x;
        double v;
        if (identity) v = planes[(long long)(sk * NQ + j) * g.plane + own];
        else {
            const WallSource ws = wall_source(x, y, j, g.NX, g.NYl);
            if (ws.dir >= 0) v = planes[(long long)(sk * NQ + ws.dir) * g.plane + (long long)(ws.y + 1) * g.pitch + ws.x];
            else if (sk & 1) v = planes[(long long)((sk - 1) * NQ + j) * g.plane + own];                      // g: post-collision f
            else v = rim[((long long)((sk >> 1) * NQ + j)) * nrim + wall_rim_index(x, y, g.NX, g.NYl)];       // f: its pre-collision value
        }
        aos[t] = v;
    }
}

struct InitParams {
    double rho_init[3], T_init[3], w[3];
};

// LBmethod::Initialize (plasma.cpp:131-158) at the cells themselves (walls: the first step does not stream)
__global__ void initialize_identity_kernel(double* __restrict__ planes, LbmGeom g, int NY, InitParams p)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= g.NX || y >= g.NYl) return;
    const bool inside = (x >= g.NX / 4 + 1) && (x < 3 * g.NX / 4) && (y >= NY / 4 + 1) && (y < 3 * NY / 4);
    const long long own = (long long)(y + 1) * g.pitch + x;
    for (int s = 0; s < 3; ++s)
        for (int i = 0; i < NQ; ++i) {
            const double w = p.w[i == 0 ? 0 : (i < 5 ? 1 : 2)];
            const bool on = (s == 2) || inside;
            planes[(long long)((s * 2 + 0) * NQ + i) * g.plane + own] = on ? __dmul_rn(w, p.rho_init[s]) : 0.0;
            planes[(long long)((s * 2 + 1) * NQ + i) * g.plane + own] = on ? __dmul_rn(w, p.T_init[s]) : 0.0;
        }
}

// plane value at storage position (xs, row rr) of direction i = initial f_i at the cell that
// pulls from there, i.e. (xs + cx_i, ys + cy_i) with periodic wrap on the GLOBAL lattice.
__global__ void initialize_kernel(double* __restrict__ planes, LbmGeom g, int NY, int y0, InitParams p)
{
    const int xs = blockIdx.x * blockDim.x + threadIdx.x;
    const int rr = blockIdx.y;                    // storage row 0 .. NYl+1
    if (xs >= g.NX) return;
    const int NX = g.NX;
    int ysg = y0 + rr - 1;                        // global row of this storage row
    ysg = ((ysg % NY) + NY) % NY;
    #pragma unroll
    for (int i = 0; i < NQ; ++i) {
        int x = xs + c_cx[i]; if (x < 0) x += NX; else if (x >= NX) x -= NX;
        int y = ysg + c_cy[i]; if (y < 0) y += NY; else if (y >= NY) y -= NY;
        const bool inside = (x >= NX / 4 + 1) && (x < 3 * NX / 4) && (y >= NY / 4 + 1) && (y < 3 * NY / 4);
        const double w = p.w[i == 0 ? 0 : (i < 5 ? 1 : 2)];
        const long long o = (long long)rr * g.pitch + xs;
        #pragma unroll
        for (int s = 0; s < 3; ++s) {
            const bool on = (s == 2) || inside;   // neutrals everywhere, charged species in the block
            planes[(long long)((s * 2 + 0) * NQ + i) * g.plane + o] = on ? __dmul_rn(w, p.rho_init[s]) : 0.0;
            planes[(long long)((s * 2 + 1) * NQ + i) * g.plane + o] = on ? __dmul_rn(w, p.T_init[s]) : 0.0;
        }
    }
}

// K5: boundary rows of the planes that leave the slab.  Directions with cy = +1 (2,5,6) of the top
// local row travel to the upper neighbour, directions with cy = -1 (4,7,8) of the bottom local row
// to the lower neighbour; 6 (species,kind) x 3 directions = 18 rows of NX doubles per side.
__constant__ int c_dir_up[3] = { 2, 5, 6 };
__constant__ int c_dir_down[3] = { 4, 7, 8 };

__global__ void halo_pack_kernel(const double* __restrict__ planes, double* __restrict__ send_lo, double* __restrict__ send_hi, LbmGeom g)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;                 // 0..17: (species*2+kind)*3 + j
    if (x >= g.NX) return;
    const int sk = r / 3, j = r % 3;
    send_hi[(size_t)r * g.NX + x] = planes[(long long)(sk * NQ + c_dir_up[j]) * g.plane + (long long)g.NYl * g.pitch + x];
    send_lo[(size_t)r * g.NX + x] = planes[(long long)(sk * NQ + c_dir_down[j]) * g.plane + (long long)1 * g.pitch + x];
}

// what the lower neighbour sent up lands in halo row 0, what the upper neighbour sent down in row NYl+1
__global__ void halo_unpack_kernel(double* __restrict__ planes, const double* __restrict__ recv_lo, const double* __restrict__ recv_hi, LbmGeom g)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (x >= g.NX) return;
    const int sk = r / 3, j = r % 3;
    planes[(long long)(sk * NQ + c_dir_up[j]) * g.plane + x] = recv_lo[(size_t)r * g.NX + x];
    planes[(long long)(sk * NQ + c_dir_down[j]) * g.plane + (long long)(g.NYl + 1) * g.pitch + x] = recv_hi[(size_t)r * g.NX + x];
}

__global__ void fill_kernel(double* p, double v, size_t n)
{
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) p[t] = v;
}

static int grid_for(long long n, int threads)
{
    long long b = (n + threads - 1) / threads;
    const long long cap = 148LL * 32;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

cudaError_t launch_aos_to_soa(const double* aos, double* planes, int species, int kind, const LbmGeom& g, cudaStream_t s)
{
    aos_to_soa_kernel<<<grid_for((long long)g.NX * g.NYl * NQ, 256), 256, 0, s>>>(aos, planes, species * 2 + kind, g);
    return cudaGetLastError();
}
cudaError_t launch_soa_to_aos(const double* planes, double* aos, int species, int kind, const LbmGeom& g, cudaStream_t s)
{
    soa_to_aos_kernel<<<grid_for((long long)g.NX * g.NYl * NQ, 256), 256, 0, s>>>(planes, aos, species * 2 + kind, g);
    return cudaGetLastError();
}
cudaError_t launch_initialize_identity(double* planes, const LbmGeom& g, int NY,
                                       const double rho_init[3], const double T_init[3], const double w[3], cudaStream_t s)
{
    InitParams p;
    for (int k = 0; k < 3; ++k) { p.rho_init[k] = rho_init[k]; p.T_init[k] = T_init[k]; p.w[k] = w[k]; }
    dim3 grid((g.NX + 127) / 128, g.NYl);
    initialize_identity_kernel<<<grid, 128, 0, s>>>(planes, g, NY, p);
    return cudaGetLastError();
}
cudaError_t launch_aos_to_soa_identity(const double* aos, double* planes, int species, int kind, const LbmGeom& g, cudaStream_t s)
{
    aos_to_soa_identity_kernel<<<grid_for((long long)g.NX * g.NYl * NQ, 256), 256, 0, s>>>(aos, planes, species * 2 + kind, g);
    return cudaGetLastError();
}
cudaError_t launch_soa_to_aos_walls(const double* planes, const double* rim, int nrim, double* aos, int species, int kind,
                                    const LbmGeom& g, int identity, cudaStream_t s)
{
    soa_to_aos_walls_kernel<<<grid_for((long long)g.NX * g.NYl * NQ, 256), 256, 0, s>>>(planes, rim, nrim, aos, species * 2 + kind, g, identity);
    return cudaGetLastError();
}
cudaError_t launch_initialize(double* planes, const LbmGeom& g, int NY, int y0,
                              const double rho_init[3], const double T_init[3], const double w[3], cudaStream_t s)
{
    InitParams p;
    for (int k = 0; k < 3; ++k) { p.rho_init[k] = rho_init[k]; p.T_init[k] = T_init[k]; p.w[k] = w[k]; }
    dim3 grid((g.NX + 127) / 128, g.NYl + 2);
    initialize_kernel<<<grid, 128, 0, s>>>(planes, g, NY, y0, p);
    return cudaGetLastError();
}
cudaError_t launch_halo_pack(const double* planes, double* send_lo, double* send_hi, const LbmGeom& g, cudaStream_t s)
{
    dim3 grid((g.NX + 255) / 256, 18);
    halo_pack_kernel<<<grid, 256, 0, s>>>(planes, send_lo, send_hi, g);
    return cudaGetLastError();
}
cudaError_t launch_halo_unpack(double* planes, const double* recv_lo, const double* recv_hi, const LbmGeom& g, cudaStream_t s)
{
    dim3 grid((g.NX + 255) / 256, 18);
    halo_unpack_kernel<<<grid, 256, 0, s>>>(planes, recv_lo, recv_hi, g);
    return cudaGetLastError();
}
cudaError_t launch_fill(double* p, double v, size_t n, cudaStream_t s)
{
    fill_kernel<<<grid_for((long long)n, 256), 256, 0, s>>>(p, v, n);
    return cudaGetLastError();
}

} // namespace plbm
