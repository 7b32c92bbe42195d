// poisson_iter.cu -- the reference's iterative Poisson solvers and wall field reconstruction.
//
//   poisson::SolvePoisson_GS       /root/reference/src/poisson.cpp:90-142    red-black Gauss-Seidel
//   poisson::SolvePoisson_SOR      /root/reference/src/poisson.cpp:216-279   red-black SOR
//   poisson::SolvePoisson_9point   /root/reference/src/poisson.cpp:429-483   4-colour 9-point Gauss-Seidel
//   poisson::SolvePoisson_{GS,SOR,9point}_Periodic  /root/reference/src/poisson.cpp:146-211, 283-354, 487-546   same sweeps, all cells, wrapped
//   poisson::ComputeElectricField  /root/reference/src/poisson.cpp:551-585   central differences + Neumann rim
//
// All three solvers sweep the interior 1..N-2 in place with phi = 0 on the rim, warm-started from
// the previous step's phi, and stop after the FIRST iteration whose largest update is below 1e-8
// (or after 5000 iterations).  Cells of one colour are independent, so a colour sweep is
// order-independent and bit-reproducible; to stop at exactly the reference's iteration without a
// host round trip per iteration, the whole solve is ONE cooperative kernel: one grid-wide barrier
// per colour sweep, the iteration's maximum update reduced with warp shuffles and one atomicMax
// per warp on the (non-negative) double's bit pattern before the last colour's barrier.
#include "poisson_iter.h"

#include <cooperative_groups.h>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace plbm {

constexpr int ITER_MAX = 5000;        // poisson.cpp:13
constexpr double ITER_TOL = 1e-8;     // poisson.cpp:14

__device__ __forceinline__ double warp_max(double v)
{
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// KIND 0 = GS, 1 = SOR (5-point, 2 colours), 2 = 9-point (4 colours).  PERIODIC: the *_Periodic variants (all cells, wrapped
// neighbours; poisson.cpp:146-211, 283-354, 487-546), else the Dirichlet ones (interior cells, phi = 0 on the rim).
// Work mapping: rows over CTAs, the cells of the current colour in a row over the CTA's threads (every second cell from the first
// one of that colour): no thread visits a cell of another colour, no division or modulo per cell.
// With an odd periodic extent two cells of one colour are neighbours across the wrap and the sweep is order-dependent -- in the
// reference's own OpenMP loop as well; even extents (the only ones the reference's default lattices have) are bit-reproducible.
constexpr int ITER_THREADS = 128;

template <int KIND, bool PERIODIC>
__global__ void __launch_bounds__(ITER_THREADS)
poisson_iter_kernel(double* phi, const double* __restrict__ rho_q, int NX, int NY, double omega,
                    unsigned long long* __restrict__ err_bits /* [2] ping-pong */, int* __restrict__ iters_out)
{
    cg::grid_group grid = cg::this_grid();
    const bool first_thread = (blockIdx.x == 0 && threadIdx.x == 0);
    constexpr int lo = PERIODIC ? 0 : 1;
    const int hx = PERIODIC ? NX : NX - 1, hy = PERIODIC ? NY : NY - 1;      // exclusive upper bounds of the swept cells
    constexpr int NCOL = (KIND == 2) ? 4 : 2;
    int iter = 0;
    for (; iter < ITER_MAX; ++iter) {
        unsigned long long* slot = err_bits + (iter & 1);
        double local = 0.0;
        for (int colour = 0; colour < NCOL; ++colour) {
            // 2 colours: (i + j) & 1 == colour, every row has cells.  4 colours: 2*(i & 1) + (j & 1) == colour, every second row.
            const int jstep = (KIND == 2) ? 2 : 1;
            const int j0 = (KIND == 2) ? lo + ((lo ^ colour) & 1) : lo;
            for (int j = j0 + jstep * (int)blockIdx.x; j < hy; j += jstep * (int)gridDim.x) {
                const int ipar = (KIND == 2) ? (colour >> 1) : ((colour ^ j) & 1);           // parity of i in this row
                const int i0 = lo + ((lo ^ ipar) & 1);
                const int jn = PERIODIC ? (j + 1 == NY ? 0 : j + 1) : j + 1, js = PERIODIC ? (j == 0 ? NY - 1 : j - 1) : j - 1;
                const double* rown = phi + (size_t)NX * jn;
                const double* rows = phi + (size_t)NX * js;
                double* row = phi + (size_t)NX * j;
                for (int i = i0 + 2 * (int)threadIdx.x; i < hx; i += 2 * ITER_THREADS) {
                    const int ie = PERIODIC ? (i + 1 == NX ? 0 : i + 1) : i + 1, iw = PERIODIC ? (i == 0 ? NX - 1 : i - 1) : i - 1;
                    const double old = row[i];
                    const double rq = rho_q[(size_t)NX * j + i];
                    double nw;
                    if (KIND == 2) {                                                     // poisson.cpp:459-466, 520-531
                        const double so = __dadd_rn(__dadd_rn(__dadd_rn(row[ie], row[iw]), rown[i]), rows[i]);
                        const double sd = __dadd_rn(__dadd_rn(__dadd_rn(rown[ie], rown[iw]), rows[ie]), rows[iw]);
                        const double num = __dadd_rn(__dadd_rn(__dmul_rn(4.0, so), sd), __dmul_rn(6.0, rq));
                        nw = PERIODIC ? __dmul_rn(num, 0.05) : __ddiv_rn(num, 20.0);     // the periodic variant multiplies by 0.05
                    } else {                                                             // poisson.cpp:106-110, 231-241, 163-171, 301-313
                        const double nb = __dadd_rn(__dadd_rn(__dadd_rn(row[ie], row[iw]), rown[i]), rows[i]);
                        const double gs = __dmul_rn(0.25, __dadd_rn(nb, rq));
                        nw = (KIND == 1) ? __dadd_rn(__dmul_rn(__dsub_rn(1.0, omega), old), __dmul_rn(omega, gs)) : gs;
                    }
                    row[i] = nw;
                    local = fmax(local, fabs(__dsub_rn(nw, old)));
                }
            }
            if (colour == NCOL - 1) {                          // the barrier that ends the last colour also completes the reduction
                local = warp_max(local);
                if ((threadIdx.x & 31) == 0) atomicMax(slot, (unsigned long long)__double_as_longlong(local));
            }
            grid.sync();                                       // the next colour (or iteration) reads this colour's updates
            // the other slot was last read after the previous iteration's final barrier: every thread is past that now
            if (colour == 0 && first_thread) err_bits[(iter + 1) & 1] = 0ull;
        }
        const double maxErr = __longlong_as_double((long long)*(volatile unsigned long long*)slot);
        if (maxErr < ITER_TOL) { ++iter; break; }              // poisson.cpp:137-140, 275-277, 479-481
    }
    if (first_thread && iters_out) *iters_out = iter;
}

// ---- the same solvers for lattices that fit ONE thread-block cluster's shared memory ----------------------------------------------
// A solve is thousands of colour sweeps of a few ten thousand cells each, separated by barriers: at the reference's default 200x200
// the grid-wide barrier of the cooperative kernel above (~2.6 us) is the whole cost (26 ms per solve).  Here one cluster of up to 16
// CTAs keeps phi and rho_q in shared memory for the entire solve: every CTA owns a band of rows, reads the two rows next to its band
// from the neighbouring CTAs' shared memory (distributed shared memory), and a sweep ends with the cluster's hardware barrier
// instead of a grid barrier.  The iteration's largest update is reduced per CTA and exchanged through one slot per CTA that every
// CTA reads after the last colour's barrier, so all CTAs stop at the same -- the reference's -- iteration.  Same arithmetic, same
// colour order: bit-identical to the cooperative kernel and to the reference.
constexpr int CL_THREADS = 512;
constexpr int CL_MAX_CTAS = 16;

template <int KIND, bool PERIODIC>
__global__ void __launch_bounds__(CL_THREADS)
poisson_iter_cluster_kernel(double* __restrict__ phi_g, const double* __restrict__ rho_g, int NX, int NY, double omega, int band,
                            int* __restrict__ iters_out)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), me = (int)cluster.block_rank();
    extern __shared__ double cl_smem[];
    double* phi = cl_smem;                                    // [band][NX]: rows me*band .. of the lattice
    double* rho = cl_smem + (size_t)band * NX;
    __shared__ unsigned long long cta_err[2];
    __shared__ double red[CL_THREADS / 32];
    const int r0 = me * band, r1 = min(NY, r0 + band);        // this CTA's rows (possibly none)
    for (int t = threadIdx.x; t < (r1 - r0) * NX; t += CL_THREADS) {
        phi[t] = phi_g[(size_t)r0 * NX + t];
        rho[t] = rho_g[(size_t)r0 * NX + t];
    }
    if (threadIdx.x < 2) cta_err[threadIdx.x] = 0ull;
    cluster.sync();
    // row j of phi wherever it lives in the cluster
    auto row_of = [&](int j) -> const double* {
        const int owner = j / band;
        const double* base = (owner == me) ? phi : cluster.map_shared_rank(phi, owner);
        return base + (size_t)(j - owner * band) * NX;
    };
    constexpr int lo = PERIODIC ? 0 : 1;
    const int hx = PERIODIC ? NX : NX - 1, hy = PERIODIC ? NY : NY - 1;
    constexpr int NCOL = (KIND == 2) ? 4 : 2;
    const int warp = threadIdx.x >> 5, nwarp = CL_THREADS / 32;
    int iter = 0;
    for (; iter < ITER_MAX; ++iter) {
        double local = 0.0;
        for (int colour = 0; colour < NCOL; ++colour) {
            // rows of the band over the warps, the cells of the colour in a row over the lanes
            for (int j = max(r0, lo) + warp; j < min(r1, hy); j += nwarp) {
                if (KIND == 2 && ((j ^ colour) & 1)) continue;                               // 4 colours: rows of the other parity
                const int ipar = (KIND == 2) ? (colour >> 1) : ((colour ^ j) & 1);           // parity of i in this row
                const int i0 = lo + ((lo ^ ipar) & 1);
                const int jn = PERIODIC ? (j + 1 == NY ? 0 : j + 1) : j + 1, js = PERIODIC ? (j == 0 ? NY - 1 : j - 1) : j - 1;
                const double* rown = row_of(jn);
                const double* rows = row_of(js);
                double* row = phi + (size_t)(j - r0) * NX;
                const double* rq_row = rho + (size_t)(j - r0) * NX;
                for (int i = i0 + 2 * (int)(threadIdx.x & 31); i < hx; i += 64) {
                    const int ie = PERIODIC ? (i + 1 == NX ? 0 : i + 1) : i + 1, iw = PERIODIC ? (i == 0 ? NX - 1 : i - 1) : i - 1;
                    const double old = row[i];
                    const double rq = rq_row[i];
                    double nw;
                    if (KIND == 2) {
                        const double so = __dadd_rn(__dadd_rn(__dadd_rn(row[ie], row[iw]), rown[i]), rows[i]);
                        const double sd = __dadd_rn(__dadd_rn(__dadd_rn(rown[ie], rown[iw]), rows[ie]), rows[iw]);
                        const double num = __dadd_rn(__dadd_rn(__dmul_rn(4.0, so), sd), __dmul_rn(6.0, rq));
                        nw = PERIODIC ? __dmul_rn(num, 0.05) : __ddiv_rn(num, 20.0);
                    } else {
                        const double nb = __dadd_rn(__dadd_rn(__dadd_rn(row[ie], row[iw]), rown[i]), rows[i]);
                        const double gs = __dmul_rn(0.25, __dadd_rn(nb, rq));
                        nw = (KIND == 1) ? __dadd_rn(__dmul_rn(__dsub_rn(1.0, omega), old), __dmul_rn(omega, gs)) : gs;
                    }
                    row[i] = nw;
                    local = fmax(local, fabs(__dsub_rn(nw, old)));
                }
            }
            if (colour == NCOL - 1) {                          // this CTA's largest update of the iteration into its slot
                local = warp_max(local);
                if ((threadIdx.x & 31) == 0) red[warp] = local;
                __syncthreads();
                if (threadIdx.x == 0) {
                    double m = 0.0;
                    for (int w = 0; w < nwarp; ++w) m = fmax(m, red[w]);
                    cta_err[iter & 1] = (unsigned long long)__double_as_longlong(m);
                }
            }
            cluster.sync();                                    // the next colour (or iteration) reads this colour's updates
        }
        // every CTA reads every slot: the same maximum, the same decision (slots alternate, two barriers lie between reuse)
        double maxErr = 0.0;
        for (int b = 0; b < C; ++b) {
            const unsigned long long* slot = cluster.map_shared_rank(cta_err, b) + (iter & 1);
            maxErr = fmax(maxErr, __longlong_as_double((long long)*(volatile const unsigned long long*)slot));
        }
        if (maxErr < ITER_TOL) { ++iter; break; }
    }
    cluster.sync();                                            // nobody leaves while a neighbour may still read its rows
    for (int t = threadIdx.x; t < (r1 - r0) * NX; t += CL_THREADS) phi_g[(size_t)r0 * NX + t] = phi[t];
    if (me == 0 && threadIdx.x == 0 && iters_out) *iters_out = iter;
}

// poisson.cpp:556-563: interior central differences
__global__ void efield_interior_kernel(const double* __restrict__ phi, double* __restrict__ Ex, double* __restrict__ Ey, int NX, int NY)
{
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 + blockIdx.y;
    if (i >= NX - 1 || j >= NY - 1) return;
    const size_t c = (size_t)i + (size_t)NX * j;
    Ex[c] = __dmul_rn(-0.5, __dsub_rn(phi[c + 1], phi[c - 1]));
    Ey[c] = __dmul_rn(-0.5, __dsub_rn(phi[c + NX], phi[c - NX]));
}
// poisson.cpp:569-575: rows 0 and NY-1 copy rows 1 and NY-2 (all columns, stale corners included)
__global__ void efield_rim_rows_kernel(double* __restrict__ Ex, double* __restrict__ Ey, int NX, int NY)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NX) return;
    Ex[i] = Ex[(size_t)NX + i];                               Ey[i] = Ey[(size_t)NX + i];
    Ex[(size_t)NX * (NY - 1) + i] = Ex[(size_t)NX * (NY - 2) + i]; Ey[(size_t)NX * (NY - 1) + i] = Ey[(size_t)NX * (NY - 2) + i];
}
// poisson.cpp:578-584: columns 0 and NX-1 copy columns 1 and NX-2 (all rows) -- runs after the rows
__global__ void efield_rim_cols_kernel(double* __restrict__ Ex, double* __restrict__ Ey, int NX, int NY)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= NY) return;
    const size_t r = (size_t)NX * j;
    Ex[r] = Ex[r + 1];                 Ey[r] = Ey[r + 1];
    Ex[r + NX - 1] = Ex[r + NX - 2];   Ey[r + NX - 1] = Ey[r + NX - 2];
}

// The cluster-resident kernel when phi and rho_q fit the shared memory of one cluster (<= 16 CTAs x ~200 KB); cudaErrorNotSupported
// tells the caller to use the cooperative kernel.  PLBM_ITER_CLUSTER=0 disables it (comparison runs).
static cudaError_t launch_poisson_cluster(int kind, bool periodic, double* phi, const double* rho_q, int NX, int NY, double omega,
                                          int* iters_out, cudaStream_t stream)
{
    if (const char* e = std::getenv("PLBM_ITER_CLUSTER")) if (e[0] == '0') return cudaErrorNotSupported;
    // measured at 200x200 on B200 (profiles/r2_summary.md): SOR 18.5 vs 28.8 ms per solve, GS 22.0 vs 30.1 -- but the 4-colour 9-point
    // sweep, with twice the barriers per iteration, is slower through the 16-CTA cluster barrier (42.8 vs 38.6): it keeps the grid kernel
    if (kind == 2) return cudaErrorNotSupported;
    void* fns[2][3] = { { (void*)poisson_iter_cluster_kernel<0, false>, (void*)poisson_iter_cluster_kernel<1, false>, (void*)poisson_iter_cluster_kernel<2, false> },
                        { (void*)poisson_iter_cluster_kernel<0, true>, (void*)poisson_iter_cluster_kernel<1, true>, (void*)poisson_iter_cluster_kernel<2, true> } };
    void* fn = fns[periodic ? 1 : 0][kind];
    const size_t smem_max = 200 * 1024;
    for (int C = CL_MAX_CTAS; C >= 8; C /= 2) {
        const int band = (NY + C - 1) / C;
        const size_t smem = sizeof(double) * 2 * (size_t)band * NX;
        if (smem > smem_max) continue;
        cudaError_t e;
        if ((e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) { cudaGetLastError(); continue; }
        if (C > 8 && (e = cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)) != cudaSuccess) { cudaGetLastError(); continue; }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(C); cfg.blockDim = dim3(CL_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, fn, &cfg) != cudaSuccess || nclusters < 1) { cudaGetLastError(); continue; }
        void* args[] = { &phi, &rho_q, &NX, &NY, &omega, (void*)&band, &iters_out };
        e = cudaLaunchKernelExC(&cfg, fn, args);
        if (e == cudaSuccess) return e;
        cudaGetLastError();
    }
    return cudaErrorNotSupported;
}

cudaError_t launch_poisson_iterative(int kind, bool periodic, double* phi, const double* rho_q, int NX, int NY, double omega,
                                     unsigned long long* err_bits, int* iters_out, cudaStream_t stream)
{
    if (launch_poisson_cluster(kind, periodic, phi, rho_q, NX, NY, omega, iters_out, stream) == cudaSuccess) return cudaSuccess;
    void* fns[2][3] = { { (void*)poisson_iter_kernel<0, false>, (void*)poisson_iter_kernel<1, false>, (void*)poisson_iter_kernel<2, false> },
                        { (void*)poisson_iter_kernel<0, true>, (void*)poisson_iter_kernel<1, true>, (void*)poisson_iter_kernel<2, true> } };
    void* fn = fns[periodic ? 1 : 0][kind];
    int dev = 0, sms = 0, per_sm = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, ITER_THREADS, 0)) != cudaSuccess) return e;
    // one CTA per swept row (per second row for the 4-colour sweep), all resident by construction; few CTAs keep the grid barrier cheap
    const int rows = periodic ? NY : (NY > 2 ? NY - 2 : 0);
    long long want = (kind == 2) ? (rows + 1) / 2 : rows;
    if (want < 1) want = 1;
    const long long cap = (long long)sms * (per_sm > 2 ? 2 : per_sm);
    const int blocks = (int)(want < cap ? want : cap);
    if ((e = cudaMemsetAsync(err_bits, 0, 2 * sizeof(unsigned long long), stream)) != cudaSuccess) return e;
    void* args[] = { &phi, &rho_q, &NX, &NY, &omega, &err_bits, &iters_out };
    return cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(ITER_THREADS), args, 0, stream);
}

cudaError_t launch_efield_walls(const double* phi, double* Ex, double* Ey, int NX, int NY, cudaStream_t stream)
{
    if (NX > 2 && NY > 2) {
        dim3 grid((NX - 2 + 255) / 256, NY - 2);
        efield_interior_kernel<<<grid, 256, 0, stream>>>(phi, Ex, Ey, NX, NY);
    }
    efield_rim_rows_kernel<<<(NX + 255) / 256, 256, 0, stream>>>(Ex, Ey, NX, NY);
    efield_rim_cols_kernel<<<(NY + 255) / 256, 256, 0, stream>>>(Ex, Ey, NX, NY);
    return cudaGetLastError();
}

} // namespace plbm
