// poisson_fft.h -- device-side plan of the spectral Poisson solve (see poisson_fft.cu).
#pragma once
#include <cuda_runtime.h>
#include "fft.cuh"

namespace plbm {

constexpr int PLBM_MAX_RANKS = 16;

// Row ranges of all slabs and this rank's share of the spectral columns.
struct SlabTable {
    int nranks;
    int y0[PLBM_MAX_RANKS + 1];   // rank s owns FFT rows [y0[s], y0[s+1])
    int nkl;                      // spectral columns owned by this rank
};

// Peer-memory transposes (several slabs, all GPUs of one node): instead of moving the half spectrum through two
// all-to-alls, the column kernel reads every slab's share of its column straight from that slab's T1 over
// NVLink and writes the result back there.  t1[s] is rank s's T1 = [NY/2+1][rows of s], mapped into this process.
struct PeerTable {
    cpx* t1[PLBM_MAX_RANKS];
};

struct PoissonFftDev {
    int n0, n1;          // n0 = NX "rows" of n1 = NY contiguous values (poisson.cpp:621-622)
    int nyl;             // local rows (n0 for a single slab)
    int k0;              // first spectral column owned by this rank
    SlabTable tab;
    FftPlan row, col;    // length n1 / length n0
    cpx* T1;             // [n1/2+1][nyl]  half spectrum of the local rows, column-major
    cpx* T2;             // [rank s][nkl][rows of s]  this rank's columns after the exchange (== T1 for one slab)
    const double* sx2;   // sin^2(pi*kx/NX) per row index i        (poisson.cpp:391-395)
    const double* sy2;   // sin^2(pi*ky/NY) per column index j
    double norm;         // 1.0 / (NX*NY)                           (poisson.cpp:415)
    unsigned* p2_flags;  // several slabs: nkl + 2 words of poisson_cols_gather_kernel (claim counters, group flags)
    unsigned p2_epoch;   // launches of that kernel so far
    int p2_per_sm;       // its CTAs per SM (0: not queried yet)
};

FftPlan make_fft_plan(int n, const cpx* tw);   // pass schedule for length n (radix schedule of oracle/fft_oracle.c, grouped)
cudaError_t poisson_fft_configure(const PoissonFftDev& p);   // after row/col plans are set
cudaError_t launch_poisson_rows_fwd(const PoissonFftDev& p, const double* rho_q, cudaStream_t stream);   // rho_q -> T1
cudaError_t launch_poisson_cols(const PoissonFftDev& p, cudaStream_t stream, const PeerTable* peer = nullptr);   // T2 in place, or every slab's T1 through peer memory
cudaError_t launch_poisson_cols_gather(PoissonFftDev& p, cudaStream_t stream, const PeerTable& peer);   // the same through peer memory with copier CTAs
cudaError_t launch_poisson_rows_inv(const PoissonFftDev& p, double* phi, cudaStream_t stream,
                                    double* first_row_copy = nullptr, double* last_row_copy = nullptr);           // T1 -> phi
cudaError_t launch_efield_periodic(const double* phi, const double* below, const double* above, double* Ex, double* Ey,
                                   int NX, int NYl, cudaStream_t stream);

} // namespace plbm
