// charge_pull.cu -- rho_q of the NEXT time step straight from the populations a step has just written.
//
// The charge density of step t (LBmethod::UpdateMacro, /root/reference/src/plasma.cpp:352-372, 452-453) is a function of the
// streamed populations alone: rho_s = sum_i f_s,i(x - c_i) in the order i = 0..8, thresholded at 1e-10, then
// q_i*rho_i/m_i + q_e*rho_e/m_e, zeroed below 1e-15.  K1 forms it on its way, but it exists as soon as the previous step has
// written its post-collision planes (and the halo rows have arrived) -- one collision earlier than K1 delivers it.  This
// kernel pulls the 18 f-planes of electrons and ions (144 B per cell) and forms exactly that value, same operations in the same
// order with IEEE divisions, so that the Poisson solve of step t can run BESIDE K1 of step t instead of after it
// (plbm_step_peer, DESIGN.md "Poisson off the critical path").
#include "charge_pull.h"

namespace plbm {

__global__ void __launch_bounds__(128)
charge_pull_kernel(const double* __restrict__ src, double* __restrict__ rho_q, const LbmGeom g,
                   double q_e, double q_i, double m_e, double m_i)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= g.NX) return;
    const int xm = (x == 0) ? g.NX - 1 : x - 1;
    const int xp = (x == g.NX - 1) ? 0 : x + 1;
    int rm = y, rp = y + 2;                            // storage rows of y-1 and y+1 (storage row = y + 1)
    if (g.wrap_y) {
        if (y == 0) rm = g.NYl;
        if (y == g.NYl - 1) rp = 1;
    }
    const int r0o = (y + 1) * g.pitch, rmo = rm * g.pitch, rpo = rp * g.pitch;
    const int off[NQ] = { r0o + x, r0o + xm, rmo + x, r0o + xp, rpo + x, rmo + xm, rmo + xp, rpo + xp, rpo + xm };
    double rho[2];
    #pragma unroll
    for (int s = 0; s < 2; ++s) {
        const double* p = src + (long long)((s * 2 + 0) * NQ) * g.plane;      // f planes of species s
        double v[NQ];
        #pragma unroll
        for (int i = 0; i < NQ; ++i) v[i] = __ldg(p + i * g.plane + off[i]);
        double r = v[0];                                                      // 0.0 + f_0 = f_0
        #pragma unroll
        for (int i = 1; i < NQ; ++i) r = __dadd_rn(r, v[i]);                  // plasma.cpp:352-372
        rho[s] = (r < 1e-10) ? 0.0 : r;                                       // plasma.cpp:373-377, 393-397
    }
    double rq = __dadd_rn(__ddiv_rn(__dmul_rn(q_i, rho[1]), m_i), __ddiv_rn(__dmul_rn(q_e, rho[0]), m_e));   // plasma.cpp:452
    if (rq < 1e-15) rq = 0.0;                                                 // plasma.cpp:453
    rho_q[(long long)y * g.NX + x] = rq;
}

cudaError_t launch_charge_pull(const double* planes, double* rho_q, const LbmGeom& g, const double q[2], const double m[2], cudaStream_t stream)
{
    dim3 grid((g.NX + 127) / 128, g.NYl);
    charge_pull_kernel<<<grid, 128, 0, stream>>>(planes, rho_q, g, q[0], q[1], m[0], m[1]);
    return cudaGetLastError();
}

} // namespace plbm
