// k1_fused.cu -- K1: one pass per time step over all three species, f and g.
//
// Replaces five full-grid sweeps of the reference with one kernel:
//   streaming::StreamingPeriodic / ThermalStreamingPeriodic   src/streaming.cpp:35-59,117-141   (as the PULL at the top)
//   LBmethod::UpdateMacro                                     src/plasma.cpp:317-456
//   LBmethod::ComputeEquilibrium                              src/plasma.cpp:162-308            (never materialised)
//   collisions::ThermalCollisions, collisions::Collisions     src/collisions.cpp:64-181
//
// State layout (HBM): 54 SoA planes of FP64, plane p = (species*2 + kind)*9 + dir, each
// [NYl + 2 halo rows][pitch]; buffer `src` holds the POST-collision populations of the previous
// step at their own cell, so streaming is the shifted read f_i(x) <- src_i(x - c_i).  The kernel
// writes post-collision values of this step to `dst` at its own cell (aligned, coalesced), plus
// rho_q for the Poisson solve and, on request, the 12 moment fields the visualiser consumes.
// Algorithmic traffic: 54*8 read + 54*8 written + Ex,Ey 16 + rho_q 8 = 888 B per cell and step.
//
// One thread owns one cell.  The 54 pulled populations are parked in shared memory (one column per
// thread, conflict-free) while the moments are formed, then re-read direction by direction, which
// keeps the register file for the ~3 k FP64 operations per cell ("K1 budget").
#include "k1_fused.h"
#include "lbm_cell.cuh"

namespace plbm {

constexpr int K1_THREADS = 128;

template <int I> struct DirInfo {
    static constexpr int wclass = (I == 0) ? 0 : (I < 5 ? 1 : 2);
};

struct K1Cell {
    CellMacro m;
    VelSet self[3], pair[3];
    D AB2[3][3];
    D rhoh[3], u2[3];
    D uE[2];
    D Ex, Ey;
};

// collide one direction I for all species; bs/bp: equilibrium brackets for the 3 self and 3 pair velocities
template <int I>
__device__ __forceinline__ void k1_direction(const K1Cell& cell, const D (&bs)[3], const D (&bp)[3], const D (&guo)[2],
                                             const D (&wr)[3], const D (&wT)[3], const D (&pref)[2],
                                             const double* stash, double* __restrict__ dst, long long plane, long long cell_off,
                                             const LbmConsts& c)
{
    static_for<3>([&](auto S) {
        constexpr int s = decltype(S)::value;
        constexpr int p0 = PAIR_SLOT[s][0], p1 = PAIR_SLOT[s][1];
        const D b[3] = { bs[s], bp[p0], bp[p1] };
        const D fv = D(stash[((s * 2 + 0) * NQ + I) * K1_THREADS]);
        const D gv = D(stash[((s * 2 + 1) * NQ + I) * K1_THREADS]);
        D force = D(0.0);
        if constexpr (s < 2) force = pref[s] * guo[s];                        // collisions.cpp:154-163
        D fnew, gnew;
        collide_species_dir<s>(fv, gv, b, wr[s], wT[s], cell.AB2[s], cell.rhoh[s], cell.u2[s], force, c, fnew, gnew);
        dst[((s * 2 + 0) * NQ + I) * plane + cell_off] = fnew.v;
        dst[((s * 2 + 1) * NQ + I) * plane + cell_off] = gnew.v;
    });
}

template <bool WRITE_MACRO>
__global__ void __launch_bounds__(K1_THREADS, 3)
k1_fused_kernel(const double* __restrict__ src, double* __restrict__ dst,
                const double* __restrict__ Exf, const double* __restrict__ Eyf,
                double* __restrict__ rho_q, const MacroOut mo,
                const __grid_constant__ LbmConsts c, const __grid_constant__ LbmGeom g)
{
    extern __shared__ double stash_all[];
    double* stash = stash_all + threadIdx.x;          // column of this thread: stash[p * K1_THREADS]
    const int x = blockIdx.x * K1_THREADS + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= g.NX) return;

    // pull offsets inside one plane: source cell (x - cx_i, y - cy_i), periodic in x; in y either
    // periodic (single slab) or through the halo rows (storage row = y + 1)
    const int xm = (x == 0) ? g.NX - 1 : x - 1;
    const int xp = (x == g.NX - 1) ? 0 : x + 1;
    int rm = y, rp = y + 2;                            // storage rows of y-1 and y+1
    if (g.wrap_y) {
        if (y == 0) rm = g.NYl;
        if (y == g.NYl - 1) rp = 1;
    }
    const int r0o = (y + 1) * g.pitch, rmo = rm * g.pitch, rpo = rp * g.pitch;
    const int off[NQ] = { r0o + x, r0o + xm, rmo + x, r0o + xp, rpo + x, rmo + xm, rmo + xp, rpo + xp, rpo + xm };

    K1Cell cell;
    D rl[3], mx[3], my[3], tl[3];
    static_for<3>([&](auto S) {
        constexpr int s = decltype(S)::value;
        D f[NQ], gg[NQ];
        const double* pf = src + (long long)((s * 2 + 0) * NQ) * g.plane;
        const double* pg = src + (long long)((s * 2 + 1) * NQ) * g.plane;
        #pragma unroll
        for (int i = 0; i < NQ; ++i) f[i] = D(__ldg(pf + i * g.plane + off[i]));
        #pragma unroll
        for (int i = 0; i < NQ; ++i) gg[i] = D(__ldg(pg + i * g.plane + off[i]));
        #pragma unroll
        for (int i = 0; i < NQ; ++i) {
            stash[((s * 2 + 0) * NQ + i) * K1_THREADS] = f[i].v;
            stash[((s * 2 + 1) * NQ + i) * K1_THREADS] = gg[i].v;
        }
        rl[s] = sum9(f); mx[s] = moment_x(f); my[s] = moment_y(f); tl[s] = sum9(gg);
    });

    const long long cidx = (long long)y * g.NX + x;    // scalar fields are flat x + NX*y
    cell.Ex = D(__ldg(Exf + cidx));
    cell.Ey = D(__ldg(Eyf + cidx));
    cell_update_macro(rl, mx, my, tl, cell.Ex, cell.Ey, c, cell.m);
    rho_q[cidx] = cell.m.rho_q.v;
    if constexpr (WRITE_MACRO) {
        static_for<3>([&](auto S) {
            constexpr int s = decltype(S)::value;
            mo.ux[s][cidx] = cell.m.ux[s].v; mo.uy[s][cidx] = cell.m.uy[s].v;
            mo.T[s][cidx] = cell.m.T[s].v;   mo.rho[s][cidx] = cell.m.rho[s].v;
        });
    }

    static_for<3>([&](auto S) {
        constexpr int s = decltype(S)::value;
        cell.self[s] = make_velset(cell.m.ux[s], cell.m.uy[s], c);
        cell.pair[s] = make_velset(cell.m.upx[s], cell.m.upy[s], c);
        thermal_cell_terms<s>(cell.m.rho[s], c, cell.AB2[s]);
        cell.rhoh[s] = D(0.5) * cell.m.rho[s];
        cell.u2[s] = cell.m.ux[s] * cell.m.ux[s] + cell.m.uy[s] * cell.m.uy[s];               // collisions.cpp:98-100
    });
    static_for<2>([&](auto S) {
        constexpr int s = decltype(S)::value;
        cell.uE[s] = cell.m.ux[s] * cell.Ex + cell.m.uy[s] * cell.Ey;                         // collisions.cpp:157,162
    });

    const long long cell_off = (long long)r0o + x;

    // ---- rest direction (weight 4/9) -------------------------------------------------------
    {
        D bs[3], bp[3], guo[2], wr[3], wT[3], pref[2];
        static_for<3>([&](auto S) {
            constexpr int s = decltype(S)::value;
            bs[s] = eq_bracket_rest(cell.self[s].K);
            bp[s] = eq_bracket_rest(cell.pair[s].K);
            wr[s] = D(c.w[0]) * cell.m.rho[s];
            wT[s] = D(c.w[0]) * cell.m.T[s];
        });
        static_for<2>([&](auto S) {
            constexpr int s = decltype(S)::value;
            pref[s] = guo_prefactor<s>(0, cell.m.rho[s], c);
            guo[s] = D(0.0) - cell.uE[s];
        });
        k1_direction<0>(cell, bs, bp, guo, wr, wT, pref, stash, dst, g.plane, cell_off, c);
    }
    // ---- the four axes: directions (1,3), (2,4) have weight 1/9; (5,7), (6,8) weight 1/36 ----
    static_for<2>([&](auto WC) {
        constexpr int wclass = decltype(WC)::value + 1;
        D wr[3], wT[3], pref[2];
        static_for<3>([&](auto S) {
            constexpr int s = decltype(S)::value;
            wr[s] = D(c.w[wclass]) * cell.m.rho[s];
            wT[s] = D(c.w[wclass]) * cell.m.T[s];
        });
        static_for<2>([&](auto S) {
            constexpr int s = decltype(S)::value;
            pref[s] = guo_prefactor<s>(wclass, cell.m.rho[s], c);
        });
        static_for<2>([&](auto AX) {
            constexpr int axis = (wclass - 1) * 2 + decltype(AX)::value;
            // first / second direction of the axis: 0:(1,3) 1:(2,4) 2:(5,7) 3:(6,8)
            constexpr int ia = (axis == 0) ? 1 : (axis == 1) ? 2 : (axis == 2) ? 5 : 6;
            constexpr int ib = (axis == 0) ? 3 : (axis == 1) ? 4 : (axis == 2) ? 7 : 8;
            D bsa[3], bsb[3], bpa[3], bpb[3], ga[2], gb[2];
            static_for<3>([&](auto S) {
                constexpr int s = decltype(S)::value;
                eq_brackets(axis_dot<axis>(cell.self[s].vx, cell.self[s].vy), cell.self[s].K, c, bsa[s], bsb[s]);
                eq_brackets(axis_dot<axis>(cell.pair[s].vx, cell.pair[s].vy), cell.pair[s].K, c, bpa[s], bpb[s]);
            });
            const D cE = axis_dot<axis>(cell.Ex, cell.Ey);
            static_for<2>([&](auto S) {
                constexpr int s = decltype(S)::value;
                guo_brackets(axis_dot<axis>(cell.self[s].vx, cell.self[s].vy), cE, cell.uE[s], c, ga[s], gb[s]);
            });
            k1_direction<ia>(cell, bsa, bpa, ga, wr, wT, pref, stash, dst, g.plane, cell_off, c);
            k1_direction<ib>(cell, bsb, bpb, gb, wr, wT, pref, stash, dst, g.plane, cell_off, c);
        });
    });
}

static constexpr size_t k1_smem_bytes() { return sizeof(double) * NPLANES * K1_THREADS; }

cudaError_t launch_k1_fused(const double* src, double* dst, const double* Ex, const double* Ey, double* rho_q,
                            const MacroOut* mo, const LbmConsts& c, const LbmGeom& g, cudaStream_t stream)
{
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k1_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1_smem_bytes());
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k1_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1_smem_bytes());
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid((g.NX + K1_THREADS - 1) / K1_THREADS, g.NYl);
    if (mo) k1_fused_kernel<true><<<grid, K1_THREADS, k1_smem_bytes(), stream>>>(src, dst, Ex, Ey, rho_q, *mo, c, g);
    else    k1_fused_kernel<false><<<grid, K1_THREADS, k1_smem_bytes(), stream>>>(src, dst, Ex, Ey, rho_q, MacroOut{}, c, g);
    return cudaGetLastError();
}

} // namespace plbm
