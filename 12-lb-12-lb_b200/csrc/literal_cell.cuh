// literal_cell.cuh -- one cell of the time step written LITERALLY as the reference writes it: every
// product with a lattice velocity, every division, every factor 2.0 and 0.5 in place, IEEE
// operations only.  This is the fallback of the fused kernel for cells whose values leave the domain
// in which the reorganised arithmetic of lbm_cell.cuh is proven identical (subnormal or huge
// intermediates, Inf, NaN): there 0*Inf = NaN, power-of-two scalings round, and only the literal
// sequence reproduces the CPU bit for bit.  It never runs on physical data, so clarity beats speed.
//   UpdateMacro          /root/reference/src/plasma.cpp:317-456
//   ComputeEquilibrium   /root/reference/src/plasma.cpp:162-308
//   ThermalCollisions    /root/reference/src/collisions.cpp:64-122
//   Collisions           /root/reference/src/collisions.cpp:128-181
#pragma once
#include "exact_math.cuh"

namespace plbm {

__device__ __forceinline__ D lit_div(D a, D b) { return D(__ddiv_rn(a.v, b.v)); }

struct LitMacro {
    D rho[3], ux[3], uy[3], T[3], upx[3], upy[3], rho_q;
};

// f, g: [species][direction] populations at the top of the time loop (already streamed)
static __device__ __noinline__ void lit_update_macro(const D (&f)[3][9], const D (&g)[3][9], D Ex, D Ey, const LbmConsts& c, LitMacro& m)
{
    const int cx[9] = { 0, 1, 0, -1, 0, 1, -1, -1, 1 }, cy[9] = { 0, 0, 1, 0, -1, 1, 1, -1, -1 };
    D rl[3];
    for (int k = 0; k < 3; ++k) {
        D r(0.0), ax(0.0), ay(0.0), t(0.0);
        for (int i = 0; i < 9; ++i) {                                         // plasma.cpp:352-372
            r = r + f[k][i];
            ax = ax + f[k][i] * D((double)cx[i]);
            ay = ay + f[k][i] * D((double)cy[i]);
            t = t + g[k][i];
        }
        rl[k] = r;
        D rho(0.0), vx(0.0), vy(0.0), T(0.0);
        if (!(r < D(1e-10))) {
            rho = r; T = t;
            if (k < 2) {                                                      // plasma.cpp:380-391, 400-411
                vx = (ax == r || ax == -r) ? D(0.0) : lit_div(ax, r);
                vy = (ay == r || ay == -r) ? D(0.0) : lit_div(ay, r);
                vx = vx + lit_div((D(0.5) * D(c.q[k])) * Ex, D(c.m[k].d));
                vy = vy + lit_div((D(0.5) * D(c.q[k])) * Ey, D(c.m[k].d));
            } else {                                                          // plasma.cpp:420-424
                vx = lit_div(ax, r);
                vy = lit_div(ay, r);
            }
        }
        m.rho[k] = rho; m.ux[k] = vx; m.uy[k] = vy; m.T[k] = T;
    }
    for (int p = 0; p < 3; ++p) {                                             // plasma.cpp:426-449
        const int a = (p == 2) ? 1 : 0, b = (p == 0) ? 1 : 2;
        if (rl[a] < D(1e-10) && rl[b] < D(1e-10)) { m.upx[p] = D(0.0); m.upy[p] = D(0.0); }
        else {
            m.upx[p] = lit_div(rl[a] * m.ux[a] + rl[b] * m.ux[b], rl[a] + rl[b]);
            m.upy[p] = lit_div(rl[a] * m.uy[a] + rl[b] * m.uy[b], rl[a] + rl[b]);
        }
    }
    D rq = lit_div(D(c.q[1]) * m.rho[1], D(c.m[1].d)) + lit_div(D(c.q[0]) * m.rho[0], D(c.m[0].d));   // plasma.cpp:452
    if (rq < D(1e-15)) rq = D(0.0);
    m.rho_q = rq;
}

__device__ __forceinline__ D lit_bracket(D cu, D u2, D invcs2)                 // plasma.cpp:196-200
{
    return ((D(1.0) + cu * invcs2) + (((cu * cu) * D(0.5)) * invcs2) * invcs2) - (u2 * D(0.5)) * invcs2;
}
__device__ __forceinline__ D lit_thermal_term(D rho, D tau, D feq)             // collisions.cpp:86-96
{
    const D a = D(1.0) - lit_div(D(1.0), tau);
    const D C = lit_div(D(9.0) * feq, tau);
    return lit_div((((D(2.0) * rho) * a) * a - (D(2.0) * a) * rho) - C, D(2.0) * (D(2.0) * a + C));
}

// direction i of all three species: fi/gi in, post-collision values out
static __device__ __noinline__ void lit_collide_direction(int i, const D (&fi)[3], const D (&gi)[3], const LitMacro& m, D Ex, D Ey,
                                                   const LbmConsts& c, D (&fo)[3], D (&go)[3])
{
    const int cxs[9] = { 0, 1, 0, -1, 0, 1, -1, -1, 1 }, cys[9] = { 0, 0, 1, 0, -1, 1, 1, -1, -1 };
    const D cx((double)cxs[i]), cy((double)cys[i]);
    const D w(c.w[i == 0 ? 0 : (i < 5 ? 1 : 2)]);
    const D invcs2(c.invcs2), cs2(c.cs2.d);
    D bs[3], bp[3];
    for (int k = 0; k < 3; ++k) {
        bs[k] = lit_bracket(cx * m.ux[k] + cy * m.uy[k], m.ux[k] * m.ux[k] + m.uy[k] * m.uy[k], invcs2);
        bp[k] = lit_bracket(cx * m.upx[k] + cy * m.upy[k], m.upx[k] * m.upx[k] + m.upy[k] * m.upy[k], invcs2);
    }
    for (int k = 0; k < 3; ++k) {
        D tau[3], feq[3], geq[3];
        for (int mm = 0; mm < 3; ++mm) {
            const int slot = (k == 0) ? (mm == 0 ? 0 : (mm == 1 ? 3 : 4)) : (k == 1) ? (mm == 0 ? 1 : (mm == 1 ? 3 : 5)) : (mm == 0 ? 2 : (mm == 1 ? 4 : 5));
            const int tv[6] = { 5, 3, 1, 6, 4, 2 };
            tau[mm] = D((double)tv[slot]);
            const int pair = (k == 0) ? (mm == 1 ? 0 : 1) : (k == 1) ? (mm == 1 ? 0 : 2) : (mm == 1 ? 1 : 2);
            const D b = (mm == 0) ? bs[k] : bp[pair];
            feq[mm] = (w * m.rho[k]) * b;                                     // plasma.cpp:195-249
            geq[mm] = (w * m.T[k]) * b;                                       // plasma.cpp:251-304
        }
        const D t0 = lit_thermal_term(m.rho[k], tau[0], feq[0]), t1 = lit_thermal_term(m.rho[k], tau[1], feq[1]),
                t2 = lit_thermal_term(m.rho[k], tau[2], feq[2]);
        const D dE = (m.rho[k] * ((t0 + t1) + t2)) * (m.ux[k] * m.ux[k] + m.uy[k] * m.uy[k]);   // collisions.cpp:98-100
        const D dT = lit_div(-dE, D(c.Kb.d));                                                  // collisions.cpp:102-104
        const D CT = (lit_div(-(gi[k] - geq[0]), tau[0]) - lit_div(gi[k] - geq[1], tau[1])) - lit_div(gi[k] - geq[2], tau[2]);
        go[k] = (gi[k] + CT) + dT;                                                              // collisions.cpp:107-114
        const D C = (lit_div(-(fi[k] - feq[0]), tau[0]) - lit_div(fi[k] - feq[1], tau[1])) - lit_div(fi[k] - feq[2], tau[2]);
        D out = fi[k] + C;                                                                      // collisions.cpp:166-173
        if (k < 2) {                                                                            // collisions.cpp:154-163
            const D cE = cx * Ex + cy * Ey, cu = cx * m.ux[k] + cy * m.uy[k];
            const D pref = lit_div(lit_div((w * D(c.q[k])) * m.rho[k], D(c.m[k].d)), cs2) * (D(1.0) - lit_div(D(1.0), D(2.0) * tau[0]));
            out = out + pref * ((cE + lit_div(cu * cE, cs2)) - (m.ux[k] * Ex + m.uy[k] * Ey));
        }
        fo[k] = out;
    }
}

} // namespace plbm
