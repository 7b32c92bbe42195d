// lbm_cell.cuh -- per-cell arithmetic of the three-species plasma LBM step, written once and used
// by the fused kernel (k1_fused.cu) and by the per-phase kernels behind the reference's free
// functions (phases.cu).  Every routine is a template on the division policy DV (exact_math.cuh).
//
// Every function states the reference expression it evaluates (file:line under /root/reference)
// and, where the evaluation is reorganised, why the result is bit-identical:
//   (E1) products with the lattice velocities c in {0,+1,-1} are exact, x + (+-0) = x, and
//        x + (-y) = x - y, so the direction sums drop their zero terms;
//   (E2) negation commutes with every rounded operation, so opposite directions share |c.u| work;
//   (E3) scaling by a power of two commutes with rounding (no overflow/underflow involved), so
//        the factors 2.0 and 0.5 of collisions.cpp:86-100 are folded into per-cell constants;
//   (E4) x / tau for tau in {1,2,4} is the exact product x * (1/tau).
// Signs of exact zeros may differ from the CPU; no non-zero value does.
#pragma once
#include <utility>
#include "exact_math.cuh"

namespace plbm {

template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>)
{
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f)
{
    static_for_impl(static_cast<F&&>(f), std::make_integer_sequence<int, N>{});
}

// x / tau for the reference's fixed relaxation times (collisions.cpp:6-7).            (E4)
template <int TAU, class DV>
__device__ __forceinline__ D div_tau(DV& dv, D x, const LbmConsts& c)
{
    if constexpr (TAU == 1) return x;
    else if constexpr (TAU == 2) return x * D(0.5);
    else if constexpr (TAU == 4) return x * D(0.25);
    else if constexpr (TAU == 3) return dv.idiv(x, c.tau3, c.inv3);
    else if constexpr (TAU == 5) return dv.idiv(x, c.tau5, c.inv5);
    else { static_assert(TAU == 6, "unknown relaxation time"); return dv.idiv(x, c.tau6, c.inv6); }
}

// (18 * feq) / tau of collisions.cpp:86-96 (doubled, see E3).  For tau in {1,2,4} the quotient of the
// rounded product by a power of two equals the single rounded product (18/tau) * feq.
template <int TAU, class DV>
__device__ __forceinline__ D c18_over_tau(DV& dv, D feq, const LbmConsts& c)
{
    if constexpr (TAU == 1) return D(18.0) * feq;
    else if constexpr (TAU == 2) return D(9.0) * feq;
    else if constexpr (TAU == 4) return D(4.5) * feq;
    else return div_tau<TAU>(dv, D(18.0) * feq, c);
}

// Sums over the nine directions in the reference's i = 0..8 order, plasma.cpp:352-372.   (E1)
__device__ __forceinline__ D sum9(const D (&v)[9])
{
    return (((((((v[0] + v[1]) + v[2]) + v[3]) + v[4]) + v[5]) + v[6]) + v[7]) + v[8];
}
__device__ __forceinline__ D moment_x(const D (&f)[9])   // sum f_i * cx_i, cx = 0,1,0,-1,0,1,-1,-1,1
{
    return ((((f[1] - f[3]) + f[5]) - f[6]) - f[7]) + f[8];
}
__device__ __forceinline__ D moment_y(const D (&f)[9])   // sum f_i * cy_i, cy = 0,0,1,0,-1,1,1,-1,-1
{
    return ((((f[2] - f[4]) + f[5]) + f[6]) - f[7]) - f[8];
}

// Macroscopic state of one cell after UpdateMacro, plasma.cpp:317-456.
struct CellMacro {
    D rho[3], ux[3], uy[3], T[3];   // stored (thresholded) moments
    D upx[3], upy[3];               // barycentric pair velocities u_ei, u_en, u_in
    D rho_q;
};

// Velocity of one species from its raw sums, plasma.cpp:373-424.  Returns the stored density
// (0 below the 1e-10 threshold) and u = sum(f c)/rho (+ 0.5 q E / m for the charged species).
template <int s, class DV>
__device__ __forceinline__ void species_velocity(DV& dv, D rl, D mx, D my, D Ex, D Ey, const LbmConsts& c,
                                                 D& rho, D& ux, D& uy)
{
    if (rl < D(1e-10)) {                                                      // plasma.cpp:373-377
        rho = D(0.0); ux = D(0.0); uy = D(0.0);
    } else {
        rho = rl;
        if constexpr (s < 2) {                                                // plasma.cpp:380-391, 400-411
            D vx = dv.xdiv(mx, rl);
            D vy = dv.xdiv(my, rl);
            if (mx == rl || mx == -rl) vx = D(0.0);
            if (my == rl || my == -rl) vy = D(0.0);
            ux = vx + dv.cdiv(D(c.hq[s]) * Ex, c.m[s]);                        // 0.5*q*Ex/m
            uy = vy + dv.cdiv(D(c.hq[s]) * Ey, c.m[s]);
        } else {                                                              // plasma.cpp:420-424
            ux = dv.xdiv(mx, rl);
            uy = dv.xdiv(my, rl);
        }
    }
}

// Barycentric velocity of a pair from the raw densities and the stored velocities, plasma.cpp:426-449.
template <class DV>
__device__ __forceinline__ void pair_velocity(DV& dv, D rla, D rlb, D uxa, D uya, D uxb, D uyb, D& upx, D& upy)
{
    if (rla < D(1e-10) && rlb < D(1e-10)) {
        upx = D(0.0); upy = D(0.0);
    } else {
        const D den = rla + rlb;
        upx = dv.xdiv(rla * uxa + rlb * uxb, den);
        upy = dv.xdiv(rla * uya + rlb * uyb, den);
    }
}

// Charge density from the stored densities, plasma.cpp:452-453.
template <class DV>
__device__ __forceinline__ D charge_density(DV& dv, D rho_e, D rho_i, const LbmConsts& c)
{
    D rq = dv.cdiv(D(c.q[1]) * rho_i, c.m[1]) + dv.cdiv(D(c.q[0]) * rho_e, c.m[0]);
    if (rq < D(1e-15)) rq = D(0.0);
    return rq;
}

// rl/mx/my/tl: the raw local sums of species 0..2.  Whole UpdateMacro of one cell.
template <class DV>
__device__ __forceinline__ void cell_update_macro(DV& dv, const D (&rl)[3], const D (&mx)[3], const D (&my)[3],
                                                  const D (&tl)[3], D Ex, D Ey, const LbmConsts& c, CellMacro& m)
{
    static_for<3>([&](auto S) {
        constexpr int s = decltype(S)::value;
        species_velocity<s>(dv, rl[s], mx[s], my[s], Ex, Ey, c, m.rho[s], m.ux[s], m.uy[s]);
        m.T[s] = (rl[s] < D(1e-10)) ? D(0.0) : tl[s];
    });
    static_for<3>([&](auto P) {
        constexpr int p = decltype(P)::value;
        constexpr int a = (p == 2) ? 1 : 0, b = (p == 0) ? 1 : 2;
        pair_velocity(dv, rl[a], rl[b], m.ux[a], m.uy[a], m.ux[b], m.uy[b], m.upx[p], m.upy[p]);
    });
    m.rho_q = charge_density(dv, m.rho[0], m.rho[1], c);
}

// ---- the same UpdateMacro with the per-cell gate of exact_math.cuh (CellGate) instead of a record per division ------------
// Noted here: raw densities (NaN detection), rho_s and T_s (upper bound; T_s also from below), the numerators mx, my of the
// velocity divisions, the velocities ux_s, uy_s (both bounds), the pair densities and numerators, the pair velocities (upper
// bound).  Divisions by the species' own density need no note: rho_s >= 1e-10 on that branch and rho_s < 2^200 is noted.
template <int s>
__device__ __forceinline__ void species_velocity(GatedDiv& dv, CellGate& gt, D rl, D mx, D my, D Ex, D Ey, const LbmConsts& c,
                                                 D& rho, D& ux, D& uy)
{
    gt.note_raw_density(rl);
    if (rl < D(1e-10)) {                                                      // plasma.cpp:373-377
        rho = D(0.0); ux = D(0.0); uy = D(0.0);
    } else {
        rho = rl;
        gt.note_scalar(rl);
        gt.note_num(mx);
        gt.note_num(my);
        if constexpr (s < 2) {                                                // plasma.cpp:380-391, 400-411
            D vx = dv.xdiv(mx, rl);
            D vy = dv.xdiv(my, rl);
            if (mx == rl || mx == -rl) vx = D(0.0);
            if (my == rl || my == -rl) vy = D(0.0);
            ux = vx + dv.cdiv(D(c.hq[s]) * Ex, c.m[s]);                        // 0.5*q*Ex/m
            uy = vy + dv.cdiv(D(c.hq[s]) * Ey, c.m[s]);
        } else {                                                              // plasma.cpp:420-424
            ux = dv.xdiv(mx, rl);
            uy = dv.xdiv(my, rl);
        }
        gt.note_velocity(ux);
        gt.note_velocity(uy);
    }
}
__device__ __forceinline__ void pair_velocity(GatedDiv& dv, CellGate& gt, D rla, D rlb, D uxa, D uya, D uxb, D uyb, D& upx, D& upy)
{
    if (rla < D(1e-10) && rlb < D(1e-10)) {                                   // plasma.cpp:426-449
        upx = D(0.0); upy = D(0.0);
    } else {
        const D den = rla + rlb;
        const D nx = rla * uxa + rlb * uxb, ny = rla * uya + rlb * uyb;
        gt.note_den(den);
        gt.note_num(nx);
        gt.note_num(ny);
        upx = dv.xdiv(nx, den);
        upy = dv.xdiv(ny, den);
        gt.note_pair_velocity(upx);
        gt.note_pair_velocity(upy);
    }
}
__device__ __forceinline__ void cell_update_macro(GatedDiv& dv, CellGate& gt, const D (&rl)[3], const D (&mx)[3], const D (&my)[3],
                                                  const D (&tl)[3], D Ex, D Ey, const LbmConsts& c, CellMacro& m)
{
    gt.note_field(Ex);
    gt.note_field(Ey);
    static_for<3>([&](auto S) {
        constexpr int s = decltype(S)::value;
        species_velocity<s>(dv, gt, rl[s], mx[s], my[s], Ex, Ey, c, m.rho[s], m.ux[s], m.uy[s]);
        m.T[s] = (rl[s] < D(1e-10)) ? D(0.0) : tl[s];
        gt.note_temperature(m.T[s]);
    });
    static_for<3>([&](auto P) {
        constexpr int p = decltype(P)::value;
        constexpr int a = (p == 2) ? 1 : 0, b = (p == 0) ? 1 : 2;
        pair_velocity(dv, gt, rl[a], rl[b], m.ux[a], m.uy[a], m.ux[b], m.uy[b], m.upx[p], m.upy[p]);
    });
    m.rho_q = charge_density(dv, m.rho[0], m.rho[1], c);
}

// Direction-independent pieces of the equilibrium bracket for one velocity (plasma.cpp:169-174,
// 196-200):  K = u2*0.5*invcs2.
struct VelSet {
    D vx, vy, K;
};
__device__ __forceinline__ VelSet make_velset(D vx, D vy, const LbmConsts& c)
{
    VelSet v;
    v.vx = vx; v.vy = vy;
    v.K = (vx * vx + vy * vy) * D(c.hinvcs2);                                 // (u2*0.5)*invcs2        (E3)
    return v;
}

// c_i . v for the first direction of an axis (the second direction is its negative):            (E1,E2)
//   axis 0: dirs (1,3) c=(1,0)   1: dirs (2,4) c=(0,1)   2: dirs (5,7) c=(1,1)   3: dirs (6,8) c=(-1,1)
//   axis 4: the rest direction, c = 0
// Branch-free: each component is kept, negated or replaced by +0 with integer masks (exact), then
// one rounded addition -- exactly cx*vx + cy*vy of plasma.cpp:185-190 up to the sign of a zero.
struct AxisSel {
    unsigned keep_x, keep_y;   // 0xffffffff keeps the component, 0 replaces it by +0
    unsigned flip_x;           // 0x80000000 negates vx (c_x = -1)
};
__device__ __forceinline__ AxisSel axis_select(int axis)
{
    AxisSel a;
    a.keep_x = (axis == 0 || axis == 2 || axis == 3) ? 0xffffffffu : 0u;
    a.keep_y = (axis >= 1 && axis <= 3) ? 0xffffffffu : 0u;
    a.flip_x = (axis == 3) ? 0x80000000u : 0u;
    return a;
}
__device__ __forceinline__ D axis_dot(const AxisSel& a, D vx, D vy)
{
    const double x = __hiloint2double((int)((((unsigned)__double2hiint(vx.v)) & a.keep_x) ^ a.flip_x), (int)(((unsigned)__double2loint(vx.v)) & a.keep_x));
    const double y = __hiloint2double((int)(((unsigned)__double2hiint(vy.v)) & a.keep_y), (int)(((unsigned)__double2loint(vy.v)) & a.keep_y));
    return D(x) + D(y);
}

// Sign-independent parts of the bracket 1 + cu*invcs2 + cu*cu*0.5*invcs2*invcs2 - K, plasma.cpp:196-200
struct BracketParts {
    D cu, P, S;
};
__device__ __forceinline__ BracketParts bracket_parts(D cu, const LbmConsts& c)
{
    BracketParts b;
    b.cu = cu;
    b.P = cu * D(c.invcs2);
    b.S = ((cu * cu) * D(c.hinvcs2)) * D(c.invcs2);                           // (E3)
    return b;
}
// bracket for cu (mask 0) or -cu (mask 0x80000000)                                               (E2)
__device__ __forceinline__ D bracket_value(const BracketParts& b, D K, unsigned mask)
{
    return ((D(1.0) + flip_sign(b.P, mask)) + b.S) - K;
}
// same with the direction's sign known at compile time: the negation rides on the DADD operand
template <bool NEG>
__device__ __forceinline__ D bracket_value(const BracketParts& b, D K)
{
    if constexpr (NEG) return ((D(1.0) - b.P) + b.S) - K;
    else return ((D(1.0) + b.P) + b.S) - K;
}

// Per-cell, direction-independent part of the thermal source, collisions.cpp:86-96:
// AB2 = 2*(2*rho*a*a - 2*a*rho) for the three relaxation times of species s.            (E3)
template <int S>
__device__ __forceinline__ void thermal_cell_terms(D rho, const LbmConsts& c, D (&AB2)[3])
{
    const D r2 = D(2.0) * rho;
    static_for<3>([&](auto M) {
        constexpr int slot = TAU_SLOT[S][decltype(M)::value];
        const D A = (r2 * D(c.a[slot])) * D(c.a[slot]);
        const D B = D(c.a2[slot]) * rho;
        AB2[decltype(M)::value] = D(2.0) * (A - B);
    });
}

// One species, one direction: thermal collision (collisions.cpp:86-114) and mass collision
// (collisions.cpp:154-173) given the three equilibrium brackets b[0..2] (self, pair 1, pair 2).
//   wr = w_i*rho_s, wT = w_i*T_s, rhoh = 0.5*rho_s, u2 = ux_s^2+uy_s^2, force = Guo term (0 for neutrals)
template <int S, class DV>
__device__ __forceinline__ void collide_species_dir(DV& dv, D fv, D gv, const D (&b)[3], D wr, D wT, const D (&AB2)[3],
                                                    D rhoh, D u2, D force, const LbmConsts& c, D& fnew, D& gnew)
{
    D feq[3], geq[3], q[3], df[3], dg[3];
    static_for<3>([&](auto M) {
        constexpr int m = decltype(M)::value;
        constexpr int slot = TAU_SLOT[S][m];
        constexpr int tau = TAU_VALUE[slot];
        feq[m] = wr * b[m];                                                   // plasma.cpp:195-249
        geq[m] = wT * b[m];                                                   // plasma.cpp:251-304
        if constexpr (tau == 1 && DV::unit_quotient_shortcut) {
            // a = 1 - 1/tau = 0: AB2 = a4 = 0 and the quotient below is (0 - C18)/(0 + C18) = -1 (GatedDiv::note_unit_quotient)
            dv.note_unit_quotient(feq[m]);
            q[m] = D(-1.0);
        } else {
            const D C18 = c18_over_tau<tau>(dv, feq[m], c);                   // 2 * (Q*feq/tau)          (E3)
            q[m] = dv.xdiv(AB2[m] - C18, D(c.a4[slot]) + C18);                // 2 * term_xy, collisions.cpp:86-96
        }
        df[m] = div_tau<tau>(dv, fv - feq[m], c);                             // collisions.cpp:166-168
        dg[m] = div_tau<tau>(dv, gv - geq[m], c);                             // collisions.cpp:107-109
    });
    const D dE = (rhoh * ((q[0] + q[1]) + q[2])) * u2;                        // collisions.cpp:98-100    (E3)
    const D dT = dv.cdiv(dE, c.Kb);                                           // -DeltaT, collisions.cpp:102-104
    gnew = (gv - ((dg[0] + dg[1]) + dg[2])) - dT;                             // collisions.cpp:112-114
    fnew = fv - ((df[0] + df[1]) + df[2]);                                    // collisions.cpp:171-173
    if constexpr (S < 2) fnew = fnew + force;
}

// Guo forcing prefactor w*q*rho/m/cs2*(1-1/(2 tau)), collisions.cpp:154,159.
template <int S, class DV>
__device__ __forceinline__ D guo_prefactor(DV& dv, int wclass, D rho, const LbmConsts& c)
{
    return dv.cdiv(dv.cdiv(D(c.wq[S][wclass]) * rho, c.m[S]), c.cs2) * D(c.gfac[S]);
}
// Guo bracket (c.E) + (c.u)(c.E)/cs2 - (u.E) for c (mask 0) and -c (mask 0x80000000),
// collisions.cpp:155-157; X = (c.u)(c.E)/cs2 is the same for both.                             (E2)
__device__ __forceinline__ D guo_bracket(D X, D cE, D uE, unsigned mask)
{
    return (flip_sign(cE, mask) + X) - uE;
}
template <bool NEG>
__device__ __forceinline__ D guo_bracket(D X, D cE, D uE)
{
    if constexpr (NEG) return (X - cE) - uE;      // (-cE) + X
    else return (cE + X) - uE;
}

} // namespace plbm
