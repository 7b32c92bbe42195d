// lbm_consts.h -- constants of the plasma LBM step shared by host and device code.
//
// Everything here is derived on the host with plain IEEE double arithmetic in the reference's
// expression order, so the device sees exactly the numbers the reference's loops compute:
//   D2Q9 tables            /root/reference/src/plasma.cpp:10-16
//   relaxation times       /root/reference/src/collisions.cpp:6-7
//   unit system            /root/reference/include/plasma.hpp:86-133
#pragma once

namespace plbm {

// A loop-invariant divisor and its refined reciprocal (see exact_math.cuh).
struct Recip {
    double d;
    double y;
};

constexpr int NQ = 9;
constexpr int NSPEC = 3;                    // 0 electrons, 1 ions, 2 neutrals
constexpr int NPLANES = NSPEC * 2 * NQ;     // SoA planes: ((species*2 + kind)*9 + dir), kind 0 = f, 1 = g

// relaxation-time slots: 0..2 = tau_e, tau_i, tau_n ; 3..5 = tau_e_i, tau_e_n, tau_i_n
constexpr int TAU_VALUE[6] = { 5, 3, 1, 6, 4, 2 };
// the three collision partners of each species in the reference's summation order
// (collisions.cpp:107-109,166-168): self, then the two pairs
constexpr int TAU_SLOT[3][3] = { { 0, 3, 4 }, { 1, 3, 5 }, { 2, 4, 5 } };
// pair velocity used by those partners: 0 = u_ei, 1 = u_en, 2 = u_in (plasma.cpp:426-449)
constexpr int PAIR_SLOT[3][2] = { { 0, 1 }, { 0, 2 }, { 1, 2 } };

struct LbmConsts {
    double invcs2;        // 1.0 / cs2                       plasma.cpp:164
    double hinvcs2;       // 0.5 * invcs2 (exact scaling)
    Recip cs2;            // divisor cs2                     collisions.cpp:154-161
    Recip Kb;             // divisor Kb                      collisions.cpp:102-104
    Recip tau3, tau5, tau6;
    // 1/tau for tau = 3, 5, 6 as an unevaluated sum hi + lo (hi = RN(1/tau), lo = RN(1/tau - hi)):
    // x/tau == fma(hi, x, RN(lo*x)) for every normal-range x, see div_tau() in lbm_cell.cuh
    double inv3[2], inv5[2], inv6[2];
    Recip m[2];           // m_e, m_i as divisors            plasma.cpp:389,409,452
    double hq[2];         // 0.5 * q_s                       plasma.cpp:389,409
    double q[2];          // q_e, q_i                        plasma.cpp:452
    double w[3];          // 4/9, 1/9, 1/36                  plasma.cpp:12-16
    double wq[2][3];      // w_class * q_s                   collisions.cpp:154,159
    double gfac[2];       // 1.0 - 1.0/(2*tau_s)             collisions.cpp:154,159
    double a[6];          // 1.0 - 1.0/tau  per slot         collisions.cpp:86-96
    double a2[6];         // 2.0 * a
    double a4[6];         // 2.0 * a2 (exact)
};

// Geometry of one rank's slab of the lattice.
struct LbmGeom {
    int NX;               // cells per row (global = local: slabs cut y only)
    int NYl;              // local rows
    int pitch;            // doubles between consecutive rows of a population plane
    long long plane;      // doubles between consecutive planes = pitch * (NYl + 2)
    int wrap_y;           // 1: single slab, pull wraps y periodically; 0: halo rows are valid
    int prefetch_rows;    // K1: distance in rows of the L2 prefetch (about one wave of CTAs ahead); 0 = none
};

} // namespace plbm
