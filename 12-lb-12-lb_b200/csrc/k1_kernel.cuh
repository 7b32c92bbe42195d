// k1_kernel.cuh -- K1: one pass per time step over all three species, f and g (kernel templates; instantiated by
// k1_fused.cu for periodic lattices and k1_walls.cu for bounce-back walls).
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
// Algorithmic traffic: 54*8 read + 54*8 written + phi 8 (or Ex,Ey 16) + rho_q 8 = 880 (888) B per cell and step.
//
// One thread owns one cell.  The 54 pulled populations are parked in shared memory (the "stash", one column per thread,
// conflict-free) while the moments are formed, then read again species by species for the collisions: a ROLLED loop over the
// four axes that carry two opposite directions (its body, ~4.5 KB of code, stays in the instruction cache) and the rest
// direction with everything resolved at compile time.  Divisions use the 2/3/8-instruction sequences of exact_math.cuh; a
// per-cell gate (CellGate) proves them exact from ~90 noted values per cell ("K1 gate"), and a cell outside
// the gate (never on physical data) is recomputed out of line with the reference's literal arithmetic (literal_cell.cuh).
//
// Three kernels share the cell code:
//   k1_fused_kernel   periodic lattices, every thread pulls its own 54 values (LDG -> STS); one CTA per 64-cell tile
//   k1_walls_kernel   bounce-back walls (walls.cuh), same structure
//   k1_pool_kernel    periodic lattices, PERSISTENT warps: the pull is done by the TMA engine into a pool of stash buffers,
//                     and a warp asks for its next tile while it still computes the current one (no load phase)
#pragma once
#include <cuda.h>
#include <cstdint>
#include "k1_fused.h"
#include "lbm_cell.cuh"
#include "literal_cell.cuh"
#include "walls.cuh"

namespace plbm {

#ifndef PLBM_K1_THREADS
#define PLBM_K1_THREADS 64
#endif
#ifndef PLBM_K1_MIN_BLOCKS
#define PLBM_K1_MIN_BLOCKS 6
#endif
#ifndef PLBM_K1_CARVEOUT
#define PLBM_K1_CARVEOUT (-1)       // preferred shared-memory carve-out in percent; -1: leave the choice to the driver
#endif
#ifndef PLBM_K1_PREFETCH
#define PLBM_K1_PREFETCH 1          // 1: every CTA asks L2 for the row segments a later CTA will pull (k1_fused_kernel)
#endif
#ifndef PLBM_K1_UNROLL
#define PLBM_K1_UNROLL 0            // 1: all five axes unrolled (20 % fewer instructions, ~60 KB of code per cell: measured slower)
#endif
#ifndef PLBM_K1_SHARE
#define PLBM_K1_SHARE 1             // 1: k1_tma_kernel computes every pair bracket once per cell and hands it on through the stash
#endif
constexpr int K1_THREADS = PLBM_K1_THREADS;

// Where a parked population lives in the stash, relative to the thread's own pointer.
//   ColumnLayout  [direction][species*2+kind][thread of the CTA]
//   PoolLayout    k1_pool_kernel: nine TMA boxes (one per direction) of six rows of POOL_BOX_W doubles; a thread reads column
//                 lane + 1 of a box whose direction has c_x != 0 (the box starts two cells early / at the tile, see there)
struct ColumnLayout {
    __host__ __device__ static constexpr int slot(int sk, int dir) { return (dir * (2 * NSPEC) + sk) * K1_THREADS; }
};
constexpr int POOL_TILE = 32;                                   // cells per tile = one warp
constexpr int POOL_BOX_W = POOL_TILE + 2;                       // box width: tile + the two cells the +-1 pull offsets need
constexpr int POOL_DIR_STRIDE = 208;                            // doubles between the boxes of two directions (6 * 34 = 204, padded to 128 B)
constexpr int POOL_BUF_DOUBLES = NQ * POOL_DIR_STRIDE;          // 1872 doubles = 14 976 B per stash buffer
struct PoolLayout {
    __host__ __device__ static constexpr int slot(int sk, int dir)
    {
        return dir * POOL_DIR_STRIDE + sk * POOL_BOX_W + ((dir == 0 || dir == 2 || dir == 4) ? 0 : 1);
    }
};

// k1_tma_kernel: nine TMA boxes (one per direction) of six rows of K1_THREADS + 2 doubles, same column rule as PoolLayout
constexpr int TMA_BOX_W = K1_THREADS + 2;
constexpr int TMA_DIR_STRIDE = ((2 * NSPEC * TMA_BOX_W * 8 + 127) / 128) * 128 / 8;      // doubles between two directions' boxes (128-byte aligned)
struct TmaLayout {
    __host__ __device__ static constexpr int slot(int sk, int dir)
    {
        return dir * TMA_DIR_STRIDE + sk * TMA_BOX_W + ((dir == 0 || dir == 2 || dir == 4) ? 0 : 1);
    }
};

struct K1Out {
    double* __restrict__ dst;     // population planes, already offset to this cell
    long long plane;
    double* rho_q;                // already offset to this cell
    MacroOut mo;                  // base pointers
    long long cidx;
};

// Second read of a parked population (the first was for the moments).  Volatile so that the compiler really reads shared
// memory again instead of keeping values alive in registers and local memory across the moments (it spills otherwise).
__device__ __forceinline__ D stash_reload(const double* p)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return D(v);
}

// Stash accesses of the collision phase when brackets are handed on through it (SHARE): volatile, so that the compiler keeps
// the program order between the last read of a population and the bracket stored over it, and re-reads shared memory.
__device__ __forceinline__ D stash_ld(const double* p) { return D(*reinterpret_cast<const volatile double*>(p)); }
__device__ __forceinline__ void stash_st(double* p, D v) { *reinterpret_cast<volatile double*>(p) = v.v; }

// Pair brackets handed on between the species (SHARE).  The equilibrium bracket of a pair velocity (u_ei, u_en, u_in) in a
// direction is the same number for both species of the pair (plasma.cpp:195-304 evaluates it once per species), so the first
// species of a pair leaves it in a stash slot whose population it has just consumed and the second reads it back:
//   u_ei: electrons -> slot f_e   (read by ions)        u_en: electrons -> slot g_e   (read by neutrals)
//   u_in: ions      -> slot f_i   (read by neutrals)
// 27 brackets per cell are not recomputed (-135 FP64 instructions, +27 STS/LDS).  b[j], j = 1, 2 are the pairs of PAIR_SLOT.
template <int S, int J> struct SharedBracket {
    static constexpr bool computed = (S == 0) || (S == 1 && J == 2);            // else: read back
    // stash distribution (species*2 + kind) that carries the bracket of pair PAIR_SLOT[S][J-1]
    static constexpr int sk = (PAIR_SLOT[S][J - 1] == 0) ? 0 : (PAIR_SLOT[S][J - 1] == 1 ? 1 : 2);
};

// c_axis . v for the first direction of a compile-time axis (lbm_cell.cuh: axis_select)            (E1)
template <int AXIS>
__device__ __forceinline__ D axis_dot_ct(D vx, D vy)
{
    if constexpr (AXIS == 0) return vx;
    else if constexpr (AXIS == 1) return vy;
    else if constexpr (AXIS == 2) return vx + vy;
    else { static_assert(AXIS == 3, "axis 4 is the rest direction"); return vy - vx; }
}

// One axis with its direction(s) known at compile time: c.v is a move, an addition or a subtraction, the weight class and the
// store offsets are constants, and the rest direction's bracket is 1 - K.
template <int S, int AXIS, class L, bool SHARE = false>
__device__ __forceinline__ void k1_axis_ct(GatedDiv& dv, CellGate& gt, std::conditional_t<SHARE, double*, const double* __restrict__> stash, const CellMacro& m,
                                           const D (&vx)[3], const D (&vy)[3], const D (&K)[3], const D (&AB2)[3], D rhoh, D u2,
                                           D uE, const D (&pref3)[3], D Ex, D Ey, const K1Out& o, const LbmConsts& c)
{
    constexpr int wclass = (AXIS < 2) ? 1 : (AXIS < 4 ? 2 : 0);
    constexpr int d0 = (AXIS == 4) ? 0 : (AXIS < 2 ? AXIS + 1 : AXIS + 3);    // first direction of the axis; the opposite is d0 + 2
    const D wr = D(c.w[wclass]) * m.rho[S];
    const D wT = D(c.w[wclass]) * m.T[S];
    auto finish = [&](const int dir, D (&b)[3], D force) {
        D fv, gv;
        if constexpr (SHARE) {
            fv = stash_ld(stash + L::slot(S * 2 + 0, dir)); gv = stash_ld(stash + L::slot(S * 2 + 1, dir));
            if constexpr (SharedBracket<S, 1>::computed) stash_st(stash + L::slot(SharedBracket<S, 1>::sk, dir), b[1]);
            else b[1] = stash_ld(stash + L::slot(SharedBracket<S, 1>::sk, dir));
            if constexpr (SharedBracket<S, 2>::computed) stash_st(stash + L::slot(SharedBracket<S, 2>::sk, dir), b[2]);
            else b[2] = stash_ld(stash + L::slot(SharedBracket<S, 2>::sk, dir));
        } else {
            fv = stash_reload(stash + L::slot(S * 2 + 0, dir)); gv = stash_reload(stash + L::slot(S * 2 + 1, dir));
        }
        D fnew, gnew;
        collide_species_dir<S>(dv, fv, gv, b, wr, wT, AB2, rhoh, u2, force, c, fnew, gnew);
        gt.note_output(fnew);
        gt.note_output(gnew);
        o.dst[((S * 2 + 0) * NQ + dir) * o.plane] = fnew.v;
        o.dst[((S * 2 + 1) * NQ + dir) * o.plane] = gnew.v;
    };
    if constexpr (AXIS == 4) {
        // rest direction: c = 0, so c.v = 0, the bracket is ((1 + 0) + 0) - K = 1 - K and the Guo bracket (0 + 0/cs2) - u.E
        D b[3];
        b[0] = D(1.0) - K[0];
        if constexpr (!SHARE || SharedBracket<S, 1>::computed) b[1] = D(1.0) - K[1];
        if constexpr (!SHARE || SharedBracket<S, 2>::computed) b[2] = D(1.0) - K[2];
        D force = D(0.0);
        if constexpr (S < 2) force = pref3[0] * (D(0.0) - uE);
        finish(d0, b, force);
    } else {
        static_assert(!SHARE, "the fully unrolled variant does not hand brackets on");
        BracketParts bp[3];
        #pragma unroll
        for (int j = 0; j < 3; ++j) bp[j] = bracket_parts(axis_dot_ct<AXIS>(vx[j], vy[j]), c);
        D X = D(0.0), cE = D(0.0);
        if constexpr (S < 2) {
            cE = axis_dot_ct<AXIS>(Ex, Ey);
            X = dv.cdiv(bp[0].cu * cE, c.cs2);                                 // (c.u)(c.E)/cs2
        }
        D b0[3], b1[3];
        #pragma unroll
        for (int j = 0; j < 3; ++j) { b0[j] = bracket_value<false>(bp[j], K[j]); b1[j] = bracket_value<true>(bp[j], K[j]); }
        D force0 = D(0.0), force1 = D(0.0);
        if constexpr (S < 2) {                                                // collisions.cpp:154-163
            force0 = pref3[wclass] * guo_bracket<false>(X, cE, uE);
            force1 = pref3[wclass] * guo_bracket<true>(X, cE, uE);
        }
        finish(d0, b0, force0);
        finish(d0 + 2, b1, force1);
    }
}

// ---- an EMPTY charged species (stored density 0: raw density below the reference's 1e-10 threshold, plasma.cpp:373-377) ---------
// With rho_s = 0 the reference's own expressions collapse exactly: u_s = 0 and T_s = 0 (plasma.cpp:373-377), so for every partner m
//   feq_m = (w*0)*b_m = 0, geq_m = (w*0)*b_m = 0            (b_m finite: the gate bounds all six velocities)
//   term_m = (0 - 0)/(a4_m + 0) = 0                         (a4_m > 0: electrons and ions have no partner with tau = 1)
//   DeltaE = (0.5*0 * 0) * 0 = 0, DeltaT = 0/Kb = 0, Guo term = (w q 0/m/cs2 (1 - 1/2tau)) * bracket = 0
// and what is left of collisions.cpp:107-114,166-173 is  f - ((f/tau_0 + f/tau_1) + f/tau_2)  and the same for g -- 16 FP64
// instructions per direction instead of ~150.  Bit-identical up to the sign of exact zeros.  In the reference's own initial
// condition (plasma.cpp:128-158) electrons and ions are empty outside the central quarter of the lattice.  Not for neutrals:
// their self collision has tau = 1 (a4 = 0), the reference divides 0 by 0 there and the full path reproduces that.
// With SHARE the species still leaves the pair brackets it owes the later species in the stash.
#ifndef PLBM_K1_LIGHT
#define PLBM_K1_LIGHT 1
#endif
template <int S, class L, bool SHARE, bool BOTH>   // BOTH: electrons and ions are both empty, nobody reads the bracket of u_ei
__device__ __forceinline__ void k1_species_empty(GatedDiv& dv, CellGate& gt, std::conditional_t<SHARE, double*, const double* __restrict__> stash,
                                                 const CellMacro& m, const K1Out& o, const LbmConsts& c)
{
    static_assert(S < 2, "neutrals take the full path");
    constexpr int T0 = TAU_VALUE[TAU_SLOT[S][0]], T1 = TAU_VALUE[TAU_SLOT[S][1]], T2 = TAU_VALUE[TAU_SLOT[S][2]];
    static_assert(T0 != 1 && T1 != 1 && T2 != 1, "a partner with tau = 1 divides 0 by 0");
    constexpr int p0 = PAIR_SLOT[S][0], p1 = PAIR_SLOT[S][1];
    constexpr bool owe1 = SHARE && SharedBracket<S, 1>::computed && !(BOTH && PAIR_SLOT[S][0] == 0);
    constexpr bool owe2 = SHARE && SharedBracket<S, 2>::computed;
    D K1 = D(0.0), K2 = D(0.0);
    if constexpr (owe1) K1 = (m.upx[p0] * m.upx[p0] + m.upy[p0] * m.upy[p0]) * D(c.hinvcs2);
    if constexpr (owe2) K2 = (m.upx[p1] * m.upx[p1] + m.upy[p1] * m.upy[p1]) * D(c.hinvcs2);
    auto relax = [&](const int dir, const D fv, const D gv) {
        const D fnew = fv - ((div_tau<T0>(dv, fv, c) + div_tau<T1>(dv, fv, c)) + div_tau<T2>(dv, fv, c));
        const D gnew = gv - ((div_tau<T0>(dv, gv, c) + div_tau<T1>(dv, gv, c)) + div_tau<T2>(dv, gv, c));
        gt.note_output(fnew);
        gt.note_output(gnew);
        o.dst[((S * 2 + 0) * NQ + dir) * o.plane] = fnew.v;
        o.dst[((S * 2 + 1) * NQ + dir) * o.plane] = gnew.v;
    };
    auto ld = [&](const int sk, const int dir) {
        if constexpr (SHARE) return stash_ld(stash + L::slot(sk, dir));
        else return D(stash[L::slot(sk, dir)]);
    };
    #pragma unroll 1
    for (int axis = 0; axis < 4; ++axis) {
        const int d0 = (axis < 2) ? axis + 1 : axis + 3;
        const D f0 = ld(S * 2 + 0, d0), g0 = ld(S * 2 + 1, d0), f1 = ld(S * 2 + 0, d0 + 2), g1 = ld(S * 2 + 1, d0 + 2);
        if constexpr (owe1 || owe2) {
            const AxisSel sel = axis_select(axis);
            if constexpr (owe1) {
                const BracketParts bp = bracket_parts(axis_dot(sel, m.upx[p0], m.upy[p0]), c);
                stash_st(stash + L::slot(SharedBracket<S, 1>::sk, d0), bracket_value<false>(bp, K1));
                stash_st(stash + L::slot(SharedBracket<S, 1>::sk, d0 + 2), bracket_value<true>(bp, K1));
            }
            if constexpr (owe2) {
                const BracketParts bp = bracket_parts(axis_dot(sel, m.upx[p1], m.upy[p1]), c);
                stash_st(stash + L::slot(SharedBracket<S, 2>::sk, d0), bracket_value<false>(bp, K2));
                stash_st(stash + L::slot(SharedBracket<S, 2>::sk, d0 + 2), bracket_value<true>(bp, K2));
            }
        }
        relax(d0, f0, g0);
        relax(d0 + 2, f1, g1);
    }
    const D fr = ld(S * 2 + 0, 0), gr = ld(S * 2 + 1, 0);
    if constexpr (owe1) stash_st(stash + L::slot(SharedBracket<S, 1>::sk, 0), D(1.0) - K1);
    if constexpr (owe2) stash_st(stash + L::slot(SharedBracket<S, 2>::sk, 0), D(1.0) - K2);
    relax(0, fr, gr);
}

struct NoHook {
    template <int S> __device__ __forceinline__ void at_axis(int) {}
};

template <class L, bool SHARE, bool CHARGED_EMPTY, class Hook>
__device__ __forceinline__ void k1_cell_collide(GatedDiv& dv, CellGate& gt, std::conditional_t<SHARE, double*, const double* __restrict__> stash,
                                                const CellMacro& m, D Ex, D Ey, const K1Out& o, const LbmConsts& c, Hook& hook);

template <bool WRITE_MACRO, class L, bool SHARE, bool CHARGED_EMPTY, class Hook>
__device__ __forceinline__ void k1_cell_rest(GatedDiv& dv, CellGate& gt, std::conditional_t<SHARE, double*, const double* __restrict__> stash,
                                             const D (&rl)[3], const D (&mx)[3], const D (&my)[3], const D (&tl)[3], D Ex, D Ey,
                                             const K1Out& o, const LbmConsts& c, Hook& hook);

// The whole cell.  L: stash layout.  `hook.at_axis<s>(axis)` runs at the top of every iteration of the axis loop of species s
// (the pool kernel asks for its next tile at one of them).
template <bool WRITE_MACRO, class L, bool SHARE, class Hook>
__device__ __forceinline__ void k1_cell_gated(GatedDiv& dv, CellGate& gt, std::conditional_t<SHARE, double*, const double* __restrict__> stash,
                                              D Ex, D Ey, const K1Out& o, const LbmConsts& c, Hook& hook)
{
    static_assert(!(SHARE && PLBM_K1_UNROLL), "the fully unrolled variant does not hand brackets on");
    // raw sums of the three species (plasma.cpp:352-372)
    D rl[3], mx[3], my[3], tl[3];
    static_for<3>([&](auto S) {
        constexpr int s = decltype(S)::value;
        D f[NQ], g[NQ];
        #pragma unroll
        for (int i = 0; i < NQ; ++i) {
            f[i] = D(stash[L::slot(s * 2 + 0, i)]);
            g[i] = D(stash[L::slot(s * 2 + 1, i)]);
        }
        #pragma unroll
        for (int i = 0; i < NQ; ++i) { gt.note_input(f[i].v); gt.note_input(g[i].v); }
        rl[s] = sum9(f); mx[s] = moment_x(f); my[s] = moment_y(f); tl[s] = sum9(g);
    });
#if PLBM_K1_LIGHT && !PLBM_K1_UNROLL
    // Two complete copies of the rest of the cell, chosen as soon as the raw densities are known: the general one, and one for
    // cells whose electrons AND ions are empty (k1_species_empty for both, the full path for the neutrals).  Whole copies, branched
    // this early, so that the general copy is scheduled exactly like the code without the short path: a branch per species, or
    // one after UpdateMacro, cost the general path 5-6 % (profiles/r2_k1_sweeps.md).
    if constexpr (std::is_same_v<Hook, NoHook>) {
        // (a warp vote: the branch is uniform, so both copies keep their constants in uniform registers; a warp with cells of
        // both kinds takes the general copy, which is correct for every cell)
        if (__all_sync(0xffffffffu, rl[0] < D(1e-10) && rl[1] < D(1e-10))) {  // the reference's own test, plasma.cpp:373,393
            k1_cell_rest<WRITE_MACRO, L, SHARE, true>(dv, gt, stash, rl, mx, my, tl, Ex, Ey, o, c, hook);
            return;
        }
    }
#endif
    k1_cell_rest<WRITE_MACRO, L, SHARE, false>(dv, gt, stash, rl, mx, my, tl, Ex, Ey, o, c, hook);
}

// UpdateMacro from the raw sums, then the collisions.  CHARGED_EMPTY: electrons and ions are below the density threshold.
template <bool WRITE_MACRO, class L, bool SHARE, bool CHARGED_EMPTY, class Hook>
__device__ __forceinline__ void k1_cell_rest(GatedDiv& dv, CellGate& gt, std::conditional_t<SHARE, double*, const double* __restrict__> stash,
                                             const D (&rl)[3], const D (&mx)[3], const D (&my)[3], const D (&tl)[3], D Ex, D Ey,
                                             const K1Out& o, const LbmConsts& c, Hook& hook)
{
    CellMacro m;
    cell_update_macro(dv, gt, rl, mx, my, tl, Ex, Ey, c, m);
    gt.close_macro();
    *o.rho_q = m.rho_q.v;
    if constexpr (WRITE_MACRO) {
        static_for<3>([&](auto S) {
            constexpr int s = decltype(S)::value;
            o.mo.ux[s][o.cidx] = m.ux[s].v; o.mo.uy[s][o.cidx] = m.uy[s].v;
            o.mo.T[s][o.cidx] = m.T[s].v;   o.mo.rho[s][o.cidx] = m.rho[s].v;
        });
    }
    k1_cell_collide<L, SHARE, CHARGED_EMPTY>(dv, gt, stash, m, Ex, Ey, o, c, hook);
}

// ---- collisions, species by species (small live state per species) -----------------------
// Every species needs three equilibrium velocities: its own and those of its two pairs (plasma.cpp:195-304).
template <class L, bool SHARE, bool CHARGED_EMPTY, class Hook>
__device__ __forceinline__ void k1_cell_collide(GatedDiv& dv, CellGate& gt, std::conditional_t<SHARE, double*, const double* __restrict__> stash,
                                                const CellMacro& m, D Ex, D Ey, const K1Out& o, const LbmConsts& c, Hook& hook)
{
    static_for<3>([&](auto S) {
        constexpr int s = decltype(S)::value;
        constexpr int p0 = PAIR_SLOT[s][0], p1 = PAIR_SLOT[s][1];
        if constexpr (CHARGED_EMPTY && s < 2) {
            k1_species_empty<s, L, SHARE, true>(dv, gt, stash, m, o, c);
            return;
        }
        const D vx[3] = { m.ux[s], m.upx[p0], m.upx[p1] };
        const D vy[3] = { m.uy[s], m.upy[p0], m.upy[p1] };
        const D u2 = vx[0] * vx[0] + vy[0] * vy[0];                           // collisions.cpp:98-100
        // brackets of the pair velocities this species forms itself (all without SHARE; else see SharedBracket)
        constexpr bool comp1 = !SHARE || SharedBracket<s, 1>::computed, comp2 = !SHARE || SharedBracket<s, 2>::computed;
        D K[3] = { D(0.0), D(0.0), D(0.0) };                                  // u2*0.5*invcs2, plasma.cpp:199   (E3)
        K[0] = u2 * D(c.hinvcs2);
        if constexpr (comp1) K[1] = (vx[1] * vx[1] + vy[1] * vy[1]) * D(c.hinvcs2);
        if constexpr (comp2) K[2] = (vx[2] * vx[2] + vy[2] * vy[2]) * D(c.hinvcs2);
        D AB2[3];
        thermal_cell_terms<s>(m.rho[s], c, AB2);
        const D rhoh = D(0.5) * m.rho[s];
        D uE = D(0.0);
        if constexpr (s < 2) uE = vx[0] * Ex + vy[0] * Ey;                    // collisions.cpp:157,162
        D pref3[3] = { D(0.0), D(0.0), D(0.0) };                              // Guo prefactor per weight class, collisions.cpp:154,159
        if constexpr (s < 2) {
            #pragma unroll
            for (int wc = 0; wc < 3; ++wc) pref3[wc] = guo_prefactor<s>(dv, wc, m.rho[s], c);
        }
#if PLBM_K1_UNROLL
        static_for<5>([&](auto AX) {
            k1_axis_ct<s, decltype(AX)::value, L>(dv, gt, stash, m, vx, vy, K, AB2, rhoh, u2, uE, pref3, Ex, Ey, o, c);
        });
#else
        // four axes with two opposite directions each in a rolled loop (its body fits the instruction cache), then the rest
        // direction with everything resolved at compile time (c = 0: no dot products, bracket 1 - K)
        #pragma unroll 1
        for (int axis = 0; axis < 4; ++axis) {
            hook.template at_axis<s>(axis);
            const AxisSel sel = axis_select(axis);
            const D w = (axis < 2) ? D(c.w[1]) : D(c.w[2]);
            const D wr = w * m.rho[s];
            const D wT = w * m.T[s];
            BracketParts bp[3];
            bp[0] = bracket_parts(axis_dot(sel, vx[0], vy[0]), c);
            if constexpr (comp1) bp[1] = bracket_parts(axis_dot(sel, vx[1], vy[1]), c);
            if constexpr (comp2) bp[2] = bracket_parts(axis_dot(sel, vx[2], vy[2]), c);
            D pref = D(0.0), X = D(0.0), cE = D(0.0);
            if constexpr (s < 2) {
                cE = axis_dot(sel, Ex, Ey);
                pref = (axis < 2) ? pref3[1] : pref3[2];
                X = dv.cdiv(bp[0].cu * cE, c.cs2);                             // (c.u)(c.E)/cs2
            }
            auto direction = [&](auto NEG, const int dir, const D fv, const D gv) {
                constexpr bool neg = decltype(NEG)::value;
                D b[3];
                b[0] = bracket_value<neg>(bp[0], K[0]);
                if constexpr (comp1) b[1] = bracket_value<neg>(bp[1], K[1]);
                if constexpr (comp2) b[2] = bracket_value<neg>(bp[2], K[2]);
                if constexpr (SHARE) {
                    // fv, gv of this direction are in registers: their slots carry the pair brackets from here on
                    if constexpr (comp1) stash_st(stash + L::slot(SharedBracket<s, 1>::sk, dir), b[1]);
                    else b[1] = stash_ld(stash + L::slot(SharedBracket<s, 1>::sk, dir));
                    if constexpr (comp2) stash_st(stash + L::slot(SharedBracket<s, 2>::sk, dir), b[2]);
                    else b[2] = stash_ld(stash + L::slot(SharedBracket<s, 2>::sk, dir));
                }
                D force = D(0.0);
                if constexpr (s < 2) force = pref * guo_bracket<neg>(X, cE, uE);   // collisions.cpp:154-163
                D fnew, gnew;
                collide_species_dir<s>(dv, fv, gv, b, wr, wT, AB2, rhoh, u2, force, c, fnew, gnew);
                gt.note_output(fnew);
                gt.note_output(gnew);
                o.dst[((s * 2 + 0) * NQ + dir) * o.plane] = fnew.v;
                o.dst[((s * 2 + 1) * NQ + dir) * o.plane] = gnew.v;
            };
            const int d0 = (axis < 2) ? axis + 1 : axis + 3;                  // first direction of the axis; the opposite is d0 + 2
            // both directions of the axis in one straight-line block: six independent division chains
            D f0, g0, f1, g1;
            if constexpr (SHARE) {
                f0 = stash_ld(stash + L::slot(s * 2 + 0, d0));     g0 = stash_ld(stash + L::slot(s * 2 + 1, d0));
                f1 = stash_ld(stash + L::slot(s * 2 + 0, d0 + 2)); g1 = stash_ld(stash + L::slot(s * 2 + 1, d0 + 2));
            } else {
                f0 = D(stash[L::slot(s * 2 + 0, d0)]);     g0 = D(stash[L::slot(s * 2 + 1, d0)]);
                f1 = D(stash[L::slot(s * 2 + 0, d0 + 2)]); g1 = D(stash[L::slot(s * 2 + 1, d0 + 2)]);
            }
            direction(std::false_type{}, d0, f0, g0);
            direction(std::true_type{}, d0 + 2, f1, g1);
        }
        k1_axis_ct<s, 4, L, SHARE>(dv, gt, stash, m, vx, vy, K, AB2, rhoh, u2, uE, pref3, Ex, Ey, o, c);
#endif
    });
}

// Recomputation of one cell with the reference's literal arithmetic (literal_cell.cuh): taken when the gate trips (operands
// outside the fast divisions' domain, non-finite values).  f, g: the cell's 54 pulled populations.
template <bool WRITE_MACRO>
static __device__ __forceinline__ void k1_literal_body(const D (&f)[3][NQ], const D (&g)[3][NQ], double Ex, double Ey, double* dst, long long plane,
                                                       double* rho_q, const MacroOut* mo, long long cidx, const LbmConsts* c)
{
    LitMacro m;
    lit_update_macro(f, g, D(Ex), D(Ey), *c, m);
    *rho_q = m.rho_q.v;
    if (WRITE_MACRO) {
        for (int s = 0; s < 3; ++s) {
            mo->ux[s][cidx] = m.ux[s].v; mo->uy[s][cidx] = m.uy[s].v;
            mo->T[s][cidx] = m.T[s].v;   mo->rho[s][cidx] = m.rho[s].v;
        }
    }
    for (int i = 0; i < NQ; ++i) {
        const D fi[3] = { f[0][i], f[1][i], f[2][i] }, gi[3] = { g[0][i], g[1][i], g[2][i] };
        D fo[3], go[3];
        lit_collide_direction(i, fi, gi, m, D(Ex), D(Ey), *c, fo, go);
        for (int s = 0; s < 3; ++s) {
            dst[((s * 2 + 0) * NQ + i) * plane] = fo[s].v;
            dst[((s * 2 + 1) * NQ + i) * plane] = go[s].v;
        }
    }
}

// Out of line, from the populations still parked in the stash.
template <bool WRITE_MACRO, class L>
static __device__ __noinline__ void k1_cell_literal(const double* stash, double Ex, double Ey, double* dst, long long plane, double* rho_q,
                                                    const MacroOut* mo, long long cidx, const LbmConsts* c)
{
    // scalar arguments only (no pointer to the caller's K1Out): the fast path keeps its addresses in registers
    D f[3][NQ], g[3][NQ];
    for (int s = 0; s < 3; ++s)
        for (int i = 0; i < NQ; ++i) {
            f[s][i] = D(stash[L::slot(s * 2 + 0, i)]);
            g[s][i] = D(stash[L::slot(s * 2 + 1, i)]);
        }
    k1_literal_body<WRITE_MACRO>(f, g, Ex, Ey, dst, plane, rho_q, mo, cidx, c);
}

// Out of line, pulling the populations from the source planes again: the fast path of k1_tma_kernel has overwritten part of the
// stash with pair brackets (SharedBracket).  One CTA per K1_THREADS-cell tile of a row, as launched by k1_tma_launch.
template <bool WRITE_MACRO>
static __device__ __noinline__ void k1_cell_literal_pull(const double* src, const LbmGeom* gp, double Ex, double Ey, double* dst, double* rho_q,
                                                         const MacroOut* mo, const LbmConsts* c)
{
    const LbmGeom& g = *gp;
    const int x = blockIdx.x * K1_THREADS + threadIdx.x, y = blockIdx.y;
    const int xm = (x == 0) ? g.NX - 1 : x - 1;
    const int xp = (x == g.NX - 1) ? 0 : x + 1;
    int rm = y, rp = y + 2;                            // storage rows of y-1 and y+1
    if (g.wrap_y) {
        if (y == 0) rm = g.NYl;
        if (y == g.NYl - 1) rp = 1;
    }
    const long long r0o = (long long)(y + 1) * g.pitch, rmo = (long long)rm * g.pitch, rpo = (long long)rp * g.pitch;
    const long long off[NQ] = { r0o + x, r0o + xm, rmo + x, r0o + xp, rpo + x, rmo + xm, rmo + xp, rpo + xp, rpo + xm };
    D f[3][NQ], gg[3][NQ];
    for (int s = 0; s < 3; ++s)
        for (int i = 0; i < NQ; ++i) {
            f[s][i] = D(__ldg(src + (long long)((s * 2 + 0) * NQ + i) * g.plane + off[i]));
            gg[s][i] = D(__ldg(src + (long long)((s * 2 + 1) * NQ + i) * g.plane + off[i]));
        }
    k1_literal_body<WRITE_MACRO>(f, gg, Ex, Ey, dst, g.plane, rho_q, mo, (long long)y * g.NX + x, c);
}

// Gate, fast path and -- outside the gate -- the literal recomputation of one cell whose populations are parked at `stash`.
// The fallback is skipped when every population input is zero or NaN and a raw density is NaN (CellGate::all_nan): then all 54
// outputs are NaN on either path (each species is coupled to both others through the pair velocities) and the moments and
// rho_q are NaN or the exact zeros of an empty species.  This keeps a lattice that the reference's own dynamics have driven to
// NaN ("Divergence of the reference") at full speed.
template <bool WRITE_MACRO, class L, class Hook>
__device__ __forceinline__ void k1_cell_checked(const double* __restrict__ stash, double Ex, double Ey, const K1Out& o, const MacroOut* mo,
                                                const LbmConsts& c, Hook& hook)
{
    CellGate gt;
    GatedDiv dv;
    k1_cell_gated<WRITE_MACRO, L, false>(dv, gt, stash, D(Ex), D(Ey), o, c, hook);
    if (!gt.ok(dv.ok())) k1_cell_literal<WRITE_MACRO, L>(stash, Ex, Ey, o.dst, o.plane, o.rho_q, mo, o.cidx, &c);
}

// The same with pair brackets handed on through the stash (k1_tma_kernel): the fallback pulls the cell's populations again.
template <bool WRITE_MACRO, class L>
__device__ __forceinline__ void k1_cell_checked_share(double* stash, double Ex, double Ey, const K1Out& o, const MacroOut* mo,
                                                      const LbmConsts& c, const double* src, const LbmGeom* g)
{
    CellGate gt;
    GatedDiv dv;
    NoHook hook;
    k1_cell_gated<WRITE_MACRO, L, true>(dv, gt, stash, D(Ex), D(Ey), o, c, hook);
    if (!gt.ok(dv.ok())) k1_cell_literal_pull<WRITE_MACRO>(src, g, Ex, Ey, o.dst, o.rho_q, mo, &c);
}

#ifdef PLBM_K1_MAXNREG
#define PLBM_K1_BOUNDS __maxnreg__(PLBM_K1_MAXNREG)
#else
#define PLBM_K1_BOUNDS __launch_bounds__(K1_THREADS, PLBM_K1_MIN_BLOCKS)
#endif

// E = -grad(phi) formed exactly as poisson::ComputeElectricField_Periodic forms it (src/poisson.cpp:589-607): phi is [NYl][NX],
// below/above the neighbouring slabs' boundary rows (nullptr: wrap inside the array).  Saves the K3 sweep and 8 B per cell.
__device__ __forceinline__ void field_from_phi(const double* __restrict__ phi, const double* __restrict__ below_row,
                                               const double* __restrict__ above_row, int x, int y, int xm, int xp, const LbmGeom& g,
                                               double& Ex, double& Ey)
{
    const double* row = phi + (long long)y * g.NX;
    const double* below = (y > 0) ? row - g.NX : (below_row ? below_row : phi + (long long)(g.NYl - 1) * g.NX);
    const double* above = (y < g.NYl - 1) ? row + g.NX : (above_row ? above_row : phi);
    Ex = __dmul_rn(-0.5, __dsub_rn(__ldg(row + xp), __ldg(row + xm)));
    Ey = __dmul_rn(-0.5, __dsub_rn(__ldg(above + x), __ldg(below + x)));
}

// E_FROM_PHI: Exf is phi, Eyf/Ezf the boundary rows below/above (field_from_phi); else Exf, Eyf are the field arrays.
template <bool WRITE_MACRO, bool E_FROM_PHI>
__global__ void PLBM_K1_BOUNDS
k1_fused_kernel(const double* __restrict__ src, double* __restrict__ dst,
                const double* __restrict__ Exf, const double* __restrict__ Eyf, const double* __restrict__ Ezf,
                double* __restrict__ rho_q, const __grid_constant__ MacroOut mo,
                const __grid_constant__ LbmConsts c, const __grid_constant__ LbmGeom g)
{
    extern __shared__ double stash_all[];
    double* stash = stash_all + threadIdx.x;          // column of this thread
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

#if PLBM_K1_PREFETCH
    // The pull is one round trip to HBM that nothing in this CTA can hide.  Shorten it for a later CTA instead: thread p < 54 asks
    // L2 for this tile's segment of plane p in the row `g.prefetch_rows` further on, so that the CTA which pulls it finds it
    // in L2.  Every segment of every row is requested at most once per step.
    if (g.prefetch_rows > 0 && threadIdx.x < NPLANES) {
        int yp = y + g.prefetch_rows;
        if (yp >= g.NYl && g.wrap_y) yp -= g.NYl;
        if (yp < g.NYl) {
            const int x0 = blockIdx.x * K1_THREADS;
            const int n = min(K1_THREADS, g.pitch - x0);      // pitch is a multiple of 16 doubles: whole 128-byte lines
            const double* seg = src + (long long)threadIdx.x * g.plane + (long long)(yp + 1) * g.pitch + x0;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(seg), "r"(n * 8) : "memory");
        }
    }
#endif
    // (ptxas hoists all 54 loads above the first store: one round trip to memory per cell)
    #pragma unroll
    for (int sk = 0; sk < 2 * NSPEC; ++sk) {
        const double* p = src + (long long)(sk * NQ) * g.plane;
        double v[NQ];
        #pragma unroll
        for (int i = 0; i < NQ; ++i) v[i] = __ldg(p + i * g.plane + off[i]);
        #pragma unroll
        for (int i = 0; i < NQ; ++i) stash[ColumnLayout::slot(sk, i)] = v[i];
    }
    const long long cidx = (long long)y * g.NX + x;    // scalar fields are flat x + NX*y
    double Ex, Ey;
    if constexpr (E_FROM_PHI) field_from_phi(Exf, Eyf, Ezf, x, y, xm, xp, g, Ex, Ey);
    else { Ex = __ldg(Exf + cidx); Ey = __ldg(Eyf + cidx); }

    K1Out o;
    o.dst = dst + (long long)r0o + x;
    o.plane = g.plane;
    o.rho_q = rho_q + cidx;
    o.mo = mo;
    o.cidx = cidx;
    NoHook hook;
    k1_cell_checked<WRITE_MACRO, ColumnLayout>(stash, Ex, Ey, o, &mo, c, hook);
}


// K1 with bounce-back walls: same cell work, the pull at the top follows walls.cuh (streaming::StreamingBounceBack,
// src/streaming.cpp:66-112,150-196, restated as a pull); single slab only.
template <bool WRITE_MACRO>
__global__ void PLBM_K1_BOUNDS
k1_walls_kernel(const double* __restrict__ src, double* __restrict__ dst,
                const double* __restrict__ Exf, const double* __restrict__ Eyf,
                double* __restrict__ rho_q, const __grid_constant__ MacroOut mo,
                const __grid_constant__ LbmConsts c, const __grid_constant__ LbmGeom g, const WallArgs wa)
{
    extern __shared__ double stash_all[];
    double* stash = stash_all + threadIdx.x;          // column of this thread
    const int x = blockIdx.x * K1_THREADS + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= g.NX) return;

    const int r0o = (y + 1) * g.pitch;
    {
        // per destination slot: the cell and direction whose post-collision value lands here, or nothing (stale)
        const bool rim = (x == 0 || x == g.NX - 1 || y == 0 || y == g.NYl - 1);
        int soff[NQ], sdir[NQ];
        unsigned stale = 0u;
        #pragma unroll
        for (int j = 0; j < NQ; ++j) {
            soff[j] = r0o + x; sdir[j] = j;
            if (!wa.identity) {
                if (!rim) soff[j] = (y + 1 - WALL_CY[j]) * g.pitch + (x - WALL_CX[j]);
                else {
                    const WallSource ws = wall_source(x, y, j, g.NX, g.NYl);
                    if (ws.dir < 0) stale |= 1u << j;
                    else { soff[j] = (ws.y + 1) * g.pitch + ws.x; sdir[j] = ws.dir; }
                }
            }
        }
        const int ridx = rim ? wall_rim_index(x, y, g.NX, g.NYl) : 0;
        #pragma unroll
        for (int sk = 0; sk < 2 * NSPEC; ++sk) {
            const double* p = src + (long long)(sk * NQ) * g.plane;
            double v[NQ];
            #pragma unroll
            for (int i = 0; i < NQ; ++i) {
                if (stale & (1u << i)) {
                    // a slot nobody streams into keeps what the reference's shared temp array held: for f its own
                    // pre-collision value, for g the post-collision f of the same species (walls.cuh)
                    v[i] = (sk & 1) ? __ldg(src + (long long)((sk - 1) * NQ + i) * g.plane + r0o + x)
                                    : __ldg(wa.rim + ((long long)((sk >> 1) * NQ + i)) * wa.nrim + ridx);
                } else {
                    v[i] = __ldg(p + (long long)sdir[i] * g.plane + soff[i]);
                }
            }
            #pragma unroll
            for (int i = 0; i < NQ; ++i) stash[ColumnLayout::slot(sk, i)] = v[i];
            if ((sk & 1) == 0 && rim) {                // this step's pre-collision f: the stale values of the next pull
                #pragma unroll
                for (int i = 0; i < NQ; ++i) wa.rim_next[((long long)((sk >> 1) * NQ + i)) * wa.nrim + ridx] = v[i];
            }
        }
    }
    const long long cidx = (long long)y * g.NX + x;    // scalar fields are flat x + NX*y
    const double Ex = __ldg(Exf + cidx), Ey = __ldg(Eyf + cidx);          // walls: the field always comes from the arrays

    K1Out o;
    o.dst = dst + (long long)r0o + x;
    o.plane = g.plane;
    o.rho_q = rho_q + cidx;
    o.mo = mo;
    o.cidx = cidx;
    NoHook hook;
    k1_cell_checked<WRITE_MACRO, ColumnLayout>(stash, Ex, Ey, o, &mo, c, hook);
}

static __host__ __device__ constexpr size_t k1_smem_bytes() { return sizeof(double) * NPLANES * K1_THREADS; }


// ---- K1 with persistent warps and a TMA-filled pool of stash buffers -----------------------------------------------------------
// What limits k1_fused_kernel once the instruction count is down: every CTA starts with one round trip to HBM that its two warps
// can only wait for (long_scoreboard 0.68 per issue = 13 % of the warp time at 12 warps per SM, profiles/r2_k1_sweeps.md), and
// a second stash per CTA to pull ahead does not fit in shared memory.  Here the warps are persistent and independent:
//   * a tile is 32 consecutive cells of a row = one warp; warp gw of the grid runs tiles gw, gw + G, gw + 2G, ...
//   * the SM's shared memory is a POOL of POOL_NBUF stash buffers for POOL_WARPS warps (15 for 12): a warp computes out of one
//     buffer and, before its last species (72 % through the tile), takes a free buffer and has the TMA engine fill it with
//     its NEXT tile; at the end of the tile it returns the old buffer.  With the three spare buffers held for the last quarter
//     of a tile each, most warps never wait for memory; a warp that finds no free buffer reuses its own and waits as before.
//   * the planes are ONE 4-D tensor [sk][direction][storage row][x] (strides 9*plane, plane, pitch, 1; x extent NX, so the
//     padding of a row and the cells beyond the periodic edge are out of bounds = zero-filled).  The pull of direction i is a box
//     {34, 1, 1, 6} at row(y - cy_i): FP64 boxes must start at an even x (16-byte alignment; an odd start raises an illegal
//     instruction on B200, scratch/tma/tma_test2.cu), so the box starts at x0 - 2 for c_x = +1 and at x0 otherwise, and a
//     thread reads column lane + 1 when c_x != 0 (PoolLayout).  One lane issues the nine boxes of a tile onto one mbarrier.
//   * the periodic wrap in x is not expressible as a box: the cells x = 0 and x = NX-1 fetch their three wrapped neighbours per
//     distribution themselves after the wait.
// No LDG/STS, address arithmetic or staging registers for the pull, no CTA-wide barrier after start-up, nothing through L1.
#ifndef PLBM_POOL_WARPS
#define PLBM_POOL_WARPS 12
#endif
constexpr int POOL_WARPS = PLBM_POOL_WARPS;                     // 12 warps x 168 registers = the SM's register file
constexpr int POOL_THREADS = POOL_WARPS * 32;
constexpr int POOL_NBUF = 15;                                   // 15 x 14 976 B + barriers <= 227 KB
constexpr unsigned POOL_TILE_BYTES = NQ * (2 * NSPEC) * POOL_BOX_W * sizeof(double);   // bytes the nine boxes deliver

// Loop state of one warp.  It lives in shared memory, not in registers: the cell code needs all 168 registers a thread can have
// at 12 warps per SM, and anything else kept live across it is paid for in recomputed FP64 products.
struct PoolWarp {
    int next_tile;                                              // the tile after the current one, -1: none
    int next_buf;                                               // buffer that tile is being pulled into, -1: none yet
    unsigned next_parity;                                       // parity of that buffer's barrier
    int pad;
    double Ex[POOL_TILE], Ey[POOL_TILE];                        // field at the cells of the next tile
};
struct PoolShared {
    alignas(128) double buf[POOL_NBUF][POOL_BUF_DOUBLES];
    alignas(8) unsigned long long full[POOL_NBUF];              // mbarrier per buffer: the tile has landed
    PoolWarp warp[POOL_WARPS];
    unsigned fills[POOL_NBUF];                                  // how often the buffer has been filled (parity of its barrier)
    unsigned free_mask;                                         // bit b: buffer b is free
    // kernel arguments the out-of-line helper needs
    const CUtensorMap* tmap;
    const LbmGeom* g;
    const double *Exf, *Eyf, *Ezf;
    int e_from_phi, tiles_per_row;
};
static __host__ __device__ constexpr size_t k1_pool_smem_bytes() { return sizeof(PoolShared) + 128; }

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_load_box4(void* dst, const void* tmap, int c0, int c1, int c2, int c3, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 :: "r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}

__device__ __forceinline__ PoolShared& pool_shared()
{
    extern __shared__ unsigned char pool_raw[];
    return *reinterpret_cast<PoolShared*>((reinterpret_cast<uintptr_t>(pool_raw) + 127) & ~(uintptr_t)127);
}

// One lane: arm buffer b's barrier and have the TMA engine pull tile `tile` into it.  Returns the parity to wait for.
__device__ __forceinline__ unsigned pool_issue(PoolShared& sh, int b, int tile)
{
    const LbmGeom& g = *sh.g;
    const int y = tile / sh.tiles_per_row, x0 = (tile - y * sh.tiles_per_row) * POOL_TILE;
    int rm = y, rp = y + 2;                            // storage rows of y-1 and y+1 (storage row = y + 1)
    if (g.wrap_y) {
        if (y == 0) rm = g.NYl;
        if (y == g.NYl - 1) rp = 1;
    }
    const unsigned parity = sh.fills[b] & 1u;
    sh.fills[b] += 1u;
    // order this thread's earlier generic-proxy accesses of the buffer (the previous tile's fix-up stores) before the engine's writes
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&sh.full[b], POOL_TILE_BYTES);
    // direction i pulls from (x - cx_i, y - cy_i): cx = 0,1,0,-1,0,1,-1,-1,1   cy = 0,0,1,0,-1,1,1,-1,-1; box start x0 - 2 where cx = +1
    const int bx[NQ] = { x0, x0 - 2, x0, x0, x0, x0 - 2, x0, x0, x0 - 2 };
    const int br[NQ] = { y + 1, y + 1, rm, y + 1, rp, rm, rm, rp, rp };
    #pragma unroll
    for (int i = 0; i < NQ; ++i) tma_load_box4(&sh.buf[b][i * POOL_DIR_STRIDE], sh.tmap, bx[i], br[i], i, 0, &sh.full[b]);
    return parity;
}

// One lane: take a free buffer (-1: none).
__device__ __forceinline__ int pool_try_acquire(PoolShared& sh)
{
    unsigned m = *(volatile unsigned*)&sh.free_mask;
    while (m) {
        const int b = __ffs(m) - 1;
        const unsigned seen = atomicCAS(&sh.free_mask, m, m & ~(1u << b));
        if (seen == m) { __threadfence_block(); return b; }
        m = seen;
    }
    return -1;
}

// The field at this lane's cell of tile `tile`, parked for the tile's start.
__device__ __forceinline__ void pool_load_field(PoolShared& sh, PoolWarp& w, int tile)
{
    const LbmGeom& g = *sh.g;
    const int lane = threadIdx.x & 31;
    const int y = tile / sh.tiles_per_row, x = (tile - y * sh.tiles_per_row) * POOL_TILE + lane;
    if (x < g.NX) {
        double Ex, Ey;
        const int xm = (x == 0) ? g.NX - 1 : x - 1, xp = (x == g.NX - 1) ? 0 : x + 1;
        if (sh.e_from_phi) field_from_phi(sh.Exf, sh.Eyf, sh.Ezf, x, y, xm, xp, g, Ex, Ey);
        else { const long long cidx = (long long)y * g.NX + x; Ex = __ldg(sh.Exf + cidx); Ey = __ldg(sh.Eyf + cidx); }
        w.Ex[lane] = Ex; w.Ey[lane] = Ey;
    }
}

// Asks for the warp's next tile: a free buffer, the nine boxes, and the field around the next cells.  Out of line and without
// arguments (everything comes from shared memory), so that the call site inside the collision loop costs no registers.
#ifndef PLBM_POOL_ASK_AXIS
#define PLBM_POOL_ASK_AXIS 2
#endif
static __device__ __noinline__ void pool_ask()
{
    PoolShared& sh = pool_shared();
    PoolWarp& w = sh.warp[threadIdx.x >> 5];
    const int tile = w.next_tile;
    if (tile < 0) return;
    if ((threadIdx.x & 31) == 0) {
        const int b = pool_try_acquire(sh);
        if (b >= 0) { w.next_parity = pool_issue(sh, b, tile); w.next_buf = b; }
    }
    pool_load_field(sh, w, tile);
}

// Called at the top of every axis iteration: late in the current tile (axis PLBM_POOL_ASK_AXIS of the last species, 80-90 %
// through it) the warp asks for its next tile.  A spare buffer is then held for the last 10-20 % of a tile, so the three spares
// serve the twelve warps (12 x 0.15 = 1.8 in use on average) and the ~2 us that remain cover the round trip to HBM.
struct PoolHook {
    template <int S> __device__ __forceinline__ void at_axis(int axis)
    {
        if constexpr (S == 2) {
            if (axis == PLBM_POOL_ASK_AXIS) pool_ask();
        }
    }
};

// ---- K1 with the pull done by the TMA engine, one CTA per tile (no persistence) ------------------------------------------------
// Same tiles, grid and cell code as k1_fused_kernel; the 54 LDG, 54 STS and their address arithmetic are replaced by nine boxes
// {K1_THREADS + 2, 1, 1, 6} that one thread issues onto one mbarrier (box start x0 - 2 where c_x = +1, else x0: FP64 boxes must start
// at an even x).  The field is loaded while the boxes are in flight.
static __host__ __device__ constexpr size_t k1_tma_smem_bytes() { return sizeof(double) * NQ * TMA_DIR_STRIDE + 16; }
constexpr unsigned TMA_TILE_BYTES = NQ * (2 * NSPEC) * TMA_BOX_W * sizeof(double);

template <bool WRITE_MACRO, bool E_FROM_PHI>
__global__ void PLBM_K1_BOUNDS
k1_tma_kernel(const __grid_constant__ CUtensorMap tmap, const double* __restrict__ src, double* __restrict__ dst,
              const double* __restrict__ Exf, const double* __restrict__ Eyf, const double* __restrict__ Ezf,
              double* __restrict__ rho_q, const __grid_constant__ MacroOut mo,
              const __grid_constant__ LbmConsts c, const __grid_constant__ LbmGeom g)
{
    extern __shared__ __align__(128) double stash_all[];
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(stash_all + NQ * TMA_DIR_STRIDE);
    const int x0 = blockIdx.x * K1_THREADS;
    const int x = x0 + threadIdx.x;
    const int y = blockIdx.y;
    int rm = y, rp = y + 2;                            // storage rows of y-1 and y+1 (storage row = y + 1)
    if (g.wrap_y) {
        if (y == 0) rm = g.NYl;
        if (y == g.NYl - 1) rp = 1;
    }
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar, TMA_TILE_BYTES);
        const int bx[NQ] = { x0, x0 - 2, x0, x0, x0, x0 - 2, x0, x0, x0 - 2 };
        const int br[NQ] = { y + 1, y + 1, rm, y + 1, rp, rm, rm, rp, rp };
        #pragma unroll
        for (int i = 0; i < NQ; ++i) tma_load_box4(stash_all + i * TMA_DIR_STRIDE, &tmap, bx[i], br[i], i, 0, bar);
    }
    // (no L2 prefetch here: with the TMA pull it costs 1 %, profiles/r2_k1_sweeps.md)
    __syncthreads();                                   // the barrier's initialisation is visible to the waiting threads
    if (x >= g.NX) return;
    const int xm = (x == 0) ? g.NX - 1 : x - 1;
    const int xp = (x == g.NX - 1) ? 0 : x + 1;
    const long long cidx = (long long)y * g.NX + x;
    double Ex, Ey;
    if constexpr (E_FROM_PHI) field_from_phi(Exf, Eyf, Ezf, x, y, xm, xp, g, Ex, Ey);
    else { Ex = __ldg(Exf + cidx); Ey = __ldg(Eyf + cidx); }
    const int r0o = (y + 1) * g.pitch;
    double* stash = stash_all + threadIdx.x;

    mbar_wait(bar, 0);
    if (x == 0 || x == g.NX - 1) {
        // periodic wrap in x: the box delivered zeros for the neighbour outside [0, NX)
        const int rmo = rm * g.pitch, rpo = rp * g.pitch;
        #pragma unroll 1
        for (int sk = 0; sk < 2 * NSPEC; ++sk) {
            const double* p = src + (long long)(sk * NQ) * g.plane;
            if (x == 0) {
                stash[TmaLayout::slot(sk, 1)] = __ldg(p + 1 * g.plane + r0o + xm);
                stash[TmaLayout::slot(sk, 5)] = __ldg(p + 5 * g.plane + rmo + xm);
                stash[TmaLayout::slot(sk, 8)] = __ldg(p + 8 * g.plane + rpo + xm);
            }
            if (x == g.NX - 1) {
                stash[TmaLayout::slot(sk, 3)] = __ldg(p + 3 * g.plane + r0o + xp);
                stash[TmaLayout::slot(sk, 6)] = __ldg(p + 6 * g.plane + rmo + xp);
                stash[TmaLayout::slot(sk, 7)] = __ldg(p + 7 * g.plane + rpo + xp);
            }
        }
    }
    K1Out o;
    o.dst = dst + (long long)r0o + x;
    o.plane = g.plane;
    o.rho_q = rho_q + cidx;
    o.mo = mo;
    o.cidx = cidx;
#if PLBM_K1_SHARE
    k1_cell_checked_share<WRITE_MACRO, TmaLayout>(stash, Ex, Ey, o, &mo, c, src, &g);
#else
    NoHook hook;
    k1_cell_checked<WRITE_MACRO, TmaLayout>(stash, Ex, Ey, o, &mo, c, hook);
#endif
}

template <bool WRITE_MACRO, bool E_FROM_PHI>
__global__ void __launch_bounds__(POOL_THREADS, 1)
k1_pool_kernel(const __grid_constant__ CUtensorMap tmap, const double* __restrict__ src, double* __restrict__ dst,
               const double* __restrict__ Exf, const double* __restrict__ Eyf, const double* __restrict__ Ezf,
               double* __restrict__ rho_q, const __grid_constant__ MacroOut mo,
               const __grid_constant__ LbmConsts c, const __grid_constant__ LbmGeom g)
{
    PoolShared& sh = pool_shared();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per_row = (g.NX + POOL_TILE - 1) / POOL_TILE;
    const int count = per_row * g.NYl;
    const int G = gridDim.x * POOL_WARPS;

    if (threadIdx.x == 0) {
        for (int b = 0; b < POOL_NBUF; ++b) { mbar_init(&sh.full[b], 1); sh.fills[b] = 0u; }
        sh.free_mask = ((1u << POOL_NBUF) - 1u) & ~((1u << POOL_WARPS) - 1u);      // warp w starts with buffer w
        sh.tmap = &tmap; sh.g = &g; sh.Exf = Exf; sh.Eyf = Eyf; sh.Ezf = Ezf;
        sh.e_from_phi = E_FROM_PHI ? 1 : 0; sh.tiles_per_row = per_row;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    int tile = blockIdx.x * POOL_WARPS + warp;         // consecutive tiles run on consecutive warps of consecutive SMs
    if (tile >= count) return;
    int buf = warp;
    unsigned parity = 0u;
    PoolWarp& ws = sh.warp[warp];
    if (lane == 0) parity = pool_issue(sh, buf, tile);
    pool_load_field(sh, ws, tile);

    #pragma unroll 1
    for (;;) {
        const int y = tile / per_row, x0 = (tile - y * per_row) * POOL_TILE;
        const int x = x0 + lane;
        const bool live = x < g.NX;
        const int xc = live ? x : g.NX - 1;            // lanes beyond the end of the row idle; keep their addresses valid
        const long long cidx = (long long)y * g.NX + xc;
        const int r0o = (y + 1) * g.pitch;
        double* stash = &sh.buf[buf][0] + lane;
        __syncwarp();
        const double Ex = ws.Ex[lane], Ey = ws.Ey[lane];  // loaded while the previous tile was computed (or before the loop)
        __syncwarp();
        if (lane == 0) { ws.next_tile = (tile + G < count) ? tile + G : -1; ws.next_buf = -1; }

        parity = __shfl_sync(0xffffffffu, parity, 0);
        mbar_wait(&sh.full[buf], parity);
        if (live && (x == 0 || x == g.NX - 1)) {
            // periodic wrap in x: the box delivered zeros for the neighbour outside [0, NX)
            const int xm = (x == 0) ? g.NX - 1 : x - 1, xp = (x == g.NX - 1) ? 0 : x + 1;
            int rm = y, rp = y + 2;
            if (g.wrap_y) {
                if (y == 0) rm = g.NYl;
                if (y == g.NYl - 1) rp = 1;
            }
            const int rmo = rm * g.pitch, rpo = rp * g.pitch;
            #pragma unroll 1
            for (int sk = 0; sk < 2 * NSPEC; ++sk) {
                const double* p = src + (long long)(sk * NQ) * g.plane;
                if (x == 0) {
                    stash[PoolLayout::slot(sk, 1)] = __ldg(p + 1 * g.plane + r0o + xm);
                    stash[PoolLayout::slot(sk, 5)] = __ldg(p + 5 * g.plane + rmo + xm);
                    stash[PoolLayout::slot(sk, 8)] = __ldg(p + 8 * g.plane + rpo + xm);
                }
                if (x == g.NX - 1) {
                    stash[PoolLayout::slot(sk, 3)] = __ldg(p + 3 * g.plane + r0o + xp);
                    stash[PoolLayout::slot(sk, 6)] = __ldg(p + 6 * g.plane + rmo + xp);
                    stash[PoolLayout::slot(sk, 7)] = __ldg(p + 7 * g.plane + rpo + xp);
                }
            }
        }
        __syncwarp();                                  // next_tile is set before any lane can reach pool_ask

        if (live) {
            K1Out o;
            o.dst = dst + (long long)r0o + x;
            o.plane = g.plane;
            o.rho_q = rho_q + cidx;
            o.mo = mo;
            o.cidx = cidx;
            PoolHook hook;
            k1_cell_checked<WRITE_MACRO, PoolLayout>(stash, Ex, Ey, o, &mo, c, hook);
        }
        __syncwarp();
        // lane 0 (always live: x0 < NX) has been through pool_ask
        const int next = ws.next_tile;
        if (next < 0) break;
        const int nb = ws.next_buf;
        if (nb >= 0) {
            if (lane == 0) { __threadfence_block(); atomicOr(&sh.free_mask, 1u << buf); }    // every lane has finished reading `buf`
            buf = nb;
            parity = ws.next_parity;
        } else {
            if (lane == 0) parity = pool_issue(sh, buf, next);                              // no spare buffer was free: pull into my own
        }
        tile = next;
    }
}

} // namespace plbm
