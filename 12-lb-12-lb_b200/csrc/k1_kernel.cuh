// k1_kernel.cuh -- K1: one pass per time step over all three species, f and g (kernel template; instantiated by
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
// Algorithmic traffic: 54*8 read + 54*8 written + Ex,Ey 16 + rho_q 8 = 888 B per cell and step.
//
// One thread owns one cell.  The 54 pulled populations are parked in shared memory (one column per
// thread, conflict-free).  The cell is then processed by k1_cell<FastDiv>: moments, then a ROLLED
// loop over the five direction axes (two opposite directions each) so that the ~2.7 k FP64
// instructions per cell come from a loop body that fits the instruction cache.  FastDiv records
// whether every division stayed inside the domain where its 2/3/9-instruction sequence is exact;
// if not (never on physical data), the cell is recomputed out of line with the reference's literal
// arithmetic (literal_cell.cuh).
#pragma once
#include <cuda.h>
#include "k1_fused.h"
#include "lbm_cell.cuh"
#include "literal_cell.cuh"
#include "walls.cuh"

namespace plbm {

#ifndef PLBM_K1_THREADS
#define PLBM_K1_THREADS 64
#endif
#ifndef PLBM_K1_PAIR
#define PLBM_K1_PAIR 1              // 1: both directions of an axis in one straight-line block
#endif
#ifndef PLBM_K1_GATE
#define PLBM_K1_GATE 1              // 1: per-cell gate (CellGate); 0: a validity record per division (FastDiv)
#endif
#ifndef PLBM_K1_MIN_BLOCKS
#define PLBM_K1_MIN_BLOCKS 6
#endif
#ifndef PLBM_K1_CARVEOUT
#define PLBM_K1_CARVEOUT (-1)       // preferred shared-memory carve-out in percent; -1: leave the choice to the driver
#endif
#ifndef PLBM_K1_PREFETCH
#define PLBM_K1_PREFETCH 1          // 1: every CTA asks L2 for the row segments a CTA one wave later will pull
#endif
#ifndef PLBM_K1_LDG
#define PLBM_K1_LDG 0               // how the populations are read: 0 ld.global.nc, 1 ld.global.cg (L2 only), 2 ld.global.cs (streaming)
#endif
constexpr int K1_THREADS = PLBM_K1_THREADS;

// Parked populations: slot of (species*2 + kind, direction) in the CTA's stash, one column per thread.  Direction-major, so that
// the six distributions that share a pull offset are one contiguous TMA box (k1_tma_kernel below).
__host__ __device__ constexpr int stash_slot(int sk, int dir) { return (dir * (2 * NSPEC) + sk) * K1_THREADS; }

__device__ __forceinline__ double k1_load(const double* p)
{
#if PLBM_K1_LDG == 1
    return __ldcg(p);
#elif PLBM_K1_LDG == 2
    return __ldcs(p);
#else
    return __ldg(p);
#endif
}

struct K1Out {
    double* __restrict__ dst;     // population planes, already offset to this cell
    long long plane;
    double* rho_q;                // already offset to this cell
    MacroOut mo;                  // base pointers
    long long cidx;
};

template <class DV, bool WRITE_MACRO>
__device__ __forceinline__ void k1_cell(DV& dv, const double* __restrict__ stash, D Ex, D Ey, const K1Out& o, const LbmConsts& c)
{
    // ---- UpdateMacro ------------------------------------------------------------------------
    CellMacro m;
    {
        D rl[3], mx[3], my[3], tl[3];
        static_for<3>([&](auto S) {
            constexpr int s = decltype(S)::value;
            D f[NQ], g[NQ];
            #pragma unroll
            for (int i = 0; i < NQ; ++i) {
                f[i] = D(stash[stash_slot(s * 2 + 0, i)]);
                g[i] = D(stash[stash_slot(s * 2 + 1, i)]);
            }
            rl[s] = sum9(f); mx[s] = moment_x(f); my[s] = moment_y(f); tl[s] = sum9(g);
        });
        cell_update_macro(dv, rl, mx, my, tl, Ex, Ey, c, m);
    }
    *o.rho_q = m.rho_q.v;
    if constexpr (WRITE_MACRO) {
        static_for<3>([&](auto S) {
            constexpr int s = decltype(S)::value;
            o.mo.ux[s][o.cidx] = m.ux[s].v; o.mo.uy[s][o.cidx] = m.uy[s].v;
            o.mo.T[s][o.cidx] = m.T[s].v;   o.mo.rho[s][o.cidx] = m.rho[s].v;
        });
    }

    // ---- collisions, species by species (small live state per species) -----------------------
    // Every species needs three equilibrium velocities: its own and those of its two pairs
    // (plasma.cpp:195-304).  Directions: axis 0..3 carry two opposite directions, axis 4 is rest.
    static_for<3>([&](auto S) {
        constexpr int s = decltype(S)::value;
        constexpr int p0 = PAIR_SLOT[s][0], p1 = PAIR_SLOT[s][1];
        const D vx[3] = { m.ux[s], m.upx[p0], m.upx[p1] };
        const D vy[3] = { m.uy[s], m.upy[p0], m.upy[p1] };
        const D u2 = vx[0] * vx[0] + vy[0] * vy[0];                           // collisions.cpp:98-100
        D K[3];                                                               // u2*0.5*invcs2, plasma.cpp:199   (E3)
        K[0] = u2 * D(c.hinvcs2);
        K[1] = (vx[1] * vx[1] + vy[1] * vy[1]) * D(c.hinvcs2);
        K[2] = (vx[2] * vx[2] + vy[2] * vy[2]) * D(c.hinvcs2);
        D AB2[3];
        thermal_cell_terms<s>(m.rho[s], c, AB2);
        const D rhoh = D(0.5) * m.rho[s];
        D uE = D(0.0);
        if constexpr (s < 2) uE = vx[0] * Ex + vy[0] * Ey;                    // collisions.cpp:157,162

        // Guo prefactor per weight class (4/9, 1/9, 1/36), collisions.cpp:154,159 -- three independent chains
        D pref3[3] = { D(0.0), D(0.0), D(0.0) };
        if constexpr (s < 2) {
            #pragma unroll
            for (int wc = 0; wc < 3; ++wc) pref3[wc] = guo_prefactor<s>(dv, wc, m.rho[s], c);
        }

        #pragma unroll 1
        for (int axis = 0; axis < 5; ++axis) {
            const int wclass = (axis < 2) ? 1 : (axis < 4 ? 2 : 0);
            const AxisSel sel = axis_select(axis);
            const D wr = D(c.w[wclass]) * m.rho[s];
            const D wT = D(c.w[wclass]) * m.T[s];
            BracketParts bp[3];
            #pragma unroll
            for (int j = 0; j < 3; ++j) bp[j] = bracket_parts(axis_dot(sel, vx[j], vy[j]), c);
            D pref = D(0.0), X = D(0.0), cE = D(0.0);
            if constexpr (s < 2) {
                cE = axis_dot(sel, Ex, Ey);
                pref = (wclass == 1) ? pref3[1] : (wclass == 2 ? pref3[2] : pref3[0]);
                X = dv.cdiv(bp[0].cu * cE, c.cs2);                             // (c.u)(c.E)/cs2
            }
            // one direction of this axis: NEG = false is the axis' first direction, true its opposite
            auto direction = [&](auto NEG, const int dir, const D fv, const D gv) {
                constexpr bool neg = decltype(NEG)::value;
                D b[3];
                #pragma unroll
                for (int j = 0; j < 3; ++j) b[j] = bracket_value<neg>(bp[j], K[j]);
                D force = D(0.0);
                if constexpr (s < 2) force = pref * guo_bracket<neg>(X, cE, uE);   // collisions.cpp:154-163
                D fnew, gnew;
                collide_species_dir<s>(dv, fv, gv, b, wr, wT, AB2, rhoh, u2, force, c, fnew, gnew);
                dv.note_output(fnew);
                dv.note_output(gnew);
                o.dst[((s * 2 + 0) * NQ + dir) * o.plane] = fnew.v;
                o.dst[((s * 2 + 1) * NQ + dir) * o.plane] = gnew.v;
            };
            const int d0 = (axis == 4) ? 0 : (axis < 2 ? axis + 1 : axis + 3);   // first direction of the axis; the opposite is d0 + 2
#if PLBM_K1_PAIR
            // both directions of the axis in one straight-line block: six independent division chains
            const D f0 = D(stash[stash_slot(s * 2 + 0, d0)]), g0 = D(stash[stash_slot(s * 2 + 1, d0)]);
            if (axis != 4) {
                const D f1 = D(stash[stash_slot(s * 2 + 0, d0 + 2)]), g1 = D(stash[stash_slot(s * 2 + 1, d0 + 2)]);
                direction(std::false_type{}, d0, f0, g0);
                direction(std::true_type{}, d0 + 2, f1, g1);
            } else {
                direction(std::false_type{}, d0, f0, g0);
            }
#else
            const int nsign = (axis == 4) ? 1 : 2;
            #pragma unroll 1
            for (int sg = 0; sg < nsign; ++sg) {
                const int dir = d0 + 2 * sg;
                const D fv = D(stash[stash_slot(s * 2 + 0, dir)]);
                const D gv = D(stash[stash_slot(s * 2 + 1, dir)]);
                if (sg) direction(std::true_type{}, dir, fv, gv);
                else direction(std::false_type{}, dir, fv, gv);
            }
#endif
        }
    });
}

// ---- the same cell with the per-cell gate (exact_math.cuh: GatedDiv + CellGate) -------------------------------------------
// No record per division: the gate notes the inputs (in the kernel's load loop), the macroscopic quantities and the
// outputs, and DESIGN.md ("K1 gate") proves every fast division of the cell exact from those.  With PLBM_K1_UNROLL the five
// direction axes are compile-time constants: c.v is a move, an addition or a subtraction, the weight class, the store offsets and
// the rest direction's bracket 1 - K are resolved by the compiler, and nothing is selected or masked at run time.
#ifndef PLBM_K1_UNROLL
#define PLBM_K1_UNROLL 0
#endif
// With the axes unrolled the collision code of one cell is ~60 KB of straight-line instructions, twice the SM's instruction
// cache: warps that drift apart fetch it from L2 again and again (measured: slower than the rolled loop although it issues
// 20 % fewer instructions).  PLBM_K1_SYNC keeps the warps of a CTA together: 1 = a CTA barrier after every species, 2 = after
// every direction axis, so that a line of code is fetched once per CTA, not once per warp.  TMA kernel only (all threads stay alive).
#ifndef PLBM_K1_SYNC
#define PLBM_K1_SYNC 0
#endif
template <int LEVEL>
__device__ __forceinline__ void k1_sync_point()
{
#if PLBM_K1_SYNC
    if constexpr (PLBM_K1_SYNC >= LEVEL) __syncthreads();
#endif
}

// Second read of a parked population (the first was for the moments).  Volatile so that the compiler really reads shared
// memory again instead of keeping all 54 values alive in registers and local memory across the moments (it spills otherwise).
__device__ __forceinline__ D stash_reload(const double* p)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return D(v);
}

// c_axis . v for the first direction of a compile-time axis (lbm_cell.cuh: axis_select)            (E1)
template <int AXIS>
__device__ __forceinline__ D axis_dot_ct(D vx, D vy)
{
    if constexpr (AXIS == 0) return vx;
    else if constexpr (AXIS == 1) return vy;
    else if constexpr (AXIS == 2) return vx + vy;
    else { static_assert(AXIS == 3, "axis 4 is the rest direction"); return vy - vx; }
}

template <int S, int AXIS>
__device__ __forceinline__ void k1_axis_ct(GatedDiv& dv, CellGate& gt, const double* __restrict__ stash, const CellMacro& m,
                                           const D (&vx)[3], const D (&vy)[3], const D (&K)[3], const D (&AB2)[3], D rhoh, D u2,
                                           D uE, const D (&pref3)[3], D Ex, D Ey, const K1Out& o, const LbmConsts& c)
{
    constexpr int wclass = (AXIS < 2) ? 1 : (AXIS < 4 ? 2 : 0);
    constexpr int d0 = (AXIS == 4) ? 0 : (AXIS < 2 ? AXIS + 1 : AXIS + 3);    // first direction of the axis; the opposite is d0 + 2
    const D wr = D(c.w[wclass]) * m.rho[S];
    const D wT = D(c.w[wclass]) * m.T[S];
    auto finish = [&](const int dir, const D (&b)[3], D force) {
        const D fv = stash_reload(stash + stash_slot(S * 2 + 0, dir)), gv = stash_reload(stash + stash_slot(S * 2 + 1, dir));
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
        #pragma unroll
        for (int j = 0; j < 3; ++j) b[j] = D(1.0) - K[j];
        D force = D(0.0);
        if constexpr (S < 2) force = pref3[0] * (D(0.0) - uE);
        finish(d0, b, force);
    } else {
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

template <bool WRITE_MACRO>
__device__ __forceinline__ void k1_cell_gated(GatedDiv& dv, CellGate& gt, const double* __restrict__ stash, D Ex, D Ey, const K1Out& o,
                                              const LbmConsts& c)
{
    // ---- UpdateMacro ------------------------------------------------------------------------
    CellMacro m;
    {
        D rl[3], mx[3], my[3], tl[3];
        static_for<3>([&](auto S) {
            constexpr int s = decltype(S)::value;
            D f[NQ], g[NQ];
            #pragma unroll
            for (int i = 0; i < NQ; ++i) {
                f[i] = D(stash[stash_slot(s * 2 + 0, i)]);
                g[i] = D(stash[stash_slot(s * 2 + 1, i)]);
            }
            #pragma unroll
            for (int i = 0; i < NQ; ++i) { gt.note_input(f[i].v); gt.note_input(g[i].v); }
            rl[s] = sum9(f); mx[s] = moment_x(f); my[s] = moment_y(f); tl[s] = sum9(g);
        });
        cell_update_macro(dv, gt, rl, mx, my, tl, Ex, Ey, c, m);
    }
    gt.close_macro();
    *o.rho_q = m.rho_q.v;
    if constexpr (WRITE_MACRO) {
        static_for<3>([&](auto S) {
            constexpr int s = decltype(S)::value;
            o.mo.ux[s][o.cidx] = m.ux[s].v; o.mo.uy[s][o.cidx] = m.uy[s].v;
            o.mo.T[s][o.cidx] = m.T[s].v;   o.mo.rho[s][o.cidx] = m.rho[s].v;
        });
    }

    // ---- collisions, species by species ---------------------------------------------------------
    static_for<3>([&](auto S) {
        constexpr int s = decltype(S)::value;
        constexpr int p0 = PAIR_SLOT[s][0], p1 = PAIR_SLOT[s][1];
        const D vx[3] = { m.ux[s], m.upx[p0], m.upx[p1] };
        const D vy[3] = { m.uy[s], m.upy[p0], m.upy[p1] };
        const D u2 = vx[0] * vx[0] + vy[0] * vy[0];                           // collisions.cpp:98-100
        D K[3];                                                               // u2*0.5*invcs2, plasma.cpp:199   (E3)
        K[0] = u2 * D(c.hinvcs2);
        K[1] = (vx[1] * vx[1] + vy[1] * vy[1]) * D(c.hinvcs2);
        K[2] = (vx[2] * vx[2] + vy[2] * vy[2]) * D(c.hinvcs2);
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
            k1_axis_ct<s, decltype(AX)::value>(dv, gt, stash, m, vx, vy, K, AB2, rhoh, u2, uE, pref3, Ex, Ey, o, c);
            k1_sync_point<2>();
        });
        k1_sync_point<1>();
#else
        // four axes with two opposite directions each in a rolled loop (its body fits the instruction cache), then the rest
        // direction with everything resolved at compile time (c = 0: no dot products, bracket 1 - K)
        #pragma unroll 1
        for (int axis = 0; axis < 4; ++axis) {
            const AxisSel sel = axis_select(axis);
            const D w = (axis < 2) ? D(c.w[1]) : D(c.w[2]);
            const D wr = w * m.rho[s];
            const D wT = w * m.T[s];
            BracketParts bp[3];
            #pragma unroll
            for (int j = 0; j < 3; ++j) bp[j] = bracket_parts(axis_dot(sel, vx[j], vy[j]), c);
            D pref = D(0.0), X = D(0.0), cE = D(0.0);
            if constexpr (s < 2) {
                cE = axis_dot(sel, Ex, Ey);
                pref = (axis < 2) ? pref3[1] : pref3[2];
                X = dv.cdiv(bp[0].cu * cE, c.cs2);                             // (c.u)(c.E)/cs2
            }
            auto direction = [&](auto NEG, const int dir, const D fv, const D gv) {
                constexpr bool neg = decltype(NEG)::value;
                D b[3];
                #pragma unroll
                for (int j = 0; j < 3; ++j) b[j] = bracket_value<neg>(bp[j], K[j]);
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
            const D f0 = D(stash[stash_slot(s * 2 + 0, d0)]), g0 = D(stash[stash_slot(s * 2 + 1, d0)]);
            const D f1 = D(stash[stash_slot(s * 2 + 0, d0 + 2)]), g1 = D(stash[stash_slot(s * 2 + 1, d0 + 2)]);
            direction(std::false_type{}, d0, f0, g0);
            direction(std::true_type{}, d0 + 2, f1, g1);
        }
        k1_axis_ct<s, 4>(dv, gt, stash, m, vx, vy, K, AB2, rhoh, u2, uE, pref3, Ex, Ey, o, c);
#endif
    });
}

// Out-of-line recomputation of one cell with the reference's literal arithmetic (literal_cell.cuh): taken when the
// validity record of the fast path trips (operands outside FastDiv's domain, non-finite values).
template <bool WRITE_MACRO>
static __device__ __noinline__ void k1_cell_literal(const double* stash, double Ex, double Ey, double* dst, long long plane, double* rho_q,
                                                    const MacroOut* mo, long long cidx, const LbmConsts* c)
{
    // scalar arguments only (no pointer to the caller's K1Out): the fast path keeps its addresses in registers
    D f[3][NQ], g[3][NQ];
    for (int s = 0; s < 3; ++s)
        for (int i = 0; i < NQ; ++i) {
            f[s][i] = D(stash[stash_slot(s * 2 + 0, i)]);
            g[s][i] = D(stash[stash_slot(s * 2 + 1, i)]);
        }
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

#ifdef PLBM_K1_MAXNREG
#define PLBM_K1_BOUNDS __maxnreg__(PLBM_K1_MAXNREG)
#else
#define PLBM_K1_BOUNDS __launch_bounds__(K1_THREADS, PLBM_K1_MIN_BLOCKS)
#endif

// E_FROM_PHI: the field is not read from the Ex/Ey arrays but formed from the potential exactly as
// poisson::ComputeElectricField_Periodic forms it (src/poisson.cpp:589-607) -- Exf is phi [NYl][NX], Eyf/Ezf the
// neighbouring slabs' boundary rows (nullptr: wrap inside the array).  Saves the K3 sweep and 8 B per cell here.
template <bool WRITE_MACRO, bool E_FROM_PHI>
__global__ void PLBM_K1_BOUNDS
k1_fused_kernel(const double* __restrict__ src, double* __restrict__ dst,
                const double* __restrict__ Exf, const double* __restrict__ Eyf, const double* __restrict__ Ezf,
                double* __restrict__ rho_q, const __grid_constant__ MacroOut mo,
                const __grid_constant__ LbmConsts c, const __grid_constant__ LbmGeom g)
{
    extern __shared__ double stash_all[];
#if PLBM_K1_SYNC
    // every thread reaches every barrier: threads beyond the end of the row shadow its last cell (same loads into the same
    // column of the stash, same results stored to the same addresses) instead of leaving
    const int col = min((int)threadIdx.x, g.NX - 1 - (int)(blockIdx.x * K1_THREADS));
#else
    const int col = threadIdx.x;
#endif
    double* stash = stash_all + col;                  // column of this thread
    const int x = blockIdx.x * K1_THREADS + col;
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
    // L2 for this tile's segment of plane p in the row `g.prefetch_rows` further on (about one wave of CTAs ahead in launch order),
    // so that the CTA which pulls it finds it in L2.  Every segment of every row is requested exactly once per step.
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
#if PLBM_K1_GATE
    CellGate gt;
#else
    unsigned nan_key = 0u;                             // max over the f inputs of 2*|high word|
#endif
    // (ptxas hoists all 54 loads above the first store: one round trip to memory per cell)
    #pragma unroll
    for (int sk = 0; sk < 2 * NSPEC; ++sk) {
        const double* p = src + (long long)(sk * NQ) * g.plane;
        double v[NQ];
        #pragma unroll
        for (int i = 0; i < NQ; ++i) v[i] = k1_load(p + i * g.plane + off[i]);
        #pragma unroll
        for (int i = 0; i < NQ; ++i) stash[stash_slot(sk, i)] = v[i];
#if !PLBM_K1_GATE
        if ((sk & 1) == 0) {                           // f populations: remember whether any of them is a NaN
            #pragma unroll
            for (int i = 0; i < NQ; ++i) { const unsigned hi = (unsigned)__double2hiint(v[i]); nan_key = max(nan_key, hi + hi); }
        }
#endif
    }
    const long long cidx = (long long)y * g.NX + x;    // scalar fields are flat x + NX*y
    double Ex, Ey;
    if constexpr (E_FROM_PHI) {
        const double* row = Exf + (long long)y * g.NX;
        const double* below = (y > 0) ? row - g.NX : (Eyf ? Eyf : Exf + (long long)(g.NYl - 1) * g.NX);
        const double* above = (y < g.NYl - 1) ? row + g.NX : (Ezf ? Ezf : Exf);
        Ex = __dmul_rn(-0.5, __dsub_rn(__ldg(row + xp), __ldg(row + xm)));
        Ey = __dmul_rn(-0.5, __dsub_rn(__ldg(above + x), __ldg(below + x)));
    } else {
        Ex = __ldg(Exf + cidx); Ey = __ldg(Eyf + cidx);
    }

    K1Out o;
    o.dst = dst + (long long)r0o + x;
    o.plane = g.plane;
    o.rho_q = rho_q + cidx;
    o.mo = mo;
    o.cidx = cidx;

#if PLBM_K1_GATE
    // Outside the gate the cell is redone with the literal arithmetic -- unless every population input is zero or NaN and a
    // raw density is NaN: then all 54 outputs are NaN on either path (each species is coupled to both others through the
    // pair velocities) and the moments and rho_q are NaN or the exact zeros of an empty species.  This keeps a lattice that the
    // reference's own dynamics have driven to NaN ("Divergence of the reference") at full speed.
    GatedDiv dv;
    k1_cell_gated<WRITE_MACRO>(dv, gt, stash, D(Ex), D(Ey), o, c);
    if (!gt.ok()) k1_cell_literal<WRITE_MACRO>(stash, Ex, Ey, o.dst, o.plane, o.rho_q, &mo, o.cidx, &c);
#else
    FastDiv dv;
    k1_cell<FastDiv, WRITE_MACRO>(dv, stash, D(Ex), D(Ey), o, c);
    // Outside FastDiv's domain the cell is redone with the literal arithmetic -- unless an f input is a (quiet) NaN and no tiny
    // numerator was seen: then every population output and rho_q is NaN on either path (each species is coupled to both
    // others through the pair velocities), and the only finite outputs, the moments of the NaN-free species, come from
    // divisions by a density >= 1e-10 whose numerators the record covers.  This keeps a lattice that the reference's own
    // dynamics have driven to NaN ("Divergence of the reference") at full speed.
    const bool nan_input = nan_key > 0xffe00000u;
    if (!dv.ok() && !(nan_input && dv.numerators_ok())) k1_cell_literal<WRITE_MACRO>(stash, Ex, Ey, o.dst, o.plane, o.rho_q, &mo, o.cidx, &c);
#endif
}


// K1 with bounce-back walls: same cell work, the pull at the top follows walls.cuh.  A separate kernel so that the periodic
// one above stays exactly as tuned (register allocation is sensitive to anything that moves in it).
// bounce-back boundaries (streaming::StreamingBounceBack, src/streaming.cpp:66-112,150-196) as the pull at the top, see
// walls.cuh; single slab only.
template <bool WRITE_MACRO>
__global__ void PLBM_K1_BOUNDS
k1_walls_kernel(const double* __restrict__ src, double* __restrict__ dst,
                const double* __restrict__ Exf, const double* __restrict__ Eyf,
                double* __restrict__ rho_q, const __grid_constant__ MacroOut mo,
                const __grid_constant__ LbmConsts c, const __grid_constant__ LbmGeom g, const WallArgs wa)
{
    extern __shared__ double stash_all[];
    double* stash = stash_all + threadIdx.x;          // column of this thread: stash[p * K1_THREADS]
    const int x = blockIdx.x * K1_THREADS + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= g.NX) return;

    const int r0o = (y + 1) * g.pitch;
#if PLBM_K1_GATE
    CellGate gt;
#else
    unsigned nan_key = 0u;                             // max over the f inputs of 2*|high word|
#endif
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
            for (int i = 0; i < NQ; ++i) stash[stash_slot(sk, i)] = v[i];
            if ((sk & 1) == 0) {
#if !PLBM_K1_GATE
                #pragma unroll
                for (int i = 0; i < NQ; ++i) { const unsigned hi = (unsigned)__double2hiint(v[i]); nan_key = max(nan_key, hi + hi); }
#endif
                if (rim) {                             // this step's pre-collision f: the stale values of the next pull
                    #pragma unroll
                    for (int i = 0; i < NQ; ++i) wa.rim_next[((long long)((sk >> 1) * NQ + i)) * wa.nrim + ridx] = v[i];
                }
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

#if PLBM_K1_GATE
    // Outside the gate the cell is redone with the literal arithmetic -- unless every population input is zero or NaN and a
    // raw density is NaN: then all 54 outputs are NaN on either path (each species is coupled to both others through the
    // pair velocities) and the moments and rho_q are NaN or the exact zeros of an empty species.  This keeps a lattice that the
    // reference's own dynamics have driven to NaN ("Divergence of the reference") at full speed.
    GatedDiv dv;
    k1_cell_gated<WRITE_MACRO>(dv, gt, stash, D(Ex), D(Ey), o, c);
    if (!gt.ok()) k1_cell_literal<WRITE_MACRO>(stash, Ex, Ey, o.dst, o.plane, o.rho_q, &mo, o.cidx, &c);
#else
    FastDiv dv;
    k1_cell<FastDiv, WRITE_MACRO>(dv, stash, D(Ex), D(Ey), o, c);
    // Outside FastDiv's domain the cell is redone with the literal arithmetic -- unless an f input is a (quiet) NaN and no tiny
    // numerator was seen: then every population output and rho_q is NaN on either path (each species is coupled to both
    // others through the pair velocities), and the only finite outputs, the moments of the NaN-free species, come from
    // divisions by a density >= 1e-10 whose numerators the record covers.  This keeps a lattice that the reference's own
    // dynamics have driven to NaN ("Divergence of the reference") at full speed.
    const bool nan_input = nan_key > 0xffe00000u;
    if (!dv.ok() && !(nan_input && dv.numerators_ok())) k1_cell_literal<WRITE_MACRO>(stash, Ex, Ey, o.dst, o.plane, o.rho_q, &mo, o.cidx, &c);
#endif
}

static __host__ __device__ constexpr size_t k1_smem_bytes() { return sizeof(double) * NPLANES * K1_THREADS; }

// ---- K1 with the pull done by the TMA engine ----------------------------------------------------------------------------------
// The population planes are described to the hardware as ONE 4-D tensor  [sk = species*2+kind][direction][storage row][x]
// (strides 9*plane, plane, pitch, 1; x extent NX, so the padding of a row is out of bounds).  The pull of direction i is then
// a box {K1_THREADS, 1, 1, 6} at (x0 - cx_i, row(y - cy_i), i, 0): six 8*K1_THREADS-byte row segments, one per distribution,
// landing contiguously in the stash.  One elected thread issues the nine boxes of the CTA's tile and everybody waits on one
// mbarrier: no LDG, no STS, no address arithmetic, no staging registers in the compute threads, and the loads do not pass
// through L1.  The periodic wrap in x is not expressible as a box (out-of-bounds elements arrive as zeros): the two cells at
// x = 0 and x = NX-1 fetch their three wrapped neighbours per distribution themselves after the wait.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_load_box4(void* dst, const void* tmap, int c0, int c1, int c2, int c3, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 :: "r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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

static constexpr size_t k1_tma_smem_bytes() { return k1_smem_bytes() + 16; }     // + the mbarrier

template <bool WRITE_MACRO, bool E_FROM_PHI>
__global__ void PLBM_K1_BOUNDS
k1_tma_kernel(const __grid_constant__ CUtensorMap tmap, const double* __restrict__ src, double* __restrict__ dst,
              const double* __restrict__ Exf, const double* __restrict__ Eyf, const double* __restrict__ Ezf,
              double* __restrict__ rho_q, const __grid_constant__ MacroOut mo,
              const __grid_constant__ LbmConsts c, const __grid_constant__ LbmGeom g)
{
    extern __shared__ __align__(128) double stash_all[];
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(stash_all + NPLANES * K1_THREADS);
    const int x0 = blockIdx.x * K1_THREADS;
#if PLBM_K1_SYNC
    // every thread reaches every barrier: threads beyond the end of the row shadow its last cell (same column of the stash, same
    // results stored to the same addresses) instead of leaving
    const int col = min((int)threadIdx.x, g.NX - 1 - x0);
#else
    const int col = threadIdx.x;
#endif
    double* stash = stash_all + col;                  // column of this thread
    const int x = x0 + col;
    const int y = blockIdx.y;
    int rm = y, rp = y + 2;                            // storage rows of y-1 and y+1 (halo rows: storage row = y + 1)
    if (g.wrap_y) {
        if (y == 0) rm = g.NYl;
        if (y == g.NYl - 1) rp = 1;
    }
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, (unsigned)k1_smem_bytes());
        // direction i pulls from (x - cx_i, y - cy_i): cx = 0,1,0,-1,0,1,-1,-1,1   cy = 0,0,1,0,-1,1,1,-1,-1
        const int bx[NQ] = { x0, x0 - 1, x0, x0 + 1, x0, x0 - 1, x0 + 1, x0 + 1, x0 - 1 };
        const int br[NQ] = { y + 1, y + 1, rm, y + 1, rp, rm, rm, rp, rp };
        #pragma unroll
        for (int i = 0; i < NQ; ++i) tma_load_box4(stash_all + stash_slot(0, i), &tmap, bx[i], br[i], i, 0, bar);
    }
    __syncthreads();                                   // the barrier's initialisation is visible to the waiting threads
#if !PLBM_K1_SYNC
    if (x >= g.NX) return;
#endif

    const long long cidx = (long long)y * g.NX + x;    // scalar fields are flat x + NX*y
    const int xm = (x == 0) ? g.NX - 1 : x - 1;
    const int xp = (x == g.NX - 1) ? 0 : x + 1;
    double Ex, Ey;
    if constexpr (E_FROM_PHI) {
        const double* row = Exf + (long long)y * g.NX;
        const double* below = (y > 0) ? row - g.NX : (Eyf ? Eyf : Exf + (long long)(g.NYl - 1) * g.NX);
        const double* above = (y < g.NYl - 1) ? row + g.NX : (Ezf ? Ezf : Exf);
        Ex = __dmul_rn(-0.5, __dsub_rn(__ldg(row + xp), __ldg(row + xm)));
        Ey = __dmul_rn(-0.5, __dsub_rn(__ldg(above + x), __ldg(below + x)));
    } else {
        Ex = __ldg(Exf + cidx); Ey = __ldg(Eyf + cidx);
    }
    const int r0o = (y + 1) * g.pitch;

    mbar_wait(bar, 0);
    if (x == 0 || x == g.NX - 1) {
        // periodic wrap in x: the box delivered zeros for the neighbour outside [0, NX)
        const int rmo = rm * g.pitch, rpo = rp * g.pitch;
        #pragma unroll 1
        for (int sk = 0; sk < 2 * NSPEC; ++sk) {
            const double* p = src + (long long)(sk * NQ) * g.plane;
            if (x == 0) {
                stash[stash_slot(sk, 1)] = __ldg(p + 1 * g.plane + r0o + xm);
                stash[stash_slot(sk, 5)] = __ldg(p + 5 * g.plane + rmo + xm);
                stash[stash_slot(sk, 8)] = __ldg(p + 8 * g.plane + rpo + xm);
            }
            if (x == g.NX - 1) {
                stash[stash_slot(sk, 3)] = __ldg(p + 3 * g.plane + r0o + xp);
                stash[stash_slot(sk, 6)] = __ldg(p + 6 * g.plane + rmo + xp);
                stash[stash_slot(sk, 7)] = __ldg(p + 7 * g.plane + rpo + xp);
            }
        }
    }
#if PLBM_K1_SYNC
    __syncthreads();                                   // shadowing threads read the edge cell's column
#endif

    K1Out o;
    o.dst = dst + (long long)r0o + x;
    o.plane = g.plane;
    o.rho_q = rho_q + cidx;
    o.mo = mo;
    o.cidx = cidx;

    CellGate gt;
    GatedDiv dv;
    k1_cell_gated<WRITE_MACRO>(dv, gt, stash, D(Ex), D(Ey), o, c);
    if (!gt.ok()) k1_cell_literal<WRITE_MACRO>(stash, Ex, Ey, o.dst, o.plane, o.rho_q, &mo, o.cidx, &c);
}


} // namespace plbm
