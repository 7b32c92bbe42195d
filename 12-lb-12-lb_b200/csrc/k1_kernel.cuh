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
#ifndef PLBM_K1_MIN_BLOCKS
#define PLBM_K1_MIN_BLOCKS 6
#endif
constexpr int K1_THREADS = PLBM_K1_THREADS;

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
                f[i] = D(stash[((s * 2 + 0) * NQ + i) * K1_THREADS]);
                g[i] = D(stash[((s * 2 + 1) * NQ + i) * K1_THREADS]);
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
            const D f0 = D(stash[((s * 2 + 0) * NQ + d0) * K1_THREADS]), g0 = D(stash[((s * 2 + 1) * NQ + d0) * K1_THREADS]);
            if (axis != 4) {
                const D f1 = D(stash[((s * 2 + 0) * NQ + d0 + 2) * K1_THREADS]), g1 = D(stash[((s * 2 + 1) * NQ + d0 + 2) * K1_THREADS]);
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
                const D fv = D(stash[((s * 2 + 0) * NQ + dir) * K1_THREADS]);
                const D gv = D(stash[((s * 2 + 1) * NQ + dir) * K1_THREADS]);
                if (sg) direction(std::true_type{}, dir, fv, gv);
                else direction(std::false_type{}, dir, fv, gv);
            }
#endif
        }
    });
}

// Out-of-line recomputation of one cell with the reference's literal arithmetic (literal_cell.cuh): taken when the
// validity record of the fast path trips (operands outside FastDiv's domain, non-finite values).
template <bool WRITE_MACRO>
static __device__ __noinline__ void k1_cell_literal(const double* stash, double Ex, double Ey, const K1Out* o, const LbmConsts* c)
{
    D f[3][NQ], g[3][NQ];
    for (int s = 0; s < 3; ++s)
        for (int i = 0; i < NQ; ++i) {
            f[s][i] = D(stash[((s * 2 + 0) * NQ + i) * K1_THREADS]);
            g[s][i] = D(stash[((s * 2 + 1) * NQ + i) * K1_THREADS]);
        }
    LitMacro m;
    lit_update_macro(f, g, D(Ex), D(Ey), *c, m);
    *o->rho_q = m.rho_q.v;
    if (WRITE_MACRO) {
        for (int s = 0; s < 3; ++s) {
            o->mo.ux[s][o->cidx] = m.ux[s].v; o->mo.uy[s][o->cidx] = m.uy[s].v;
            o->mo.T[s][o->cidx] = m.T[s].v;   o->mo.rho[s][o->cidx] = m.rho[s].v;
        }
    }
    for (int i = 0; i < NQ; ++i) {
        const D fi[3] = { f[0][i], f[1][i], f[2][i] }, gi[3] = { g[0][i], g[1][i], g[2][i] };
        D fo[3], go[3];
        lit_collide_direction(i, fi, gi, m, D(Ex), D(Ey), *c, fo, go);
        for (int s = 0; s < 3; ++s) {
            o->dst[((s * 2 + 0) * NQ + i) * o->plane] = fo[s].v;
            o->dst[((s * 2 + 1) * NQ + i) * o->plane] = go[s].v;
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

    unsigned nan_key = 0u;                             // max over the f inputs of 2*|high word|
    #pragma unroll
    for (int sk = 0; sk < 2 * NSPEC; ++sk) {
        const double* p = src + (long long)(sk * NQ) * g.plane;
        double v[NQ];
        #pragma unroll
        for (int i = 0; i < NQ; ++i) v[i] = __ldg(p + i * g.plane + off[i]);
        #pragma unroll
        for (int i = 0; i < NQ; ++i) stash[(sk * NQ + i) * K1_THREADS] = v[i];
        if ((sk & 1) == 0) {                           // f populations: remember whether any of them is a NaN
            #pragma unroll
            for (int i = 0; i < NQ; ++i) { const unsigned hi = (unsigned)__double2hiint(v[i]); nan_key = max(nan_key, hi + hi); }
        }
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

    FastDiv dv;
    k1_cell<FastDiv, WRITE_MACRO>(dv, stash, D(Ex), D(Ey), o, c);
    // Outside FastDiv's domain the cell is redone with the literal arithmetic -- unless an f input is a (quiet) NaN and no tiny
    // numerator was seen: then every population output and rho_q is NaN on either path (each species is coupled to both
    // others through the pair velocities), and the only finite outputs, the moments of the NaN-free species, come from
    // divisions by a density >= 1e-10 whose numerators the record covers.  This keeps a lattice that the reference's own
    // dynamics have driven to NaN ("Divergence of the reference") at full speed.
    const bool nan_input = nan_key > 0xffe00000u;
    if (!dv.ok() && !(nan_input && dv.numerators_ok())) k1_cell_literal<WRITE_MACRO>(stash, Ex, Ey, &o, &c);
}


// K1 with bounce-back walls: same cell work, the pull at the top follows walls.cuh.  A separate kernel so that the periodic
// one above stays exactly as tuned (register allocation is sensitive to anything that moves in it).
// bounce-back boundaries (streaming::StreamingBounceBack, src/streaming.cpp:66-112,150-196) as the pull at the top, see
// walls.cuh; single slab only.
template <bool WRITE_MACRO>
__global__ void PLBM_K1_BOUNDS
k1_walls_kernel(const double* __restrict__ src, double* __restrict__ dst,
                const double* __restrict__ Exf, const double* __restrict__ Eyf,
                double* __restrict__ rho_q, const MacroOut mo,
                const __grid_constant__ LbmConsts c, const __grid_constant__ LbmGeom g, const WallArgs wa)
{
    extern __shared__ double stash_all[];
    double* stash = stash_all + threadIdx.x;          // column of this thread: stash[p * K1_THREADS]
    const int x = blockIdx.x * K1_THREADS + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= g.NX) return;

    const int r0o = (y + 1) * g.pitch;
    unsigned nan_key = 0u;                             // max over the f inputs of 2*|high word|
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
            for (int i = 0; i < NQ; ++i) stash[(sk * NQ + i) * K1_THREADS] = v[i];
            if ((sk & 1) == 0) {
                #pragma unroll
                for (int i = 0; i < NQ; ++i) { const unsigned hi = (unsigned)__double2hiint(v[i]); nan_key = max(nan_key, hi + hi); }
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

    FastDiv dv;
    k1_cell<FastDiv, WRITE_MACRO>(dv, stash, D(Ex), D(Ey), o, c);
    // Outside FastDiv's domain the cell is redone with the literal arithmetic -- unless an f input is a (quiet) NaN and no tiny
    // numerator was seen: then every population output and rho_q is NaN on either path (each species is coupled to both
    // others through the pair velocities), and the only finite outputs, the moments of the NaN-free species, come from
    // divisions by a density >= 1e-10 whose numerators the record covers.  This keeps a lattice that the reference's own
    // dynamics have driven to NaN ("Divergence of the reference") at full speed.
    const bool nan_input = nan_key > 0xffe00000u;
    if (!dv.ok() && !(nan_input && dv.numerators_ok())) k1_cell_literal<WRITE_MACRO>(stash, Ex, Ey, &o, &c);
}

static constexpr size_t k1_smem_bytes() { return sizeof(double) * NPLANES * K1_THREADS; }


} // namespace plbm
