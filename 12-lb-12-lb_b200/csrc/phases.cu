// phases.cu -- the reference's individual sweeps as separate device kernels on the HOST-API layout
// (AoS i + 9*(x + NX*y), reference include/utils.hpp:6-11).
//
// These back the free functions of include/collisions.hpp, streaming.hpp and poisson.hpp, which
// take caller-owned arrays (equilibria, moments, fields) and therefore cannot use the fused kernel.
// They follow the reference expressions literally with IEEE operations (a second, independent
// formulation of the arithmetic that k1_fused.cu reorganises); speed is not the point here.
//   UpdateMacro          src/plasma.cpp:317-456        ComputeEquilibrium  src/plasma.cpp:162-308
//   ThermalCollisions    src/collisions.cpp:64-122     Collisions          src/collisions.cpp:128-181
//   Streaming*Periodic   src/streaming.cpp:35-59,117-141
#include "phases.h"
#include "exact_math.cuh"

namespace plbm {

__constant__ int p_cx[NQ] = { 0, 1, 0, -1, 0, 1, -1, -1, 1 };
__constant__ int p_cy[NQ] = { 0, 0, 1, 0, -1, 1, 1, -1, -1 };
__constant__ double p_tau_self[3] = { 5.0, 3.0, 1.0 };
__constant__ double p_tau_pair[3] = { 6.0, 4.0, 2.0 };
__constant__ int p_pair_of[3][2] = { { 0, 1 }, { 0, 2 }, { 1, 2 } };

__device__ __forceinline__ D div(D a, D b) { return D(__ddiv_rn(a.v, b.v)); }
__device__ __forceinline__ D dint(int i) { return D((double)i); }

__global__ void update_macro_kernel(PhaseArrays a, PhaseUnits u, int N)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    D rl[3], mx[3], my[3], tl[3];
    for (int k = 0; k < 3; ++k) {
        D r(0.0), ax(0.0), ay(0.0), t(0.0);
        for (int i = 0; i < NQ; ++i) {                                        // plasma.cpp:352-372
            const D fi(a.f[k][(size_t)c * NQ + i]);
            r = r + fi;
            ax = ax + fi * dint(p_cx[i]);
            ay = ay + fi * dint(p_cy[i]);
            t = t + D(a.g[k][(size_t)c * NQ + i]);
        }
        rl[k] = r; mx[k] = ax; my[k] = ay; tl[k] = t;
    }
    const D Ex(a.Ex[c]), Ey(a.Ey[c]);
    for (int k = 0; k < 3; ++k) {
        D rho(0.0), vx(0.0), vy(0.0), T(0.0);
        if (!(rl[k] < D(1e-10))) {
            rho = rl[k]; T = tl[k];
            if (k < 2) {                                                      // plasma.cpp:380-391
                vx = (mx[k] == rl[k] || mx[k] == -rl[k]) ? D(0.0) : div(mx[k], rl[k]);
                vy = (my[k] == rl[k] || my[k] == -rl[k]) ? D(0.0) : div(my[k], rl[k]);
                vx = vx + div((D(0.5) * D(u.q[k])) * Ex, D(u.m[k]));
                vy = vy + div((D(0.5) * D(u.q[k])) * Ey, D(u.m[k]));
            } else {
                vx = div(mx[k], rl[k]);
                vy = div(my[k], rl[k]);
            }
        }
        a.rho[k][c] = rho.v; a.ux[k][c] = vx.v; a.uy[k][c] = vy.v; a.T[k][c] = T.v;
    }
    for (int p = 0; p < 3; ++p) {                                             // plasma.cpp:426-449
        const int sa = (p == 2) ? 1 : 0, sb = (p == 0) ? 1 : 2;
        D px(0.0), py(0.0);
        if (!(rl[sa] < D(1e-10) && rl[sb] < D(1e-10))) {
            px = div(rl[sa] * D(a.ux[sa][c]) + rl[sb] * D(a.ux[sb][c]), rl[sa] + rl[sb]);
            py = div(rl[sa] * D(a.uy[sa][c]) + rl[sb] * D(a.uy[sb][c]), rl[sa] + rl[sb]);
        }
        a.upx[p][c] = px.v; a.upy[p][c] = py.v;
    }
    D rq = div(D(u.q[1]) * D(a.rho[1][c]), D(u.m[1])) + div(D(u.q[0]) * D(a.rho[0][c]), D(u.m[0]));   // plasma.cpp:452
    if (rq < D(1e-15)) rq = D(0.0);
    a.rho_q[c] = rq.v;
}

__device__ __forceinline__ D eq_bracket(D cu, D u2, D invcs2)                  // plasma.cpp:196-200
{
    return ((D(1.0) + cu * invcs2) + (((cu * cu) * D(0.5)) * invcs2) * invcs2) - (u2 * D(0.5)) * invcs2;
}

__global__ void equilibrium_kernel(PhaseArrays a, PhaseUnits u, int N)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    const D invcs2 = div(D(1.0), D(u.cs2));                                   // plasma.cpp:164
    D u2s[3], u2p[3];
    for (int k = 0; k < 3; ++k) {
        u2s[k] = D(a.ux[k][c]) * D(a.ux[k][c]) + D(a.uy[k][c]) * D(a.uy[k][c]);
        u2p[k] = D(a.upx[k][c]) * D(a.upx[k][c]) + D(a.upy[k][c]) * D(a.upy[k][c]);
    }
    const D w[3] = { div(D(4.0), D(9.0)), div(D(1.0), D(9.0)), div(D(1.0), D(36.0)) };
    for (int i = 0; i < NQ; ++i) {
        const D wi = w[i == 0 ? 0 : (i < 5 ? 1 : 2)];
        D bs[3], bp[3];
        for (int k = 0; k < 3; ++k) {
            const D cus = dint(p_cx[i]) * D(a.ux[k][c]) + dint(p_cy[i]) * D(a.uy[k][c]);
            const D cup = dint(p_cx[i]) * D(a.upx[k][c]) + dint(p_cy[i]) * D(a.upy[k][c]);
            bs[k] = eq_bracket(cus, u2s[k], invcs2);
            bp[k] = eq_bracket(cup, u2p[k], invcs2);
        }
        const size_t q = (size_t)c * NQ + i;
        for (int k = 0; k < 3; ++k) {
            a.feq[k][0][q] = ((wi * D(a.rho[k][c])) * bs[k]).v;
            a.geq[k][0][q] = ((wi * D(a.T[k][c])) * bs[k]).v;
            for (int m = 0; m < 2; ++m) {
                a.feq[k][1 + m][q] = ((wi * D(a.rho[k][c])) * bp[p_pair_of[k][m]]).v;
                a.geq[k][1 + m][q] = ((wi * D(a.T[k][c])) * bp[p_pair_of[k][m]]).v;
            }
        }
    }
}

__device__ __forceinline__ D thermal_term(D rho, D tau, D feq)                 // collisions.cpp:86-96
{
    const D a = D(1.0) - div(D(1.0), tau);
    const D C = div(D(9.0) * feq, tau);
    return div((((D(2.0) * rho) * a) * a - (D(2.0) * a) * rho) - C, D(2.0) * (D(2.0) * a + C));
}

__global__ void thermal_collisions_kernel(PhaseArrays a, PhaseUnits u, int N)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * NQ) return;
    const int c = t / NQ;
    const size_t q = (size_t)t;
    for (int k = 0; k < 3; ++k) {
        const D tau[3] = { D(p_tau_self[k]), D(p_tau_pair[p_pair_of[k][0]]), D(p_tau_pair[p_pair_of[k][1]]) };
        const D rho(a.rho[k][c]), ux(a.ux[k][c]), uy(a.uy[k][c]);
        const D t0 = thermal_term(rho, tau[0], D(a.feq[k][0][q]));
        const D t1 = thermal_term(rho, tau[1], D(a.feq[k][1][q]));
        const D t2 = thermal_term(rho, tau[2], D(a.feq[k][2][q]));
        const D dE = (rho * ((t0 + t1) + t2)) * (ux * ux + uy * uy);          // collisions.cpp:98-100
        const D dT = div(-dE, D(u.Kb));                                       // collisions.cpp:102-104
        const D gk(a.g[k][q]);
        const D CT = (div(-(gk - D(a.geq[k][0][q])), tau[0]) - div(gk - D(a.geq[k][1][q]), tau[1]))
                     - div(gk - D(a.geq[k][2][q]), tau[2]);                   // collisions.cpp:107-109
        a.tmp[k][q] = ((gk + CT) + dT).v;                                     // collisions.cpp:112-114
    }
}

__global__ void collisions_kernel(PhaseArrays a, PhaseUnits u, int N)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * NQ) return;
    const int c = t / NQ, i = t % NQ;
    const size_t q = (size_t)t;
    const D Ex(a.Ex[c]), Ey(a.Ey[c]);
    const D w[3] = { div(D(4.0), D(9.0)), div(D(1.0), D(9.0)), div(D(1.0), D(36.0)) };
    const D wi = w[i == 0 ? 0 : (i < 5 ? 1 : 2)];
    const D cs2(u.cs2);
    for (int k = 0; k < 3; ++k) {
        const D tau[3] = { D(p_tau_self[k]), D(p_tau_pair[p_pair_of[k][0]]), D(p_tau_pair[p_pair_of[k][1]]) };
        const D fk(a.f[k][q]);
        const D C = (div(-(fk - D(a.feq[k][0][q])), tau[0]) - div(fk - D(a.feq[k][1][q]), tau[1]))
                    - div(fk - D(a.feq[k][2][q]), tau[2]);                    // collisions.cpp:166-168
        D out = fk + C;
        if (k < 2) {                                                          // collisions.cpp:154-163
            const D ux(a.ux[k][c]), uy(a.uy[k][c]);
            const D cE = dint(p_cx[i]) * Ex + dint(p_cy[i]) * Ey;
            const D cu = dint(p_cx[i]) * ux + dint(p_cy[i]) * uy;
            const D pref = div(div((wi * D(u.q[k])) * D(a.rho[k][c]), D(u.m[k])), cs2) * (D(1.0) - div(D(1.0), D(2.0) * tau[0]));
            const D F = pref * ((cE + div(cu * cE, cs2)) - (ux * Ex + uy * Ey));
            out = out + F;
        }
        a.tmp[k][q] = out.v;
    }
}

// push with periodic wrap, streaming.cpp:44-52
__global__ void stream_periodic_kernel(const double* __restrict__ s0, const double* __restrict__ s1, const double* __restrict__ s2,
                                       double* __restrict__ d0, double* __restrict__ d1, double* __restrict__ d2, int NX, int NY)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= NX * NY * NQ) return;
    const int i = t % NQ, c = t / NQ, x = c % NX, y = c / NX;
    const int xs = (x + NX + p_cx[i]) % NX, ys = (y + NY + p_cy[i]) % NY;
    const size_t to = (size_t)i + NQ * ((size_t)xs + (size_t)NX * ys);
    d0[to] = s0[t]; d1[to] = s1[t]; d2[to] = s2[t];
}

// streaming.cpp:66-112 / :150-196 -- the reference's bounce-back push, restated per DESTINATION slot.
// The reference's loop runs serially in (x outer, y, i) order and several sources can target one slot,
// the last one winning; slots nobody targets keep what `dst` held before (stale temp_* contents).
// A slot (X, Y, j) can be written by
//   (1) plain streaming from (X - cx_j, Y - cy_j), direction j, when that cell exists;
//   (2) a "y out only" reflection of direction o = opp(j) at cell (X + cx_j, Y): needs Y - cy_j outside;
//   (3) an "x out only" reflection of direction o at cell (X, Y + cy_j): needs X - cx_j outside;
//   (4) a corner reflection of direction o at (X, Y) itself: needs both outside.
__constant__ int p_opp[NQ] = { 0, 3, 4, 1, 2, 7, 8, 5, 6 };

__global__ void stream_bounceback_kernel(const double* __restrict__ s0, const double* __restrict__ s1, const double* __restrict__ s2,
                                         double* __restrict__ d0, double* __restrict__ d1, double* __restrict__ d2, int NX, int NY)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= NX * NY * NQ) return;
    const int j = t % NQ, c = t / NQ, X = c % NX, Y = c / NX;
    const int o = p_opp[j];
    const bool xin = (X - p_cx[j] >= 0 && X - p_cx[j] < NX), yin = (Y - p_cy[j] >= 0 && Y - p_cy[j] < NY);
    int bx = -1, by = -1, bi = -1;                       // winning source: the last in (x, y, i) order
    auto offer = [&](int sx, int sy, int si) {
        if (sx > bx || (sx == bx && (sy > by || (sy == by && si > bi)))) { bx = sx; by = sy; bi = si; }
    };
    if (xin && yin) offer(X - p_cx[j], Y - p_cy[j], j);                                  // (1)
    if (!yin && X + p_cx[j] >= 0 && X + p_cx[j] < NX) offer(X + p_cx[j], Y, o);          // (2) source moves to x_str = X (inside), y_str outside
    if (!xin && Y + p_cy[j] >= 0 && Y + p_cy[j] < NY) offer(X, Y + p_cy[j], o);          // (3)
    if (!xin && !yin) offer(X, Y, o);                                                    // (4)
    if (bi < 0) return;                                  // nobody writes this slot: it keeps its stale value
    const size_t from = (size_t)bi + NQ * ((size_t)bx + (size_t)NX * by);
    d0[t] = s0[from]; d1[t] = s1[from]; d2[t] = s2[from];
}

// LBmethod::Initialize in the host layout, plasma.cpp:131-158
__global__ void init_aos_kernel(double* f0, double* f1, double* f2, double* g0, double* g1, double* g2, int NX, int NY,
                                double r0, double r1, double r2, double T0, double T1, double T2)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= NX * NY * NQ) return;
    const int i = t % NQ, c = t / NQ, x = c % NX, y = c / NX;
    const double w = (i == 0) ? __ddiv_rn(4.0, 9.0) : (i < 5 ? __ddiv_rn(1.0, 9.0) : __ddiv_rn(1.0, 36.0));
    const bool inside = (x >= NX / 4 + 1) && (x < 3 * NX / 4) && (y >= NY / 4 + 1) && (y < 3 * NY / 4);
    f0[t] = inside ? __dmul_rn(w, r0) : 0.0; g0[t] = inside ? __dmul_rn(w, T0) : 0.0;
    f1[t] = inside ? __dmul_rn(w, r1) : 0.0; g1[t] = inside ? __dmul_rn(w, T1) : 0.0;
    f2[t] = __dmul_rn(w, r2);                g2[t] = __dmul_rn(w, T2);
}

static inline int blocks_for(long long n, int th) { return (int)((n + th - 1) / th); }

cudaError_t launch_update_macro(const PhaseArrays& a, const PhaseUnits& u, int N, cudaStream_t s)
{
    update_macro_kernel<<<blocks_for(N, 128), 128, 0, s>>>(a, u, N);
    return cudaGetLastError();
}
cudaError_t launch_equilibrium(const PhaseArrays& a, const PhaseUnits& u, int N, cudaStream_t s)
{
    equilibrium_kernel<<<blocks_for(N, 128), 128, 0, s>>>(a, u, N);
    return cudaGetLastError();
}
cudaError_t launch_thermal_collisions(const PhaseArrays& a, const PhaseUnits& u, int N, cudaStream_t s)
{
    thermal_collisions_kernel<<<blocks_for((long long)N * NQ, 128), 128, 0, s>>>(a, u, N);
    return cudaGetLastError();
}
cudaError_t launch_collisions(const PhaseArrays& a, const PhaseUnits& u, int N, cudaStream_t s)
{
    collisions_kernel<<<blocks_for((long long)N * NQ, 128), 128, 0, s>>>(a, u, N);
    return cudaGetLastError();
}
cudaError_t launch_stream_periodic(const double* const src[3], double* const dst[3], int NX, int NY, cudaStream_t s)
{
    stream_periodic_kernel<<<blocks_for((long long)NX * NY * NQ, 256), 256, 0, s>>>(src[0], src[1], src[2], dst[0], dst[1], dst[2], NX, NY);
    return cudaGetLastError();
}

cudaError_t launch_stream_bounceback(const double* const src[3], double* const dst[3], int NX, int NY, cudaStream_t s)
{
    stream_bounceback_kernel<<<blocks_for((long long)NX * NY * NQ, 256), 256, 0, s>>>(src[0], src[1], src[2], dst[0], dst[1], dst[2], NX, NY);
    return cudaGetLastError();
}
cudaError_t launch_init_aos(double* const f[3], double* const g[3], int NX, int NY, const double rho_init[3], const double T_init[3], cudaStream_t s)
{
    init_aos_kernel<<<blocks_for((long long)NX * NY * NQ, 256), 256, 0, s>>>(f[0], f[1], f[2], g[0], g[1], g[2], NX, NY,
                                                                             rho_init[0], rho_init[1], rho_init[2], T_init[0], T_init[1], T_init[2]);
    return cudaGetLastError();
}

} // namespace plbm
