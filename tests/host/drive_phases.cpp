// tests/host/drive_phases.cpp -- the reference's loop body (src/plasma.cpp:476-513) executed phase
// by phase through the FREE FUNCTIONS of include/collisions.hpp, streaming.hpp and poisson.hpp on
// host vectors in the reference layout (UpdateMacro / ComputeEquilibrium, private in the reference,
// through their C-ABI counterparts).  Dumps the 15 fields of every step listed in $PLBM_DUMP_STEPS.
#include "collisions.hpp"
#include "plbm.h"
#include "poisson.hpp"
#include "streaming.hpp"
#include "visualize.hpp"

#include <cstdio>
#include <cstdlib>
#include <exception>
#include <stdexcept>
#include <string>
#include <vector>

using V = std::vector<double>;

int main(int argc, char** argv)
{
    if (argc < 6) { std::fprintf(stderr, "usage: %s NX NY NSTEPS poisson bc\n", argv[0]); return 2; }
    const int NX = std::atoi(argv[1]), NY = std::atoi(argv[2]), NSTEPS = std::atoi(argv[3]);
    const auto ptype = static_cast<poisson::PoissonType>(std::atoi(argv[4]));
    const auto btype = static_cast<streaming::BCType>(std::atoi(argv[5]));
    const size_t N = static_cast<size_t>(NX) * NY, NQ9 = N * Q;
    try {
        plbm_config u = {};
        if (plbm_units_from_si(1, 1, 1e-2, 0.0, 1e4, 300, 300, 1e11, 1e18, &u)) throw std::runtime_error(plbm_last_error());
        const std::array<int, Q> cx = { { 0, 1, 0, -1, 0, 1, -1, -1, 1 } }, cy = { { 0, 0, 1, 0, -1, 1, 1, -1, -1 } };
        const std::array<double, Q> w = { { 4.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 9.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0 } };
        V f[3], g[3], tmp[3], feq[9], geq[9], rho[3], ux[3], uy[3], T[3], upx[3], upy[3];
        for (int k = 0; k < 3; ++k) {
            f[k].assign(NQ9, 0.0); g[k].assign(NQ9, 0.0); tmp[k].assign(NQ9, 0.0);
            rho[k].assign(N, 0.0); ux[k].assign(N, 0.0); uy[k].assign(N, 0.0); T[k].assign(N, 0.0); upx[k].assign(N, 0.0); upy[k].assign(N, 0.0);
        }
        for (int k = 0; k < 9; ++k) { feq[k].assign(NQ9, 0.0); geq[k].assign(NQ9, 0.0); }
        V Ex(N, u.Ex_ext), Ey(N, u.Ey_ext), rho_q(N, 0.0);
        // initial condition of reference src/plasma.cpp:131-158
        for (int y = 0; y < NY; ++y) for (int x = 0; x < NX; ++x) for (int i = 0; i < Q; ++i) {
            const size_t k = INDEX(x, y, i, NX, Q);
            const bool in = x >= NX / 4 + 1 && x < 3 * NX / 4 && y >= NY / 4 + 1 && y < 3 * NY / 4;
            if (in) { f[0][k] = w[i] * u.rho_init[0]; g[0][k] = w[i] * u.T_init[0]; f[1][k] = w[i] * u.rho_init[1]; g[1][k] = w[i] * u.T_init[1]; }
            f[2][k] = w[i] * u.rho_init[2]; g[2][k] = w[i] * u.T_init[2];
        }
        auto cptr = [](V* v, int n, const double** out) { for (int k = 0; k < n; ++k) out[k] = v[k].data(); };
        auto mptr = [](V* v, int n, double** out) { for (int k = 0; k < n; ++k) out[k] = v[k].data(); };
        visualize::InitVisualization(NX, NY, NSTEPS);
        for (int t = 0; t < NSTEPS; ++t) {
            const double *fc[3], *gc[3], *rc[3], *uxc[3], *uyc[3], *Tc[3], *pxc[3], *pyc[3];
            double *rm[3], *uxm[3], *uym[3], *Tm[3], *pxm[3], *pym[3], *feqm[9], *geqm[9];
            cptr(f, 3, fc); cptr(g, 3, gc); mptr(rho, 3, rm); mptr(ux, 3, uxm); mptr(uy, 3, uym); mptr(T, 3, Tm); mptr(upx, 3, pxm); mptr(upy, 3, pym);
            if (plbm_host_update_macro(NX, NY, &u, fc, gc, Ex.data(), Ey.data(), rm, uxm, uym, Tm, pxm, pym, rho_q.data())) throw std::runtime_error(plbm_last_error());
            cptr(rho, 3, rc); cptr(ux, 3, uxc); cptr(uy, 3, uyc); cptr(T, 3, Tc); cptr(upx, 3, pxc); cptr(upy, 3, pyc); mptr(feq, 9, feqm); mptr(geq, 9, geqm);
            if (plbm_host_equilibrium(NX, NY, &u, rc, uxc, uyc, Tc, pxc, pyc, feqm, geqm)) throw std::runtime_error(plbm_last_error());
            collisions::Collide(g[0], g[1], g[2], geq[0], geq[1], geq[2], geq[3], geq[4], geq[5], geq[6], geq[7], geq[8],
                                f[0], f[1], f[2], feq[0], feq[1], feq[2], feq[3], feq[4], feq[5], feq[6], feq[7], feq[8],
                                rho[0], rho[1], rho[2], ux[0], uy[0], ux[1], uy[1], ux[2], uy[2], Ex, Ey,
                                u.q[0], u.q[1], u.m[0], u.m[1], tmp[0], tmp[1], tmp[2], cx, cy, w, NX, NY, u.Kb, u.cs2);
            streaming::Stream(f[0], f[1], f[2], tmp[0], tmp[1], tmp[2], g[0], g[1], g[2], cx, cy, NX, NY, btype);
            poisson::SolvePoisson(Ex, Ey, rho_q, NX, NY, 1.8, ptype, btype);
            visualize::UpdateVisualization(t, NX, NY, ux[0], uy[0], ux[1], uy[1], ux[2], uy[2], T[0], T[1], T[2], rho[0], rho[1], rho[2], rho_q, Ex, Ey);
        }
        visualize::CloseVisualization();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
