// collisions.cpp -- collisions::{Collide, ThermalCollisions, Collisions} with the reference's
// signatures (include/collisions.hpp) over plbm_host_thermal_collisions / plbm_host_collisions.
#include "collisions.hpp"
#include "plbm.h"

#include <stdexcept>
#include <string>

namespace collisions {

namespace {
void need(int rc, const char* what)
{
    if (rc) throw std::runtime_error(std::string(what) + ": " + plbm_last_error());
}
// the relaxation times are compile-time constants of the reference (src/collisions.cpp:6-7) and of
// the device code; only Kb, cs2, charges and masses travel
plbm_config units(double Kb, double cs2, double q_e, double q_i, double m_e, double m_i)
{
    plbm_config c = {};
    c.Kb = Kb; c.cs2 = cs2;
    c.q[0] = q_e; c.q[1] = q_i; c.q[2] = 0.0;
    c.m[0] = m_e; c.m[1] = m_i; c.m[2] = m_i;
    return c;
}
}

void ThermalCollisions(
    std::vector<double>& g_e, std::vector<double>& g_i, std::vector<double>& g_n,
    const std::vector<double>& g_eq_e, const std::vector<double>& g_eq_i, const std::vector<double>& g_eq_n,
    const std::vector<double>& g_eq_e_i, const std::vector<double>& g_eq_e_n, const std::vector<double>& g_eq_i_n,
    const std::vector<double>& g_eq_i_e, const std::vector<double>& g_eq_n_e, const std::vector<double>& g_eq_n_i,
    const std::vector<double>& f_eq_e, const std::vector<double>& f_eq_i, const std::vector<double>& f_eq_n,
    const std::vector<double>& f_eq_e_i, const std::vector<double>& f_eq_e_n, const std::vector<double>& f_eq_i_n,
    const std::vector<double>& f_eq_i_e, const std::vector<double>& f_eq_n_e, const std::vector<double>& f_eq_n_i,
    const std::vector<double>& rho_e, const std::vector<double>& rho_i, const std::vector<double>& rho_n,
    const std::vector<double>& ux_e, const std::vector<double>& uy_e,
    const std::vector<double>& ux_i, const std::vector<double>& uy_i,
    const std::vector<double>& ux_n, const std::vector<double>& uy_n,
    std::vector<double>& temp_e, std::vector<double>& temp_i, std::vector<double>& temp_n,
    const int NX, const int NY, const double Kb)
{
    const plbm_config c = units(Kb, 1.0, 0.0, 0.0, 1.0, 1.0);
    const double* g[3] = { g_e.data(), g_i.data(), g_n.data() };
    const double* geq[9] = { g_eq_e.data(), g_eq_i.data(), g_eq_n.data(), g_eq_e_i.data(), g_eq_e_n.data(), g_eq_i_n.data(),
                             g_eq_i_e.data(), g_eq_n_e.data(), g_eq_n_i.data() };
    const double* feq[9] = { f_eq_e.data(), f_eq_i.data(), f_eq_n.data(), f_eq_e_i.data(), f_eq_e_n.data(), f_eq_i_n.data(),
                             f_eq_i_e.data(), f_eq_n_e.data(), f_eq_n_i.data() };
    const double* rho[3] = { rho_e.data(), rho_i.data(), rho_n.data() };
    const double* ux[3] = { ux_e.data(), ux_i.data(), ux_n.data() };
    const double* uy[3] = { uy_e.data(), uy_i.data(), uy_n.data() };
    double* out[3] = { temp_e.data(), temp_i.data(), temp_n.data() };
    need(plbm_host_thermal_collisions(NX, NY, &c, g, geq, feq, rho, ux, uy, out), "collisions::ThermalCollisions");
    g_e.swap(temp_e); g_i.swap(temp_i); g_n.swap(temp_n);       // reference src/collisions.cpp:119-121
}

void Collisions(
    std::vector<double>& f_e, std::vector<double>& f_i, std::vector<double>& f_n,
    const std::vector<double>& f_eq_e, const std::vector<double>& f_eq_i, const std::vector<double>& f_eq_n,
    const std::vector<double>& f_eq_e_i, const std::vector<double>& f_eq_e_n, const std::vector<double>& f_eq_i_n,
    const std::vector<double>& f_eq_i_e, const std::vector<double>& f_eq_n_e, const std::vector<double>& f_eq_n_i,
    const std::vector<double>& rho_e, const std::vector<double>& rho_i,
    const std::vector<double>& ux_e, const std::vector<double>& uy_e,
    const std::vector<double>& ux_i, const std::vector<double>& uy_i,
    const std::vector<double>& Ex, const std::vector<double>& Ey,
    const double q_e, const double q_i, const double m_e, const double m_i,
    std::vector<double>& temp_e, std::vector<double>& temp_i, std::vector<double>& temp_n,
    const std::array<int, Q>&, const std::array<int, Q>&, const std::array<double, Q>&,
    const int NX, const int NY, const double cs2)
{
    const plbm_config c = units(1.0, cs2, q_e, q_i, m_e, m_i);
    const double* f[3] = { f_e.data(), f_i.data(), f_n.data() };
    const double* feq[9] = { f_eq_e.data(), f_eq_i.data(), f_eq_n.data(), f_eq_e_i.data(), f_eq_e_n.data(), f_eq_i_n.data(),
                             f_eq_i_e.data(), f_eq_n_e.data(), f_eq_n_i.data() };
    const double* rho[3] = { rho_e.data(), rho_i.data(), nullptr };
    const double* ux[3] = { ux_e.data(), ux_i.data(), nullptr };
    const double* uy[3] = { uy_e.data(), uy_i.data(), nullptr };
    double* out[3] = { temp_e.data(), temp_i.data(), temp_n.data() };
    need(plbm_host_collisions(NX, NY, &c, f, feq, rho, ux, uy, Ex.data(), Ey.data(), out), "collisions::Collisions");
    f_e.swap(temp_e); f_i.swap(temp_i); f_n.swap(temp_n);       // reference src/collisions.cpp:178-180
}

void Collide(
    std::vector<double>& g_e, std::vector<double>& g_i, std::vector<double>& g_n,
    const std::vector<double>& g_eq_e, const std::vector<double>& g_eq_i, const std::vector<double>& g_eq_n,
    const std::vector<double>& g_eq_e_i, const std::vector<double>& g_eq_e_n, const std::vector<double>& g_eq_i_n,
    const std::vector<double>& g_eq_i_e, const std::vector<double>& g_eq_n_e, const std::vector<double>& g_eq_n_i,
    std::vector<double>& f_e, std::vector<double>& f_i, std::vector<double>& f_n,
    const std::vector<double>& f_eq_e, const std::vector<double>& f_eq_i, const std::vector<double>& f_eq_n,
    const std::vector<double>& f_eq_e_i, const std::vector<double>& f_eq_e_n, const std::vector<double>& f_eq_i_n,
    const std::vector<double>& f_eq_i_e, const std::vector<double>& f_eq_n_e, const std::vector<double>& f_eq_n_i,
    const std::vector<double>& rho_e, const std::vector<double>& rho_i, const std::vector<double>& rho_n,
    const std::vector<double>& ux_e, const std::vector<double>& uy_e,
    const std::vector<double>& ux_i, const std::vector<double>& uy_i,
    const std::vector<double>& ux_n, const std::vector<double>& uy_n,
    const std::vector<double>& Ex, const std::vector<double>& Ey,
    const double q_e, const double q_i, const double m_e, const double m_i,
    std::vector<double>& temp_e, std::vector<double>& temp_i, std::vector<double>& temp_n,
    const std::array<int, Q>& cx, const std::array<int, Q>& cy, const std::array<double, Q>& w,
    const int NX, const int NY, const double Kb, const double cs2)
{
    // thermal first: it reads f_eq, rho and u of the pre-collision state and never f (reference src/collisions.cpp:35-55)
    ThermalCollisions(g_e, g_i, g_n, g_eq_e, g_eq_i, g_eq_n, g_eq_e_i, g_eq_e_n, g_eq_i_n, g_eq_i_e, g_eq_n_e, g_eq_n_i,
                      f_eq_e, f_eq_i, f_eq_n, f_eq_e_i, f_eq_e_n, f_eq_i_n, f_eq_i_e, f_eq_n_e, f_eq_n_i,
                      rho_e, rho_i, rho_n, ux_e, uy_e, ux_i, uy_i, ux_n, uy_n, temp_e, temp_i, temp_n, NX, NY, Kb);
    Collisions(f_e, f_i, f_n, f_eq_e, f_eq_i, f_eq_n, f_eq_e_i, f_eq_e_n, f_eq_i_n, f_eq_i_e, f_eq_n_e, f_eq_n_i,
               rho_e, rho_i, ux_e, uy_e, ux_i, uy_i, Ex, Ey, q_e, q_i, m_e, m_i, temp_e, temp_i, temp_n, cx, cy, w, NX, NY, cs2);
}

} // namespace collisions
