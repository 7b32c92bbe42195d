// collisions.hpp -- public surface of the reference's collision module (reference
// include/collisions.hpp:10-70), implemented on the B200 through include/plbm.h.
// Same borrow-and-swap convention as the reference: after a call g_*/f_* hold the post-collision
// populations and temp_* the previous ones (reference src/collisions.cpp:119-121, 178-180).
#pragma once

#include "utils.hpp"

#include <array>
#include <vector>

namespace collisions {

// thermal collision first, then mass collision (reference src/collisions.cpp:15-56)
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
    const double q_e, const double q_i,
    const double m_e, const double m_i,
    std::vector<double>& temp_e, std::vector<double>& temp_i, std::vector<double>& temp_n,
    const std::array<int, Q>& cx, const std::array<int, Q>& cy, const std::array<double, Q>& w,
    const int NX, const int NY, const double Kb, const double cs2);

// reference src/collisions.cpp:64-122
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
    const int NX, const int NY, const double Kb);

// reference src/collisions.cpp:128-181
void Collisions(
    std::vector<double>& f_e, std::vector<double>& f_i, std::vector<double>& f_n,
    const std::vector<double>& f_eq_e, const std::vector<double>& f_eq_i, const std::vector<double>& f_eq_n,
    const std::vector<double>& f_eq_e_i, const std::vector<double>& f_eq_e_n, const std::vector<double>& f_eq_i_n,
    const std::vector<double>& f_eq_i_e, const std::vector<double>& f_eq_n_e, const std::vector<double>& f_eq_n_i,
    const std::vector<double>& rho_e, const std::vector<double>& rho_i,
    const std::vector<double>& ux_e, const std::vector<double>& uy_e,
    const std::vector<double>& ux_i, const std::vector<double>& uy_i,
    const std::vector<double>& Ex, const std::vector<double>& Ey,
    const double q_e, const double q_i,
    const double m_e, const double m_i,
    std::vector<double>& temp_e, std::vector<double>& temp_i, std::vector<double>& temp_n,
    const std::array<int, Q>& cx, const std::array<int, Q>& cy, const std::array<double, Q>& w,
    const int NX, const int NY, const double cs2);

} // namespace collisions
