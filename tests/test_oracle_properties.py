"""Size-independent properties of the path, checked on the CPU restatement."""
import numpy as np
import pytest

from helpers import assert_same_bits


def test_periodic_streaming_is_a_permutation(oracle):
    """NX applications of the periodic streaming step return every population to its cell
    (reference src/streaming.cpp:44-52)."""
    N = 12
    o = oracle.PortOracle(N, N, poisson="none", initialize=False)
    rng = np.random.default_rng(0)
    f0 = [rng.uniform(0, 1, size=(N, N, 9)) for _ in range(3)]
    g0 = [rng.uniform(0, 1, size=(N, N, 9)) for _ in range(3)]
    for s in range(3):
        o.f(s)[...] = f0[s]; o.g(s)[...] = g0[s]
    for _ in range(N):
        o.stream()
    for s in range(3):
        assert_same_bits(o.f(s), f0[s], "f after N streams")
        assert_same_bits(o.g(s), g0[s], "g after N streams")


def test_streaming_moves_each_direction_by_its_velocity(oracle):
    N = 8
    cx = [0, 1, 0, -1, 0, 1, -1, -1, 1]; cy = [0, 0, 1, 0, -1, 1, 1, -1, -1]
    o = oracle.PortOracle(N, N, poisson="none", initialize=False)
    o.f(0)[...] = 0.0
    o.f(0)[3, 2, :] = np.arange(1, 10)
    o.stream()
    for i in range(9):
        assert o.f(0)[(3 + cy[i]) % N, (2 + cx[i]) % N, i] == i + 1


def test_mass_collision_conserves_species_mass(oracle):
    """Every D2Q9 equilibrium with cs2 = 1/3 sums to its density, so Collisions conserves sum_i f_i
    per species wherever rho >= 1e-10, up to rounding (no net Guo source: sum_i F_i = 0)."""
    o = oracle.PortOracle(24, 24, poisson="none")
    o.step(3)
    o.update_macro(); o.compute_equilibrium()
    before = [o.f(s).sum(axis=2).copy() for s in range(3)]
    o.thermal_collisions(); o.collisions()
    for s in range(3):
        after = o.f(s).sum(axis=2)
        scale = np.abs(before[s]).max()
        assert np.abs(after - before[s]).max() <= 1e-12 * scale


def test_fft_poisson_satisfies_the_five_point_identity(oracle):
    """phi(E)+phi(W)+phi(N)+phi(S)-4 phi = -(rho_q - mean rho_q): the symbol of reference src/poisson.cpp:393-401."""
    N = 40
    o = oracle.PortOracle(N, N, poisson="fft")
    o.step(5)
    phi = o.scalar(oracle.PO_PHI); rq = o.scalar(oracle.PO_RHO_Q)
    lap = np.roll(phi, 1, 0) + np.roll(phi, -1, 0) + np.roll(phi, 1, 1) + np.roll(phi, -1, 1) - 4 * phi
    assert np.abs(lap + (rq - rq.mean())).max() <= 1e-12 * max(np.abs(rq).max(), 1e-30)
    Ex = o.scalar(oracle.PO_EX)
    assert_same_bits(Ex, -0.5 * (np.roll(phi, -1, 1) - np.roll(phi, 1, 1)), "Ex = -d(phi)/dx")


def test_poisson_none_zeroes_the_field_after_the_first_step(oracle):
    """reference src/poisson.cpp:34-43: E_ext acts in step 0 only."""
    o = oracle.PortOracle(16, 16, poisson="none")
    assert o.scalar(oracle.PO_EX)[0, 0] == o.units.Ex_ext
    o.step(1)
    assert not o.scalar(oracle.PO_EX).any() and not o.scalar(oracle.PO_EY).any()


def test_charge_density_clamps_negative_values(oracle):
    """reference src/plasma.cpp:452-453: everything below 1e-15, negatives included, becomes 0."""
    o = oracle.PortOracle(20, 20, poisson="fft")
    o.step(8)
    rq = o.scalar(oracle.PO_RHO_Q)
    assert (rq >= 0).all() and ((rq == 0) | (rq >= 1e-15)).all()
