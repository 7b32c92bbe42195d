"""The spectral Poisson solve on its own (poisson::SolvePoisson, FFT + periodic field) against the CPU checker,
bit for bit, over the sequence lengths the pass scheduler treats differently: pure powers of four, 4..4,2 tails,
single 4 / 2 passes, odd primes after the 4/2 passes (radix 3, radix 5, both mixed, larger primes), odd-only lengths,
and the benchmark's sizes."""
import numpy as np
import pytest

from helpers import assert_same_bits

pytestmark = pytest.mark.gpu

SIZES = [3, 4, 5, 8, 15, 16, 17, 25, 27, 30, 32, 45, 49, 60, 64, 75, 96, 100, 121, 128, 200, 225, 243, 256, 360, 384, 500, 512, 720,
         1000, 1024, 1536]


def solve_both(oracle, plbm, NX, NY, seed):
    rng = np.random.default_rng(seed)
    rho = rng.standard_normal((NY, NX)) * 10.0 ** rng.integers(-3, 4)
    o = oracle.PortOracle(NX, NY, poisson="fft")
    o.scalar(oracle.PO_RHO_Q)[...] = rho
    o.solve_poisson()
    want = {"phi": o.scalar(oracle.PO_PHI).copy(), "Ex": o.scalar(oracle.PO_EX).copy(), "Ey": o.scalar(oracle.PO_EY).copy()}
    o.close()
    with plbm.PlasmaLBM(NX, NY, poisson="fft", fields_only=True) as sim:
        Ex, Ey = sim.solve_poisson(rho)
        phi = sim.fields(["phi"])["phi"]
    assert_same_bits(phi, want["phi"], f"{NX}x{NY}: phi")
    assert_same_bits(Ex, want["Ex"], f"{NX}x{NY}: Ex")
    assert_same_bits(Ey, want["Ey"], f"{NX}x{NY}: Ey")


@pytest.mark.parametrize("n", SIZES)
def test_square_lattices(oracle, plbm, n):
    solve_both(oracle, plbm, n, n, seed=n)


@pytest.mark.parametrize("NX,NY", [(64, 48), (48, 64), (200, 120), (81, 256), (2048, 8), (6, 1024)])
def test_rectangular_lattices(oracle, plbm, NX, NY):
    # NX != NY keeps the reference's reshape quirk (src/poisson.cpp:621): NX rows of NY values
    solve_both(oracle, plbm, NX, NY, seed=NX * 7 + NY)


@pytest.mark.parametrize("n", [2048, 3072, 4096])
def test_benchmark_sizes(oracle, plbm, n):
    solve_both(oracle, plbm, n, n, seed=n)


@pytest.mark.parametrize("NX,NY", [(8192, 16), (16, 8192), (6144, 12), (12, 6144), (12288, 4), (4, 12288), (10007, 3),
                                   (5760, 8), (8, 5760), (2880, 6), (7680, 4), (4, 7500)])
def test_longest_sequences(oracle, plbm, NX, NY):
    solve_both(oracle, plbm, NX, NY, seed=NX + NY)
