"""Parity of the CUDA path (through the C ABI of include/plbm.h) against the CPU checker.

Bar: BIT-EXACT (up to the sign of exact zeros) for every visualised field, the potential and all
54 populations, at every compared step -- stronger than the <= 1e-10 field-normalised tolerance
BASELINE.json states, and the only way to hold that tolerance over hundreds of steps
(noise-floor table).
"""
import numpy as np
import pytest

from pathlib import Path

from helpers import assert_fields_same, assert_same_bits

GOLDEN = Path(__file__).resolve().parent / "golden"

pytestmark = pytest.mark.gpu


def run_both(oracle, plbm, NX, NY, poisson, nsteps, check_steps, **kw):
    o = oracle.PortOracle(NX, NY, poisson=poisson, **kw)
    with plbm.PlasmaLBM(NX, NY, poisson=poisson, **kw) as sim:
        for t in range(nsteps):
            o.step(1)
            want = t in check_steps
            sim.step(1, want_fields=want)
            if want:
                assert_fields_same(sim.fields(), o.fields(), f"{NX}x{NY}/{poisson}/t={t}")
        f, g = sim.download_state()
    for s in range(3):
        assert_same_bits(f[s], o.f(s), f"{NX}x{NY}/{poisson}: f[{s}] after {nsteps} steps")
        assert_same_bits(g[s], o.g(s), f"{NX}x{NY}/{poisson}: g[{s}] after {nsteps} steps")
    o.close()


@pytest.mark.parametrize("NX,NY", [(64, 64), (50, 70), (129, 33), (16, 16)])
def test_fused_step_without_poisson(oracle, plbm, NX, NY):
    run_both(oracle, plbm, NX, NY, "none", 24, {0, 1, 2, 5, 23})


@pytest.mark.parametrize("NX,NY,nsteps", [(64, 64, 24), (60, 60, 24), (71, 71, 24), (45, 30, 24),
                                          # NX != NY: the reference reshapes the flat array as NX rows of NY
                                          # (src/poisson.cpp:621); 48x64 diverges to Inf after ~18 steps on the CPU too
                                          (48, 64, 12)])
def test_fused_step_with_fft_poisson(oracle, plbm, NX, NY, nsteps):
    run_both(oracle, plbm, NX, NY, "fft", nsteps, {0, 1, 2, 5, nsteps - 1})


def test_reference_default_case_200x200(oracle, plbm):
    """configs[1] of BASELINE.json: 200x200, 200 steps, FFT Poisson, periodic."""
    run_both(oracle, plbm, 200, 200, "fft", 200, {0, 1, 10, 50, 100, 150, 199})


def test_benchmark_workload_2048x2048_whole_step(oracle, plbm):
    """configs[2] of BASELINE.json, the lattice bench.py times: 8 whole time steps (K1 + spectral Poisson + field) at 2048x2048
    against the UNMODIFIED reference compiled in place (parity build) where oracle/_ref holds it, else the C restatement:
    all 15 visualised fields and the potential, bit for bit."""
    NX, steps = 2048, 8
    if oracle.have_reference("parity"):
        _, dumps, _ = oracle.run_reference(NX, NX, steps, poisson="fft", kind="parity", dump_steps=[steps - 1])
        want = dumps[steps - 1]
    else:
        o = oracle.PortOracle(NX, NX, poisson="fft")
        o.step(steps)
        want = o.fields()
        o.close()
    with plbm.PlasmaLBM(NX, NX, poisson="fft") as sim:
        sim.step(steps, want_fields=True)
        assert_fields_same(sim.fields(), want, f"{NX}x{NX}/fft/t={steps - 1}")


@pytest.mark.parametrize("NX,NY", [(64, 64), (50, 70), (129, 33)])
def test_persistent_warp_kernel_with_tma_pool(oracle, plbm, monkeypatch, NX, NY):
    """PLBM_K1_POOL=1: K1 as persistent warps whose pulls are done by the TMA engine into a pool of stash buffers
    (csrc/k1_kernel.cuh: k1_pool_kernel).  Not the default (measured slower), but it must stay bit-identical."""
    monkeypatch.setenv("PLBM_K1_POOL", "1")
    run_both(oracle, plbm, NX, NY, "fft", 12, {0, 1, 5, 11})


@pytest.mark.parametrize("NX,NY", [(64, 64), (50, 70), (130, 33)])
def test_tile_kernel_with_per_thread_loads(oracle, plbm, monkeypatch, NX, NY):
    """PLBM_K1_TMA=0: the per-tile kernel in which every thread pulls its own populations (k1_fused_kernel) instead of the default
    k1_tma_kernel, whose pull is nine TMA boxes."""
    monkeypatch.setenv("PLBM_K1_TMA", "0")
    run_both(oracle, plbm, NX, NY, "fft", 12, {0, 1, 5, 11})


def test_other_physical_parameters(oracle, plbm):
    run_both(oracle, plbm, 40, 40, "fft", 12, {0, 11}, Z_ion=2, A_ion=4, Ex_SI=3e-2, Ey_SI=-1e-2, T_i_SI=500.0)


def test_random_state_one_step(oracle, plbm):
    """Arbitrary (seeded) positive populations and field: one fused step against the checker."""
    NX, NY = 96, 40
    rng = np.random.default_rng(1234)
    f = rng.uniform(0.05, 1.0, size=(3, NY, NX, 9))
    g = rng.uniform(0.01, 0.5, size=(3, NY, NX, 9))
    f[1] *= 1800.0
    f[2] *= 1e9
    f[0][:, : NX // 3] = 0.0          # an empty electron region (threshold branch)
    f[1][: NY // 2] *= 1e-14          # ions below the 1e-10 density threshold
    Ex = rng.normal(0, 1e-3, size=(NY, NX))
    Ey = rng.normal(0, 1e-3, size=(NY, NX))
    o = oracle.PortOracle(NX, NY, poisson="none", initialize=False)
    for s in range(3):
        o.f(s)[...] = f[s]
        o.g(s)[...] = g[s]
    o.scalar(oracle.PO_EX)[...] = Ex
    o.scalar(oracle.PO_EY)[...] = Ey
    with plbm.PlasmaLBM(NX, NY, poisson="none", initialize=False) as sim:
        sim.upload_state(f, g)
        f2, g2 = sim.download_state()
        assert_same_bits(f2, f, "upload/download round trip f")
        assert_same_bits(g2, g, "upload/download round trip g")
        sim.set_efield(Ex, Ey)
        for t in range(3):
            o.step(1)
            sim.step(1, want_fields=True)
            assert_fields_same(sim.fields(), o.fields(), f"random/t={t}")
        f3, g3 = sim.download_state()
    for s in range(3):
        assert_same_bits(f3[s], o.f(s), f"random f[{s}]")
        assert_same_bits(g3[s], o.g(s), f"random g[{s}]")


def test_tiny_and_huge_values_take_the_exact_fallback(oracle, plbm):
    """Populations far outside the domain where the fast division sequence is proven exact
    (subnormals, 1e-300, 1e+200): the kernel must notice and recompute those cells with IEEE
    divisions, so the result is still bit-identical (Inf/NaN included)."""
    NX, NY = 64, 24
    rng = np.random.default_rng(99)
    f = rng.uniform(0.05, 1.0, size=(3, NY, NX, 9))
    g = rng.uniform(0.01, 0.5, size=(3, NY, NX, 9))
    f[2] *= 1e9
    f[0][:, 0:8] *= 1e-300            # tiny but non-zero electrons
    f[0][:, 8:16] *= 5e-324 / 0.05    # deep subnormals
    g[1][:, 16:24] *= 1e-310
    f[1][:, 24:32] *= 1e200           # overflow in the collision terms
    g[0][:, 32:40] = 0.0
    Ex = rng.normal(0, 1e-3, size=(NY, NX))
    Ey = rng.normal(0, 1e-3, size=(NY, NX))
    Ex[:, 40:48] *= 1e-300
    o = oracle.PortOracle(NX, NY, poisson="none", initialize=False)
    for s in range(3):
        o.f(s)[...] = f[s]
        o.g(s)[...] = g[s]
    o.scalar(oracle.PO_EX)[...] = Ex
    o.scalar(oracle.PO_EY)[...] = Ey
    with plbm.PlasmaLBM(NX, NY, poisson="none", initialize=False) as sim:
        sim.upload_state(f, g)
        sim.set_efield(Ex, Ey)
        with np.errstate(all="ignore"):
            for t in range(2):
                o.step(1)
                sim.step(1, want_fields=True)
                assert_fields_same(sim.fields(), o.fields(), f"extreme/t={t}")
            f3, g3 = sim.download_state()
            for s in range(3):
                assert_same_bits(f3[s], o.f(s), f"extreme f[{s}]")
                assert_same_bits(g3[s], o.g(s), f"extreme g[{s}]")


def _one_step_against_checker(oracle, plbm, f, g, Ex, Ey, label, steps=1):
    NY, NX = Ex.shape
    o = oracle.PortOracle(NX, NY, poisson="none", initialize=False)
    for s in range(3):
        o.f(s)[...] = f[s]
        o.g(s)[...] = g[s]
    o.scalar(oracle.PO_EX)[...] = Ex
    o.scalar(oracle.PO_EY)[...] = Ey
    with plbm.PlasmaLBM(NX, NY, poisson="none", initialize=False) as sim:
        sim.upload_state(f, g)
        sim.set_efield(Ex, Ey)
        with np.errstate(all="ignore"):
            for t in range(steps):
                o.step(1)
                sim.step(1, want_fields=True)
                assert_fields_same(sim.fields(), o.fields(), f"{label}/t={t}")
            f3, g3 = sim.download_state()
            for s in range(3):
                assert_same_bits(f3[s], o.f(s), f"{label} f[{s}]")
                assert_same_bits(g3[s], o.g(s), f"{label} g[{s}]")
    o.close()


def test_values_at_the_edges_of_the_gate(oracle, plbm):
    """The per-cell gate of K1 (csrc/exact_math.cuh: CellGate; DESIGN.md "K1 gate") accepts the fast divisions when every noted
    value is inside its bounds.  Columns of cells sit just inside and just outside every bound -- population inputs at 2^-950,
    field at 2^-400, temperatures at 2^-300, velocities at 2^-240, densities at 2^200, a pair density that cancels to zero --
    and either way the result must be the checker's, bit for bit (Inf/NaN included)."""
    NX, NY = 160, 16
    rng = np.random.default_rng(2024)
    f = rng.uniform(0.05, 1.0, size=(3, NY, NX, 9)); g = rng.uniform(0.01, 0.5, size=(3, NY, NX, 9))
    f[2] *= 1e9
    Ex = rng.normal(0, 1e-3, size=(NY, NX)); Ey = rng.normal(0, 1e-3, size=(NY, NX))
    two = lambda e: np.ldexp(1.0, e)
    col = iter(range(0, NX, 4))
    def block():
        c = next(col); return slice(c, c + 4)
    for e in (-949, -950, -951, -960, -970, -1000, -1040, -1074):          # one tiny population among normal ones
        b = block(); f[0][:, b, 3] = two(e) * rng.uniform(1.0, 1.9, size=(NY, 4))
    for e in (-945, -951, -965, -1030):                                   # a whole species tiny (empty-species branch, divisions by tau)
        b = block(); f[0][:, b, :] = two(e) * rng.uniform(1.0, 1.9, size=(NY, 4, 9)); g[0][:, b, :] = two(e) * rng.uniform(1.0, 1.9, size=(NY, 4, 9))
    for e in (-299, -301, -700, -949, -952):                              # temperatures
        b = block(); g[1][:, b, :] = two(e) * rng.uniform(1.0, 1.9, size=(NY, 4, 9))
    for e in (-399, -401, -600, -1000):                                   # field
        b = block(); Ex[:, b] = two(e) * rng.uniform(1.0, 1.9, size=(NY, 4)); Ey[:, b] = -two(e)
    for e in (-239, -241, -300, -500, -559, -561):                        # velocity = tiny momentum over O(1) density
        b = block(); f[0][:, b, :] = 0.0; f[0][:, b, 0] = 1.0; f[0][:, b, 1] = two(e); Ex[:, b] = 0.0; Ey[:, b] = 0.0
    for e in (190, 199, 201, 260, 500):                                   # huge densities
        b = block(); f[2][:, b, :] *= two(e) / 1e9
    b = block(); f[1][:, b, :] = -f[0][:, b, :]                           # rho_e + rho_i == 0 exactly: pair density zero
    b = block(); f[2][:, b, :] = 0.0                                      # neutrals empty: 0/0 in their thermal term (tau_n = 1)
    _one_step_against_checker(oracle, plbm, f, g, Ex, Ey, "gate edges", steps=2)



@pytest.mark.gpu
def test_empty_charged_species_short_path(oracle, plbm):
    """Cells whose electrons AND ions are below the density threshold take K1's second copy of the cell code (k1_species_empty:
    f - ((f/tau_0 + f/tau_1) + f/tau_2) per direction), chosen per warp.  Whole warps of such cells, warps that mix them with
    dense cells, one species empty only, populations that are tiny but not zero (below the threshold with non-zero f and g), zero
    g beside non-zero f and the reverse, and empty neutrals (the reference divides 0 by 0 there): all bit-identical to the checker."""
    NX, NY = 320, 12
    rng = np.random.default_rng(777)
    f = rng.uniform(0.05, 1.0, size=(3, NY, NX, 9)); g = rng.uniform(0.01, 0.5, size=(3, NY, NX, 9))
    f[1] *= 1800.0
    f[2] *= 1e9
    Ex = rng.normal(0, 1e-3, size=(NY, NX)); Ey = rng.normal(0, 1e-3, size=(NY, NX))
    for s in (0, 1):                                   # whole warps and whole tiles of exact zeros
        f[s][:, 0:128] = 0.0; g[s][:, 0:128] = 0.0
    f[0][:, 128:160] = 0.0; g[0][:, 128:160] = 0.0     # electrons empty only
    f[1][:, 160:192] = 0.0; g[1][:, 160:192] = 0.0     # ions empty only
    for s in (0, 1):                                   # below the threshold but not zero, in a warp of their own ...
        f[s][:, 192:224] *= 1e-14 / (1800.0 if s else 1.0); g[s][:, 192:224] *= 1e-14
    for s in (0, 1):                                   # ... and patches narrower than a warp among dense cells
        f[s][:, 230:237] = 0.0; g[s][:, 230:237] = 0.0
        f[s][3:5, 250:253] *= 1e-20
    f[0][:, 260:270] = 0.0; f[1][:, 260:270] = 0.0     # zero f, non-zero g
    g[0][:, 270:280] = 0.0; g[1][:, 270:280] = 0.0     # zero g, non-zero f (dense: the full path)
    f[2][:, 64:70] = 0.0; g[2][:, 64:70] = 0.0         # empty neutrals inside the all-empty region
    f[2][:, 290:296] = 0.0                             # and among dense charged species
    Ex[:, 100:110] = 0.0; Ey[:, 100:110] = 0.0
    _one_step_against_checker(oracle, plbm, f, g, Ex, Ey, "empty species", steps=3)

@pytest.mark.parametrize("seed", [11, 12, 13])
def test_random_exponents_everywhere(oracle, plbm, seed):
    """Fuzz: every one of the 54 populations and both field components of every cell gets an independent magnitude from 2^-1074
    to 2^+300 (mostly moderate, a few extreme, some exact zeros, some negative): whichever side of the gate a cell lands on, one
    time step equals the checker's bit for bit."""
    NX, NY = 128, 64
    rng = np.random.default_rng(seed)
    def draw(shape, centre):
        kind = rng.random(shape)
        e = np.where(kind < 0.80, rng.integers(centre - 12, centre + 12, shape),
            np.where(kind < 0.90, rng.integers(-400, 120, shape), rng.integers(-1074, 300, shape)))
        v = np.ldexp(rng.uniform(1.0, 2.0, shape), e)
        v = np.where(rng.random(shape) < 0.03, 0.0, v)
        return np.where(rng.random(shape) < 0.02, -v, v)
    f = np.stack([draw((NY, NX, 9), -3), draw((NY, NX, 9), 8), draw((NY, NX, 9), 30)])
    g = np.stack([draw((NY, NX, 9), -3), draw((NY, NX, 9), -6), draw((NY, NX, 9), -6)])
    # most cells moderate throughout, so that single extreme values meet otherwise ordinary cells
    calm = rng.random((NY, NX)) < 0.5
    fm = np.stack([rng.uniform(0.05, 1.0, (NY, NX, 9)), 1800 * rng.uniform(0.05, 1.0, (NY, NX, 9)), 1e9 * rng.uniform(0.05, 1.0, (NY, NX, 9))])
    one = rng.integers(0, 9, (NY, NX))
    for s in range(3):
        keep = np.zeros((NY, NX, 9), bool)
        keep[np.arange(NY)[:, None], np.arange(NX)[None, :], one] = True      # one extreme direction survives in a calm cell
        f[s] = np.where(calm[..., None] & ~keep, fm[s], f[s])
    Ex = draw((NY, NX), -10) * np.where(rng.random((NY, NX)) < 0.5, 1, -1); Ey = draw((NY, NX), -10)
    _one_step_against_checker(oracle, plbm, f, g, Ex, Ey, f"fuzz seed {seed}")


@pytest.mark.parametrize("NX,NY,poisson,bc,nsteps", [
    (24, 24, "sor", "bounceback", 6), (24, 24, "gs", "periodic", 6), (24, 24, "nps", "bounceback", 6),
    (24, 24, "fft", "bounceback", 6), (30, 22, "none", "bounceback", 8), (32, 32, "sor", "periodic", 5),
    (40, 40, "nps", "periodic", 4), (33, 27, "gs", "bounceback", 5),
])
def test_walls_and_iterative_solvers(oracle, plbm, NX, NY, poisson, bc, nsteps):
    """bounce-back streaming with the reference's corner quirks (stale temp_* slots, last-writer-wins),
    Dirichlet GS / SOR / 9-point solvers with warm start and the reference's exact stopping iteration, Neumann rim."""
    run_both(oracle, plbm, NX, NY, poisson, nsteps, {0, 1, nsteps - 1}, bc=bc)


@pytest.mark.parametrize("name", ["n24_sor_bounceback", "n24_gs_periodic", "n24_nps_bounceback", "n24_fft_bounceback",
                                  "n32_fft_periodic", "n32_none_periodic", "n30x20_fft_periodic"])
def test_against_golden_vectors_of_the_unmodified_reference(plbm, name):
    from pathlib import Path
    z = np.load(Path(__file__).resolve().parent / "golden" / f"{name}.npz")
    NX, NY, steps = int(z["NX"]), int(z["NY"]), int(z["steps"])
    dumps = [int(t) for t in z["dump_steps"]]
    with plbm.PlasmaLBM(NX, NY, poisson=str(z["poisson"]), bc=str(z["bc"])) as sim:
        for t in range(steps):
            sim.step(1, want_fields=t in dumps)
            if t in dumps:
                got = sim.fields()
                for fname in plbm.FIELD_NAMES:
                    assert_same_bits(got[fname], z[f"t{t}_{fname}"], f"{name}: {fname} at step {t}")
        f, g = sim.download_state()
    for s in range(3):
        assert_same_bits(f[s], z["pops_f"][s], f"{name}: f[{s}]")
        assert_same_bits(g[s], z["pops_g"][s], f"{name}: g[{s}]")


@pytest.mark.parametrize("name,solver", [("n32x24_gs_periodic_solver", "gs_periodic"), ("n32x24_sor_periodic_solver", "sor_periodic"),
                                         ("n32x24_nps_periodic_solver", "nps_periodic")])
def test_periodic_iterative_solvers_against_the_unmodified_reference(plbm, name, solver):
    """poisson::SolvePoisson_{GS,SOR,9point}_Periodic (public in include/poisson.hpp, src/poisson.cpp:146-211, 283-354, 487-546):
    phi after two warm-started calls on a fixed right-hand side, bit for bit against vectors the unmodified reference produced
    (tests/golden/make_golden.py: periodic_solvers)."""
    z = np.load(GOLDEN / f"{name}.npz")
    NX, NY = int(z["NX"]), int(z["NY"])
    with plbm.PlasmaLBM(NX, NY, poisson="gs", fields_only=True) as sim:
        sim.solve_poisson(z["rho_q"] * 0.0)                   # the reference sizes and zeroes phi in its first SolvePoisson call
        for _ in range(int(z["calls"])):
            sim.poisson_solver(solver, z["rho_q"], omega=float(z["omega"]))
        assert_same_bits(sim.fields(["phi"])["phi"], z["phi"], f"{name}: phi")


def test_host_poisson_dispatch_is_per_call(plbm):
    """type, boundary and omega are arguments of every poisson:: call in the reference; a fields-only context must not
    keep the first call's values.  GS then SOR(omega = 1.8) must differ from GS then SOR(omega = 0) (which leaves phi unchanged),
    and an FFT request on a context created for GS builds its plan on demand."""
    NX = NY = 32
    rng = np.random.default_rng(5)
    rho_q = rng.normal(0, 1e-3, size=(NY, NX))
    with plbm.PlasmaLBM(NX, NY, poisson="gs", fields_only=True) as sim:
        sim.poisson_solver("gs", rho_q)
        after_gs = sim.fields(["phi"])["phi"].copy()
        sim.poisson_solver("sor", rho_q, omega=1.8)
        after_sor = sim.fields(["phi"])["phi"].copy()
        assert not np.array_equal(after_gs, after_sor)
        sim.poisson_solver("fft", rho_q)
        after_fft = sim.fields(["phi"])["phi"]
        assert np.isfinite(after_fft).all() and not np.array_equal(after_fft, after_sor)


def test_nan_and_inf_cells(oracle, plbm):
    """Cells whose inputs already hold NaN / Inf (what the reference's own dynamics produce on large lattices):
    NaN cells skip the IEEE fallback, Inf cells take it; either way the result equals the checker's, NaN for NaN."""
    NX, NY = 64, 16
    rng = np.random.default_rng(7)
    f = rng.uniform(0.05, 1.0, size=(3, NY, NX, 9)); g = rng.uniform(0.01, 0.5, size=(3, NY, NX, 9))
    f[2] *= 1e9
    f[0][:, 0:8, 3] = np.nan          # electrons NaN in one direction
    f[1][:, 8:16, 0] = np.inf         # ions Inf
    f[2][:, 16:24, 5] = -np.inf
    g[0][:, 24:32, 2] = np.nan        # NaN only in a thermal population
    f[1][:, 32:40, :] = np.nan
    f[0][:, 40:44, 1] = np.nan; f[1][:, 40:44, :] *= 1e-310   # NaN cell that ALSO has tiny numerators elsewhere
    Ex = rng.normal(0, 1e-3, size=(NY, NX)); Ey = rng.normal(0, 1e-3, size=(NY, NX))
    Ex[:, 48:52] = np.nan
    o = oracle.PortOracle(NX, NY, poisson="none", initialize=False)
    for s in range(3):
        o.f(s)[...] = f[s]; o.g(s)[...] = g[s]
    o.scalar(oracle.PO_EX)[...] = Ex; o.scalar(oracle.PO_EY)[...] = Ey
    with plbm.PlasmaLBM(NX, NY, poisson="none", initialize=False) as sim:
        sim.upload_state(f, g); sim.set_efield(Ex, Ey)
        with np.errstate(all="ignore"):
            for t in range(3):
                o.step(1); sim.step(1, want_fields=True)
                assert_fields_same(sim.fields(), o.fields(), f"nan/t={t}")
            f3, g3 = sim.download_state()
            for s in range(3):
                assert_same_bits(f3[s], o.f(s), f"nan f[{s}]"); assert_same_bits(g3[s], o.g(s), f"nan g[{s}]")


@pytest.mark.parametrize("poisson,bc", [("fft", "periodic"), ("sor", "bounceback")])
def test_pipelined_fetch_delivers_every_step(oracle, plbm, poisson, bc):
    """plbm_fetch_begin / plbm_fetch_wait (LBmethod::Run_simulation's loop): the copy of step t is pending while
    step t+1 already runs, and must still deliver step t's fields -- all 16, every step."""
    NX = NY = 48
    nsteps = 9
    o = oracle.PortOracle(NX, NY, poisson=poisson, bc=bc)
    bufs = [np.full((16, NY, NX), np.nan), np.full((16, NY, NX), np.nan)]
    with plbm.PlasmaLBM(NX, NY, poisson=poisson, bc=bc) as sim:
        sim.step(1, want_fields=True)
        sim.fetch_begin(bufs[0])
        for t in range(nsteps):
            if t + 1 < nsteps:
                sim.step(1, want_fields=True)            # issued while the fetch of step t is pending
            got = sim.fetch_wait()
            assert got is bufs[t & 1]
            if t + 1 < nsteps:
                sim.fetch_begin(bufs[(t + 1) & 1])
            o.step(1)
            assert_fields_same({n: got[k] for k, n in enumerate(plbm.FIELD_NAMES)}, o.fields(), f"pipelined {poisson}/{bc} t={t}")
        sim.fetch_begin(bufs[0])
        with pytest.raises(plbm.PlbmError):
            sim.fetch_begin(bufs[1])                     # a second fetch before fetch_wait is refused
        sim.fetch_wait()
    o.close()


@pytest.mark.parametrize("poisson", ["none", "sor"])
def test_walls_through_the_unfused_sweep_sequence(oracle, plbm, monkeypatch, poisson):
    """PLBM_UNFUSED_WALLS=1 keeps the first wall implementation reachable: the reference's own sequence of sweeps on
    its own array set (the default is the fused kernel with the walls pull, csrc/walls.cuh)."""
    monkeypatch.setenv("PLBM_UNFUSED_WALLS", "1")
    run_both(oracle, plbm, 30, 26, poisson, 6, {0, 1, 5}, bc="bounceback")


@pytest.mark.parametrize("poisson,bc,nsteps", [("fft", "periodic", 41), ("none", "periodic", 30), ("sor", "bounceback", 26), ("none", "bounceback", 33)])
def test_long_runs_replayed_from_a_cuda_graph(oracle, plbm, poisson, bc, nsteps):
    """plbm_step with many steps on a small lattice records two steps into a CUDA graph and replays it; the result must be
    the one of step-by-step execution (odd and even step counts, moments of the last step included)."""
    NX, NY = 40, 36
    o = oracle.PortOracle(NX, NY, poisson=poisson, bc=bc)
    o.step(nsteps)
    with plbm.PlasmaLBM(NX, NY, poisson=poisson, bc=bc) as sim:
        sim.step(nsteps, want_fields=True)
        assert_fields_same(sim.fields(), o.fields(), f"graph replay {poisson}/{bc}/{nsteps}")
        f, g = sim.download_state()
        sim.step(nsteps + 1)                         # a second long call on the same context
        f2, g2 = sim.download_state()
    for s in range(3):
        assert_same_bits(f[s], o.f(s), f"graph replay: f[{s}]")
        assert_same_bits(g[s], o.g(s), f"graph replay: g[{s}]")
    o.step(nsteps + 1)
    for s in range(3):
        assert_same_bits(f2[s], o.f(s), f"graph replay, second call: f[{s}]")
        assert_same_bits(g2[s], o.g(s), f"graph replay, second call: g[{s}]")
    o.close()
