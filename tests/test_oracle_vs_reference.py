"""The CPU restatement against the reference itself, run live (only where oracle/_ref was built,
i.e. where /root/reference is mounted or its binaries travelled).  Bar: bit-exact."""
import pytest

from helpers import assert_fields_same, assert_same_bits


def _need_ref(oracle):
    if not oracle.have_reference("parity"):
        pytest.skip("oracle/_ref/ref_plasma_parity not built (needs /root/reference)")


@pytest.mark.parametrize("NX,NY,steps,poisson,bc", [
    (40, 40, 12, "fft", "periodic"),
    (36, 28, 6, "none", "periodic"),
    (20, 20, 5, "sor", "periodic"),
    (20, 20, 5, "gs", "bounceback"),
    (21, 21, 5, "nps", "periodic"),
    (18, 22, 5, "none", "bounceback"),
])
def test_port_equals_live_reference(oracle, NX, NY, steps, poisson, bc):
    _need_ref(oracle)
    dumps = sorted({0, steps // 2, steps - 1})
    info, ref, pops = oracle.run_reference(NX, NY, steps, poisson=poisson, bc=bc, threads=2, dump_steps=dumps, pops=True)
    o = oracle.PortOracle(NX, NY, poisson=poisson, bc=bc)
    got = o.run_with_dumps(steps, dumps)
    for t in dumps:
        assert_fields_same(got[t], ref[t], f"{NX}x{NY}/{poisson}/{bc}/t={t}")
    for s in range(3):
        assert_same_bits(o.f(s), pops["f"][s], f"f[{s}]")
        assert_same_bits(o.g(s), pops["g"][s], f"g[{s}]")


def test_reference_is_thread_count_invariant(oracle):
    """The reference's loops are element-independent: 1 and 4 OpenMP threads give identical bits."""
    _need_ref(oracle)
    _, a, _ = oracle.run_reference(32, 32, 6, poisson="fft", threads=1, dump_steps=[5])
    _, b, _ = oracle.run_reference(32, 32, 6, poisson="fft", threads=4, dump_steps=[5])
    assert_fields_same(a[5], b[5], "threads 1 vs 4")


def test_timing_build_stays_within_noise_floor(oracle):
    """The reference compiled with its own flags (FMA contraction allowed) differs from the parity
    build only at rounding level over a few steps (BASELINE.md noise-floor table)."""
    _need_ref(oracle)
    if not oracle.have_reference("timing"):
        pytest.skip("timing build missing")
    _, a, _ = oracle.run_reference(48, 48, 10, poisson="fft", threads=2, dump_steps=[9], kind="parity")
    _, b, _ = oracle.run_reference(48, 48, 10, poisson="fft", threads=2, dump_steps=[9], kind="timing")
    for name in ("rho_e", "ux_e", "T_e", "rho_q", "Ex", "phi"):
        assert oracle.max_norm_err(b[9][name], a[9][name]) < 1e-12, name
