"""The CPU restatement (oracle/plasma_oracle.c) against golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py).  Bar: bit-exact.  Covers every Poisson type and both
boundary types the reference offers, the NX != NY reshape quirk, and the default 200x200 case."""
from pathlib import Path

import numpy as np
import pytest

from helpers import assert_same_bits

GOLD = Path(__file__).resolve().parent / "golden"
SMALL = ["n32_none_periodic", "n32_fft_periodic", "n30x20_fft_periodic", "n24_sor_bounceback",
         "n24_gs_periodic", "n24_nps_bounceback", "n24_fft_bounceback"]


@pytest.mark.parametrize("name", SMALL)
def test_port_matches_reference_golden(oracle, name):
    z = np.load(GOLD / f"{name}.npz")
    NX, NY, steps = int(z["NX"]), int(z["NY"]), int(z["steps"])
    o = oracle.PortOracle(NX, NY, poisson=str(z["poisson"]), bc=str(z["bc"]))
    for k in ("cs2", "Kb"):
        assert getattr(o.units, k) == float(z[f"units_{k}"])
    assert o.units.m[1] == float(z["units_m_i"]) and o.units.Ex_ext == float(z["units_Ex_ext"])
    assert o.units.rho_init[2] == float(z["units_rho_n_init"]) and o.units.T_init[1] == float(z["units_T_i_init"])
    got = o.run_with_dumps(steps, [int(t) for t in z["dump_steps"]])
    for t, fields in got.items():
        for fname, v in fields.items():
            assert_same_bits(v, z[f"t{t}_{fname}"], f"{name}: {fname} at step {t}")
    for s in range(3):
        assert_same_bits(o.f(s), z["pops_f"][s], f"{name}: f[{s}] after {steps} steps")
        assert_same_bits(o.g(s), z["pops_g"][s], f"{name}: g[{s}] after {steps} steps")


def test_port_matches_reference_default_case(oracle):
    """200x200, 200 steps, FFT, periodic (reference src/main_plasma.cpp:16-51): sample points and sums."""
    z = np.load(GOLD / "n200_fft_periodic_default.npz")
    pts = [tuple(p) for p in z["points"]]
    dumps = [int(t) for t in z["dump_steps"]]
    o = oracle.PortOracle(200, 200, poisson="fft")
    got = o.run_with_dumps(200, dumps)
    for t in dumps:
        for fname, v in got[t].items():
            assert_same_bits(np.array([v[y, x] for (x, y) in pts]), z[f"t{t}_{fname}_points"], f"{fname} points at step {t}")
            stats = np.array([v.sum(), v.min(), v.max(), np.abs(v).sum()])
            assert_same_bits(stats, z[f"t{t}_{fname}_stats"], f"{fname} stats at step {t}")
