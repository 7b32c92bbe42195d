#!/usr/bin/env python
"""Generates the golden vectors in this directory by RUNNING THE UNMODIFIED REFERENCE
(oracle/_ref/ref_plasma_parity = /root/reference sources compiled in place with -ffp-contract=off,
FFTW replaced by oracle/fft_oracle.c).  The reference ships no tests or vectors of its own
(SURVEY.md section 4), so these are the pins of the CPU restatement and of the CUDA path.

    python tests/golden/make_golden.py          # needs /root/reference (only in the build container)
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402

HERE = Path(__file__).resolve().parent

SMALL = [  # name, NX, NY, steps, poisson, bc, dump steps
    ("n32_none_periodic", 32, 32, 16, "none", "periodic", (0, 3, 15)),
    ("n32_fft_periodic", 32, 32, 16, "fft", "periodic", (0, 3, 15)),
    ("n30x20_fft_periodic", 30, 20, 8, "fft", "periodic", (0, 7)),
    ("n24_sor_bounceback", 24, 24, 6, "sor", "bounceback", (0, 5)),
    ("n24_gs_periodic", 24, 24, 6, "gs", "periodic", (0, 5)),
    ("n24_nps_bounceback", 24, 24, 6, "nps", "bounceback", (0, 5)),
    ("n24_fft_bounceback", 24, 24, 6, "fft", "bounceback", (0, 5)),
]
POINTS_200 = [(100, 100), (50, 50), (150, 50), (50, 150), (150, 150), (100, 60), (100, 140), (60, 100), (140, 100), (3, 7)]


PERIODIC_SOLVERS = [  # name, harness code, NX, NY, calls, omega -- even sizes: with an odd extent cells of one colour are neighbours
    ("n32x24_gs_periodic_solver", 1, 32, 24, 2, 0.0),          # across the wrap and the reference's own OpenMP sweep is order-dependent
    ("n32x24_sor_periodic_solver", 2, 32, 24, 2, 1.8),
    ("n32x24_nps_periodic_solver", 4, 32, 24, 2, 0.0),
]


def periodic_rho_q(NX, NY):
    """Deterministic zero-mean right-hand side for the periodic solvers (no RNG: the same array anywhere)."""
    y, x = np.mgrid[0:NY, 0:NX]
    r = 1e-3 * (np.sin(2 * np.pi * x / NX) * np.cos(4 * np.pi * y / NY) + 0.5 * np.cos(6 * np.pi * x / NX + 2 * np.pi * y / NY))
    r[NY // 3, NX // 2] += 2e-3
    r[2 * NY // 3, NX // 4] -= 2e-3
    return np.ascontiguousarray(r - r.mean())


def periodic_solvers():
    """poisson::SolvePoisson_{GS,SOR,9point}_Periodic of the unmodified reference (never called by its own loop): phi after
    `calls` warm-started calls on a fixed rho_q, through the harness mode --periodic-solver."""
    import subprocess, tempfile
    for name, code, NX, NY, calls, omega in PERIODIC_SOLVERS:
        rho_q = periodic_rho_q(NX, NY)
        with tempfile.TemporaryDirectory(prefix="plbm_ref_") as tmp:
            rho_q.tofile(f"{tmp}/rhoq.f64")
            subprocess.run([str(O.ref_binary("parity")), "--nx", str(NX), "--ny", str(NY), "--threads", "2", "--omega", repr(omega),
                            "--periodic-solver", str(code), "--calls", str(calls), "--rhoq", f"{tmp}/rhoq.f64", "--out", tmp], check=True,
                           capture_output=True)
            phi = np.fromfile(f"{tmp}/phi_periodic.f64", dtype=np.float64).reshape(NY, NX)
        np.savez_compressed(HERE / f"{name}.npz", NX=NX, NY=NY, calls=calls, omega=omega, solver=code, rho_q=rho_q, phi=phi)
        print("wrote", name, "max|phi| =", np.abs(phi).max())


def main():
    O.build(ref=True)
    if "--periodic-solvers-only" in sys.argv:
        periodic_solvers()
        return
    periodic_solvers()
    for name, NX, NY, steps, poisson, bc, dumps in SMALL:
        info, fields, pops = O.run_reference(NX, NY, steps, poisson=poisson, bc=bc, threads=2, dump_steps=dumps, pops=True)
        out = {"NX": NX, "NY": NY, "steps": steps, "poisson": poisson, "bc": bc, "dump_steps": np.array(dumps)}
        for t in dumps:
            for k, v in fields[t].items():
                out[f"t{t}_{k}"] = v
        out["pops_f"] = pops["f"]
        out["pops_g"] = pops["g"]
        for k, v in info["units"].items():
            out[f"units_{k}"] = v
        np.savez_compressed(HERE / f"{name}.npz", **out)
        print("wrote", name)
    # the reference's default case (src/main_plasma.cpp:16-51): sample points, sums and extrema
    dumps = (0, 1, 10, 50, 100, 150, 199)
    info, fields, _ = O.run_reference(200, 200, 200, poisson="fft", bc="periodic", threads=8, dump_steps=dumps)
    out = {"dump_steps": np.array(dumps), "points": np.array(POINTS_200)}
    for t in dumps:
        for k, v in fields[t].items():
            out[f"t{t}_{k}_points"] = np.array([v[y, x] for (x, y) in POINTS_200])
            out[f"t{t}_{k}_stats"] = np.array([v.sum(), v.min(), v.max(), np.abs(v).sum()])
    np.savez_compressed(HERE / "n200_fft_periodic_default.npz", **out)
    print("wrote n200_fft_periodic_default")


if __name__ == "__main__":
    main()
