"""The reference-facing C++ surface (include/plasma.hpp, collisions.hpp, streaming.hpp, poisson.hpp):
builds everywhere; on a GPU, LBmethod::Run_simulation and the free functions reproduce the CPU
checker bit for bit, and the reference's own src/main_plasma.cpp (compiled UNCHANGED against this
repo's headers, where /root/reference is mounted at build time) reproduces the golden vectors."""
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np
import pytest

from helpers import assert_same_bits

ROOT = Path(__file__).resolve().parent.parent
HOST = ROOT / "tests" / "host"
BUILD = HOST / "_build"
NAMES15 = ("ux_e", "uy_e", "ux_i", "uy_i", "ux_n", "uy_n", "T_e", "T_i", "T_n", "rho_e", "rho_i", "rho_n", "rho_q", "Ex", "Ey")


def build_drivers():
    subprocess.run(["make", "-s", "-j8", "-C", str(ROOT / "12-lb-12-lb_b200"), "host"], check=True)
    subprocess.run(["make", "-s", "-C", str(HOST), "all"], check=True)
    if Path("/root/reference/src/main_plasma.cpp").exists():
        subprocess.run(["make", "-s", "-C", str(HOST), "_build/reference_main"], check=True)


def test_cpp_surface_builds(plbm):
    """libplasma_b200.so and the drivers compile and link; where the reference is mounted, its
    main_plasma.cpp builds unchanged against include/*.hpp."""
    build_drivers()
    assert (BUILD / "drive_lbmethod").exists() and (BUILD / "drive_phases").exists()
    if Path("/root/reference/src/main_plasma.cpp").exists():
        assert (BUILD / "reference_main").exists()


def test_reference_renderer_type_checks_against_our_header():
    """the reference's renderer src/visualize.cpp (OpenCV, out of scope) keeps compiling UNCHANGED against
    this repo's include/visualize.hpp.  The image has no OpenCV, so the check is g++ -fsyntax-only with declarations of the OpenCV
    calls the renderer makes (tests/host/opencv_shim, OpenCV's own signatures): every symbol the renderer needs from visualize.hpp
    (namespace, constants, the three entry points and its helper prototypes) is there with a compatible type."""
    src = Path("/root/reference/src/visualize.cpp")
    if not src.exists():
        pytest.skip("needs /root/reference (only in the build container)")
    r = subprocess.run(["g++", "-std=c++20", "-fsyntax-only", "-fopenmp", f"-I{HOST / 'opencv_shim'}", f"-I{ROOT / 'include'}", str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_lbmethod_fails_loudly_without_gpu(plbm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    build_drivers()
    r = subprocess.run([str(BUILD / "drive_lbmethod"), "16", "16", "2", "3", "0"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr


def run_driver(exe, args, dump_steps, NX, NY, cwd=None, extra_env=None):
    with tempfile.TemporaryDirectory(prefix="plbm_cpp_") as tmp:
        env = dict(os.environ, PLBM_DUMP_DIR=tmp, PLBM_DUMP_STEPS=",".join(map(str, dump_steps)), **(extra_env or {}))
        r = subprocess.run([str(exe)] + [str(a) for a in args], capture_output=True, text=True, env=env, cwd=cwd or tmp)
        assert r.returncode == 0, r.stderr
        out = {}
        for t in dump_steps:
            raw = np.fromfile(os.path.join(tmp, f"fields_t{t:05d}.f64"), dtype=np.float64).reshape(15, NY, NX)
            out[t] = {n: raw[k] for k, n in enumerate(NAMES15)}
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("NX,NY,steps,poisson,bc", [(64, 64, 20, "fft", "periodic"), (48, 40, 8, "none", "periodic"),
                                                     (28, 28, 5, "sor", "bounceback"), (26, 30, 5, "nps", "periodic")])
def test_lbmethod_run_simulation(oracle, plbm, NX, NY, steps, poisson, bc):
    build_drivers() if not (BUILD / "drive_lbmethod").exists() else None
    dumps = sorted({0, 1, steps - 1})
    got = run_driver(BUILD / "drive_lbmethod", [NX, NY, steps, oracle.POISSON[poisson], oracle.BC[bc]], dumps, NX, NY)
    o = oracle.PortOracle(NX, NY, poisson=poisson, bc=bc)
    want = o.run_with_dumps(steps, dumps)
    for t in dumps:
        for n in NAMES15:
            assert_same_bits(got[t][n], want[t][n], f"LBmethod {NX}x{NY}/{poisson}/{bc} step {t}: {n}")


def expected_frames_and_series(f: dict, NX: int, NY: int):
    """What reference src/visualize.cpp:168-315 derives from the FP64 fields before its first OpenCV call."""
    mag = lambda a, b: np.sqrt(a * a + b * b)                       # double, as in the reference
    frames = [f["rho_e"], f["rho_i"], f["rho_q"], f["ux_e"], f["uy_e"], mag(f["ux_e"], f["uy_e"]),
              f["ux_i"], f["uy_i"], mag(f["ux_i"], f["uy_i"]), f["T_e"], f["T_i"], f["T_n"]]
    frames = np.stack([a.astype(np.float32) for a in frames])
    cx, cy, dx, dy = NX // 2, NY // 2, NX // 4, NY // 4
    pts = [(cx, cy), (cx + dx, cy), (cx - dx, cy), (cx, cy + dy), (cx, cy - dy),
           (cx + dx, cy + dy), (cx + dx, cy - dy), (cx - dx, cy + dy), (cx - dx, cy - dy)]
    at = lambda a: np.array([a[j, i] for i, j in pts])
    rows = []
    for sp in "ein":
        rows += [at(f[f"ux_{sp}"]), at(f[f"uy_{sp}"]), at(mag(f[f"ux_{sp}"], f[f"uy_{sp}"]))]
    rows += [at(f[f"T_{sp}"]) for sp in "ein"] + [at(f[f"rho_{sp}"]) for sp in "ein"]
    rows += [at(f["rho_q"]), at(f["Ex"]), at(f["Ey"]), at(mag(f["Ex"], f["Ey"]))]
    return frames, np.stack(rows)


@pytest.mark.gpu
@pytest.mark.parametrize("NX,NY,steps,poisson,bc", [(64, 64, 12, "fft", "periodic"), (36, 28, 5, "sor", "bounceback")])
def test_lbmethod_alternate_output_path(oracle, plbm, NX, NY, steps, poisson, bc):
    """Run_simulation_frames: the visualiser's CV_32F matrices and sample series formed on the device are, bit for
    bit, what the unchanged visualiser would have derived from the FP64 fields."""
    build_drivers() if not (BUILD / "drive_lbmethod").exists() else None
    dumps = sorted({0, 1, steps - 1})
    with tempfile.TemporaryDirectory(prefix="plbm_frames_") as tmp:
        env = dict(os.environ, PLBM_DUMP_DIR=tmp, PLBM_DUMP_STEPS=",".join(map(str, dumps)))
        r = subprocess.run([str(BUILD / "drive_lbmethod"), str(NX), str(NY), str(steps), str(oracle.POISSON[poisson]), str(oracle.BC[bc]), "1"],
                           capture_output=True, text=True, env=env, cwd=tmp)
        assert r.returncode == 0, r.stderr
        got = {}
        for t in dumps:
            raw = np.fromfile(os.path.join(tmp, f"frames_t{t:05d}.bin"), dtype=np.uint8)
            nf = 12 * NX * NY * 4
            got[t] = (raw[:nf].view(np.float32).reshape(12, NY, NX), raw[nf:].view(np.float64).reshape(19, 9))
    o = oracle.PortOracle(NX, NY, poisson=poisson, bc=bc)
    want = o.run_with_dumps(steps, dumps)
    for t in dumps:
        frames, series = expected_frames_and_series(want[t], NX, NY)
        assert got[t][0].tobytes() == frames.tobytes() or np.array_equal(got[t][0], frames), f"frames differ at step {t}"
        assert_same_bits(got[t][1], series, f"series at step {t}")


@pytest.mark.gpu
@pytest.mark.parametrize("NX,NY,steps,poisson,bc", [(40, 40, 6, "fft", "periodic"), (24, 36, 4, "none", "periodic"),
                                                     (22, 22, 4, "gs", "bounceback"), (24, 20, 4, "sor", "periodic")])
def test_free_functions_phase_by_phase(oracle, plbm, NX, NY, steps, poisson, bc):
    """collisions::Collide, streaming::Stream, poisson::SolvePoisson on host vectors."""
    build_drivers() if not (BUILD / "drive_phases").exists() else None
    dumps = sorted({0, steps - 1})
    got = run_driver(BUILD / "drive_phases", [NX, NY, steps, oracle.POISSON[poisson], oracle.BC[bc]], dumps, NX, NY)
    o = oracle.PortOracle(NX, NY, poisson=poisson, bc=bc)
    want = o.run_with_dumps(steps, dumps)
    for t in dumps:
        for n in NAMES15:
            assert_same_bits(got[t][n], want[t][n], f"free functions {NX}x{NY}/{poisson}/{bc} step {t}: {n}")


@pytest.mark.gpu
def test_unchanged_reference_main_on_the_gpu_library():
    """reference src/main_plasma.cpp (200x200, 200 steps, FFT, periodic) linked against this repo:
    sample points and sums of every field against the golden vectors of the unmodified CPU reference."""
    exe = BUILD / "reference_main"
    if not exe.exists():
        pytest.skip("reference_main was not built (needs /root/reference at build time)")
    z = np.load(ROOT / "tests" / "golden" / "n200_fft_periodic_default.npz")
    dumps = [int(t) for t in z["dump_steps"]]
    with tempfile.TemporaryDirectory(prefix="plbm_main_") as cwd:
        os.makedirs(os.path.join(cwd, "build"))           # main_plasma.cpp appends its timing CSV there
        got = run_driver(exe, [1], dumps, 200, 200, cwd=cwd)
        assert "200x200,200,1,3,0," in Path(cwd, "build", "simulation_time_plasma_details.csv").read_text()
    pts = [tuple(p) for p in z["points"]]
    for t in dumps:
        for n in NAMES15:
            v = got[t][n]
            assert_same_bits(np.array([v[y, x] for (x, y) in pts]), z[f"t{t}_{n}_points"], f"{n} points at step {t}")
            stats = np.array([v.sum(), v.min(), v.max(), np.abs(v).sum()])
            assert_same_bits(stats, z[f"t{t}_{n}_stats"], f"{n} stats at step {t}")


def _gpu_count():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.gpu
def test_unchanged_reference_main_on_several_gpus():
    """The same unchanged src/main_plasma.cpp with PLBM_DEVICES=N in the environment: LBmethod cuts the 200x200 lattice into N
    y-slabs on N GPUs of the box (one process, one host thread per device, peer memory between the slabs) and still reproduces
    the golden vectors of the CPU reference bit for bit.  Needs >= 2 GPUs."""
    n = _gpu_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    exe = BUILD / "reference_main"
    if not exe.exists():
        pytest.skip("reference_main was not built (needs /root/reference at build time)")
    z = np.load(ROOT / "tests" / "golden" / "n200_fft_periodic_default.npz")
    dumps = [int(t) for t in z["dump_steps"]]
    pts = [tuple(p) for p in z["points"]]
    for ndev in sorted({2, min(n, 8)}):
        with tempfile.TemporaryDirectory(prefix="plbm_main_") as cwd:
            os.makedirs(os.path.join(cwd, "build"))
            got = run_driver(exe, [1], dumps, 200, 200, cwd=cwd, extra_env={"PLBM_DEVICES": str(ndev)})
        for t in dumps:
            for name in NAMES15:
                v = got[t][name]
                assert_same_bits(np.array([v[y, x] for (x, y) in pts]), z[f"t{t}_{name}_points"], f"{ndev} GPUs: {name} points at step {t}")
                stats = np.array([v.sum(), v.min(), v.max(), np.abs(v).sum()])
                assert_same_bits(stats, z[f"t{t}_{name}_stats"], f"{ndev} GPUs: {name} stats at step {t}")


@pytest.mark.gpu
@pytest.mark.parametrize("NX,NY,steps", [(64, 64, 12), (96, 96, 6)])      # several slabs need NX == NY (the reference's FFT reshape quirk)
def test_lbmethod_on_several_gpus(oracle, NX, NY, steps):
    """LBmethod::Run_simulation with PLBM_DEVICES=2 (and 4 where present) against the single-domain CPU checker, all 15 fields."""
    n = _gpu_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    build_drivers() if not (BUILD / "drive_lbmethod").exists() else None
    dumps = sorted({0, 1, steps - 1})
    o = oracle.PortOracle(NX, NY, poisson="fft")
    want = o.run_with_dumps(steps, dumps)
    for ndev in sorted({2, min(n, 4)}):
        got = run_driver(BUILD / "drive_lbmethod", [NX, NY, steps, oracle.POISSON["fft"], oracle.BC["periodic"]], dumps, NX, NY,
                         extra_env={"PLBM_DEVICES": str(ndev)})
        for t in dumps:
            for name in NAMES15:
                assert_same_bits(got[t][name], want[t][name], f"LBmethod on {ndev} GPUs {NX}x{NY} step {t}: {name}")
