"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

Python access to the two CPU checkers of the plasma LBM hot path:

* ``PortOracle``  -- ctypes wrapper of ``liboracle_port.so`` (oracle/plasma_oracle.c, the C
  restatement of /root/reference/src/{plasma,collisions,streaming,poisson}.cpp).
* ``run_reference`` -- runs ``oracle/_ref/ref_plasma_{parity,timing}``, the UNMODIFIED reference
  sources compiled in place by oracle/Makefile, and loads the fields its visualisation hook dumped.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (12-lb-12-lb_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
PORT_LIB = HERE / "liboracle_port.so"
REF_DIR = HERE / "_ref"

# order in which the reference hands its fields to visualize::UpdateVisualization
# (/root/reference/include/visualize.hpp:53-61), plus phi from the harness accessor
FIELD_NAMES = ("ux_e", "uy_e", "ux_i", "uy_i", "ux_n", "uy_n", "T_e", "T_i", "T_n",
               "rho_e", "rho_i", "rho_n", "rho_q", "Ex", "Ey", "phi")

POISSON = {"none": 0, "gs": 1, "sor": 2, "fft": 3, "nps": 4}
BC = {"periodic": 0, "bounceback": 1}

# default physical parameters, /root/reference/src/main_plasma.cpp:16-51
DEFAULT_SI = dict(Z_ion=1, A_ion=1, Ex_SI=1e-2, Ey_SI=0.0, T_e_SI=1e4, T_i_SI=300.0, T_n_SI=300.0,
                  n_e_SI=1e11, n_n_SI=1e18)

PO_F, PO_G, PO_TMP, PO_FEQ_SELF, PO_GEQ_SELF, PO_FEQ_CROSS, PO_GEQ_CROSS = range(7)
PO_RHO, PO_UX, PO_UY, PO_T, PO_UX_PAIR, PO_UY_PAIR, PO_EX, PO_EY, PO_RHO_Q, PO_PHI = range(7, 17)


class Units(C.Structure):
    _fields_ = [("cs2", C.c_double), ("Kb", C.c_double), ("Ex_ext", C.c_double), ("Ey_ext", C.c_double),
                ("T_init", C.c_double * 3), ("m", C.c_double * 3), ("q", C.c_double * 3),
                ("rho_init", C.c_double * 3)]


def build(ref: bool = True) -> None:
    """Compile the checker(s).  `make ref` needs /root/reference and is skipped without it."""
    subprocess.run(["make", "-s", "-C", str(HERE), "port"], check=True)
    if ref and Path(os.environ.get("PLBM_REFERENCE", "/root/reference")).is_dir():
        subprocess.run(["make", "-s", "-C", str(HERE), "ref"], check=True)


_lib = None


def port_lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not PORT_LIB.exists():
            build(ref=False)
        lib = C.CDLL(str(PORT_LIB))
        lib.po_units_from_si.argtypes = [C.c_int, C.c_int] + [C.c_double] * 7 + [C.POINTER(Units)]
        lib.po_create.restype = C.c_void_p
        lib.po_create.argtypes = [C.c_int, C.c_int, C.POINTER(Units), C.c_int, C.c_int, C.c_double]
        lib.po_field.restype = C.POINTER(C.c_double)
        lib.po_field.argtypes = [C.c_void_p, C.c_int, C.c_int]
        for name in ("po_destroy", "po_initialize", "po_update_macro", "po_compute_equilibrium",
                     "po_thermal_collisions", "po_collisions", "po_stream", "po_solve_poisson"):
            getattr(lib, name).argtypes = [C.c_void_p]
            getattr(lib, name).restype = None
        lib.po_step.argtypes = [C.c_void_p, C.c_int]
        lib.po_step.restype = None
        lib.offt_plan2d_create.restype = C.c_void_p
        lib.offt_plan2d_create.argtypes = [C.c_int, C.c_int]
        lib.offt_plan2d_destroy.argtypes = [C.c_void_p]
        lib.offt_r2c_2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.offt_c2r_2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.offt_rows_fwd.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        lib.offt_rows_inv.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        lib.offt_col.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.offt_factorize.argtypes = [C.c_int, C.POINTER(C.c_int)]
        lib.offt_factorize.restype = C.c_int
        _lib = lib
    return _lib


def units_from_si(**si) -> Units:
    p = dict(DEFAULT_SI)
    p.update(si)
    u = Units()
    port_lib().po_units_from_si(p["Z_ion"], p["A_ion"], p["Ex_SI"], p["Ey_SI"], p["T_e_SI"], p["T_i_SI"],
                                p["T_n_SI"], p["n_e_SI"], p["n_n_SI"], C.byref(u))
    return u


class PortOracle:
    """One simulation state of the C restatement.  Arrays are numpy views into its memory."""

    def __init__(self, NX: int, NY: int, poisson: str = "fft", bc: str = "periodic", omega: float = 1.8,
                 units: Units | None = None, initialize: bool = True, **si):
        self.lib = port_lib()
        self.NX, self.NY = NX, NY
        self.units = units if units is not None else units_from_si(**si)
        self._h = self.lib.po_create(NX, NY, C.byref(self.units), POISSON[poisson], BC[bc], omega)
        if initialize:
            self.lib.po_initialize(self._h)

    def close(self):
        if self._h:
            self.lib.po_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _view(self, what: int, idx: int, n: int) -> np.ndarray:
        ptr = self.lib.po_field(self._h, what, idx)
        return np.ctypeslib.as_array(ptr, shape=(n,))

    def q_array(self, what: int, idx: int) -> np.ndarray:
        """AoS view [NY, NX, 9] of a population array (note: pointers move after swaps)."""
        return self._view(what, idx, self.NX * self.NY * 9).reshape(self.NY, self.NX, 9)

    def scalar(self, what: int, idx: int = 0) -> np.ndarray:
        return self._view(what, idx, self.NX * self.NY).reshape(self.NY, self.NX)

    def f(self, s): return self.q_array(PO_F, s)
    def g(self, s): return self.q_array(PO_G, s)

    def fields(self) -> dict:
        """The 15 visualised fields + phi, copied, keyed like FIELD_NAMES."""
        out = {}
        for s, sp in enumerate("ein"):
            out[f"ux_{sp}"] = self.scalar(PO_UX, s).copy()
            out[f"uy_{sp}"] = self.scalar(PO_UY, s).copy()
            out[f"T_{sp}"] = self.scalar(PO_T, s).copy()
            out[f"rho_{sp}"] = self.scalar(PO_RHO, s).copy()
        out["rho_q"] = self.scalar(PO_RHO_Q).copy()
        out["Ex"] = self.scalar(PO_EX).copy()
        out["Ey"] = self.scalar(PO_EY).copy()
        out["phi"] = self.scalar(PO_PHI).copy()
        return out

    def update_macro(self): self.lib.po_update_macro(self._h)
    def compute_equilibrium(self): self.lib.po_compute_equilibrium(self._h)
    def thermal_collisions(self): self.lib.po_thermal_collisions(self._h)
    def collisions(self): self.lib.po_collisions(self._h)
    def stream(self): self.lib.po_stream(self._h)
    def solve_poisson(self): self.lib.po_solve_poisson(self._h)
    def step(self, n: int = 1): self.lib.po_step(self._h, n)

    def run_with_dumps(self, nsteps: int, dump_steps) -> dict:
        """Advance nsteps; return {t: fields} as the reference's hook sees them at step t."""
        dump_steps = set(dump_steps)
        out = {}
        for t in range(nsteps):
            self.step(1)
            if t in dump_steps:
                out[t] = self.fields()
        return out


def ref_binary(kind: str = "parity") -> Path:
    return REF_DIR / f"ref_plasma_{kind}"


def have_reference(kind: str = "parity") -> bool:
    return ref_binary(kind).exists()


def run_reference(NX: int, NY: int, steps: int, poisson: str = "fft", bc: str = "periodic", threads: int = 0,
                  dump_steps=(), pops: bool = False, kind: str = "parity", phases: bool = False,
                  omega: float = 1.8, extra_args=(), timeout: float | None = None):
    """Run the compiled reference.  Returns (info_json, {t: {name: [NY,NX] array}}, pops or None)."""
    exe = ref_binary(kind)
    if not exe.exists():
        raise FileNotFoundError(f"{exe} missing: run `make -C oracle ref` where /root/reference is mounted")
    threads = threads or os.cpu_count() or 1
    with tempfile.TemporaryDirectory(prefix="plbm_ref_") as tmp:
        cmd = [str(exe), "--nx", str(NX), "--ny", str(NY), "--steps", str(steps), "--poisson", str(POISSON[poisson]),
               "--bc", str(BC[bc]), "--threads", str(threads), "--omega", repr(omega)]
        dump_steps = sorted(set(dump_steps))
        if dump_steps or pops:
            cmd += ["--out", tmp]
        if dump_steps:
            cmd += ["--dump", ",".join(str(t) for t in dump_steps)]
        if pops:
            cmd += ["--pops"]
        if phases:
            cmd += ["--phases"]
        cmd += list(extra_args)
        res = subprocess.run(cmd, check=True, capture_output=True, text=True, timeout=timeout)
        info = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
        dumps = {}
        for t in dump_steps:
            raw = np.fromfile(os.path.join(tmp, f"fields_t{t:05d}.f64"), dtype=np.float64)
            raw = raw.reshape(len(FIELD_NAMES), NY, NX)
            dumps[t] = {name: raw[k].copy() for k, name in enumerate(FIELD_NAMES)}
        pp = None
        if pops:
            raw = np.fromfile(os.path.join(tmp, "pops_final.f64"), dtype=np.float64).reshape(6, NY, NX, 9)
            pp = {"f": raw[0:3].copy(), "g": raw[3:6].copy()}
    return info, dumps, pp


def same_bits(a: np.ndarray, b: np.ndarray) -> bool:
    """Bit-identical up to the sign of zero (NaNs must match positionally)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return False
    eq = (a == b) | (np.isnan(a) & np.isnan(b))
    return bool(eq.all())


def max_norm_err(a: np.ndarray, ref: np.ndarray) -> float:
    """Field-normalised L-inf error max|a-ref| / max|ref|."""
    denom = float(np.max(np.abs(ref)))
    num = float(np.max(np.abs(np.asarray(a) - np.asarray(ref))))
    return num / denom if denom > 0 else num
