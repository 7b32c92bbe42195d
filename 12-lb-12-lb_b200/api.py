"""ctypes binding of include/plbm.h.  No arithmetic lives here; a missing library or a missing
GPU raises PlbmError -- there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent

# order of visualize::UpdateVisualization's parameters (reference include/visualize.hpp:53-61) + phi
NUM_FRAMES, NUM_SERIES, NUM_POINTS = 12, 19, 9          # plbm.h: PLBM_NUM_FRAMES / SERIES / POINTS
FRAME_NAMES = ("rho_e", "rho_i", "rho_q", "ux_e", "uy_e", "ue_mag", "ux_i", "uy_i", "ui_mag", "T_e", "T_i", "T_n")
SERIES_NAMES = ("ux_e", "uy_e", "ue_mag", "ux_i", "uy_i", "ui_mag", "ux_n", "uy_n", "un_mag", "T_e", "T_i", "T_n",
                "rho_e", "rho_i", "rho_n", "rho_q", "Ex", "Ey", "E_mag")
FIELD_NAMES = ("ux_e", "uy_e", "ux_i", "uy_i", "ux_n", "uy_n", "T_e", "T_i", "T_n",
               "rho_e", "rho_i", "rho_n", "rho_q", "Ex", "Ey", "phi")
POISSON = {"none": 0, "gs": 1, "sor": 2, "fft": 3, "nps": 4}   # poisson::PoissonType, reference include/poisson.hpp:15-21
BC = {"periodic": 0, "bounceback": 1}                          # streaming::BCType, reference include/streaming.hpp:10-13
# reference src/main_plasma.cpp:16-51
DEFAULT_SI = dict(Z_ion=1, A_ion=1, Ex_SI=1e-2, Ey_SI=0.0, T_e_SI=1e4, T_i_SI=300.0, T_n_SI=300.0,
                  n_e_SI=1e11, n_n_SI=1e18)


class PlbmError(RuntimeError):
    pass


class PlbmConfig(C.Structure):
    _fields_ = [("NX", C.c_int), ("NY", C.c_int), ("poisson_type", C.c_int), ("bc_type", C.c_int),
                ("omega_sor", C.c_double), ("cs2", C.c_double), ("Kb", C.c_double),
                ("Ex_ext", C.c_double), ("Ey_ext", C.c_double),
                ("T_init", C.c_double * 3), ("m", C.c_double * 3), ("q", C.c_double * 3), ("rho_init", C.c_double * 3),
                ("rank", C.c_int), ("nranks", C.c_int), ("y0", C.c_int), ("NY_local", C.c_int), ("device", C.c_int),
                ("fields_only", C.c_int)]


def library_path() -> Path:
    """In-tree library; PLBM_LIBRARY selects a differently tuned build of the same sources (benchmark sweeps)."""
    import os
    return Path(os.environ["PLBM_LIBRARY"]) if os.environ.get("PLBM_LIBRARY") else PKG_DIR / "libplbm.so"


def build_library(force: bool = False) -> Path:
    """Compile libplbm.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.run(["make", "-s", "-C", str(PKG_DIR), "clean"], check=True)
    subprocess.run(["make", "-s", "-j8", "-C", str(PKG_DIR)], check=True)
    return library_path()


_lib = None


def load_library() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not path.exists():
        raise PlbmError(f"{path} is missing: run `make -C {PKG_DIR}` (or __graft_entry__.build()); there is no fallback path")
    lib = C.CDLL(str(path))
    dp = C.POINTER(C.c_double)
    lib.plbm_last_error.restype = C.c_char_p
    lib.plbm_units_from_si.argtypes = [C.c_int, C.c_int] + [C.c_double] * 7 + [C.POINTER(PlbmConfig)]
    lib.plbm_create.argtypes = [C.POINTER(PlbmConfig), C.POINTER(C.c_void_p)]
    lib.plbm_destroy.argtypes = [C.c_void_p]
    lib.plbm_destroy.restype = None
    lib.plbm_initialize.argtypes = [C.c_void_p]
    lib.plbm_upload_state.argtypes = [C.c_void_p, C.POINTER(dp), C.POINTER(dp)]
    lib.plbm_download_state.argtypes = [C.c_void_p, C.POINTER(dp), C.POINTER(dp)]
    lib.plbm_set_efield.argtypes = [C.c_void_p, dp, dp]
    lib.plbm_step.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.plbm_sync.argtypes = [C.c_void_p]
    lib.plbm_download_fields.argtypes = [C.c_void_p, C.POINTER(dp)]
    lib.plbm_fetch_begin.argtypes = [C.c_void_p, C.POINTER(dp)]
    lib.plbm_fetch_wait.argtypes = [C.c_void_p]
    lib.plbm_frames_begin.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_float)), dp]
    lib.plbm_pin_host.argtypes = [C.c_void_p, C.c_size_t]
    lib.plbm_unpin_host.argtypes = [C.c_void_p]
    lib.plbm_host_solve_poisson.argtypes = [C.c_void_p, dp, dp, dp]
    lib.plbm_host_poisson_config.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
    lib.plbm_host_poisson_solver.argtypes = [C.c_void_p, C.c_int, dp]
    lib.plbm_step_timed.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                    C.POINTER(C.c_float), C.POINTER(C.c_longlong)]
    lib.plbm_local_rows.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.plbm_device_bytes.argtypes = [C.c_void_p]
    lib.plbm_device_bytes.restype = C.c_longlong
    lib.plbm_stream.argtypes = [C.c_void_p]
    lib.plbm_stream.restype = C.c_void_p
    _lib = lib
    return lib


def _check(lib, rc: int, what: str) -> None:
    if rc != 0:
        raise PlbmError(f"{what}: {lib.plbm_last_error().decode()}")


def units_from_si(cfg: PlbmConfig | None = None, **si) -> PlbmConfig:
    lib = load_library()
    cfg = cfg if cfg is not None else PlbmConfig()
    p = dict(DEFAULT_SI)
    p.update(si)
    _check(lib, lib.plbm_units_from_si(p["Z_ion"], p["A_ion"], p["Ex_SI"], p["Ey_SI"], p["T_e_SI"], p["T_i_SI"],
                                       p["T_n_SI"], p["n_e_SI"], p["n_n_SI"], C.byref(cfg)), "plbm_units_from_si")
    return cfg


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class PlasmaLBM:
    """One simulation context on one GPU (optionally one y-slab of a larger lattice)."""

    def __init__(self, NX: int, NY: int, poisson: str = "fft", bc: str = "periodic", omega: float = 1.8,
                 rank: int = 0, nranks: int = 1, y0: int = 0, NY_local: int | None = None, device: int = -1,
                 initialize: bool = True, fields_only: bool = False, **si):
        self.lib = load_library()
        cfg = PlbmConfig()
        cfg.fields_only = 1 if fields_only else 0
        cfg.NX, cfg.NY = NX, NY
        cfg.poisson_type, cfg.bc_type, cfg.omega_sor = POISSON[poisson], BC[bc], omega
        cfg.rank, cfg.nranks, cfg.y0 = rank, nranks, y0
        cfg.NY_local = NY if NY_local is None else NY_local
        cfg.device = device
        units_from_si(cfg, **si)
        self.cfg = cfg
        self._h = C.c_void_p()
        _check(self.lib, self.lib.plbm_create(C.byref(cfg), C.byref(self._h)), "plbm_create")
        self.NX, self.NY = NX, NY
        y0_, nyl = C.c_int(), C.c_int()
        self.lib.plbm_local_rows(self._h, C.byref(y0_), C.byref(nyl))
        self.y0, self.NY_local = y0_.value, nyl.value
        if initialize and not fields_only:
            self.initialize()

    def close(self):
        if getattr(self, "_h", None):
            self.lib.plbm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def initialize(self):
        _check(self.lib, self.lib.plbm_initialize(self._h), "plbm_initialize")

    def step(self, nsteps: int = 1, want_fields: bool = False):
        _check(self.lib, self.lib.plbm_step(self._h, nsteps, int(want_fields)), "plbm_step")

    def sync(self):
        _check(self.lib, self.lib.plbm_sync(self._h), "plbm_sync")

    def step_timed(self, nsteps: int, want_fields: bool = False) -> dict:
        tot, k1, ps, nl = C.c_float(), C.c_float(), C.c_float(), C.c_longlong()
        _check(self.lib, self.lib.plbm_step_timed(self._h, nsteps, int(want_fields), C.byref(tot), C.byref(k1),
                                                  C.byref(ps), C.byref(nl)), "plbm_step_timed")
        return {"ms_total": tot.value, "ms_k1": k1.value, "ms_poisson": ps.value, "launches": nl.value}

    def upload_state(self, f, g):
        """f, g: arrays [3, NY_local, NX, 9] (AoS, populations at the top of the time loop)."""
        f = [np.ascontiguousarray(f[s], dtype=np.float64) for s in range(3)]
        g = [np.ascontiguousarray(g[s], dtype=np.float64) for s in range(3)]
        shape = (self.NY_local, self.NX, 9)
        for a in f + g:
            if a.shape != shape:
                raise ValueError(f"population array has shape {a.shape}, expected {shape}")
        dp = C.POINTER(C.c_double)
        fa = (dp * 3)(*[_dptr(a) for a in f])
        ga = (dp * 3)(*[_dptr(a) for a in g])
        _check(self.lib, self.lib.plbm_upload_state(self._h, fa, ga), "plbm_upload_state")

    def download_state(self):
        f = np.empty((3, self.NY_local, self.NX, 9), dtype=np.float64)
        g = np.empty_like(f)
        dp = C.POINTER(C.c_double)
        fa = (dp * 3)(*[_dptr(f[s]) for s in range(3)])
        ga = (dp * 3)(*[_dptr(g[s]) for s in range(3)])
        _check(self.lib, self.lib.plbm_download_state(self._h, fa, ga), "plbm_download_state")
        return f, g

    def set_efield(self, Ex, Ey):
        Ex = np.ascontiguousarray(Ex, dtype=np.float64)
        Ey = np.ascontiguousarray(Ey, dtype=np.float64)
        _check(self.lib, self.lib.plbm_set_efield(self._h, _dptr(Ex), _dptr(Ey)), "plbm_set_efield")

    def fields(self, names=FIELD_NAMES) -> dict:
        """Download the named fields ([NY_local, NX] each).  Moment fields need step(..., want_fields=True)."""
        dp = C.POINTER(C.c_double)
        out = {n: np.empty((self.NY_local, self.NX), dtype=np.float64) for n in names}
        ptrs = (dp * len(FIELD_NAMES))()
        for k, n in enumerate(FIELD_NAMES):
            ptrs[k] = _dptr(out[n]) if n in out else dp()
        _check(self.lib, self.lib.plbm_download_fields(self._h, ptrs), "plbm_download_fields")
        return out

    def solve_poisson(self, rho_q, Ex=None, Ey=None):
        """poisson::SolvePoisson on host arrays (plbm_host_solve_poisson): returns (Ex, Ey); the potential stays in the
        context (fields(["phi"]))."""
        rho_q = np.ascontiguousarray(rho_q, dtype=np.float64)
        Ex = np.zeros_like(rho_q) if Ex is None else np.ascontiguousarray(Ex, dtype=np.float64).copy()
        Ey = np.zeros_like(rho_q) if Ey is None else np.ascontiguousarray(Ey, dtype=np.float64).copy()
        _check(self.lib, self.lib.plbm_host_solve_poisson(self._h, _dptr(rho_q), _dptr(Ex), _dptr(Ey)), "plbm_host_solve_poisson")
        return Ex, Ey

    # poisson::SolvePoisson_{GS,SOR,FFT,9point} and their *_Periodic variants on host arrays (plbm_host_poisson_solver)
    SOLVER_CODES = {"gs": 1, "sor": 2, "fft": 3, "nps": 4, "gs_periodic": 5, "sor_periodic": 6, "nps_periodic": 7}

    def poisson_solver(self, solver: str, rho_q, omega: float = 1.8, bc: str = "periodic"):
        """One call of the named solver on a fields-only context (rho_q -> phi, warm-started; the potential stays in the
        context: fields(["phi"])).  Type, boundary and omega are per call, as in the reference."""
        code = self.SOLVER_CODES[solver]
        rho_q = np.ascontiguousarray(rho_q, dtype=np.float64)
        base = {5: 1, 6: 2, 7: 4}.get(code, code)          # the PoissonType the reference's caller would pass
        _check(self.lib, self.lib.plbm_host_poisson_config(self._h, base, BC[bc], omega), "plbm_host_poisson_config")
        _check(self.lib, self.lib.plbm_host_poisson_solver(self._h, code, _dptr(rho_q)), "plbm_host_poisson_solver")

    def fetch_begin(self, out: np.ndarray, nfields: int = len(FIELD_NAMES)):
        """Start copying the last step's first `nfields` fields into out[nfields, NY_local, NX] (ideally pinned memory)
        and return at once; the next step() may be issued right away.  Complete it with fetch_wait()."""
        if out.dtype != np.float64 or not out.flags["C_CONTIGUOUS"] or out.shape != (nfields, self.NY_local, self.NX):
            raise ValueError("fetch_begin: out must be a C-contiguous float64 array [nfields, NY_local, NX]")
        dp = C.POINTER(C.c_double)
        ptrs = (dp * len(FIELD_NAMES))()
        for k in range(len(FIELD_NAMES)):
            ptrs[k] = _dptr(out[k]) if k < nfields else dp()
        self._fetch_target = out                     # keep the destination alive until fetch_wait
        _check(self.lib, self.lib.plbm_fetch_begin(self._h, ptrs), "plbm_fetch_begin")

    def frames_begin(self, frames: np.ndarray | None, series: np.ndarray | None):
        """Alternate output path (plbm_frames_begin): frames[12, NY_local, NX] float32 and/or series[19, 9] float64
        receive what the visualiser derives from the last step's fields; complete with fetch_wait()."""
        fp = C.POINTER(C.c_float)
        ptrs = None
        if frames is not None:
            if frames.dtype != np.float32 or not frames.flags["C_CONTIGUOUS"] or frames.shape != (NUM_FRAMES, self.NY_local, self.NX):
                raise ValueError("frames_begin: frames must be a C-contiguous float32 array [12, NY_local, NX]")
            ptrs = (fp * NUM_FRAMES)(*[frames[k].ctypes.data_as(fp) for k in range(NUM_FRAMES)])
        sp = None
        if series is not None:
            if series.dtype != np.float64 or not series.flags["C_CONTIGUOUS"] or series.shape != (NUM_SERIES, NUM_POINTS):
                raise ValueError("frames_begin: series must be a C-contiguous float64 array [19, 9]")
            sp = _dptr(series)
        self._fetch_target = (frames, series)
        _check(self.lib, self.lib.plbm_frames_begin(self._h, ptrs, sp), "plbm_frames_begin")

    def fetch_wait(self):
        _check(self.lib, self.lib.plbm_fetch_wait(self._h), "plbm_fetch_wait")
        out, self._fetch_target = getattr(self, "_fetch_target", None), None
        return out

    @property
    def device_bytes(self) -> int:
        return int(self.lib.plbm_device_bytes(self._h))

    @property
    def stream(self) -> int:
        return int(self.lib.plbm_stream(self._h) or 0)
