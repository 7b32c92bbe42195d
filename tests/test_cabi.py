"""The drop-in boundary: libplbm.so loads, exports every symbol include/plbm.h declares, and fails
loudly (no CPU fallback) when there is no CUDA device.  No compute calls here."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "plbm.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(plbm_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for must in ("plbm_create", "plbm_destroy", "plbm_step", "plbm_upload_state", "plbm_download_fields", "plbm_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(plbm):
    lib = plbm.load_library()
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"include/plbm.h declares symbols the library does not export: {missing}"


def test_config_struct_matches_header(plbm):
    """Field order/types of the ctypes mirror follow the C struct (spot check by size: 4 ints, 17 doubles, 6 ints)."""
    assert C.sizeof(plbm.PlbmConfig) == 4 * 4 + 8 * 17 + 6 * 4


def test_units_match_the_reference(plbm, oracle):
    """plbm_units_from_si (host code of the product) against the checker's restatement of
    reference include/plasma.hpp:86-133 and against the values the reference itself printed
    (tests/golden, generated from the unmodified sources)."""
    import numpy as np
    cfg = plbm.units_from_si()
    u = oracle.units_from_si()
    assert cfg.cs2 == u.cs2 and cfg.Kb == u.Kb and cfg.Ex_ext == u.Ex_ext and cfg.Ey_ext == u.Ey_ext
    for k in range(3):
        assert cfg.m[k] == u.m[k] and cfg.q[k] == u.q[k] and cfg.T_init[k] == u.T_init[k] and cfg.rho_init[k] == u.rho_init[k]
    z = np.load(ROOT / "tests" / "golden" / "n32_fft_periodic.npz")
    assert cfg.cs2 == float(z["units_cs2"]) and cfg.Kb == float(z["units_Kb"]) and cfg.m[1] == float(z["units_m_i"])
    assert cfg.rho_init[2] == float(z["units_rho_n_init"]) and cfg.Ex_ext == float(z["units_Ex_ext"])
    cfg2 = plbm.units_from_si(Z_ion=2, A_ion=4, T_e_SI=2e4, n_e_SI=3e11)
    u2 = oracle.units_from_si(Z_ion=2, A_ion=4, T_e_SI=2e4, n_e_SI=3e11)
    assert cfg2.cs2 == u2.cs2 and cfg2.m[1] == u2.m[1] and cfg2.q[1] == u2.q[1] and cfg2.rho_init[1] == u2.rho_init[1]


def test_no_cpu_fallback(plbm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(plbm.PlbmError, match="no CUDA device"):
        plbm.PlasmaLBM(16, 16)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the product package or include/ may reference it."""
    bad = []
    for p in list((ROOT / "12-lb-12-lb_b200").rglob("*")) + list((ROOT / "include").rglob("*")):
        if p.is_file() and p.suffix in {".py", ".cu", ".cuh", ".h", ".hpp", ".cpp"} and "build" not in p.parts:
            text = p.read_text(errors="ignore")
            for ln in text.splitlines():
                if re.search(r'^\s*(#\s*include|import|from)\b.*\boracle\b', ln):
                    bad.append(f"{p}: {ln.strip()}")
    assert not bad, bad


def test_slab_cut_invariants(plbm):
    """plbm_slab_of (the library owns the cut): slabs are even-sized, contiguous, cover NY, and at least two rows each, for the
    weighted default (rows of the reference's central block weigh more) on every lattice the benchmark and the tests
    cut; small lattices fall back to the equal split.  No GPU needed."""
    lib = plbm.load_library()
    lib.plbm_slab_of.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    def cut(NY, R):
        out = []
        for r in range(R):
            y0, n = C.c_int(), C.c_int()
            assert lib.plbm_slab_of(NY, r, R, C.byref(y0), C.byref(n)) == 0
            out.append((y0.value, n.value))
        return out
    for NY in (16, 20, 60, 64, 96, 200, 256, 1536, 2048, 3072, 4096, 6144, 8192, 12288):
        for R in (2, 3, 4, 8):
            if NY // 2 < R:
                continue
            c = cut(NY, R)
            assert c[0][0] == 0 and sum(n for _, n in c) == NY
            assert all(n >= 2 and n % 2 == 0 for _, n in c), (NY, R, c)
            assert all(c[i][0] + c[i][1] == c[i + 1][0] for i in range(R - 1))
    assert [n for _, n in cut(8192, 2)] == [4096, 4096]
    rows = [n for _, n in cut(8192, 8)]
    assert rows[0] > rows[3] and rows[0] == rows[7]                                            # outer slabs hold more rows
