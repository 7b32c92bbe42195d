"""oracle/fft_oracle.c (the stand-in for the absent FFTW3) against numpy.fft, plus its radix rule.
Tolerance: 1e-13 field-normalised (both are double-precision FFTs; they need not agree bitwise)."""
import ctypes as C

import numpy as np
import pytest


@pytest.mark.parametrize("n0,n1", [(8, 8), (200, 200), (12, 10), (7, 9), (64, 64), (71, 71), (15, 20), (20, 15),
                                   (5, 4), (141, 141), (3, 3), (2, 2), (1, 6), (6, 1), (100, 36)])
def test_r2c_c2r_against_numpy(oracle, n0, n1):
    lib = oracle.port_lib()
    rng = np.random.default_rng(n0 * 1000 + n1)
    a = rng.standard_normal((n0, n1))
    out = np.zeros((n0, n1 // 2 + 1), dtype=np.complex128)
    plan = lib.offt_plan2d_create(n0, n1)
    try:
        lib.offt_r2c_2d(plan, a.ctypes.data, out.ctypes.data)
        ref = np.fft.rfft2(a)
        assert np.abs(out - ref).max() <= 1e-13 * max(np.abs(ref).max(), 1.0)
        spec = ref.copy()
        back = np.zeros((n0, n1))
        lib.offt_c2r_2d(plan, spec.ctypes.data, back.ctypes.data)
        assert np.abs(back / (n0 * n1) - a).max() <= 1e-13 * np.abs(a).max()
    finally:
        lib.offt_plan2d_destroy(plan)


def test_c2r_ignores_imaginary_part_of_self_conjugate_bins(oracle):
    """FFTW's c2r contract: Im of the DC and Nyquist bins along the half axis is not used."""
    lib = oracle.port_lib()
    n0, n1 = 6, 8
    rng = np.random.default_rng(5)
    spec = np.fft.rfft2(rng.standard_normal((n0, n1)))
    dirty = spec.copy()
    dirty[0, 0] += 3.0j
    dirty[0, n1 // 2] -= 2.0j
    a = np.zeros((n0, n1)); b = np.zeros((n0, n1))
    plan = lib.offt_plan2d_create(n0, n1)
    s1 = spec.copy(); s2 = dirty.copy()
    lib.offt_c2r_2d(plan, s1.ctypes.data, a.ctypes.data)
    lib.offt_c2r_2d(plan, s2.ctypes.data, b.ctypes.data)
    lib.offt_plan2d_destroy(plan)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("n,expect", [(1, []), (2, [2]), (8, [4, 2]), (200, [4, 2, 5, 5]), (2048, [4] * 5 + [2]),
                                      (8192, [4] * 6 + [2]), (5760, [4, 4, 4, 2, 3, 3, 5]), (71, [71]), (141, [3, 47])])
def test_radix_schedule(oracle, n, expect):
    lib = oracle.port_lib()
    buf = (C.c_int * 64)()
    k = lib.offt_factorize(n, buf)
    assert list(buf[:k]) == expect
