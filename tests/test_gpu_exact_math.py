"""The fast division primitives of csrc/exact_math.cuh against IEEE division (numpy), on millions of
operands: wherever the primitive reports its operands in-domain the quotient must be the correctly
rounded one, bit for bit.  This pins the exactness claims of DESIGN.md section 3."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run(plbm, a, b, mode):
    lib = plbm.load_library()
    dp = C.POINTER(C.c_double)
    lib.plbm_selftest_division.argtypes = [dp, dp, C.c_int, C.c_int, dp, C.POINTER(C.c_int)]
    a = np.ascontiguousarray(a, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
    out = np.empty_like(a); ok = np.empty(a.size, dtype=np.int32)
    rc = lib.plbm_selftest_division(a.ctypes.data_as(dp), b.ctypes.data_as(dp), a.size, mode, out.ctypes.data_as(dp),
                                    ok.ctypes.data_as(C.POINTER(C.c_int)))
    assert rc == 0, lib.plbm_last_error().decode()
    return out, ok.astype(bool)


def operands(n, seed):
    """Mixed magnitudes, random mantissas, exact powers of two, values next to powers of two, zeros, signs."""
    rng = np.random.default_rng(seed)
    mant = rng.integers(0, 1 << 52, size=n, dtype=np.uint64)
    expo = rng.integers(1023 - 300, 1023 + 300, size=n, dtype=np.uint64)
    sign = rng.integers(0, 2, size=n, dtype=np.uint64) << np.uint64(63)
    v = (sign | (expo << np.uint64(52)) | mant).view(np.float64)
    k = n // 16
    v[:k] = np.ldexp(1.0, rng.integers(-200, 200, size=k))
    v[k:2 * k] = np.nextafter(np.ldexp(1.0, rng.integers(-200, 200, size=k)), 0.0)
    v[2 * k:3 * k] = 0.0
    v[3 * k:4 * k] = rng.uniform(-2, 2, size=k)
    return v


def same(got, want):
    return (got == want) | (np.isnan(got) & np.isnan(want))


@pytest.mark.parametrize("tau", [3.0, 5.0, 6.0])
def test_division_by_relaxation_times(plbm, tau):
    a = operands(4_000_000, int(tau))
    with np.errstate(all="ignore"):
        want = a / tau
    for mode in (1, 2):
        got, ok = run(plbm, a, np.array([tau]), mode)
        assert ok.mean() > 0.9
        bad = ok & ~same(got, want)
        assert not bad.any(), f"tau={tau} mode={mode}: {int(bad.sum())} wrong quotients, first a={a[bad][0]!r}"


def test_division_by_lattice_constants(plbm, oracle):
    u = oracle.units_from_si()
    for d in (u.cs2, u.Kb, u.m[1], 1.0, 4.0 * u.m[1]):
        a = operands(2_000_000, 17)
        with np.errstate(all="ignore"):
            want = a / d
        got, ok = run(plbm, a, np.array([d]), 2)
        bad = ok & ~same(got, want)
        assert not bad.any(), f"d={d!r}: {int(bad.sum())} wrong quotients, first a={a[bad][0]!r}"


def test_data_dependent_division(plbm):
    a = operands(4_000_000, 5)
    b = operands(4_000_000, 6)
    b[b == 0] = 1.5
    with np.errstate(all="ignore"):
        want = a / b
    got, ok = run(plbm, a, b, 0)
    assert ok.mean() > 0.5
    bad = ok & ~same(got, want)
    assert not bad.any(), f"{int(bad.sum())} wrong quotients, first {a[bad][0]!r} / {b[bad][0]!r}"


def test_out_of_domain_operands_are_flagged(plbm):
    """Tiny, subnormal, huge and non-finite operands must never be reported in-domain with a wrong result."""
    a = np.array([1e-300, 5e-324, 2.5e-310, 1e300, np.inf, -np.inf, np.nan, 1e-170, 1.0, 0.0, -0.0, 3.0])
    b = np.array([3.0, 3.0, 7.0, 1e-300, 2.0, 2.0, 2.0, 1e150, 0.0, 5.0, 5.0, np.inf])
    with np.errstate(all="ignore"):
        want = a / b
    got, ok = run(plbm, a, b, 0)
    bad = ok & ~same(got, want)
    assert not bad.any(), (a[bad], b[bad], got[bad], want[bad])
    assert not ok[:9].any()          # all of those are outside the accepted domain
    assert ok[9] and ok[10]          # exact zeros are accepted
