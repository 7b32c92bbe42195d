"""Shared comparison helpers for the parity tests (test infrastructure)."""
import numpy as np

SPECIES = "ein"


def assert_same_bits(got, ref, what):
    """Bit-identical up to the sign of zero; NaNs must coincide."""
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    eq = (got == ref) | (np.isnan(got) & np.isnan(ref))
    if not eq.all():
        bad = np.argwhere(~eq)
        k = tuple(bad[0])
        denom = np.max(np.abs(ref)) or 1.0
        raise AssertionError(f"{what}: {len(bad)} of {ref.size} values differ; first at {k}: got {got[k]!r} want {ref[k]!r}; "
                             f"max|diff|/max|ref| = {np.nanmax(np.abs(got - ref)) / denom:.3e}")


def assert_fields_same(got: dict, ref: dict, label: str, names=None):
    for name in (names or ref.keys()):
        assert_same_bits(got[name], ref[name], f"{label}:{name}")


def norm_err(got, ref):
    denom = float(np.max(np.abs(ref)))
    num = float(np.max(np.abs(np.asarray(got) - np.asarray(ref))))
    return num / denom if denom > 0 else num
