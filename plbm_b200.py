"""Importable alias of the product package directory ``12-lb-12-lb_b200`` (not a Python identifier)."""
import importlib
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
_pkg = importlib.import_module("12-lb-12-lb_b200")
globals().update({k: getattr(_pkg, k) for k in _pkg.__all__})
__all__ = list(_pkg.__all__)
