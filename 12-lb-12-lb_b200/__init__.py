"""12-lb-12-lb_b200 -- B200-native plasma lattice-Boltzmann step (one hot path of AMSC-24-25/12-lb-12-lb).

The product is the CUDA library ``libplbm.so`` behind the C ABI of ``include/plbm.h`` plus the C++
headers in ``include/`` that keep the reference's public surface.  This Python package is only
the thin ctypes binding the tests and ``bench.py`` use; it holds no arithmetic and it never
imports ``oracle/``.  (The directory name is not a Python identifier: import it with
``importlib.import_module("12-lb-12-lb_b200")`` or through the top-level alias ``plbm_b200``.)
"""
from .api import (PlasmaLBM, PlbmConfig, PlbmError, FIELD_NAMES, POISSON, BC, DEFAULT_SI,
                  build_library, load_library, library_path, units_from_si)

from .distributed import CudaSlabBackend, SlabDriver, slab_of  # noqa: E402

__all__ = ["CudaSlabBackend", "SlabDriver", "slab_of", "PlasmaLBM", "PlbmConfig", "PlbmError", "FIELD_NAMES", "POISSON", "BC", "DEFAULT_SI",
           "build_library", "load_library", "library_path", "units_from_si"]
