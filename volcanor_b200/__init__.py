"""volcanor_b200 -- B200-native (sm_100a, FP64) Biot-Savart hot path of VOLCANOR.

The product is the C-ABI shared library ``libvolcanor_b200.so`` (include/volcanor_b200.h);
this package is the thin host-side mirror used by tests and bench.py.  There is no CPU path:
importing works anywhere (so the build can be checked without a GPU), but creating a
:class:`Context` without a CUDA device raises.
"""
from .api import Context, VlcError, lib_path, load_library, build_library, DECLARED_SYMBOLS  # noqa: F401

__all__ = ["Context", "VlcError", "lib_path", "load_library", "build_library", "DECLARED_SYMBOLS"]
