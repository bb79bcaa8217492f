"""optiland_b200 -- B200-native batched real-ray trace behind Optiland's backend registry.

Only the hot path lives here: ``SurfaceGroup.trace`` and everything it
calls per surface, as hand-written sm_100a CUDA reached through the C ABI in
``include/olb.h``.  Importing the package does not touch the GPU; the CUDA library is
loaded on first use and its absence is a hard error (there is no CPU fallback).
"""
from . import table  # noqa: F401
from .table import SurfaceSpec, SurfaceTable  # noqa: F401

__all__ = ["table", "SurfaceSpec", "SurfaceTable"]
__version__ = "0.1.0"
