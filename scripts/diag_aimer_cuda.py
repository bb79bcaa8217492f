"""Diagnostic (GPU box): why do the iterative aimer's subset traces not reach the CUDA engine?  Prints plugin.stats()
and the device / dtype / shape of every ray array the aimer hands to Surface.trace."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import import_reference  # noqa: E402

import_reference()
import optiland.backend as be  # noqa: E402
from optiland.samples.objectives import CookeTriplet  # noqa: E402

from optiland_b200 import plugin as P  # noqa: E402

be.set_backend("torch")
be.set_device("cuda")
be.set_precision("float64")
be.grad_mode.disable()
eng = P.CudaEngine()
P.install(engine=eng)
from optiland.rays.ray_aiming.iterative import IterativeRayAimer  # noqa: E402

wrapped = IterativeRayAimer._trace_subset


def spy(self, x, y, z, L, M, N, wl, stop, is_inf):
    for name, v in (("x", x), ("y", y), ("z", z), ("L", L), ("M", M), ("N", N), ("wl", wl)):
        print("  subset arg", name, type(v).__name__, getattr(v, "device", None), getattr(v, "dtype", None), getattr(v, "shape", None),
              getattr(v, "requires_grad", None))
    n0 = len(eng.calls)
    out = wrapped(self, x, y, z, L, M, N, wl, stop, is_inf)
    print("  -> engine calls:", eng.calls[n0:], "stats:", P.stats())
    return out


IterativeRayAimer._trace_subset = spy
for fused in (True, False):
    P._state["fuse_aimer"] = fused
    P.stats(reset=True)
    lens = CookeTriplet()
    lens.set_ray_aiming("iterative", max_iter=10, tol=1e-9)
    print("fused", fused)
    rays = lens.trace(0.0, 0.7, 0.55, 4, "hexapolar")
    print("final stats", P.stats(), "calls", [c[:2] for c in eng.calls][-12:])
