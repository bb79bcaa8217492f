"""cProfile of the plugin's per-call host work at ~1000 rays with live Optiland objects on the GPU (where does the
time of a small Optic.trace / SurfaceGroup.trace go?).  python scripts/profile_small_trace.py [optic|group|changed]"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle.ref_import import import_reference  # noqa: E402

import_reference()
import optiland.backend as be  # noqa: E402
from optiland.rays import RealRays  # noqa: E402
from optiland.samples.objectives import DoubleGauss  # noqa: E402

from optiland_b200 import plugin as P  # noqa: E402

be.set_backend("torch")
be.set_device("cuda")
be.set_precision("float64")
be.grad_mode.disable()
P.install()
lens = DoubleGauss()
what = sys.argv[1] if len(sys.argv) > 1 else "optic"
r0 = lens.trace(0.0, 0.7, 0.5876, 18, "hexapolar")
gen = lens.ray_tracer.ray_generator.generate_rays(0.0, 0.7, P._distribution(be, "hexapolar", 18).x, P._distribution(be, "hexapolar", 18).y, 0.5876)
keep = {k: getattr(gen, k).clone() for k in ("x", "y", "z", "L", "M", "N", "i", "w")}
k = [0]
r3 = float(lens.surfaces.surfaces[3].geometry.radius)


def step():
    if what == "optic":
        lens.trace(0.0, 0.7, 0.5876, 18, "hexapolar")
    elif what == "group":
        lens.surfaces.trace(RealRays(*[keep[q] for q in ("x", "y", "z", "L", "M", "N", "i", "w")]))
    else:
        k[0] += 1
        lens.surfaces.surfaces[3].geometry.radius = be.array(r3 * (1 + 1e-9 * k[0]))
        lens.trace(0.0, 0.7, 0.5876, 18, "hexapolar")


for _ in range(20):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    step()
torch.cuda.synchronize()
print(f"{what}: {(time.perf_counter() - t0) / 200 * 1e3:.3f} ms per call")
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
