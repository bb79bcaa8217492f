"""Secondary measurements (not the headline bench): the other BASELINE.json configurations at their
full sizes on one GPU -- forward fp32/fp64 with full records, polarized forward, and the
forward+backward step of the autograd configuration.  Prints one JSON line per configuration."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optiland_b200 import autograd as AG  # noqa: E402
from optiland_b200.trace import DeviceTable, PolarizedRays, RealRays, trace_device  # noqa: E402
from tests._util import Case  # noqa: E402


def timeit(fn, k=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / k


def resample(c, n, dtype, cls=RealRays):
    dev = torch.device("cuda:0")
    idx = torch.randint(0, c.n, (n,), device=dev, generator=torch.Generator(device=dev).manual_seed(0))
    rr = {k: torch.from_numpy(v).to(dev)[idx] for k, v in c.rays.items()}
    return cls(rr["x"], rr["y"], rr["z"], rr["L"], rr["M"], rr["N"], rr["i"], rr["w"], dtype=dtype, device=dev)


def forward_case(name, n, label, cls=RealRays):
    c = Case(name)
    dtab = DeviceTable(c.table, "cuda:0")
    S = c.table.num_surfaces
    out = {"config": label, "system": name, "rays": n, "surfaces": S}
    for dtype, tag, es in ((torch.float32, "f32", 4), (torch.float64, "f64", 8)):
        base = resample(c, n, dtype, cls)

        def step():
            rr = cls.__new__(cls)
            rr.__dict__.update(base.__dict__)
            if cls is PolarizedRays:
                rr.p = None
            trace_device(dtab, rr, 0, S, record=True)

        ms = timeit(step)
        bytes_ = es * ((8 if c.table.n_wl == 1 else 9) + 8 * S) * n + (18 * es * n if cls is PolarizedRays else 0)
        out[tag] = {"ms": round(ms, 4), "ray_surfaces_per_s": round(n * (S - 1) / ms * 1e3, 0),
                    "algorithmic_GBps": round(bytes_ / ms / 1e6, 1)}
        del base
        torch.cuda.empty_cache()
    print(json.dumps(out), flush=True)


def autograd_case(n):
    c = Case("telephoto_c3_tol1e-6")
    S = c.table.num_surfaces
    out = {"config": "C3 reverse telephoto + 2 even aspheres: forward + backward (d RMS spot / d all parameters)",
           "rays": n, "surfaces": S}
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        base = resample(c, n, dtype)
        params = AG.table_to_params(c.table).requires_grad_(True)

        def step():
            rr = RealRays.__new__(RealRays)
            rr.__dict__.update(base.__dict__)
            rec = AG.trace_differentiable(c.table, params, rr, rows=(-1,))
            x, y = rec["x"], rec["y"]
            loss = torch.sqrt(torch.mean((x - x.mean()) ** 2 + (y - y.mean()) ** 2))
            params.grad = None
            loss.backward()

        ms = timeit(step, k=10)
        out[tag] = {"ms_fwd_bwd": round(ms, 3), "ray_surfaces_per_s": round(n * (S - 1) / ms * 1e3, 0)}
        del base
        torch.cuda.empty_cache()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    forward_case("dgauss_c2", 10_000_000, "C2 Double-Gauss 10M rays (resampled fixture rays)")
    forward_case("telephoto_c3_tol1e-6", 4_000_000, "C3 reverse telephoto + 2 even aspheres, 4M rays")
    forward_case("hubble_c4", 16_000_000, "C4 Hubble (conic mirrors + obscuration), 16M rays")
    forward_case("zernike_fringe", 4_000_000, "C5 geometry: Zernike freeform singlet, 4M rays")
    forward_case("zernike_polarized_c5", 4_000_000, "C5: Zernike + Fresnel coatings + polarized, 3 wavelengths, 4M rays/GPU", PolarizedRays)
    autograd_case(4_000_000)
