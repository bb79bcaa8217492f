"""Per-call host overhead of the plugin path at small ray counts (where it dominates): packing live objects
whose parameters sit on the GPU, preparing + uploading a changed table, and a cached trace."""
import json
import os
import sys
import time
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optiland_b200 import pack as PK  # noqa: E402
from optiland_b200.plugin import CudaEngine  # noqa: E402
from tests._fake_optiland import fake_surfaces  # noqa: E402
from tests._util import Case  # noqa: E402


def timeit(fn, n=200, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def main():
    c = Case("dgauss_c2")
    group = types.SimpleNamespace(surfaces=fake_surfaces(c.table, "cuda"))
    wl = c.table.wavelengths
    out = {"system": "dgauss_c2 (13 surfaces), parameters resident on cuda:0"}
    out["pack_ms_one_copy"] = timeit(lambda: PK.pack_surface_group(group, wl))
    enter = PK._Prefetch.__enter__
    PK._Prefetch.__enter__ = lambda self: self
    out["pack_ms_scalar_by_scalar"] = timeit(lambda: PK.pack_surface_group(group, wl))
    PK._Prefetch.__enter__ = enter
    eng = CudaEngine()
    n = 1000
    rays0 = {k: torch.from_numpy(c.rays[k][:n]).cuda() for k in ("x", "y", "z", "L", "M", "N", "i", "w")}

    def mk():
        r = types.SimpleNamespace(**rays0)
        r.opd = torch.zeros_like(r.x)
        return r

    tab = PK.pack_surface_group(group, wl)
    out["trace_ms_cached_table_1000_rays"] = timeit(lambda: eng.trace(tab, mk(), 0, tab.num_surfaces))
    k = [0]

    def changed():
        k[0] += 1
        group.surfaces[3].geometry.radius = torch.tensor(c.table.surfaces[3].radius * (1 + 1e-9 * k[0]), dtype=torch.float64, device="cuda")
        t = PK.pack_surface_group(group, wl)
        eng.trace(t, mk(), 0, t.num_surfaces)

    out["pack_upload_trace_ms_changed_table_1000_rays"] = timeit(changed)

    # differentiable step as the plugin runs it for the reference's torch optimiser: pack live objects,
    # parameters as autograd inputs, forward + adjoint kernels, d(loss)/d(radius of surface 3)
    from optiland_b200 import plugin as P

    leaf = torch.tensor(c.table.surfaces[3].radius, dtype=torch.float64, device="cuda", requires_grad=True)
    stages = {"pack": 0.0, "live_params": 0.0, "forward": 0.0, "loss_backward": 0.0}

    def grad_step(acc=True):
        group.surfaces[3].geometry.radius = leaf * 1.0
        t0 = time.perf_counter()
        t = PK.pack_surface_group(group, wl)
        t1 = time.perf_counter()
        params = P._live_params(group.surfaces, t, float(wl[0]))
        t2 = time.perf_counter()
        rec = eng.trace_grad(t, params, mk())
        t3 = time.perf_counter()
        x, y = rec["x"][-1], rec["y"][-1]
        loss = (x * x + y * y).mean()
        leaf.grad = None
        loss.backward()
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        if acc:
            for k, v in zip(stages, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
                stages[k] += v

    out["grad_step_ms_1000_rays"] = timeit(grad_step, n=100)
    out["grad_step_stages_ms"] = {k: round(v / 120 * 1e3, 4) for k, v in stages.items()}
    assert leaf.grad is not None and float(leaf.grad.abs()) > 0
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in out.items()}))


if __name__ == "__main__":
    main()
