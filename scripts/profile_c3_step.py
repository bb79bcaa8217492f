"""Where the time of the differentiable configuration-3 step goes (4 M rays, reverse telephoto + 2 even aspheres):
host table build + upload, forward kernel, the loss in eager torch, its autograd, the adjoint kernel, the D2H of the
gradient block.  Device segments are timed with CUDA events, host segments with perf_counter after a synchronize.

    python scripts/profile_c3_step.py [n_rays [golden case]]     # e.g. 4000000 zernike_fringe
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optiland_b200 import autograd as AG  # noqa: E402
from optiland_b200.trace import DeviceTable, RealRays  # noqa: E402
from scripts.bench_configs import resample  # noqa: E402
from tests._util import Case  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    name = sys.argv[2] if len(sys.argv) > 2 else "telephoto_c3_tol1e-6"
    c = Case(name)
    out = {"system": name, "rays": n, "surfaces": c.table.num_surfaces}
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        base = resample(c, n, dtype)
        params = AG.table_to_params(c.table).requires_grad_(True)
        coefs = AG.table_to_coefs(c.table)
        if coefs is not None:
            coefs = coefs.requires_grad_(True)
        seg = {}

        def mark(name, t0):
            torch.cuda.synchronize()
            seg[name] = seg.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
            return time.perf_counter()

        def step(measure):
            rr = RealRays.__new__(RealRays)
            rr.__dict__.update(base.__dict__)
            t = time.perf_counter()
            if measure:
                # the two host pieces of trace_differentiable's forward, timed on their own
                if coefs is None:
                    packed = AG._packed_from_params(c.table, params)
                    t = mark("host_pack_params", t)
                    DeviceTable(c.table, "cuda:0", packed=packed)
                else:
                    table = AG.params_to_table(c.table, params, coefs)
                    t = mark("host_pack_params", t)
                    DeviceTable(table, "cuda:0")
                t = mark("host_prepare_upload", t)
            rec = AG.trace_differentiable(c.table, params, rr, rows=(-1,), coefs=coefs)
            if measure:
                t = mark("forward_total(incl. the two above again)", t)
            x, y = rec["x"], rec["y"]
            if coefs is not None:       # (a freeform fixture has rays that miss: NaN in band)
                m = torch.isfinite(x) & torch.isfinite(y)
                x, y = torch.where(m, x, 0), torch.where(m, y, 0)
            loss = torch.sqrt(torch.mean((x - x.mean()) ** 2 + (y - y.mean()) ** 2))
            if measure:
                t = mark("loss_eager", t)
            params.grad = None
            if coefs is not None:
                coefs.grad = None
            loss.backward()
            if measure:
                t = mark("backward_total(autograd of the loss + adjoint kernel + D2H)", t)

        for _ in range(3):
            step(False)
        torch.cuda.synchronize()
        K = 10
        t0 = time.perf_counter()
        for _ in range(K):
            step(False)
        torch.cuda.synchronize()
        whole = (time.perf_counter() - t0) / K * 1e3
        for _ in range(K):
            step(True)
        out[tag] = {"ms_step_unsegmented": round(whole, 3), **{k: round(v / K, 3) for k, v in seg.items()}}
        # kernel-only times from the profiler
        from torch.profiler import ProfilerActivity, profile

        try:
          with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(3):
                step(False)
            torch.cuda.synchronize()
          ks = {}
          for e in prof.key_averages():
            dt = getattr(e, "device_time_total", 0) or getattr(e, "cuda_time_total", 0)
            if dt > 0:
                ks[e.key[:70]] = round(dt / 3 / 1e3, 4)
          out[tag]["device_ms_by_kernel"] = dict(sorted(ks.items(), key=lambda kv: -kv[1])[:16])
        except Exception as e:  # (CUPTI not available on the box)
          out[tag]["device_ms_by_kernel"] = f"profiler unavailable: {e}"
        del base
        torch.cuda.empty_cache()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
