"""The GPU bar SURVEY.md 8(d) asks for: the UNMODIFIED reference's own torch-CUDA eager path
(`be.set_backend("torch"); be.set_device("cuda")`, optiland/backend/torch_backend.py:64-78) timed on the same
B200 next to the fused kernel reached through the plugin -- same live `DoubleGauss()` object, same launch rays
(the reference's own RayGenerator), `optic.surfaces.trace(rays)` as the timed call (t_trace).

    python scripts/bench_reference_torch_cuda.py [--rays 1000000 10000000] [--precision float32 float64]

Needs the reference (oracle/_ref on the GPU box: scripts/make_ref.sh).  One JSON line per (precision, N).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, nargs="+", default=[1_000_000, 10_000_000])
    ap.add_argument("--precision", nargs="+", default=["float32", "float64"])
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()

    import torch

    from oracle.ref_import import import_reference

    import_reference()
    import optiland.backend as be
    from optiland.samples.objectives import DoubleGauss

    from optiland_b200 import _lib
    from optiland_b200 import plugin as P

    assert torch.cuda.is_available()
    be.set_backend("torch")
    be.set_device("cuda")
    be.grad_mode.disable()
    lib = _lib.load()

    for prec in args.precision:
        be.set_precision(prec)
        for n in args.rays:
            rng = np.random.default_rng(0)
            r = np.sqrt(rng.random(n))
            th = 2 * np.pi * rng.random(n)
            Px, Py = be.array(r * np.cos(th)), be.array(r * np.sin(th))
            zeros = be.zeros_like(Px)
            lens = DoubleGauss()
            S = lens.surfaces.num_surfaces

            def launch():
                return lens.ray_tracer.ray_generator.generate_rays(zeros, zeros, Px, Py, 0.5876)

            def timed(fn, reps):
                fn()                       # warm-up (allocator, caches, table upload)
                fn()
                torch.cuda.synchronize()
                ts = []
                for _ in range(reps):
                    rays = launch()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    fn(rays)
                    torch.cuda.synchronize()
                    ts.append(time.perf_counter() - t0)
                return float(np.median(ts)), float(np.min(ts))

            def trace(rays=None):
                rays = rays if rays is not None else launch()
                lens.surfaces.trace(rays)
                return rays

            # ---- stock reference: eager torch ops on the GPU -----------------------------------
            P.uninstall()
            t_ref, t_ref_min = timed(trace, args.reps)
            ref_rec = {k: getattr(lens.surfaces, k).clone() for k in ("x", "y", "opd", "L")}
            peak_ref = torch.cuda.max_memory_allocated() / 1e9
            torch.cuda.reset_peak_memory_stats()

            # ---- same call, plugin installed: one fused launch -----------------------------------
            P.install()
            P.stats(reset=True)
            l0 = lib.olb_launch_count()
            t_ours, t_ours_min = timed(trace, args.reps)
            launches = lib.olb_launch_count() - l0
            declines = P.stats()
            err = {k: float((getattr(lens.surfaces, k) - ref_rec[k]).abs().max()) for k in ref_rec}
            peak_ours = torch.cuda.max_memory_allocated() / 1e9
            P.uninstall()
            del ref_rec
            torch.cuda.empty_cache()
            print(json.dumps({
                "what": "SurfaceGroup.trace on DoubleGauss (13 surfaces / 12 traced), launch rays from the reference's "
                        "RayGenerator, wall clock around optic.surfaces.trace(rays) with synchronize on both sides",
                "precision": prec, "rays": n,
                "reference_torch_cuda_ms": 1e3 * t_ref, "reference_torch_cuda_ms_min": 1e3 * t_ref_min,
                "reference_ray_surfaces_per_s": n * (S - 1) / t_ref,
                "plugin_ms": 1e3 * t_ours, "plugin_ms_min": 1e3 * t_ours_min,
                "plugin_ray_surfaces_per_s": n * (S - 1) / t_ours,
                "speedup": t_ref / t_ours, "olb_launches": int(launches), "declines": declines,
                "max_abs_diff_vs_reference": err, "peak_mem_gb": {"reference": peak_ref, "plugin": peak_ours},
            }), flush=True)


if __name__ == "__main__":
    main()
