"""Per-call cost of the drop-in at SMALL ray counts, with LIVE Optiland objects (torch backend, device cuda) -- the size
the reference's optimisation loops live at, where host work, not the kernel, is the time:

  * `Optic.trace(Hx, Hy, wl, num_rays, "hexapolar")`            (fused launch; packing, launch scalars, upload, launch)
  * `optic.surfaces.trace(rays)` on pre-made launch rays         (SurfaceGroup.trace capability)
  * the differentiable step of the torch optimiser's inner loop  (be.grad_mode on: trace -> RMS spot -> backward)
  * the same three on the STOCK reference (eager torch ops on the same GPU)

    python scripts/host_overhead_live.py [--rays-rings 18]        # 18 hexapolar rings = 1027 rays
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rings", type=int, default=18)
    ap.add_argument("--reps", type=int, default=100)
    ap.add_argument("--cpu-smoke", action="store_true", help="script self-check without a GPU (test-only oracle engine)")
    args = ap.parse_args()
    import torch

    from oracle.ref_import import import_reference

    import_reference()
    import optiland.backend as be
    from optiland.samples.objectives import DoubleGauss

    from oracle.make_golden import reverse_telephoto_asphere
    from optiland_b200 import _lib
    from optiland_b200 import plugin as P

    be.set_backend("torch")
    if not args.cpu_smoke:
        be.set_device("cuda")
    be.set_precision("float64")
    lib = _lib.load()
    sync = torch.cuda.synchronize if not args.cpu_smoke else (lambda: None)

    def timeit(fn, n):
        for _ in range(5):
            fn()
        sync()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        sync()
        return (time.perf_counter() - t0) / n * 1e3

    out = {"rays": 1 + 3 * args.rings * (args.rings + 1), "precision": "float64", "device": torch.cuda.get_device_name(0) if not args.cpu_smoke else "cpu (smoke)"}

    def measure(tag, reps):
        res = {}
        be.grad_mode.disable()
        lens = DoubleGauss()
        res["optic_trace_ms"] = timeit(lambda: lens.trace(0.0, 0.7, 0.5876, args.rings, "hexapolar"), reps)
        rays0 = lens.ray_tracer.ray_generator.generate_rays(0.0, 0.7, lens_px(lens, args.rings)[0], lens_px(lens, args.rings)[1], 0.5876)
        keep = {k: getattr(rays0, k).clone() for k in ("x", "y", "z", "L", "M", "N", "i", "w")}
        from optiland.rays import RealRays

        def sg():
            r = RealRays(*[keep[k] for k in ("x", "y", "z", "L", "M", "N", "i", "w")])
            lens.surfaces.trace(r)

        res["surface_group_trace_ms"] = timeit(sg, reps)
        # a parameter changes before every call (what an optimiser does): nothing about the table can be reused
        k = [0]

        def changed():
            k[0] += 1
            lens.surfaces.surfaces[3].geometry.radius = be.array(float(lens_r3) * (1 + 1e-9 * k[0]))
            lens.trace(0.0, 0.7, 0.5876, args.rings, "hexapolar")

        lens_r3 = float(lens.surfaces.surfaces[3].geometry.radius)
        res["optic_trace_changed_parameter_ms"] = timeit(changed, reps)
        # differentiable step (config 3's system)
        be.grad_mode.enable()
        try:
            tele = reverse_telephoto_asphere(1e-10)

            def grad_step():
                tele.trace(0.0, 0.7, 0.5876, args.rings, "hexapolar")
                x = tele.surfaces.x[-1, :]
                y = tele.surfaces.y[-1, :]
                loss = torch.sqrt(torch.mean((x - torch.mean(x)) ** 2 + (y - torch.mean(y)) ** 2))
                loss.backward()
                for s in tele.surfaces.surfaces[1:]:
                    g = s.geometry
                    for t in (getattr(g, "radius", None), g.cs.z):
                        if getattr(t, "grad", None) is not None:
                            t.grad = None

            res["grad_step_ms"] = timeit(grad_step, max(10, reps // 4))
        finally:
            be.grad_mode.disable()
        out[tag] = res

    def lens_px(lens, rings):
        from optiland.distribution import create_distribution

        d = create_distribution("hexapolar")
        d.generate_points(rings)
        return d.x, d.y

    P.uninstall()
    measure("reference_torch_cuda", max(5, args.reps // 10))
    if args.cpu_smoke:
        from oracle.oracle_engine import OracleEngine

        P.install(engine=OracleEngine())
    else:
        P.install()
    P.stats(reset=True)
    l0 = lib.olb_launch_count()
    measure("plugin", args.reps)
    out["plugin"]["olb_launches"] = int(lib.olb_launch_count() - l0)
    out["plugin"]["declines"] = P.stats()
    out["speedup"] = {k: out["reference_torch_cuda"][k] / out["plugin"][k] for k in out["reference_torch_cuda"]}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
