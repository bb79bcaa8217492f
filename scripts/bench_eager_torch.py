"""The "GPU bar" of SURVEY.md 8(d): what an EAGER torch-CUDA evaluation of the same trace costs on the same
B200.  The reference's torch backend runs the per-surface Python walk with one small CUDA kernel per
element-wise op; the reference itself cannot travel to the GPU box, so this script restates that op
sequence (plane / conic surfaces, refraction, absorbing media, full records; the Double-Gauss of
bench.py) with plain torch ops -- NO Python-side material lookups, cache keys or object churn, i.e. a
LOWER bound on the reference's eager time -- and times it beside the fused kernel on identical launch rays.

Op order follows optiland/surfaces/standard_surface.py:232-248 (localize, distance, propagate, OPD, normal,
refract, globalize, record), geometries/standard.py:97-175, rays/real_rays.py:163-187, 535-571.
Prints one JSON line; the two results are also compared (max |dx| at the image surface)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WAVELENGTH, load_case  # noqa: E402
from optiland_b200 import table as T  # noqa: E402
from optiland_b200.launch import launch_infinite_angle  # noqa: E402
from optiland_b200.trace import DeviceTable, RealRays, trace_device  # noqa: E402


def eager_trace(table, x, y, z, L, M, N, inten, wavelength):
    """Eager restatement for unrotated plane / standard surfaces and one wavelength."""
    opd = torch.zeros_like(x)
    recs = []
    for s in table.surfaces:
        if s.kind == T.GEOM_NOOP:
            recs.append([t.clone() for t in (x, y, z, L, M, N, inten, opd)])
            continue
        tx, ty, tz = (float(v) for v in s.t)
        x, y, z = x - tx, y - ty, z - tz                                  # localize
        if s.kind == T.GEOM_PLANE:
            t = -z / N
        else:
            R, k = float(s.radius), float(s.conic)
            a = k * N**2 + L**2 + M**2 + N**2
            b = 2 * k * N * z + 2 * L * x + 2 * M * y - 2 * N * R + 2 * N * z
            c = k * z**2 - 2 * R * z + x**2 + y**2 + z**2
            d = b**2 - 4 * a * c
            d = torch.where(d < 0, torch.full_like(d, float("nan")), d)
            t1 = (-b + torch.sqrt(d)) / (2 * a)
            t2 = (-b - torch.sqrt(d)) / (2 * a)
            z1, z2 = z + t1 * N, z + t2 * N
            t = torch.where(torch.abs(z1) <= torch.abs(z2), t1, t2)
        x, y, z = x + t * L, y + t * M, z + t * N                          # propagate
        if float(s.k1[0]) > 0:                                            # homogeneous.py:45-53
            inten = inten * torch.exp(-(4 * np.pi * float(s.k1[0]) / wavelength) * t * 1e3)
        opd = opd + torch.abs(t * float(s.n1[0]))
        if s.aperture is not None and int(s.aperture[0]) == T.AP_RADIAL:
            r2 = x**2 + y**2
            inside = (r2 <= float(s.aperture[1]) ** 2) & (r2 >= float(s.aperture[2]) ** 2)
            inten = torch.where(inside, inten, torch.zeros_like(inten))
        if s.kind == T.GEOM_PLANE:
            nx, ny, nz = torch.zeros_like(x), torch.zeros_like(x), torch.ones_like(x)
        else:
            r2 = x**2 + y**2
            denom = R * torch.sqrt(1 - (1 + k) * r2 / R**2)
            dfdx, dfdy, dfdz = x / denom, y / denom, -1.0
            mag = torch.sqrt(dfdx**2 + dfdy**2 + dfdz**2)
            nx, ny, nz = dfdx / mag, dfdy / mag, dfdz / mag
        dot = L * nx + M * ny + N * nz                                      # align the normal
        sgn = torch.sign(dot)
        nx, ny, nz = nx * sgn, ny * sgn, nz * sgn
        dot = torch.abs(dot)
        if s.reflective:
            L, M, N = L - 2 * dot * nx, M - 2 * dot * ny, N - 2 * dot * nz
        else:
            u = float(s.n1[0]) / float(s.n2[0])
            root = torch.sqrt(1 - u**2 * (1 - dot**2))
            L = u * L + nx * root - u * nx * dot
            M = u * M + ny * root - u * ny * dot
            N = u * N + nz * root - u * nz * dot
        x, y, z = x + tx, y + ty, z + tz                                    # globalize
        recs.append([t_.clone() for t_ in (x, y, z, L, M, N, inten, opd)])  # record
    return (x, y, z, L, M, N, inten, opd), recs


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
    dev = torch.device("cuda:0")
    c, sc = load_case()
    assert all(not s.rotated and s.kind in (T.GEOM_NOOP, T.GEOM_PLANE, T.GEOM_STANDARD) for s in c.table.surfaces)
    S = c.table.num_surfaces
    out = {"workload": f"double_gauss_13surf_{n}rays_full_records"}
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        g = torch.Generator(device=dev).manual_seed(0)
        r = torch.rand(n, generator=g, device=dev, dtype=torch.float64).sqrt()
        th = 2 * np.pi * torch.rand(n, generator=g, device=dev, dtype=torch.float64)
        x0, y0, z0, L, M, N = (t.to(dtype) for t in launch_infinite_angle(r * torch.cos(th), r * torch.sin(th), sc))
        inten = torch.ones_like(x0)
        ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

        def timed(fn, reps):
            # results are dropped before the next call and the warm-up is long enough for the caching
            # allocator to stop growing: a cudaMalloc inside the loop would be timed as GPU idle time
            res = None
            for _ in range(4):
                res = None
                res = fn()
            torch.cuda.synchronize()
            a, b = ev(), ev()
            a.record()
            for _ in range(reps):
                res = None
                res = fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps, res

        ms_eager, (fin, recs) = timed(lambda: eager_trace(c.table, x0, y0, z0, L, M, N, inten, WAVELENGTH), 5)
        dtab = DeviceTable(c.table, dev)

        base = RealRays(x0, y0, z0, L, M, N, inten, WAVELENGTH, dtype=dtype, device=dev)

        def ours():
            rr = RealRays.__new__(RealRays)      # trace_device re-points the attributes at the record rows
            rr.__dict__.update(base.__dict__)
            return trace_device(dtab, rr, 0, S)

        ms_ours, rec = timed(ours, 20)
        dx = float((rec["x"][-1] - recs[-1][0]).abs().max())
        rs = n * (S - 1)
        out[tag] = {"eager_torch_ms": round(ms_eager, 3), "eager_torch_ray_surfaces_per_s": rs / ms_eager * 1e3,
                    "fused_kernel_ms": round(ms_ours, 4), "fused_ray_surfaces_per_s": rs / ms_ours * 1e3,
                    "speedup": ms_eager / ms_ours, "max_abs_dx_image_mm": dx}
        del recs, rec, fin
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
