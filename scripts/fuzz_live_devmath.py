"""Fuzz campaign with LIVE reference objects through the plugin over the device-math engine (oracle/devmath_engine.py:
the kernel's arithmetic compiled for the host) -- wider than tests/test_zzz_live_fuzz.py: every geometry family, planes,
hyperbolas, mirrors, catalogue glasses (never the same glass on both sides of a curved surface: there the REFERENCE's
polarization basis is rounding noise), decenters / tilts, radial / rectangular / elliptical / boolean
apertures, four system-aperture types, wide fields with vignetting factors, three wavelengths, polarized and unpolarized
states, Gaussian apodization; per system three Optic.trace calls (random pupil distribution) and one trace_generic call
with per-ray field / pupil arrays, every record + final intensity + P matrices against the NumPy reference.

    python scripts/fuzz_live_devmath.py <first seed> <last seed>      (CPU only; ~0.25 s per system)

Summary of the campaign run for round 2: profiles/r2c_live_fuzz.txt."""
import os
import sys
import traceback
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import numpy as np
import tests.test_zzz_live_fuzz as F
from oracle.ref_import import import_reference
import_reference()
import optiland.backend as be
from optiland_b200 import plugin as P
from oracle.devmath_engine import DeviceMathEngine
from oracle.oracle_engine import OracleEngine
eng = DeviceMathEngine()

def build(seed):
    from optiland import optic as _optic
    from optiland import physical_apertures as PA
    from optiland.rays import PolarizationState
    from optiland.apodization import GaussianApodization
    rng = np.random.default_rng(seed)
    build._prev = None
    lens = _optic.Optic()
    finite = rng.random() < 0.3
    lens.surfaces.add(index=0, radius=be.inf, thickness=(float(rng.uniform(80, 200)) if finite else be.inf))
    n = int(rng.integers(3, 7))
    stop = int(rng.integers(1, n + 1))
    kinds, in_glass = [], False
    for i in range(1, n + 1):
        radius = float(rng.choice([-1, 1]) * rng.uniform(20, 120))
        r = rng.random()
        if r < 0.1:
            kind, kw = "plane", dict(radius=be.inf)
        elif r < 0.0:
            kind, kw = "parabola", dict(radius=radius, conic=-1.0)
        elif r < 0.3:
            kind, kw = "hyperbola", dict(radius=radius, conic=float(rng.uniform(-3, -1.3)))
        else:
            kind, kw = F._surface(rng, radius)
        mirror = (i < n) and (not in_glass) and rng.random() < 0.1 and kind in ("standard", "conic", "plane", "parabola", "hyperbola", "even_asphere")
        if mirror:
            kw["material"] = "mirror"; kind += "+mirror"
        else:
            in_glass = (not in_glass) if (rng.random() < 0.8 or not in_glass) else in_glass
            if in_glass:
                prev_glass = getattr(build, "_prev", None)
                gl = str(rng.choice([g for g in F.GLASSES if g != prev_glass]))
                build._prev = gl
                kw["material"] = gl
        kinds.append(kind)
        kw["thickness"] = float(rng.uniform(2.0, 6.0)) if i < n else float(rng.uniform(30, 80))
        if mirror: kw["thickness"] = -kw["thickness"] if False else kw["thickness"]
        if rng.random() < 0.25:
            kw["dx"], kw["dy"] = float(rng.normal(0, 0.2)), float(rng.normal(0, 0.2))
        if rng.random() < 0.25:
            kw["rx"], kw["ry"] = float(rng.normal(0, 0.02)), float(rng.normal(0, 0.02))
        a = rng.random()
        if a < 0.15:
            kw["aperture"] = PA.RadialAperture(r_max=float(rng.uniform(2.5, 6.0)), r_min=float(rng.choice([0.0, 0.5])))
        elif a < 0.25:
            kw["aperture"] = PA.RectangularAperture(x_min=-4.0, x_max=3.5, y_min=-3.0, y_max=4.0)
        elif a < 0.32:
            kw["aperture"] = PA.EllipticalAperture(a=4.0, b=3.0, offset_x=0.2, offset_y=-0.1)
        elif a < 0.38:
            kw["aperture"] = PA.RadialAperture(r_max=5.0) - PA.RectangularAperture(x_min=-0.5, x_max=0.5, y_min=-6, y_max=6)
        elif a < 0.42:
            kw["aperture"] = PA.RectangularAperture(x_min=-4.0, x_max=0.0, y_min=-3.0, y_max=4.0) | PA.RadialAperture(r_max=2.0)
        lens.surfaces.add(index=i, is_stop=(i == stop), **kw)
    lens.surfaces.add(index=n + 1)
    at = str(rng.choice(["EPD", "imageFNO", "objectNA" if finite else "EPD", "float_by_stop_size"]))
    val = {"EPD": float(rng.uniform(4.0, 9.0)), "imageFNO": float(rng.uniform(4, 10)), "objectNA": float(rng.uniform(0.01, 0.04)),
           "float_by_stop_size": float(rng.uniform(2.0, 4.0))}[at]
    lens.set_aperture(aperture_type=at, value=val)
    lens.fields.set_type(field_type="object_height" if finite else "angle")
    fmax = float(rng.choice([3.0, 10.0, 25.0])) if not finite else float(rng.choice([3.0, 10.0]))
    lens.fields.add(y=0.0)
    lens.fields.add(y=fmax, x=float(rng.choice([0.0, fmax / 2])))
    if rng.random() < 0.3:
        lens.fields.fields[-1].vx = 0.1; lens.fields.fields[-1].vy = 0.2
    lens.wavelengths.add(value=0.4861); lens.wavelengths.add(value=0.5876, is_primary=True); lens.wavelengths.add(value=0.6563)
    pol = rng.random() < 0.35
    if pol:
        if rng.random() < 0.5:
            lens.set_polarization(PolarizationState(is_polarized=False))
        else:
            lens.set_polarization(PolarizationState(is_polarized=True, Ex=1.0, Ey=float(rng.uniform(0, 1)), phase_x=0.0, phase_y=float(rng.uniform(0, 1))))
    apod = rng.random() < 0.2
    if apod:
        lens.set_apodization(GaussianApodization(sigma=float(rng.uniform(0.5, 2.0))))
    return lens, kinds + [at, "finite" if finite else "inf", "pol" if pol else "", "apod" if apod else "", f"f{fmax}"]

REC = F.REC
def run_calls(lens, rng_seed):
    rng = np.random.default_rng(rng_seed)
    outs = []
    dist = str(rng.choice(["ring", "hexapolar", "uniform", "line_y", "cross", "random"]))
    if dist == "random": dist = "hexapolar"
    wl = float(rng.choice([0.4861, 0.5876, 0.6563]))
    for (hx, hy) in ((0.0, 0.0), (0.5, 1.0), (0.0, -0.7)):
        rays = lens.trace(hx, hy, wl, 6, dist)
        o = {k: np.array(be.to_numpy(getattr(lens.surfaces, k))) for k in REC}
        o["fin_i"] = np.array(be.to_numpy(rays.i)); o["fin_opd"] = np.array(be.to_numpy(rays.opd))
        if hasattr(rays, "p"): o["p"] = np.array(be.to_numpy(rays.p))
        outs.append(o)
    # trace_generic with per-ray arrays
    m = 24
    Hx = be.array(rng.uniform(-0.5, 0.5, m)); Hy = be.array(rng.uniform(-1, 1, m)); Px = be.array(rng.uniform(-0.7, 0.7, m)); Py = be.array(rng.uniform(-0.7, 0.7, m))
    rays = lens.trace_generic(Hx, Hy, Px, Py, wl)
    o = {k: np.array(be.to_numpy(getattr(lens.surfaces, k))) for k in REC}
    o["fin_i"] = np.array(be.to_numpy(rays.i))
    if hasattr(rays, "p"): o["p"] = np.array(be.to_numpy(rays.p))
    outs.append(o)
    return outs

lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0; skipped = 0; declines = {}; nanpat = {}
for seed in range(lo, hi):
    try:
        be.set_backend("numpy")
        try:
            ref, kinds = build(seed)
            want = run_calls(ref, seed)
        except Exception as e:
            skipped += 1; continue
        be.set_backend("torch"); be.set_precision("float64"); be.grad_mode.disable()
        P.install(engine=eng); P.stats(reset=True)
        lens, kinds = build(seed)
        got = run_calls(lens, seed)
        worst = 0.0; wk = None
        for ci, (g_, w_) in enumerate(zip(got, want)):
            for k, v in w_.items():
                g = g_[k]
                if g.shape != v.shape:
                    print(seed, "SHAPE", k, g.shape, v.shape); worst = 1; continue
                gn = np.isnan(g) if not np.iscomplexobj(g) else np.isnan(g.real) | np.isnan(g.imag)
                vn = np.isnan(v) if not np.iscomplexobj(v) else np.isnan(v.real) | np.isnan(v.imag)
                if not np.array_equal(gn, vn):
                    nanpat.setdefault(seed, []).append((ci, k, int(gn.sum()), int(vn.sum())))
                m = ~gn & ~vn
                if m.any():
                    sc = max(1.0, float(np.max(np.abs(v[m])))) if k in ("x", "y", "z", "opd", "fin_opd") else 1.0
                    e = float(np.max(np.abs(g[m] - v[m]))) / sc
                    if e > worst: worst, wk = e, (ci, k)
        for k, v in P.stats().items(): declines[k] = declines.get(k, 0) + v
        if worst > 1e-9:
            bad += 1; print(seed, "MISMATCH", f"{worst:.1e}", wk, kinds, P.stats())
    except Exception as e:
        bad += 1; print(seed, "EXCEPTION", type(e).__name__, str(e)[:300]); traceback.print_exc(limit=4)
    finally:
        if P._state.get("installed"): P.uninstall()
        be.set_backend("numpy")
print("nan-pattern differences:", {k: v[:2] for k, v in nanpat.items()})
print("bad", bad, "skipped", skipped, "declines", declines)
