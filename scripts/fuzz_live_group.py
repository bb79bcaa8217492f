"""Live fuzz, the capability boundary itself: ``SurfaceGroup.trace(rays[, skip])`` and ``Surface.trace(rays)`` called directly with
user-made ``RealRays`` (arbitrary launch points and directions, per-ray wavelengths, intensities in [0, 1] incl. exact zeros,
a non-zero initial OPD) on the unpolarized systems of scripts/fuzz_live_devmath.py; records and the final ray state against
the unmodified reference on NumPy.

    python scripts/fuzz_live_group.py <first seed> <last seed>      (CPU only; summary: profiles/r2c_live_fuzz.txt)"""
import os, sys, traceback, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); warnings.filterwarnings("ignore")
import numpy as np
import tests.test_zzz_live_fuzz as F
from oracle.ref_import import import_reference
import_reference()
import optiland.backend as be
from optiland_b200 import plugin as P
from oracle.devmath_engine import DeviceMathEngine
exec(open(os.path.join(ROOT, "scripts", "fuzz_live_devmath.py")).read().split("REC = F.REC")[0].split("eng = DeviceMathEngine()")[1])
eng = DeviceMathEngine()
REC = F.REC
def calls(lens, seed):
    from optiland.rays import RealRays
    rng = np.random.default_rng(seed)
    n = 40
    z0 = -5.0 if bool(lens.object_surface.is_infinite) else float(be.to_numpy(be.atleast_1d(lens.surfaces.surfaces[0].geometry.cs.z))[0])
    x, y = rng.uniform(-3, 3, n), rng.uniform(-3, 3, n)
    L, M = rng.normal(0, 0.03, n), rng.normal(0, 0.03, n)
    N = np.sqrt(1 - L**2 - M**2)
    inten = rng.uniform(0.2, 1.0, n); inten[rng.integers(0, n, 3)] = 0.0
    wl = rng.choice([0.4861, 0.5876, 0.6563], n) if rng.random() < 0.5 else np.full(n, 0.5876)
    outs = []
    for skip in (0, 1, 2):
        rays = RealRays(be.array(x), be.array(y), be.array(np.full(n, z0)), be.array(L), be.array(M), be.array(N), be.array(inten), be.array(wl))
        rays.opd = be.array(rng.uniform(0, 1, n))
        if skip == 2:     # single-surface calls
            for s in lens.surfaces.surfaces[1:3]:
                s.trace(rays)
            o = {k: np.array(be.to_numpy(getattr(lens.surfaces.surfaces[2], k))) for k in ("x", "y", "z", "L", "M", "N", "opd", "intensity")}
        else:
            lens.surfaces.trace(rays, skip) if skip else lens.surfaces.trace(rays)
            o = {k: np.array(be.to_numpy(getattr(lens.surfaces, k))) for k in REC}
        for k in ("x", "y", "z", "L", "M", "N", "i", "opd"):
            o["fin_" + k] = np.array(be.to_numpy(getattr(rays, k)))
        outs.append(o)
    return outs
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = skipped = 0; declines = {}; ncalls = 0
for seed in range(lo, hi):
    try:
        be.set_backend("numpy")
        try:
            ref, kinds = build(seed)
            if ref.polarization != "ignore": skipped += 1; continue
            want = calls(ref, seed)
        except Exception as e:
            skipped += 1; continue
        be.set_backend("torch"); be.set_precision("float64"); be.grad_mode.disable()
        P.install(engine=eng); P.stats(reset=True); n0 = len(eng.calls)
        lens, kinds = build(seed)
        got = calls(lens, seed)
        ncalls += len(eng.calls) - n0
        worst = 0.0; wk = None
        for ci, (g_, w_) in enumerate(zip(got, want)):
            for k, v in w_.items():
                g = g_[k]
                if g.shape != v.shape: print(seed, "SHAPE", ci, k, g.shape, v.shape); worst = 1; continue
                gn, vn = np.isnan(g), np.isnan(v)
                if not np.array_equal(gn, vn): print(seed, "NANPAT", ci, k, int(gn.sum()), int(vn.sum()))
                m = ~gn & ~vn
                if m.any():
                    sc = max(1.0, float(np.max(np.abs(v[m]))))
                    e = float(np.max(np.abs(g[m] - v[m]))) / sc
                    if e > worst: worst, wk = e, (ci, k)
        for k, v in P.stats().items(): declines[k] = declines.get(k, 0) + v
        if worst > 1e-9:
            bad += 1; print(seed, "MISMATCH", f"{worst:.1e}", wk, kinds, P.stats())
    except Exception as e:
        bad += 1; print(seed, "EXCEPTION", type(e).__name__, str(e)[:300]); traceback.print_exc(limit=5)
    finally:
        if P._state.get("installed"): P.uninstall()
        be.set_backend("numpy")
print("bad", bad, "skipped", skipped, "capability calls", ncalls, "declines", declines)
