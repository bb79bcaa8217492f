"""Analysis-level fuzz campaign with LIVE reference objects through the plugin over the device-math engine (the systems of
scripts/fuzz_live_devmath.py): per random system a ``Wavefront`` over all fields x wavelengths with a random strategy
(chief_ray -> the fused wavefront epilogue; centroid_sphere / best_fit_sphere -> fused launch + the reference's own
reductions), a ``SpotDiagram`` (rms_spot_radius / centroid from moment launches, geometric_spot_radius from the lazily
materialised spots; local / global coordinates, chief-ray / centroid reference) and a ``trace_generic`` call with per-ray
field, pupil AND wavelength arrays -- every number against the unmodified reference on its NumPy backend; an exception on
one side must be an exception on the other (the Zernike / Chebyshev range errors).

    python scripts/fuzz_live_analyses.py <first seed> <last seed>      (CPU only)

Summary of the round-2 campaign: profiles/r2c_live_fuzz.txt."""
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
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fuzz_live_devmath.py")).read().split("REC = F.REC")[0].split("eng = DeviceMathEngine()")[1])
eng = DeviceMathEngine()

def analyses(lens, seed):
    from optiland.wavefront import Wavefront
    from optiland.analysis import SpotDiagram
    rng = np.random.default_rng(seed)
    out = {}
    strat = str(rng.choice(["chief_ray", "centroid_sphere", "best_fit_sphere"]))
    try:
        wf = Wavefront(lens, fields="all", wavelengths="all", num_rays=int(rng.choice([6, 9])), distribution=str(rng.choice(["hexapolar", "uniform"])), strategy=strat)
        for (fld, wl), d in wf.data.items():
            out[f"wf_opd_{fld}_{wl}"] = np.array(be.to_numpy(d.opd)); out[f"wf_i_{fld}_{wl}"] = np.array(be.to_numpy(d.intensity))
            out[f"wf_px_{fld}_{wl}"] = np.array(be.to_numpy(d.pupil_x)); out[f"wf_r_{fld}_{wl}"] = np.array(be.to_numpy(be.atleast_1d(d.radius)))
    except Exception as e:
        out["wf_error"] = np.array([1.0])
    coords = str(rng.choice(["local", "global"])); refp = str(rng.choice(["chief_ray", "centroid"]))
    try:
        sd = SpotDiagram(lens, num_rings=int(rng.choice([3, 5])), distribution=str(rng.choice(["hexapolar", "random" if False else "ring"])), coordinates=coords, reference=refp)
        out["sd_rms"] = np.array([[float(be.to_numpy(v).reshape(-1)[0]) for v in row] for row in sd.rms_spot_radius()])
        out["sd_cen"] = np.array([[float(be.to_numpy(c).reshape(-1)[0]) for c in pair] for pair in sd.centroid()])
        out["sd_geo"] = np.array([[float(be.to_numpy(v).reshape(-1)[0]) for v in row] for row in sd.geometric_spot_radius()])
    except Exception as e:
        out["sd_error"] = np.array([1.0])
    m = 30
    Hx = be.array(rng.uniform(-0.5, 0.5, m)); Hy = be.array(rng.uniform(-1, 1, m)); Px = be.array(rng.uniform(-0.7, 0.7, m)); Py = be.array(rng.uniform(-0.7, 0.7, m))
    W = be.array(rng.choice([0.4861, 0.5876, 0.6563], m))
    try:
        rays = lens.trace_generic(Hx, Hy, Px, Py, W)
        for k in ("x", "y", "opd", "intensity"):
            out["gen_" + k] = np.array(be.to_numpy(getattr(lens.surfaces, k)))
        out["gen_fin_i"] = np.array(be.to_numpy(rays.i))
    except Exception as e:
        out["gen_error"] = np.array([1.0])
    return out, (strat, coords, refp)

lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0; skipped = 0; declines = {}; nanpat = {}
for seed in range(lo, hi):
    try:
        be.set_backend("numpy")
        try:
            ref, kinds = build(seed)
            want, cfg = analyses(ref, seed)
        except Exception as e:
            skipped += 1; continue
        be.set_backend("torch"); be.set_precision("float64"); be.grad_mode.disable()
        P.install(engine=eng); P.stats(reset=True)
        lens, kinds = build(seed)
        got, _ = analyses(lens, seed)
        worst = 0.0; wk = None
        if set(got) != set(want):
            print(seed, "KEYS DIFFER", sorted(set(got) ^ set(want)), kinds, cfg); bad += 1; continue
        for k, v in want.items():
            g = got[k]
            if g.shape != v.shape:
                print(seed, "SHAPE", k, g.shape, v.shape); worst = 1; continue
            gn, vn = np.isnan(g), np.isnan(v)
            if not np.array_equal(gn, vn):
                nanpat.setdefault(seed, []).append((k, int(gn.sum()), int(vn.sum())))
            m = ~gn & ~vn
            if m.any():
                sc = max(1.0, float(np.max(np.abs(v[m])))) if not k.startswith("wf_opd") else 1.0
                e = float(np.max(np.abs(g[m] - v[m]))) / sc
                if e > worst: worst, wk = e, k
        for k, v in P.stats().items(): declines[k] = declines.get(k, 0) + v
        if worst > 1e-7:
            bad += 1; print(seed, "MISMATCH", f"{worst:.1e}", wk, kinds, cfg, P.stats())
    except Exception as e:
        bad += 1; print(seed, "EXCEPTION", type(e).__name__, str(e)[:300]); traceback.print_exc(limit=5)
    finally:
        if P._state.get("installed"): P.uninstall()
        be.set_backend("numpy")
print("nan-pattern differences:", {k: v[:3] for k, v in nanpat.items()})
print("bad", bad, "skipped", skipped, "declines", declines)
