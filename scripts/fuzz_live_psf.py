"""Live fuzz, PSFs: FFTPSF (fused wavefront epilogue + the two gridding kernels' arithmetic) and HuygensPSF (the summation
kernel's arithmetic) on the systems of scripts/fuzz_live_devmath.py against the unmodified reference on NumPy.

    python scripts/fuzz_live_psf.py <first seed> <last seed>        (CPU only; summary: profiles/r2c_live_fuzz.txt)"""
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
def calls(lens, seed):
    from optiland.psf import FFTPSF, HuygensPSF
    rng = np.random.default_rng(seed)
    out = {}
    fld = [(0.0, 0.0), (0.0, 1.0)][int(rng.integers(0, 2))]
    wl = "primary" if rng.random() < 0.5 else 0.5876
    try:
        psf = FFTPSF(lens, fld, wl, num_rays=int(rng.choice([16, 24])), grid_size=int(rng.choice([32, 48])))
        out["fft"] = np.array(be.to_numpy(psf.psf)); out["strehl"] = np.array([float(be.to_numpy(be.atleast_1d(psf.strehl_ratio()))[0])])
    except Exception as e:
        out["fft_error"] = np.array([1.0]); out["_e1"] = f"{type(e).__name__}: {str(e)[:80]}"
    try:
        h = HuygensPSF(lens, fld, 0.5876, num_rays=12, image_size=16)
        out["huygens"] = np.array(be.to_numpy(h.psf))
    except Exception as e:
        out["huy_error"] = np.array([1.0]); out["_e2"] = f"{type(e).__name__}: {str(e)[:80]}"
    return out
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = skipped = 0; declines = {}; kinds_seen = {}
for seed in range(lo, hi):
    try:
        be.set_backend("numpy")
        try:
            ref, kinds = build(seed)
            want = calls(ref, seed)
        except Exception as e:
            skipped += 1; continue
        be.set_backend("torch"); be.set_precision("float64"); be.grad_mode.disable()
        P.install(engine=eng); P.stats(reset=True); n0 = len(eng.calls)
        lens, kinds = build(seed)
        got = calls(lens, seed)
        for c in eng.calls[n0:]:
            kinds_seen[c[0] if isinstance(c[0], str) else "trace"] = kinds_seen.get(c[0] if isinstance(c[0], str) else "trace", 0) + 1
        ks = lambda d: {k for k in d if not k.startswith("_")}
        if ks(got) != ks(want):
            bad += 1; print(seed, "KEYS", sorted(ks(got) ^ ks(want)), {k: v for k, v in {**want, **got}.items() if k.startswith("_")}, kinds); continue
        worst = 0.0; wk = None
        for k in ks(want):
            g, v = got[k], want[k]
            if g.shape != v.shape: print(seed, "SHAPE", k, g.shape, v.shape); worst = 1; continue
            sc = max(1e-12, float(np.nanmax(np.abs(v))))
            m = np.isfinite(v) & np.isfinite(g)
            if not np.array_equal(np.isfinite(v), np.isfinite(g)): print(seed, "NANPAT", k)
            e = float(np.max(np.abs(g[m] - v[m]))) / sc if m.any() else 0.0
            if e > worst: worst, wk = e, k
        for k, v in P.stats().items(): declines[k] = declines.get(k, 0) + v
        if worst > 1e-6:
            bad += 1; print(seed, "MISMATCH", f"{worst:.1e}", wk, kinds, P.stats())
    except Exception as e:
        bad += 1; print(seed, "EXCEPTION", type(e).__name__, str(e)[:300]); traceback.print_exc(limit=5)
    finally:
        if P._state.get("installed"): P.uninstall()
        be.set_backend("numpy")
print("bad", bad, "skipped", skipped, "engine calls", kinds_seen, "declines", declines)
