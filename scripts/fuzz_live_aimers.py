"""Live fuzz, ray aimers: the systems of scripts/fuzz_live_devmath.py with the reference's iterative / robust ray aimer
configured (every solver iteration is a subset trace through the SurfaceGroup capability, one packed table per aiming call:
plugin._FrozenTables), two fields, every record against the unmodified reference on NumPy.  The aimers converge to their own
tolerance (1e-8 here), so agreement is asserted at 1e-7 of the system scale.

    python scripts/fuzz_live_aimers.py <first seed> <last seed>     (CPU only; summary: profiles/r2c_live_fuzz.txt)"""
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
    rng = np.random.default_rng(seed)
    mode = str(rng.choice(["iterative", "robust"]))
    lens.ray_tracer.set_aiming(mode, 10, 1e-8)
    outs = []
    for (hx, hy) in ((0.0, 0.0), (0.0, 1.0)):
        rays = lens.trace(hx, hy, 0.5876, 3, "hexapolar")
        o = {k: np.array(be.to_numpy(getattr(lens.surfaces, k))) for k in REC}
        o["fin_i"] = np.array(be.to_numpy(rays.i))
        outs.append(o)
    return outs, mode
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = skipped = 0; declines = {}; ncalls = 0
for seed in range(lo, hi):
    try:
        be.set_backend("numpy")
        try:
            ref, kinds = build(seed)
            want, mode = calls(ref, seed)
        except Exception as e:
            skipped += 1; continue
        be.set_backend("torch"); be.set_precision("float64"); be.grad_mode.disable()
        P.install(engine=eng); P.stats(reset=True); n0 = len(eng.calls)
        lens, kinds = build(seed)
        got, _ = calls(lens, seed)
        ncalls += len(eng.calls) - n0
        worst = 0.0; wk = None
        for ci, (g_, w_) in enumerate(zip(got, want)):
            for k, v in w_.items():
                g = g_[k]
                if g.shape != v.shape: print(seed, "SHAPE", k); worst = 1; continue
                gn, vn = np.isnan(g), np.isnan(v)
                if not np.array_equal(gn, vn): print(seed, "NANPAT", ci, k, int(gn.sum()), int(vn.sum()))
                m = ~gn & ~vn
                if m.any():
                    sc = max(1.0, float(np.max(np.abs(v[m]))))
                    e = float(np.max(np.abs(g[m] - v[m]))) / sc
                    if e > worst: worst, wk = e, (ci, k)
        for k, v in P.stats().items(): declines[k] = declines.get(k, 0) + v
        if worst > 1e-7:
            bad += 1; print(seed, "MISMATCH", f"{worst:.1e}", wk, mode, kinds, P.stats())
    except Exception as e:
        bad += 1; print(seed, "EXCEPTION", type(e).__name__, str(e)[:300]); traceback.print_exc(limit=5)
    finally:
        if P._state.get("installed"): P.uninstall()
        be.set_backend("numpy")
print("bad", bad, "skipped", skipped, "capability calls", ncalls, "declines", declines)
