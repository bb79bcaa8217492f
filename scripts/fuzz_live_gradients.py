"""Live fuzz, gradients: the random systems of tests/test_zzz_live_fuzz.py with every float surface parameter a leaf that
requires grad; one differentiable step (two fields, spot + OPD loss); the plugin's gradients (forward kernel + hand-derived
adjoint; test-only engine = oracle forward + the host instantiation of the device adjoint) against the STOCK reference's eager
autograd.  Systems the stock reference cannot differentiate (Zernike: aten::floor_divide) are reported and skipped; parameters
whose reference gradient is NaN (odd asphere hit on its vertex) are skipped.

    python scripts/fuzz_live_gradients.py <first seed> <last seed>  (CPU only; summary: profiles/r2c_live_fuzz.txt)"""
import os, sys, traceback, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); warnings.filterwarnings("ignore")
import numpy as np
import tests.test_zzz_live_fuzz as F
from oracle.ref_import import import_reference
import_reference()
import torch
import optiland.backend as be
from optiland_b200 import plugin as P
from oracle.oracle_engine import OracleEngine
eng = OracleEngine()
be.set_backend("torch"); be.set_precision("float64")

def params_of(lens):
    ps = []
    for i, s in enumerate(lens.surfaces.surfaces[1:-1], start=1):
        g = s.geometry
        for nm in ("radius", "k", "Rx", "Ry", "kx", "ky"):
            v = getattr(g, nm, None)
            if torch.is_tensor(v) and v.dtype.is_floating_point and bool(torch.isfinite(v).all()):
                v = v.detach().clone().requires_grad_(True); setattr(g, nm, v); ps.append((i, nm, v))
        for nm in ("coefficients", "c"):
            v = getattr(g, nm, None)
            if torch.is_tensor(v) and v.dtype.is_floating_point and v.numel():
                v = v.detach().clone().requires_grad_(True); setattr(g, nm, v); ps.append((i, nm, v))
        cs = g.cs
        for nm in ("x", "y", "z", "rx", "ry"):
            v = getattr(cs, nm)
            if torch.is_tensor(v):
                v = v.detach().clone().requires_grad_(True); setattr(cs, nm, v); ps.append((i, "cs." + nm, v))
    return ps

def run(seed):
    lens, kinds = F._build(be, seed)
    ps = params_of(lens)
    with be.grad_mode.temporary_enable():
        loss = 0.0
        for hy in (0.0, 1.0):
            lens.trace(0.0, hy, 0.55, 6, "hexapolar")
            x, y, o, i = lens.surfaces.x[-1], lens.surfaces.y[-1], lens.surfaces.opd[-1], lens.surfaces.intensity[-1]
            m = torch.isfinite(x) & (i > 0)
            if m.sum() < 3: return None
            loss = loss + (x[m] ** 2).mean() + ((y[m] - y[m].mean()) ** 2).mean() + 1e-3 * ((o[m] - o[m].mean()) ** 2).mean()
        loss.backward()
    return float(loss.detach()), [(i, nm, None if v.grad is None else v.grad.detach().clone().numpy()) for i, nm, v in ps], kinds

lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0; served = 0; declined = {}
for seed in range(lo, hi):
    try:
        be.grad_mode.disable()
        try:
            ref = run(seed)
        except Exception as e:
            print(seed, "stock reference fails:", type(e).__name__, str(e)[:100]); continue
        if ref is None: continue
        P.install(engine=eng); P.stats(reset=True); n0 = len(eng.calls)
        got = run(seed)
        ng = sum(1 for c in eng.calls[n0:] if c[0] == "grad"); served += ng
        for k, v in P.stats().items(): declined[k] = declined.get(k, 0) + v
        nanref = False
        gscale = 1e-6 * max([float(np.max(np.abs(b))) for _, _, b in ref[1] if b is not None and np.all(np.isfinite(b))] + [1e-30])
        worst = abs(got[0] - ref[0]) / max(1e-30, abs(ref[0])); wk = "loss"
        for (i, nm, a), (_, _, b) in zip(got[1], ref[1]):
            if a is None or b is None:
                if not (a is None and b is None): print(seed, "GRAD NONE", i, nm, a is None, b is None)
                continue
            if not np.all(np.isfinite(b)):
                nanref = True; continue
            sc = max(gscale, float(np.max(np.abs(b))))
            e = float(np.max(np.abs(a - b))) / sc
            if not np.isfinite(e): e = 0.0 if np.array_equal(np.isnan(a), np.isnan(b)) else 1.0
            if e > worst: worst, wk = e, (i, nm)
        if worst > 5e-6:
            bad += 1; print(seed, "MISMATCH", f"{worst:.1e}", wk, got[2], "grad calls", ng, P.stats())
    except Exception as e:
        bad += 1; print(seed, "EXCEPTION", type(e).__name__, str(e)[:300]); traceback.print_exc(limit=4)
    finally:
        if P._state.get("installed"): P.uninstall()
print("bad", bad, "grad calls served", served, "declines", declined)
