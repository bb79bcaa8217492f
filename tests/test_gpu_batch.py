"""batched many-systems traces.  Every system's segment of the one-launch batch must be
bit-identical to a single-system trace of the same perturbed table (same kernel, same prepared bytes),
and agree with the NumPy oracle on that perturbed table."""
import numpy as np
import pytest
import torch

from optiland_b200 import _lib
from optiland_b200 import table as T
from tests._util import REC, Case, max_abs_err

pytestmark = pytest.mark.gpu


def _np(t):
    return t.double().cpu().numpy()


def _perturbed(template, B, seed, tilt=True):
    """Monte-Carlo style perturbations: decenter, despace, tilt, curvature, conic, index, asphere terms."""
    from optiland_b200.batch import template_params

    rng = np.random.default_rng(seed)
    p0 = template_params(template)
    P = np.repeat(p0[None], B, axis=0)
    for b in range(1, B):                       # system 0 stays the nominal one
        for s, spec in enumerate(template.surfaces):
            if spec.kind == T.GEOM_NOOP or s == template.num_surfaces - 1:
                continue
            P[b, s, _lib.BP_TX:_lib.BP_TX + 3] += rng.normal(0, 0.02, 3)
            if tilt:
                R = T.rotation_matrix(*rng.normal(0, 2e-3, 3))
                P[b, s, _lib.BP_R:_lib.BP_R + 9] = (R @ p0[s, _lib.BP_R:_lib.BP_R + 9].reshape(3, 3)).reshape(9)
            if spec.kind != T.GEOM_PLANE:
                P[b, s, _lib.BP_CURV] *= 1 + rng.normal(0, 1e-3)
                P[b, s, _lib.BP_CONIC] += rng.normal(0, 1e-3)
            if spec.kind == T.GEOM_EVEN_ASPHERE:
                k = len(spec.coefficients)
                P[b, s, _lib.BP_COEF:_lib.BP_COEF + k] *= 1 + rng.normal(0, 1e-2, k)
        # a glass melt: every index of the system moves together, surface by surface consistently
        dn = rng.normal(0, 2e-4)
        for s in range(template.num_surfaces):
            for q in (_lib.BP_N1, _lib.BP_N2):
                if P[b, s, q] != 1.0:
                    P[b, s, q] += dn
    return P


def _one_wavelength(c):
    """Golden case restricted to its first wavelength (batched tables hold one)."""
    if c.table.n_wl == 1:
        return c.table, c.rays
    w0 = c.table.wavelengths[0]
    specs = []
    import dataclasses
    for s in c.table.surfaces:
        ch = {k: getattr(s, k)[:1].copy() for k in ("n1", "n2", "k1")}
        specs.append(dataclasses.replace(s, **ch))
    rays = dict(c.rays)
    rays["w"] = np.full_like(rays["w"], w0)
    return T.SurfaceTable(specs, np.array([w0])), rays


@pytest.mark.parametrize("name", ["cooke_c1", "aspheric_singlet", "tilted_fold", "hubble_c4"])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("m", [1000, 1023])
def test_batch_segments_equal_single_system_traces(name, dtype, m):
    from oracle import trace_oracle as O
    from optiland_b200.batch import BatchedTable, system_table, trace_batch
    from optiland_b200.trace import RealRays, SurfaceGroup

    c = Case(name)
    table, rays_np = _one_wavelength(c)
    B = 5
    P = _perturbed(table, B, seed=len(name) + m)
    rng = np.random.default_rng(m)
    idx = rng.integers(0, c.n, size=B * m)
    sub = {k: v[idx] for k, v in rays_np.items()}
    mk = lambda r: RealRays(r["x"], r["y"], r["z"], r["L"], r["M"], r["N"], r["i"], r["w"], dtype=dtype)

    bt = BatchedTable(table, P)
    recs, mom = trace_batch(bt, mk(sub), m, moments=True)
    S = table.num_surfaces
    assert recs["x"].shape == (S, B, m) and mom.shape == (B, 8)
    ntol = max([s.tol for s in table.surfaces if s.kind in T.NEWTON_KINDS] or [0.0])
    # same bars as test_gpu_parity: 1e-11 x scale (+ the reference's Newton stopping residual) in fp64
    tol = (1e-11 * c.scale if dtype == torch.float64 else 3e-5 * c.scale) + 2.0 * ntol
    for b in range(B):
        tb = system_table(table, P[b])
        seg = {k: v[b * m:(b + 1) * m] for k, v in sub.items()}
        sg = SurfaceGroup(tb)
        sg.trace(mk(seg))
        for k in REC:
            got, ref = recs[k][:, b], getattr(sg, k)
            if b == 0:   # the nominal system alone may select a leaner kernel variant (no rotations)
                assert max_abs_err(_np(got), _np(ref)) <= tol, (k, b)
            else:
                assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(ref, nan=-7.0)), (k, b)
        if b in (0, B - 1):
            _, orec, _ = O.trace(tb, seg)
            for k in REC:
                assert max_abs_err(_np(recs[k][:, b]), orec[k]) <= tol, (k, b)
        # per-system moments = moments of that system's image-surface record row
        x, y, i = (_np(recs[k][-1, b]) for k in ("x", "y", "intensity"))
        ok = np.isfinite(x) & np.isfinite(y) & (i > 0)
        mm = mom[b].cpu().numpy()
        assert mm[0] == ok.sum()
        if ok.any():
            rel = 1e-9 if dtype == torch.float64 else 1e-5
            assert abs(mm[1] - x[ok].sum()) <= rel * max(1.0, np.abs(x[ok]).sum())
            assert abs(mm[3] - (x[ok] ** 2 + y[ok] ** 2).sum()) <= rel * max(1.0, (x[ok] ** 2 + y[ok] ** 2).sum())


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_batch_shared_input_equals_replicated_rays(dtype):
    from optiland_b200.batch import BatchedTable, trace_batch
    from optiland_b200.trace import RealRays

    c = Case("cooke_c1")
    table, rays_np = _one_wavelength(c)
    B, m = 37, 2048
    P = _perturbed(table, B, seed=3)
    idx = np.random.default_rng(1).integers(0, c.n, size=m)
    one = {k: v[idx] for k, v in rays_np.items()}
    rep = {k: np.tile(v, B) for k, v in one.items()}
    mk = lambda r: RealRays(r["x"], r["y"], r["z"], r["L"], r["M"], r["N"], r["i"], r["w"], dtype=dtype)
    bt = BatchedTable(table, P)
    shared = mk(one)
    keep = shared.x.clone()
    r1, m1 = trace_batch(bt, shared, m, shared_input=True, moments=True)
    r2, m2 = trace_batch(bt, mk(rep), m, moments=True)
    assert torch.equal(shared.x, keep)                       # shared launch rays are only read
    for k in REC:
        assert torch.equal(torch.nan_to_num(r1[k], nan=-7.0), torch.nan_to_num(r2[k], nan=-7.0)), k
    assert torch.allclose(m1, m2, rtol=1e-12, atol=0)
    # moments only (no per-ray output at all)
    _, m3 = trace_batch(bt, shared, m, shared_input=True, record=False, moments=True)
    assert torch.allclose(m1, m3, rtol=1e-12, atol=0)
    # the nominal system differs from a perturbed one
    assert not torch.equal(r1["x"][-1, 0], r1["x"][-1, 1])


def test_batch_in_place_final_state_and_errors():
    from optiland_b200.batch import BatchedTable, system_table, trace_batch
    from optiland_b200.trace import DeviceTable, RealRays, trace_device

    c = Case("cooke_c1")
    table, rays_np = _one_wavelength(c)
    B, m = 3, 512
    P = _perturbed(table, B, seed=9)
    idx = np.arange(B * m) % c.n
    sub = {k: v[idx] for k, v in rays_np.items()}
    mk = lambda r: RealRays(r["x"], r["y"], r["z"], r["L"], r["M"], r["N"], r["i"], r["w"], dtype=torch.float64)
    bt = BatchedTable(table, P)
    rays = mk(sub)
    recs, mom = trace_batch(bt, rays, m, record=False)
    assert recs is None and mom is None
    for b in range(B):
        seg = mk({k: v[b * m:(b + 1) * m] for k, v in sub.items()})
        trace_device(DeviceTable(system_table(table, P[b])), seg, 0, table.num_surfaces, record=False)
        for k in ("x", "y", "L", "opd", "i"):
            assert torch.equal(torch.nan_to_num(getattr(rays, k)[b * m:(b + 1) * m], nan=-7.0),
                               torch.nan_to_num(getattr(seg, k), nan=-7.0)), (k, b)
    # a batched table refuses the single-system entry point, and a wrong ray count
    with pytest.raises(RuntimeError, match="several systems"):
        class _D:   # minimal DeviceTable look-alike
            pass
        d = _D()
        d.lib, d.c, d.device, d.table, d.has_zernike = bt.lib, bt.c, bt.device, table, False
        trace_device(d, mk(sub), 0, table.num_surfaces)
    with pytest.raises(ValueError):
        trace_batch(bt, mk(sub), m + 1)
    # structure-changing parameters are rejected at upload
    Pbad = P.copy()
    Pbad[1, 1, _lib.BP_R:_lib.BP_R + 9] = np.nan
    with pytest.raises(RuntimeError):
        BatchedTable(table, Pbad)
