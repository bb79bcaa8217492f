"""The randomised systems of tests/test_fuzz_hostcheck.py through the sm_100a kernel (C ABI) against the oracle,
and the backward kernel against finite differences of the forward kernel on random systems."""
import numpy as np
import pytest
import torch

from tests._util import REC
from tests.test_fuzz_hostcheck import random_system

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(16))
def test_random_systems_kernel_vs_oracle(seed):
    from oracle import trace_oracle as O
    from optiland_b200.trace import RealRays, SurfaceGroup

    rng = np.random.default_rng(1000 + seed)
    table = random_system(rng, int(rng.integers(4, 10)))
    n = 1000 + seed          # ragged sizes: scalar and vector paths
    x, y = rng.uniform(-5, 5, n), rng.uniform(-5, 5, n)
    L, M = rng.normal(0, 0.05, n), rng.normal(0, 0.05, n)
    N = np.sqrt(1 - L**2 - M**2)
    rays = dict(x=x, y=y, z=np.full(n, -5.0), L=L, M=M, N=N, i=np.ones(n), w=rng.choice(table.wavelengths, n))
    _, orec, _ = O.trace(table, rays)
    scale = max(1.0, float(np.nanmax(np.abs(np.where(np.isfinite(orec["z"]), orec["z"], 0)))))
    sg = SurfaceGroup(table)
    sg.trace(RealRays(x, y, rays["z"], L, M, N, rays["i"], rays["w"], dtype=torch.float64))
    for k in REC:
        a, b = getattr(sg, k).cpu().numpy(), orec[k]
        assert np.array_equal(np.isnan(a), np.isnan(b)), (seed, k)
        m = np.isfinite(b)
        if m.any():
            assert np.max(np.abs(a[m] - b[m])) <= 1e-11 * scale, (seed, k, float(np.max(np.abs(a[m] - b[m]))))
    sg32 = SurfaceGroup(table)
    sg32.trace(RealRays(x, y, rays["z"], L, M, N, rays["i"], rays["w"], dtype=torch.float32))
    a, b = sg32.x[-1].double().cpu().numpy(), orec["x"][-1]
    fin = np.isfinite(a) & np.isfinite(b)
    if fin.sum() > n // 2:
        assert np.percentile(np.abs(a[fin] - b[fin]), 90) <= 2e-4 * scale


@pytest.mark.parametrize("seed", range(6))
def test_random_systems_backward_kernel_vs_finite_differences(seed):
    """Adjoint kernel on random plane / conic / even- and odd-asphere systems with tilts, decenters, mirrors, aperture
    trees and simple coatings: parameter gradients of a random linear functional of the image-surface records
    against central differences of the forward kernel (fp64)."""
    from optiland_b200 import autograd as AG
    from optiland_b200 import table as T
    from optiland_b200.trace import RealRays

    rng = np.random.default_rng(5000 + seed)
    full = random_system(rng, int(rng.integers(4, 8)))
    specs = []
    for s in full.surfaces:      # one wavelength, kinds inside the adjoint's scope
        import dataclasses
        ch = {k: getattr(s, k)[:1].copy() for k in ("n1", "n2", "k1")}
        if s.kind == T.GEOM_POLYNOMIAL:
            ch.update(kind=T.GEOM_STANDARD, coefficients=np.zeros(0))
        if s.kind in (T.GEOM_EVEN_ASPHERE, T.GEOM_ODD_ASPHERE):
            ch["tol"] = 1e-14
        specs.append(dataclasses.replace(s, **ch))
    table = T.SurfaceTable(specs, full.wavelengths[:1])
    n = 512
    x, y = rng.uniform(-3, 3, n), rng.uniform(-3, 3, n)
    L, M = rng.normal(0, 0.03, n), rng.normal(0, 0.03, n)
    N = np.sqrt(1 - L**2 - M**2)
    wts = {k: torch.from_numpy(rng.normal(size=n)).cuda() for k in ("x", "y", "L", "opd")}

    def loss(params):
        rr = RealRays(x, y, np.full(n, -5.0), L, M, N, np.ones(n), np.full(n, table.wavelengths[0]), dtype=torch.float64)
        rec = AG.trace_differentiable(table, params, rr, rows=(-1,))
        tot = 0.0
        for k, w in wts.items():
            v = rec[k]
            tot = tot + (torch.where(torch.isfinite(v), v, torch.zeros_like(v)) * w).sum()
        return tot

    p0 = AG.table_to_params(table)
    pr = p0.clone().requires_grad_(True)
    loss(pr).backward()
    g = pr.grad
    gmax = float(g.abs().max())
    checked = 0
    for s_, spec in enumerate(table.surfaces):
        if spec.kind == T.GEOM_NOOP:
            continue
        slots = [AG.GP_TZ, AG.GP_TX]
        if spec.kind != T.GEOM_PLANE:
            slots += [AG.GP_CURV, AG.GP_CONIC]
        if spec.kind in (T.GEOM_EVEN_ASPHERE, T.GEOM_ODD_ASPHERE):
            slots.append(AG.GP_COEF + 1)
        if spec.rotated:
            slots.append(AG.GP_R + 1)
        for q in slots:
            h = 1e-6 * (abs(float(p0[s_, q])) + 1.0) if q not in (AG.GP_CURV,) else 1e-6 * abs(float(p0[s_, q]))
            pa, pb = p0.clone(), p0.clone()
            pa[s_, q] += h
            pb[s_, q] -= h
            fd = (float(loss(pa)) - float(loss(pb))) / (2 * h)
            assert g[s_, q].item() == pytest.approx(fd, rel=5e-4, abs=2e-6 * gmax), (seed, s_, q, g[s_, q].item(), fd)
            checked += 1
    assert checked >= 6


@pytest.mark.parametrize("seed", range(12))
def test_random_freeform_systems_kernel_vs_oracle(seed):
    """Random Chebyshev / biconic / toroidal / Forbes / Zernike / polynomial systems through the kernel."""
    from oracle import trace_oracle as O
    from optiland_b200.trace import RealRays, SurfaceGroup
    from tests.test_fuzz_hostcheck import random_freeform_system

    rng = np.random.default_rng(3000 + seed)
    table = random_freeform_system(rng)
    n = 777
    x, y = rng.uniform(-4, 4, n), rng.uniform(-4, 4, n)
    L, M = rng.normal(0, 0.03, n), rng.normal(0, 0.03, n)
    N = np.sqrt(1 - L**2 - M**2)
    rays = dict(x=x, y=y, z=np.full(n, -3.0), L=L, M=M, N=N, i=np.ones(n), w=np.full(n, 0.55))
    _, orec, _ = O.trace(table, rays)
    scale = max(1.0, float(np.nanmax(np.abs(np.where(np.isfinite(orec["z"]), orec["z"], 0)))))
    sg = SurfaceGroup(table)
    sg.trace(RealRays(x, y, rays["z"], L, M, N, rays["i"], rays["w"], dtype=torch.float64))
    for k in REC:
        a, b = getattr(sg, k).cpu().numpy(), orec[k]
        assert np.array_equal(np.isnan(a), np.isnan(b)), (seed, k)
        m = np.isfinite(b)
        assert np.max(np.abs(a[m] - b[m])) <= 1e-11 * scale + 1e-10, (seed, k, float(np.max(np.abs(a[m] - b[m]))))
