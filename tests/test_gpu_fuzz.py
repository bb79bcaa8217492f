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
