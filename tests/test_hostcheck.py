"""CPU instantiation of the DEVICE arithmetic (olb_math.cuh + olb_prep.h compiled by g++,
tests/hostcheck/hostcheck.cpp) against the reference-generated golden fixtures.

This lets the GPU-less build container verify the kernel's math; the same comparisons
run on the B200 through the C ABI in tests/test_gpu_parity.py.  The host-check library
is test infrastructure only -- the product has no CPU path.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from optiland_b200 import _lib
from optiland_b200 import table as T
from tests._util import ERROR_CASES, REAL_CASES, REC, Case, max_abs_err

from oracle.hostcheck_api import SO, load, run_hostcheck  # noqa: E402,F401


@pytest.fixture(scope="module")
def hc():
    return load()


def newton_tol(c):
    tols = [s.tol for s in c.table.surfaces if s.kind in T.NEWTON_KINDS]
    return max(tols) if tols else 0.0


@pytest.mark.parametrize("name", REAL_CASES)
def test_device_math_f64_vs_reference(hc, name):
    c = Case(name)
    out, rec, status = run_hostcheck(hc, c.table, c.rays, np.float64, want_l0=True)
    assert status == 0
    # fp64 tolerance: 1e-11 x system scale (FMA contraction, flattened poses, stable conic
    # roots) + the reference's own Newton stopping residual where a Newton surface exists.
    tol = 1e-11 * c.scale + 2.0 * newton_tol(c)
    for k in REC:
        assert max_abs_err(rec[k], c.rec[k]) <= tol, (k, max_abs_err(rec[k], c.rec[k]), tol)
    for k in ("x", "y", "z", "L", "M", "N", "i", "opd", "L0", "M0", "N0"):
        assert max_abs_err(out[k], c.out[k]) <= tol, k


@pytest.mark.parametrize("name", REAL_CASES)
def test_device_math_f32_vs_reference(hc, name):
    """fp32 arithmetic vs the reference's fp64 records: every entry within 3x of the error this arithmetic achieves
    on the fixture (tests/golden/f32_achieved.json, scripts/f32_achieved.py) -- not a scale-based guess."""
    from tests._util import f32_bounds, fp32_errors

    c = Case(name)
    _, rec, status = run_hostcheck(hc, c.table, c.rays, np.float32)
    assert status == 0
    got = fp32_errors(rec, c.rec)
    bound = f32_bounds(name)
    for k, v in got.items():
        assert v <= 3.0 * bound[k] + 1e-9, (k, v, bound[k])


@pytest.mark.parametrize("name", ERROR_CASES)
def test_device_math_flags_zernike_range(hc, name):
    c = Case(name)
    _, _, status = run_hostcheck(hc, c.table, c.rays, np.float64)
    assert status & (T.ST_CHEBYSHEV_RANGE if "chebyshev" in name else T.ST_ZERNIKE_RANGE)


def test_partial_range_and_noop(hc):
    """Trace surfaces [3, 9) only (SurfaceGroup.trace(rays, skip=3) semantics + early stop)."""
    from oracle import trace_oracle as O

    c = Case("dgauss_c2")
    mid, _, _ = O.trace(c.table, c.rays, 0, 3)
    mid_in = {k: mid[k] for k in ("x", "y", "z", "L", "M", "N", "i", "w", "opd")}
    ref_out, ref_rec, _ = O.trace(c.table, mid_in, 3, 9)
    out, rec, _ = run_hostcheck(hc, c.table, mid_in, np.float64, first=3, last=9)
    for k in REC:
        assert rec[k].shape == (6, c.n)
        assert max_abs_err(rec[k], ref_rec[k]) <= 1e-11 * c.scale, k
    assert max_abs_err(out["opd"], ref_out["opd"]) <= 1e-11 * c.scale


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_degenerate_conic_roots(hc, dtype):
    """a == 0 (axial rays onto a paraboloid: the reference's `-c/b` branch, standard.py:144-146),
    a plane-like infinite radius StandardGeometry (N_safe guard :108-111) and rays that miss
    (NaN in band) all follow the oracle without special-case branches in the kernel."""
    from oracle import trace_oracle as O

    n = 64
    rng = np.random.default_rng(5)
    x = rng.uniform(-3, 3, n)
    y = rng.uniform(-3, 3, n)
    x[-4:] = 60.0  # misses the R = 40 sphere below -> NaN
    surfaces = [
        T.SurfaceSpec(kind=T.GEOM_STANDARD, radius=-50.0, conic=-1.0, reflective=True, t=[0, 0, 10.0]),
        T.SurfaceSpec(kind=T.GEOM_STANDARD, radius=float("inf"), t=[0, 0, 2.0], n1=[1.0], n2=[1.5]),
        T.SurfaceSpec(kind=T.GEOM_STANDARD, radius=40.0, conic=0.0, t=[0, 0, -5.0], n1=[1.5], n2=[1.0]),
    ]
    tab = T.SurfaceTable(surfaces, [0.55])
    rays = dict(x=x, y=y, z=np.zeros(n), L=np.zeros(n), M=np.zeros(n), N=np.ones(n), i=np.ones(n), w=np.full(n, 0.55))
    _, orec, _ = O.trace(tab, rays)
    _, rec, _ = run_hostcheck(hc, tab, rays, dtype)
    assert np.isnan(orec["x"][-1][-4:]).all() and np.isfinite(orec["x"][-1][:-4]).all()
    tol = 1e-12 * 60 if dtype == np.float64 else 5e-6 * 60  # steep rays at x = 60 mm amplify fp32 rounding
    for k in REC:
        assert max_abs_err(rec[k], orec[k]) <= tol, (k, max_abs_err(rec[k], orec[k]))


@pytest.mark.parametrize("name", ["zernike_polarized_c5", "cooke_polarized", "tilted_fold_polarized"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_device_math_polarized_vs_reference(hc, name, dtype):
    """P-matrix update (PolarizedRays.update) + Fresnel Jones matrices vs. the reference."""
    from tests._util import Case as _Case

    c = _Case(name)
    p0 = np.tile(np.eye(3, dtype=np.complex128), (c.n, 1, 1))
    out, rec, status = run_hostcheck(hc, c.table, c.rays, dtype, pmat=p0)
    assert status == 0
    from tests._util import f32_bounds

    b32 = f32_bounds(name)
    f64 = dtype == np.float64
    tol, ptol = 1e-11 * c.scale + 2 * newton_tol(c), (1e-11 if f64 else 3 * b32["p"])
    for k in ("x", "y", "opd"):
        assert max_abs_err(rec[k], c.rec[k]) <= (tol if f64 else 3 * b32["opd" if k == "opd" else "pos"]), k
    assert np.max(np.abs(out["p"] - c.out["p"])) <= ptol


@pytest.mark.parametrize("name", ["zernike_polarized_c5", "cooke_polarized", "tilted_fold_polarized"])
def test_device_math_polarized_intensity_epilogue(hc, name):
    """PolarizedRays.update_intensity as the kernels evaluate it (olb_math.cuh::polarized_intensity) against the
    reference's value (fixture) and the oracle, for the fixture's own state and for a second, elliptical one."""
    from oracle import trace_oracle as O
    from oracle.hostcheck_api import run_pol_intensity
    from tests._util import Case as _Case

    c = _Case(name)
    P, k0, i0 = c.out["p"], c.extra("k0"), c.extra("i0")
    state = tuple(c.extra("state")) if "x_state" in c.z else None
    want = c.extra("final_intensity") if state is not None else c.extra("final_intensity_unpolarized")
    got, st = run_pol_intensity(hc, P, k0, i0, state)
    assert st == 0 and np.max(np.abs(got - want)) <= 1e-13
    for other in ((0.3, 1.0, 0.7, -0.4), None):
        got, _ = run_pol_intensity(hc, P, k0, i0, other)
        mag = np.hypot(other[0], other[1]) if other else 1.0
        want = O.polarized_intensity(P, k0[0], k0[1], k0[2], i0, (other[0] / mag, other[1] / mag, other[2], other[3]) if other else None)
        assert np.max(np.abs(got - want)) <= 1e-13
    # a launch direction along x: the reference raises, the kernel sets a status bit
    kx = np.array([[1.0], [0.0], [0.0]])
    _, st = run_pol_intensity(hc, P[:1], kx, i0[:1], None)
    assert st & T.ST_K_PARALLEL_X
