"""The NumPy oracle against fixtures produced by the UNMODIFIED reference (tests/golden,
generator: oracle/make_golden.py) and against the reference's own hard-coded known-answer
vectors.  This is what pins the oracle (CPU only)."""
import numpy as np
import pytest

from oracle import trace_oracle as O
from optiland_b200 import table as T
from tests._util import ERROR_CASES, POLARIZED_CASES, REAL_CASES, REC, Case, max_abs_err


@pytest.mark.parametrize("name", REAL_CASES)
def test_oracle_matches_reference_records(name):
    c = Case(name)
    out, rec, status = O.trace(c.table, c.rays)
    assert status == 0
    # same operations in the same order as the reference => agreement to rounding.
    # (rotated poses are flattened to one matrix, so allow a few ulp of the system scale)
    tol = 1e-12 * c.scale if any(s.rotated for s in c.table.surfaces) else 0.0
    for k in REC:
        assert max_abs_err(rec[k], c.rec[k]) <= tol, k
    for k in ("x", "y", "z", "L", "M", "N", "i", "opd", "L0", "M0", "N0"):
        assert max_abs_err(out[k], c.out[k]) <= tol, k


@pytest.mark.parametrize("name", POLARIZED_CASES)
def test_oracle_matches_reference_polarized(name):
    c = Case(name)
    out, rec, status = O.trace(c.table, c.rays, polarized=True)
    for k in REC:
        assert max_abs_err(rec[k], c.rec[k]) <= 1e-13 * c.scale, k
    assert np.max(np.abs(out["p"] - c.out["p"])) < 1e-13
    k0 = c.extra("k0")
    if "x_state" in c.z:
        inten = O.polarized_intensity(out["p"], k0[0], k0[1], k0[2], c.extra("i0"), tuple(c.extra("state")))
        assert np.max(np.abs(inten - c.extra("final_intensity"))) < 1e-13
    elif "x_final_intensity_unpolarized" in c.z:
        inten = O.polarized_intensity(out["p"], k0[0], k0[1], k0[2], c.extra("i0"), None)
        assert np.max(np.abs(inten - c.extra("final_intensity_unpolarized"))) < 1e-13
    # (trace_generic-shaped fixtures hold no updated intensity: that call never runs update_intensity)


@pytest.mark.parametrize("name", ERROR_CASES)
def test_oracle_flags_zernike_range(name):
    c = Case(name)
    _, _, status = O.trace(c.table, c.rays)
    if "chebyshev" in name:
        assert "Chebyshev input coordinates must be normalized" in c.error
        assert status & T.ST_CHEBYSHEV_RANGE
    else:
        assert "Zernike coordinates must be normalized" in c.error
        assert status & T.ST_ZERNIKE_RANGE


def test_cooke_spot_radius_golden():
    """Config 1: RMS spot radius from the traced image-surface record must reproduce the
    reference's golden numbers (/root/reference/tests/test_analysis.py:88-102)."""
    c = Case("cooke_c1")
    _, rec, _ = O.trace(c.table, c.rays)
    npup = int(c.extra("n_pupil"))
    golden_055 = [0.004293689564257, 0.016918412809703662, 0.012116688566406967]  # test_analysis.py:93,97,101
    rms_ref = c.extra("spot_rms")
    for f in range(3):
        x = rec["x"][-1, f * npup:(f + 1) * npup]
        y = rec["y"][-1, f * npup:(f + 1) * npup]
        # SpotDiagram centres on the chief ray (pupil point 0 of hexapolar = centre)
        r2 = (x - x[0]) ** 2 + (y - y[0]) ** 2
        rms = np.sqrt(np.mean(r2))
        assert abs(rms - rms_ref[f][1]) < 1e-12
        assert abs(rms - golden_055[f]) < 5e-13


def _a(*v):
    return np.array(v, dtype=np.float64)


def test_known_answers_standard_geometry():
    """/root/reference/tests/test_geometries.py:182-222 (StandardGeometry distance & normal)."""
    t = O.conic_distance(_a(1.0, 2.0), _a(2.0, 3.0), _a(-3.0, -4.0), _a(0, 0), _a(0, 0), _a(1, 1), -12.0, 0.5)
    np.testing.assert_allclose(t, [2.7888809636986154, 3.4386378681404657], rtol=1e-13)
    L, M = 0.359, -0.229
    N = np.sqrt(1 - L**2 - M**2)
    t = O.conic_distance(_a(1.0), _a(2.0), _a(-10.2), _a(L), _a(M), _a(N), -12.0, 0.5)
    np.testing.assert_allclose(t, [10.201933401020467], rtol=1e-13)
    nx, ny, nz = O.conic_normal(_a(1.0), _a(2.0), 10.0, 0.5)
    np.testing.assert_allclose([nx[0], ny[0], nz[0]],
                               [0.10127393670836665, 0.2025478734167333, -0.9740215340114144], rtol=1e-13)


def test_known_answers_even_asphere():
    """/root/reference/tests/test_geometries.py:284-334 (EvenAsphere distance & normal)."""
    spec = T.SurfaceSpec(kind=T.GEOM_EVEN_ASPHERE, radius=-41.1, conic=0.0, coefficients=[1e-3, -1e-5, 1e-7],
                         tol=1e-10, max_iter=100)
    sag, normal = O._sag_and_normal_fns(spec, [0])
    t = O.newton_distance(_a(1.0, 2.0), _a(2.0, 3.0), _a(-3.0, -4.0), _a(0, 0), _a(0, 0), _a(1, 1), spec, sag, normal)
    np.testing.assert_allclose(t, [2.9438901710409624, 3.8530733934173256], rtol=1e-12)
    L, M = 0.222, -0.229
    N = np.sqrt(1 - L**2 - M**2)
    t = O.newton_distance(_a(1.0), _a(2.0), _a(-10.2), _a(L), _a(M), _a(N), spec, sag, normal)
    np.testing.assert_allclose(t, [10.625463223037386], rtol=1e-12)
    nx, ny, nz = O.even_normal(_a(1.0), _a(2.0), 10.0, 0.5, [1e-2])
    np.testing.assert_allclose([nx[0], ny[0], nz[0]],
                               [0.11946945186789681, 0.23893890373579363, -0.9636572265862595], rtol=1e-13)


def test_known_answers_reflect():
    """/root/reference/tests/test_rays.py:367-392 style: reflection off a plane flips N."""
    spec = T.SurfaceSpec(kind=T.GEOM_PLANE, reflective=True)
    tab = T.SurfaceTable([spec], [0.55])
    L, M = 0.1, -0.2
    N = np.sqrt(1 - L * L - M * M)
    rays = dict(x=_a(0.3), y=_a(-0.1), z=_a(-2.0), L=_a(L), M=_a(M), N=_a(N), i=_a(1.0), w=_a(0.55))
    out, rec, _ = O.trace(tab, rays)
    np.testing.assert_allclose([out["L"][0], out["M"][0], out["N"][0]], [L, M, -N], rtol=1e-15)
    np.testing.assert_allclose(out["z"], [0.0], atol=1e-16)
    np.testing.assert_allclose(out["opd"], [2.0 / N], rtol=1e-15)


def test_known_answers_p_matrix_update():
    """/root/reference/tests/test_rays.py:599-640 (PolarizedRays.update): identity Jones matrix with k0 == k1, with
    k0 tilted 0.1 in y against k1 = +z, and the jones=None form."""
    eye = np.eye(3)[None].astype(complex)
    one = np.array
    P = O.polarized_update(eye.copy(), one([0.0]), one([0.0]), one([1.0]), one([0.0]), one([0.0]), one([1.0]), eye.copy())
    np.testing.assert_allclose(P, eye, atol=1e-15)
    n0 = np.sqrt(1 - 0.1**2)
    want = np.array([[[1.0, 0.0, 0.0], [0.0, 0.99498744, -0.1], [0.0, 0.1, 0.99498744]]])
    P2 = O.polarized_update(P.copy(), one([0.0]), one([0.1]), one([n0]), one([0.0]), one([0.0]), one([1.0]), eye.copy())
    np.testing.assert_allclose(P2.real, want, atol=1e-8)
    np.testing.assert_allclose(P2.imag, 0, atol=1e-15)
    P3 = O.polarized_update(P2.copy(), one([0.0]), one([0.0]), one([1.0]), one([0.0]), one([0.0]), one([1.0]), None)
    np.testing.assert_allclose(P3.real, want, atol=1e-8)


def test_known_answers_fresnel_jones():
    """/root/reference/tests/test_jones.py:37-62 (JonesFresnel.calculate_matrix, n1 = 1.0, n2 = 1.5, aoi = 0.2)."""
    aoi = np.array([0.2])
    J = O.fresnel_jones(aoi, np.array([1.0]), np.array([1.5]), True, 1)
    np.testing.assert_allclose(J[0, 0, 0].real, -0.20541108217641596, rtol=1e-12)
    np.testing.assert_allclose(J[0, 1, 1].real, -0.19457669033430527, rtol=1e-12)
    np.testing.assert_allclose(J[0, 2, 2].real, -1.0)
    J = O.fresnel_jones(aoi, np.array([1.0]), np.array([1.5]), False, 1)
    np.testing.assert_allclose(J[0, 0, 0].real, 0.7945889178235841, rtol=1e-12)
    np.testing.assert_allclose(J[0, 1, 1].real, 0.7963844602228702, rtol=1e-12)
    np.testing.assert_allclose(J[0, 2, 2].real, 1.0)
