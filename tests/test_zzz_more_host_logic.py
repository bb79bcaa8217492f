"""Host logic written in the LAST session of round 2 (no GPU minutes left): Python in front of hardware-verified kernels; the
``[cuda]`` variants have not run on a B200, so the file sorts behind every other test (the driver runs ``pytest -x``)."""
import numpy as np
import pytest

from oracle.ref_import import reference_available
from tests.test_plugin_reference import _numpy_reference, plugin  # noqa: F401  (fixture: [oracle] on CPU, [cuda] on the B200)

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference not present on this box")


def test_trace_generic_with_apodization_and_vignetting_factors(plugin):
    """``trace_generic`` on an apodized optic whose fields carry vignetting factors: the reference evaluates the apodization
    on the pupil point scaled ONCE by (1 - v) (real_ray_tracer.py:132-141 -> ray_generator.py:83-85) while its aimer scales
    the launch geometry a second time; the fused launch does the latter in the kernel and the former as three eager ops."""
    from optiland.apodization import GaussianApodization
    from optiland.samples.objectives import CookeTriplet

    P, eng, be = plugin

    def make():
        lens = CookeTriplet()
        for f, (vx, vy) in zip(lens.fields.fields, ((0.0, 0.0), (0.1, 0.2), (0.15, 0.3))):
            f.vx, f.vy = vx, vy
        lens.set_apodization(GaussianApodization(sigma=0.8))
        return lens

    rng = np.random.default_rng(3)
    n = 64
    arrs = [rng.uniform(-0.3, 0.3, n), rng.uniform(-1, 1, n), rng.uniform(-0.7, 0.7, n), rng.uniform(-0.7, 0.7, n)]
    want, fin = _numpy_reference(make, lambda lens: lens.trace_generic(*[be.array(a) for a in arrs], 0.55))
    lens = make()
    P.stats(reset=True)
    n0 = len(eng.calls)
    rays = lens.trace_generic(*[be.array(a) for a in arrs], 0.55)
    assert [c[0] for c in eng.calls[n0:]] == ["pupil"] and not P.stats(), (eng.calls[n0:], P.stats())
    for k, v in want.items():
        np.testing.assert_allclose(be.to_numpy(getattr(lens.surfaces, k)), v, rtol=0, atol=1e-10, err_msg=k)
    np.testing.assert_allclose(be.to_numpy(rays.i), fin["i"], rtol=0, atol=1e-13)
    assert float(np.ptp(fin["i"])) > 0.05          # the apodization really varies over the sample


def test_spot_statistics_of_an_apodized_pupil_from_moment_launches(plugin):
    """``SpotDiagram`` on an apodized (unpolarized) optic: the apodization only scales the intensity and the statistics mask
    ``i > 0``, so ``rms_spot_radius`` / ``centroid`` come from the moment launches of the unit-intensity pupil; the lazily
    materialised per-ray data (``geometric_spot_radius``, ``data[..].intensity``) carry the apodization factor."""
    from optiland.analysis import SpotDiagram
    from optiland.apodization import GaussianApodization
    from optiland.samples.objectives import CookeTriplet

    P, eng, be = plugin

    def make():
        lens = CookeTriplet()
        lens.set_apodization(GaussianApodization(sigma=0.7))
        return lens

    def numbers(lens):
        sd = SpotDiagram(lens, num_rings=4)
        rms = [[float(np.asarray(be.to_numpy(v)).reshape(-1)[0]) for v in row] for row in sd.rms_spot_radius()]
        cen = [[float(np.asarray(be.to_numpy(c)).reshape(-1)[0]) for c in pair] for pair in sd.centroid()]
        calls_before_geo = len(eng.calls)
        geo = [[float(np.asarray(be.to_numpy(v)).reshape(-1)[0]) for v in row] for row in sd.geometric_spot_radius()]
        inten = np.asarray(be.to_numpy(sd.data[1][0].intensity))
        return np.array(rms), np.array(cen), np.array(geo), inten, calls_before_geo

    be.set_backend("numpy")
    want = numbers(make())
    be.set_backend("torch")
    P.stats(reset=True)
    n0 = len(eng.calls)
    got = numbers(make())
    before = eng.calls[n0:got[4]]
    # statistics: moment launches only -- the one-ray "pupil" launches are the chief-ray reference's own traces
    assert any(c[0] == "moments" for c in before) and all(c[0] == "moments" or (c[0] == "pupil" and c[2] == 1) for c in before), before
    assert not P.stats(), P.stats()
    assert any(c[0] == "pupil" and c[2] > 1 for c in eng.calls[got[4]:])              # the per-ray data: traced when read
    for a, b, what in zip(got[:4], want[:4], ("rms", "centroid", "geometric", "intensity")):
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12, err_msg=what)
    assert float(np.ptp(want[3])) > 0.1
