"""Every sample system the reference ships (``optiland.samples``: 29 classes -- objectives, eyepieces, microscopes incl. a
reflecting one, the 44-surface lithography lens, infrared triplets, the Navarro eye, the Hubble, wide-angle lenses that need
the robust / iterative ray aimer): ``Optic.trace`` for every field at the primary wavelength under the plugin == the NumPy
reference on all eight record arrays of all surfaces, identical NaN patterns, no decline other than the aimer's.

``[oracle]``: CPU (``profiles/r2b_samples_sweep.txt``: worst relative difference 5e-15).  ``[cuda]``: the product engine; written
after the round's GPU budget was spent, hence in a last-sorted file."""
import importlib
import inspect
import pkgutil

import numpy as np
import pytest

from oracle.ref_import import reference_available
from tests.test_plugin_reference import plugin  # noqa: F401  (fixture)

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference not present on this box")


def _sample_classes():
    if not reference_available():
        return []
    from oracle.ref_import import import_reference

    import_reference()
    import optiland.samples as S
    from optiland.optic import Optic

    out = []
    for m in pkgutil.iter_modules(S.__path__):
        mod = importlib.import_module("optiland.samples." + m.name)
        for name, obj in inspect.getmembers(mod, inspect.isclass):
            if issubclass(obj, Optic) and obj is not Optic and obj.__module__ == mod.__name__:
                out.append(f"{m.name}.{name}")
    return sorted(out)


@pytest.mark.parametrize("qualname", _sample_classes())
def test_sample_system_traces_like_the_numpy_reference(plugin, qualname):
    P, eng, be = plugin
    modname, name = qualname.split(".")
    cls = getattr(importlib.import_module("optiland.samples." + modname), name)
    be.set_backend("numpy")
    ref = cls()
    fields = [tuple(float(v) for v in f) for f in ref.fields.get_field_coords()]
    wl = ref.primary_wavelength
    want = []
    for hx, hy in fields:
        ref.trace(hx, hy, wl, 4, "hexapolar")
        want.append({k: np.array(getattr(ref.surfaces, k)) for k in ("x", "y", "z", "L", "M", "N", "opd", "intensity")})
    be.set_backend("torch")
    P.stats(reset=True)
    n0 = len(eng.calls)
    lens = cls()
    for (hx, hy), w in zip(fields, want):
        lens.trace(hx, hy, wl, 4, "hexapolar")
        scale = max(1.0, float(np.nanmax(np.abs(np.where(np.isfinite(w["z"]), w["z"], 0.0)))))
        for k, v in w.items():
            g = be.to_numpy(getattr(lens.surfaces, k))
            assert g.shape == v.shape, (k, g.shape, v.shape)
            assert np.array_equal(np.isnan(g), np.isnan(v)), (qualname, k, "NaN pattern")
            m = np.isfinite(v)
            assert not m.any() or float(np.max(np.abs(g[m] - v[m]))) <= 1e-10 * scale, (qualname, k)
    assert len(eng.calls) - n0 >= len(fields)
    assert set(P.stats()) <= {"fused launch: non-paraxial ray aiming"}, P.stats()
