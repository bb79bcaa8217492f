"""Host-logic additions of the second round-2 session, each against the unmodified reference (NumPy backend / its own eager
autograd): apodized pupils through the fused launches and the wavefront epilogue, size-1 arrays in ``trace_generic``, nested
coordinate frames (forward + gradients), systems imported from the reference's Zemax test files.

They add no device code -- they are Python in front of kernels that tests/test_plugin_reference.py and tests/test_gpu_*.py
verify on the B200 -- and were written after that session's GPU budget had been spent, so their ``[cuda]`` variants have not
run on hardware yet (their ``[oracle]`` twins are green).  The file sorts behind every hardware-verified test for that
reason."""
import numpy as np
import pytest

from oracle.ref_import import reference_available
from tests.test_plugin_reference import _numpy_reference, plugin  # noqa: F401  (fixture: [oracle] on CPU, [cuda] on the B200)

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference not present on this box")


def test_apodized_pupil_goes_through_the_fused_launch(plugin):
    """An apodized pupil (``optic.apodization``: the ray generator launches with intensity
    ``apodization.get_intensity(Px, Py)``, rays/ray_generator.py:83-87).  The kernel launches with unit intensity and every
    operation on the intensity along the path is a product or a reset to zero, so the plugin scales the intensity records
    by the per-ray factor afterwards -- ``Optic.trace`` and ``trace_generic`` stay ONE fused launch, unpolarized and
    polarized (Fresnel coatings: the factor also multiplies update_intensity's result), and equal the NumPy reference."""
    P, eng, be = plugin
    from optiland.apodization import GaussianApodization
    from optiland.samples.objectives import DoubleGauss

    rng = np.random.default_rng(8)
    n = 40
    Pxs, Pys = rng.uniform(-0.7, 0.7, n), rng.uniform(-0.7, 0.7, n)
    Hys = rng.uniform(0.0, 1.0, n)

    def make(polarized):
        def f():
            lens = DoubleGauss()
            lens.set_apodization(GaussianApodization(sigma=0.6))
            lens.surfaces.surfaces[3].aperture = None
            if polarized:
                from optiland.rays import PolarizationState

                lens.surfaces.set_fresnel_coatings()
                lens.set_polarization(PolarizationState(is_polarized=False))
            return lens
        return f

    for polarized in (False, True):
        for trace in (lambda lens: lens.trace(0.0, 0.7, 0.5876, 6, "hexapolar"),
                      lambda lens: lens.trace_generic(be.array(np.zeros(n)), be.array(Hys), be.array(Pxs), be.array(Pys), 0.5876)):
            ref_rec, ref_fin = _numpy_reference(make(polarized), trace)
            assert ref_rec["intensity"][0].min() < 0.9 * ref_rec["intensity"][0].max()      # the pupil really is apodized
            lens = make(polarized)()
            n0 = len(eng.calls)
            P.stats(reset=True)
            rays = trace(lens)
            assert any(c[0] == "pupil" for c in eng.calls[n0:]) and not P.stats(), (eng.calls[n0:], P.stats())
            for k, v in ref_rec.items():
                np.testing.assert_allclose(be.to_numpy(getattr(lens.surfaces, k)), v, rtol=0, atol=1e-10, err_msg=k)
            np.testing.assert_allclose(be.to_numpy(rays.i), ref_fin["i"], rtol=0, atol=1e-11)


def test_wavefront_epilogue_with_an_apodized_pupil(plugin):
    """Wavefront(strategy='chief_ray') of an apodized system through the fused epilogue: same OPD and pupil intercepts, the
    intensity output scaled by the per-ray factor; equal to the NumPy reference."""
    P, eng, be = plugin
    from optiland.apodization import GaussianApodization
    from optiland.samples.objectives import CookeTriplet
    from optiland.wavefront import Wavefront

    def make():
        lens = CookeTriplet()
        lens.set_apodization(GaussianApodization(sigma=0.7))
        return lens

    field, wl = (0.0, 1.0), 0.55

    def run(lens):
        w = Wavefront(lens, fields=[field], wavelengths=[wl], num_rays=8, distribution="hexapolar", strategy="chief_ray")
        d = w.get_data(field, wl)
        return {k: np.array(be.to_numpy(getattr(d, k)), dtype=np.float64) for k in ("opd", "pupil_x", "pupil_y", "pupil_z", "intensity")}

    be.set_backend("numpy")
    want = run(make())
    be.set_backend("torch")
    n0 = len(eng.calls)
    P.stats(reset=True)
    got = run(make())
    assert any(c[0] == "wavefront" for c in eng.calls[n0:]) and not P.stats(), (eng.calls[n0:], P.stats())
    assert want["intensity"].min() < 0.9 * want["intensity"].max()
    for k in want:
        np.testing.assert_allclose(got[k], want[k], rtol=0, atol=1e-6 if k == "opd" else 1e-10, err_msg=k)


def test_trace_generic_broadcasts_size_one_arrays(plugin):
    """Python floats, 0-d and 1-element arrays beside (n,) arrays in ``trace_generic``: the reference's element-wise ops
    broadcast them (real_ray_tracer.py:175-194 expands Python numbers only); the fused launch brings them to (n,)."""
    P, eng, be = plugin
    from optiland.samples.objectives import DoubleGauss

    rng = np.random.default_rng(3)
    n = 60
    Px = rng.uniform(-0.7, 0.7, n)

    def trace_mixed(lens):
        return lens.trace_generic(0.0, be.array(0.7), be.array(Px), be.array([0.25]), 0.5876)

    ref_rec, ref_fin = _numpy_reference(DoubleGauss, trace_mixed)
    lens = DoubleGauss()
    n0 = len(eng.calls)
    P.stats(reset=True)
    rays = trace_mixed(lens)
    assert ("pupil", 13, n) in [c[:3] for c in eng.calls[n0:]] and not P.stats(), (eng.calls[n0:], P.stats())
    for k, v in ref_rec.items():
        np.testing.assert_allclose(be.to_numpy(getattr(lens.surfaces, k)), v, rtol=0, atol=1e-10, err_msg=k)
    for k, v in ref_fin.items():
        np.testing.assert_allclose(be.to_numpy(getattr(rays, k)), v, rtol=0, atol=1e-10, err_msg=k)


def test_nested_coordinate_frames_forward_and_autograd(plugin):
    """Frames defined relative to another frame (``CoordinateSystem.reference_cs``: the coordinate breaks of imported
    systems, fileio/zemax/reader/converter.py:120-190).  Forward: the packer flattens the chain
    (``get_effective_transform``, coordinate_system.py:145-165) and the records equal the NumPy reference.  Gradients
    (be.grad_mode): the effective pose t = t_p + R_p t_c, R = R_p R_c is composed from the LIVE tensors of every level
    (``plugin._live_frame``), so d(RMS spot)/d(parent tilt, parent decenter, child tilt, child z) through the adjoint
    kernel's dLoss/dt and dLoss/dR equal the reference's own eager autograd."""
    import torch

    P, eng, be = plugin
    from optiland import optic as _optic
    from optiland.coordinate_system import CoordinateSystem

    def make():
        lens = _optic.Optic()
        lens.surfaces.add(index=0, radius=be.inf, thickness=be.inf)
        lens.surfaces.add(index=1, radius=40.0, thickness=5.0, material="N-BK7", is_stop=True)
        lens.surfaces.add(index=2, radius=-55.0, thickness=3.0, conic=-0.5)
        lens.surfaces.add(index=3, radius=-30.0, thickness=4.0, material="SF5")
        lens.surfaces.add(index=4, radius=-80.0, thickness=45.0)
        lens.surfaces.add(index=5)
        lens.set_aperture(aperture_type="EPD", value=9.0)
        lens.fields.set_type(field_type="angle")
        lens.fields.add(y=0)
        lens.fields.add(y=3)
        lens.wavelengths.add(value=0.6, is_primary=True)
        # surfaces 3 and 4 live in a tilted / decentered carrier frame; surface 4 is tilted once more inside it
        carrier = CoordinateSystem(x=0.15, y=-0.1, z=8.0, rx=0.02, ry=-0.015, rz=0.3)
        lens.surfaces.surfaces[3].geometry.cs = CoordinateSystem(x=0.0, y=0.0, z=0.0, reference_cs=carrier)
        lens.surfaces.surfaces[4].geometry.cs = CoordinateSystem(x=-0.05, y=0.02, z=4.0, rx=-0.01, ry=0.025, rz=0.0,
                                                               reference_cs=carrier)
        return lens, carrier

    def trace(lens):
        return lens.trace(0.0, 1.0, 0.6, 6, "hexapolar")

    ref_rec, ref_fin = _numpy_reference(lambda: make()[0], trace)
    lens, _ = make()
    n0 = len(eng.calls)
    P.stats(reset=True)
    rays = trace(lens)
    assert len(eng.calls) > n0 and not P.stats(), P.stats()
    for k, v in ref_rec.items():
        np.testing.assert_allclose(be.to_numpy(getattr(lens.surfaces, k)), v, rtol=0, atol=1e-10, err_msg=k)
    np.testing.assert_allclose(be.to_numpy(rays.opd), ref_fin["opd"], rtol=0, atol=1e-10)

    def run():
        lens, carrier = make()
        c4 = lens.surfaces.surfaces[4].geometry.cs
        trace(lens)
        x, y = lens.surfaces.x[-1, :], lens.surfaces.y[-1, :]
        loss = torch.sqrt(torch.mean((x - torch.mean(x)) ** 2 + (y - torch.mean(y)) ** 2)) + 1e-3 * torch.mean(lens.surfaces.opd[-1, :])
        loss.backward()
        out = {"loss": float(loss.detach())}
        for name, t in (("carrier.x", carrier.x), ("carrier.z", carrier.z), ("carrier.rx", carrier.rx), ("carrier.ry", carrier.ry),
                        ("carrier.rz", carrier.rz), ("c4.x", c4.x), ("c4.z", c4.z), ("c4.rx", c4.rx), ("c4.ry", c4.ry),
                        ("radius4", lens.surfaces.surfaces[4].geometry.radius)):
            out[name] = float(t.grad)
        return out

    be.grad_mode.enable()
    try:
        n1 = len(eng.calls)
        P.stats(reset=True)
        got = run()
        assert any(c[0] == "grad" for c in eng.calls[n1:]) and not P.stats(), (eng.calls[n1:], P.stats())
        P.uninstall()                       # the reference's own eager graph
        ref = run()
    finally:
        be.grad_mode.disable()
    assert got["loss"] == pytest.approx(ref["loss"], rel=1e-10)
    scale = max(abs(v) for k, v in ref.items() if k != "loss")
    for k in ref:
        assert got[k] == pytest.approx(ref[k], rel=5e-6, abs=1e-9 * scale), (k, got[k], ref[k])


@pytest.mark.parametrize("name", ["simple_fold_mirror_up.zmx", "one_mirror_up_45deg.zmx", "complicated_fold_mirrors_setup_v2.zmx",
                                  "lens1.zmx", "thorlabs_lj1598l1.zmx"])
def test_imported_zemax_systems(plugin, name):
    """Systems read from the reference's own Zemax test files (fileio/zemax: coordinate breaks turned into tilted /
    decentered frames, fold mirrors, a multi-field lens, a toroidal cylinder lens): ``Optic.trace`` for every field under
    the plugin == the NumPy reference, and the trace went through the capability without a decline."""
    import os

    P, eng, be = plugin
    from optiland.fileio import load_zemax_file

    from oracle.ref_import import REFERENCE_TESTS

    path = os.path.join(REFERENCE_TESTS, "zemax_files", name)

    def fields(lens):
        return [tuple(float(v) for v in f) for f in lens.fields.get_field_coords()]

    be.set_backend("numpy")
    ref_lens = load_zemax_file(path)
    want = []
    for hx, hy in fields(ref_lens):
        ref_lens.trace(hx, hy, ref_lens.primary_wavelength, 5, "hexapolar")
        want.append({k: np.array(getattr(ref_lens.surfaces, k)) for k in ("x", "y", "z", "L", "M", "N", "opd", "intensity")})
    be.set_backend("torch")
    lens = load_zemax_file(path)
    P.stats(reset=True)
    n0 = len(eng.calls)
    for (hx, hy), w in zip(fields(lens), want):
        lens.trace(hx, hy, lens.primary_wavelength, 5, "hexapolar")
        scale = max(1.0, float(np.nanmax(np.abs(np.where(np.isfinite(w["z"]), w["z"], 0.0)))))
        for k, v in w.items():
            np.testing.assert_allclose(be.to_numpy(getattr(lens.surfaces, k)), v, rtol=0, atol=1e-11 * scale + 1e-10, err_msg=f"{name} {k}")
    assert len(eng.calls) - n0 == len(want) and not P.stats(), (eng.calls[n0:], P.stats())


FIELD_TYPES = ("angle", "object_height", "paraxial_image_height", "real_image_height")
APERTURES = {"EPD": 8.0, "imageFNO": 6.0, "objectNA": 0.06, "float_by_stop_size": 3.5}
DISTRIBUTIONS = (("hexapolar", 4), ("uniform", 7), ("line_x", 9), ("line_y", 9), ("positive_line_x", 5),
                 ("positive_line_y", 5), ("cross", 7), ("ring", 8))


@pytest.mark.parametrize("finite", [False, True], ids=["infinite_object", "finite_object"])
@pytest.mark.parametrize("field_type", FIELD_TYPES)
def test_every_field_type_aperture_type_and_pupil_distribution(plugin, finite, field_type):
    """The launch side of ``Optic.trace`` across the reference's options: 2 conjugates x 4 field types
    (fields/field_types/*.py) x 4 system-aperture types (aperture/*.py) x 8 deterministic pupil distributions
    (distribution.py:415-441; 'random' / 'sobol' draw from different generators on the NumPy and torch backends) on a
    doublet -- records under the plugin == the NumPy reference, all through the in-kernel launch generation: angle and object-height
    fields in closed form, the image-height field types with the object angle / height taken from ONE probe of the
    reference's own solve (pack.launch_scalars, mode 3).  The one degenerate combination (object-space NA with the
    object at infinity) declines the fused launch and runs on the SurfaceGroup capability."""
    P, eng, be = plugin
    from optiland import optic as _optic

    if field_type == "object_height" and not finite:
        pytest.skip("the reference raises: an object-height field needs a finite object")

    def build(ap_type):
        lens = _optic.Optic()
        lens.surfaces.add(index=0, radius=be.inf, thickness=(60.0 if finite else be.inf))
        lens.surfaces.add(index=1, radius=32.0, thickness=5.0, material="N-BK7")
        lens.surfaces.add(index=2, radius=-48.0, thickness=2.0, is_stop=True)
        lens.surfaces.add(index=3, radius=-30.0, thickness=2.5, material="SF5")
        lens.surfaces.add(index=4, radius=-90.0, thickness=55.0)
        lens.surfaces.add(index=5)
        lens.set_aperture(aperture_type=ap_type, value=APERTURES[ap_type])
        lens.fields.set_type(field_type=field_type)
        lens.fields.add(y=0.0)
        lens.fields.add(y={"angle": 4.0, "object_height": 3.0, "paraxial_image_height": 2.5, "real_image_height": 2.5}[field_type])
        lens.wavelengths.add(value=0.55, is_primary=True)
        return lens

    keys = ("x", "y", "z", "L", "M", "N", "opd", "intensity")
    for ap_type in APERTURES:
        be.set_backend("numpy")
        ref = build(ap_type)
        want = {}
        for dn, nr in DISTRIBUTIONS:
            ref.trace(0.0, 1.0, 0.55, nr, dn)
            want[dn] = {k: np.array(getattr(ref.surfaces, k)) for k in keys}
        be.set_backend("torch")
        lens = build(ap_type)
        P.stats(reset=True)
        n0 = len(eng.calls)
        for dn, nr in DISTRIBUTIONS:
            lens.trace(0.0, 1.0, 0.55, nr, dn)
            for k, v in want[dn].items():
                g = be.to_numpy(getattr(lens.surfaces, k))
                assert g.shape == v.shape and np.array_equal(np.isnan(g), np.isnan(v)), (ap_type, dn, k)
                np.testing.assert_allclose(g, v, rtol=0, atol=1e-10, err_msg=f"{ap_type} {dn} {k}")
        fused = [c for c in eng.calls[n0:] if c[0] == "pupil"]
        degenerate = ap_type == "objectNA" and not finite and field_type.endswith("image_height")
        if not degenerate:
            assert len(fused) == len(DISTRIBUTIONS) and not P.stats(), (ap_type, P.stats())
        else:
            # an object-space NA with the object at infinity: the probed origin is not the expected function of the pupil
            # point -> the launch scalars decline and the SurfaceGroup capability carries the trace
            assert len(eng.calls) - n0 >= len(DISTRIBUTIONS)
            assert all(k.startswith("fused launch: unsupported: launch_scalars: field type") for k in P.stats()), P.stats()


def test_catalogue_glass_indices_are_memoised_and_other_materials_are_not():
    """``pack.catalogue_value``: n / k of a catalogue glass (``Material`` / ``MaterialFile``: functions of the data file
    alone) are remembered on the material object, so that packing under ``be.grad_mode`` -- where the reference's own
    per-wavelength cache is bypassed (materials/base.py:118-121) -- does not re-evaluate a dispersion formula per surface
    side per call; the values are those a fresh evaluation gives; ``IdealMaterial`` / ``AbbeMaterial`` (parameters an
    optimiser may drive) are asked every time."""
    from oracle.ref_import import import_reference

    import_reference()
    import optiland.backend as be
    from optiland.materials import AbbeMaterial, IdealMaterial, Material
    from optiland.samples.objectives import CookeTriplet

    from optiland_b200 import pack as PK

    be.set_backend("torch")
    be.set_precision("float64")
    be.grad_mode.enable()
    try:
        lens = CookeTriplet()
        glass = lens.surfaces.surfaces[1].material_post
        assert type(glass).__name__ in PK._CATALOGUE_MATERIALS
        calls = {"n": 0}
        orig = type(glass)._calculate_n

        def counted(self, *a, **k):
            calls["n"] += 1
            return orig(self, *a, **k)

        type(glass)._calculate_n = counted
        try:
            t1 = PK.pack_surface_group(lens.surfaces, np.array([0.55]))
            first = calls["n"]
            t2 = PK.pack_surface_group(lens.surfaces, np.array([0.55]))
            assert first > 0 and calls["n"] == first          # the second pack evaluated no dispersion formula
            t3 = PK.pack_surface_group(lens.surfaces, np.array([0.48, 0.55]))
            assert calls["n"] > first                          # a new wavelength is asked (once)
        finally:
            type(glass)._calculate_n = orig
        s1, p1 = t1.pack()
        s2, p2 = t2.pack()
        assert s1.tobytes() == s2.tobytes() and p1.tobytes() == p2.tobytes()
        fresh = float(Material(glass.name).n(0.55)) if hasattr(glass, "name") else None
        if fresh is not None:
            assert PK.catalogue_value(glass, "n", 0.55) == fresh == t3.surfaces[1].n2[1]
        for other in (IdealMaterial(n=1.5), AbbeMaterial(1.6, 50.0, model="polynomial")):
            PK._index_table(other, [0.55], "n")
            assert PK.catalogue_value(other, "n", 0.55) is None and "_olb_index_memo" not in other.__dict__
    finally:
        be.grad_mode.disable()
        be.set_backend("numpy")
