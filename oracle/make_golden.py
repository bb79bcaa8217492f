"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python -m oracle.make_golden

Each fixture holds, for one optical system and one seeded ray batch:
  * the packed surface table (``optiland_b200.pack.pack_surface_group`` of the LIVE
    reference objects -> ``SurfaceTable.to_arrays``),
  * the launch rays the reference generated (x,y,z,L,M,N,i,w),
  * what the reference's own ``SurfaceGroup.trace`` produced with the NumPy backend
    (fp64): the stacked per-surface records and the final ray state,
  * (polarized case) the P matrices and the final intensity.
Nothing here is a restatement: the outputs come from the reference's code path
(optiland/surfaces/surface_group.py:245-257).  The GPU box has no reference; it reads
these files.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_import import import_reference  # noqa: E402

import_reference()

import optiland.backend as be  # noqa: E402
from optiland import optic as _optic  # noqa: E402
from optiland import physical_apertures as pa  # noqa: E402
from optiland.coatings import SimpleCoating  # noqa: E402
from optiland.materials import IdealMaterial  # noqa: E402
from optiland.rays import PolarizationState, RealRays  # noqa: E402
from optiland.samples.objectives import CookeTriplet, DoubleGauss, ReverseTelephoto  # noqa: E402
from optiland.samples.simple import AsphericSinglet  # noqa: E402
from optiland.samples.telescopes import HubbleTelescope  # noqa: E402

from optiland_b200.pack import launch_scalars, pack_surface_group  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
REC = ("x", "y", "z", "L", "M", "N", "intensity", "opd")


def disk(n, seed=0, rmax=1.0):
    rng = np.random.default_rng(seed)
    r = rmax * np.sqrt(rng.random(n))
    th = 2 * np.pi * rng.random(n)
    return r * np.cos(th), r * np.sin(th)


def run_case(name, lens, rays, wavelengths, extra=None, polarized=False, expect_error=None):
    """Trace `rays` through lens.surfaces with the reference and save everything."""
    inp = {k: np.array(getattr(rays, k), dtype=np.float64) for k in "xyzLMNiw"}
    tab = pack_surface_group(lens.surfaces, wavelengths)
    out = dict(tab.to_arrays())
    for k, v in inp.items():
        out["in_" + k] = v
    err = ""
    try:
        lens.surfaces.trace(rays)
    except ValueError as e:  # Zernike range check raises in the reference
        err = str(e)
        if expect_error is None:
            raise
    out["ref_error"] = np.array(err)
    if not err:
        S = lens.surfaces
        for k in REC:
            out["rec_" + k] = np.array(getattr(S, k), dtype=np.float64)
        for k in "xyzLMN":
            out["out_" + k] = np.array(getattr(rays, k), dtype=np.float64)
        out["out_i"] = np.array(rays.i, dtype=np.float64)
        out["out_opd"] = np.array(rays.opd, dtype=np.float64)
        for k in ("L0", "M0", "N0"):
            out["out_" + k] = np.array(getattr(rays, k), dtype=np.float64)
        if polarized:
            out["out_p"] = np.array(rays.p)
    for k, v in (extra or {}).items():
        out["x_" + k] = np.asarray(v)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    nan_frac = float(np.mean(np.isnan(out["out_x"]))) if not err else -1
    print(f"{name:28s} S={tab.num_surfaces:2d} N={inp['x'].size:5d} n_wl={tab.n_wl} "
          f"nan={nan_frac:.3f} err={err[:30]!r} {os.path.getsize(path) / 1024:.0f} KiB")


def gen(lens, Hx, Hy, Px, Py, wl):
    return lens.ray_tracer.ray_generator.generate_rays(Hx, Hy, Px, Py, wl)


# ---------------------------------------------------------------------------

def case_cooke():
    """Config 1: Cooke triplet, 3 fields x hexapolar(6 rings)=127 pts, 0.55 um; also the
    reference's golden spot radii (tests/test_analysis.py:88-102)."""
    from optiland.analysis import SpotDiagram
    from optiland.distribution import create_distribution

    lens = CookeTriplet()
    d = create_distribution("hexapolar")
    d.generate_points(6)
    Px, Py = np.array(d.x), np.array(d.y)
    fields = [0.0, 0.7, 1.0]
    Hy = np.repeat(fields, Px.size)
    rays = gen(lens, np.zeros_like(Hy), Hy, np.tile(Px, 3), np.tile(Py, 3), 0.55)
    spot = SpotDiagram(CookeTriplet())
    rms = np.array(spot.rms_spot_radius())  # (fields, wavelengths)
    geo = np.array(spot.geometric_spot_radius())
    run_case("cooke_c1", lens, rays, [0.55],
             extra={"spot_rms": rms, "spot_geo": geo, "n_pupil": Px.size, "Px": Px, "Py": Py})


def case_dgauss():
    """Config 2 (small N): Double-Gauss, on-axis, 0.5876 um, uniform-in-disk pupil."""
    lens = DoubleGauss()
    Px, Py = disk(512, seed=0)
    rays = gen(lens, 0.0, 0.0, Px, Py, 0.5876)
    sc = launch_scalars(lens, 0.0, 0.0)
    run_case("dgauss_c2", lens, rays, [0.5876],
             extra={"Px": Px, "Py": Py, **{"launch_" + k: v for k, v in sc.items()}})
    # off-axis, three wavelengths interleaved per ray (trace_generic-style batches)
    lens = DoubleGauss()
    Px, Py = disk(600, seed=1)
    wl = np.tile([0.4861, 0.5876, 0.6563], 200)
    rays = gen(lens, 0.0, 0.7, Px, Py, wl)
    run_case("dgauss_multiwl", lens, rays, [0.4861, 0.5876, 0.6563])
    # NaN in band: rays far outside the pupil miss surfaces / suffer TIR
    lens = DoubleGauss()
    Px, Py = disk(300, seed=2, rmax=4.0)
    rays = gen(lens, 0.0, 1.0, Px, Py, 0.5876)
    run_case("dgauss_nan", lens, rays, [0.5876])


def reverse_telephoto_asphere(tol):
    lens = _optic.Optic()
    ref = ReverseTelephoto()
    lens.surfaces.add(index=0, radius=be.inf, thickness=be.inf)
    spec = [
        (1.69111096, 0.08259680, "N-SK10"), (0.94414496, 0.8, None), (4.32100401, 0.080256, "SK15"),
        (1.78117621, 0.5, None), (2.64050282, 0.27638160, "BASF2"), (-3.86177348, 0.1, None),
        (1.05627661, 0.2, "FK3"), (-4.06933311, 0.2001384, None), (np.inf, 0.06688, None),
        (-2.61246583, 0.064372, ("SF15", "hikari")), (0.99117409, 0.3, None),
        (9.03045960, 0.18743120, "N-LAK12"), (-1.35680743, 2.35130547, None),
    ]
    for j, (R, th, mat) in enumerate(spec, start=1):
        kw = dict(index=j, radius=R, thickness=th)
        if mat is not None:
            kw["material"] = mat
        if j == 9:
            kw["is_stop"] = True
        if j == 1:
            kw.update(surface_type="even_asphere", conic=0.0, coefficients=[1e-3, -2e-3, 5e-4], tol=tol)
        if j == 13:
            kw.update(surface_type="even_asphere", conic=-0.3, coefficients=[-2e-3, 1e-3, -4e-4], tol=tol)
        lens.surfaces.add(**kw)
    lens.surfaces.add(index=14)
    lens.set_aperture(aperture_type="EPD", value=0.3)
    lens.fields.set_type(field_type="angle")
    for f in (0, 21, 30):
        lens.fields.add(y=f)
    lens.wavelengths.add(value=0.5876, is_primary=True)
    del ref
    return lens


def case_telephoto():
    """Config 3: reverse telephoto with two even aspheres (Newton path)."""
    for tag, tol in (("tol1e-10", 1e-10), ("tol1e-6", 1e-6)):
        lens = reverse_telephoto_asphere(tol)
        Px, Py = disk(400, seed=3)
        rays = gen(lens, 0.0, 0.7, Px, Py, 0.5876)
        run_case(f"telephoto_c3_{tag}", lens, rays, [0.5876])
    lens = AsphericSinglet()
    Px, Py = disk(300, seed=4)
    rays = gen(lens, 0.0, 0.0, Px, Py, 0.587)
    run_case("aspheric_singlet", lens, rays, [0.587])


def case_hubble():
    """Config 4: Hubble (conic mirrors, radial obscuration on the primary)."""
    lens = HubbleTelescope()
    Px, Py = disk(600, seed=5)
    rays = gen(lens, 0.0, 1.0, Px, Py, 0.55)
    sc = launch_scalars(lens, 0.0, 1.0)
    run_case("hubble_c4", lens, rays, [0.55],
             extra={"Px": Px, "Py": Py, **{"launch_" + k: v for k, v in sc.items()}})


def zernike_singlet(zernike_type="fringe", coefficients=None, norm_radius=12.0, fresnel=False):
    lens = _optic.Optic()
    if coefficients is None:
        coefficients = [0.0, 2e-3, -1e-3, 4e-3, 1e-3, -2e-3, 5e-4, -7e-4, 3e-3, 2e-4, -3e-4, 1e-4,
                        6e-4, -2e-4, 1e-4, 4e-4]
    lens.surfaces.add(index=0, radius=be.inf, thickness=be.inf)
    lens.surfaces.add(index=1, radius=40.0, thickness=6.0, material="N-BK7", is_stop=True,
                      surface_type="zernike", zernike_type=zernike_type, conic=-0.5,
                      coefficients=coefficients, norm_radius=norm_radius, tol=1e-10)
    lens.surfaces.add(index=2, radius=-60.0, thickness=45.0)
    lens.surfaces.add(index=3)
    lens.set_aperture(aperture_type="EPD", value=16.0)
    lens.fields.set_type(field_type="angle")
    lens.fields.add(y=0)
    lens.fields.add(y=5)
    lens.wavelengths.add(value=0.55, is_primary=True)
    if fresnel:
        lens.surfaces.set_fresnel_coatings()
        lens.set_polarization(PolarizationState(is_polarized=False))
    return lens


def case_zernike():
    """Config 5 geometry: Zernike freeform (fringe / standard / noll), Newton path."""
    for ztype in ("fringe", "standard", "noll"):
        lens = zernike_singlet(ztype)
        Px, Py = disk(400, seed=6)
        Px[0] = 0.0
        Py[0] = 0.0  # exact axial ray: exercises the rho == 0 branch of the normal
        rays = gen(lens, 0.0, 0.0 if ztype == "fringe" else 0.7, Px, Py, 0.55)
        run_case(f"zernike_{ztype}", lens, rays, [0.55])
    # out-of-range coordinates: the reference raises ValueError (zernike.py:254-266)
    lens = zernike_singlet("fringe", norm_radius=6.0)
    Px, Py = disk(50, seed=7)
    rays = gen(lens, 0.0, 0.0, Px, Py, 0.55)
    run_case("zernike_range_error", lens, rays, [0.55], expect_error=True)


def case_polarized():
    """Config 5 polarization: Zernike surface + Fresnel coatings + unpolarized PolarizedRays,
    three wavelengths."""
    lens = zernike_singlet("fringe", fresnel=True)
    Px, Py = disk(300, seed=8)
    wl = np.tile([0.48, 0.55, 0.65], 100)
    rays = gen(lens, 0.0, 1.0, Px, Py, wl)
    assert type(rays).__name__ == "PolarizedRays"
    i0 = np.array(rays._i0)
    k0 = np.stack([np.array(rays._L0), np.array(rays._M0), np.array(rays._N0)])
    # what RealRayTracer.trace does after the loop (real_ray_tracer.py:112-113)
    import copy
    probe = copy.deepcopy(rays)
    lens2 = zernike_singlet("fringe", fresnel=True)
    lens2.surfaces.trace(probe)
    probe.update_intensity(lens2.polarization_state)
    run_case("zernike_polarized_c5", lens, rays, [0.48, 0.55, 0.65], polarized=True,
             extra={"i0": i0, "k0": k0, "final_intensity_unpolarized": np.array(probe.i)})
    # plain (uncoated) polarized trace on the Cooke triplet, polarized input state
    lens = CookeTriplet()
    lens.set_polarization(PolarizationState(is_polarized=True, Ex=1.0, Ey=0.5, phase_x=0.0, phase_y=0.3))
    Px, Py = disk(200, seed=9)
    rays = gen(lens, 0.0, 0.7, Px, Py, 0.55)
    i0 = np.array(rays._i0)
    k0 = np.stack([np.array(rays._L0), np.array(rays._M0), np.array(rays._N0)])
    probe = copy.deepcopy(rays)
    lens2 = CookeTriplet()
    lens2.set_polarization(PolarizationState(is_polarized=True, Ex=1.0, Ey=0.5, phase_x=0.0, phase_y=0.3))
    lens2.surfaces.trace(probe)
    probe.update_intensity(lens2.polarization_state)
    run_case("cooke_polarized", lens, rays, [0.55], polarized=True,
             extra={"i0": i0, "k0": k0, "final_intensity": np.array(probe.i),
                    # PolarizationState normalises (Ex, Ey) (optiland/rays/polarization_state.py:53-56)
                    "state": np.array([1.0 / np.sqrt(1.25), 0.5 / np.sqrt(1.25), 0.0, 0.3])})


def tilted_fold_lens(fresnel=False):
    lens = _optic.Optic()
    lens.surfaces.add(index=0, radius=be.inf, thickness=be.inf)
    lens.surfaces.add(index=1, radius=50.0, thickness=5.0, material="N-BK7", is_stop=True,
                      dx=0.3, dy=-0.2, rx=0.02, ry=-0.015)
    lens.surfaces.add(index=2, radius=-80.0, thickness=30.0, rz=0.4, rx=-0.01, conic=-0.8)
    lens.surfaces.add(index=3, radius=be.inf, thickness=-25.0, material="mirror", rx=np.pi / 4)
    lens.surfaces.add(index=4, radius=be.inf, thickness=0.0, rx=np.pi / 2, dy=0.5,
                      aperture=pa.RectangularAperture(-3.0, 3.5, -2.0, 4.0))
    lens.set_aperture(aperture_type="EPD", value=10.0)
    lens.fields.set_type(field_type="angle")
    lens.fields.add(y=0)
    lens.fields.add(y=3)
    lens.wavelengths.add(value=0.6, is_primary=True)
    if fresnel:
        lens.surfaces.set_fresnel_coatings()
        lens.set_polarization(PolarizationState(is_polarized=True, Ex=0.6, Ey=0.8, phase_x=0.2, phase_y=-0.4))
    return lens


def case_polarized_tilted():
    """Polarization through tilted / decentered surfaces and a fold mirror with Fresnel coatings: pins the
    reflect branch of the Jones matrices (jones.py:71-117) and the reference's local-frame P update across
    rotated poses."""
    import copy

    lens = tilted_fold_lens(fresnel=True)
    Px, Py = disk(300, seed=14)
    rays = gen(lens, 0.0, 1.0, Px, Py, 0.6)
    assert type(rays).__name__ == "PolarizedRays"
    i0 = np.array(rays._i0)
    k0 = np.stack([np.array(rays._L0), np.array(rays._M0), np.array(rays._N0)])
    probe = copy.deepcopy(rays)
    lens2 = tilted_fold_lens(fresnel=True)
    lens2.surfaces.trace(probe)
    probe.update_intensity(lens2.polarization_state)
    n = np.sqrt(0.6**2 + 0.8**2)
    run_case("tilted_fold_polarized", lens, rays, [0.6], polarized=True,
             extra={"i0": i0, "k0": k0, "final_intensity": np.array(probe.i),
                    "state": np.array([0.6 / n, 0.8 / n, 0.2, -0.4])})


def case_tilted():
    """Decentered / tilted surfaces incl. a fold mirror (rotated poses, reflect)."""
    lens = _optic.Optic()
    lens.surfaces.add(index=0, radius=be.inf, thickness=be.inf)
    lens.surfaces.add(index=1, radius=50.0, thickness=5.0, material="N-BK7", is_stop=True,
                      dx=0.3, dy=-0.2, rx=0.02, ry=-0.015)
    lens.surfaces.add(index=2, radius=-80.0, thickness=30.0, rz=0.4, rx=-0.01, conic=-0.8)
    lens.surfaces.add(index=3, radius=be.inf, thickness=-25.0, material="mirror", rx=np.pi / 4)
    lens.surfaces.add(index=4, radius=be.inf, thickness=0.0, rx=np.pi / 2, dy=0.5,
                      aperture=pa.RectangularAperture(-3.0, 3.5, -2.0, 4.0))
    lens.set_aperture(aperture_type="EPD", value=10.0)
    lens.fields.set_type(field_type="angle")
    lens.fields.add(y=0)
    lens.fields.add(y=3)
    lens.wavelengths.add(value=0.6, is_primary=True)
    Px, Py = disk(400, seed=10)
    rays = gen(lens, 0.0, 1.0, Px, Py, 0.6)
    run_case("tilted_fold", lens, rays, [0.6])


def case_misc():
    """Odd asphere, polynomial, aperture trees, simple coating, absorbing medium."""
    lens = _optic.Optic()
    glass = IdealMaterial(n=1.6, k=2e-6)  # absorbing: exercises Beer-Lambert attenuation
    lens.surfaces.add(index=0, radius=be.inf, thickness=be.inf)
    lens.surfaces.add(index=1, radius=45.0, thickness=6.0, material=glass, is_stop=True,
                      surface_type="odd_asphere", conic=0.1, coefficients=[1e-3, -2e-4, 3e-5, 1e-6], tol=1e-10,
                      aperture=pa.UnionAperture(pa.RadialAperture(r_max=5.0, r_min=1.0),
                                                pa.OffsetRadialAperture(r_max=3.0, r_min=0.0,
                                                                        offset_x=5.0, offset_y=1.0)))
    lens.surfaces.add(index=2, radius=-70.0, thickness=4.0, coating=SimpleCoating(0.9, 0.05),
                      surface_type="polynomial", conic=0.0, tol=1e-10,
                      coefficients=[[0.0, 1e-3, 2e-4, -1e-5], [-2e-3, 5e-4, 1e-5, 0.0], [3e-4, -1e-5, 0.0, 2e-6]])
    lens.surfaces.add(index=3, radius=30.0, thickness=5.0, material="N-SF11",
                      aperture=pa.DifferenceAperture(pa.EllipticalAperture(6.0, 4.5, 0.2, -0.1),
                                                     pa.RectangularAperture(-1.0, 1.0, -0.5, 0.7)))
    lens.surfaces.add(index=4, radius=be.inf, thickness=20.0,
                      aperture=pa.IntersectionAperture(pa.RadialAperture(r_max=6.0),
                                                       pa.RectangularAperture(-5.0, 5.0, -4.0, 6.0)))
    lens.surfaces.add(index=5)
    lens.set_aperture(aperture_type="EPD", value=14.0)
    lens.fields.set_type(field_type="angle")
    lens.fields.add(y=0)
    lens.fields.add(y=4)
    lens.wavelengths.add(value=0.55, is_primary=True)
    Px, Py = disk(500, seed=11)
    # (no exact chief ray here: the r^1 odd term makes the vertex a cone tip, where the slope
    #  is discontinuous and any rounding difference flips it)
    rays = gen(lens, 0.0, 0.5, Px, Py, 0.55)
    run_case("misc_apertures_coatings", lens, rays, [0.55])


def case_more_geometries():
    """Chebyshev, biconic and toroidal surfaces (the remaining Newton-family geometries)."""
    lens = _optic.Optic()
    lens.surfaces.add(index=0, radius=be.inf, thickness=be.inf)
    lens.surfaces.add(index=1, radius=60.0, thickness=5.0, material="N-BK7", is_stop=True, conic=-0.3,
                      surface_type="chebyshev", tol=1e-10, norm_x=9.0, norm_y=11.0,
                      coefficients=[[0.0, 2e-3, -1e-3, 4e-4], [1e-3, -5e-4, 2e-4, 0.0], [-3e-3, 1e-4, 0.0, 6e-5],
                                    [2e-4, 0.0, -8e-5, 0.0]])
    lens.surfaces.add(index=2, thickness=4.0, material="N-SF11", surface_type="biconic", radius_x=-45.0,
                      radius_y=-70.0, conic_x=0.4, conic_y=-1.2, tol=1e-10)
    lens.surfaces.add(index=3, thickness=30.0, surface_type="toroidal", radius_x=-120.0, radius_y=-55.0,
                      conic=-0.6, toroidal_coeffs_poly_y=[1e-5, -2e-7], tol=1e-10)
    lens.surfaces.add(index=4)
    lens.set_aperture(aperture_type="EPD", value=12.0)
    lens.fields.set_type(field_type="angle")
    lens.fields.add(y=0)
    lens.fields.add(y=3)
    lens.wavelengths.add(value=0.55, is_primary=True)
    Px, Py = disk(400, seed=12)
    rays = gen(lens, 0.0, 0.6, Px, Py, 0.55)
    run_case("cheb_biconic_toroidal", lens, rays, [0.55])
    # Chebyshev range error: the reference raises ValueError (chebyshev.py:230-244)
    lens = _optic.Optic()
    lens.surfaces.add(index=0, radius=be.inf, thickness=be.inf)
    lens.surfaces.add(index=1, radius=60.0, thickness=5.0, material="N-BK7", is_stop=True, surface_type="chebyshev",
                      norm_x=4.0, norm_y=4.0, coefficients=[[0.0, 1e-3], [1e-3, 0.0]])
    lens.surfaces.add(index=2, thickness=30.0)
    lens.surfaces.add(index=3)
    lens.set_aperture(aperture_type="EPD", value=12.0)
    lens.fields.set_type(field_type="angle")
    lens.fields.add(y=0)
    lens.wavelengths.add(value=0.55, is_primary=True)
    Px, Py = disk(50, seed=13)
    rays = gen(lens, 0.0, 0.0, Px, Py, 0.55)
    run_case("chebyshev_range_error", lens, rays, [0.55], expect_error=True)


def case_huygens():
    """f-3: the reference's own Numba Huygens-Fresnel summation on a seeded pupil / image grid, and the
    HuygensPSF of the Cooke triplet (pupil data + image grid + PSF) end to end."""
    from optiland.psf import HuygensPSF
    from optiland.psf.huygens_fresnel_strategies import NumbaSummation

    rng = np.random.default_rng(21)
    npup = 700
    r = 9.0 * np.sqrt(rng.random(npup))
    th = 2 * np.pi * rng.random(npup)
    Rp = -80.0  # sign: n_Q = Q / Rp must point from the pupil towards the image (obliquity ~ 1)
    pu, pv = r * np.cos(th), r * np.sin(th)
    pw = -np.sqrt(Rp**2 - pu**2 - pv**2)  # exit-pupil reference sphere centred on the image point
    amp = 0.5 + rng.random(npup)
    opd = 2e-4 * rng.normal(size=npup)  # mm
    gx, gy = np.meshgrid(np.linspace(-0.02, 0.02, 24), np.linspace(-0.02, 0.02, 24))
    gz = np.zeros_like(gx)
    psf = NumbaSummation().compute(gx, gy, gz, pu, pv, pw, amp, opd, 0.55e-3, Rp)
    out = dict(image_x=gx, image_y=gy, image_z=gz, pupil_x=pu, pupil_y=pv, pupil_z=pw, pupil_amp=amp, pupil_opd=opd,
               wavelength=0.55e-3, Rp=Rp, psf=np.array(psf))
    # system level: Cooke triplet, field (0, 0.7), 32 x 32 pupil, 32 x 32 image
    lens = CookeTriplet()
    h = HuygensPSF(lens, field=(0, 0.7), wavelength=0.55, num_rays=32, image_size=32)
    data = h.get_data((0, 0.7), 0.55)
    ix, iy, iz = h._get_image_coordinates()
    out.update(sys_image_x=np.array(ix), sys_image_y=np.array(iy), sys_image_z=np.array(iz),
               sys_pupil_x=np.array(data.pupil_x), sys_pupil_y=np.array(data.pupil_y), sys_pupil_z=np.array(data.pupil_z),
               sys_pupil_amp=np.sqrt(np.array(data.intensity)), sys_pupil_opd=np.array(data.opd) * 0.55e-3,
               sys_Rp=float(data.radius), sys_norm=float(h.normalization), sys_psf=np.array(h.psf))
    path = os.path.join(OUT, "huygens_psf_ref.npz")
    np.savez_compressed(path, **out)
    print("huygens_psf_ref", psf.shape, float(psf.max()), "system strehl-scaled peak", float(np.max(h.psf)),
          f"{os.path.getsize(path) / 1024:.0f} KiB")


def case_generic_fields():
    """trace_generic-shaped batches (raytrace/real_ray_tracer.py:120-154): per-ray field points (and wavelengths):
    infinite-object angle fields (config 5's shape: several fields x several wavelengths in one call), an
    object-height field on a finite object, and the telecentric lithography lens."""
    from optiland.samples.lithography import UVProjectionLens

    rng = np.random.default_rng(77)
    n = 450
    for name, lens, wls in (("generic_dgauss", DoubleGauss(), [0.4861, 0.5876, 0.6563]),
                            ("generic_finite_height", finite_relay("object_height"), [0.5876]),
                            ("generic_finite_angle", finite_relay("angle"), [0.5876]),
                            ("generic_litho", UVProjectionLens(), [0.248])):
        Px, Py = disk(n, seed=12)
        Hx = np.round(rng.uniform(-0.6, 0.6, n), 3)
        Hy = np.round(rng.uniform(-1.0, 1.0, n), 3)
        keep = Hx**2 + Hy**2 <= 1.0
        Hx, Hy = np.where(keep, Hx, 0.0), np.where(keep, Hy, 0.5)
        wl = np.asarray(wls)[rng.integers(0, len(wls), n)]
        rays = gen(lens, Hx, Hy, Px, Py, wl)
        sc = launch_scalars(lens, 0.0, 0.0)
        run_case(name, lens, rays, wls, extra={"Px": Px, "Py": Py, "Hx": Hx, "Hy": Hy,
                                                **{"launch_" + k: v for k, v in sc.items()}})


def case_forbes():
    """Forbes Q (slope-orthogonal) radial aspheres (geometries/forbes/geometry.py:187-366): a singlet with two
    forbes_qbfs surfaces (one with a conic base), rays reaching beyond the normalisation radius on the second
    (departure switched off there) and a chief ray through the vertex."""
    lens = _optic.Optic()
    lens.surfaces.add(index=0, radius=be.inf, thickness=be.inf)
    lens.surfaces.add(index=1, radius=22.0, thickness=6.0, material="N-BK7", is_stop=True, conic=-0.4,
                      radial_terms={0: 0.12, 1: -0.041, 2: 0.013, 3: -0.006, 5: 0.002}, norm_radius=9.0,
                      surface_type="forbes_qbfs", tol=1e-10)
    lens.surfaces.add(index=2, radius=-31.0, thickness=28.0, conic=0.0,
                      radial_terms={0: -0.27, 1: 0.087, 2: -0.048, 3: 0.026, 4: -0.012}, norm_radius=6.5,
                      surface_type="forbes_qbfs", tol=1e-10)
    lens.surfaces.add(index=3)
    lens.set_aperture(aperture_type="EPD", value=15.0)
    lens.fields.set_type(field_type="angle")
    lens.fields.add(y=0)
    lens.fields.add(y=4)
    lens.wavelengths.add(value=0.55, is_primary=True)
    Px, Py = disk(500, seed=31)
    Px[0] = Py[0] = 0.0
    rays = gen(lens, np.zeros(Px.size), np.concatenate([[0.0], np.full(Px.size - 1, 0.6)]), Px, Py, 0.55)
    run_case("forbes_qbfs", lens, rays, [0.55])


def wavefront_ref_scalars(lens, strategy, field, wl):
    """Reference-sphere scalars of ChiefRayStrategy.compute_wavefront_data (steps 1-2,
    optiland/wavefront/strategy.py:160-170), computed with the reference's own methods."""
    chief = lens.trace_generic(*field, Px=0.0, Py=0.0, wavelength=wl)
    geometry = strategy._create_reference_geometry(chief)
    opd_img_ref = geometry.path_length(chief, strategy.n_image)
    opd_ref = strategy._correct_tilt(field, chief.opd - opd_img_ref, x=0, y=0)
    tilt = (0.0, 0.0)
    fd = lens.fields.field_definition
    if type(fd).__name__ == "AngleField" and lens.object_surface.is_infinite:   # strategy.py:112-138
        tx = np.tan(np.deg2rad(field[0] * float(lens.fields.max_field)))
        ty = np.tan(np.deg2rad(field[1] * float(lens.fields.max_field)))
        uz = 1.0 / np.sqrt(1.0 + tx**2 + ty**2)
        epd = float(lens.paraxial.EPD())
        tilt = (tx * uz * epd / 2, ty * uz * epd / 2)
    return {"center": np.array([float(v) for v in geometry.center]), "radius": float(geometry.radius),
            "n_image": float(np.ravel(strategy.n_image)[0]), "tilt": np.array(tilt),
            "opd_ref": float(np.ravel(opd_ref)[0]), "wavelength_um": float(wl)}


def case_wavefront():
    """f-2 (second half): the reference's Wavefront analysis, chief-ray strategy, spherical reference."""
    from optiland.wavefront import Wavefront

    out = {}
    for tag, lens, field, wl in (("dgauss", DoubleGauss(), (0.0, 0.7), 0.5876),
                                 ("cooke", CookeTriplet(), (0.0, 1.0), 0.55),
                                 ("finite", finite_relay("object_height"), (2.0 / 9.0, 1.0), 0.5876),
                                 ("hubble", HubbleTelescope(), (0.0, 1.0), 0.55)):
        w = Wavefront(lens, fields=[field], wavelengths=[wl], num_rays=12, distribution="hexapolar", strategy="chief_ray")
        data = w.get_data(field, wl)
        ref = wavefront_ref_scalars(lens, w.strategy, field, wl)
        tab = pack_surface_group(lens.surfaces, [wl])
        sc = launch_scalars(lens, field[0], field[1])
        out.update({f"{tag}_{k}": v for k, v in tab.to_arrays().items()})
        out.update({f"{tag}_launch_{k}": v for k, v in sc.items()})
        out.update({f"{tag}_ref_{k}": v for k, v in ref.items()})
        out.update({f"{tag}_Px": np.array(w.distribution.x), f"{tag}_Py": np.array(w.distribution.y),
                    f"{tag}_opd": np.array(data.opd), f"{tag}_pupil_x": np.array(data.pupil_x),
                    f"{tag}_pupil_y": np.array(data.pupil_y), f"{tag}_pupil_z": np.array(data.pupil_z),
                    f"{tag}_intensity": np.array(data.intensity), f"{tag}_radius": float(data.radius)})
        print(f"wavefront {tag:8s} N={np.size(data.opd)} opd rms {float(np.std(np.array(data.opd))):.4f} waves "
              f"R={float(data.radius):.3f} tilt={ref['tilt']}")
    path = os.path.join(OUT, "wavefront_chief_ray_ref.npz")
    np.savez_compressed(path, **out)
    print("wavefront_chief_ray_ref", f"{os.path.getsize(path) / 1024:.0f} KiB")


def finite_relay(field_type):
    """Finite-conjugate 1:1-ish relay doublet: object 120 mm in front, `field_type` fields."""
    lens = _optic.Optic()
    lens.surfaces.add(index=0, radius=be.inf, thickness=120.0)
    lens.surfaces.add(index=1, radius=61.0, thickness=7.0, material="N-BK7")
    lens.surfaces.add(index=2, radius=-44.0, thickness=2.5, material="SF2", is_stop=True)
    lens.surfaces.add(index=3, radius=-129.0, thickness=110.0)
    lens.surfaces.add(index=4)
    lens.set_aperture(aperture_type="EPD", value=18.0)
    lens.fields.set_type(field_type=field_type)
    lens.fields.add(y=0)
    lens.fields.add(y=6.0)
    lens.fields.add(y=9.0, x=2.0) if field_type == "object_height" else lens.fields.add(y=4.0, x=1.0)
    lens.wavelengths.add(value=0.5876, is_primary=True)
    return lens


def case_finite_objects():
    """f-1 beyond infinite-object angle fields: finite object with object-height and angle fields
    (object_height.py:17-46, angle.py:49-58) and the object-space telecentric lithography sample
    (paraxial.py:82-88; 44 surfaces)."""
    from optiland.samples.lithography import UVProjectionLens

    for name, lens, H in (("finite_object_height", finite_relay("object_height"), (2.0 / 9.0, 1.0)),
                          ("finite_object_angle", finite_relay("angle"), (0.25, 1.0)),
                          ("litho_telecentric", UVProjectionLens(), (0.0, 1.0))):
        Px, Py = disk(400, seed=11)
        wl = float(lens.primary_wavelength)
        # per-ray field arrays, as RealRayTracer.trace passes them (real_ray_tracer.py:90-103): with scalar H
        # the object-point origin would come back with size 1 and the object-surface record would be ragged
        rays = gen(lens, np.full(Px.size, H[0]), np.full(Px.size, H[1]), Px, Py, wl)
        sc = launch_scalars(lens, H[0], H[1])
        run_case(name, lens, rays, [wl], extra={"Px": Px, "Py": Py, **{"launch_" + k: v for k, v in sc.items()}})


def main():
    be.set_backend("numpy")
    case_cooke()
    case_dgauss()
    case_telephoto()
    case_hubble()
    case_zernike()
    case_polarized()
    case_polarized_tilted()
    case_tilted()
    case_misc()
    case_more_geometries()
    case_huygens()
    case_finite_objects()
    case_generic_fields()
    case_forbes()
    case_wavefront()
    case_autograd()
    case_config_shapes()
    case_c3_grad_full_size()


def case_autograd():
    """Config 3 gradient golden: d(RMS spot about the centroid)/d(radius, conic, coefficients, cs.z)
    from the REFERENCE's own torch-CPU fp64 autograd (loss.backward() through its eager graph),
    on the reverse telephoto with two even aspheres; same rays as telephoto_c3_tol1e-10."""
    import torch

    be.set_backend("torch")
    be.set_precision("float64")
    be.grad_mode.enable()
    try:
        lens = reverse_telephoto_asphere(1e-10)
        Px, Py = disk(400, seed=3)
        rays = gen(lens, 0.0, 0.7, be.array(Px), be.array(Py), 0.5876)
        # isolate the hot path: the launch state also depends on the radii through paraxial ray
        # aiming (EPL/EPD); detach it so the golden is d(trace)/d(parameter) at fixed launch rays
        rays = RealRays(*[getattr(rays, k).detach() for k in ("x", "y", "z", "L", "M", "N", "i", "w")])
        lens.surfaces.trace(rays)
        x = lens.surfaces.x[-1, :]
        y = lens.surfaces.y[-1, :]
        loss = torch.sqrt(torch.mean((x - torch.mean(x)) ** 2 + (y - torch.mean(y)) ** 2))
        loss.backward()
        out = {"loss": float(loss)}
        for s in (1, 2, 13):
            g = lens.surfaces.surfaces[s].geometry
            out[f"d_radius_{s}"] = float(g.radius.grad)
            if g.cs.z.grad is not None:
                out[f"d_z_{s}"] = float(g.cs.z.grad)
            if hasattr(g, "coefficients"):
                out[f"d_conic_{s}"] = float(g.k.grad)
        # coefficients are python floats in the reference (not leaves): finite-difference them there
    finally:
        be.grad_mode.disable()
        be.set_backend("numpy")
    path = os.path.join(OUT, "telephoto_c3_grad.npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in out.items()})
    print("telephoto_c3_grad", out)


def case_config_shapes():
    """BASELINE.json's configurations at their STATED shapes (round-2 fixtures; each holds only what a full-size
    comparison needs -- expected statistics, not the per-ray records):

    * ``c1_cooke_64rings``: config 1 as stated -- Cooke triplet, 3 fields x 1 wavelength, hexapolar pupil with 64 rings
      (12 481 points per field): the reference's SpotDiagram RMS / geometric radii and centroids;
    * ``c5_opd_maps``: config 5 as stated -- Zernike freeform + Fresnel coatings + unpolarized PolarizedRays, 5 fields x
      3 wavelengths: the reference's Wavefront (chief-ray strategy) OPD map, exit-pupil intercepts, intensity and P
      matrices for every (field, wavelength), with the reference-sphere scalars of steps 1-2;
    * ``generic_polarized_c5``: the same system in trace_generic's call shape (per-ray fields and wavelengths),
      with launch scalars -- what the 32 M-ray sharded run replays from global ray indices;
    * ``telephoto_c3_4M_grad``: config 3's gradient at its stated size -- d(RMS spot)/d(radius, conic, z) over the
      3 998 611-ray hexapolar pupil (1 154 rings) from the reference's own torch-CPU fp64 autograd, accumulated over
      chunks (the loss depends on the rays only through sum x, sum y, sum (x^2 + y^2), which are additive)."""
    from optiland.analysis import SpotDiagram
    from optiland.distribution import create_distribution
    from optiland.wavefront import Wavefront

    # ---- C1 at 64 rings ---------------------------------------------------------------------------
    lens = CookeTriplet()
    spot = SpotDiagram(lens, wavelengths=[0.55], num_rings=64, distribution="hexapolar")
    rms = np.array(spot.rms_spot_radius(), dtype=np.float64)          # (3 fields, 1 wavelength)
    geo = np.array(spot.geometric_spot_radius(), dtype=np.float64)
    cen = np.array([[np.asarray(c, dtype=np.float64) for c in spot.centroid()]])
    chief = np.array([[float(np.ravel(v)[0]) for v in c] for c in spot._get_reference_centers(spot.data)])
    spot_c = SpotDiagram(lens, wavelengths=[0.55], num_rings=64, distribution="hexapolar", reference="centroid")
    rms_c = np.array(spot_c.rms_spot_radius(), dtype=np.float64)
    d = create_distribution("hexapolar")
    d.generate_points(64)
    out = dict(pack_surface_group(lens.surfaces, [0.55]).to_arrays())
    fields = lens.fields.get_field_coords()
    for j, (hx, hy) in enumerate(fields):
        out.update({f"f{j}_launch_{k}": v for k, v in launch_scalars(lens, float(hx), float(hy)).items()})
    out.update(spot_rms=rms, spot_geo=geo, centroid=np.squeeze(cen), chief_center=chief, spot_rms_centroid=rms_c,
               n_rings=64, n_pupil=np.size(d.x),
               fields=np.array(fields, dtype=np.float64), wavelength=0.55)
    path = os.path.join(OUT, "c1_cooke_64rings_ref.npz")
    np.savez_compressed(path, **out)
    print("c1_cooke_64rings", rms.ravel(), f"{np.size(d.x)} pupil points per field, {os.path.getsize(path) / 1024:.0f} KiB")

    # ---- C5: OPD maps, 5 fields x 3 wavelengths -------------------------------------------------
    fields5 = [(0.0, 0.0), (0.0, 0.5), (0.0, 1.0), (0.5, 0.5), (-0.7, 0.3)]
    wls = [0.48, 0.55, 0.65]
    lens = zernike_singlet("fringe", fresnel=True)
    w = Wavefront(lens, fields=fields5, wavelengths=wls, num_rays=8, distribution="hexapolar", strategy="chief_ray")
    out = {"fields": np.array(fields5), "wavelengths": np.array(wls), "Px": np.array(w.distribution.x),
           "Py": np.array(w.distribution.y)}
    worst_rms = 0.0
    for fi, f in enumerate(fields5):
        for wi, wl in enumerate(wls):
            tag = f"f{fi}w{wi}"
            data = w.get_data(f, wl)
            ref = wavefront_ref_scalars(lens, w.strategy, f, wl)
            if fi == 0:
                out.update({f"w{wi}_{k}": v for k, v in pack_surface_group(lens.surfaces, [wl]).to_arrays().items()})
            out.update({f"{tag}_launch_{k}": v for k, v in launch_scalars(lens, f[0], f[1]).items()})
            out.update({f"{tag}_ref_{k}": v for k, v in ref.items()})
            out.update({f"{tag}_opd": np.array(data.opd), f"{tag}_pupil_x": np.array(data.pupil_x),
                        f"{tag}_pupil_y": np.array(data.pupil_y), f"{tag}_pupil_z": np.array(data.pupil_z),
                        f"{tag}_intensity": np.array(data.intensity), f"{tag}_p": np.array(data.prt_matrix)})
            worst_rms = max(worst_rms, float(np.std(np.array(data.opd))))
    path = os.path.join(OUT, "c5_opd_maps_ref.npz")
    np.savez_compressed(path, **out)
    print(f"c5_opd_maps: 15 maps x {np.size(w.distribution.x)} points, largest OPD rms {worst_rms:.3f} waves, "
          f"{os.path.getsize(path) / 1024:.0f} KiB")

    # ---- C5 in trace_generic's call shape ----------------------------------------------------------
    lens = zernike_singlet("fringe", fresnel=True)
    Pxg, Pyg = disk(40, seed=21, rmax=0.95)
    Hx = np.repeat([f[0] for f in fields5], 3 * Pxg.size)
    Hy = np.repeat([f[1] for f in fields5], 3 * Pxg.size)
    wl = np.tile(np.repeat(wls, Pxg.size), 5)
    Px, Py = np.tile(Pxg, 15), np.tile(Pyg, 15)
    rays = gen(lens, Hx, Hy, Px, Py, wl)
    sc = launch_scalars(lens, 0.0, 0.0)
    k0 = np.stack([np.array(rays._L0), np.array(rays._M0), np.array(rays._N0)])
    run_case("generic_polarized_c5", lens, rays, wls, polarized=True,
             extra={"Px": Px, "Py": Py, "Hx": Hx, "Hy": Hy, "i0": np.array(rays._i0), "k0": k0,
                    **{"launch_" + k: v for k, v in sc.items()}})


def case_c3_grad_full_size():
    """See case_config_shapes: the config-3 gradient over the full 1 154-ring hexapolar pupil, chunked."""
    import torch
    from optiland.distribution import create_distribution

    be.set_backend("numpy")
    d = create_distribution("hexapolar")
    d.generate_points(1154)
    Px_all, Py_all = np.array(d.x), np.array(d.y)
    n = Px_all.size
    be.set_backend("torch")
    be.set_precision("float64")
    be.grad_mode.enable()
    try:
        lens = reverse_telephoto_asphere(1e-10)
        # launch rays at fixed values (the hot path's gradient, as in case_autograd)
        names = [("radius", 1), ("radius", 2), ("radius", 13), ("k", 1), ("k", 13), ("z", 1), ("z", 13)]

        def leaf(kind, s):
            g = lens.surfaces.surfaces[s].geometry
            return g.cs.z if kind == "z" else getattr(g, kind)

        leaves = [leaf(k, s) for k, s in names]
        sums = np.zeros(3)
        gsum = np.zeros((3, len(leaves)))
        chunk = 250_000
        for lo in range(0, n, chunk):
            rays = gen(lens, 0.0, 0.7, be.array(Px_all[lo:lo + chunk]), be.array(Py_all[lo:lo + chunk]), 0.5876)
            rays = RealRays(*[getattr(rays, k).detach() for k in ("x", "y", "z", "L", "M", "N", "i", "w")])
            lens.surfaces.trace(rays)
            x = lens.surfaces.x[-1, :]
            y = lens.surfaces.y[-1, :]
            parts = [x.sum(), y.sum(), (x * x + y * y).sum()]
            for q, v in enumerate(parts):
                g = torch.autograd.grad(v, leaves, retain_graph=q < 2, allow_unused=True)
                gsum[q] += [0.0 if t is None else float(t) for t in g]
                sums[q] += float(v.detach())
            print(f"  c3 grad chunk {lo // chunk + 1}/{(n + chunk - 1) // chunk}", flush=True)
        cx, cy = sums[0] / n, sums[1] / n
        var = sums[2] / n - cx * cx - cy * cy
        loss = np.sqrt(var)
        # d var = d(S2)/n - 2 cx d(Sx)/n - 2 cy d(Sy)/n ;  d loss = d var / (2 loss)
        dvar = gsum[2] / n - 2 * cx * gsum[0] / n - 2 * cy * gsum[1] / n
        dloss = dvar / (2 * loss)
        out = {"loss": loss, "n_rays": n, "n_rings": 1154, "centroid": np.array([cx, cy])}
        for (kind, s), v in zip(names, dloss):
            out[f"d_{'conic' if kind == 'k' else kind}_{s}"] = v
    finally:
        be.grad_mode.disable()
        be.set_backend("numpy")
    out.update({"launch_" + k: v for k, v in launch_scalars(reverse_telephoto_asphere(1e-10), 0.0, 0.7).items()})
    path = os.path.join(OUT, "telephoto_c3_4M_grad.npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in out.items()})
    print("telephoto_c3_4M_grad", {k: (float(v) if np.ndim(v) == 0 else v) for k, v in out.items()})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "shapes":
        be.set_backend("numpy")
        case_config_shapes()
    elif len(sys.argv) > 1 and sys.argv[1] == "c3full":
        case_c3_grad_full_size()
    elif len(sys.argv) > 1 and sys.argv[1] == "huygens":
        be.set_backend("numpy")
        case_huygens()
    elif len(sys.argv) > 1 and sys.argv[1] == "more":
        be.set_backend("numpy")
        case_more_geometries()
    elif len(sys.argv) > 1 and sys.argv[1] == "autograd":
        case_autograd()
    elif len(sys.argv) > 1 and sys.argv[1] == "poltilt":
        be.set_backend("numpy")
        case_polarized_tilted()
    elif len(sys.argv) > 1 and sys.argv[1] == "generic":
        be.set_backend("numpy")
        case_generic_fields()
    elif len(sys.argv) > 1 and sys.argv[1] == "forbes":
        be.set_backend("numpy")
        case_forbes()
    elif len(sys.argv) > 1 and sys.argv[1] == "wavefront":
        be.set_backend("numpy")
        case_wavefront()
    elif len(sys.argv) > 1 and sys.argv[1] == "finite":
        be.set_backend("numpy")
        case_finite_objects()
    else:
        main()
