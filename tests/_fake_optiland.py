"""Duck-typed stand-ins for Optiland's live objects (Surface, geometry, coordinate system, materials,
apertures) whose scalar attributes are torch tensors on a chosen device -- what ``optiland_b200.pack`` sees
under ``be.set_backend("torch"); be.set_device("cuda")``.  Built FROM a SurfaceTable so that packing them
must reproduce that table.  Only the class NAMES matter to pack.py (it dispatches on them)."""
import numpy as np
import torch

from optiland_b200 import table as T


def _obj(name, **attrs):
    o = type(name, (), {})()
    for k, v in attrs.items():
        setattr(o, k, v)
    return o


class _Material:
    """n(wl) / k(wl) with a per-wavelength cache, like optiland/materials/base.py:98-149."""

    def __init__(self, wavelengths, n, k, dev):
        self._n = {float(w): torch.tensor(float(v), dtype=torch.float64, device=dev) for w, v in zip(wavelengths, n)}
        self._k = {float(w): torch.tensor(float(v), dtype=torch.float64, device=dev) for w, v in zip(wavelengths, k)}

    def n(self, wl):
        return self._n[float(wl)]

    def k(self, wl):
        return self._k[float(wl)]


def _aperture(prog, sc):
    op = int(prog[0])
    if op == T.AP_RADIAL:
        return _obj("RadialAperture", r_max=sc(prog[1]), r_min=sc(prog[2]))
    raise NotImplementedError(op)


def fake_surfaces(table: T.SurfaceTable, device, angles=None):
    """List of fake Surface objects for an UNROTATED plane / standard / even-asphere table.  ``angles``:
    optional {surface index: (rx, ry, rz)} to tilt surfaces (then the packed R must equal
    table.rotation_matrix of those angles)."""
    dev = torch.device(device)
    sc = lambda v: torch.tensor(float(v), dtype=torch.float64, device=dev)  # noqa: E731
    wl = table.wavelengths
    out = []
    for s, spec in enumerate(table.surfaces):
        if spec.kind == T.GEOM_NOOP:
            out.append(_obj("ObjectSurface"))
            continue
        assert not spec.rotated and spec.kind in (T.GEOM_PLANE, T.GEOM_STANDARD, T.GEOM_EVEN_ASPHERE)
        rx, ry, rz = (angles or {}).get(s, (0.0, 0.0, 0.0))
        cs = _obj("CoordinateSystem", x=sc(spec.t[0]), y=sc(spec.t[1]), z=sc(spec.t[2]), rx=sc(rx), ry=sc(ry), rz=sc(rz),
                  reference_cs=None)
        if spec.kind == T.GEOM_PLANE:
            g = _obj("Plane", cs=cs)
        elif spec.kind == T.GEOM_STANDARD:
            g = _obj("StandardGeometry", cs=cs, radius=sc(spec.radius), k=sc(spec.conic))
        else:
            g = _obj("EvenAsphere", cs=cs, radius=sc(spec.radius), k=sc(spec.conic), tol=spec.tol, max_iter=spec.max_iter,
                     coefficients=[sc(c) for c in spec.coefficients])
        im = _obj("RefractiveReflectiveModel", is_reflective=spec.reflective, coating=None, bsdf=None)
        if spec.coating == T.COAT_SIMPLE:
            im.coating = _obj("SimpleCoating", transmittance=sc(spec.coat_t), reflectance=sc(spec.coat_r))
        ap = None
        if spec.aperture is not None:
            ap = _aperture(spec.aperture, sc)
        name = "ImageSurface" if s == table.num_surfaces - 1 else "Surface"
        out.append(_obj(name, geometry=g, interaction_model=im, aperture=ap,
                        material_pre=_Material(wl, spec.n1, spec.k1, dev),
                        material_post=_Material(wl, spec.n2, np.zeros_like(spec.n2), dev)))
    return out
