"""Launch state of one field point as a closed form of (Px, Py): infinite-object angle fields, finite
objects (angle / object-height fields) and object-space telecentric systems.

Host-side mirror of the step immediately before the hot path, for boxes without
the reference (bench, GPU tests): ``AngleField.get_ray_origins``
(/root/reference/optiland/fields/field_types/angle.py:17-58) followed by
``ParaxialRayAimer.aim_rays`` (optiland/rays/ray_aiming/paraxial.py:33-106).  The
scalars come from ``optiland_b200.pack.launch_scalars`` (stored in the golden fixtures).
Works on numpy arrays and torch tensors alike (only arithmetic operators + ``sqrt``).
"""
from __future__ import annotations

import math


def _sqrt(v):
    return v.sqrt() if hasattr(v, "sqrt") else v ** 0.5


def _tan(v):
    if hasattr(v, "tan"):
        return v.tan()
    import numpy as np

    return np.tan(v)


def launch_infinite_angle(Px, Py, sc: dict):
    """Return x0, y0, z0, L, M, N for pupil coordinates (Px, Py) and launch scalars ``sc``."""
    EPL, EPD, offset = sc["EPL"], sc["EPD"], sc["offset"]
    field_x = sc["max_field"] * sc["Hx"]
    field_y = sc["max_field"] * sc["Hy"]
    xo = -math.tan(math.radians(field_x)) * (offset + EPL)
    yo = -math.tan(math.radians(field_y)) * (offset + EPL)
    zo = sc["z1"] - offset
    x0 = Px * EPD / 2 * sc["vx"] + xo
    y0 = Py * EPD / 2 * sc["vy"] + yo
    z0 = Px * 0 + zo
    x1 = Px * EPD * sc["vx"] / 2
    y1 = Py * EPD * sc["vy"] / 2
    z1 = Px * 0 + EPL
    mag = _sqrt((x1 - x0) ** 2 + (y1 - y0) ** 2 + (z1 - z0) ** 2)
    L = (x1 - x0) / mag
    M = (y1 - y0) / mag
    N = (z1 - z0) / mag
    return x0, y0, z0, L, M, N


def pupil_affine_infinite_angle(sc: dict) -> dict:
    """The same launch state as ``launch_infinite_angle`` in the affine form the kernel evaluates
    (include/olb.h ``OlbPupilLaunch``): origin = origin0 + origin_scale * (Px, Py), target likewise,
    direction = normalised (target - origin)."""
    EPL, EPD, offset = sc["EPL"], sc["EPD"], sc["offset"]
    xo = -math.tan(math.radians(sc["max_field"] * sc["Hx"])) * (offset + EPL)
    yo = -math.tan(math.radians(sc["max_field"] * sc["Hy"])) * (offset + EPL)
    zo = sc["z1"] - offset
    sx, sy = EPD / 2 * sc["vx"], EPD / 2 * sc["vy"]
    return {"origin0": (xo, yo, zo), "origin_scale": (sx, sy), "target0": (0.0, 0.0, EPL),
            "target_scale": (EPD * sc["vx"] / 2, EPD * sc["vy"] / 2), "intensity": 1.0}


def pupil_affine(sc: dict) -> dict:
    """Affine launch form for any mode of ``pack.launch_scalars`` (see there).  All three are
    origin = origin0 + origin_scale * (Px, Py), target = target0 + target_scale * (Px, Py)."""
    mode = int(sc.get("mode", 0))
    if mode == 0:
        return pupil_affine_infinite_angle(sc)
    o0 = (sc["x0"], sc["y0"], sc["z0"])
    if mode == 1:        # paraxial.py:90-96: aim at the paraxial entrance pupil
        return {"origin0": o0, "origin_scale": (0.0, 0.0), "target0": (0.0, 0.0, sc["EPL"]),
                "target_scale": (sc["EPD"] * sc["vx"] / 2, sc["EPD"] * sc["vy"] / 2), "intensity": 1.0}
    if mode == 3:        # infinite object, image-height field types: the origin slides with the pupil point (mode 0's form)
        return {"origin0": o0, "origin_scale": (sc["EPD"] / 2 * sc["vx"], sc["EPD"] / 2 * sc["vy"]), "target0": (0.0, 0.0, sc["EPL"]),
                "target_scale": (sc["EPD"] * sc["vx"] / 2, sc["EPD"] * sc["vy"] / 2), "intensity": 1.0}
    if mode == 2:        # paraxial.py:82-88: telecentric object space
        sin = sc["sin"]
        z1 = math.sqrt(1 - sin ** 2) / sin + sc["z0"]
        return {"origin0": o0, "origin_scale": (0.0, 0.0), "target0": (sc["x0"], sc["y0"], z1),
                "target_scale": (sc["vx"], sc["vy"]), "intensity": 1.0}
    raise ValueError(f"unknown launch mode {mode}")


def launch_from_affine(Px, Py, aff: dict):
    """x0, y0, z0, L, M, N from the affine form -- the arithmetic of the kernel's ``pupil_launch``
    (olb_math.cuh), on numpy arrays or torch tensors (incl. the per-ray field offsets of pupil_affine_fields)."""
    ofx = ofy = tfx = tfy = 0.0
    if aff.get("fields") is not None:
        Hx, Hy = aff["fields"]
        if aff.get("vig") is not None:       # nearest defined field's vignetting factors (numpy arrays only: test mirror)
            import numpy as np

            tab, power = aff["vig"]
            tab = np.asarray(tab, dtype=np.float64).reshape(-1, 4)
            hx, hy = np.asarray(Hx, dtype=np.float64), np.asarray(Hy, dtype=np.float64)
            d2 = (hx[:, None] - tab[None, :, 0]) ** 2 + (hy[:, None] - tab[None, :, 1]) ** 2
            best = np.argmin(d2, axis=1)
            for _ in range(int(power)):
                Px = Px * (1 - tab[best, 2])
                Py = Py * (1 - tab[best, 3])
        if int(aff["field_mode"]) == 1:
            gx, gy = _tan(Hx * aff["field_arg"]), _tan(Hy * aff["field_arg"])
        else:
            gx, gy = Hx, Hy
        ofx, ofy = aff["origin_field"][0] * gx, aff["origin_field"][1] * gy
        tfx, tfy = aff["target_field"][0] * gx, aff["target_field"][1] * gy
    x0 = Px * aff["origin_scale"][0] + aff["origin0"][0] + ofx
    y0 = Py * aff["origin_scale"][1] + aff["origin0"][1] + ofy
    z0 = Px * 0 + aff["origin0"][2]
    dx = Px * aff["target_scale"][0] + aff["target0"][0] + tfx - x0
    dy = Py * aff["target_scale"][1] + aff["target0"][1] + tfy - y0
    dz = aff["target0"][2] - z0
    mag = _sqrt(dx ** 2 + dy ** 2 + dz ** 2)
    return x0, y0, z0, dx / mag, dy / mag, dz / mag


def pupil_affine_fields(sc: dict, Hx, Hy) -> dict:
    """Per-ray field points (``RealRayTracer.trace_generic``, raytrace/real_ray_tracer.py:120-154): the affine
    form of the H = 0 field plus the field-dependent offsets of origin and target (include/olb.h
    ``OlbPupilLaunch.Hx``).  ``sc`` = ``pack.launch_scalars(optic, 0, 0)`` of an optic WITHOUT vignetting factors
    (they would make the pupil scale field-dependent).  ``Hx``, ``Hy``: arrays of the kernel's element type."""
    if sc["vx"] != 1.0 or sc["vy"] != 1.0:
        raise ValueError("per-ray fields need an optic without vignetting factors")
    if int(sc.get("mode", 0)) == 3 or float(sc.get("field_kind", 2.0)) == 0.0:
        raise ValueError("per-ray fields need an angle or object-height field (image-height fields: one field point per launch)")
    base = pupil_affine({**sc, "Hx": 0.0, "Hy": 0.0} if int(sc.get("mode", 0)) == 0 else sc)
    mode = int(sc.get("mode", 0))
    aff = dict(base)
    aff["fields"] = (Hx, Hy)
    if mode == 0:                      # infinite object, angle field: origin slides by -tan(field) (offset + EPL)
        k = -(sc["offset"] + sc["EPL"])
        aff.update(field_mode=1, field_arg=math.radians(sc["max_field"]), origin_field=(k, k), target_field=(0.0, 0.0))
    elif float(sc.get("field_kind", 2.0)) == 1.0:   # finite object, angle field (angle.py:49-58): x0 = -tan(field) (EPL - z0)
        k = -(sc["EPL"] - sc["z0"])
        aff.update(field_mode=1, field_arg=math.radians(sc["max_field"]), origin_field=(k, k),
                   target_field=(k, k) if mode == 2 else (0.0, 0.0))
    else:                              # object-height field (object_height.py:37-44): x0 = max_field * Hx
        k = sc["max_field"]
        aff.update(field_mode=2, field_arg=0.0, origin_field=(k, k), target_field=(k, k) if mode == 2 else (0.0, 0.0))
    return aff
