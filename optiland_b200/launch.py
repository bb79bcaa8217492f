"""Launch state of an infinite-object angle field as a closed form of (Px, Py).

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
