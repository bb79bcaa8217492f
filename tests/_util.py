"""Shared helpers for the test-suite: golden fixtures and comparison metrics."""
from __future__ import annotations

import glob
import os

import numpy as np

from optiland_b200.table import SurfaceTable

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REC = ("x", "y", "z", "L", "M", "N", "intensity", "opd")
RAY_IN = ("x", "y", "z", "L", "M", "N", "i", "w")

ALL_CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                   if not p.endswith("_grad.npz") and not p.endswith("_ref.npz"))
POLARIZED_CASES = [c for c in ALL_CASES if "polarized" in c]
ERROR_CASES = [c for c in ALL_CASES if "error" in c]
REAL_CASES = [c for c in ALL_CASES if c not in POLARIZED_CASES and c not in ERROR_CASES]


class Case:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
        self.table = SurfaceTable.from_arrays(self.z)
        self.rays = {k: self.z["in_" + k] for k in RAY_IN}
        self.n = self.rays["x"].size
        self.error = str(self.z["ref_error"])
        if not self.error:
            self.rec = {k: self.z["rec_" + k] for k in REC}
            self.out = {k: self.z["out_" + k] for k in ("x", "y", "z", "L", "M", "N", "i", "opd", "L0", "M0", "N0")}
            if "out_p" in self.z:
                self.out["p"] = self.z["out_p"]

    def extra(self, key):
        return self.z["x_" + key]

    @property
    def scale(self):
        """Characteristic length (mm) of the system: tolerances scale with it."""
        s = 1.0
        for k in ("x", "y", "z", "opd"):
            a = self.rec[k]
            a = a[np.isfinite(a)]
            if a.size:
                s = max(s, float(np.max(np.abs(a))))
        return s


def max_abs_err(a, b):
    """max |a-b| over entries where the reference is finite; NaN patterns must agree."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), f"NaN pattern differs: {nan_a.sum()} vs {nan_b.sum()}"
    m = ~nan_b
    if not m.any():
        return 0.0
    return float(np.max(np.abs(a[m] - b[m])))


def f32_bounds(name):
    """What the fp32 arithmetic ACHIEVES on fixture ``name`` against the reference's fp64 records
    (tests/golden/f32_achieved.json, measured by scripts/f32_achieved.py on the CPU instantiation of the device math
    and on the B200 kernel; the larger of the two): max |error| of intercepts (mm), OPD (mm), direction cosines,
    intensity, P-matrix entries.  The parity tests assert <= 3x these numbers -- not a scale-based guess."""
    import json

    with open(os.path.join(GOLDEN, "f32_achieved.json")) as f:
        return json.load(f)["cases"][name]


def fp32_errors(rec, want):
    """max |error| per group over entries finite in both; fp32 may turn a grazing ray into NaN where fp64 does not
    (and vice versa): at most 2 % such disagreements."""
    out = {}
    for tag, keys in (("pos", ("x", "y", "z")), ("opd", ("opd",)), ("dir", ("L", "M", "N")), ("intensity", ("intensity",))):
        w = 0.0
        for k in keys:
            a, b = np.asarray(rec[k], dtype=np.float64), want[k]
            assert np.mean(np.isfinite(a) != np.isfinite(b)) <= 0.02, k
            m = np.isfinite(a) & np.isfinite(b)
            if m.any():
                w = max(w, float(np.max(np.abs(a[m] - b[m]))))
        out[tag] = w
    return out
