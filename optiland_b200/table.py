"""Host-side surface table: the data contract between Optiland objects and libolb.

A ``SurfaceTable`` is the flattened, plain-data description of what the hot path
reads from a live ``SurfaceGroup`` (reference: SURVEY.md Appendix B; the
attributes read by ``Surface._trace_real``,
``/root/reference/optiland/surfaces/standard_surface.py:232-248``).  It packs into
the ``OlbSurface`` array + ``pool`` of ``include/olb.h``.

Nothing here touches a GPU; the module is shared by the product path
(``optiland_b200.trace``), the Optiland plugin (``optiland_b200.plugin``) and the
test oracle.
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from typing import Sequence

import numpy as np

# ---- enums: keep in sync with include/olb.h --------------------------------
GEOM_NOOP = 0
GEOM_PLANE = 1
GEOM_STANDARD = 2
GEOM_EVEN_ASPHERE = 3
GEOM_ZERNIKE = 4
GEOM_ODD_ASPHERE = 5
GEOM_POLYNOMIAL = 6
GEOM_CHEBYSHEV = 7
GEOM_BICONIC = 8
GEOM_TOROIDAL = 9
GEOM_FORBES_QBFS = 10

SF_REFLECT = 1 << 0
SF_ROTATED = 1 << 1
SF_APERTURE = 1 << 2
SF_ABSORBING = 1 << 3
SF_NORECORD = 1 << 4

COAT_NONE = 0
COAT_SIMPLE = 1
COAT_FRESNEL = 2

AP_RADIAL = 1
AP_OFFSET_RADIAL = 2
AP_RECT = 3
AP_ELLIPSE = 4
AP_UNION = 16
AP_INTERSECT = 17
AP_DIFFERENCE = 18
_AP_OPERANDS = {AP_RADIAL: 2, AP_OFFSET_RADIAL: 4, AP_RECT: 4, AP_ELLIPSE: 4,
                AP_UNION: 0, AP_INTERSECT: 0, AP_DIFFERENCE: 0}

TF_POLARIZED = 1 << 0
ST_ZERNIKE_RANGE = 1 << 0
ST_CHEBYSHEV_RANGE = 1 << 1
ST_K_PARALLEL_X = 1 << 2

MAX_SURFACES = 64
MAX_WAVELENGTHS = 16

NEWTON_KINDS = (GEOM_EVEN_ASPHERE, GEOM_ZERNIKE, GEOM_ODD_ASPHERE, GEOM_POLYNOMIAL, GEOM_CHEBYSHEV, GEOM_BICONIC,
                GEOM_TOROIDAL, GEOM_FORBES_QBFS)

# numpy mirror of `struct OlbSurface` (192 bytes)
OLB_SURFACE_DTYPE = np.dtype(
    [
        ("kind", "<i4"), ("flags", "<u4"), ("n_coef", "<i4"), ("coef_off", "<i4"),
        ("aper_off", "<i4"), ("aper_len", "<i4"), ("max_iter", "<i4"), ("coating", "<i4"),
        ("media_off", "<i4"), ("aux0", "<i4"), ("reserved", "<i4", (2,)),
        ("t", "<f8", (3,)), ("R", "<f8", (9,)),
        ("radius", "<f8"), ("conic", "<f8"), ("tol", "<f8"),
        ("coat_t", "<f8"), ("coat_r", "<f8"), ("norm_radius", "<f8"),
    ],
    align=False,
)
assert OLB_SURFACE_DTYPE.itemsize == 192


@dataclass
class SurfaceSpec:
    """One surface as the trace loop sees it.  Arrays are fp64 numpy."""

    kind: int = GEOM_PLANE
    t: np.ndarray = field(default_factory=lambda: np.zeros(3))
    R: np.ndarray = field(default_factory=lambda: np.eye(3))
    radius: float = float("inf")
    conic: float = 0.0
    tol: float = 1e-10
    max_iter: int = 100
    # EVEN/ODD: (n,) ; POLYNOMIAL: (rows, cols) ; ZERNIKE: (n_terms, 4) {n, m, c*N_nm, c}
    coefficients: np.ndarray = field(default_factory=lambda: np.zeros(0))
    norm_radius: float = 1.0   # Zernike norm_radius; Chebyshev norm_x
    norm_y: float = 1.0        # Chebyshev norm_y
    radius_y: float = float("inf")  # biconic Ry; toroidal: radius of rotation R_rot
    conic_y: float = 0.0            # biconic ky; toroidal: conic of the Y-Z curve
    reflective: bool = False
    aperture: np.ndarray | None = None  # postfix program, see include/olb.h
    n1: np.ndarray = field(default_factory=lambda: np.ones(1))
    n2: np.ndarray = field(default_factory=lambda: np.ones(1))
    k1: np.ndarray = field(default_factory=lambda: np.zeros(1))
    coating: int = COAT_NONE
    coat_t: float = 1.0
    coat_r: float = 0.0
    coat_n1: np.ndarray | None = None
    coat_n2: np.ndarray | None = None
    record: bool = True
    # ZERNIKE only, host side only (not packed): the normalisation constants N_nm per term, for mapping table
    # gradients back to the coefficients when a coefficient is exactly 0 (c * N_nm then does not reveal N_nm)
    zernike_norms: np.ndarray | None = None

    def __post_init__(self):
        self.t = np.asarray(self.t, dtype=np.float64).reshape(3)
        self.R = np.asarray(self.R, dtype=np.float64).reshape(3, 3)
        self.coefficients = np.asarray(self.coefficients, dtype=np.float64)
        self.n1 = np.atleast_1d(np.asarray(self.n1, dtype=np.float64))
        self.n2 = np.atleast_1d(np.asarray(self.n2, dtype=np.float64))
        self.k1 = np.atleast_1d(np.asarray(self.k1, dtype=np.float64))
        if self.aperture is not None:
            self.aperture = np.asarray(self.aperture, dtype=np.float64).ravel()
        if self.coat_n1 is not None:
            self.coat_n1 = np.atleast_1d(np.asarray(self.coat_n1, dtype=np.float64))
        if self.coat_n2 is not None:
            self.coat_n2 = np.atleast_1d(np.asarray(self.coat_n2, dtype=np.float64))

    @property
    def rotated(self) -> bool:
        return not np.array_equal(self.R, np.eye(3))

    @property
    def absorbing(self) -> bool:
        return bool(np.any(self.k1 > 0))

    @property
    def flags(self) -> int:
        f = 0
        if self.reflective:
            f |= SF_REFLECT
        if self.rotated:
            f |= SF_ROTATED
        if self.aperture is not None:
            f |= SF_APERTURE
        if self.absorbing:
            f |= SF_ABSORBING
        if not self.record:
            f |= SF_NORECORD
        return f


def validate_aperture_program(prog: np.ndarray) -> None:
    """Check a postfix aperture program is well formed (stack depth ends at 1)."""
    i, depth = 0, 0
    n = len(prog)
    while i < n:
        op = int(prog[i])
        if op not in _AP_OPERANDS or prog[i] != op:
            raise ValueError(f"bad aperture opcode {prog[i]!r} at {i}")
        nops = _AP_OPERANDS[op]
        if op >= AP_UNION:
            if depth < 2:
                raise ValueError("aperture program stack underflow")
            depth -= 1
        else:
            depth += 1
            if depth > 8:
                raise ValueError("aperture program too deep (max 8)")
        i += 1 + nops
    if i != n or depth != 1:
        raise ValueError("malformed aperture program")


@dataclass
class SurfaceTable:
    """All surfaces of a system + the distinct wavelengths their media are tabulated at."""

    surfaces: list[SurfaceSpec]
    wavelengths: np.ndarray  # (n_wl,) micrometres; exact values that appear in rays.w

    def __post_init__(self):
        self.wavelengths = np.atleast_1d(np.asarray(self.wavelengths, dtype=np.float64))
        n_wl = len(self.wavelengths)
        if not 1 <= n_wl <= MAX_WAVELENGTHS:
            raise ValueError(f"n_wl must be in [1, {MAX_WAVELENGTHS}], got {n_wl}")
        if not 1 <= len(self.surfaces) <= MAX_SURFACES:
            raise ValueError(f"number of surfaces must be in [1, {MAX_SURFACES}]")
        for s in self.surfaces:
            for name in ("n1", "n2", "k1"):
                if len(getattr(s, name)) != n_wl:
                    raise ValueError(f"surface media '{name}' must have {n_wl} entries")
            if s.aperture is not None:
                validate_aperture_program(s.aperture)
            if s.coating == COAT_FRESNEL and (s.coat_n1 is None or s.coat_n2 is None):
                raise ValueError("Fresnel coating needs coat_n1/coat_n2")

    @property
    def num_surfaces(self) -> int:
        return len(self.surfaces)

    @property
    def n_wl(self) -> int:
        return len(self.wavelengths)

    # ---- packing into the C ABI layout ------------------------------------
    def pack(self) -> tuple[np.ndarray, np.ndarray]:
        """Return (surfaces: OLB_SURFACE_DTYPE[n], pool: float64[m])."""
        n_wl = self.n_wl
        n = len(self.surfaces)
        surf = np.zeros(n, dtype=OLB_SURFACE_DTYPE)
        pool: list[float] = []

        def push(values) -> int:
            off = len(pool)
            pool.extend(np.asarray(values, dtype=np.float64).ravel().tolist())
            # keep every block 16-byte aligned for vector loads
            if len(pool) % 2:
                pool.append(0.0)
            return off

        # integer / offset columns are filled per surface, then every struct field is assigned ONCE for all
        # surfaces (a per-element structured assignment costs ~1 us; this runs on every plugin call)
        ints = {k: [0] * n for k in ("n_coef", "aux0", "coef_off", "aper_off", "aper_len", "media_off")}
        for j, s in enumerate(self.surfaces):
            coef = s.coefficients
            extra_head = None
            if s.kind in (GEOM_POLYNOMIAL, GEOM_CHEBYSHEV):
                coef = np.atleast_2d(coef)
                ints["n_coef"][j] = coef.size
                ints["aux0"][j] = coef.shape[1]
                if s.kind == GEOM_CHEBYSHEV:
                    extra_head = [s.norm_radius, s.norm_y]
            elif s.kind in (GEOM_BICONIC, GEOM_TOROIDAL):
                ints["n_coef"][j] = coef.size
                extra_head = [s.radius_y, s.conic_y]
            elif s.kind == GEOM_ZERNIKE:
                coef = coef.reshape(-1, 4)
                ints["n_coef"][j] = coef.shape[0]
            else:
                ints["n_coef"][j] = coef.size
            if extra_head is not None:
                ints["coef_off"][j] = push(np.concatenate([np.asarray(extra_head, float), coef.ravel()]))
            else:
                ints["coef_off"][j] = push(coef) if coef.size else 0
            if s.aperture is not None:
                ints["aper_off"][j] = push(s.aperture)
                ints["aper_len"][j] = len(s.aperture)
            cn1 = s.coat_n1 if s.coat_n1 is not None else s.n1
            cn2 = s.coat_n2 if s.coat_n2 is not None else s.n2
            ints["media_off"][j] = push(np.concatenate([s.n1, s.n2, s.k1, cn1, cn2]))
            assert len(s.n1) == n_wl
        for k, v in ints.items():
            surf[k] = v
        sl = self.surfaces
        surf["kind"] = [s.kind for s in sl]
        surf["flags"] = [s.flags for s in sl]
        surf["max_iter"] = [s.max_iter for s in sl]
        surf["coating"] = [s.coating for s in sl]
        if n:
            surf["t"] = np.stack([s.t for s in sl])
            surf["R"] = np.stack([s.R.ravel() for s in sl])
        for k, attr in (("radius", "radius"), ("conic", "conic"), ("tol", "tol"), ("coat_t", "coat_t"),
                        ("coat_r", "coat_r"), ("norm_radius", "norm_radius")):
            surf[k] = [getattr(s, attr) for s in sl]
        if not pool:
            pool.append(0.0)
            pool.append(0.0)
        return surf, np.asarray(pool, dtype=np.float64)

    def packed(self):
        """``pack()`` computed once per table object (a table is not mutated after it is built)."""
        pk = self.__dict__.get("_packed")
        if pk is None:
            pk = self.pack()
            self.__dict__["_packed"] = pk
        return pk

    def content_key(self) -> bytes:
        """Every value the kernels read, as bytes: equal keys <=> identical prepared tables."""
        key = self.__dict__.get("_content_key")
        if key is None:
            surf, pool = self.packed()
            key = surf.tobytes() + pool.tobytes() + np.asarray(self.wavelengths, dtype=np.float64).tobytes()
            self.__dict__["_content_key"] = key
        return key

    # ---- (de)serialisation for the golden fixtures ------------------------
    def to_arrays(self, prefix: str = "tab_") -> dict[str, np.ndarray]:
        surf, pool = self.pack()
        return {
            prefix + "surfaces": surf.view(np.uint8).reshape(len(surf), -1),
            prefix + "pool": pool,
            prefix + "wavelengths": self.wavelengths,
        }

    @classmethod
    def from_arrays(cls, arrays, prefix: str = "tab_") -> "SurfaceTable":
        raw = np.ascontiguousarray(arrays[prefix + "surfaces"], dtype=np.uint8)
        surf = raw.reshape(-1).view(OLB_SURFACE_DTYPE)
        pool = np.asarray(arrays[prefix + "pool"], dtype=np.float64)
        wl = np.asarray(arrays[prefix + "wavelengths"], dtype=np.float64)
        return cls.unpack(surf, pool, wl)

    @classmethod
    def unpack(cls, surf: np.ndarray, pool: np.ndarray, wavelengths: np.ndarray) -> "SurfaceTable":
        n_wl = len(wavelengths)
        specs = []
        for r in surf:
            kind = int(r["kind"])
            n_coef = int(r["n_coef"])
            off = int(r["coef_off"])
            head = None
            if kind == GEOM_ZERNIKE:
                coef = pool[off: off + 4 * n_coef].reshape(-1, 4).copy()
            elif kind == GEOM_POLYNOMIAL:
                cols = max(int(r["aux0"]), 1)
                coef = pool[off: off + n_coef].reshape(-1, cols).copy()
            elif kind == GEOM_CHEBYSHEV:
                cols = max(int(r["aux0"]), 1)
                head = pool[off: off + 2].copy()
                coef = pool[off + 2: off + 2 + n_coef].reshape(-1, cols).copy()
            elif kind in (GEOM_BICONIC, GEOM_TOROIDAL):
                head = pool[off: off + 2].copy()
                coef = pool[off + 2: off + 2 + n_coef].copy()
            else:
                coef = pool[off: off + n_coef].copy()
            flags = int(r["flags"])
            aper = None
            if flags & SF_APERTURE:
                a0 = int(r["aper_off"])
                aper = pool[a0: a0 + int(r["aper_len"])].copy()
            m0 = int(r["media_off"])
            media = pool[m0: m0 + 5 * n_wl].reshape(5, n_wl)
            coating = int(r["coating"])
            specs.append(
                SurfaceSpec(
                    kind=kind, t=r["t"].copy(), R=r["R"].reshape(3, 3).copy(),
                    radius=float(r["radius"]), conic=float(r["conic"]), tol=float(r["tol"]),
                    max_iter=int(r["max_iter"]), coefficients=coef,
                    norm_radius=float(head[0]) if kind == GEOM_CHEBYSHEV else float(r["norm_radius"]),
                    norm_y=float(head[1]) if kind == GEOM_CHEBYSHEV else 1.0,
                    radius_y=float(head[0]) if kind in (GEOM_BICONIC, GEOM_TOROIDAL) else float("inf"),
                    conic_y=float(head[1]) if kind in (GEOM_BICONIC, GEOM_TOROIDAL) else 0.0,
                    reflective=bool(flags & SF_REFLECT), aperture=aper,
                    n1=media[0].copy(), n2=media[1].copy(), k1=media[2].copy(),
                    coating=coating, coat_t=float(r["coat_t"]), coat_r=float(r["coat_r"]),
                    coat_n1=media[3].copy() if coating == COAT_FRESNEL else None,
                    coat_n2=media[4].copy() if coating == COAT_FRESNEL else None,
                    record=not (flags & SF_NORECORD),
                )
            )
        return cls(specs, wavelengths)

    def replace_surface(self, index: int, **changes) -> "SurfaceTable":
        surfaces = list(self.surfaces)
        surfaces[index] = dataclasses.replace(surfaces[index], **changes)
        return SurfaceTable(surfaces, self.wavelengths)


# ---- small constructors used by tests / bench (no Optiland needed) ---------

def aperture_radial(r_max: float, r_min: float = 0.0) -> np.ndarray:
    return np.array([AP_RADIAL, r_max, r_min], dtype=np.float64)


def aperture_offset_radial(r_max, r_min, dx, dy) -> np.ndarray:
    return np.array([AP_OFFSET_RADIAL, r_max, r_min, dx, dy], dtype=np.float64)


def aperture_rect(x_min, x_max, y_min, y_max) -> np.ndarray:
    return np.array([AP_RECT, x_min, x_max, y_min, y_max], dtype=np.float64)


def aperture_ellipse(a, b, dx=0.0, dy=0.0) -> np.ndarray:
    return np.array([AP_ELLIPSE, a, b, dx, dy], dtype=np.float64)


def aperture_combine(op: int, a: Sequence[float], b: Sequence[float]) -> np.ndarray:
    return np.concatenate([np.asarray(a, float), np.asarray(b, float), [float(op)]])


def rotation_matrix(rx: float, ry: float, rz: float) -> np.ndarray:
    """R = Rz @ Ry @ Rx  (reference: optiland/coordinate_system.py:121-143)."""
    cx, sx = np.cos(rx), np.sin(rx)
    cy, sy = np.cos(ry), np.sin(ry)
    cz, sz = np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def zernike_monomials(n: int, m: int, width: int) -> np.ndarray:
    """Monomial expansion of the UNIT Zernike term R_n^|m|(rho) {cos, sin}(|m| phi) (m >= 0: cos, m < 0: sin) as a
    (width, width) table M[i, j] ~ xn^i yn^j -- the Python twin of olb_prep.h::zernike_add_monomials (exact integer
    arithmetic in doubles).  The prepared sag table of a Zernike surface is S = sum_k c_k N_k M_k and its slope table
    D = sum_k c_k M_k (the reference's derivative path omits N_k), so table gradients map back to the coefficients by
    dL/dc_k = N_k <M_k, dL/dS> + <M_k, dL/dD>  (optiland_b200.autograd)."""
    from math import comb, factorial

    ma = abs(m)
    out = np.zeros((width, width), dtype=np.float64)
    for k in range((n - ma) // 2 + 1):
        rc = (-1.0 if k & 1 else 1.0) * factorial(n - k) / (factorial(k) * factorial((n + ma) // 2 - k) * factorial((n - ma) // 2 - k))
        q = (n - ma) // 2 - k
        for a in range(q + 1):
            ca = comb(q, a)
            for j in range(ma + 1):
                imag = (j & 1) != 0
                if (m >= 0) == imag:
                    continue
                sgn = -1.0 if (j >> 1) & 1 else 1.0
                out[2 * a + (ma - j), 2 * (q - a) + j] += rc * ca * comb(ma, j) * sgn
    return out
