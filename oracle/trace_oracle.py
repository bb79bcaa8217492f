"""TEST INFRASTRUCTURE ONLY -- CPU restatement (NumPy, fp64) of Optiland's real-ray trace loop.

This is the parity oracle for libolb.  It is NOT part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (its parity gate, and the
``cpu_baseline_port`` / fallback ``--impl reference`` legs) may import it; the product path
(``optiland_b200.*``) never does and fails loudly without the CUDA library.

Parity is PINNED: ``tests/test_oracle_golden.py`` checks this restatement against the
committed fixtures in ``tests/golden/``, which were produced by the unmodified reference
(``oracle/make_golden.py``, run in the build container where /root/reference is
importable), and against the reference's own hard-coded known-answer vectors
(tests/test_geometries.py etc.); ``tests/test_plugin_reference.py[oracle]`` runs it
under the live reference's own objects (test-only engine, ``oracle/oracle_engine.py``)
and compares with the reference's NumPy backend call by call.

Each function cites the reference file:line (relative to /root/reference) whose
arithmetic it restates, operation for operation, so that fp64 results agree with
the reference NumPy backend to rounding.  It consumes the same plain-data
``SurfaceTable`` the CUDA path consumes.
"""
from __future__ import annotations

import math

import numpy as np

from optiland_b200 import table as T


# --------------------------------------------------------------------------
# geometry: base conic
# --------------------------------------------------------------------------

def conic_distance(x, y, z, L, M, N, radius, k):
    """optiland/geometries/standard.py:97-148 (StandardGeometry.distance)."""
    if math.isinf(radius):
        N_safe = np.where(np.abs(N) > 1e-14, N, 1e-14)
        return -z / N_safe
    a = k * N**2 + L**2 + M**2 + N**2
    b = 2 * k * N * z + 2 * L * x + 2 * M * y - 2 * N * radius + 2 * N * z
    c = k * z**2 - 2 * radius * z + x**2 + y**2 + z**2
    d = b**2 - 4 * a * c
    with np.errstate(all="ignore"):
        t1 = (-b + np.sqrt(d)) / (2 * a)
        t2 = (-b - np.sqrt(d)) / (2 * a)
        z1 = z + t1 * N
        z2 = z + t2 * N
        t = np.where(np.abs(z1) <= np.abs(z2), t1, t2)
        t = np.where(a == 0, -c / b, t)
    return t


def plane_distance(z, N):
    """optiland/geometries/plane.py:72-88."""
    with np.errstate(all="ignore"):
        return -z / N


def conic_sag(x, y, radius, k):
    """optiland/geometries/standard.py:80-95."""
    r2 = x**2 + y**2
    with np.errstate(all="ignore"):
        return r2 / (radius * (1 + np.sqrt(1 - (1 + k) * r2 / radius**2)))


def conic_normal(x, y, radius, k):
    """optiland/geometries/standard.py:150-175."""
    r2 = x**2 + y**2
    with np.errstate(all="ignore"):
        denom = radius * np.sqrt(1 - (1 + k) * r2 / radius**2)
        dfdx = x / denom
        dfdy = y / denom
        dfdz = -1
        mag = np.sqrt(dfdx**2 + dfdy**2 + dfdz**2)
        return dfdx / mag, dfdy / mag, dfdz / mag


# --------------------------------------------------------------------------
# geometry: Newton-Raphson family
# --------------------------------------------------------------------------

def even_sag(x, y, radius, k, coefs):
    """optiland/geometries/even_asphere.py:93-109."""
    r2 = x**2 + y**2
    with np.errstate(all="ignore"):
        z = r2 / (radius * (1 + np.sqrt(1 - (1 + k) * r2 / radius**2)))
    for i, Ci in enumerate(coefs):
        z = z + Ci * r2 ** (i + 1)
    return z


def even_normal(x, y, radius, k, coefs):
    """optiland/geometries/even_asphere.py:111-140."""
    r2 = x**2 + y**2
    with np.errstate(all="ignore"):
        denom = radius * np.sqrt(1 - (1 + k) * r2 / radius**2)
        dfdx = x / denom
        dfdy = y / denom
        for i, Ci in enumerate(coefs):
            dfdx = dfdx + 2 * (i + 1) * x * Ci * r2**i
            dfdy = dfdy + 2 * (i + 1) * y * Ci * r2**i
        mag = np.sqrt(dfdx**2 + dfdy**2 + 1)
        return dfdx / mag, dfdy / mag, -1 / mag


def odd_sag(x, y, radius, k, coefs):
    """optiland/geometries/odd_asphere.py:86-101."""
    r2 = np.asarray(x**2 + y**2)
    r = np.sqrt(r2)
    with np.errstate(all="ignore"):
        z = r2 / (radius * (1 + np.sqrt(1 - (1 + k) * r2 / radius**2)))
    for i, Ci in enumerate(coefs):
        z = z + Ci * r ** (i + 1)
    return z


def odd_normal(x, y, radius, k, coefs):
    """optiland/geometries/odd_asphere.py:103-142."""
    r2 = x**2 + y**2
    r = np.sqrt(r2)
    with np.errstate(all="ignore"):
        denom = radius * np.sqrt(1 - (1 + k) * r2 / radius**2)
        dfdx = x / denom
        dfdy = y / denom
        for i, Ci in enumerate(coefs):
            x_term = np.asarray((i + 1) * x * Ci * r ** (i - 1), dtype=np.float64).copy()
            y_term = np.asarray((i + 1) * y * Ci * r ** (i - 1), dtype=np.float64).copy()
            x_term[~np.isfinite(x_term)] = 0
            y_term[~np.isfinite(y_term)] = 0
            dfdx = dfdx + x_term
            dfdy = dfdy + y_term
        mag = np.sqrt(dfdx**2 + dfdy**2 + 1)
        return dfdx / mag, dfdy / mag, -1 / mag


def poly_sag(x, y, radius, k, C):
    """optiland/geometries/polynomial.py:105-121."""
    r2 = x**2 + y**2
    with np.errstate(all="ignore"):
        z = r2 / (radius * (1 + np.sqrt(1 - (1 + k) * r2 / radius**2)))
    for i in range(C.shape[0]):
        for j in range(C.shape[1]):
            z = z + C[i, j] * (x**i) * (y**j)
    return z


def poly_normal(x, y, radius, k, C):
    """optiland/geometries/polynomial.py:123-155."""
    r2 = x**2 + y**2
    with np.errstate(all="ignore"):
        denom = radius * np.sqrt(1 - (1 + k) * r2 / radius**2)
        dzdx = x / denom
        dzdy = y / denom
        for i in range(1, C.shape[0]):
            for j in range(C.shape[1]):
                dzdx = dzdx + i * C[i, j] * (x ** (i - 1)) * (y**j)
        for i in range(C.shape[0]):
            for j in range(1, C.shape[1]):
                dzdy = dzdy + j * C[i, j] * (x**i) * (y ** (j - 1))
        norm = np.sqrt(dzdx**2 + dzdy**2 + 1)
        return dzdx / norm, dzdy / norm, -1 / norm


# ---- Chebyshev / biconic / toroidal ------------------------------------------

def _cheb(n, x):
    """optiland/geometries/chebyshev.py:193-204."""
    return np.cos(n * np.arccos(x))


def _cheb_d(n, x):
    """optiland/geometries/chebyshev.py:206-228."""
    return n * np.sin(n * np.arccos(x)) / np.sqrt(1 - x**2)


def cheb_sag(x, y, radius, k, C, norm_x, norm_y, status):
    """optiland/geometries/chebyshev.py:126-150."""
    x_norm, y_norm = x / norm_x, y / norm_y
    with np.errstate(all="ignore"):
        if np.any(np.abs(x_norm) > 1) or np.any(np.abs(y_norm) > 1):
            status[0] |= T.ST_CHEBYSHEV_RANGE  # reference raises ValueError (chebyshev.py:230-244)
        r2 = x**2 + y**2
        z = r2 / (radius * (1 + np.sqrt(1 - (1 + k) * r2 / radius**2)))
        for i, j in np.argwhere(C != 0):
            z = z + C[i, j] * _cheb(i, x_norm) * _cheb(j, y_norm)
    return z


def cheb_normal(x, y, radius, k, C, norm_x, norm_y, status):
    """optiland/geometries/chebyshev.py:152-191.  (The chain-rule factor 1/norm is NOT applied by the
    reference: d/dx of T_i(x/norm_x) is taken as T_i'(x/norm_x); reproduced.)"""
    x_norm, y_norm = x / norm_x, y / norm_y
    with np.errstate(all="ignore"):
        if np.any(np.abs(x_norm) > 1) or np.any(np.abs(y_norm) > 1):
            status[0] |= T.ST_CHEBYSHEV_RANGE
        r2 = x**2 + y**2
        denom = radius * np.sqrt(1 - (1 + k) * r2 / radius**2)
        dzdx = x / denom
        dzdy = y / denom
        for i, j in np.argwhere(C != 0):
            dzdx = dzdx + (_cheb_d(i, x_norm) * C[i, j] * _cheb(j, y_norm))
            dzdy = dzdy + (_cheb_d(j, y_norm) * C[i, j] * _cheb(i, x_norm))
        norm = np.sqrt(dzdx**2 + dzdy**2 + 1)
        return dzdx / norm, dzdy / norm, -1 / norm


def _curv(R):
    return 0.0 if (math.isinf(R) or R == 0) else 1.0 / R


def biconic_sag(x, y, Rx, kx, Ry, ky):
    """optiland/geometries/biconic.py:72-105."""
    cx, cy = _curv(Rx), _curv(Ry)
    zx, zy = np.zeros_like(x), np.zeros_like(y)
    with np.errstate(all="ignore"):
        if cx != 0:
            v = 1.0 - (1.0 + kx) * cx**2 * x**2
            rt = np.where(v < 1e-14, 0.0, v)
            den = 1.0 + np.sqrt(rt)
            zx = (cx * x**2) / np.where(np.abs(den) < 1e-14, 1e-14, den)
        if cy != 0:
            v = 1.0 - (1.0 + ky) * cy**2 * y**2
            rt = np.where(v < 1e-14, 0.0, v)
            den = 1.0 + np.sqrt(rt)
            zy = (cy * y**2) / np.where(np.abs(den) < 1e-14, 1e-14, den)
    return zx + zy


def biconic_normal(x, y, Rx, kx, Ry, ky):
    """optiland/geometries/biconic.py:107-160."""
    cx, cy = _curv(Rx), _curv(Ry)
    with np.errstate(all="ignore"):
        if cx == 0:
            dfdx = np.zeros_like(x)
        else:
            v = 1.0 - (1.0 + kx) * cx**2 * x**2
            sq = np.sqrt(np.where(v < 1e-14, 1e-14, v))
            dfdx = (cx * x) / np.where(np.abs(sq) < 1e-14, 1e-14, sq)
        if cy == 0:
            dfdy = np.zeros_like(y)
        else:
            v = 1.0 - (1.0 + ky) * cy**2 * y**2
            sq = np.sqrt(np.where(v < 1e-14, 1e-14, v))
            dfdy = (cy * y) / np.where(np.abs(sq) < 1e-14, 1e-14, sq)
        mag = np.sqrt(dfdx**2 + dfdy**2 + 1.0)
        mag = np.where(mag < 1e-14, 1.0, mag)
        return dfdx / mag, dfdy / mag, -1.0 / mag


def _tor_zy(y, R_yz, k_yz, coefs):
    """optiland/geometries/toroidal.py:87-121."""
    y2 = y**2
    z_y = np.zeros_like(y)
    with np.errstate(all="ignore"):
        if math.isfinite(R_yz) and R_yz != 0:
            c = 1.0 / R_yz
            v = 1.0 - (1.0 + k_yz) * c**2 * y2
            den = 1.0 + np.sqrt(np.where(v < 0, 0.0, v))
            z_y = (c * y2) / np.where(np.abs(den) < 1e-14, 1e-14, den)
        if len(coefs) > 0:
            poly, p = np.zeros_like(y), y2
            for coeff in coefs:
                poly = poly + coeff * p
                p = p * y2
            z_y = z_y + poly
    return z_y


def _tor_dzy(y, R_yz, k_yz, coefs):
    """optiland/geometries/toroidal.py:123-160."""
    y2 = y**2
    dz = np.zeros_like(y)
    with np.errstate(all="ignore"):
        if math.isfinite(R_yz) and R_yz != 0:
            c = 1.0 / R_yz
            v = 1.0 - (1.0 + k_yz) * c**2 * y2
            sq = np.sqrt(np.where(v < 1e-14, 1e-14, v))
            dz = (c * y) / np.where(np.abs(sq) < 1e-14, 1e-14, sq)
        if len(coefs) > 0:
            poly, p = np.zeros_like(y), y
            for i, coeff in enumerate(coefs):
                poly = poly + coeff * (2.0 * (i + 1.0)) * p
                p = p * y2
            dz = dz + poly
    return dz


def toroidal_sag(x, y, R_rot, R_yz, k_yz, coefs):
    """optiland/geometries/toroidal.py:162-186."""
    z_y = _tor_zy(y, R_yz, k_yz, coefs)
    if math.isinf(R_rot):
        return z_y
    with np.errstate(all="ignore"):
        term = (R_rot - z_y) ** 2 - x**2
        return np.where(term < 0, np.nan, z_y + ((R_rot - z_y) - np.sign(R_rot - z_y) * np.sqrt(term)))


def toroidal_normal(x, y, R_rot, R_yz, k_yz, coefs):
    """optiland/geometries/toroidal.py:188-232."""
    z_y = _tor_zy(y, R_yz, k_yz, coefs)
    dz_dy = _tor_dzy(y, R_yz, k_yz, coefs)
    eps = 1e-14
    with np.errstate(all="ignore"):
        if math.isinf(R_rot):
            fx, fy = np.zeros_like(x), dz_dy
            term = np.full_like(x, np.inf)
        else:
            term = (R_rot - z_y) ** 2 - x**2
            valid = term >= 0
            sq = np.sqrt(np.where(valid, term, eps))
            sq = np.where(np.abs(sq) < eps, eps, sq)
            fx = np.where(valid, np.sign(R_rot) * x / sq, 0.0)
            fy = np.where(valid, np.sign(R_rot) * (R_rot - z_y) * dz_dy / sq, 0.0)
        mag = np.sqrt(fx**2 + fy**2 + 1.0)
        mag = np.where(mag < eps, 1.0, mag)
        nx, ny, nz = fx / mag, fy / mag, -1.0 / mag
        ok = term >= 0
        return np.where(ok, nx, 0.0), np.where(ok, ny, 0.0), np.where(ok, nz, -1.0)


# ---- Zernike ---------------------------------------------------------------

def _fact(n: int) -> float:
    return float(math.factorial(int(n)))


def zernike_radial(n: int, m: int, r):
    """optiland/zernike/base.py:217-243 (BaseZernike._radial_term)."""
    m_abs = abs(m)
    value = np.zeros_like(r)
    for k in range((n - m_abs) // 2 + 1):
        num = _fact(n - k)
        denom = _fact(k) * _fact((n + m_abs) // 2 - k) * _fact((n - m_abs) // 2 - k)
        coeff = (-1) ** k * num / denom
        value = value + coeff * (r ** (n - 2 * k))
    return value


def zernike_radial_derivative(n: int, m: int, r):
    """optiland/zernike/base.py:264-299 (BaseZernike._radial_derivative)."""
    value = np.zeros_like(r)
    for k in range((n - abs(m)) // 2 + 1):
        numerator = _fact(n - k)
        denominator = _fact(k) * _fact((n + m) // 2 - k) * _fact((n - m) // 2 - k)
        factor = n - 2 * k
        if factor < 0:
            continue
        power_term = r ** (n - 2 * k - 1) if (n - 2 * k - 1) >= 0 else 0
        value = value + (-1) ** k * (numerator / denominator) * factor * power_term
    return value


def zernike_poly(terms, rho, phi):
    """optiland/zernike/base.py:42-99 (get_term / terms / poly); c*N_nm is pre-multiplied
    by the packer, which is the same product ``coeff * _norm_constant`` the reference forms."""
    total = 0
    for n, m, cN, _c in terms:
        n, m = int(n), int(m)
        az = np.cos(m * phi) if m >= 0 else np.sin(abs(m) * phi)
        total = total + cN * zernike_radial(n, m, rho) * az
    return total


def zernike_sag(x, y, radius, k, terms, norm_radius, status):
    """optiland/geometries/zernike.py:153-180."""
    x_norm = x / norm_radius
    y_norm = y / norm_radius
    with np.errstate(all="ignore"):
        if np.any(np.abs(x_norm) > 1) or np.any(np.abs(y_norm) > 1):
            status[0] |= T.ST_ZERNIKE_RANGE  # reference raises ValueError (zernike.py:254-266)
        rho = np.sqrt(x_norm**2 + y_norm**2)
        phi = np.arctan2(y_norm, x_norm)
        r2 = x**2 + y**2
        z = r2 / (radius * (1 + np.sqrt(1 - (1 + k) * r2 / radius**2)))
        z = z + zernike_poly(terms, rho, phi)
    return z


def zernike_normal(x, y, radius, k, terms, norm_radius):
    """optiland/geometries/zernike.py:182-252 (derivatives WITHOUT N_nm: reference quirk)."""
    with np.errstate(all="ignore"):
        r2 = x**2 + y**2
        denominator = radius * np.sqrt(1 - (1 + k) * r2 / radius**2)
        dzdx = x / denominator
        dzdy = y / denominator
        eps = 1e-14
        x_norm = x / norm_radius
        y_norm = y / norm_radius
        rho = np.sqrt(x_norm**2 + y_norm**2)
        phi = np.arctan2(y_norm, x_norm)
        if np.all(rho == 0):
            drho_dx = np.zeros_like(x)
            drho_dy = np.zeros_like(y)
        else:
            drho_dx = (x / (norm_radius**2)) / (rho + eps)
            drho_dy = (y / (norm_radius**2)) / (rho + eps)
        dphi_dx = -(y_norm) / (rho**2 + eps) * (1.0 / norm_radius)
        dphi_dy = +(x_norm) / (rho**2 + eps) * (1.0 / norm_radius)
        for n, m, _cN, c in terms:
            if c == 0:
                continue
            n, m = int(n), int(m)
            # optiland/zernike/base.py:101-137 (get_derivative)
            radial_term = zernike_radial(n, abs(m), rho)
            radial_derivative = zernike_radial_derivative(n, abs(m), rho)
            if m == 0:
                dZdrho, dZdphi = radial_derivative, 0.0
            elif m > 0:
                dZdrho = radial_derivative * np.cos(m * phi)
                dZdphi = -m * radial_term * np.sin(m * phi)
            else:
                dZdrho = radial_derivative * np.sin(abs(m) * phi)
                dZdphi = abs(m) * radial_term * np.cos(abs(m) * phi)
            dzdx = dzdx + c * (dZdrho * drho_dx + dZdphi * dphi_dx)
            dzdy = dzdy + c * (dZdrho * drho_dy + dZdphi * dphi_dy)
        nx, ny = dzdx, dzdy
        norm = np.sqrt(nx**2 + ny**2 + 1)
        norm = np.where(norm < eps, 1.0, norm)
        return nx / norm, ny / norm, -np.ones_like(x) / norm


def _sag_and_normal_fns(s: T.SurfaceSpec, status):
    k, R = s.conic, s.radius
    if s.kind == T.GEOM_EVEN_ASPHERE:
        c = list(s.coefficients)
        return (lambda x, y: even_sag(x, y, R, k, c)), (lambda x, y: even_normal(x, y, R, k, c))
    if s.kind == T.GEOM_ODD_ASPHERE:
        c = list(s.coefficients)
        return (lambda x, y: odd_sag(x, y, R, k, c)), (lambda x, y: odd_normal(x, y, R, k, c))
    if s.kind == T.GEOM_POLYNOMIAL:
        C = np.atleast_2d(s.coefficients)
        return (lambda x, y: poly_sag(x, y, R, k, C)), (lambda x, y: poly_normal(x, y, R, k, C))
    if s.kind == T.GEOM_CHEBYSHEV:
        C = np.atleast_2d(s.coefficients)
        nx_, ny_ = s.norm_radius, s.norm_y
        return (lambda x, y: cheb_sag(x, y, R, k, C, nx_, ny_, status)), (lambda x, y: cheb_normal(x, y, R, k, C, nx_, ny_, status))
    if s.kind == T.GEOM_BICONIC:
        Ry, ky = s.radius_y, s.conic_y
        return (lambda x, y: biconic_sag(x, y, R, k, Ry, ky)), (lambda x, y: biconic_normal(x, y, R, k, Ry, ky))
    if s.kind == T.GEOM_TOROIDAL:
        Rr, kyz, cf = s.radius_y, s.conic_y, list(s.coefficients)
        return (lambda x, y: toroidal_sag(x, y, Rr, R, kyz, cf)), (lambda x, y: toroidal_normal(x, y, Rr, R, kyz, cf))
    if s.kind == T.GEOM_FORBES_QBFS:
        a, nr = list(s.coefficients), s.norm_radius
        return (lambda x, y: forbes_qbfs_sag(x, y, R, k, a, nr)), (lambda x, y: forbes_qbfs_normal(x, y, R, k, a, nr))
    if s.kind == T.GEOM_ZERNIKE:
        terms = s.coefficients.reshape(-1, 4)
        nr = s.norm_radius
        return (
            lambda x, y: zernike_sag(x, y, R, k, terms, nr, status),
            lambda x, y: zernike_normal(x, y, R, k, terms, nr),
        )
    raise ValueError(f"not a Newton geometry: {s.kind}")


# ---- Forbes Q (slope-orthogonal, "Q^bfs") radial surfaces -------------------------------------------------
def _qbfs_fgh(nmax):
    """Recurrence coefficients of the Q^bfs basis (optiland/geometries/forbes/qpoly.py:56-84; G. W. Forbes,
    Opt. Express 18, 19700 (2010))."""
    f, g, h = {0: 2.0, 1: np.sqrt(19.0) / 2}, {0: -0.5}, {}
    for n in range(2, nmax + 1):
        h[n - 2] = -n * (n - 1) / (2 * f[n - 2])
        g[n - 1] = -(1 + g[n - 2] * h[n - 2]) / f[n - 1]
        f[n] = np.sqrt(n * (n + 1) + 3 - g[n - 1] ** 2 - h[n - 2] ** 2)
    return f, g, h


def _qbfs_change_basis(cs):
    """qpoly.py:87-115."""
    m = len(cs) - 1
    if m < 0:
        return []
    f, g, h = _qbfs_fgh(m)
    bs = [0.0] * (m + 1)
    bs[m] = cs[m] / f[m]
    if m >= 1:
        bs[m - 1] = (cs[m - 1] - g[m - 1] * bs[m]) / f[m - 1]
    for i in range(m - 2, -1, -1):
        bs[i] = (cs[i] - g[i] * bs[i + 1] - h[i] * bs[i + 2]) / f[i]
    return bs


def _qbfs_sum(cs, usq, want_derivative=False):
    """clenshaw_qbfs / clenshaw_qbfs_der (qpoly.py:131-143, 146-162, 185-212): S and dS/d(usq)."""
    bs = _qbfs_change_basis(cs)
    m = len(bs) - 1
    if m < 0:
        z = np.zeros_like(usq)
        return (z, z) if want_derivative else z
    prefix = 2 - 4 * usq
    al = [None] * (m + 1)
    al[m] = bs[m] + np.zeros_like(usq)
    if m > 0:
        al[m - 1] = bs[m - 1] + prefix * al[m]
    for i in range(m - 2, -1, -1):
        al[i] = bs[i] + prefix * al[i + 1] - al[i + 2]
    S = 2 * (al[0] + al[1]) if m > 0 else 2 * al[0]
    if not want_derivative:
        return S
    d = [np.zeros_like(usq) for _ in range(m + 1)]
    if m - 1 >= 0:
        d[m - 1] = -4 * al[m]
    if m - 2 >= 0:
        d[m - 2] = prefix * d[m - 1] - 4 * al[m - 1]
    for n in range(m - 3, -1, -1):
        d[n] = prefix * d[n + 1] - d[n + 2] - 4 * al[n + 1]
    dS = 2 * (d[0] + d[1]) if m > 0 else 2 * d[0]
    return S, dS


def _forbes_phi(r2, radius, k):
    """_conic_correction_factor, forbes/geometry.py:152-181."""
    if np.isinf(radius):
        return 1.0, 0.0
    c2 = (1.0 / radius) ** 2
    rho = np.sqrt(r2)
    num, den = 1 - k * c2 * r2, 1 - (k + 1) * c2 * r2
    Nn = np.sqrt(np.where(num > 0, num, 1e-12))
    D = np.sqrt(np.where(den > 0, den, 1e-12))
    return Nn / D, (c2 * rho) / (Nn * D**3)


def forbes_qbfs_sag(x, y, radius, k, a, norm_radius):
    """ForbesQNormalSlopeGeometry.sag, forbes/geometry.py:276-298."""
    with np.errstate(all="ignore"):
        r2 = x**2 + y**2
        if np.isinf(radius):
            zb = np.zeros_like(r2)
        else:
            arg = 1 - (1 + k) * r2 / radius**2
            zb = r2 / (radius * (1 + np.sqrt(np.where(arg < 0, 0, arg))))
        usq = r2 / norm_radius**2
        phi, _ = _forbes_phi(r2, radius, k)
        dep = usq * (1 - usq) * phi * _qbfs_sum(a, usq)
        return zb + np.where(usq > 1, 0.0, dep)


def forbes_qbfs_normal(x, y, radius, k, a, norm_radius):
    """_surface_normal / _surface_normal_analytical (NumPy branch), forbes/geometry.py:300-366."""
    eps = 1e-12
    with np.errstate(all="ignore"):
        r2 = x**2 + y**2
        rho = np.sqrt(r2 + eps**2)
        if np.isinf(radius) or radius == 0:
            dbase = np.zeros_like(rho)
        else:
            c = 1.0 / radius
            arg = 1 - (k + 1) * c**2 * r2
            dbase = c * rho / np.sqrt(np.where(arg > 0, arg, 1e-12))
        if len(a) == 0 or all(v == 0 for v in a):
            df = dbase
        else:
            u = rho / norm_radius
            Sx, dS = _qbfs_sum(a, u**2, want_derivative=True)
            dpoly_du = dS * 2 * u
            dpref = (2 * u - 4 * u**3) / norm_radius
            dpoly_drho = dpoly_du / norm_radius
            phi, dphi = _forbes_phi(r2, radius, k)
            usq = u**2
            dep = dpref * phi * Sx + (usq - usq**2) * dphi * Sx + (usq - usq**2) * phi * dpoly_drho
            df = dbase + np.where(u >= 1, 0.0, dep)
        dfdx, dfdy = df * (x / rho), df * (y / rho)
        mag = np.sqrt(dfdx**2 + dfdy**2 + 1)
        mag = np.where(mag < eps, 1.0, mag)
        return dfdx / mag, dfdy / mag, -1 / mag


def newton_distance(x, y, z, L, M, N, s: T.SurfaceSpec, sag, normal):
    """optiland/geometries/newton_raphson.py:119-168 -- including the GLOBAL
    convergence test ``max over all rays |f| < tol`` (:147-149)."""
    t = conic_distance(x, y, z, L, M, N, s.radius, s.conic)
    with np.errstate(all="ignore"):
        for _ in range(s.max_iter):
            x_int = x + t * L
            y_int = y + t * M
            z_int = z + t * N
            f_t = sag(x_int, y_int) - z_int
            if np.max(np.abs(f_t)) < s.tol:
                break
            nx, ny, nz = normal(x_int, y_int)
            nz_safe = np.where(np.abs(nz) > 1e-14, nz, 1e-14)
            fx = -nx / nz_safe
            fy = -ny / nz_safe
            df_dt = fx * L + fy * M - N
            safe_df_dt = np.where(np.abs(df_dt) > 1e-14, df_dt, 1e-14)
            t = t - f_t / safe_df_dt
    return t


# --------------------------------------------------------------------------
# apertures
# --------------------------------------------------------------------------

def aperture_inside(prog, x, y):
    """Postfix evaluation of optiland/physical_apertures: radial.py:56-70,
    offset_radial.py contains, rectangular.py contains, elliptical.py contains,
    base.py:259-340 (Union / Intersection / Difference)."""
    stack = []
    i = 0
    with np.errstate(all="ignore"):
        while i < len(prog):
            op = int(prog[i])
            if op == T.AP_RADIAL:
                r_max, r_min = prog[i + 1: i + 3]
                radius2 = x**2 + y**2
                stack.append((radius2 <= r_max**2) & (radius2 >= r_min**2))
                i += 3
            elif op == T.AP_OFFSET_RADIAL:
                r_max, r_min, dx, dy = prog[i + 1: i + 5]
                radius2 = (x - dx) ** 2 + (y - dy) ** 2
                stack.append(np.logical_and(radius2 <= r_max**2, radius2 >= r_min**2))
                i += 5
            elif op == T.AP_RECT:
                x_min, x_max, y_min, y_max = prog[i + 1: i + 5]
                stack.append((x_min <= x) & (x <= x_max) & (y_min <= y) & (y <= y_max))
                i += 5
            elif op == T.AP_ELLIPSE:
                a, b, dx, dy = prog[i + 1: i + 5]
                xs, ys = x - dx, y - dy
                stack.append((xs**2 / a**2 + ys**2 / b**2) <= 1)
                i += 5
            elif op in (T.AP_UNION, T.AP_INTERSECT, T.AP_DIFFERENCE):
                b_ = stack.pop()
                a_ = stack.pop()
                if op == T.AP_UNION:
                    stack.append(np.logical_or(a_, b_))
                elif op == T.AP_INTERSECT:
                    stack.append(np.logical_and(a_, b_))
                else:
                    stack.append(np.logical_and(a_, np.logical_not(b_)))
                i += 1
            else:
                raise ValueError(f"bad aperture opcode {op}")
    assert len(stack) == 1
    return stack[0]


# --------------------------------------------------------------------------
# polarization
# --------------------------------------------------------------------------

def local_basis(k0, k1):
    """optiland/rays/polarized_rays.py:136-176 (get_local_basis)."""
    with np.errstate(all="ignore"):
        s = np.cross(k0, k1)
        mag = np.linalg.norm(s, axis=1)
        mask = mag == 0
        if np.any(mask):
            xh = np.broadcast_to(np.array([1.0, 0.0, 0.0]), k0[mask].shape)
            p_fallback = np.cross(k0[mask], xh)
            p_norms = np.linalg.norm(p_fallback, axis=1)
            yh = np.broadcast_to(np.array([0.0, 1.0, 0.0]), k0[mask].shape)
            p_fallback = np.where((p_norms == 0)[..., None], np.cross(k0[mask], yh), p_fallback)
            s[mask] = np.cross(p_fallback, k0[mask])
            mag = np.linalg.norm(s, axis=1)
        s = s / mag[..., None]
        p0 = np.cross(k0, s)
        p1 = np.cross(k1, s)
        o_in = np.stack((s, p0, k0), axis=1)
        o_out = np.stack((s, p1, k1), axis=2)
    return o_in, o_out


def fresnel_jones(aoi, n1, n2, reflect, n_rays):
    """optiland/jones.py:71-117 (JonesFresnel.calculate_matrix)."""
    with np.errstate(all="ignore"):
        cos_theta_i = np.cos(aoi)
        n = n2 / n1
        radicand = (n**2 - np.sin(aoi) ** 2).astype(np.complex128)
        root = np.sqrt(radicand)
        J = np.zeros((n_rays, 3, 3), dtype=np.complex128)
        if reflect:
            s = (cos_theta_i - root) / (cos_theta_i + root)
            p = (n**2 * cos_theta_i - root) / (n**2 * cos_theta_i + root)
            J[:, 0, 0] = s
            J[:, 1, 1] = -p
            J[:, 2, 2] = -1
        else:
            s = 2 * cos_theta_i / (cos_theta_i + root)
            p = 2 * n * cos_theta_i / (n**2 * cos_theta_i + root)
            J[:, 0, 0] = s
            J[:, 1, 1] = p
            J[:, 2, 2] = 1
    return J


def polarized_update(P, L0, M0, N0, L, M, N, jones=None):
    """optiland/rays/polarized_rays.py:178-202 (PolarizedRays.update)."""
    k0 = np.stack([L0, M0, N0]).T
    k1 = np.stack([L, M, N]).T
    o_in, o_out = local_basis(k0, k1)
    with np.errstate(all="ignore"):
        if jones is None:
            p = np.matmul(o_out, o_in)
        else:
            dt = np.result_type(o_out, jones, o_in)
            p = np.matmul(np.matmul(o_out.astype(dt), jones.astype(dt)), o_in.astype(dt))
        return np.matmul(p, P)


def polarized_intensity(P, L_init, M_init, N_init, i0, state=None):
    """optiland/rays/polarized_rays.py:57-133 + 204-233: update_intensity for a
    polarization state; ``state`` None => unpolarized (mean of two orthogonal inputs).
    ``state`` is (Ex, Ey, phase_x, phase_y)."""
    k = np.stack([L_init, M_init, N_init]).T
    xh = np.broadcast_to(np.array([1.0, 0.0, 0.0]), k.shape)
    p = np.cross(k, xh)
    norms = np.linalg.norm(p, axis=1)
    if np.any(norms == 0):
        raise ValueError("k-vector parallel to x-axis is not currently supported.")
    p = p / norms[..., None]
    s = np.cross(p, k)

    def field(Ex, Ey, phx, phy):
        E0 = Ex * np.exp(1j * phx) * s + Ey * np.exp(1j * phy) * p
        return np.squeeze(np.matmul(P, E0[:, :, None]), axis=2)

    fields = [field(*state)] if state is not None else [field(1.0, 0.0, 0.0, 0.0), field(0.0, 1.0, 0.0, 0.0)]
    intensity = np.zeros_like(i0)
    for E1 in fields:
        intensity = intensity + np.sum(np.abs(E1) ** 2, axis=1)
    return intensity * i0 / len(fields)


# --------------------------------------------------------------------------
# the loop
# --------------------------------------------------------------------------

RECORD_KEYS = ("x", "y", "z", "L", "M", "N", "intensity", "opd")


def wavelength_index(w, wavelengths):
    """Map each ray's wavelength to its column of the media table (exact match)."""
    idx = np.full(w.shape, -1, dtype=np.int64)
    for j, wl in enumerate(wavelengths):
        idx[w == wl] = j
    if np.any(idx < 0):
        raise ValueError("ray wavelength not present in table.wavelengths")
    return idx


def trace(table: T.SurfaceTable, rays: dict, first: int = 0, last: int | None = None,
          polarized: bool = False):
    """SurfaceGroup.trace(rays, skip=first) restricted to surfaces [first, last).

    optiland/surfaces/surface_group.py:245-257; Surface.trace
    optiland/surfaces/standard_surface.py:200-215; Surface._trace_real :232-248.

    ``rays``: dict with fp64 arrays x,y,z,L,M,N,i,w and optional opd, p (N,3,3 complex).
    Returns (rays_out: dict incl. L0,M0,N0[,p]; records: dict of (rows, N) arrays; status int).
    """
    last = table.num_surfaces if last is None else last
    f8 = np.float64
    x, y, z = (np.array(rays[k], dtype=f8) for k in "xyz")
    L, M, N = (np.array(rays[k], dtype=f8) for k in "LMN")
    inten = np.array(rays["i"], dtype=f8)
    w = np.array(rays["w"], dtype=f8)
    opd = np.array(rays["opd"], dtype=f8) if "opd" in rays else np.zeros_like(x)
    P = None
    if polarized:
        P = np.array(rays["p"]) if "p" in rays else np.tile(np.eye(3), (x.size, 1, 1))
    L0 = M0 = N0 = None
    widx = wavelength_index(w, table.wavelengths)
    status = [0]
    rec = {k: [] for k in RECORD_KEYS}

    for si in range(first, last):
        s = table.surfaces[si]
        if s.kind != T.GEOM_NOOP:
            # -- localize: optiland/coordinate_system.py:73-89 (flattened pose) --
            x, y, z = x - s.t[0], y - s.t[1], z - s.t[2]
            if s.rotated:
                Rm = s.R
                x, y, z = (Rm[0, c] * x + Rm[1, c] * y + Rm[2, c] * z for c in range(3))
                L, M, N = (Rm[0, c] * L + Rm[1, c] * M + Rm[2, c] * N for c in range(3))
            # -- distance --
            if s.kind == T.GEOM_PLANE:
                t = plane_distance(z, N)
            elif s.kind == T.GEOM_STANDARD:
                t = conic_distance(x, y, z, L, M, N, s.radius, s.conic)
            else:
                sag, normal = _sag_and_normal_fns(s, status)
                t = newton_distance(x, y, z, L, M, N, s, sag, normal)
            n1 = s.n1[widx]
            n2 = s.n2[widx]
            with np.errstate(all="ignore"):
                # -- propagate: optiland/propagation/homogeneous.py:30-57 --
                x = x + t * L
                y = y + t * M
                z = z + t * N
                k1 = s.k1[widx]
                if np.any(k1 > 0):
                    alpha = 4 * np.pi * k1 / w
                    inten = inten * np.exp(-alpha * t * 1e3)
                # -- OPD: optiland/surfaces/standard_surface.py:244 --
                opd = opd + np.abs(t * n1)
                # -- aperture: standard_surface.py:245-246, real_rays.py:154-161 --
                if s.aperture is not None:
                    inside = aperture_inside(s.aperture, x, y)
                    inten = np.where(~inside, np.zeros_like(inten), inten)
                # -- interaction: interactions/refractive_reflective_model.py:32-55 --
                if s.kind == T.GEOM_PLANE:
                    nx, ny, nz = np.zeros_like(x), np.zeros_like(x), np.ones_like(x)
                elif s.kind == T.GEOM_STANDARD:
                    if math.isinf(s.radius):
                        # 1 - (1+k) r2/inf**2 = 1 ; denom = inf ; dfdx = 0
                        nx, ny, nz = conic_normal(x, y, s.radius, s.conic)
                    else:
                        nx, ny, nz = conic_normal(x, y, s.radius, s.conic)
                else:
                    nx, ny, nz = normal(x, y)
                L0, M0, N0 = L.copy(), M.copy(), N.copy()
                # _align_surface_normal: optiland/rays/real_rays.py:535-571
                dot = L0 * nx + M0 * ny + N0 * nz
                sgn = np.sign(dot)
                nx, ny, nz = nx * sgn, ny * sgn, nz * sgn
                dot = np.abs(dot)
                if s.reflective:
                    # optiland/rays/real_rays.py:189-205
                    L = L - 2 * dot * nx
                    M = M - 2 * dot * ny
                    N = N - 2 * dot * nz
                else:
                    # optiland/rays/real_rays.py:163-187
                    u = n1 / n2
                    root = np.sqrt(1 - u**2 * (1 - dot**2))
                    L = u * L0 + nx * root - u * nx * dot
                    M = u * M0 + ny * root - u * ny * dot
                    N = u * N0 + nz * root - u * nz * dot
                # -- coating: optiland/interactions/base.py:111-128 --
                if s.coating == T.COAT_SIMPLE:
                    inten = inten * (s.coat_r if s.reflective else s.coat_t)
                elif s.coating == T.COAT_FRESNEL:
                    if P is None:
                        raise ValueError("Fresnel coating requires polarized rays")
                    # optiland/coatings.py:72-93 (_compute_aoi), :285-331
                    d = np.abs(nx * L0 + ny * M0 + nz * N0)
                    aoi = np.arccos(np.clip(d, -1, 1))
                    J = fresnel_jones(aoi, s.coat_n1[widx], s.coat_n2[widx], s.reflective, x.size)
                    P = polarized_update(P, L0, M0, N0, L, M, N, J)
                elif P is not None:
                    P = polarized_update(P, L0, M0, N0, L, M, N, None)
            # -- globalize: optiland/coordinate_system.py:91-107 --
            if s.rotated:
                Rm = s.R
                x, y, z = (Rm[r, 0] * x + Rm[r, 1] * y + Rm[r, 2] * z for r in range(3))
                L, M, N = (Rm[r, 0] * L + Rm[r, 1] * M + Rm[r, 2] * N for r in range(3))
            x, y, z = x + s.t[0], y + s.t[1], z + s.t[2]
        # -- record: optiland/surfaces/standard_surface.py:260-274 --
        if s.record:
            for key, val in zip(RECORD_KEYS, (x, y, z, L, M, N, inten, opd)):
                rec[key].append(val.copy())
        else:
            for key in RECORD_KEYS:
                rec[key].append(np.full_like(x, np.nan))

    out = {"x": x, "y": y, "z": z, "L": L, "M": M, "N": N, "i": inten, "w": w, "opd": opd,
           "L0": L0, "M0": M0, "N0": N0}
    if P is not None:
        out["p"] = P
    records = {k: (np.stack(v) if v else np.zeros((0, x.size))) for k, v in rec.items()}
    return out, records, status[0]


# --------------------------------------------------------------------------
# Huygens-Fresnel PSF summation (consumer of the path)
# --------------------------------------------------------------------------

def huygens_fresnel_psf(image_x, image_y, image_z, pupil_x, pupil_y, pupil_z, pupil_amp, pupil_opd, wavelength, Rp):
    """optiland/psf/huygens_fresnel_strategies.py:97-160 (NumbaSummation._huygens_fresnel_summation),
    vectorised over the pupil for each image point; returns (psf, field)."""
    k = 2.0 * np.pi / wavelength
    shape = np.shape(image_x)
    ix, iy, iz = (np.asarray(a, dtype=np.float64).ravel() for a in (image_x, image_y, image_z))
    u, v, w = (np.asarray(a, dtype=np.float64).ravel() for a in (pupil_x, pupil_y, pupil_z))
    amp = np.asarray(pupil_amp).ravel()
    phase = np.exp(-1j * k * np.asarray(pupil_opd, dtype=np.float64).ravel())
    field = np.zeros(ix.size, dtype=np.complex128)
    for p in range(ix.size):
        dx, dy, dz = ix[p] - u, iy[p] - v, iz[p] - w
        R = np.sqrt(dx * dx + dy * dy + dz * dz)
        wave = np.exp(1j * k * R) / R
        dot = dx * (u / Rp) + dy * (v / Rp) + dz * (w / Rp)
        field[p] = np.sum(amp * phase * wave * (0.5 * (1.0 + dot / R)))
    field = field.reshape(shape)
    return np.abs(field) ** 2, field


def wavefront_reference_sphere(fin: dict, Px, Py, ref: dict) -> dict:
    """Steps 4-5 of ChiefRayStrategy.compute_wavefront_data (optiland/wavefront/strategy.py:179-190) for a
    spherical reference: SphericalReference.path_length (optiland/wavefront/reference_geometry.py:55-82), the
    launch-plane tilt term of _correct_tilt (strategy.py:93-139, pre-multiplied: tilt = (ux EPD/2, uy EPD/2)),
    the OPD in waves and the exit-pupil intercepts.  ``fin``: final GLOBAL ray state (x y z L M N opd i)."""
    f8 = np.float64
    xr, yr, zr = (np.asarray(fin[k], dtype=f8) for k in "xyz")
    Lf, Mf, Nf = (np.asarray(fin[k], dtype=f8) for k in "LMN")
    xc, yc, zc = (float(v) for v in ref["center"])
    R, n_img = float(ref["radius"]), float(ref["n_image"])
    L, M, N = -Lf, -Mf, -Nf
    with np.errstate(all="ignore"):
        a = L**2 + M**2 + N**2
        b = 2 * (L * (xr - xc) + M * (yr - yc) + N * (zr - zc))
        c = xr**2 + yr**2 + zr**2 - 2 * (xr * xc + yr * yc + zr * zc) + xc**2 + yc**2 + zc**2 - R**2
        d = b**2 - 4 * a * c
        d = np.where(d < 0, 0, d)
        t1 = (-b - np.sqrt(d)) / (2 * a)
        t2 = (-b + np.sqrt(d)) / (2 * a)
        t = np.where(t1 < 0, t2, t1)
        opd_img = n_img * t
        tilt = ref.get("tilt", (0.0, 0.0))
        opd = np.asarray(fin["opd"], dtype=f8) - opd_img + (float(tilt[0]) * np.asarray(Px, dtype=f8)
                                                            + float(tilt[1]) * np.asarray(Py, dtype=f8))
        opd_wv = (float(ref["opd_ref"]) - opd) / (float(ref["wavelength_um"]) * 1e-3)
        tt = opd_img / n_img
        return {"opd": opd_wv, "pupil_x": xr - tt * Lf, "pupil_y": yr - tt * Mf, "pupil_z": zr - tt * Nf,
                "intensity": np.asarray(fin["i"], dtype=f8)}
