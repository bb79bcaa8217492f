// olb_prep.h -- device-side ("prepared") surface table and its host-side construction.
//
// The host table of include/olb.h (fp64, reference vocabulary) is turned once per
// upload into one contiguous blob per element type T (float / double):
//
//     [ PrepHeader ][ PrepSurface<T> x n_surf ][ T pool[...] ]
//
// which the trace kernel pulls into shared memory with a single TMA bulk copy.
// Everything that is uniform over rays is precomputed here in fp64: flattened poses
// and surface-to-surface relative transforms, curvature, index ratios n1/n2 per
// wavelength, Beer-Lambert coefficients, and the monomial form of Zernike sums.
//
// Reference data contract: SURVEY.md Appendix B; citations inline.
#ifndef OLB_PREP_H_
#define OLB_PREP_H_

#include <stdint.h>

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/olb.h"

namespace olb {

// Feature bits: which code paths a table needs (selects the kernel instantiation).
enum : uint32_t {
  FEAT_ROT = 1u << 0,      // some surface has a rotated pose
  FEAT_NEWTON = 1u << 1,   // a Newton-iteration surface (alone: even / odd aspheres only)
  FEAT_EXTRA = 1u << 2,    // non-radial aperture programs, simple coatings, L0/M0/N0 output
  FEAT_POL = 1u << 3,      // polarized rays (P matrix) / Fresnel coatings
  FEAT_FREEFORM = 1u << 4, // polynomial / Zernike / Chebyshev / biconic / toroidal / Forbes surfaces (with NEWTON)
};

struct PrepHeader {
  int32_t n_surf;
  int32_t n_wl;
  int32_t pool_len;     // in elements of T
  uint32_t features;    // FEAT_* needed by this table
  int32_t blob_bytes;   // total bytes of this blob (multiple of 16)
  int32_t elem_size;    // sizeof(T)
  int32_t pad[2];
};  // 32 bytes

// Media block per surface in the pool: n_wl records of MEDIA_STRIDE values each.
enum { MED_N1 = 0, MED_U = 1, MED_ALPHA = 2, MED_CN = 3, MED_STRIDE = 4 };
//   MED_N1    n1                           (OPD: standard_surface.py:244)
//   MED_U     n1 / n2                      (refract: real_rays.py:174)
//   MED_ALPHA 4*pi*k1/lambda * 1e3         (homogeneous.py:45-53)
//   MED_CN    coating n2 / coating n1      (jones.py:95)

template <typename T>
struct PrepSurface {
  int32_t kind;
  uint32_t flags;      // OLB_SF_* plus the PSF_* bits below
  int32_t n_coef;      // even/odd: number of coefficients
  int32_t coef_off;    // pool offset (elements of T)
  int32_t aper_off;
  int32_t aper_len;
  int32_t max_iter;
  int32_t coating;
  int32_t media_off;
  int32_t poly_rows;   // bivariate tables: rows (x powers) and cols (y powers)
  int32_t poly_cols;
  int32_t poly_d_off;  // pool offset of the derivative-source table (Zernike quirk)
  int32_t gslot;       // backward: first per-thread gradient accumulator slot of this surface
  int32_t gslots;      //           number of slots (7 + n_coef [+ 9 for a tilted pose]; 0 for NOOP)
  int32_t pad16[2];    // keeps sizeof a multiple of 16 (256 / 448 bytes): the pool behind the array stays 16-byte aligned
  // incoming transform from GLOBAL coordinates: p_loc = Ag * p + bg
  T Ag[9], bg[3];
  // incoming transform from the PREVIOUS surface's local frame: p_loc = Ar * p + br
  T Ar[9], br[3];
  // outgoing transform to global: p_glob = R * p_loc + t
  T R[9], t[3];
  T radius, curv, conic, kp1;   // curv = 1/radius (0 for infinite radius), kp1 = 1 + k
  T tol, coat_t, coat_r, inv_norm;  // inv_norm = 1 / norm_radius (Chebyshev: 1 / norm_x)
  T inv_norm_y, curv_y, kp1_y, r_rot;  // Chebyshev 1/norm_y; biconic cy, 1+ky; toroidal: c_yz, 1+k_yz, R_rot
};
// the pool starts right behind the PrepSurface array: 16-byte alignment of its (4-element aligned) blocks needs this
static_assert(sizeof(PrepSurface<float>) % 16 == 0 && sizeof(PrepSurface<double>) % 16 == 0, "PrepSurface must keep the pool 16-byte aligned");
static_assert(sizeof(PrepHeader) % 16 == 0, "PrepHeader must keep the table 16-byte aligned");


enum : uint32_t {
  PSF_ROT_IN_G = 1u << 8,    // Ag != I
  PSF_ROT_IN_R = 1u << 9,    // Ar != I
  PSF_RADIUS_INF = 1u << 10, // StandardGeometry with infinite radius (standard.py:108-111)
  PSF_APER_RADIAL = 1u << 11,// aperture program is a single RADIAL op (fast path)
  PSF_POLY_TRI = 1u << 12,   // bivariate table is triangular (i + j <= rows-1)
};

static inline void mat3_mul(const double* A, const double* B, double* C) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += A[3 * r + k] * B[3 * k + c];
      C[3 * r + c] = s;
    }
}
static inline void mat3_transpose(const double* A, double* At) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) At[3 * c + r] = A[3 * r + c];
}
static inline bool mat3_is_identity(const double* A) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      if (A[3 * r + c] != (r == c ? 1.0 : 0.0)) return false;
  return true;
}

// ---- Zernike -> monomials in (xn, yn) -----------------------------------------
// Z_n^m = R_n^|m|(rho) * {cos(m phi) | sin(|m| phi)},  optiland/zernike/base.py:42-68,
// R_n^m(rho) = sum_k (-1)^k (n-k)! / (k! ((n+m)/2-k)! ((n-m)/2-k)!) rho^(n-2k)  (:217-243)
// rho^j cos(m phi) etc. are polynomials in xn = rho cos(phi), yn = rho sin(phi):
//   rho^(n-2k) {cos|sin}(m phi) = (xn^2+yn^2)^((n-m)/2-k) * {Re|Im} (xn + i yn)^m
static inline double fact(int n) {
  double f = 1;
  for (int i = 2; i <= n; ++i) f *= i;
  return f;
}
static inline double binom(int n, int k) { return fact(n) / (fact(k) * fact(n - k)); }

// Adds coef * Z_n^m to the (deg+1)x(deg+1) row-major table tab[i*(deg+1)+j] ~ xn^i yn^j.
static inline void zernike_add_monomials(int n, int m, double coef, int deg, double* tab) {
  const int ma = m < 0 ? -m : m;
  const int W = deg + 1;
  for (int k = 0; k <= (n - ma) / 2; ++k) {
    double rc = ((k & 1) ? -1.0 : 1.0) * fact(n - k) /
                (fact(k) * fact((n + ma) / 2 - k) * fact((n - ma) / 2 - k));
    const int q = (n - ma) / 2 - k;  // power of (xn^2 + yn^2)
    for (int a = 0; a <= q; ++a) {   // (x^2+y^2)^q = sum_a C(q,a) x^(2a) y^(2(q-a))
      double ca = binom(q, a);
      for (int j = 0; j <= ma; ++j) {  // (x+iy)^m = sum_j C(m,j) x^(m-j) (iy)^j
        // i^j: j%4==0 -> 1, 1 -> i, 2 -> -1, 3 -> -i
        const bool imag = (j & 1) != 0;
        if ((m >= 0) == imag) continue;  // cos wants real part, sin wants imaginary part
        double sgn = ((j >> 1) & 1) ? -1.0 : 1.0;
        int px = 2 * a + (ma - j);
        int py = 2 * (q - a) + j;
        tab[px * W + py] += coef * rc * ca * binom(ma, j) * sgn;
      }
    }
  }
}

enum : uint32_t { HINT_POLY_NEWTON = 1u };   // a polynomial-family / biconic / toroidal Newton surface is present
struct PrepResult {
  std::vector<unsigned char> blob_f64, blob_f32;
  uint32_t features = 0;
  uint32_t hints = 0;          // HINT_* (launch policy only, never semantics)
  bool bwd_supported = true;   // every surface is covered by surface_backward (olb_math.cuh)
  bool bwd_tables = false;     // ... and some surface is a polynomial / Zernike one: its TABLE gradients are wanted
                               // (olb_trace_bwd_tables_*)
  int total_gslots = 0;        // per-thread gradient accumulator slots the backward kernel needs
  std::string error;
};

template <typename T>
static void build_blob(const OlbTable& tab, const std::vector<std::vector<double>>& pools,
                       const std::vector<PrepSurface<double>>& ps, uint32_t features,
                       std::vector<unsigned char>& out) {
  // flatten per-surface pools into one pool, fixing offsets
  std::vector<PrepSurface<T>> surf(ps.size());
  std::vector<T> pool;
  for (size_t s = 0; s < ps.size(); ++s) {
    const PrepSurface<double>& a = ps[s];
    PrepSurface<T>& b = surf[s];
    const int base = (int)pool.size();
    b.kind = a.kind; b.flags = a.flags; b.n_coef = a.n_coef;
    b.coef_off = a.coef_off + base; b.aper_off = a.aper_off + base; b.aper_len = a.aper_len;
    b.max_iter = a.max_iter; b.coating = a.coating; b.media_off = a.media_off + base;
    b.poly_rows = a.poly_rows; b.poly_cols = a.poly_cols; b.poly_d_off = a.poly_d_off + base;
    b.gslot = a.gslot; b.gslots = a.gslots;
    for (int i = 0; i < 9; ++i) { b.Ag[i] = (T)a.Ag[i]; b.Ar[i] = (T)a.Ar[i]; b.R[i] = (T)a.R[i]; }
    for (int i = 0; i < 3; ++i) { b.bg[i] = (T)a.bg[i]; b.br[i] = (T)a.br[i]; b.t[i] = (T)a.t[i]; }
    b.radius = (T)a.radius; b.curv = (T)a.curv; b.conic = (T)a.conic; b.kp1 = (T)a.kp1;
    b.tol = (T)a.tol; b.coat_t = (T)a.coat_t; b.coat_r = (T)a.coat_r; b.inv_norm = (T)a.inv_norm;
    b.inv_norm_y = (T)a.inv_norm_y; b.curv_y = (T)a.curv_y; b.kp1_y = (T)a.kp1_y; b.r_rot = (T)a.r_rot;
    for (double v : pools[s]) pool.push_back((T)v);
    while (pool.size() % 4) pool.push_back((T)0);
  }
  // wavelengths at the end of the pool
  const int wl_off = (int)pool.size();
  for (int j = 0; j < tab.n_wl; ++j) pool.push_back((T)tab.wavelengths[j]);
  while (pool.size() % 4) pool.push_back((T)0);

  PrepHeader h{};
  h.n_surf = (int32_t)ps.size();
  h.n_wl = tab.n_wl;
  h.pool_len = (int32_t)pool.size();
  h.features = features;
  h.elem_size = (int32_t)sizeof(T);
  h.pad[0] = wl_off;
  size_t bytes = sizeof(PrepHeader) + surf.size() * sizeof(PrepSurface<T>) + pool.size() * sizeof(T);
  bytes = (bytes + 15) & ~size_t(15);
  h.blob_bytes = (int32_t)bytes;
  out.assign(bytes, 0);
  unsigned char* p = out.data();
  std::memcpy(p, &h, sizeof(h)); p += sizeof(h);
  std::memcpy(p, surf.data(), surf.size() * sizeof(PrepSurface<T>)); p += surf.size() * sizeof(PrepSurface<T>);
  std::memcpy(p, pool.data(), pool.size() * sizeof(T));
}

static inline int aperture_operands(int op) {
  switch (op) {
    case OLB_AP_RADIAL: return 2;
    case OLB_AP_OFFSET_RADIAL: case OLB_AP_RECT: case OLB_AP_ELLIPSE: return 4;
    case OLB_AP_UNION: case OLB_AP_INTERSECT: case OLB_AP_DIFFERENCE: return 0;
    default: return -1;
  }
}

// Validate + prepare. Returns empty error string on success.
static inline PrepResult prepare_table(const OlbTable& tab) {
  PrepResult res;
  if (!tab.surfaces || !tab.wavelengths || !tab.pool) { res.error = "NULL table member"; return res; }
  if (tab.n_surfaces < 1 || tab.n_surfaces > OLB_MAX_SURFACES) { res.error = "n_surfaces out of range"; return res; }
  if (tab.n_wl < 1 || tab.n_wl > OLB_MAX_WAVELENGTHS) { res.error = "n_wl out of range"; return res; }
  const int n_wl = tab.n_wl;
  std::vector<PrepSurface<double>> ps(tab.n_surfaces);
  std::vector<std::vector<double>> pools(tab.n_surfaces);
  uint32_t features = 0;
  int prev = -1;  // previous surface with a frame (non-NOOP)
  auto in_pool = [&](int off, int len) { return off >= 0 && len >= 0 && (int64_t)off + len <= tab.pool_len; };

  for (int s = 0; s < tab.n_surfaces; ++s) {
    const OlbSurface& in = tab.surfaces[s];
    PrepSurface<double>& o = ps[s];
    std::memset(&o, 0, sizeof(o));
    std::vector<double>& pool = pools[s];
    o.kind = in.kind;
    o.flags = in.flags & 0xffu;
    o.max_iter = in.max_iter;
    o.coating = in.coating;
    if (in.kind < OLB_GEOM_NOOP || in.kind > OLB_GEOM_FORBES_QBFS) { res.error = "unknown geometry kind"; return res; }

    // ---- pose --------------------------------------------------------------
    if (in.kind != OLB_GEOM_NOOP) {
      bool finite = std::isfinite(in.t[0]) && std::isfinite(in.t[1]) && std::isfinite(in.t[2]);
      for (int i = 0; i < 9; ++i) finite = finite && std::isfinite(in.R[i]);
      if (!finite) { res.error = "non-finite surface pose"; return res; }
    }
    double Rt[9];
    mat3_transpose(in.R, Rt);
    for (int i = 0; i < 9; ++i) { o.R[i] = in.R[i]; o.Ag[i] = Rt[i]; }
    for (int i = 0; i < 3; ++i) o.t[i] = in.t[i];
    const bool rotated = !mat3_is_identity(in.R);
    for (int r = 0; r < 3; ++r) {
      if (rotated) o.bg[r] = -(Rt[3 * r] * in.t[0] + Rt[3 * r + 1] * in.t[1] + Rt[3 * r + 2] * in.t[2]);
      else o.bg[r] = -in.t[r];
    }
    if (rotated) { o.flags |= OLB_SF_ROTATED | PSF_ROT_IN_G; features |= FEAT_ROT; }
    else o.flags &= ~uint32_t(OLB_SF_ROTATED);
    if (in.kind != OLB_GEOM_NOOP) {
      if (prev >= 0) {
        const OlbSurface& pv = tab.surfaces[prev];
        const bool prev_rot = !mat3_is_identity(pv.R);
        if (!rotated && !prev_rot) {
          for (int i = 0; i < 9; ++i) o.Ar[i] = (i % 4 == 0) ? 1.0 : 0.0;
          for (int r = 0; r < 3; ++r) o.br[r] = pv.t[r] - in.t[r];
        } else {
          mat3_mul(Rt, pv.R, o.Ar);  // R_cur^T R_prev
          double d[3] = {pv.t[0] - in.t[0], pv.t[1] - in.t[1], pv.t[2] - in.t[2]};
          for (int r = 0; r < 3; ++r) o.br[r] = Rt[3 * r] * d[0] + Rt[3 * r + 1] * d[1] + Rt[3 * r + 2] * d[2];
          if (!mat3_is_identity(o.Ar)) { o.flags |= PSF_ROT_IN_R; features |= FEAT_ROT; }
        }
      } else {
        for (int i = 0; i < 9; ++i) o.Ar[i] = o.Ag[i];
        for (int r = 0; r < 3; ++r) o.br[r] = o.bg[r];
        if (rotated) o.flags |= PSF_ROT_IN_R;
      }
      prev = s;
    }

    // ---- geometry ----------------------------------------------------------
    o.radius = in.radius; o.conic = in.conic; o.kp1 = 1.0 + in.conic;
    o.tol = in.tol;
    o.coat_t = in.coat_t; o.coat_r = in.coat_r;
    o.inv_norm = 1.0; o.inv_norm_y = 1.0; o.curv_y = 0; o.kp1_y = 1.0; o.r_rot = INFINITY;
    if (in.kind == OLB_GEOM_PLANE || in.kind == OLB_GEOM_NOOP) { o.radius = INFINITY; o.curv = 0; }
    else if (std::isinf(in.radius)) { o.curv = 0; o.flags |= PSF_RADIUS_INF; }
    else { o.curv = 1.0 / in.radius; }
    const bool newton = in.kind >= OLB_GEOM_EVEN_ASPHERE;
    if (newton) {
      features |= FEAT_NEWTON;
      if (in.kind != OLB_GEOM_EVEN_ASPHERE && in.kind != OLB_GEOM_ODD_ASPHERE) { res.hints |= HINT_POLY_NEWTON; features |= FEAT_FREEFORM; }
      if (in.max_iter < 0) { res.error = "negative max_iter"; return res; }
    }
    if (in.kind == OLB_GEOM_EVEN_ASPHERE || in.kind == OLB_GEOM_ODD_ASPHERE) {
      if (!in_pool(in.coef_off, in.n_coef)) { res.error = "coefficient block outside pool"; return res; }
      o.n_coef = in.n_coef;
      o.coef_off = (int)pool.size();
      for (int i = 0; i < in.n_coef; ++i) pool.push_back(tab.pool[in.coef_off + i]);
      while (pool.size() % 4) pool.push_back(0);
      // slope coefficients, so that the Newton loop does one FMA per term: 2(i+1) C_i (even), (i+1) C_i (odd)
      o.poly_d_off = (int)pool.size();
      const double step = in.kind == OLB_GEOM_EVEN_ASPHERE ? 2.0 : 1.0;
      for (int i = 0; i < in.n_coef; ++i) pool.push_back(step * (i + 1) * tab.pool[in.coef_off + i]);
      while (pool.size() % 4) pool.push_back(0);
    } else if (in.kind == OLB_GEOM_POLYNOMIAL) {
      const int cols = in.aux0 > 0 ? in.aux0 : 1;
      if (in.n_coef % cols || !in_pool(in.coef_off, in.n_coef)) { res.error = "bad polynomial block"; return res; }
      o.poly_rows = in.n_coef / cols; o.poly_cols = cols;
      if (o.poly_rows > 24 || cols > 24) { res.error = "polynomial order too high (max 23)"; return res; }
      o.coef_off = (int)pool.size();
      for (int i = 0; i < in.n_coef; ++i) pool.push_back(tab.pool[in.coef_off + i]);
      while (pool.size() % 4) pool.push_back(0);
      o.poly_d_off = o.coef_off;  // derivative of the same polynomial (polynomial.py:123-155)
      o.inv_norm = 1.0;
    } else if (in.kind == OLB_GEOM_CHEBYSHEV) {
      // conic + sum C_ij T_i(x/nx) T_j(y/ny)  (chebyshev.py:126-150): expanded into monomials of
      // (xn, yn) with the integer coefficients of T_n (T_0 = 1, T_1 = x, T_{n+1} = 2x T_n - T_{n-1}).
      const int cols = in.aux0 > 0 ? in.aux0 : 1;
      if (in.n_coef % cols || !in_pool(in.coef_off, in.n_coef + 2)) { res.error = "bad Chebyshev block"; return res; }
      const int rows = in.n_coef / cols;
      if (rows > 24 || cols > 24) { res.error = "Chebyshev order too high (max 23)"; return res; }
      const double nx = tab.pool[in.coef_off], ny = tab.pool[in.coef_off + 1];
      if (!(nx > 0) || !(ny > 0)) { res.error = "Chebyshev norms must be positive"; return res; }
      const int M = rows > cols ? rows : cols;
      std::vector<std::vector<double>> Tc(M, std::vector<double>(M, 0.0));  // Tc[n][p]: coefficient of x^p in T_n
      Tc[0][0] = 1.0;
      if (M > 1) Tc[1][1] = 1.0;
      for (int n = 2; n < M; ++n)
        for (int p = 0; p < M; ++p) Tc[n][p] = (p > 0 ? 2.0 * Tc[n - 1][p - 1] : 0.0) - Tc[n - 2][p];
      std::vector<double> Pm(rows * cols, 0.0);
      for (int i = 0; i < rows; ++i)
        for (int j = 0; j < cols; ++j) {
          const double c = tab.pool[in.coef_off + 2 + i * cols + j];
          if (c == 0) continue;
          for (int p = 0; p <= i; ++p)
            for (int q = 0; q <= j; ++q) Pm[p * cols + q] += c * Tc[i][p] * Tc[j][q];
        }
      o.poly_rows = rows; o.poly_cols = cols;
      o.coef_off = (int)pool.size();
      pool.insert(pool.end(), Pm.begin(), Pm.end());
      while (pool.size() % 4) pool.push_back(0);
      o.poly_d_off = o.coef_off;
      o.inv_norm = 1.0 / nx; o.inv_norm_y = 1.0 / ny;
    } else if (in.kind == OLB_GEOM_BICONIC) {
      if (!in_pool(in.coef_off, 2)) { res.error = "biconic block outside pool"; return res; }
      const double Ry = tab.pool[in.coef_off], ky = tab.pool[in.coef_off + 1];
      // cx / cy = 0 for an infinite or zero radius (biconic.py:63-64)
      if (std::isinf(in.radius) || in.radius == 0) o.curv = 0;
      o.curv_y = (std::isinf(Ry) || Ry == 0) ? 0.0 : 1.0 / Ry;
      o.kp1_y = 1.0 + ky;
    } else if (in.kind == OLB_GEOM_TOROIDAL) {
      if (!in_pool(in.coef_off, 2 + in.n_coef) || in.n_coef < 0) { res.error = "toroidal block outside pool"; return res; }
      if (in.conic != 0) { res.error = "toroidal: OlbSurface.conic must be 0 (the Newton start sphere)"; return res; }
      o.r_rot = tab.pool[in.coef_off];
      o.kp1_y = 1.0 + tab.pool[in.coef_off + 1];
      o.curv_y = (std::isfinite(in.radius) && in.radius != 0) ? 1.0 / in.radius : 0.0;  // c_yz (toroidal.py:82-84)
      o.n_coef = in.n_coef;
      o.coef_off = (int)pool.size();
      for (int i = 0; i < in.n_coef; ++i) pool.push_back(tab.pool[in.coef_off + 2 + i]);
      while (pool.size() % 4) pool.push_back(0);
    } else if (in.kind == OLB_GEOM_FORBES_QBFS) {
      if (in.n_coef < 0 || in.n_coef > 64 || !in_pool(in.coef_off, in.n_coef)) { res.error = "bad Forbes coefficient block"; return res; }
      if (!(in.norm_radius > 0)) { res.error = "Forbes norm_radius must be positive"; return res; }
      // Change of basis a_m -> b_m of the orthonormal polynomials the Clenshaw recurrence runs on
      // (G. W. Forbes, Opt. Express 18, 19700 (2010), eqs. A.14-A.16; geometries/forbes/qpoly.py:56-115):
      //   f_0 = 2, f_1 = sqrt(19)/2, g_0 = -1/2, h_{n-2} = -n(n-1) / (2 f_{n-2}),
      //   g_{n-1} = -(1 + g_{n-2} h_{n-2}) / f_{n-1}, f_n = sqrt(n(n+1) + 3 - g_{n-1}^2 - h_{n-2}^2)
      const int nc = in.n_coef;
      std::vector<double> f(nc + 2), g(nc + 2), h(nc + 2), b(nc, 0.0);
      for (int n = 0; n < nc; ++n) {
        if (n == 0) f[0] = 2.0;
        else if (n == 1) { g[0] = -0.5; f[1] = std::sqrt(19.0) / 2.0; }
        else {
          h[n - 2] = -(double)n * (n - 1) / (2.0 * f[n - 2]);
          g[n - 1] = -(1.0 + g[n - 2] * h[n - 2]) / f[n - 1];
          f[n] = std::sqrt((double)n * (n + 1) + 3.0 - g[n - 1] * g[n - 1] - h[n - 2] * h[n - 2]);
        }
      }
      const double* a = tab.pool + in.coef_off;
      const int m = nc - 1;
      bool all_zero = true;
      for (int i = 0; i < nc; ++i) all_zero = all_zero && a[i] == 0.0;
      if (m >= 0) b[m] = a[m] / f[m];
      if (m >= 1) b[m - 1] = (a[m - 1] - g[m - 1] * b[m]) / f[m - 1];
      for (int i = m - 2; i >= 0; --i) b[i] = (a[i] - g[i] * b[i + 1] - h[i] * b[i + 2]) / f[i];
      o.n_coef = all_zero ? 0 : nc;    // no / all-zero terms: the reference's slope takes the base-conic branch
      o.coef_off = (int)pool.size();
      for (int i = 0; i < nc; ++i) pool.push_back(b[i]);
      while (pool.size() % 4) pool.push_back(0);
      o.inv_norm = 1.0 / in.norm_radius;
      features |= FEAT_EXTRA;          // the Forbes code lives in the general kernel only (olb_math.cuh::newton_sag)
    } else if (in.kind == OLB_GEOM_ZERNIKE) {
      if (!in_pool(in.coef_off, 4 * in.n_coef)) { res.error = "Zernike block outside pool"; return res; }
      if (!(in.norm_radius > 0)) { res.error = "Zernike norm_radius must be positive"; return res; }
      int deg = 0;
      for (int i = 0; i < in.n_coef; ++i) {
        const double* tm = tab.pool + in.coef_off + 4 * i;
        int n = (int)tm[0], m = (int)tm[1];
        if (n < 0 || (m < 0 ? -m : m) > n || ((n - m) & 1) || n > 23) { res.error = "bad Zernike (n, m)"; return res; }
        if (n > deg) deg = n;
      }
      // table width padded to one of the compile-time sizes of olb_math.cuh::poly_tri_value (zero rows / columns:
      // the nested Horner is unchanged); wider tables run the runtime-sized loops
      int degp = deg;
      for (int wp : {4, 8, 12})
        if (deg + 1 <= wp) { degp = wp - 1; break; }
      const int W = degp + 1;
      std::vector<double> S(W * W, 0.0), D(W * W, 0.0);
      for (int i = 0; i < in.n_coef; ++i) {
        const double* tm = tab.pool + in.coef_off + 4 * i;
        zernike_add_monomials((int)tm[0], (int)tm[1], tm[2], degp, S.data());  // c * N_nm
        zernike_add_monomials((int)tm[0], (int)tm[1], tm[3], degp, D.data());  // c (quirk: no N_nm)
      }
      o.poly_rows = W; o.poly_cols = W; o.flags |= PSF_POLY_TRI;
      while (pool.size() % 4) pool.push_back(0);     // 16-byte aligned rows: the kernel loads them as vectors
      o.coef_off = (int)pool.size();
      pool.insert(pool.end(), S.begin(), S.end());
      while (pool.size() % 4) pool.push_back(0);
      o.poly_d_off = (int)pool.size();
      pool.insert(pool.end(), D.begin(), D.end());
      while (pool.size() % 4) pool.push_back(0);
      o.inv_norm = 1.0 / in.norm_radius; o.inv_norm_y = o.inv_norm;
    }

    // ---- aperture ------------------------------------------------------------
    if (in.flags & OLB_SF_APERTURE) {
      if (!in_pool(in.aper_off, in.aper_len) || in.aper_len < 1) { res.error = "aperture program outside pool"; return res; }
      int i = 0, depth = 0;
      while (i < in.aper_len) {
        int op = (int)tab.pool[in.aper_off + i];
        int nops = aperture_operands(op);
        if (nops < 0) { res.error = "bad aperture opcode"; return res; }
        if (op >= OLB_AP_UNION) { if (depth < 2) { res.error = "aperture stack underflow"; return res; } depth -= 1; }
        else { depth += 1; if (depth > 8) { res.error = "aperture program too deep"; return res; } }
        i += 1 + nops;
      }
      if (i != in.aper_len || depth != 1) { res.error = "malformed aperture program"; return res; }
      o.aper_off = (int)pool.size();
      o.aper_len = in.aper_len;
      for (int k = 0; k < in.aper_len; ++k) pool.push_back(tab.pool[in.aper_off + k]);
      while (pool.size() % 4) pool.push_back(0);
      if ((int)tab.pool[in.aper_off] == OLB_AP_RADIAL && in.aper_len == 3) {
        o.flags |= PSF_APER_RADIAL;
        // store squared radii right after the program for the fast path
        double rmax = tab.pool[in.aper_off + 1], rmin = tab.pool[in.aper_off + 2];
        pool[o.aper_off + 1] = rmax * rmax;  // reference compares r2 <= r_max**2 (radial.py:68-69)
        pool[o.aper_off + 2] = rmin * rmin;
      } else {
        features |= FEAT_EXTRA;
      }
    }

    // ---- media -----------------------------------------------------------------
    if (!in_pool(in.media_off, 5 * n_wl)) { res.error = "media block outside pool"; return res; }
    o.media_off = (int)pool.size();
    bool absorbing = false;
    for (int j = 0; j < n_wl; ++j) {
      const double n1 = tab.pool[in.media_off + 0 * n_wl + j];
      const double n2 = tab.pool[in.media_off + 1 * n_wl + j];
      const double k1 = tab.pool[in.media_off + 2 * n_wl + j];
      const double c1 = tab.pool[in.media_off + 3 * n_wl + j];
      const double c2 = tab.pool[in.media_off + 4 * n_wl + j];
      if (k1 > 0) absorbing = true;
      pool.push_back(n1);
      pool.push_back(n1 / n2);
      pool.push_back(4.0 * M_PI * k1 / tab.wavelengths[j] * 1e3);
      pool.push_back(c2 / c1);
    }
    if (absorbing) o.flags |= OLB_SF_ABSORBING;
    else o.flags &= ~uint32_t(OLB_SF_ABSORBING);
    if (in.coating == OLB_COAT_SIMPLE) features |= FEAT_EXTRA;
    else if (in.coating == OLB_COAT_FRESNEL) features |= FEAT_POL;
    else if (in.coating != OLB_COAT_NONE) { res.error = "unknown coating"; return res; }
  }
  res.features = features;
  int gslot = 0;
  for (int s = 0; s < tab.n_surfaces; ++s) {
    ps[s].gslot = gslot;
    // 7 scalars + even-asphere coefficients + (tilted pose) the 9 entries of dLoss/dR
    // (a Forbes Q^bfs surface carries the gradients of its Clenshaw-basis coefficients in the same slots)
    const bool asph = ps[s].kind == OLB_GEOM_EVEN_ASPHERE || ps[s].kind == OLB_GEOM_ODD_ASPHERE ||
                      ps[s].kind == OLB_GEOM_FORBES_QBFS;
    ps[s].gslots = ps[s].kind == OLB_GEOM_NOOP ? 0 : 7 + (asph ? ps[s].n_coef : 0) +
                                                         ((ps[s].flags & OLB_SF_ROTATED) ? 9 : 0);
    gslot += ps[s].gslots;
  }
  res.total_gslots = gslot;
  for (int s = 0; s < tab.n_surfaces; ++s) {
    const PrepSurface<double>& o = ps[s];
    const bool polyfam = (o.kind == OLB_GEOM_POLYNOMIAL || o.kind == OLB_GEOM_ZERNIKE || o.kind == OLB_GEOM_CHEBYSHEV) &&
                         o.poly_rows <= 12 && o.poly_cols <= 12;
    const bool kind_ok = o.kind == OLB_GEOM_NOOP || o.kind == OLB_GEOM_PLANE || o.kind == OLB_GEOM_STANDARD ||
                         ((o.kind == OLB_GEOM_EVEN_ASPHERE || o.kind == OLB_GEOM_ODD_ASPHERE) && o.n_coef <= 12) || polyfam;
    // Forbes Q^bfs: covered by the general (olb_trace_bwd_tables_*) variant of the adjoint kernel, at most 12 terms
    const bool forbes = o.kind == OLB_GEOM_FORBES_QBFS && tab.surfaces[s].n_coef <= 12;
    if (polyfam || forbes) res.bwd_tables = true;
    if (!(kind_ok || forbes) || o.coating == OLB_COAT_FRESNEL || tab.n_wl != 1)
      res.bwd_supported = false;
  }
  build_blob<double>(tab, pools, ps, features, res.blob_f64);
  build_blob<float>(tab, pools, ps, features, res.blob_f32);
  return res;
}

// ---- batched systems: B perturbed copies of one template -----------------------------------
// `params`: n_systems x n_surfaces blocks of OLB_BP_COUNT doubles with ABSOLUTE values (include/olb.h).  Every
// system is prepared like a single table; the blobs must come out the same size (same structure) and are laid
// side by side; each header is stamped with the UNION of the feature bits (one kernel variant serves them all).
struct BatchPrep {
  std::vector<unsigned char> all64, all32;
  uint32_t features = 0, hints = 0;
  int32_t bytes_f64 = 0, bytes_f32 = 0;
  bool unsupported = false;
  std::string error;
};

static BatchPrep prepare_batch(const OlbTable& tmpl, const double* params, int n_systems) {
  BatchPrep out;
  if (tmpl.n_wl != 1) { out.error = "batched tables support one wavelength"; out.unsupported = true; return out; }
  const int S = tmpl.n_surfaces;
  std::vector<OlbSurface> surf(tmpl.surfaces, tmpl.surfaces + S);
  std::vector<double> pool(tmpl.pool, tmpl.pool + tmpl.pool_len);
  OlbTable t = tmpl;
  t.surfaces = surf.data();
  t.pool = pool.data();
  for (int b = 0; b < n_systems; ++b) {
    for (int s = 0; s < S; ++s) {
      const double* p = params + ((size_t)b * S + s) * OLB_BP_COUNT;
      OlbSurface& o = surf[s];
      const OlbSurface& o0 = tmpl.surfaces[s];
      if (o0.kind == OLB_GEOM_NOOP) continue;
      o.t[0] = p[OLB_BP_TX]; o.t[1] = p[OLB_BP_TY]; o.t[2] = p[OLB_BP_TZ];
      for (int q = 0; q < 9; ++q) o.R[q] = p[OLB_BP_R + q];
      if (o0.kind != OLB_GEOM_PLANE) {
        o.radius = p[OLB_BP_CURV] == 0 ? INFINITY : 1.0 / p[OLB_BP_CURV];
        if (o0.kind != OLB_GEOM_TOROIDAL) o.conic = p[OLB_BP_CONIC];
      }
      // media block (one wavelength): {n1, n2, k1, coating n1, coating n2}; without a Fresnel coating the last two
      // mirror n1 / n2 (table.py::pack)
      pool[o0.media_off + 0] = p[OLB_BP_N1];
      pool[o0.media_off + 1] = p[OLB_BP_N2];
      if (o0.coating != OLB_COAT_FRESNEL) {
        pool[o0.media_off + 3] = p[OLB_BP_N1];
        pool[o0.media_off + 4] = p[OLB_BP_N2];
      }
      if (o0.kind == OLB_GEOM_EVEN_ASPHERE)
        for (int j = 0; j < o0.n_coef && j < OLB_BP_MAX_COEF; ++j) pool[o0.coef_off + j] = p[OLB_BP_COEF + j];
    }
    PrepResult pr = prepare_table(t);
    if (!pr.error.empty()) { out.error = "system " + std::to_string(b) + ": " + pr.error; return out; }
    if (b == 0) {
      out.bytes_f64 = (int32_t)pr.blob_f64.size();
      out.bytes_f32 = (int32_t)pr.blob_f32.size();
    } else if ((int32_t)pr.blob_f64.size() != out.bytes_f64 || (int32_t)pr.blob_f32.size() != out.bytes_f32) {
      out.error = "batched systems must share one table structure";
      return out;
    }
    out.features |= pr.features;
    out.hints |= pr.hints;
    out.all64.insert(out.all64.end(), pr.blob_f64.begin(), pr.blob_f64.end());
    out.all32.insert(out.all32.end(), pr.blob_f32.begin(), pr.blob_f32.end());
  }
  for (int b = 0; b < n_systems; ++b) {
    reinterpret_cast<PrepHeader*>(out.all64.data() + (size_t)b * out.bytes_f64)->features = out.features;
    reinterpret_cast<PrepHeader*>(out.all32.data() + (size_t)b * out.bytes_f32)->features = out.features;
  }
  return out;
}

}  // namespace olb
#endif  // OLB_PREP_H_
