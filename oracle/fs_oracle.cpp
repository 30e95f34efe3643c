/*
 * fs_oracle.cpp -- CPU oracle (restatement of the reference's algorithm).
 *
 * TEST INFRASTRUCTURE -- see fs_oracle.h.  Written in C-style C++ so that the
 * Xrange operator overloads of the reference (numpy_utils/numba_xr.py) can be
 * restated one-to-one and the model formulas written once for the plain-double
 * and the Xrange instantiation, exactly as numba specialises them.
 *
 * Build: g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).
 * No -ffast-math: every operation below is a single IEEE-754 operation.
 */
#include "fs_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

/* ------------------------------------------------------------------------ */
/* bit helpers (numba_xr.py:674-706)                                         */
static inline int64_t d2b(double x) { int64_t b; memcpy(&b, &x, 8); return b; }
static inline double b2d(int64_t b) { double x; memcpy(&x, &b, 8); return x; }

/* numba_xr.py:390-411 : |biased_exponent - 1023| > 100 */
static inline bool need_renorm(double m)
{
    int64_t e = ((d2b(m) >> 52) & 0x7ff) - 1023;
    if (e < 0) e = -e;
    return e > 100;
}

/* numba_xr.py:674-685 : mantissa forced into [1, 2), raw exponent returned */
static inline void xr_frexp(double m, double *nm, int32_t *ne)
{
    int64_t bits = d2b(m);
    *nm = b2d((int64_t)(((uint64_t)bits & 0x8000000000000000ULL)
                        + (0x3ffULL << 52)
                        + ((uint64_t)bits & 0xfffffffffffffULL)));
    *ne = (int32_t)(((bits >> 52) & 0x7ff) - 0x3ff);
}

/* numba_xr.py:687-694 */
static inline void normalize_real(double m, int32_t exp, double *nm, int32_t *ne)
{
    if (m == 0.) { *nm = m; *ne = 0; return; }
    int32_t e;
    xr_frexp(m, nm, &e);
    *ne = exp + e;
}

/* numba_xr.py:696-706 : add to the exponent FIELD, clamp at 0, keep mantissa */
static inline double exp2_shift(double m, int32_t shift)
{
    int64_t bits = d2b(m);
    int64_t e = ((bits >> 52) & 0x7ff) + (int64_t)shift;
    if (e < 0) e = 0;
    return b2d((int64_t)(((uint64_t)bits & 0x8000000000000000ULL)
                         + ((uint64_t)e << 52)
                         + ((uint64_t)bits & 0xfffffffffffffULL)));
}

/* python-semantics max / min as numba lowers them (b > a ? b : a) */
static inline double pymax(double a, double b) { return (b > a) ? b : a; }
static inline double pymin(double a, double b) { return (b < a) ? b : a; }

/* ------------------------------------------------------------------------ */
/* complex128 with numba's lowering                                          */
struct C { double re, im; };
static inline C mkC(double r, double i) { C c; c.re = r; c.im = i; return c; }
static inline C operator+(C a, C b) { return mkC(a.re + b.re, a.im + b.im); }
static inline C operator-(C a, C b) { return mkC(a.re - b.re, a.im - b.im); }
static inline C operator*(C a, C b)
{
    return mkC(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
static inline C operator*(double s, C a) { return mkC(s * a.re, s * a.im); }
static inline C operator+(C a, double b) { return mkC(a.re + b, a.im); }
static inline bool is0(C a) { return a.re == 0. && a.im == 0.; }

/* ------------------------------------------------------------------------ */
/* Xrange scalars (numba_xr.py:79-116)                                       */
struct XF { double m; int32_t e; };
struct XC { C m; int32_t e; };
static inline XF mkXF(double m, int32_t e) { XF x; x.m = m; x.e = e; return x; }
static inline XC mkXC(C m, int32_t e) { XC x; x.m = m; x.e = e; return x; }

/* _coexp_ufunc, real implementation (numba_xr.py:716-733) */
static inline void coexp_f(double m0, int32_t e0, double m1, int32_t e1,
                           double *o0, double *o1, int32_t *oe)
{
    double c0 = m0, c1 = m1;
    int32_t d = e0 - e1, e;
    if (m0 == 0.) e = e1;
    else if (m1 == 0.) e = e0;
    else if (e1 > e0) { c0 = exp2_shift(c0, d); e = e1; }
    else if (e0 > e1) { c1 = exp2_shift(c1, -d); e = e0; }
    else e = e0;
    *o0 = c0; *o1 = c1; *oe = e;
}

/* _coexp_ufunc, complex implementation (numba_xr.py:735-754) */
static inline void coexp_c(C m0, int32_t e0, C m1, int32_t e1, C *o0, C *o1,
                           int32_t *oe)
{
    C c0 = m0, c1 = m1;
    int32_t d = e0 - e1, e;
    if (is0(m0)) e = e1;
    else if (is0(m1)) e = e0;
    else if (e1 > e0) { c0 = mkC(exp2_shift(c0.re, d), exp2_shift(c0.im, d)); e = e1; }
    else if (e0 > e1) { c1 = mkC(exp2_shift(c1.re, -d), exp2_shift(c1.im, -d)); e = e0; }
    else e = e0;
    *o0 = c0; *o1 = c1; *oe = e;
}

/* _normalize (numba_xr.py:650-672) */
static inline XF normalize(double m, int32_t e)
{
    XF r;
    normalize_real(m, e, &r.m, &r.e);
    return r;
}
static inline XC normalize(C m, int32_t e)
{
    double nre, nim, cre, cim;
    int32_t ere, eim, ce;
    normalize_real(m.re, e, &nre, &ere);
    normalize_real(m.im, e, &nim, &eim);
    coexp_f(nre, ere, nim, eim, &cre, &cim, &ce);
    return mkXC(mkC(cre, cim), ce);
}
static inline bool need_renorm(C m) { return need_renorm(m.re) || need_renorm(m.im); }

/* to_Xrange_scalar of a number (numba_xr.py:784-800) */
static inline XF to_xr(double v) { return normalize(v, 0); }
static inline XC to_xr(C v) { return normalize(v, 0); }

/* exact 2**e as a double, 0 / inf outside the range (np.ldexp(1., e)) */
static inline double ldexp1(int32_t e) { return ldexp(1., e); }

/* to_standard (numba_xr.py:802-829) */
static inline double to_std(XF x) { return ldexp(x.m, x.e); }
static inline C to_std(XC x)
{
    XC n = normalize(x.m, x.e);
    double s = ldexp1(n.e);
    return mkC(n.m.re * s, n.m.im * s);
}

/* ---- add / sub (numba_xr.py:318-388) ---- */
static inline XF operator+(XF a, XF b)
{
    double x, y; int32_t e;
    coexp_f(a.m, a.e, b.m, b.e, &x, &y, &e);
    return mkXF(x + y, e);
}
static inline XF operator-(XF a, XF b)
{
    double x, y; int32_t e;
    coexp_f(a.m, a.e, b.m, b.e, &x, &y, &e);
    return mkXF(x - y, e);
}
static inline XF as_operand(double v)
{
    /* base-type operand of a mixed add/sub: normalised only if needed */
    if (need_renorm(v)) return normalize(v, 0);
    return mkXF(v, 0);
}
static inline XC as_operand(C v)
{
    if (need_renorm(v)) return normalize(v, 0);
    return mkXC(v, 0);
}
static inline XF operator+(XF a, double b) { return a + as_operand(b); }
static inline XF operator-(XF a, double b) { return a - as_operand(b); }
static inline XF operator+(double a, XF b) { return as_operand(a) + b; }
static inline XF operator-(double a, XF b) { return as_operand(a) - b; }
static inline XF operator-(XF a) { return mkXF(-a.m, a.e); }

static inline XC operator+(XC a, XC b)
{
    C x, y; int32_t e;
    coexp_c(a.m, a.e, b.m, b.e, &x, &y, &e);
    return mkXC(x + y, e);
}
static inline XC operator+(XC a, C b) { return a + as_operand(b); }
/* complex Xrange + real Xrange (dZndc path: "+ scale_deriv_xr[0]") */
static inline XC operator+(XC a, XF b)
{
    C x, y; int32_t e;
    coexp_c(a.m, a.e, mkC(b.m, 0.), b.e, &x, &y, &e);
    return mkXC(x + y, e);
}

/* ---- mul (numba_xr.py:416-444) ---- */
static inline XF xr_pack(double m, int32_t e)
{
    if (need_renorm(m)) return normalize(m, e);
    return mkXF(m, e);
}
static inline XC xr_pack(C m, int32_t e)
{
    if (need_renorm(m)) return normalize(m, e);
    return mkXC(m, e);
}
static inline XF operator*(XF a, XF b) { return xr_pack(a.m * b.m, a.e + b.e); }
static inline XF operator*(XF a, double b) { return xr_pack(a.m * b, a.e); }
static inline XF operator*(double a, XF b) { return xr_pack(a * b.m, b.e); }
static inline XC operator*(XC a, XC b) { return xr_pack(a.m * b.m, a.e + b.e); }
static inline XC operator*(double a, XC b) { return xr_pack(a * b.m, b.e); }
static inline XC operator*(C a, XC b) { return xr_pack(a * b.m, b.e); }
static inline XC operator*(XF a, C b) { return xr_pack(a.m * b, a.e); }

/* ---- div (numba_xr.py:446-474), only used by the unit tests ---- */
static inline C cdiv(C a, C b)
{
    /* numba complex division: Smith-free textbook formula is NOT what numba
     * uses; it follows CPython's algorithm (complexobject.c _Py_c_quot). */
    double abs_breal = b.re < 0 ? -b.re : b.re;
    double abs_bimag = b.im < 0 ? -b.im : b.im;
    C r;
    if (abs_breal >= abs_bimag) {
        if (abs_breal == 0.) { r.re = NAN; r.im = NAN; }
        else {
            double ratio = b.im / b.re;
            double denom = b.re + b.im * ratio;
            r.re = (a.re + a.im * ratio) / denom;
            r.im = (a.im - a.re * ratio) / denom;
        }
    } else {
        double ratio = b.re / b.im;
        double denom = b.re * ratio + b.im;
        r.re = (a.re * ratio + a.im) / denom;
        r.im = (a.im * ratio - a.re) / denom;
    }
    return r;
}
static inline XF operator/(XF a, XF b) { return xr_pack(a.m / b.m, a.e - b.e); }
static inline XC operator/(XC a, XC b) { return xr_pack(cdiv(a.m, b.m), a.e - b.e); }

/* ---- compare (numba_xr.py:476-510) ---- */
static inline bool xr_le(XF a, XF b)
{
    double x, y; int32_t e;
    coexp_f(a.m, a.e, b.m, b.e, &x, &y, &e);
    return x <= y;
}
static inline bool xr_lt(XF a, double b)
{
    double x, y; int32_t e;
    coexp_f(a.m, a.e, b, 0, &x, &y, &e);
    return x < y;
}
static inline bool operator>=(XF a, double b)
{
    double x, y; int32_t e;
    coexp_f(a.m, a.e, b, 0, &x, &y, &e);
    return x >= y;
}
static inline bool operator<=(XF a, double b)
{
    double x, y; int32_t e;
    coexp_f(a.m, a.e, b, 0, &x, &y, &e);
    return x <= y;
}

/* extended_abs2 (numba_xr.py:531-561) */
static inline XF abs2(XC a)
{
    return mkXF(a.m.re * a.m.re + a.m.im * a.m.im, a.e + a.e);
}
static inline double fabs_(double x) { return fabs(x); }
/* np.abs of a real Xrange (numba_xr.py:563-572) */
static inline XF fabs_(XF x) { return mkXF(fabs(x.m), x.e); }

} /* namespace */

/* |x + iy| -- the reference uses np.abs(complex) = hypot.  The oracle and the
 * CUDA kernels share this exact definition (power-of-two pre-scaling, then
 * sqrt(a*a + b*b) with individually rounded operations) so that both sides
 * give bit-identical BLA radii; it is within 1 ulp of a correctly rounded
 * hypot. */
extern "C" double fso_hypot(double x, double y)
{
    double a = fabs(x), b = fabs(y);
    if (a < b) { double t = a; a = b; b = t; }
    if (!(a == a) || !(b == b)) return NAN;
    if (a == 0.) return 0.;
    if (a > 1.7976931348623157e308) return a;
    int64_t ea = (d2b(a) >> 52) & 0x7ff;
    double up = 1., down = 1.;
    if (ea > 1023 + 500) { up = 0x1p-600; down = 0x1p600; }
    else if (ea < 1023 - 500) { up = 0x1p600; down = 0x1p-600; }
    a = a * up;
    b = b * up;
    double s = a * a + b * b;
    return sqrt(s) * down;
}

namespace {

static inline double cabs_(C z) { return fso_hypot(z.re, z.im); }

/* ------------------------------------------------------------------------ */
/* pixel -> c                                                                */

/* core.py:3161-3194 (c_from_pix, lin_proj_impl; Cartesian proj = identity) */
static inline C c_from_pix(C pix, const double *lm, double dx, C center)
{
    double x1 = lm[0] * pix.re + lm[1] * pix.im;
    double y1 = lm[2] * pix.re + lm[3] * pix.im;
    return center + (dx * mkC(x1, y1));
}

/* perturbation.py:2214-2230,2666-2672 */
static inline XC c_xr_from_pix(C pix, const double *lm, XF lin_scale, XC drift)
{
    double x1 = lm[0] * pix.re + lm[1] * pix.im;
    double y1 = lm[2] * pix.re + lm[3] * pix.im;
    return (lin_scale * mkC(x1, y1)) + drift;
}

/* ------------------------------------------------------------------------ */
/* Standard loops                                                            */

static inline C m2_iterate(C z, C c) { return z * z + c; } /* mandelbrot_M2.py:13-15 */

/* core.py:2965-3004 + mandelbrot_M2.py:310-334 */
static void std_m2_pixel(C c, int64_t max_iter, double Mdiv_sq, double epscv_sq,
                         int calc_d2, int calc_orbit, int64_t backshift,
                         double *Z, int64_t stride, int8_t *stop, int32_t *niter)
{
    C zn = mkC(0., 0.), dzndz = zn, dzndc = zn, d2 = zn;
    int64_t n_iter = 0;
    int64_t div_shift = 0, orbit_i1 = 0, orbit_i2 = 0;
    C orbit_zn1 = zn, orbit_zn2 = zn;
    int8_t reason = -1;
    for (;;) {
        n_iter += 1;
        int ret = 0;
        if (n_iter >= max_iter) {
            reason = 0;
            ret = 1;
        } else {
            if (calc_d2) d2 = 2. * (d2 * zn + dzndc * dzndc);
            dzndc = (2. * dzndc) * zn + 1.;
            dzndz = (2. * dzndz) * zn;
            zn = zn * zn + c;
            if (n_iter == 1) dzndz = mkC(1., 0.);
            if (zn.re * zn.re + zn.im * zn.im > Mdiv_sq) { reason = 1; ret = 1; }
            else if (dzndz.re * dzndz.re + dzndz.im * dzndz.im < epscv_sq) { reason = 2; ret = 1; }
        }
        if (calc_orbit) {
            int64_t div = n_iter / backshift;
            if (div > div_shift) {
                div_shift = div;
                orbit_i2 = orbit_i1;
                orbit_zn2 = orbit_zn1;
                orbit_i1 = n_iter;
                orbit_zn1 = zn;
            }
        }
        if (ret) break;
    }
    int row = 0;
    Z[2 * stride * row] = zn.re; Z[2 * stride * row + 1] = zn.im; row++;
    Z[2 * stride * row] = dzndz.re; Z[2 * stride * row + 1] = dzndz.im; row++;
    Z[2 * stride * row] = dzndc.re; Z[2 * stride * row + 1] = dzndc.im; row++;
    if (calc_d2) { Z[2 * stride * row] = d2.re; Z[2 * stride * row + 1] = d2.im; row++; }
    if (calc_orbit) {
        C zo = orbit_zn2;
        while (orbit_i2 < n_iter - backshift) { zo = m2_iterate(zo, c); orbit_i2 += 1; }
        Z[2 * stride * row] = zo.re; Z[2 * stride * row + 1] = zo.im; row++;
    }
    *stop = reason;
    *niter = (int32_t)n_iter;
}

/* Power-N standard loop: core.py:2965-3004 + mandelbrot_Mn.py:300-350.
 * `_zn ** deg_m1` is numba's complex power: a product when the exponent is 2,
 * else numba_cpow -> CPython's _Py_c_pow, the polar form on top of the C
 * library (hypot, pow, atan2, cos, sin).
 *   use_cpow = 1  that polar form: what the reference executes (pinned by the
 *                 strict fixtures);
 *   use_cpow = 0  the power as a left-to-right product chain: the platform-
 *                 independent definition the CUDA library evaluates (identical
 *                 for N = 3 without d2zndc2, a few ulp away otherwise). */
static inline C c_pow_int(C a, int n, int use_cpow)
{
    if (n == 2) return a * a;                       /* numbers.py:1008-1022 */
    if (use_cpow) {
        if (n == 0) return mkC(1., 0.);
        if (a.re == 0. && a.im == 0.) return mkC(0., 0.);
        double vabs = hypot(a.re, a.im);
        double len = pow(vabs, (double)n);
        double at = atan2(a.im, a.re);
        double phase = at * (double)n;
        return mkC(len * cos(phase), len * sin(phase));
    }
    if (n == 0) return mkC(1., 0.);
    C r = a;
    for (int k = 1; k < n; k++) r = r * a;
    return r;
}

static void std_mn_pixel(int deg, int use_cpow, C c, int64_t max_iter, double Mdiv_sq,
                         double epscv_sq, int calc_d2, int calc_orbit, int64_t backshift,
                         double *Z, int64_t stride, int8_t *stop, int32_t *niter)
{
    C zn = mkC(0., 0.), dzndz = zn, dzndc = zn, d2 = zn;
    const double fdeg = (double)deg, fdeg_m1 = (double)(deg - 1);
    int64_t n_iter = 0, div_shift = 0, orbit_i1 = 0, orbit_i2 = 0;
    C orbit_zn1 = zn, orbit_zn2 = zn;
    int8_t reason = -1;
    for (;;) {
        n_iter += 1;
        int ret = 0;
        if (n_iter >= max_iter) { reason = 0; ret = 1; }
        else {
        C zn_m1, zn_m;
        if (calc_d2) {
            C zn_m2 = c_pow_int(zn, deg - 2, use_cpow);
            zn_m1 = zn_m2 * zn;
            zn_m = zn_m1 * zn;
            d2 = fdeg * (d2 * zn_m1 + ((fdeg_m1 * dzndz) * dzndc) * zn_m2);
        } else {
            zn_m1 = c_pow_int(zn, deg - 1, use_cpow);
            zn_m = zn_m1 * zn;
        }
        dzndc = (fdeg * dzndc) * zn_m1 + 1.;
        dzndz = (fdeg * dzndz) * zn_m1;
        zn = zn_m + c;
        if (n_iter == 1) dzndz = mkC(1., 0.);
        if (zn.re * zn.re + zn.im * zn.im > Mdiv_sq) { reason = 1; ret = 1; }
        else if (dzndz.re * dzndz.re + dzndz.im * dzndz.im < epscv_sq) { reason = 2; ret = 1; }
        }
        if (calc_orbit) {                       /* core.py:3034-3046 */
            int64_t div = n_iter / backshift;
            if (div > div_shift) {
                div_shift = div;
                orbit_i2 = orbit_i1; orbit_zn2 = orbit_zn1;
                orbit_i1 = n_iter; orbit_zn1 = zn;
            }
        }
        if (ret) break;
    }
    int row = 0;
    Z[2 * stride * row] = zn.re; Z[2 * stride * row + 1] = zn.im; row++;
    Z[2 * stride * row] = dzndz.re; Z[2 * stride * row + 1] = dzndz.im; row++;
    Z[2 * stride * row] = dzndc.re; Z[2 * stride * row + 1] = dzndc.im; row++;
    if (calc_d2) { Z[2 * stride * row] = d2.re; Z[2 * stride * row + 1] = d2.im; row++; }
    if (calc_orbit) {       /* back-shift with zn_iterate = zn ** N + c (mandelbrot_Mn.py:13-17) */
        C zo = orbit_zn2;
        while (orbit_i2 < n_iter - backshift) { zo = c_pow_int(zo, deg, use_cpow) + c; orbit_i2 += 1; }
        Z[2 * stride * row] = zo.re; Z[2 * stride * row + 1] = zo.im; row++;
    }
    *stop = reason;
    *niter = (int32_t)n_iter;
}

static inline double sgn(double x) { return (x < 0.) ? -1. : 1.; } /* burning_ship.py:12-17 */

/* burning_ship.py:82-122 */
static inline void bs_iterate(int flavor, double xn, double yn, double a, double b,
                              double *ox, double *oy)
{
    switch (flavor) {
    case 1: *ox = xn * xn - yn * yn + a; *oy = 2. * fabs(xn * yn) - b; break;
    case 2: *ox = xn * xn - yn * yn + a; *oy = 2. * xn * fabs(yn) - b; break;
    case 3: *ox = xn * xn - yn * fabs(yn) + a; *oy = 2. * xn * yn - b; break;
    case 4: *ox = fabs(xn * xn - yn * yn) + a; *oy = 2. * xn * yn - b; break;
    default: *ox = fabs(xn * xn - yn * yn) + a; *oy = 2. * fabs(xn * yn) - b; break;
    }
}

/* core.py:3007-3056 + burning_ship.py:351-423 */
static void std_bs_pixel(int flavor, C c, int64_t max_iter, double Mdiv_sq,
                         int calc_orbit, int64_t backshift, double *Z,
                         int64_t stride, int8_t *stop, int32_t *niter)
{
    double a = c.re, b = c.im;
    double X = 0., Y = 0., dXdA = 0., dXdB = 0., dYdA = 0., dYdB = 0.;
    int64_t n_iter = 0, div_shift = 0, orbit_i1 = 0, orbit_i2 = 0;
    double oxn1 = 0., oxn2 = 0., oyn1 = 0., oyn2 = 0.;
    int8_t reason = -1;
    for (;;) {
        n_iter += 1;
        int ret = 0;
        if (n_iter >= max_iter) { reason = 0; ret = 1; }
        else {
            double nx, ny, ndxa, ndxb, ndya, ndyb;
            switch (flavor) {
            case 1:
                nx = X * X - Y * Y + a;
                ny = 2. * fabs(X * Y) - b;
                ndxa = 2. * (X * dXdA - Y * dYdA) + 1.;
                ndxb = 2. * (X * dXdB - Y * dYdB);
                ndya = 2. * (fabs(X) * sgn(Y) * dYdA + sgn(X) * dXdA * fabs(Y));
                ndyb = 2. * (fabs(X) * sgn(Y) * dYdB + sgn(X) * dXdB * fabs(Y)) - 1.;
                break;
            case 2:
                nx = X * X - Y * Y + a;
                ny = 2. * X * fabs(Y) - b;
                ndxa = 2. * (X * dXdA - Y * dYdA) + 1.;
                ndxb = 2. * (X * dXdB - Y * dYdB);
                ndya = 2. * (X * sgn(Y) * dYdA + dXdA * fabs(Y));
                ndyb = 2. * (X * sgn(Y) * dYdB + dXdB * fabs(Y)) - 1.;
                break;
            case 3:
                nx = X * X - Y * fabs(Y) + a;
                ny = 2. * X * Y - b;
                ndxa = 2. * (X * dXdA - fabs(Y) * dYdA) + 1.;
                ndxb = 2. * (X * dXdB - fabs(Y) * dYdB);
                ndya = 2. * (dXdA * Y + X * dYdA);
                ndyb = 2. * (dXdB * Y + X * dYdB) - 1.;
                break;
            case 4: {
                double x2my2 = X * X - Y * Y;
                nx = fabs(x2my2) + a;
                ny = 2. * X * Y - b;
                ndxa = 2. * sgn(x2my2) * (X * dXdA - Y * dYdA);
                ndxb = 2. * sgn(x2my2) * (X * dXdB - Y * dYdB);
                ndya = 2. * (dXdA * Y + X * dYdA);
                ndyb = 2. * (dXdB * Y + X * dYdB) - 1.;
                break;
            }
            default: {
                double x2my2 = X * X - Y * Y;
                nx = fabs(x2my2) + a;
                ny = 2. * fabs(X * Y) - b;
                ndxa = 2. * sgn(x2my2) * (X * dXdA - Y * dYdA);
                ndxb = 2. * sgn(x2my2) * (X * dXdB - Y * dYdB);
                ndya = 2. * (fabs(X) * sgn(Y) * dYdA + sgn(X) * dXdA * fabs(Y));
                ndyb = 2. * (fabs(X) * sgn(Y) * dYdB + sgn(X) * dXdB * fabs(Y)) - 1.;
                break;
            }
            }
            X = nx; Y = ny; dXdA = ndxa; dXdB = ndxb; dYdA = ndya; dYdB = ndyb;
            if (X * X + Y * Y > Mdiv_sq) { reason = 1; ret = 1; }
        }
        if (calc_orbit) {
            int64_t div = n_iter / backshift;
            if (div > div_shift) {
                div_shift = div;
                orbit_i2 = orbit_i1; oxn2 = oxn1; oyn2 = oyn1;
                orbit_i1 = n_iter; oxn1 = X; oyn1 = Y;
            }
        }
        if (ret) break;
    }
    Z[0 * stride] = X; Z[1 * stride] = Y;
    Z[2 * stride] = dXdA; Z[3 * stride] = dXdB; Z[4 * stride] = dYdA; Z[5 * stride] = dYdB;
    if (calc_orbit) {
        double xo = oxn2, yo = oyn2;
        while (orbit_i2 < n_iter - backshift) {
            double tx, ty;
            bs_iterate(flavor, xo, yo, a, b, &tx, &ty);
            xo = tx; yo = ty; orbit_i2 += 1;
        }
        Z[6 * stride] = xo; Z[7 * stride] = yo;
    }
    *stop = reason;
    *niter = (int32_t)n_iter;
}

/* ------------------------------------------------------------------------ */
/* Reference path access                                                     */

/* perturbation.py:2519-2588 : stateless restatement of the cursor -- returns
 * the position of idx in the sorted ref_index_xr, or -1. */
static inline int64_t xr_find(const int32_t *index, int64_t n, int64_t idx)
{
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (index[mid] < idx) lo = mid + 1; else hi = mid;
    }
    if (lo < n && index[lo] == idx) return lo;
    return -1;
}

static inline C path_c(const double *p, int64_t i) { return mkC(p[2 * i], p[2 * i + 1]); }

/* ------------------------------------------------------------------------ */
/* BLA lookup, perturbation.py:2108-2170                                     */
static inline int64_t bla_index(int64_t i, int stg) { return 2 * i + (((int64_t)1 << stg) - 1); }

static inline int64_t ref_bla_get(const double *r_bla, int stages_bla, C zn,
                                  int64_t n_iter, int64_t first_invalid,
                                  int64_t *index_out)
{
    if (stages_bla <= 3) return 0;
    int64_t it = n_iter >> 3;
    int stages = stages_bla - 1;
    for (int s = 3; s < stages_bla; s++) {
        if (it & 1) { stages = s; break; }
        it >>= 1;
    }
    int64_t invalid_step = first_invalid - n_iter;
    for (int stg = stages; stg > 2; stg--) {
        int64_t step = (int64_t)1 << stg;
        if (step >= invalid_step) continue;
        int64_t ib = bla_index(n_iter / 8, stg - 3);
        double r = r_bla[ib];
        if (cabs_(zn) < r) { *index_out = ib; return step; }
    }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Holomorphic model formulas, written once (mandelbrot_M2.py:591-627)       */
template <class T, class R>
static inline T p_iter_zn(T z, R ref_zn, T c) { return z * (z + 2. * ref_zn) + c; }
template <class T, class R, class D>
static inline T p_iter_deriv(T z, T dz, R ref_zn, D ref_d)
{
    return 2. * ((ref_zn + z) * dz + ref_d * z);
}

/* Power-N Mandelbrot, models/mandelbrot_Mn.py:628-742: full binomial
 * expansions, C_binom[k] = comb(N, k) as float64 */
template <class T>
static inline T mn_dfdz(int nexp, T z)               /* :643-649 */
{
    T tmp = z;
    for (int k = 2; k < nexp; k++) tmp = tmp * z;
    return (double)nexp * tmp;
}
template <class T> static inline T dfdz_(int nexp, T z) { return nexp ? mn_dfdz(nexp, z) : 2. * z; }
template <class T, class R>
static inline T mn_iter_zn(int nexp, const double *Cb, T z, R ref_zn, T c)   /* :656-668 */
{
    T tmp = z * (z + Cb[1] * ref_zn);
    R pk = ref_zn;
    for (int k = 2; k < nexp; k++) {
        pk = pk * ref_zn;
        tmp = z * (tmp + Cb[k] * pk);
    }
    return tmp + c;
}
template <class T, class R, class D>
static inline T mn_iter_deriv(int nexp, const double *Cb, T z, T dz, R ref_zn, D ref_d) /* :670-728 */
{
    T mul = z + Cb[1] * ref_zn;
    T tmp = z * mul;
    T dtmp = dz * mul + z * (dz + Cb[1] * ref_d);
    R pk = ref_zn;
    for (int k = 2; k < nexp; k++) {
        D dpk = ((double)k * pk) * ref_d;
        pk = pk * ref_zn;
        mul = tmp + Cb[k] * pk;
        dtmp = dz * mul + z * (dtmp + Cb[k] * dpk);
        tmp = z * mul;
    }
    return dtmp;
}
static inline void binomials(int nexp, double *Cb)
{
    Cb[0] = 1.;
    for (int k = 1; k <= nexp; k++) Cb[k] = Cb[k - 1] * (double)(nexp - k + 1) / (double)k;
}
#define FSO_MAX_NEXP 32

/* perturbation.py:1065-1400 */
template <bool XR>
static void perturb_m2_pixel(const fso_frame_m2 *f, C pix, double *Z,
                             int64_t stride, int32_t *U, int8_t *stop_out,
                             int32_t *niter_out, int64_t *cnt)
{
    const int64_t L = f->L;
    const bool has_xr = f->n_xr > 0;
    const bool calc_dzndc = f->calc_dzndc, calc_dzndz = f->calc_dzndz;
    const int64_t ref_order = f->ref_order, ref_div_iter = f->ref_div_iter;
    const int64_t max_iter = f->max_iter;
    const int nexp = f->nexp;           /* 0: Perturbation_mandelbrot ; N: ..._mandelbrot_N */
    double Cb[FSO_MAX_NEXP + 1];
    if (nexp > 0) binomials(nexp, Cb);

    /* perturbation.py:1026-1031 */
    XC c_xr = c_xr_from_pix(pix, f->lin_mat, mkXF(f->lin_scale, f->lin_scale_e),
                            mkXC(mkC(f->drift[0], f->drift[1]), f->drift_e));
    C c = to_std(c_xr);

    C zn = mkC(0., 0.), dzndc = zn, dzndz = zn;      /* Z[...]    */
    XC zn_x = to_xr(zn), dzndc_x = zn_x, dzndz_x = zn_x; /* Z_xr[...] */
    const XC record_zero = mkXC(mkC(0., 0.), 0);

    int64_t w_iter = 0, n_iter = 0;
    if (w_iter >= ref_order) w_iter = w_iter % ref_order;

    int64_t div_shift = 0, orbit_i1 = 0, orbit_i2 = 0;
    C orbit_zn1 = zn, orbit_zn2 = zn;

    bool nullify_dZndz = false;
    const int64_t w_wraped = L; /* len(dZndz_path) - 1 */

    int64_t first_invalid = L;
    if (ref_div_iter < first_invalid) first_invalid = ref_div_iter;
    if (ref_order < first_invalid) first_invalid = ref_order;

    bool bool_dyn_rebase = true;
    int8_t stop = -1;

#define DZNDC_X(i) mkXC(path_c(f->dZndc, (i)), f->dZndc_e[(i)])
#define DZNDZ_X(i) mkXC(path_c(f->dZndz, (i)), f->dZndz_e[(i)])

    for (;;) {
        /* ---- BLA step, :1121-1154 ---- */
        if (f->bla_activated && (w_iter & 7) == 0) {
            int64_t ib = 0;
            int64_t step = ref_bla_get(f->r_bla, f->stages_bla, zn, w_iter,
                                       first_invalid, &ib);
            if (step != 0) {
                C A = path_c(f->M_bla, 2 * ib), B = path_c(f->M_bla, 2 * ib + 1);
                n_iter += step;
                w_iter = (w_iter + step) % ref_order;
                if (XR) {
                    zn_x = A * zn_x + B * c_xr;
                    zn = to_std(zn_x);
                    if (calc_dzndc) dzndc_x = A * dzndc_x;
                    if (calc_dzndz) dzndz_x = A * dzndz_x;
                } else {
                    zn = A * zn + B * c;
                    if (calc_dzndc) dzndc = A * dzndc;
                    if (calc_dzndz) dzndz = A * dzndz;
                }
                cnt[1] += 1;
                continue;
            }
        }

        /* ---- full perturbation iteration, :1158-1209 ---- */
        n_iter += 1;
        cnt[0] += 1;
        C ref_zn = path_c(f->Zn_path, w_iter);
        XC ref_zn_x = record_zero;
        if (XR) {
            int64_t k = has_xr ? xr_find(f->ref_index_xr, f->n_xr, w_iter) : -1;
            if (k >= 0) ref_zn_x = mkXC(path_c(f->ref_xr, k), f->ref_xr_e[k]);
            else ref_zn_x = to_xr(ref_zn);
        }

        if (calc_dzndc) {
            if (XR) {
                XC ref_d = bool_dyn_rebase ? record_zero : DZNDC_X(w_iter);
                dzndc_x = nexp ? mn_iter_deriv(nexp, Cb, zn_x, dzndc_x, ref_zn_x, ref_d)
                               : p_iter_deriv(zn_x, dzndc_x, ref_zn_x, ref_d);
            } else {
                C ref_d = bool_dyn_rebase ? mkC(0., 0.) : path_c(f->dZndc, w_iter);
                dzndc = nexp ? mn_iter_deriv(nexp, Cb, zn, dzndc, ref_zn, ref_d)
                             : p_iter_deriv(zn, dzndc, ref_zn, ref_d);
            }
        }
        if (calc_dzndz) {
            int64_t i = nullify_dZndz ? 0 : w_iter;
            if (XR) dzndz_x = nexp ? mn_iter_deriv(nexp, Cb, zn_x, dzndz_x, ref_zn_x, DZNDZ_X(i))
                                   : p_iter_deriv(zn_x, dzndz_x, ref_zn_x, DZNDZ_X(i));
            else dzndz = nexp ? mn_iter_deriv(nexp, Cb, zn, dzndz, ref_zn, path_c(f->dZndz, i))
                              : p_iter_deriv(zn, dzndz, ref_zn, path_c(f->dZndz, i));
        }
        if (XR) {
            zn_x = nexp ? mn_iter_zn(nexp, Cb, zn_x, ref_zn_x, c_xr) : p_iter_zn(zn_x, ref_zn_x, c_xr);
            zn = to_std(zn_x);
        } else {
            zn = nexp ? mn_iter_zn(nexp, Cb, zn, ref_zn, c) : p_iter_zn(zn, ref_zn, c);
        }

        w_iter += 1;
        if (w_iter >= ref_order) w_iter = w_iter % ref_order;

        /* ---- stop: max_iter, :1218 ---- */
        if (n_iter >= max_iter) { stop = 0; break; }

        /* ---- stop: interior, :1224-1246 ---- */
        if (calc_dzndz) {
            int64_t i = 0;
            if (!nullify_dZndz) {
                i = w_iter;
                if (n_iter == ref_order) i = w_wraped;
            }
            bool stationnary;
            if (XR) {
                XC ZdZ = dzndz_x + DZNDZ_X(i);
                stationnary = xr_lt(abs2(ZdZ), f->epsilon_stationnary_sq);
            } else {
                C ZdZ = dzndz + path_c(f->dZndz, i);
                stationnary = (ZdZ.re * ZdZ.re + ZdZ.im * ZdZ.im
                               < f->epsilon_stationnary_sq);
            }
            if (stationnary) { stop = 2; break; }
        }

        /* ---- stop: divergence, :1252-1279 ---- */
        C ref_zn_next = path_c(f->Zn_path, w_iter);
        int64_t knext = -1;
        if (XR && has_xr) knext = xr_find(f->ref_index_xr, f->n_xr, w_iter);
        C ZZ = zn + ref_zn_next;
        double full_sq_norm = ZZ.re * ZZ.re + ZZ.im * ZZ.im;

        if (f->calc_orbit) {
            int64_t div = n_iter / f->backshift;
            if (div > div_shift) {
                div_shift = div;
                orbit_i2 = orbit_i1; orbit_zn2 = orbit_zn1;
                orbit_i1 = n_iter; orbit_zn1 = ZZ;
            }
        }
        if (full_sq_norm > f->M_divergence_sq) { stop = 1; break; }

        /* ---- rebase: reference diverging, :1283-1313 ---- */
        if (w_iter >= ref_div_iter - 1) {
            zn = ZZ;
            if (XR) {
                zn_x = to_xr(ZZ);
                if (calc_dzndc) dzndc_x = dzndc_x + DZNDC_X(w_iter);
                if (calc_dzndz) {
                    if (!nullify_dZndz) {
                        int64_t i = w_iter;
                        if (n_iter == ref_order) i = w_wraped;
                        dzndz_x = dzndz_x + DZNDZ_X(i);
                    }
                    nullify_dZndz = true;
                }
            } else {
                if (calc_dzndc) dzndc = dzndc + path_c(f->dZndc, w_iter);
                if (calc_dzndz) {
                    if (!nullify_dZndz) {
                        int64_t i = (n_iter == ref_order) ? w_wraped : w_iter;
                        dzndz = dzndz + path_c(f->dZndz, i);
                    }
                    nullify_dZndz = true;
                }
            }
            w_iter = 0;
            cnt[2] += 1;
            continue;
        }

        /* ---- rebase: dynamic glitch, :1317-1372 ---- */
        bool_dyn_rebase = (fabs(ZZ.re) <= fabs(zn.re)) && (fabs(ZZ.im) <= fabs(zn.im));
        if (bool_dyn_rebase) {
            if (XR) {
                XC Z_xrn = zn_x, ZZ_xr;
                if (knext >= 0) ZZ_xr = Z_xrn + mkXC(path_c(f->ref_xr, knext), f->ref_xr_e[knext]);
                else ZZ_xr = Z_xrn + ref_zn_next;
                if (xr_le(abs2(ZZ_xr), abs2(Z_xrn))) {
                    zn_x = ZZ_xr;
                    zn = to_std(ZZ_xr);
                    if (calc_dzndc) dzndc_x = dzndc_x + DZNDC_X(w_iter);
                    if (calc_dzndz) {
                        if (!nullify_dZndz) {
                            int64_t i = w_iter;
                            if (n_iter == ref_order) i = w_wraped;
                            dzndz_x = dzndz_x + DZNDZ_X(i);
                        }
                        nullify_dZndz = true;
                    }
                    w_iter = 0;
                    cnt[2] += 1;
                    continue;
                }
            } else {
                zn = ZZ;
                if (calc_dzndc) dzndc = dzndc + path_c(f->dZndc, w_iter);
                if (calc_dzndz) {
                    if (!nullify_dZndz) {
                        int64_t i = (n_iter == ref_order) ? w_wraped : w_iter;
                        dzndz = dzndz + path_c(f->dZndz, i);
                    }
                    nullify_dZndz = true;
                }
                w_iter = 0;
                cnt[2] += 1;
                continue;
            }
        }
    }

    /* ---- epilogue, :1374-1398 ----
     * Quirk: a pixel that walks the whole stored orbit by BLA steps without a
     * single rebase ends with w_iter == L (stop_iter = max_iter + 1) and the
     * reference then reads Zn_path[L] / dZndc_path[L], one element past the
     * arrays (numba does not bounds-check): undefined in the reference.  The
     * oracle and the CUDA path define that element as 0. */
    U[0] = (int32_t)w_iter;
    const bool oob = (w_iter >= L);
    if (XR) {
        zn = to_std(zn_x) + (oob ? mkC(0., 0.) : path_c(f->Zn_path, w_iter));
        if (calc_dzndc) dzndc = to_std(dzndc_x + (oob ? record_zero : DZNDC_X(w_iter)));
    } else {
        zn = zn + (oob ? mkC(0., 0.) : path_c(f->Zn_path, w_iter));
        if (calc_dzndc) dzndc = dzndc + (oob ? mkC(0., 0.) : path_c(f->dZndc, w_iter));
    }
#undef DZNDC_X
#undef DZNDZ_X
    int row = 0;
    Z[2 * stride * row] = zn.re; Z[2 * stride * row + 1] = zn.im; row++;
    if (calc_dzndz) { Z[2 * stride * row] = dzndz.re; Z[2 * stride * row + 1] = dzndz.im; row++; }
    if (calc_dzndc) { Z[2 * stride * row] = dzndc.re; Z[2 * stride * row + 1] = dzndc.im; row++; }
    if (f->calc_orbit) {
        C zo = orbit_zn2;
        C CC = c + path_c(f->Zn_path, 1);
        while (orbit_i2 < n_iter - f->backshift) {
            /* zn_iterate: zn * zn + c, or zn ** N + c for Perturbation_mandelbrot_N
             * (polar form when f->use_cpow, as the reference runs it) */
            zo = (f->nexp > 2) ? c_pow_int(zo, f->nexp, f->use_cpow) + CC : m2_iterate(zo, CC);
            orbit_i2 += 1;
        }
        Z[2 * stride * row] = zo.re; Z[2 * stride * row + 1] = zo.im; row++;
    }
    *stop_out = stop;
    *niter_out = (int32_t)n_iter;
}

/* ------------------------------------------------------------------------ */
/* Burning-ship family: formulas written once for double and XF              */
/* burning_ship.py:19-60 */
template <class T> static inline T diffabs(T X, T x)
{
    if (X >= 0.) {
        if ((X + x) >= 0.) return 1. * x;
        return -(2. * X + x);
    }
    if ((X + x) <= 0.) return -x;
    return (2. * X + x);
}
template <class T> static inline double ddiffabsdX(T X, T x)
{
    if (X >= 0.) { if ((X + x) >= 0.) return 0.; return -2.; }
    if ((X + x) <= 0.) return 0.;
    return 2.;
}
template <class T> static inline double ddiffabsdx(T X, T x)
{
    if (X >= 0.) { if ((X + x) >= 0.) return 1.; return -1.; }
    if ((X + x) <= 0.) return -1.;
    return 1.;
}
static inline double sgn_(double x) { return sgn(x); }
static inline double sgn_(XF x) { return (x.m < 0.) ? -1. : 1.; }

/* burning_ship.py:535-619 */
template <class T>
static inline void bs_p_iter_zn(int flavor, T &x, T &y, T rx, T ry, T a, T b)
{
    T nx, ny;
    switch (flavor) {
    case 1: {
        T rxy = rx * ry;
        nx = x * (x + 2. * rx) - y * (y + 2. * ry) + a;
        ny = 2. * diffabs(rxy, x * y + x * ry + y * rx) - b;
        break;
    }
    case 2:
        nx = x * (x + 2. * rx) - y * (y + 2. * ry) + a;
        ny = 2. * (rx * diffabs(ry, y) + x * fabs_(ry + y)) - b;
        break;
    case 3:
        nx = x * (x + 2. * rx) - ry * diffabs(ry, y) - y * fabs_(ry + y) + a;
        ny = 2. * (rx * y + ry * x + x * y) - b;
        break;
    case 4: {
        T r2 = rx * rx - ry * ry;
        nx = diffabs(r2, x * (x + 2. * rx) - y * (y + 2. * ry)) + a;
        ny = 2. * (rx * y + ry * x + x * y) - b;
        break;
    }
    default: {
        T rxy = rx * ry;
        T r2 = rx * rx - ry * ry;
        nx = diffabs(r2, x * (x + 2. * rx) - y * (y + 2. * ry)) + a;
        ny = 2. * diffabs(rxy, x * y + x * ry + y * rx) - b;
        break;
    }
    }
    x = nx; y = ny;
}

/* burning_ship.py:622-859 */
template <class T>
static inline void bs_p_iter_hessian(int flavor, T x, T y, T &dxa, T &dxb, T &dya,
                                     T &dyb, T rx, T ry, T rdxa, T rdxb, T rdya,
                                     T rdyb)
{
    T ndxa, ndxb, ndya, ndyb;
    switch (flavor) {
    case 1: {
        T opX = rx * ry;
        T dXa = rdxa * ry + rx * rdya;
        T dXb = rdxb * ry + rx * rdyb;
        T opx = x * y + x * ry + y * rx;
        T dxa_ = dxa * y + x * dya + dxa * ry + x * rdya + dya * rx + y * rdxa;
        T dxb_ = dxb * y + x * dyb + dxb * ry + x * rdyb + dyb * rx + y * rdxb;
        double dX = ddiffabsdX(opX, opx), dx = ddiffabsdx(opX, opx);
        ndxa = 2. * ((rx + x) * dxa + rdxa * x) - 2. * ((ry + y) * dya + rdya * y);
        ndxb = 2. * ((rx + x) * dxb + rdxb * x) - 2. * ((ry + y) * dyb + rdyb * y);
        ndya = 2. * (dX * dXa + dx * dxa_);
        ndyb = 2. * (dX * dXb + dx * dxb_);
        break;
    }
    case 2: {
        T da = diffabs(ry, y);
        double dX = ddiffabsdX(ry, y), dx = ddiffabsdx(ry, y);
        T Yy = ry + y;
        T ab = fabs_(Yy);
        double sg = sgn_(Yy);
        ndxa = 2. * (((rx + x) * dxa + rdxa * x) - ((ry + y) * dya + rdya * y));
        ndxb = 2. * (((rx + x) * dxb + rdxb * x) - ((ry + y) * dyb + rdyb * y));
        ndya = 2. * (rdxa * da + rx * (dX * rdya + dx * dya) + dxa * ab + x * sg * (rdya + dya));
        ndyb = 2. * (rdxb * da + rx * (dX * rdyb + dx * dyb) + dxb * ab + x * sg * (rdyb + dyb));
        break;
    }
    case 3: {
        T da = diffabs(ry, y);
        double dX = ddiffabsdX(ry, y), dx = ddiffabsdx(ry, y);
        T Yy = ry + y;
        T ab = fabs_(Yy);
        double sg = sgn_(Yy);
        ndxa = dxa * (x + 2. * rx) + x * (dxa + 2. * rdxa) - rdya * da
               - ry * (rdya * dX + dya * dx) - dya * ab - y * sg * (rdya + dya);
        ndxb = dxb * (x + 2. * rx) + x * (dxb + 2. * rdxb) - rdyb * da
               - ry * (rdyb * dX + dyb * dx) - dyb * ab - y * sg * (rdyb + dyb);
        ndya = 2. * (rdxa * y + rx * dya + rdya * x + ry * dxa + dxa * y + x * dya);
        ndyb = 2. * (rdxb * y + rx * dyb + rdyb * x + ry * dxb + dxb * y + x * dyb);
        break;
    }
    case 4: {
        T opX = rx * rx - ry * ry;
        T dXa = 2. * (rx * rdxa - ry * rdya);
        T dXb = 2. * (rx * rdxb - ry * rdyb);
        T opx = x * (x + 2. * rx) - y * (y + 2. * ry);
        T dxa_ = dxa * (x + 2. * rx) + x * (dxa + 2. * rdxa) - dya * (y + 2. * ry) - y * (dya + 2. * rdya);
        T dxb_ = dxb * (x + 2. * rx) + x * (dxb + 2. * rdxb) - dyb * (y + 2. * ry) - y * (dyb + 2. * rdyb);
        double dX = ddiffabsdX(opX, opx), dx = ddiffabsdx(opX, opx);
        ndxa = dX * dXa + dx * dxa_;
        ndxb = dX * dXb + dx * dxb_;
        ndya = 2. * (rdxa * y + rx * dya + rdya * x + ry * dxa + dxa * y + x * dya);
        ndyb = 2. * (rdxb * y + rx * dyb + rdyb * x + ry * dxb + dxb * y + x * dyb);
        break;
    }
    default: {
        T opX = rx * rx - ry * ry;
        T dXa = 2. * (rx * rdxa - ry * rdya);
        T dXb = 2. * (rx * rdxb - ry * rdyb);
        T opx = x * (x + 2. * rx) - y * (y + 2. * ry);
        T dxa_ = dxa * (x + 2. * rx) + x * (dxa + 2. * rdxa) - dya * (y + 2. * ry) - y * (dya + 2. * rdya);
        T dxb_ = dxb * (x + 2. * rx) + x * (dxb + 2. * rdxb) - dyb * (y + 2. * ry) - y * (dyb + 2. * rdyb);
        double dX = ddiffabsdX(opX, opx), dx = ddiffabsdx(opX, opx);
        ndxa = dX * dXa + dx * dxa_;
        ndxb = dX * dXb + dx * dxb_;
        T opX2 = rx * ry;
        T dXa2 = rdxa * ry + rx * rdya;
        T dXb2 = rdxb * ry + rx * rdyb;
        T opx2 = x * y + x * ry + y * rx;
        T dxa2 = dxa * y + x * dya + dxa * ry + x * rdya + dya * rx + y * rdxa;
        T dxb2 = dxb * y + x * dyb + dxb * ry + x * rdyb + dyb * rx + y * rdxb;
        double dX2 = ddiffabsdX(opX2, opx2), dx2 = ddiffabsdx(opX2, opx2);
        ndya = 2. * (dX2 * dXa2 + dx2 * dxa2);
        ndyb = 2. * (dX2 * dXb2 + dx2 * dxb2);
        break;
    }
    }
    dxa = ndxa; dxb = ndxb; dya = ndya; dyb = ndyb;
}

/* burning_ship.py:441-532 */
static inline void bs_jac(int flavor, double x, double y, double *fxx, double *fxy,
                          double *fyx, double *fyy)
{
    switch (flavor) {
    case 1: *fxx = 2. * x; *fxy = -2. * y; *fyx = 2. * sgn(x) * fabs(y); *fyy = 2. * sgn(y) * fabs(x); break;
    case 2: *fxx = 2. * x; *fxy = -2. * y; *fyx = 2. * fabs(y); *fyy = 2. * sgn(y) * x; break;
    case 3: *fxx = 2. * x; *fxy = -2. * fabs(y); *fyx = 2. * y; *fyy = 2. * x; break;
    case 4: { double s = sgn(x * x - y * y); *fxx = 2. * s * x; *fxy = -2. * s * y; *fyx = 2. * y; *fyy = 2. * x; break; }
    default: { double s = sgn(x * x - y * y); *fxx = 2. * s * x; *fxy = -2. * s * y; *fyx = 2. * sgn(x) * fabs(y); *fyy = 2. * sgn(y) * fabs(x); break; }
    }
}
/* same, Xrange operands (numba specialises the closures on XF) */
static inline void bs_jac(int flavor, XF x, XF y, XF *fxx, XF *fxy, XF *fyx, XF *fyy)
{
    switch (flavor) {
    case 1: *fxx = 2. * x; *fxy = -2. * y; *fyx = 2. * sgn_(x) * fabs_(y); *fyy = 2. * sgn_(y) * fabs_(x); break;
    case 2: *fxx = 2. * x; *fxy = -2. * y; *fyx = 2. * fabs_(y); *fyy = 2. * sgn_(y) * x; break;
    case 3: *fxx = 2. * x; *fxy = -2. * fabs_(y); *fyx = 2. * y; *fyy = 2. * x; break;
    case 4: { double s = sgn_(x * x - y * y); *fxx = 2. * s * x; *fxy = -2. * s * y; *fyx = 2. * y; *fyy = 2. * x; break; }
    default: { double s = sgn_(x * x - y * y); *fxx = 2. * s * x; *fxy = -2. * s * y; *fyx = 2. * sgn_(x) * fabs_(y); *fyy = 2. * sgn_(y) * fabs_(x); break; }
    }
}

/* perturbation.py:1793-1811 */
template <class T>
static inline void apply_bla_bs(const double *M, T &x, T &y, T a, T b)
{
    T nx = M[0] * x + M[1] * y + M[4] * a + M[5] * b;
    T ny = M[2] * x + M[3] * y + M[6] * a + M[7] * b;
    x = nx; y = ny;
}
template <class T>
static inline void apply_bla_deriv_bs(const double *M, T &dxa, T &dxb, T &dya, T &dyb)
{
    T a = M[0] * dxa + M[1] * dya;
    T b = M[0] * dxb + M[1] * dyb;
    T c = M[2] * dxa + M[3] * dya;
    T d = M[2] * dxb + M[3] * dyb;
    dxa = a; dxb = b; dya = c; dyb = d;
}

/* perturbation.py:1468-1791 */
template <bool XR>
static void perturb_bs_pixel(const fso_frame_bs *f, C pix, double *Z,
                             int64_t stride, int32_t *U, int8_t *stop_out,
                             int32_t *niter_out, int64_t *cnt)
{
    const int64_t L = f->L;
    const bool has_xr = f->n_xr > 0;
    const bool hess = f->calc_hessian;
    const int flavor = f->flavor;
    const int64_t ref_order = f->ref_order, ref_div_iter = f->ref_div_iter;

    /* perturbation.py:2260-2280 */
    XC c_xr;
    {
        double x1 = f->lin_mat[0] * pix.re + f->lin_mat[1] * pix.im;
        double y1 = f->lin_mat[2] * pix.re + f->lin_mat[3] * pix.im;
        c_xr = mkXF(f->lin_scale, f->lin_scale_e) * mkC(x1, y1);
    }
    XF a_x = mkXF(c_xr.m.re, c_xr.e) + mkXF(f->driftx, f->driftx_e);
    XF b_x = mkXF(c_xr.m.im, c_xr.e) + mkXF(f->drifty, f->drifty_e);
    double a = to_std(a_x), b = to_std(b_x);

    double x = 0., y = 0., dxa = 0., dxb = 0., dya = 0., dyb = 0.;
    XF x_x = to_xr(0.), y_x = x_x, dxa_x = x_x, dxb_x = x_x, dya_x = x_x, dyb_x = x_x;
    const XF record_zero = mkXF(0., 0);

    int64_t w_iter = 0, n_iter = 0;
    if (w_iter >= ref_order) w_iter = w_iter % ref_order;
    int64_t div_shift = 0, orbit_i1 = 0, orbit_i2 = 0;
    double oxn1 = 0., oxn2 = 0., oyn1 = 0., oyn2 = 0.;

    int64_t first_invalid = L;
    if (ref_div_iter < first_invalid) first_invalid = ref_div_iter;
    if (ref_order < first_invalid) first_invalid = ref_order;
    bool bool_dyn_rebase = true;
    int8_t stop = -1;

#define D_X(arr, i) mkXF(f->arr[(i)], f->arr##_e[(i)])

    for (;;) {
        if (f->bla_activated && (w_iter & 7) == 0) {
            int64_t ib = 0;
            int64_t step = ref_bla_get(f->r_bla, f->stages_bla, mkC(x, y), w_iter,
                                       first_invalid, &ib);
            if (step != 0) {
                const double *M = f->M_bla + 8 * ib;
                n_iter += step;
                w_iter = (w_iter + step) % ref_order;
                if (XR) {
                    apply_bla_bs(M, x_x, y_x, a_x, b_x);
                    x = to_std(x_x);
                    y = to_std(y_x);
                    if (hess) apply_bla_deriv_bs(M, dxa_x, dxb_x, dya_x, dyb_x);
                } else {
                    apply_bla_bs(M, x, y, a, b);
                    if (hess) apply_bla_deriv_bs(M, dxa, dxb, dya, dyb);
                }
                cnt[1] += 1;
                continue;
            }
        }

        n_iter += 1;
        cnt[0] += 1;
        C ref_zn = path_c(f->Zn_path, w_iter);
        XF rx_x = record_zero, ry_x = record_zero;
        if (XR) {
            int64_t k = has_xr ? xr_find(f->ref_index_xr, f->n_xr, w_iter) : -1;
            if (k >= 0) {
                rx_x = mkXF(f->refx_xr[k], f->refx_xr_e[k]);
                ry_x = mkXF(f->refy_xr[k], f->refy_xr_e[k]);
            } else {
                rx_x = to_xr(ref_zn.re);
                ry_x = to_xr(ref_zn.im);
            }
        }

        if (hess) {
            if (XR) {
                XF ra = record_zero, rb = record_zero, rc = record_zero, rd = record_zero;
                if (!bool_dyn_rebase) {
                    ra = D_X(dXnda, w_iter); rb = D_X(dXndb, w_iter);
                    rc = D_X(dYnda, w_iter); rd = D_X(dYndb, w_iter);
                }
                bs_p_iter_hessian(flavor, x_x, y_x, dxa_x, dxb_x, dya_x, dyb_x,
                                  rx_x, ry_x, ra, rb, rc, rd);
            } else {
                double ra = 0., rb = 0., rc = 0., rd = 0.;
                if (!bool_dyn_rebase) {
                    ra = f->dXnda[w_iter]; rb = f->dXndb[w_iter];
                    rc = f->dYnda[w_iter]; rd = f->dYndb[w_iter];
                }
                bs_p_iter_hessian(flavor, x, y, dxa, dxb, dya, dyb, ref_zn.re,
                                  ref_zn.im, ra, rb, rc, rd);
            }
        }
        if (XR) {
            bs_p_iter_zn(flavor, x_x, y_x, rx_x, ry_x, a_x, b_x);
            x = to_std(x_x);
            y = to_std(y_x);
        } else {
            bs_p_iter_zn(flavor, x, y, ref_zn.re, ref_zn.im, a, b);
        }

        /* max_iter BEFORE w_iter += 1 (:1616-1625) */
        if (n_iter >= f->max_iter) { stop = 0; break; }

        w_iter += 1;
        if (w_iter >= ref_order) w_iter = w_iter % ref_order;

        C ref_next = path_c(f->Zn_path, w_iter);
        int64_t knext = -1;
        if (XR && has_xr) knext = xr_find(f->ref_index_xr, f->n_xr, w_iter);
        double XX = x + ref_next.re, YY = y + ref_next.im;
        double full_sq_norm = XX * XX + YY * YY;
        if (f->calc_orbit) {
            int64_t div = n_iter / f->backshift;
            if (div > div_shift) {
                div_shift = div;
                orbit_i2 = orbit_i1; oxn2 = oxn1; oyn2 = oyn1;
                orbit_i1 = n_iter; oxn1 = XX; oyn1 = YY;
            }
        }
        if (full_sq_norm > f->M_divergence_sq) { stop = 1; break; }

        if (w_iter >= ref_div_iter - 1) {
            x = XX; y = YY;
            if (XR) {
                x_x = to_xr(XX); y_x = to_xr(YY);
                if (hess) {
                    dxa_x = dxa_x + D_X(dXnda, w_iter); dxb_x = dxb_x + D_X(dXndb, w_iter);
                    dya_x = dya_x + D_X(dYnda, w_iter); dyb_x = dyb_x + D_X(dYndb, w_iter);
                }
            } else if (hess) {
                dxa += f->dXnda[w_iter]; dxb += f->dXndb[w_iter];
                dya += f->dYnda[w_iter]; dyb += f->dYndb[w_iter];
            }
            w_iter = 0;
            cnt[2] += 1;
            continue;
        }

        bool_dyn_rebase = (fabs(XX) <= fabs(x)) && (fabs(YY) <= fabs(y));
        if (bool_dyn_rebase) {
            if (XR) {
                XF Xn = x_x, Yn = y_x, XXx, YYx;
                if (knext >= 0) {
                    XXx = Xn + mkXF(f->refx_xr[knext], f->refx_xr_e[knext]);
                    YYx = Yn + mkXF(f->refy_xr[knext], f->refy_xr_e[knext]);
                } else {
                    XXx = Xn + ref_next.re;
                    YYx = Yn + ref_next.im;
                }
                if (xr_le(XXx * XXx + YYx * YYx, Xn * Xn + Yn * Yn)) {
                    x_x = XXx; y_x = YYx;
                    x = to_std(XXx); y = to_std(YYx);
                    if (hess) {
                        dxa_x = dxa_x + D_X(dXnda, w_iter); dxb_x = dxb_x + D_X(dXndb, w_iter);
                        dya_x = dya_x + D_X(dYnda, w_iter); dyb_x = dyb_x + D_X(dYndb, w_iter);
                    }
                    w_iter = 0;
                    cnt[2] += 1;
                    continue;
                }
            } else {
                x = XX; y = YY;
                if (hess) {
                    dxa += f->dXnda[w_iter]; dxb += f->dXndb[w_iter];
                    dya += f->dYnda[w_iter]; dyb += f->dYndb[w_iter];
                }
                w_iter = 0;
                cnt[2] += 1;
                continue;
            }
        }
    }

    U[0] = (int32_t)w_iter;
    C ref_zn = path_c(f->Zn_path, w_iter);
    if (XR) {
        x = to_std(x_x + ref_zn.re);
        y = to_std(y_x + ref_zn.im);
        if (hess) {
            dxa = to_std(dxa_x + D_X(dXnda, w_iter)); dxb = to_std(dxb_x + D_X(dXndb, w_iter));
            dya = to_std(dya_x + D_X(dYnda, w_iter)); dyb = to_std(dyb_x + D_X(dYndb, w_iter));
        }
    } else {
        x += ref_zn.re; y += ref_zn.im;
        if (hess) {
            dxa += f->dXnda[w_iter]; dxb += f->dXndb[w_iter];
            dya += f->dYnda[w_iter]; dyb += f->dYndb[w_iter];
        }
    }
#undef D_X
    int row = 0;
    Z[stride * row++] = x; Z[stride * row++] = y;
    if (hess) { Z[stride * row++] = dxa; Z[stride * row++] = dxb; Z[stride * row++] = dya; Z[stride * row++] = dyb; }
    if (f->calc_orbit) {
        double xo = oxn2, yo = oyn2;
        double AA = a + f->Zn_path[2], BB = b + f->Zn_path[3];
        while (orbit_i2 < n_iter - f->backshift) {
            double tx, ty;
            bs_iterate(flavor, xo, yo, AA, BB, &tx, &ty);
            xo = tx; yo = ty; orbit_i2 += 1;
        }
        Z[stride * row++] = xo; Z[stride * row++] = yo;
    }
    *stop_out = stop;
    *niter_out = (int32_t)n_iter;
}

static int stages_bla_of(int64_t L)
{
    /* perturbation.py:1976-1981 : int(ceil(log2(L))) */
    int s = 0;
    while (((int64_t)1 << s) < L) s++;
    return s;
}

/* perturbation.py:1983-2024 */
static void combine_bla(C *M, double *r, double kc_std, int stg, int64_t len, double eps)
{
    int64_t step = (int64_t)1 << stg;
    for (int64_t i = 0; i < len - step + 1; i += step) {
        int64_t ii = i + step / 2;
        if (ii >= len) break;
        int64_t i1 = bla_index(i, stg - 1), i2 = bla_index(ii, stg - 1), ir = bla_index(i, stg);
        C A1 = M[2 * i1], B1 = M[2 * i1 + 1], A2 = M[2 * i2], B2 = M[2 * i2 + 1];
        M[2 * ir] = A2 * A1;
        M[2 * ir + 1] = A2 * B1 + B2;
        double r1 = r[i1], r2 = r[i2];
        double mA1 = cabs_(A1), mB1 = cabs_(B1);
        double r2_backw = 0.95 * pymax(0., (r2 - mB1 * kc_std) / pymax(mA1, eps));
        r[ir] = pymin(r1, r2_backw);
    }
}

/* perturbation.py:2027-2105 */
static void combine_bla_bs(double *M, double *r, double kc_std, int stg, int64_t len, double eps)
{
    int64_t step = (int64_t)1 << stg;
    for (int64_t i = 0; i < len - step + 1; i += step) {
        int64_t ii = i + step / 2;
        if (ii >= len) break;
        int64_t i1 = bla_index(i, stg - 1), i2 = bla_index(ii, stg - 1), ir = bla_index(i, stg);
        const double *M1 = M + 8 * i1, *M2 = M + 8 * i2;
        double R[8];
        R[0] = M2[0] * M1[0] + M2[1] * M1[2];
        R[1] = M2[0] * M1[1] + M2[1] * M1[3];
        R[2] = M2[2] * M1[0] + M2[3] * M1[2];
        R[3] = M2[2] * M1[1] + M2[3] * M1[3];
        R[4] = M2[0] * M1[4] + M2[1] * M1[6] + M2[4];
        R[5] = M2[0] * M1[5] + M2[1] * M1[7] + M2[5];
        R[6] = M2[2] * M1[4] + M2[3] * M1[6] + M2[6];
        R[7] = M2[2] * M1[5] + M2[3] * M1[7] + M2[7];
        double r1 = r[i1], r2 = r[i2];
        double mA1 = pymax(pymax(pymax(fabs(M1[0]), fabs(M1[1])), fabs(M1[2])), fabs(M1[3]));
        double mB1 = pymax(pymax(pymax(fabs(M1[4]), fabs(M1[5])), fabs(M1[6])), fabs(M1[7]));
        double r2_backw = 0.95 * pymax(0., (r2 - mB1 * kc_std) / pymax(mA1, eps));
        for (int d = 0; d < 8; d++) M[8 * ir + d] = R[d];
        r[ir] = pymin(r1, r2_backw);
    }
}

static int resolve_threads(int nthreads)
{
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    return nthreads;
}

} /* namespace */

/* ======================================================================== */
extern "C" {

int fso_std_m2(int64_t npts, const double *c_pix, double center_re,
               double center_im, double dx, const double *lin_mat,
               int64_t max_iter, double Mdiv_sq, double epscv_sq,
               int calc_d2zndc2, int calc_orbit, int64_t backshift, double *Z,
               int8_t *stop_reason, int32_t *stop_iter, int nthreads)
{
    nthreads = resolve_threads(nthreads);
    C center = mkC(center_re, center_im);
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
    for (int64_t i = 0; i < npts; i++) {
        C c = c_from_pix(path_c(c_pix, i), lin_mat, dx, center);
        std_m2_pixel(c, max_iter, Mdiv_sq, epscv_sq, calc_d2zndc2, calc_orbit,
                     backshift, Z + 2 * i, npts, stop_reason + i, stop_iter + i);
    }
    return 0;
}

int fso_std_mn(int nexp, int use_cpow, int64_t npts, const double *c_pix, double center_re,
               double center_im, double dx, const double *lin_mat,
               int64_t max_iter, double Mdiv_sq, double epscv_sq,
               int calc_d2zndc2, int calc_orbit, int64_t backshift, double *Z,
               int8_t *stop_reason, int32_t *stop_iter, int nthreads)
{
    C center = mkC(center_re, center_im);
    int nt = resolve_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt)
    for (int64_t i = 0; i < npts; i++) {
        C c = c_from_pix(path_c(c_pix, i), lin_mat, dx, center);
        std_mn_pixel(nexp, use_cpow, c, max_iter, Mdiv_sq, epscv_sq, calc_d2zndc2, calc_orbit,
                     backshift, Z + 2 * i, npts, stop_reason + i, stop_iter + i);
    }
    return 0;
}

int fso_std_bs(int flavor, int64_t npts, const double *c_pix, double center_re,
               double center_im, double dx, const double *lin_mat,
               int64_t max_iter, double Mdiv_sq, int calc_orbit,
               int64_t backshift, double *Z, int8_t *stop_reason,
               int32_t *stop_iter, int nthreads)
{
    nthreads = resolve_threads(nthreads);
    C center = mkC(center_re, center_im);
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads)
    for (int64_t i = 0; i < npts; i++) {
        C c = c_from_pix(path_c(c_pix, i), lin_mat, dx, center);
        std_bs_pixel(flavor, c, max_iter, Mdiv_sq, calc_orbit, backshift, Z + i,
                     npts, stop_reason + i, stop_iter + i);
    }
    return 0;
}

int fso_perturb_m2(const fso_frame_m2 *f, int64_t npts, const double *c_pix,
                   double *Z, int32_t *U, int8_t *stop_reason,
                   int32_t *stop_iter, int nthreads, int64_t *counters)
{
    nthreads = resolve_threads(nthreads);
    int64_t c0 = 0, c1 = 0, c2 = 0;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads) reduction(+ : c0, c1, c2)
    for (int64_t i = 0; i < npts; i++) {
        int64_t cnt[3] = {0, 0, 0};
        if (f->xr_detect)
            perturb_m2_pixel<true>(f, path_c(c_pix, i), Z + 2 * i, npts, U + i,
                                   stop_reason + i, stop_iter + i, cnt);
        else
            perturb_m2_pixel<false>(f, path_c(c_pix, i), Z + 2 * i, npts, U + i,
                                    stop_reason + i, stop_iter + i, cnt);
        c0 += cnt[0]; c1 += cnt[1]; c2 += cnt[2];
    }
    if (counters) { counters[0] = c0; counters[1] = c1; counters[2] = c2; }
    return 0;
}

int fso_perturb_bs(const fso_frame_bs *f, int64_t npts, const double *c_pix,
                   double *Z, int32_t *U, int8_t *stop_reason,
                   int32_t *stop_iter, int nthreads, int64_t *counters)
{
    nthreads = resolve_threads(nthreads);
    int64_t c0 = 0, c1 = 0, c2 = 0;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads) reduction(+ : c0, c1, c2)
    for (int64_t i = 0; i < npts; i++) {
        int64_t cnt[3] = {0, 0, 0};
        if (f->xr_detect)
            perturb_bs_pixel<true>(f, path_c(c_pix, i), Z + i, npts, U + i,
                                   stop_reason + i, stop_iter + i, cnt);
        else
            perturb_bs_pixel<false>(f, path_c(c_pix, i), Z + i, npts, U + i,
                                    stop_reason + i, stop_iter + i, cnt);
        c0 += cnt[0]; c1 += cnt[1]; c2 += cnt[2];
    }
    if (counters) { counters[0] = c0; counters[1] = c1; counters[2] = c2; }
    return 0;
}

/* perturbation.py:1819-1881 */
int fso_make_bla_m2(const double *Zn_path, int64_t L, double kc_m,
                    int32_t kc_e, double eps, double *M_bla, double *r_bla)
{
    return fso_make_bla_mn(0, Zn_path, L, kc_m, kc_e, eps, M_bla, r_bla);
}

/* nexp = 0: Perturbation_mandelbrot (dfdz = 2 z); N >= 2: ..._mandelbrot_N
 * (dfdz = N z^(N-1) by repeated products, mandelbrot_Mn.py:643-649) */
int fso_make_bla_mn(int nexp, const double *Zn_path, int64_t L, double kc_m,
                    int32_t kc_e, double eps, double *M_bla, double *r_bla)
{
    double kc_std = to_std(mkXF(kc_m, kc_e));
    const int k_comp = 8;
    int64_t comp_len = L / k_comp;
    C *M = (C *)M_bla;
    memset(M_bla, 0, sizeof(double) * 2 * 2 * 2 * comp_len);
    memset(r_bla, 0, sizeof(double) * 2 * comp_len);
    for (int64_t i = 0; i < comp_len; i++) {
        C tM[2 * 2 * 8];
        double tr[2 * 8];
        memset(tM, 0, sizeof(tM));
        memset(tr, 0, sizeof(tr));
        for (int j = 0; j < k_comp; j++) {
            C Zn_i = path_c(Zn_path, i * k_comp + j);
            tM[2 * (2 * j)] = nexp ? mn_dfdz(nexp, Zn_i)
                                   : 2. * Zn_i; /* dfdz, mandelbrot_M2.py:599-602 */
            tM[2 * (2 * j) + 1] = mkC(1., 0.);
            tr[2 * j] = eps * cabs_(tM[2 * (2 * j)]);
        }
        for (int stg = 1; stg <= 3; stg++) combine_bla(tM, tr, kc_std, stg, k_comp, eps);
        M[2 * (2 * i)] = tM[2 * 7];
        M[2 * (2 * i) + 1] = tM[2 * 7 + 1];
        r_bla[2 * i] = tr[7];
    }
    int stages = stages_bla_of(L);
    for (int stg = 1; stg < stages - 3; stg++) combine_bla(M, r_bla, kc_std, stg, comp_len, eps);
    return stages;
}

/* perturbation.py:1884-1973 */
int fso_make_bla_bs(int flavor, const double *Zn_path, int64_t L, double kc_m,
                    int32_t kc_e, double eps, double *M_bla, double *r_bla)
{
    double kc_std = to_std(mkXF(kc_m, kc_e));
    const int k_comp = 8;
    int64_t comp_len = L / k_comp;
    memset(M_bla, 0, sizeof(double) * 8 * 2 * comp_len);
    memset(r_bla, 0, sizeof(double) * 2 * comp_len);
    for (int64_t i = 0; i < comp_len; i++) {
        double tM[8 * 2 * 8], tr[2 * 8];
        memset(tM, 0, sizeof(tM));
        memset(tr, 0, sizeof(tr));
        for (int j = 0; j < k_comp; j++) {
            double X = Zn_path[2 * (i * k_comp + j)], Y = Zn_path[2 * (i * k_comp + j) + 1];
            double *m = tM + 8 * (2 * j);
            bs_jac(flavor, X, Y, &m[0], &m[1], &m[2], &m[3]);
            m[4] = 1.; m[5] = 0.; m[6] = 0.; m[7] = -1.;
            tr[2 * j] = eps * pymin(fabs(X), fabs(Y));
        }
        for (int stg = 1; stg <= 3; stg++) combine_bla_bs(tM, tr, kc_std, stg, k_comp, eps);
        for (int d = 0; d < 8; d++) M_bla[8 * (2 * i) + d] = tM[8 * 7 + d];
        r_bla[2 * i] = tr[7];
    }
    int stages = stages_bla_of(L);
    for (int stg = 1; stg < stages - 3; stg++) combine_bla_bs(M_bla, r_bla, kc_std, stg, comp_len, eps);
    return stages;
}

/* perturbation.py:2282-2336 */
int fso_dzndc_path_m2(const double *Zn_path, int64_t L, int64_t n_xr,
                      const int32_t *ref_index_xr, const double *ref_xr,
                      const int32_t *ref_xr_e, int64_t ref_div_iter,
                      int64_t ref_order, double scale_m, int32_t scale_e,
                      int xr_detect, double *out, int32_t *out_e)
{
    return fso_dzndc_path_mn(0, Zn_path, L, n_xr, ref_index_xr, ref_xr, ref_xr_e,
                             ref_div_iter, ref_order, scale_m, scale_e, xr_detect, out, out_e);
}

int fso_dzndc_path_mn(int nexp, const double *Zn_path, int64_t L, int64_t n_xr,
                      const int32_t *ref_index_xr, const double *ref_xr,
                      const int32_t *ref_xr_e, int64_t ref_div_iter,
                      int64_t ref_order, double scale_m, int32_t scale_e,
                      int xr_detect, double *out, int32_t *out_e)
{
    int64_t valid_pts = L < ref_div_iter ? L : ref_div_iter;
    XF scale_x = mkXF(scale_m, scale_e);
    double scale = to_std(scale_x);
    C *o = (C *)out;
    for (int64_t i = 0; i < L; i++) { o[i] = mkC(0., 0.); if (xr_detect) out_e[i] = 0; }
    if (valid_pts < 2) return 0;
    int64_t i = 1;
    if (xr_detect) {
        for (i = 1; i < valid_pts; i++) {
            C ref_zn = path_c(Zn_path, i - 1);
            int64_t k = n_xr > 0 ? xr_find(ref_index_xr, n_xr, i - 1) : -1;
            XC rz = (k >= 0) ? mkXC(path_c(ref_xr, k), ref_xr_e[k]) : to_xr(ref_zn);
            XC v = dfdz_(nexp, rz) * mkXC(o[i - 1], out_e[i - 1]) + scale_x;
            o[i] = v.m; out_e[i] = v.e;
        }
        i = valid_pts - 1;
        if (i == ref_order - 1) {
            XC v = dfdz_(nexp, path_c(Zn_path, i)) * mkXC(o[i], out_e[i]) + scale_x;
            o[0] = v.m; out_e[0] = v.e;
        }
    } else {
        for (i = 1; i < valid_pts; i++)
            o[i] = dfdz_(nexp, path_c(Zn_path, i - 1)) * o[i - 1] + scale;
        i = valid_pts - 1;
        if (i == ref_order - 1)
            o[0] = dfdz_(nexp, path_c(Zn_path, i)) * o[i] + scale;
    }
    return 0;
}

/* perturbation.py:2466-2516 */
int fso_dzndz_path_m2(const double *Zn_path, int64_t L, int64_t n_xr,
                      const int32_t *ref_index_xr, const double *ref_xr,
                      const int32_t *ref_xr_e, int64_t ref_div_iter,
                      int64_t ref_order, int xr_detect, double *out,
                      int32_t *out_e)
{
    return fso_dzndz_path_mn(0, Zn_path, L, n_xr, ref_index_xr, ref_xr, ref_xr_e,
                             ref_div_iter, ref_order, xr_detect, out, out_e);
}

int fso_dzndz_path_mn(int nexp, const double *Zn_path, int64_t L, int64_t n_xr,
                      const int32_t *ref_index_xr, const double *ref_xr,
                      const int32_t *ref_xr_e, int64_t ref_div_iter,
                      int64_t ref_order, int xr_detect, double *out,
                      int32_t *out_e)
{
    (void)ref_order;
    int64_t valid_pts = L < ref_div_iter ? L : ref_div_iter;
    C *o = (C *)out;
    for (int64_t i = 0; i < L + 1; i++) { o[i] = mkC(0., 0.); if (xr_detect) out_e[i] = 0; }
    o[1] = mkC(1., 0.);
    if (valid_pts < 3) return 0;
    int64_t i;
    if (xr_detect) {
        for (i = 2; i < valid_pts; i++) {
            C ref_zn = path_c(Zn_path, i - 1);
            int64_t k = n_xr > 0 ? xr_find(ref_index_xr, n_xr, i - 1) : -1;
            XC rz = (k >= 0) ? mkXC(path_c(ref_xr, k), ref_xr_e[k]) : to_xr(ref_zn);
            XC v = dfdz_(nexp, rz) * mkXC(o[i - 1], out_e[i - 1]);
            o[i] = v.m; out_e[i] = v.e;
        }
        i = valid_pts - 1;
        C ref_zn = path_c(Zn_path, i);
        int64_t k = n_xr > 0 ? xr_find(ref_index_xr, n_xr, i) : -1;
        XC rz = (k >= 0) ? mkXC(path_c(ref_xr, k), ref_xr_e[k]) : to_xr(ref_zn);
        XC v = dfdz_(nexp, rz) * mkXC(o[i], out_e[i]);
        o[L] = v.m; out_e[L] = v.e;
    } else {
        for (i = 2; i < valid_pts; i++)
            o[i] = dfdz_(nexp, path_c(Zn_path, i - 1)) * o[i - 1];
        i = valid_pts - 1;
        o[L] = dfdz_(nexp, path_c(Zn_path, i)) * o[i];
    }
    return 0;
}

/* perturbation.py:2339-2463 ; out4 = [dXnda | dXndb | dYnda | dYndb], L each */
int fso_dzndc_path_bs(int flavor, const double *Zn_path, int64_t L,
                      int64_t n_xr, const int32_t *ref_index_xr,
                      const double *refx_xr, const int32_t *refx_xr_e,
                      const double *refy_xr, const int32_t *refy_xr_e,
                      int64_t ref_div_iter, int64_t ref_order, double scale_m,
                      int32_t scale_e, int xr_detect, double *out4,
                      int32_t *out4_e)
{
    int64_t valid_pts = L < ref_div_iter ? L : ref_div_iter;
    XF scale_x = mkXF(scale_m, scale_e);
    double scale = to_std(scale_x);
    double *A = out4, *B = out4 + L, *Cc = out4 + 2 * L, *D = out4 + 3 * L;
    int32_t *Ae = out4_e, *Be = out4_e ? out4_e + L : 0, *Ce = out4_e ? out4_e + 2 * L : 0,
            *De = out4_e ? out4_e + 3 * L : 0;
    for (int64_t i = 0; i < 4 * L; i++) { out4[i] = 0.; if (xr_detect) out4_e[i] = 0; }
    if (valid_pts < 2) return 0;
    int64_t n_steps = valid_pts - 1;
    bool wrap = ((valid_pts - 1) == ref_order - 1);
    for (int64_t s = 0; s < n_steps + (wrap ? 1 : 0); s++) {
        int64_t from_i = (s < n_steps) ? s : valid_pts - 1;
        int64_t to_i = (s < n_steps) ? s + 1 : 0;
        double X = Zn_path[2 * from_i], Y = Zn_path[2 * from_i + 1];
        if (xr_detect) {
            int64_t k = n_xr > 0 ? xr_find(ref_index_xr, n_xr, from_i) : -1;
            XF rx = (k >= 0) ? mkXF(refx_xr[k], refx_xr_e[k]) : to_xr(X);
            XF ry = (k >= 0) ? mkXF(refy_xr[k], refy_xr_e[k]) : to_xr(Y);
            XF fxx, fxy, fyx, fyy;
            bs_jac(flavor, rx, ry, &fxx, &fxy, &fyx, &fyy);
            XF a = mkXF(A[from_i], Ae[from_i]), b = mkXF(B[from_i], Be[from_i]);
            XF c = mkXF(Cc[from_i], Ce[from_i]), d = mkXF(D[from_i], De[from_i]);
            XF na = fxx * a + fxy * c + scale_x;
            XF nb = fxx * b + fxy * d;
            XF nc = fyx * a + fyy * c;
            XF nd = fyx * b + fyy * d - scale_x;
            A[to_i] = na.m; Ae[to_i] = na.e; B[to_i] = nb.m; Be[to_i] = nb.e;
            Cc[to_i] = nc.m; Ce[to_i] = nc.e; D[to_i] = nd.m; De[to_i] = nd.e;
        } else {
            double fxx, fxy, fyx, fyy;
            bs_jac(flavor, X, Y, &fxx, &fxy, &fyx, &fyy);
            double a = A[from_i], b = B[from_i], c = Cc[from_i], d = D[from_i];
            A[to_i] = fxx * a + fxy * c + scale;
            B[to_i] = fxx * b + fxy * d;
            Cc[to_i] = fyx * a + fyy * c;
            D[to_i] = fyx * b + fyy * d - scale;
        }
    }
    return 0;
}

/* ---- Xrange unit-test entry points ---- */
void fso_xr_binop_c(int op, int64_t n, const double *a, const int32_t *ae,
                    const double *b, const int32_t *be, double *out,
                    int32_t *oute)
{
    for (int64_t i = 0; i < n; i++) {
        XC x = mkXC(path_c(a, i), ae[i]), y = mkXC(path_c(b, i), be[i]), r;
        switch (op) {
        case 0: r = x + y; break;
        case 1: { C p, q; int32_t e; coexp_c(x.m, x.e, y.m, y.e, &p, &q, &e); r = mkXC(p - q, e); break; }
        case 2: r = x * y; break;
        default: r = x / y; break;
        }
        out[2 * i] = r.m.re; out[2 * i + 1] = r.m.im; oute[i] = r.e;
    }
}

void fso_xr_binop_f(int op, int64_t n, const double *a, const int32_t *ae,
                    const double *b, const int32_t *be, double *out,
                    int32_t *oute)
{
    for (int64_t i = 0; i < n; i++) {
        XF x = mkXF(a[i], ae[i]), y = mkXF(b[i], be[i]), r;
        switch (op) {
        case 0: r = x + y; break;
        case 1: r = x - y; break;
        case 2: r = x * y; break;
        default: r = x / y; break;
        }
        out[i] = r.m; oute[i] = r.e;
    }
}

void fso_xr_compare_f(int cmp, int64_t n, const double *a, const int32_t *ae,
                      const double *b, const int32_t *be, uint8_t *out)
{
    for (int64_t i = 0; i < n; i++) {
        double x, y; int32_t e;
        coexp_f(a[i], ae[i], b[i], be[i], &x, &y, &e);
        bool r;
        switch (cmp) {
        case 0: r = x < y; break;
        case 1: r = x <= y; break;
        case 2: r = x == y; break;
        case 3: r = x != y; break;
        case 4: r = x >= y; break;
        default: r = x > y; break;
        }
        out[i] = r ? 1 : 0;
    }
}

void fso_xr_to_standard_c(int64_t n, const double *a, const int32_t *ae, double *out)
{
    for (int64_t i = 0; i < n; i++) {
        C r = to_std(mkXC(path_c(a, i), ae[i]));
        out[2 * i] = r.re; out[2 * i + 1] = r.im;
    }
}

void fso_xr_to_standard_f(int64_t n, const double *a, const int32_t *ae, double *out)
{
    for (int64_t i = 0; i < n; i++) out[i] = to_std(mkXF(a[i], ae[i]));
}

void fso_xr_normalize_c(int64_t n, const double *a, const int32_t *ae,
                        double *out, int32_t *oute)
{
    for (int64_t i = 0; i < n; i++) {
        XC r = normalize(path_c(a, i), ae[i]);
        out[2 * i] = r.m.re; out[2 * i + 1] = r.m.im; oute[i] = r.e;
    }
}

void fso_xr_abs2_c(int64_t n, const double *a, const int32_t *ae, double *out,
                   int32_t *oute)
{
    for (int64_t i = 0; i < n; i++) {
        XF r = abs2(mkXC(path_c(a, i), ae[i]));
        out[i] = r.m; oute[i] = r.e;
    }
}


/* ======================================================================== */
/* Projections: projection.py:363-373 (Expmap.make_f_impl), :455-471
 * (Expmap.make_dzndc_modifier), :205-219 (Cartesian.make_dzndc_modifier) and
 * the final `Z[dzndc] *= proj_dzndc_modifier(c_pix)` of the perturbation loops
 * (perturbation.py:1387-1388, 1772-1776).
 *
 * The reference evaluates them with numba's complex exp (cmathimpl.exp_impl:
 * r = exp(x), c = cos(y), s = sin(y) -> (r c, r s)) on top of the C library.
 *   det = 0  the C library functions, i.e. what the reference executes on this
 *            host: pinned bit for bit by the strict fixtures;
 *   det = 1  exp / sin / cos by a fixed sequence of individually rounded
 *            operations (Cody-Waite reduction + the classic fdlibm kernels,
 *            error < 1 ulp): a platform-independent definition, which the CUDA
 *            library evaluates with the same sequence -> bit-exact GPU parity.
 * The two differ by at most 1 ulp per function value (checked in tests/). */

static const double LN2HI = 6.93147180369123816490e-01, LN2LO = 1.90821492927058770002e-10,
                    INVLN2 = 1.44269504088896338700e+00;

static double pow2i(int e)         /* exact 2^e, 0 below the subnormal range */
{
    return ldexp(1., e);
}

double fso_det_exp(double x)
{
    static const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03,
                        P3 = 6.61375632143793436117e-05, P4 = -1.65339022054652515390e-06,
                        P5 = 4.13813679705723846039e-08;
    if (x != x) return x;
    if (x > 7.09782712893383973096e+02) return INFINITY;
    if (x < -7.45133219101941108420e+02) return 0.;
    double ax = fabs(x), hi = x, lo = 0.;
    int k = 0;
    if (ax > 0.34657359027997264) {
        if (ax < 1.0397207708399179) {
            if (x < 0.) { k = -1; hi = x + LN2HI; lo = -LN2LO; }
            else { k = 1; hi = x - LN2HI; lo = LN2LO; }
        } else {
            double h = (x < 0.) ? -0.5 : 0.5;
            double kk = INVLN2 * x;
            k = (int)(kk + h);
            double t = (double)k;
            double th = t * LN2HI;
            hi = x - th;
            lo = t * LN2LO;
        }
        x = hi - lo;
    } else if (ax < 3.725290298461914e-09) {
        return 1. + x;
    }
    double t = x * x;
    double q = t * P5; q = P4 + q;
    q = t * q; q = P3 + q;
    q = t * q; q = P2 + q;
    q = t * q; q = P1 + q;
    double tq = t * q;
    double c = x - tq;
    double xc = x * c;
    if (k == 0) {
        double d = c - 2.;
        double r = xc / d;
        r = r - x;
        return 1. - r;
    }
    double d = 2. - c;
    double r = xc / d;
    r = lo - r;
    r = r - hi;
    double y = 1. - r;
    int k1 = k / 2;
    y = y * pow2i(k1);
    return y * pow2i(k - k1);
}

static int expfield_(double x) { return (int)((d2b(x) >> 52) & 0x7ff); }

static int det_rem_pio2(double x, double *y0, double *y1)
{
    static const double invpio2 = 6.36619772367581382433e-01,
        pio2_1 = 1.57079632673412561417e+00, pio2_1t = 6.07710050650619224932e-11,
        pio2_2 = 6.07710050630396597660e-11, pio2_2t = 2.02226624879595063154e-21,
        pio2_3 = 2.02226624871116645580e-21, pio2_3t = 8.47842766036889956997e-32;
    double ax = fabs(x);
    if (ax <= 0.78539816339744830962) { *y0 = x; *y1 = 0.; return 0; }
    double nn = ax * invpio2;
    int n = (int)(nn + 0.5);
    double fn = (double)n;
    double p = fn * pio2_1;
    double r = ax - p;
    double w = fn * pio2_1t;
    int j = expfield_(ax);
    double z0 = r - w;
    if (j - expfield_(z0) > 16) {
        double t = r;
        w = fn * pio2_2;
        r = t - w;
        double u = t - r; u = u - w;
        w = fn * pio2_2t; w = w - u;
        z0 = r - w;
        if (j - expfield_(z0) > 49) {
            t = r;
            w = fn * pio2_3;
            r = t - w;
            u = t - r; u = u - w;
            w = fn * pio2_3t; w = w - u;
            z0 = r - w;
        }
    }
    double z1 = r - z0; z1 = z1 - w;
    if (x < 0.) { *y0 = -z0; *y1 = -z1; return -n; }
    *y0 = z0; *y1 = z1;
    return n;
}

static double det_ksin(double x, double y)
{
    static const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
        S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
        S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    if (fabs(x) < 7.450580596923828e-09) return x;
    double z = x * x, v = z * x;
    double r = z * S6; r = S5 + r;
    r = z * r; r = S4 + r;
    r = z * r; r = S3 + r;
    r = z * r; r = S2 + r;
    double hy = 0.5 * y, vr = v * r;
    double a = hy - vr;
    a = z * a;
    a = a - y;
    double b = v * S1;
    a = a - b;
    return x - a;
}

static double det_kcos(double x, double y)
{
    static const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
        C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
        C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double ax = fabs(x);
    if (ax < 7.450580596923828e-09) return 1.;
    double z = x * x;
    double r = z * C6; r = C5 + r;
    r = z * r; r = C4 + r;
    r = z * r; r = C3 + r;
    r = z * r; r = C2 + r;
    r = z * r; r = C1 + r;
    r = z * r;
    double zr = z * r, xy = x * y;
    double e = zr - xy;
    double hz = 0.5 * z;
    if (ax < 0.3) { double a = hz - e; return 1. - a; }
    double qx;
    if (ax > 0.78125) qx = 0.28125;
    else qx = b2d((int64_t)((uint64_t)(((d2b(ax) >> 32) - 0x00200000) & 0xffffffffLL) << 32));
    hz = hz - qx;
    double a = 1. - qx;
    double b = hz - e;
    return a - b;
}

void fso_det_sincos(double x, double *s, double *c)
{
    if (!(fabs(x) < 1.6e6)) { *s = *c = NAN; return; }
    double y0, y1;
    int n = det_rem_pio2(x, &y0, &y1);
    double ks = det_ksin(y0, y1), kc = det_kcos(y0, y1);
    switch (n & 3) {
    case 0: *s = ks; *c = kc; break;
    case 1: *s = kc; *c = -ks; break;
    case 2: *s = -ks; *c = -kc; break;
    default: *s = -kc; *c = ks; break;
    }
}

void fso_proj_expmap(int64_t n, const double *pix, double hmoy, double k_re, double k_im,
                     int det, double *out)
{
    const C k = mkC(k_re, k_im);
    for (int64_t i = 0; i < n; i++) {
        C ht = k * path_c(pix, i);           /* pix_to_ht * pix (float factor cast to complex) */
        double x = hmoy + ht.re;             /* float + complex: (hmoy + re, 0 + im) */
        double y = 0. + ht.im;
        double r, s, c;
        if (det) { r = fso_det_exp(x); fso_det_sincos(y, &s, &c); }
        else { c = cos(y); s = sin(y); r = exp(x); }
        out[2 * i] = r * c;
        out[2 * i + 1] = r * s;
    }
}

void fso_modifier_expmap(int64_t n, const double *pix, double k_re, double k_im, double hshift,
                         int det, double *out)
{
    for (int64_t i = 0; i < n; i++) {
        double a = k_re * pix[2 * i], b = k_im * pix[2 * i + 1];
        double h = a - b;                    /* np.real(pix_to_ht * cpix) */
        h = h + hshift;
        out[i] = det ? fso_det_exp(h) : exp(h);
    }
}

void fso_modifier_seam(int64_t n, const double *pix, double seam, int det, double *out)
{
    for (int64_t i = 0; i < n; i++) {
        double re = pix[2 * i] + 1.e-6, im = pix[2 * i + 1] + 0.;
        double a = det ? fso_hypot(re, im) : hypot(re, im);    /* np.abs(complex) */
        out[i] = a * seam;
    }
}

/* complex128 row *= float64 : numba casts the float to complex first */
void fso_apply_modifier_c(int64_t n, double *row, const double *mod)
{
    for (int64_t i = 0; i < n; i++) {
        C z = path_c(row, i);
        C m = mkC(mod[i], 0.);
        C r = z * m;
        row[2 * i] = r.re; row[2 * i + 1] = r.im;
    }
}

void fso_apply_modifier_f(int64_t n, double *row, const double *mod)
{
    for (int64_t i = 0; i < n; i++) row[i] = row[i] * mod[i];
}

} /* extern "C" */
