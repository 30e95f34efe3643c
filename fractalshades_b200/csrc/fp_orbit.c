/*
 * fp_orbit.c -- full-precision reference orbit on the host (MPFR / MPC).
 *
 * The reference orbit of a perturbation frame is a serial arbitrary-precision
 * recurrence; it stays on the host and is uploaded once per frame.  This file
 * is the native replacement for the Cython extension of the reference
 *   src/fractalshades/mpmath_utils/FP_loop.pyx:237-421  (holomorphic, z^2+c)
 *   src/fractalshades/mpmath_utils/FP_loop.pyx:1343-1455,1828-1979
 *                                          (burning-ship family, 5 flavours)
 * with the same contract: fill a double orbit, stop at |z| > M, register the
 * orbit points that underflow a double ("Xrange" points) as (mantissa, exp).
 *
 * It issues the same library calls at the same precision as the reference
 * (mpc_sqr + mpc_add, MPC_RNDNN; mpfr_sqr/mul/sub/add/abs/mul_si for the
 * non-holomorphic flavours) so that the stored doubles are bit-identical to
 * a gmpy2-built reference.
 *
 * The image has the MPFR/MPC runtime libraries but not their headers, so the
 * handful of prototypes used are declared here by hand (x86-64 SysV layout
 * of mpfr 4.x: { long prec; int sign; long exp; limb* d }).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/fsb200_orbit.h"

typedef struct {
    long _mpfr_prec;
    int _mpfr_sign;
    long _mpfr_exp;
    void *_mpfr_d;
} fsb_mpfr_struct;
typedef fsb_mpfr_struct fsb_mpfr_t[1];
typedef struct {
    fsb_mpfr_t re;
    fsb_mpfr_t im;
} fsb_mpc_struct;
typedef fsb_mpc_struct fsb_mpc_t[1];

#define RNDN 0   /* MPFR_RNDN */
#define RNDNN 0  /* MPC_RNDNN = MPC_RND(MPFR_RNDN, MPFR_RNDN) */

extern void mpfr_init2(fsb_mpfr_struct *, long);
extern void mpfr_clear(fsb_mpfr_struct *);
extern int mpfr_set_str(fsb_mpfr_struct *, const char *, int, int);
extern int mpfr_set_si(fsb_mpfr_struct *, long, int);
extern double mpfr_get_d(const fsb_mpfr_struct *, int);
extern double mpfr_get_d_2exp(long *, const fsb_mpfr_struct *, int);
extern int mpfr_sqr(fsb_mpfr_struct *, const fsb_mpfr_struct *, int);
extern int mpfr_mul(fsb_mpfr_struct *, const fsb_mpfr_struct *, const fsb_mpfr_struct *, int);
extern int mpfr_add(fsb_mpfr_struct *, const fsb_mpfr_struct *, const fsb_mpfr_struct *, int);
extern int mpfr_sub(fsb_mpfr_struct *, const fsb_mpfr_struct *, const fsb_mpfr_struct *, int);
extern int mpfr_mul_si(fsb_mpfr_struct *, const fsb_mpfr_struct *, long, int);
/* mpfr_abs is a macro over mpfr_set4(rop, op, rnd, sign=+1) */
extern int mpfr_set4(fsb_mpfr_struct *, const fsb_mpfr_struct *, int, int);
#define fsb_mpfr_abs(r, o) mpfr_set4((r), (o), RNDN, 1)

extern void mpc_init2(fsb_mpc_struct *, long);
extern void mpc_clear(fsb_mpc_struct *);
extern int mpc_set_fr_fr(fsb_mpc_struct *, const fsb_mpfr_struct *, const fsb_mpfr_struct *, int);
extern int mpc_set_si_si(fsb_mpc_struct *, long, long, int);
extern int mpc_sqr(fsb_mpc_struct *, const fsb_mpc_struct *, int);
extern int mpc_add(fsb_mpc_struct *, const fsb_mpc_struct *, const fsb_mpc_struct *, int);
extern int mpc_pow_ui(fsb_mpc_struct *, const fsb_mpc_struct *, unsigned long, int);

extern int mpc_mul(fsb_mpc_struct *, const fsb_mpc_struct *, const fsb_mpc_struct *, int);
extern int mpc_mul_si(fsb_mpc_struct *, const fsb_mpc_struct *, long, int);
extern int mpc_mul_fr(fsb_mpc_struct *, const fsb_mpc_struct *, const fsb_mpfr_struct *, int);
extern int mpc_add_ui(fsb_mpc_struct *, const fsb_mpc_struct *, unsigned long, int);
extern int mpc_sub(fsb_mpc_struct *, const fsb_mpc_struct *, const fsb_mpc_struct *, int);
extern int mpc_div(fsb_mpc_struct *, const fsb_mpc_struct *, const fsb_mpc_struct *, int);
extern int mpc_abs(fsb_mpfr_struct *, const fsb_mpc_struct *, int);
extern void mpc_swap(fsb_mpc_struct *, fsb_mpc_struct *);
extern int mpfr_ui_div(fsb_mpfr_struct *, unsigned long, const fsb_mpfr_struct *, int);
extern int mpfr_cmp_d(const fsb_mpfr_struct *, double);
extern int mpfr_greaterequal_p(const fsb_mpfr_struct *, const fsb_mpfr_struct *);
extern int mpc_mul_ui(fsb_mpc_struct *, const fsb_mpc_struct *, unsigned long, int);
extern int mpfr_add_si(fsb_mpfr_struct *, const fsb_mpfr_struct *, long, int);
extern int mpfr_sgn(const fsb_mpfr_struct *);
extern int mpfr_hypot(fsb_mpfr_struct *, const fsb_mpfr_struct *, const fsb_mpfr_struct *, int);
extern char *mpfr_get_str(char *, long *, int, size_t, const fsb_mpfr_struct *, int);
extern void mpfr_free_str(char *);

/* Thresholds: fs.settings.newton_zoom_level / xrange_zoom_level
 * (reference settings.py:14,22; captured at FP_loop.pyx:141-143). */
static const double XR_TSHOLD = 1.e-300;

static int xr_push(fsb_orbit_xr *out, int64_t cap, int64_t *n, int64_t idx,
                   const fsb_mpfr_struct *x, const fsb_mpfr_struct *y)
{
    if (*n >= cap) return -1;
    long ex = 0, ey = 0;
    fsb_orbit_xr *o = &out[*n];
    o->index = idx;
    /* FP_loop.pyx:436-437: mpfr_get_d_2exp -> mantissa in [0.5, 1) */
    o->mx = mpfr_get_d_2exp(&ex, x, RNDN);
    o->my = mpfr_get_d_2exp(&ey, y, RNDN);
    o->ex = (int32_t)ex;
    o->ey = (int32_t)ey;
    *n += 1;
    return 0;
}

/* FP_loop.pyx:274-421 */
int64_t fsb_orbit_mandelbrot(double *orbit, int64_t max_iter, uint32_t exponent,
                             int need_xrange, double M, const char *seed_x,
                             const char *seed_y, int64_t prec_bits,
                             fsb_orbit_xr *xr_out, int64_t xr_cap,
                             int64_t *xr_count)
{
    fsb_mpc_t z, c, tmp;
    fsb_mpfr_t x_t, y_t;
    int64_t i = 0, nxr = 0;
    double abs_i = 0.;
    int overflow = 0;

    if (exponent < 2) return -2;
    mpc_init2(z, prec_bits);
    mpc_init2(c, prec_bits);
    mpc_init2(tmp, prec_bits);
    mpfr_init2(x_t, prec_bits);
    mpfr_init2(y_t, prec_bits);
    if (mpfr_set_str(x_t, seed_x, 10, RNDN) != 0 ||
        mpfr_set_str(y_t, seed_y, 10, RNDN) != 0) {
        i = -3;
        goto done;
    }
    mpc_set_fr_fr(c, x_t, y_t, RNDNN);
    mpc_set_si_si(z, 0, 0, RNDNN);
    orbit[0] = 0.;
    orbit[1] = 0.;

    for (i = 1; i <= max_iter; i++) {
        if (exponent == 2) {
            mpc_sqr(tmp, z, RNDNN);
        } else {
            mpc_pow_ui(tmp, z, exponent, RNDNN);
        }
        mpc_add(z, tmp, c, RNDNN);
        double x = mpfr_get_d(z->re, RNDN);
        double y = mpfr_get_d(z->im, RNDN);
        orbit[2 * i] = x;
        orbit[2 * i + 1] = y;
        abs_i = hypot(x, y);
        if (abs_i > M) break;
        if (need_xrange && abs_i < XR_TSHOLD) {
            if (xr_push(xr_out, xr_cap, &nxr, i, z->re, z->im) != 0) overflow = 1;
        }
    }
    /* never escaped: the first invalid index is max_iter + 1 */
    if (i > max_iter) i = max_iter + 1;
    if (overflow) i = -4;
done:
    if (xr_count) *xr_count = nxr;
    mpc_clear(z);
    mpc_clear(c);
    mpc_clear(tmp);
    mpfr_clear(x_t);
    mpfr_clear(y_t);
    return i;
}

/* One burning-ship-family step, FP_loop.pyx:1358-1455 (same call sequences). */
static void bs_step(int kind, fsb_mpfr_struct *xn, fsb_mpfr_struct *yn,
                    const fsb_mpfr_struct *a, const fsb_mpfr_struct *b,
                    fsb_mpfr_struct *xsq, fsb_mpfr_struct *ysq, fsb_mpfr_struct *xy)
{
    switch (kind) {
    case FSB_FLAVOR_BURNING_SHIP:
        mpfr_sqr(xsq, xn, RNDN);
        mpfr_sqr(ysq, yn, RNDN);
        mpfr_mul(xy, xn, yn, RNDN);
        mpfr_sub(xn, xsq, ysq, RNDN);
        mpfr_add(xn, xn, a, RNDN);
        fsb_mpfr_abs(xy, xy);
        mpfr_mul_si(xy, xy, 2, RNDN);
        mpfr_sub(yn, xy, b, RNDN);
        break;
    case FSB_FLAVOR_PERPENDICULAR_BS:
        mpfr_sqr(xsq, xn, RNDN);
        mpfr_sqr(ysq, yn, RNDN);
        fsb_mpfr_abs(xy, yn);
        mpfr_mul(xy, xn, xy, RNDN);
        mpfr_mul_si(xy, xy, 2, RNDN);
        mpfr_sub(xn, xsq, ysq, RNDN);
        mpfr_add(xn, xn, a, RNDN);
        mpfr_sub(yn, xy, b, RNDN);
        break;
    case FSB_FLAVOR_SHARK_FIN:
        mpfr_sqr(xsq, xn, RNDN);
        fsb_mpfr_abs(ysq, yn);
        mpfr_mul(ysq, ysq, yn, RNDN);
        mpfr_mul(xy, xn, yn, RNDN);
        mpfr_mul_si(xy, xy, 2, RNDN);
        mpfr_sub(xn, xsq, ysq, RNDN);
        mpfr_add(xn, xn, a, RNDN);
        mpfr_sub(yn, xy, b, RNDN);
        break;
    case FSB_FLAVOR_CELTIC:
        mpfr_sqr(xsq, xn, RNDN);
        mpfr_sqr(ysq, yn, RNDN);
        mpfr_mul(xy, xn, yn, RNDN);
        mpfr_mul_si(xy, xy, 2, RNDN);
        mpfr_sub(xn, xsq, ysq, RNDN);
        fsb_mpfr_abs(xn, xn);
        mpfr_add(xn, xn, a, RNDN);
        mpfr_sub(yn, xy, b, RNDN);
        break;
    default: /* FSB_FLAVOR_BUFFALO */
        mpfr_sqr(xsq, xn, RNDN);
        mpfr_sqr(ysq, yn, RNDN);
        mpfr_mul(xy, xn, yn, RNDN);
        mpfr_sub(xn, xsq, ysq, RNDN);
        fsb_mpfr_abs(xn, xn);
        mpfr_add(xn, xn, a, RNDN);
        fsb_mpfr_abs(xy, xy);
        mpfr_mul_si(xy, xy, 2, RNDN);
        mpfr_sub(yn, xy, b, RNDN);
        break;
    }
}

/* FP_loop.pyx:1828-1979 */
int64_t fsb_orbit_burning_ship(double *orbit, int64_t max_iter, int flavor,
                               int need_xrange, double M, const char *seed_x,
                               const char *seed_y, int64_t prec_bits,
                               fsb_orbit_xr *xr_out, int64_t xr_cap,
                               int64_t *xr_count)
{
    fsb_mpfr_t xn, yn, a, b, xsq, ysq, xy;
    int64_t i = 0, nxr = 0;
    double abs_i = 0.;
    int overflow = 0;

    if (flavor < FSB_FLAVOR_BURNING_SHIP || flavor > FSB_FLAVOR_BUFFALO) return -2;
    mpfr_init2(xn, prec_bits);
    mpfr_init2(yn, prec_bits);
    mpfr_init2(a, prec_bits);
    mpfr_init2(b, prec_bits);
    mpfr_init2(xsq, prec_bits);
    mpfr_init2(ysq, prec_bits);
    mpfr_init2(xy, prec_bits);
    if (mpfr_set_str(a, seed_x, 10, RNDN) != 0 ||
        mpfr_set_str(b, seed_y, 10, RNDN) != 0) {
        i = -3;
        goto done;
    }
    mpfr_set_si(xn, 0, RNDN);
    mpfr_set_si(yn, 0, RNDN);
    orbit[0] = 0.;
    orbit[1] = 0.;

    for (i = 1; i <= max_iter; i++) {
        bs_step(flavor, xn, yn, a, b, xsq, ysq, xy);
        double x = mpfr_get_d(xn, RNDN);
        double y = mpfr_get_d(yn, RNDN);
        orbit[2 * i] = x;
        orbit[2 * i + 1] = y;
        abs_i = hypot(x, y);
        if (abs_i > M) break;
        if (need_xrange && (fabs(x) < XR_TSHOLD || fabs(y) < XR_TSHOLD)) {
            if (xr_push(xr_out, xr_cap, &nxr, i, xn, yn) != 0) overflow = 1;
        }
    }
    if (i > max_iter) i = max_iter + 1;
    if (overflow) i = -4;
done:
    if (xr_count) *xr_count = nxr;
    mpfr_clear(xn);
    mpfr_clear(yn);
    mpfr_clear(a);
    mpfr_clear(b);
    mpfr_clear(xsq);
    mpfr_clear(ysq);
    mpfr_clear(xy);
    return i;
}


/* ======================================================================== */
/* Period (ball method) and nucleus (Newton) of the reference point           */

/* one step of z <- z^2 + c with its derivative dz/dc <- 2 z dz/dc + 1, in the
 * call order of iter_deriv_M2 / iter_M2 (FP_loop.pyx:158-190) */
static void m2_step_deriv(fsb_mpc_struct *z, fsb_mpc_struct *dz, const fsb_mpc_struct *c,
                          fsb_mpc_struct *tmp)
{
    mpc_mul(tmp, z, dz, RNDNN);
    mpc_mul_si(dz, tmp, 2, RNDNN);
    mpc_add_ui(dz, dz, 1, RNDNN);
    mpc_sqr(tmp, z, RNDNN);
    mpc_add(z, tmp, c, RNDNN);
}

/* FP_loop.pyx:605-758 : first i with |z_i / (dz_i/dc)| < px, -1 if none */
int64_t fsb_ball_method_mandelbrot(const char *seed_x, const char *seed_y, int64_t prec_bits,
                                   const char *seed_px, int64_t maxiter, double M_divergence)
{
    fsb_mpc_t c, z, dz, tmp, r;
    fsb_mpfr_t ar, x_t, y_t, pix, inv_pix;
    int64_t ret = -1, i;
    mpc_init2(c, prec_bits); mpc_init2(z, prec_bits); mpc_init2(dz, prec_bits);
    mpc_init2(tmp, prec_bits); mpc_init2(r, prec_bits);
    mpfr_init2(ar, 54); mpfr_init2(x_t, prec_bits); mpfr_init2(y_t, prec_bits);
    mpfr_init2(pix, prec_bits); mpfr_init2(inv_pix, prec_bits);
    if (mpfr_set_str(x_t, seed_x, 10, RNDN) != 0 || mpfr_set_str(y_t, seed_y, 10, RNDN) != 0 ||
        mpfr_set_str(pix, seed_px, 10, RNDN) != 0) {
        ret = -3;
        goto done;
    }
    mpc_set_fr_fr(c, x_t, y_t, RNDNN);
    mpfr_ui_div(inv_pix, 1, pix, RNDN);
    mpc_set_si_si(z, 0, 0, RNDNN);
    mpc_set_si_si(dz, 0, 0, RNDNN);
    for (i = 1; i <= maxiter; i++) {
        m2_step_deriv(z, dz, c, tmp);
        mpc_div(r, z, dz, RNDNN);
        mpc_mul_fr(r, r, inv_pix, RNDNN);
        if (hypot(mpfr_get_d(z->re, RNDN), mpfr_get_d(z->im, RNDN)) > M_divergence) break;
        mpc_abs(ar, r, RNDN);
        if (mpfr_cmp_d(ar, 1.) < 0) { ret = i; break; }
    }
done:
    mpc_clear(c); mpc_clear(z); mpc_clear(dz); mpc_clear(tmp); mpc_clear(r);
    mpfr_clear(ar); mpfr_clear(x_t); mpfr_clear(y_t); mpfr_clear(pix); mpfr_clear(inv_pix);
    return ret;
}

/* "<sign>0.<digits>e<exp>" with enough digits for an exact round trip */
static int put_decimal(char *out, int64_t cap, const fsb_mpfr_struct *v)
{
    long e = 0;
    char *d = mpfr_get_str(NULL, &e, 10, 0, v, RNDN);
    if (!d) return -1;
    const char *digits = d;
    int neg = (d[0] == '-');
    if (neg) digits++;
    size_t need = strlen(digits) + 40;
    int rc = -1;
    if ((int64_t)need <= cap) {
        char ebuf[32];
        int k = 0;
        if (neg) out[k++] = '-';
        out[k++] = '0'; out[k++] = '.';
        strcpy(out + k, digits);
        k += (int)strlen(digits);
        out[k++] = 'e';
        /* long to string */
        long ev = e; int m = 0, j;
        if (ev < 0) { out[k++] = '-'; ev = -ev; }
        do { ebuf[m++] = (char)('0' + ev % 10); ev /= 10; } while (ev > 0);
        for (j = m - 1; j >= 0; j--) out[k++] = ebuf[j];
        out[k] = 0;
        rc = 0;
    }
    mpfr_free_str(d);
    return rc;
}

/* FP_loop.pyx:900-1118 (divide by the roots of the divisors of `order`) and
 * :1159-1340 (any_nucleus != 0: plain Newton on z_order(c)).  Returns 1 when the
 * descent converged to a point whose |z_order| passes eps_valid, 0 otherwise. */
int fsb_find_nucleus_mandelbrot(const char *seed_x, const char *seed_y, int64_t prec_bits,
                                int64_t order, int64_t max_newton, const char *seed_eps_cv,
                                const char *seed_eps_valid, int any_nucleus, char *out_x,
                                char *out_y, int64_t out_cap)
{
    fsb_mpc_t c, zr, dzr, h, dh, f, df, t1, t2;
    fsb_mpfr_t x_t, y_t, abs_diff, eps;
    int cv = 0;
    int64_t i_newton, i;
    if (order < 1) return -2;
    mpc_init2(c, prec_bits); mpc_init2(zr, prec_bits); mpc_init2(dzr, prec_bits);
    mpc_init2(h, prec_bits); mpc_init2(dh, prec_bits); mpc_init2(f, prec_bits);
    mpc_init2(df, prec_bits); mpc_init2(t1, prec_bits); mpc_init2(t2, prec_bits);
    mpfr_init2(x_t, prec_bits); mpfr_init2(y_t, prec_bits);
    mpfr_init2(abs_diff, 54); mpfr_init2(eps, 54);
    if (mpfr_set_str(x_t, seed_x, 10, RNDN) != 0 || mpfr_set_str(y_t, seed_y, 10, RNDN) != 0 ||
        mpfr_set_str(eps, seed_eps_cv, 10, RNDN) != 0) {
        cv = -3;
        goto done;
    }
    mpc_set_fr_fr(c, x_t, y_t, RNDNN);
    mpfr_mul_si(eps, eps, 64, RNDN);
    for (i_newton = 0; i_newton < max_newton; i_newton++) {
        mpc_set_si_si(zr, 0, 0, RNDNN);
        mpc_set_si_si(dzr, 0, 0, RNDNN);
        if (any_nucleus) {
            for (i = 1; i <= order; i++) m2_step_deriv(zr, dzr, c, t1);
            mpc_div(t1, zr, dzr, RNDNN);                 /* Newton step f / f' */
            mpc_sub(c, c, t1, RNDNN);
        } else {
            mpc_set_si_si(h, 1, 0, RNDNN);
            mpc_set_si_si(dh, 0, 0, RNDNN);
            for (i = 1; i <= order; i++) {
                mpc_mul_si(t1, dzr, 2, RNDNN);
                mpc_mul(t2, t1, zr, RNDNN);
                mpc_add_ui(dzr, t2, 1, RNDNN);
                mpc_sqr(t1, zr, RNDNN);
                mpc_add(zr, t1, c, RNDNN);
                if (i < order && order % i == 0) {       /* h *= z_i ; dh += z_i' / z_i */
                    mpc_mul(t1, h, zr, RNDNN);
                    mpc_swap(t1, h);
                    mpc_div(t1, dzr, zr, RNDNN);
                    mpc_add(t2, t1, dh, RNDNN);
                    mpc_swap(t2, dh);
                }
            }
            mpc_div(f, zr, h, RNDNN);                    /* f = z / h */
            mpc_mul(t1, zr, dh, RNDNN);
            mpc_sub(t2, dzr, t1, RNDNN);
            mpc_div(df, t2, h, RNDNN);                   /* f' = (z' - z dh) / h */
            mpc_div(t1, f, df, RNDNN);
            mpc_sub(t2, c, t1, RNDNN);
            mpc_swap(t2, c);
        }
        mpc_abs(abs_diff, t1, RNDN);
        if (mpfr_greaterequal_p(eps, abs_diff)) {
            mpc_abs(abs_diff, zr, RNDN);
            if (mpfr_set_str(eps, seed_eps_valid, 10, RNDN) != 0) { cv = -3; goto done; }
            cv = mpfr_greaterequal_p(eps, abs_diff) ? 1 : 0;
            break;
        }
    }
    if (cv == 1 && (put_decimal(out_x, out_cap, c->re) != 0 || put_decimal(out_y, out_cap, c->im) != 0))
        cv = -4;
done:
    mpc_clear(c); mpc_clear(zr); mpc_clear(dzr); mpc_clear(h); mpc_clear(dh); mpc_clear(f);
    mpc_clear(df); mpc_clear(t1); mpc_clear(t2);
    mpfr_clear(x_t); mpfr_clear(y_t); mpfr_clear(abs_diff); mpfr_clear(eps);
    return cv;
}


/* ======================================================================== */
/* Period and nucleus of the reference point, burning-ship family           */

/* FP_loop.pyx:1778-1783: never 0 */
static int sign_pm(const fsb_mpfr_struct *op) { return mpfr_sgn(op) >= 0 ? 1 : -1; }

/* One step of the Jacobian d(xn, yn)/d(a, b) (var_ab_xy = 0) -- FP_loop.pyx:
 * iter_J_BS :1458-1512, iter_J_pBS :1514-1565, iter_J_sharkfin :1567-1617,
 * iter_J_celtic :1619-1675, iter_J_buffalo :1677-1743 -- the same MPFR calls in
 * the same order.  (xx, xy, yx, yy) = (dxnda, dxndb, dynda, dyndb), updated
 * in place from the CURRENT (xn, yn): called before the orbit step. */
static void bs_jacobian_step(int kind, const fsb_mpfr_struct *xn, const fsb_mpfr_struct *yn,
                             fsb_mpfr_struct *xx, fsb_mpfr_struct *xy, fsb_mpfr_struct *yx,
                             fsb_mpfr_struct *yy, fsb_mpfr_struct *abs_xn, fsb_mpfr_struct *abs_yn,
                             fsb_mpfr_struct *t_xx, fsb_mpfr_struct *t_xy, fsb_mpfr_struct *t_yx,
                             fsb_mpfr_struct *t_yy, fsb_mpfr_struct *tmp)
{
    int sgn_xn = 1, sgn_yn = 1, sgn_d = 1;
    const int first_abs = (kind == FSB_FLAVOR_CELTIC || kind == FSB_FLAVOR_BUFFALO);
    const int second_abs_xy = (kind == FSB_FLAVOR_BURNING_SHIP || kind == FSB_FLAVOR_BUFFALO);
    if (first_abs) {                       /* sign of x^2 - y^2 */
        mpfr_sqr(tmp, xn, RNDN);
        mpfr_sqr(t_xy, yn, RNDN);
        mpfr_sub(t_xy, tmp, t_xy, RNDN);
        sgn_d = sign_pm(t_xy);
    }
    if (second_abs_xy) {
        if (kind == FSB_FLAVOR_BURNING_SHIP) {
            fsb_mpfr_abs(abs_xn, xn); fsb_mpfr_abs(abs_yn, yn);
            sgn_xn = sign_pm(xn); sgn_yn = sign_pm(yn);
        } else {
            sgn_xn = sign_pm(xn); sgn_yn = sign_pm(yn);
            fsb_mpfr_abs(abs_xn, xn); fsb_mpfr_abs(abs_yn, yn);
        }
    } else if (kind == FSB_FLAVOR_PERPENDICULAR_BS) {
        fsb_mpfr_abs(abs_yn, yn);
        sgn_yn = sign_pm(yn);
    } else if (kind == FSB_FLAVOR_SHARK_FIN) {
        fsb_mpfr_abs(abs_yn, yn);
    }
    /* first row: 2 [sgn] (xn dx - Y dy), Y = yn (|yn| for the shark fin) */
    {
        const fsb_mpfr_struct *Y = (kind == FSB_FLAVOR_SHARK_FIN) ? abs_yn : yn;
        mpfr_mul(t_xx, xn, xx, RNDN);
        mpfr_mul(tmp, Y, yx, RNDN);
        mpfr_sub(t_xx, t_xx, tmp, RNDN);
        if (first_abs) mpfr_mul_si(t_xx, t_xx, sgn_d, RNDN);
        mpfr_mul(t_xy, xn, xy, RNDN);
        mpfr_mul(tmp, Y, yy, RNDN);
        mpfr_sub(t_xy, t_xy, tmp, RNDN);
        if (first_abs) mpfr_mul_si(t_xy, t_xy, sgn_d, RNDN);
    }
    /* second row */
    if (second_abs_xy) {                   /* 2 (|xn| sgn_yn dy + sgn_xn dx |yn|) */
        mpfr_mul(t_yx, abs_xn, yx, RNDN);
        mpfr_mul_si(t_yx, t_yx, sgn_yn, RNDN);
        mpfr_mul(tmp, abs_yn, xx, RNDN);
        mpfr_mul_si(tmp, tmp, sgn_xn, RNDN);
        mpfr_add(t_yx, t_yx, tmp, RNDN);
        mpfr_mul(t_yy, abs_xn, yy, RNDN);
        mpfr_mul_si(t_yy, t_yy, sgn_yn, RNDN);
        mpfr_mul(tmp, abs_yn, xy, RNDN);
        mpfr_mul_si(tmp, tmp, sgn_xn, RNDN);
        mpfr_add(t_yy, t_yy, tmp, RNDN);
    } else if (kind == FSB_FLAVOR_PERPENDICULAR_BS) {   /* 2 (xn sgn_yn dy + dx |yn|) */
        mpfr_mul(t_yx, xn, yx, RNDN);
        mpfr_mul_si(t_yx, t_yx, sgn_yn, RNDN);
        mpfr_mul(tmp, abs_yn, xx, RNDN);
        mpfr_add(t_yx, t_yx, tmp, RNDN);
        mpfr_mul(t_yy, xn, yy, RNDN);
        mpfr_mul_si(t_yy, t_yy, sgn_yn, RNDN);
        mpfr_mul(tmp, abs_yn, xy, RNDN);
        mpfr_add(t_yy, t_yy, tmp, RNDN);
    } else {                               /* shark fin, celtic: 2 (xn dy + dx yn) */
        mpfr_mul(t_yx, xn, yx, RNDN);
        mpfr_mul(tmp, yn, xx, RNDN);
        mpfr_add(t_yx, t_yx, tmp, RNDN);
        mpfr_mul(t_yy, xn, yy, RNDN);
        mpfr_mul(tmp, yn, xy, RNDN);
        mpfr_add(t_yy, t_yy, tmp, RNDN);
    }
    mpfr_mul_si(xx, t_xx, 2, RNDN);
    mpfr_mul_si(xy, t_xy, 2, RNDN);
    mpfr_mul_si(yx, t_yx, 2, RNDN);
    mpfr_mul_si(yy, t_yy, 2, RNDN);
    mpfr_add_si(xx, xx, 1, RNDN);          /* derivatives with respect to (a, b) */
    mpfr_add_si(yy, yy, -1, RNDN);
}

/* FP_loop.pyx:1786-1825: (x, y) with [a b; c d] (x, y)^T = (e, f)^T */
static void matsolve2(fsb_mpfr_struct *x_res, fsb_mpfr_struct *y_res, const fsb_mpfr_struct *a,
                      const fsb_mpfr_struct *b, const fsb_mpfr_struct *c, const fsb_mpfr_struct *d,
                      const fsb_mpfr_struct *e, const fsb_mpfr_struct *f, fsb_mpfr_struct *delta,
                      fsb_mpfr_struct *tmp)
{
    mpfr_mul(delta, a, d, RNDN);
    mpfr_mul(tmp, c, b, RNDN);
    mpfr_sub(delta, delta, tmp, RNDN);
    mpfr_ui_div(delta, 1, delta, RNDN);
    mpfr_mul(x_res, d, e, RNDN);
    mpfr_mul(tmp, b, f, RNDN);
    mpfr_sub(x_res, x_res, tmp, RNDN);
    mpfr_mul(x_res, x_res, delta, RNDN);
    mpfr_mul(y_res, a, f, RNDN);
    mpfr_mul(tmp, c, e, RNDN);
    mpfr_sub(y_res, y_res, tmp, RNDN);
    mpfr_mul(y_res, y_res, delta, RNDN);
}

#define BS_NVARS 22
/* FP_loop.pyx:2357-2465: first i <= maxiter with |J_i^-1 (x_i, y_i)| < px */
int64_t fsb_ball_method_burning_ship(int flavor, const char *seed_x, const char *seed_y,
                                     int64_t prec_bits, const char *seed_px, int64_t maxiter,
                                     double M_divergence)
{
    fsb_mpfr_t v[BS_NVARS], lowp;
    int64_t ret = -1, i;
    int k;
    if (flavor < FSB_FLAVOR_BURNING_SHIP || flavor > FSB_FLAVOR_BUFFALO) return -2;
    for (k = 0; k < BS_NVARS; k++) mpfr_init2(v[k], prec_bits);
    mpfr_init2(lowp, 54);
#define xn v[0]
#define yn v[1]
#define a_t v[2]
#define b_t v[3]
#define xsq v[4]
#define ysq v[5]
#define xy_t v[6]
#define dxa v[7]
#define dxb v[8]
#define dya v[9]
#define dyb v[10]
#define delta v[11]
#define abs_xn v[12]
#define abs_yn v[13]
#define t_xx v[14]
#define t_xy v[15]
#define t_yx v[16]
#define t_yy v[17]
#define rx v[18]
#define ry v[19]
#define inv_pix v[20]
#define tmp v[21]
    if (mpfr_set_str(a_t, seed_x, 10, RNDN) != 0 || mpfr_set_str(b_t, seed_y, 10, RNDN) != 0 ||
        mpfr_set_str(tmp, seed_px, 10, RNDN) != 0) {
        ret = -3;
        goto done;
    }
    mpfr_ui_div(inv_pix, 1, tmp, RNDN);
    mpfr_set_si(xn, 0, RNDN); mpfr_set_si(yn, 0, RNDN);
    mpfr_set_si(dxa, 0, RNDN); mpfr_set_si(dxb, 0, RNDN);
    mpfr_set_si(dya, 0, RNDN); mpfr_set_si(dyb, 0, RNDN);
    for (i = 1; i <= maxiter; i++) {
        bs_jacobian_step(flavor, xn, yn, dxa, dxb, dya, dyb, abs_xn, abs_yn, t_xx, t_xy, t_yx,
                         t_yy, tmp);
        bs_step(flavor, xn, yn, a_t, b_t, xsq, ysq, xy_t);
        matsolve2(rx, ry, dxa, dxb, dya, dyb, xn, yn, delta, tmp);
        mpfr_mul(rx, rx, inv_pix, RNDN);
        mpfr_mul(ry, ry, inv_pix, RNDN);
        if (hypot(mpfr_get_d(xn, RNDN), mpfr_get_d(yn, RNDN)) > M_divergence) break;
        fsb_mpfr_abs(lowp, rx);
        if (mpfr_cmp_d(lowp, 1.) < 0) {
            fsb_mpfr_abs(lowp, ry);
            if (mpfr_cmp_d(lowp, 1.) < 0
                && hypot(mpfr_get_d(rx, RNDN), mpfr_get_d(ry, RNDN)) < 1.) { ret = i; break; }
        }
    }
done:
    for (k = 0; k < BS_NVARS; k++) mpfr_clear(v[k]);
    mpfr_clear(lowp);
    return ret;
}

/* FP_loop.pyx:2564-2755: Newton on (x_order, y_order)(a, b) with the full Jacobian;
 * divisors of `order` are not excluded.  Same return convention as
 * fsb_find_nucleus_mandelbrot. */
int fsb_find_any_nucleus_burning_ship(int flavor, const char *seed_x, const char *seed_y,
                                      int64_t prec_bits, int64_t order, int64_t max_newton,
                                      const char *seed_eps_cv, const char *seed_eps_valid,
                                      char *out_x, char *out_y, int64_t out_cap)
{
    fsb_mpfr_t v[BS_NVARS], eps, abs_diff;
    int cv = 0, k;
    int64_t i_newton, i;
    if (flavor < FSB_FLAVOR_BURNING_SHIP || flavor > FSB_FLAVOR_BUFFALO || order < 1) return -2;
    for (k = 0; k < BS_NVARS; k++) mpfr_init2(v[k], prec_bits);
    mpfr_init2(eps, 54); mpfr_init2(abs_diff, 54);
    if (mpfr_set_str(a_t, seed_x, 10, RNDN) != 0 || mpfr_set_str(b_t, seed_y, 10, RNDN) != 0 ||
        mpfr_set_str(eps, seed_eps_cv, 10, RNDN) != 0) {
        cv = -3;
        goto done;
    }
    mpfr_mul_si(eps, eps, 64, RNDN);
    for (i_newton = 0; i_newton < max_newton; i_newton++) {
        mpfr_set_si(xn, 0, RNDN); mpfr_set_si(yn, 0, RNDN);
        mpfr_set_si(dxa, 0, RNDN); mpfr_set_si(dxb, 0, RNDN);
        mpfr_set_si(dya, 0, RNDN); mpfr_set_si(dyb, 0, RNDN);
        for (i = 1; i <= order; i++) {
            bs_jacobian_step(flavor, xn, yn, dxa, dxb, dya, dyb, abs_xn, abs_yn, t_xx, t_xy,
                             t_yx, t_yy, tmp);
            bs_step(flavor, xn, yn, a_t, b_t, xsq, ysq, xy_t);
        }
        /* (da, db) = J^-1 (xn, yn) ; (a, b) -= (da, db) */
        matsolve2(rx, ry, dxa, dxb, dya, dyb, xn, yn, delta, tmp);
        mpfr_sub(a_t, a_t, rx, RNDN);
        mpfr_sub(b_t, b_t, ry, RNDN);
        mpfr_hypot(abs_diff, rx, ry, RNDN);
        if (mpfr_greaterequal_p(eps, abs_diff)) {
            mpfr_hypot(abs_diff, xn, yn, RNDN);
            if (mpfr_set_str(eps, seed_eps_valid, 10, RNDN) != 0) { cv = -3; goto done; }
            cv = mpfr_greaterequal_p(eps, abs_diff) ? 1 : 0;
            break;
        }
    }
    if (cv == 1 && (put_decimal(out_x, out_cap, a_t) != 0 || put_decimal(out_y, out_cap, b_t) != 0))
        cv = -4;
done:
    for (k = 0; k < BS_NVARS; k++) mpfr_clear(v[k]);
    mpfr_clear(eps); mpfr_clear(abs_diff);
    return cv;
}
#undef xn
#undef yn
#undef a_t
#undef b_t
#undef xsq
#undef ysq
#undef xy_t
#undef dxa
#undef dxb
#undef dya
#undef dyb
#undef delta
#undef abs_xn
#undef abs_yn
#undef t_xx
#undef t_xy
#undef t_yx
#undef t_yy
#undef rx
#undef ry
#undef inv_pix
#undef tmp


/* ======================================================================== */
/* Period and nucleus of the reference point, z^N + c (N > 2)               */

/* iter_deriv_Mn then iter_Mn, FP_loop.pyx:167-211 */
static void mn_step_deriv(unsigned long exponent, fsb_mpc_struct *z, fsb_mpc_struct *dz,
                          const fsb_mpc_struct *c, fsb_mpc_struct *tmp)
{
    if (exponent == 2) { m2_step_deriv(z, dz, c, tmp); return; }
    mpc_pow_ui(tmp, z, exponent - 1, RNDNN);
    mpc_mul(tmp, dz, tmp, RNDNN);
    mpc_mul_ui(dz, tmp, exponent, RNDNN);
    mpc_add_ui(dz, dz, 1, RNDNN);
    mpc_pow_ui(tmp, z, exponent, RNDNN);
    mpc_add(z, tmp, c, RNDNN);
}

/* perturbation_mandelbrotN_select_ball_method, FP_loop.pyx:631-758 */
int64_t fsb_ball_method_mandelbrot_n(uint32_t exponent, const char *seed_x, const char *seed_y,
                                     int64_t prec_bits, const char *seed_px, int64_t maxiter,
                                     double M_divergence)
{
    fsb_mpc_t c, z, dz, tmp, r;
    fsb_mpfr_t ar, x_t, y_t, pix, inv_pix;
    int64_t ret = -1, i;
    if (exponent < 2) return -2;
    mpc_init2(c, prec_bits); mpc_init2(z, prec_bits); mpc_init2(dz, prec_bits);
    mpc_init2(tmp, prec_bits); mpc_init2(r, prec_bits);
    mpfr_init2(ar, 54); mpfr_init2(x_t, prec_bits); mpfr_init2(y_t, prec_bits);
    mpfr_init2(pix, prec_bits); mpfr_init2(inv_pix, prec_bits);
    if (mpfr_set_str(x_t, seed_x, 10, RNDN) != 0 || mpfr_set_str(y_t, seed_y, 10, RNDN) != 0 ||
        mpfr_set_str(pix, seed_px, 10, RNDN) != 0) {
        ret = -3;
        goto done;
    }
    mpc_set_fr_fr(c, x_t, y_t, RNDNN);
    mpfr_ui_div(inv_pix, 1, pix, RNDN);
    mpc_set_si_si(z, 0, 0, RNDNN);
    mpc_set_si_si(dz, 0, 0, RNDNN);
    for (i = 1; i <= maxiter; i++) {
        mn_step_deriv(exponent, z, dz, c, tmp);
        mpc_div(r, z, dz, RNDNN);
        mpc_mul_fr(r, r, inv_pix, RNDNN);
        if (hypot(mpfr_get_d(z->re, RNDN), mpfr_get_d(z->im, RNDN)) > M_divergence) break;
        mpc_abs(ar, r, RNDN);
        if (mpfr_cmp_d(ar, 1.) < 0) { ret = i; break; }
    }
done:
    mpc_clear(c); mpc_clear(z); mpc_clear(dz); mpc_clear(tmp); mpc_clear(r);
    mpfr_clear(ar); mpfr_clear(x_t); mpfr_clear(y_t); mpfr_clear(pix); mpfr_clear(inv_pix);
    return ret;
}

/* perturbation_mandelbrotN_select_find_any_nucleus, FP_loop.pyx:1159-1340 */
int fsb_find_any_nucleus_mandelbrot_n(uint32_t exponent, const char *seed_x, const char *seed_y,
                                      int64_t prec_bits, int64_t order, int64_t max_newton,
                                      const char *seed_eps_cv, const char *seed_eps_valid,
                                      char *out_x, char *out_y, int64_t out_cap)
{
    fsb_mpc_t c, zr, dzr, t1;
    fsb_mpfr_t x_t, y_t, abs_diff, eps;
    int cv = 0;
    int64_t i_newton, i;
    if (order < 1 || exponent < 2) return -2;
    mpc_init2(c, prec_bits); mpc_init2(zr, prec_bits); mpc_init2(dzr, prec_bits);
    mpc_init2(t1, prec_bits);
    mpfr_init2(x_t, prec_bits); mpfr_init2(y_t, prec_bits);
    mpfr_init2(abs_diff, 54); mpfr_init2(eps, 54);
    if (mpfr_set_str(x_t, seed_x, 10, RNDN) != 0 || mpfr_set_str(y_t, seed_y, 10, RNDN) != 0 ||
        mpfr_set_str(eps, seed_eps_cv, 10, RNDN) != 0) {
        cv = -3;
        goto done;
    }
    mpc_set_fr_fr(c, x_t, y_t, RNDNN);
    mpfr_mul_si(eps, eps, 64, RNDN);
    for (i_newton = 0; i_newton < max_newton; i_newton++) {
        mpc_set_si_si(zr, 0, 0, RNDNN);
        mpc_set_si_si(dzr, 0, 0, RNDNN);
        for (i = 1; i <= order; i++) mn_step_deriv(exponent, zr, dzr, c, t1);
        mpc_div(t1, zr, dzr, RNDNN);
        mpc_sub(c, c, t1, RNDNN);
        mpc_abs(abs_diff, t1, RNDN);
        if (mpfr_greaterequal_p(eps, abs_diff)) {
            mpc_abs(abs_diff, zr, RNDN);
            if (mpfr_set_str(eps, seed_eps_valid, 10, RNDN) != 0) { cv = -3; goto done; }
            cv = mpfr_greaterequal_p(eps, abs_diff) ? 1 : 0;
            break;
        }
    }
    if (cv == 1 && (put_decimal(out_x, out_cap, c->re) != 0 || put_decimal(out_y, out_cap, c->im) != 0))
        cv = -4;
done:
    mpc_clear(c); mpc_clear(zr); mpc_clear(dzr); mpc_clear(t1);
    mpfr_clear(x_t); mpfr_clear(y_t); mpfr_clear(abs_diff); mpfr_clear(eps);
    return cv;
}
