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
