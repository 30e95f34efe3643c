/*
 * fsb200_orbit.h -- C ABI of libfsb200_orbit.so (host, no CUDA).
 *
 * Native full-precision reference orbit.  Each entry point replaces one
 * function of the reference's Cython extension `fractalshades.mpmath_utils.
 * FP_loop` (the only compiled component of the reference):
 *
 *   fsb_orbit_mandelbrot     <- perturbation_mandelbrot_FP_loop /
 *                               perturbation_mandelbrotN_FP_loop
 *                               (mpmath_utils/FP_loop.pyx:237-421)
 *   fsb_orbit_burning_ship   <- perturbation_nonholomorphic_FP_loop
 *                               (mpmath_utils/FP_loop.pyx:1828-1979)
 *
 * Plain pointers and sizes only.  All buffers are caller-owned.
 */
#ifndef FSB200_ORBIT_H
#define FSB200_ORBIT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Flavours of the non-holomorphic family (FP_loop.pyx:1343-1347). */
enum {
    FSB_FLAVOR_BURNING_SHIP = 1,
    FSB_FLAVOR_PERPENDICULAR_BS = 2,
    FSB_FLAVOR_SHARK_FIN = 3,
    FSB_FLAVOR_CELTIC = 4,
    FSB_FLAVOR_BUFFALO = 5
};

/* One orbit point that underflows a double: value = (mx * 2^ex, my * 2^ey),
 * mantissas in [0.5, 1) as returned by mpfr_get_d_2exp (FP_loop.pyx:424-454). */
typedef struct fsb_orbit_xr {
    int64_t index;
    double mx;
    double my;
    int32_t ex;
    int32_t ey;
} fsb_orbit_xr;

/*
 * orbit      : out, 2*(max_iter+1) doubles (re, im interleaved); orbit[0:2]=0
 * exponent   : 2 for z^2+c, >2 for z^n+c
 * need_xrange: register points with |z| < 1e-300 into xr_out
 * M          : escape radius for the reference point
 * seed_x/y   : decimal strings of the reference point, parsed at prec_bits
 * xr_out     : out, capacity xr_cap entries; *xr_count receives the number used
 * returns    : first invalid orbit index (escape index, or max_iter+1 if the
 *              orbit never escaped); < 0 on error (-2 bad argument, -3 bad
 *              number string, -4 xr_out too small)
 */
int64_t fsb_orbit_mandelbrot(double *orbit, int64_t max_iter, uint32_t exponent,
                             int need_xrange, double M, const char *seed_x,
                             const char *seed_y, int64_t prec_bits,
                             fsb_orbit_xr *xr_out, int64_t xr_cap,
                             int64_t *xr_count);

/* Same contract; `flavor` is one of FSB_FLAVOR_*; a point is registered as
 * Xrange when |x| < 1e-300 or |y| < 1e-300 (FP_loop.pyx:1933). */
int64_t fsb_orbit_burning_ship(double *orbit, int64_t max_iter, int flavor,
                               int need_xrange, double M, const char *seed_x,
                               const char *seed_y, int64_t prec_bits,
                               fsb_orbit_xr *xr_out, int64_t xr_cap,
                               int64_t *xr_count);

/* ---- period and nucleus of the reference point (holomorphic z^2 + c) ----
 *   fsb_ball_method_mandelbrot   <- perturbation_mandelbrot_ball_method
 *                                   (FP_loop.pyx:605-758): first iteration i
 *                                   <= maxiter with |z_i / (dz_i/dc)| < px;
 *                                   -1 when the point escapes first or no such
 *                                   i exists, -3 on a bad number string
 *   fsb_find_nucleus_mandelbrot  <- perturbation_mandelbrot_find_nucleus
 *                                   (:900-1118, any_nucleus = 0: the roots of
 *                                   the divisors of `order` are divided out) /
 *                                   perturbation_mandelbrot_find_any_nucleus
 *                                   (:1119-1340, any_nucleus = 1).
 *                                   eps_cv is multiplied by 64 as in the
 *                                   reference; converged and |z_order| <=
 *                                   eps_valid -> returns 1 and writes the
 *                                   nucleus as decimal strings (exact round
 *                                   trip at prec_bits) into out_x / out_y
 *                                   (capacity out_cap bytes each); 0 = did not
 *                                   converge; < 0 error (-4: out_cap too small)
 */
int64_t fsb_ball_method_mandelbrot(const char *seed_x, const char *seed_y, int64_t prec_bits,
                                   const char *seed_px, int64_t maxiter, double M_divergence);
int fsb_find_nucleus_mandelbrot(const char *seed_x, const char *seed_y, int64_t prec_bits,
                                int64_t order, int64_t max_newton, const char *seed_eps_cv,
                                const char *seed_eps_valid, int any_nucleus, char *out_x,
                                char *out_y, int64_t out_cap);

/* ---- the same for the burning-ship family --------------------------------
 *   fsb_ball_method_burning_ship       <- perturbation_nonholomorphic_ball_method
 *                                         (FP_loop.pyx:2357-2465; the per-flavour
 *                                         wrappers :2217-2355): first iteration i with
 *                                         |J_i^-1 (x_i, y_i)| < px, J the Jacobian of
 *                                         (x_i, y_i) with respect to (a, b)
 *                                         (iter_J_* :1458-1743); -1 / -3 as above,
 *                                         -2 unknown flavour
 *   fsb_find_any_nucleus_burning_ship  <- perturbation_nonholomorphic_find_any_nucleus
 *                                         (:2564-2755): Newton descent with the full
 *                                         Jacobian (matsolve :1799-1825); divisors of
 *                                         `order` are not excluded (the reference has
 *                                         no other form for this family,
 *                                         burning_ship.py:1189-1200); returns as
 *                                         fsb_find_nucleus_mandelbrot
 */
int64_t fsb_ball_method_burning_ship(int flavor, const char *seed_x, const char *seed_y,
                                     int64_t prec_bits, const char *seed_px, int64_t maxiter,
                                     double M_divergence);
int fsb_find_any_nucleus_burning_ship(int flavor, const char *seed_x, const char *seed_y,
                                      int64_t prec_bits, int64_t order, int64_t max_newton,
                                      const char *seed_eps_cv, const char *seed_eps_valid,
                                      char *out_x, char *out_y, int64_t out_cap);

/* ---- and for z^N + c --------------------------------------------------------
 *   fsb_ball_method_mandelbrot_n       <- perturbation_mandelbrotN_select_ball_method
 *                                         (FP_loop.pyx:631-758; iter_deriv_Mn / iter_Mn
 *                                         :167-211)
 *   fsb_find_any_nucleus_mandelbrot_n  <- perturbation_mandelbrotN_select_find_any_nucleus
 *                                         (:1159-1340); the reference has no
 *                                         divide-by-divisors form for N > 2
 *                                         (mandelbrot_Mn.py:765-779)
 */
int64_t fsb_ball_method_mandelbrot_n(uint32_t exponent, const char *seed_x, const char *seed_y,
                                     int64_t prec_bits, const char *seed_px, int64_t maxiter,
                                     double M_divergence);
int fsb_find_any_nucleus_mandelbrot_n(uint32_t exponent, const char *seed_x, const char *seed_y,
                                      int64_t prec_bits, int64_t order, int64_t max_newton,
                                      const char *seed_eps_cv, const char *seed_eps_valid,
                                      char *out_x, char *out_y, int64_t out_cap);

#ifdef __cplusplus
}
#endif
#endif /* FSB200_ORBIT_H */
