/*
 * bs_fused_check.cpp -- TEST INFRASTRUCTURE: the fused tiny-perturbation forms of
 * the burning-ship iteration (fsb_lane.cuh: bs_tiny_f1_zn / bs_tiny_f1_hessian)
 * against the chain of Xrange operators the kernels otherwise run
 * (bs_p_iter_zn<XF> / bs_p_iter_hessian<XF>), on random states, bit for bit on
 * the VALUES (to_std of mantissa ratios is not enough at 1e-500: values are
 * compared as normalised (mantissa, exponent) pairs).
 * Returns the number of mismatches.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../fractalshades_b200/csrc/fsb_lane.cuh"
using namespace fsb;

static uint64_t rng_state = 88172645463325252ULL;
static double urand() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return (double)(rng_state >> 11) / 9007199254740992.0; }
static double srand1() { double v = 0.5 + urand(); return (urand() < 0.5) ? -v : v; }
static XF rxf(int e_lo, int e_hi, bool allow_zero)
{
    if (allow_zero && urand() < 0.05) return mkXF(0., 0);
    /* un-normalised representations as the lazy renormalisation leaves them */
    const int k = (int)(urand() * 60) - 30;
    return mkXF(ldexp(srand1(), k), e_lo + (int)(urand() * (e_hi - e_lo)) - k);
}
static bool same(XF a, XF b)
{
    const XF na = normalize(a.m, a.e), nb = normalize(b.m, b.e);
    if (na.m == 0. && nb.m == 0.) return true;
    return na.m == nb.m && na.e == nb.e;
}

extern "C" int bs_fused_check(int n, int verbose)
{
    int bad = 0, used = 0;
    for (int it = 0; it < n; it++) {
        const double rx = ldexp(srand1(), (int)(urand() * 40) - 30), ry = ldexp(srand1(), (int)(urand() * 40) - 30);
        const int base = -400 - (int)(urand() * 1500);
        XF x = rxf(base - 40, base, true), y = rxf(base - 40, base, true);
        XF a = rxf(base - 60, base + 10, true), b = rxf(base - 60, base + 10, true);
        XF dxa = rxf(-300, 300, true), dxb = rxf(-300, 300, true), dya = rxf(-300, 300, true), dyb = rxf(-300, 300, true);
        XF ra = rxf(-200, 600, true), rb = rxf(-200, 600, true), rc = rxf(-200, 600, true), rd = rxf(-200, 600, true);
        if (!bs_tiny_f1_ok(rx, ry, x, y)) continue;
        used++;
        /* operator chain */
        XF gx = x, gy = y, ga = dxa, gb = dxb, gc = dya, gd = dyb;
        bs_p_iter_hessian(1, gx, gy, ga, gb, gc, gd, to_xr(rx), to_xr(ry), ra, rb, rc, rd);
        bs_p_iter_zn(1, gx, gy, to_xr(rx), to_xr(ry), a, b);
        /* fused */
        XF fx = x, fy = y, fa = dxa, fb = dxb, fc = dya, fd = dyb;
        bs_tiny_f1_hessian(fx, fy, fa, fb, fc, fd, rx, ry, ra, rb, rc, rd);
        bs_tiny_f1_zn(fx, fy, rx, ry, a, b);
        const bool ok = same(gx, fx) && same(gy, fy) && same(ga, fa) && same(gb, fb) && same(gc, fc) && same(gd, fd);
        if (!ok) {
            bad++;
            if (verbose && bad < 6)
                printf("mismatch %d: x %d y %d dxa %d dxb %d dya %d dyb %d  (rx %g ry %g)\n", it, same(gx, fx), same(gy, fy),
                       same(ga, fa), same(gb, fb), same(gc, fc), same(gd, fd), rx, ry);
        }
    }
    if (verbose) printf("checked %d of %d random states, %d mismatches\n", used, n, bad);
    return used > n / 4 ? bad : -1;
}
