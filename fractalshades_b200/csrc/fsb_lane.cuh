/*
 * fsb_lane.cuh -- per-frame device structures and the per-pixel ("lane") pieces
 * of the perturbation loops, written once for the device kernels and for the
 * host: table lookups (reference orbit, BLA tree), the model formulas, the
 * fused Xrange forms, and the event-driven lane state machine of the
 * persistent holomorphic kernel `k_perturb_m2_v2` (fsb_kernels.cuh).
 *
 * Everything here is `__host__ __device__`: the CPU test-suite compiles the very
 * same state machine (tests/emul/) and checks it against the oracle bit for bit
 * -- test infrastructure only; the product runs it on the GPU.
 */
#pragma once
#include "fsb_math.cuh"

/* FSB_XF_FUSED_DOT=1: burning-ship Xrange BLA step, dot products with ONE alignment
 * (xf_dot2 / xf_dot4).  Bit-exact (strict build against the oracle, full-size config 4 tiles)
 * and half the instructions of the chain on paper -- but the chain has to stay as the guarded
 * fallback, and this kernel is bound by instruction fetch: 23.5 ms inlined, 25.4 ms with the
 * fallback out of line (more spills), against 22.3 ms.  Off. */
#ifndef FSB_XF_FUSED_DOT
#define FSB_XF_FUSED_DOT 0
#endif
/* FSB_BLA_LOOKUP4=1: radius tests of the BLA lookup decided on the larger component alone
 * where that is conclusive (ref_bla_get4_).  Same decisions; measured slower (config 2 / 3:
 * 10.91 / 24.69 ms against 10.75 / 23.52): one more table in the walk, and the square test it
 * saves was not the expensive part.  Off. */
#ifndef FSB_BLA_LOOKUP4
#define FSB_BLA_LOOKUP4 0
#endif

namespace fsb {

#ifdef __CUDA_ARCH__
template <class T> __device__ __forceinline__ T ldg_(const T *p) { return __ldg(p); }
__device__ __forceinline__ int ffs_(int x) { return __ffs(x); }
__device__ __forceinline__ int clz_(int x) { return __clz(x); }
#else
template <class T> inline T ldg_(const T *p) { return *p; }
inline int ffs_(int x) { return __builtin_ffs(x); }
inline int clz_(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
#endif
FSB_HD int imax(int a, int b) { return a > b ? a : b; }
FSB_HD int imin(int a, int b) { return a < b ? a : b; }

/* Pixel projection (projection.py) and the derivative modifier applied at the
 * end of the perturbation loops (perturbation.py:1387-1388, 1772-1776).
 * kind: 0 Cartesian (identity), 1 Expmap; mod_kind: 0 none, 1 Expmap
 * exp(Re(k pix) + mod_param), 2 Cartesian(expmap_seam) |pix + 1e-6| mod_param */
struct ProjDev {
    int kind, mod_kind;
    double hmoy, k_re, k_im, mod_param;
};

struct FrameDev {
    long long L;
    const C *Zn;
    /* holomorphic */
    const C *dZndc; const int *dZndc_e;
    const C *dZndc_std;   /* Xrange frames: flushed fp64 mirror of dZndc (fast path) */
    const C *dZndz; const int *dZndz_e;
    const C *ref_xr; const int *ref_xr_e;
    /* burning ship family */
    const double *dP[4]; const int *dP_e[4];
    const double *dP_std[4];   /* Xrange frames: flushed fp64 mirrors (fast path) */
    const double *refx_xr; const int *refx_xr_e;
    const double *refy_xr; const int *refy_xr_e;
    long long n_xr; const int *ref_index_xr;
    long long ref_div_iter, ref_order;
    double drift[2]; int drift_e[2];
    double lin_scale; int lin_scale_e;
    double lin_mat[4];
    const double *M_bla; const double *r_bla;
    long long bla_len; int stages_bla;
    long long max_iter;
    double Mdiv_sq, eps_sq;
    int calc_orbit; long long backshift;
    int flavor;
    /* 32-bit mirrors used by the pixel kernels (every orbit index fits) */
    int Li, ref_div_i, order_i /* 0: not a cycle */, first_invalid_i, max_iter_i, n_xr_i;
    long long zstride;   /* row stride of the output planes (>= npts of the launch) */
    /* Perturbation_mandelbrot_N (appended: the offsets above are those of every
     * other kernel): exponent and comb(N, k) as doubles, k = 0..N */
    int nexp;
    const double *cbinom;
    int ref_div_m1_i;    /* ref_div_i - 1: the rebase test compares against a constant-bank operand */
    /* k_perturb_m2_v2: interleaved orbit table, one 32-byte record per index
     *   T2[i] = {Zn[i+1], FSB_TSCALE * dZndc[i]}
     * (dZndc: the fp64 mirror in Xrange frames) -- all an iteration reads, in ONE 32-byte
     * load: the L1 -> register write-back path (32 lanes x bytes per load, broadcast or
     * not) is what bounds the hot loop next to the FP64 pipe, measured at 90 % with
     * 64-byte records; h3[h3_slot(w)] = high word of the stage-3 BLA radius of index w where
     * the loop looks the tree up (w a multiple of 8), else 0; and the exponent bound of the escape pre-test (both
     * parts of Z + z below 2^k with 2^(2k+1) <= Mdiv_sq) */
    const double *T2;
    const unsigned *h3;
    unsigned esc_hi;
    /* high words of fl((r up)^2) per BLA node, up = 1 and 2^600 (bla_r2hi) */
    const int *r2hi, *r2hi_up;
    /* high word of r per BLA node, 0 when |z| < r can never hold (bla_rhi) */
    const int *rhi;
};

struct StdDev {
    double center_re, center_im, dx;
    double lin_mat[4];
    long long max_iter;
    double Mdiv_sq, eps_sq;
    int calc_d2, calc_orbit;
    long long backshift;
    int flavor;
    long long zstride;   /* row stride of the output planes */
    int nexp;            /* k_std_mn: exponent of Mandelbrot_N */
};

/* Work units.  A launch covers units [unit_lo, unit_hi); one warp takes one unit
 * at a time from a global counter (warp-level work stealing) and its 32 lanes
 * take the unit's 32 points.
 *   flat list   (tiles == nullptr): unit u = points [32u, 32u + 32)
 *   tile list   the point list is a concatenation of row-major tiles
 *               (core.py:1767-1830); a unit is an 8 x 4 pixel patch of one tile,
 *               lane l -> (row l >> 3, column l & 7).  Neighbouring pixels leave the
 *               loop at nearby iteration counts, and a compact footprint keeps
 *               more lanes alive than a 32 x 1 strip (measured 4-8 %).
 * Each tile descriptor is {first unit, first point, width, height}. */
struct Tiling {
    const int4 *tiles;
    int n_tiles;
    int unit_lo, unit_hi;
};

FSB_HD C ldC(const C *p, long long i)
{
#ifdef __CUDA_ARCH__
    double2 v = __ldg(reinterpret_cast<const double2 *>(p) + i);
    return mkC(v.x, v.y);
#else
    return p[i];
#endif
}
FSB_HD void stC(double *Z, long long row, long long npts, long long i, C v)
{
#ifdef __CUDA_ARCH__
    reinterpret_cast<double2 *>(Z)[row * npts + i] = make_double2(v.re, v.im);
#else
    Z[2 * (row * npts + i)] = v.re; Z[2 * (row * npts + i) + 1] = v.im;
#endif
}

/* ======================================================================== */
/* Reference-path access                                                     */

/* Position of idx in the sorted ref_index_xr or -1: stateless equivalent of
 * the cursor of perturbation.py:2519-2588. */
FSB_HD int xr_find(const int *index, int n, int idx)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (ldg_(index + mid) < idx) lo = mid + 1; else hi = mid;
    }
    if (lo < n && ldg_(index + lo) == idx) return lo;
    return -1;
}

FSB_HD int bla_index(int i, int stg)
{
    return 2 * i + ((1 << stg) - 1);
}
FSB_HD long long bla_index64(long long i, int stg)
{
    return 2 * i + ((1LL << stg) - 1);
}

/* perturbation.py:2116-2170.  Returns the step (0 = no BLA applicable) and
 * the node index.  All orbit indices fit 32 bits (the host checks it).
 *
 * The reference walks the stages from the highest admissible one down and
 * takes the first node with |z| < r.  A merged node's radius is
 * min(r_first_half, ...) (perturbation.py:2021), so along one start index the
 * radii never grow with the stage (and a NaN radius stays NaN upwards): a point
 * that fails the lowest stored stage fails them all.  That stage is tested
 * first -- about two checks in three end there -- and |z| >= max(|re|, |im|)
 * (the rounded hypot is never below its larger argument) rejects most of those
 * before the hypot is even evaluated.  Same node as the top-down walk. */
FSB_HD int ref_bla_get(const double *__restrict__ r_bla,
                                           int stages_bla, C zn, int n_iter,
                                           int first_invalid, int &index_out)
{
    const int it = n_iter >> 3;
    const int invalid_step = first_invalid - n_iter;
    /* skip the levels whose step cannot fit before the first invalid index */
    if (invalid_step <= 8 || stages_bla < 4) return 0;
    const int base = 2 * it - 1;
    const double r3 = ldg_(r_bla + base + 1);
    if (!(fabs(zn.re) < r3 && fabs(zn.im) < r3)) return 0;
    const double az = cabs_rn(zn);
    if (!(az < r3)) return 0;
    int stages = stages_bla - 1;
    if (it != 0) {
        int s = 3 + (ffs_(it) - 1);
        if (s < stages) stages = s;
    }
    const int top = 31 - clz_(invalid_step - 1);   /* largest stg with 2^stg < invalid_step */
    if (stages > top) stages = top;
    /* Measured and dropped: bisecting the stage (the predicate is monotone, same node
     * bit for bit) -- the walk from the top usually ends at its first or second node;
     * config 2 13.58 -> 14.48 ms, config 3 28.9 -> 29.7 ms. */
    for (int stg = stages; stg > 3; stg--) {
        const int ib = base + (1 << (stg - 3));
        if (az < ldg_(r_bla + ib)) { index_out = ib; return 1 << stg; }
    }
    index_out = base + 1;
    return 8;
}

/* ======================================================================== */
/* Holomorphic perturbation (Mandelbrot power 2)                             */

template <class T, class R>
FSB_HD T p_iter_zn(T z, R ref_zn, T c)
{
    return z * (z + 2. * ref_zn) + c; /* mandelbrot_M2.py:607-610 */
}
template <class T, class R, class D>
FSB_HD T p_iter_deriv(T z, T dz, R ref_zn, D ref_d)
{
    return 2. * ((ref_zn + z) * dz + ref_d * z); /* mandelbrot_M2.py:611-622 */
}

#if !defined(FSB_STRICT) && defined(FSB_FMA_CHAIN)
/* Default build, fp64 operands: the same two formulas as explicit FMA chains
 * (the reference's loops are numba fastmath: LLVM contracts and re-associates
 * them on FMA hosts, so no particular rounding sequence is "the" reference). */
FSB_HD C p_iter_zn(C z, C ref_zn, C c)
{
    const double tr = fma(2., ref_zn.re, z.re), ti = fma(2., ref_zn.im, z.im);
    return mkC(fma(z.re, tr, fma(-z.im, ti, c.re)), fma(z.re, ti, fma(z.im, tr, c.im)));
}
FSB_HD C p_iter_deriv(C z, C dz, C ref_zn, C ref_d)
{
    /* s = 2 (Z + z) = (z + 2 Z) + z ; d = 2 Z' : both doublings are exact */
    const double sr = fma(2., ref_zn.re, z.re) + z.re, si = fma(2., ref_zn.im, z.im) + z.im;
    const double dr = ref_d.re + ref_d.re, di = ref_d.im + ref_d.im;
    return mkC(fma(sr, dz.re, fma(-si, dz.im, fma(dr, z.re, -(di * z.im)))),
               fma(sr, dz.im, fma(si, dz.re, fma(dr, z.im, di * z.re))));
}
#endif

/* Power-N Mandelbrot (models/mandelbrot_Mn.py:628-742): full binomial
 * expansions, written once for complex128 and Xrange like the reference's
 * numba closures.  Cb[k] = comb(N, k) as float64. */
template <class T>
FSB_HD T mn_dfdz(int nexp, T z)              /* :643-649 */
{
    T tmp = z;
    for (int k = 2; k < nexp; k++) tmp = tmp * z;
    return (double)nexp * tmp;
}
template <class T, class R>
FSB_HD T mn_iter_zn(int nexp, const double *__restrict__ Cb, T z, R ref_zn, T c)
{
    T tmp = z * (z + ldg_(Cb + 1) * ref_zn);                     /* :656-668 */
    R pk = ref_zn;
    for (int k = 2; k < nexp; k++) {
        pk = pk * ref_zn;
        tmp = z * (tmp + ldg_(Cb + k) * pk);
    }
    return tmp + c;
}
template <class T, class R, class D>
FSB_HD T mn_iter_deriv(int nexp, const double *__restrict__ Cb, T z, T dz,
                                           R ref_zn, D ref_d)   /* :670-728 */
{
    const double c1 = ldg_(Cb + 1);
    T mul = z + c1 * ref_zn;
    T tmp = z * mul;
    T dtmp = dz * mul + z * (dz + c1 * ref_d);
    R pk = ref_zn;
    for (int k = 2; k < nexp; k++) {
        const double ck = ldg_(Cb + k);
        D dpk = ((double)k * pk) * ref_d;
        pk = pk * ref_zn;
        mul = tmp + ck * pk;
        dtmp = dz * mul + z * (dtmp + ck * dpk);
        tmp = z * mul;
    }
    return dtmp;
}

/* Fast path of the Xrange kernels.  Xrange arithmetic is fp64 arithmetic with
 * an unbounded exponent: every operation is the correctly rounded result of
 * the same real operation.  While every live component of the pixel state is
 * a normal double comfortably inside the range, the plain fp64 operation
 * sequence therefore produces bit-identical values (scaling by 2^k is exact,
 * rounding is scale invariant, and the sub-1e-300 addends -- c, the Xrange
 * reference points, the tiny dZndc entries -- are below half an ulp of every
 * sum they enter).  `in_fast_range` is the guard: biased exponent within
 * [1023-460, 1023+900] for each component (zeros, denormals, inf and NaN all
 * fail).  When it fails the iteration is redone in exact Xrange arithmetic. */
#ifndef FSB_FAST_LO      /* narrowed by the test-suite to exercise the guard paths */
#define FSB_FAST_LO 460
#define FSB_FAST_HI 900
#endif
FSB_HD bool in_fast_range(double x)
{
    /* 2 * hi drops the sign bit; the exponent field then sits in bits 21-31 */
    const unsigned lo = (unsigned)(1023 - FSB_FAST_LO) << 21,
                   span = (unsigned)(FSB_FAST_LO + FSB_FAST_HI + 1) << 21;
    return 2u * (unsigned)hi32(x) - lo < span;
}
FSB_HD bool in_fast_range(C z)
{
#ifdef FSB_SC_GUARDS
    return in_fast_range(z.re) && in_fast_range(z.im);
#else
    return in_fast_range(z.re) & in_fast_range(z.im);
#endif
}

/* Fused Xrange forms of the BLA step (perturbation.py:1139-1153).  Same real
 * operations, in the same order and with the same roundings as the operator
 * chain `A * z + B * c` of numba_xr.py (product mantissas, alignment to the
 * larger exponent by exponent-field arithmetic clamped at 0, one rounded add
 * per component) -- an Xrange value does not depend on how its mantissa /
 * exponent split was normalised on the way, so only the bookkeeping is fused:
 * one alignment instead of three normalisations. */
FSB_HD int cexp_field(C v)
{
    return imax(expfield(v.re), expfield(v.im));
}
/* m * 2^shift on the exponent field (clamped at 0, mantissa bits kept, as
 * _exp2_shift does); zeros pass through */
FSB_HD double xshift(double m, int shift)
{
    const int hi = hi32(m);
    const int fld = (hi >> 20) & 0x7ff;
    int nf = fld + shift;
    nf = (fld == 0 || nf < 0) ? 0 : nf;
    return mk64((hi & (int)0x800fffff) | (nf << 20), lo32(m));
}
/* p 2^pe + q 2^qe: the aligned Xrange addition of two mantissa pairs.  When
 * every part is a normal double and stays one after its shift, the shift is an
 * addition on the high word (same bits as xshift). */
FSB_HD XC xr_sum2(C p, int pe, C q, int qe)
{
    const int hpr = hi32(p.re), hpi = hi32(p.im), hqr = hi32(q.re), hqi = hi32(q.im);
    const int fpr = (hpr >> 20) & 0x7ff, fpi = (hpi >> 20) & 0x7ff;
    const int fqr = (hqr >> 20) & 0x7ff, fqi = (hqi >> 20) & 0x7ff;
    const int fp = imax(fpr, fpi), fq = imax(fqr, fqi);
    const int ep = pe + (fp - 1023), eq = qe + (fq - 1023);
    int e = imax(ep, eq);
    if (fp == 0) e = eq;              /* a zero product does not set the exponent */
    if (fq == 0) e = ep;
    const int sp = (1023 - fp) - (e - ep), sq = (1023 - fq) - (e - eq);
#ifdef FSB_FAST_ALIGN   /* measured: slower (config 3 24.9 -> 26.1 ms: code size), off */
    if (imin(imin(fpr, fpi) + sp, imin(fqr, fqi) + sq) >= 1 && imin(imin(fpr, fpi), imin(fqr, fqi)) >= 1) {
        const int ap = sp << 20, aq = sq << 20;      /* (max field + shift <= 1023 by construction) */
        return mkXC(mkC(mk64(hpr + ap, lo32(p.re)) + mk64(hqr + aq, lo32(q.re)),
                        mk64(hpi + ap, lo32(p.im)) + mk64(hqi + aq, lo32(q.im))), e);
    }
#endif
    return mkXC(mkC(xshift(p.re, sp) + xshift(q.re, sq), xshift(p.im, sp) + xshift(q.im, sq)), e);
}
FSB_HD XC xr_lin(C A, XC z, C B, XC c)
{
    return xr_sum2(A * z.m, z.e, B * c.m, c.e);
}
FSB_HD XC xr_mulc(C A, XC d)
{
    const C p = A * d.m;
    const int hr = hi32(p.re), hi_ = hi32(p.im);
    const int fr = (hr >> 20) & 0x7ff, fi = (hi_ >> 20) & 0x7ff;
    const int fp = imax(fr, fi), sh = 1023 - fp;
#ifdef FSB_FAST_ALIGN   /* measured: slower (config 3 24.9 -> 26.1 ms: code size), off */
    if (imin(fr, fi) + sh >= 1 && imin(fr, fi) >= 1)
        return mkXC(mkC(mk64(hr + (sh << 20), lo32(p.re)), mk64(hi_ + (sh << 20), lo32(p.im))),
                    d.e + (fp - 1023));
#endif
    return mkXC(mkC(xshift(p.re, sh), xshift(p.im, sh)), d.e + (fp - 1023));
}
/* to_standard of a value whose mantissa parts are below 4 in magnitude */
FSB_HD C to_std_small(XC x)
{
    if (x.e < -1200)        /* rounds to (signed) zero */
        return mkC(mk64(hi32(x.m.re) & (int)0x80000000, 0), mk64(hi32(x.m.im) & (int)0x80000000, 0));
    return to_std(x);
}


/* flushed fp64 mirror of an Xrange table (fast path of the Xrange kernels):
 * exact for normal components, 0 below the normal range, NaN when too large */
FSB_HD double flush_component(double m, int e)
{
    if (m == 0.) return 0.;
    double nm; int ne;
    normalize_real(m, e, nm, ne);
    if (!(m == m) || ne > 1000) return mk64(0x7ff80000, 0);
    if (ne < -1022) return 0.;
    return ldexp(nm, ne);
}

/* ======================================================================== */
/* Burning-ship family: model formulas and fused Xrange forms                */

/* burning_ship.py:19-60 */
template <class T> FSB_HD T diffabs(T X, T x)
{
    if (X >= 0.) {
        if ((X + x) >= 0.) return 1. * x;
        return -(2. * X + x);
    }
    if ((X + x) <= 0.) return -x;
    return (2. * X + x);
}
template <class T> FSB_HD double ddiffabsdX(T X, T x)
{
    if (X >= 0.) { if ((X + x) >= 0.) return 0.; return -2.; }
    if ((X + x) <= 0.) return 0.;
    return 2.;
}
template <class T> FSB_HD double ddiffabsdx(T X, T x)
{
    if (X >= 0.) { if ((X + x) >= 0.) return 1.; return -1.; }
    if ((X + x) <= 0.) return -1.;
    return 1.;
}

/* burning_ship.py:535-619 */
template <class T>
FSB_HD void bs_p_iter_zn(int flavor, T &x, T &y, T rx, T ry, T a, T b)
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
FSB_HD void bs_p_iter_hessian(int flavor, T x, T y, T &dxa, T &dxb,
                                                  T &dya, T &dyb, T rx, T ry, T rdxa,
                                                  T rdxb, T rdya, T rdyb)
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

/* Fused Xrange forms of apply_BLA_BS / apply_BLA_deriv_BS
 * (perturbation.py:1793-1811), same idea as xr_lin: every product and every
 * add of the reference's left-to-right chain is performed once, on mantissas
 * aligned by exponent-field arithmetic; zero terms pass through an addition as
 * in _coexp_ufunc (numba_xr.py:716-733). */
#define XR_ZERO_E (-(1 << 28))
FSB_HD XF xf_prod(double M, XF v)
{
    const double p = M * v.m;
    const int fld = expfield(p);
    return mkXF(xshift(p, 1023 - fld), (fld == 0) ? XR_ZERO_E : v.e + (fld - 1023));
}
FSB_HD XF xf_sum(XF s, XF p)
{
    const int se = (s.m == 0.) ? XR_ZERO_E : s.e;
    const int e = imax(se, p.e);
    return mkXF(xshift(s.m, se - e) + xshift(p.m, p.e - e), e);
}
#if defined(__CUDA_ARCH__) && FSB_XF_FUSED_DOT
#define FSB_CHAIN_FN static __device__ __noinline__      /* the rare fallback stays out of line */
#else
#define FSB_CHAIN_FN FSB_HD
#endif
FSB_CHAIN_FN XF xf_dot2_chain(double m0, XF u, double m1, XF v)
{
    return xf_sum(xf_prod(m0, u), xf_prod(m1, v));
}
FSB_CHAIN_FN XF xf_dot4_chain(double m0, XF u, double m1, XF v, double m2, XF a,
                                      double m3, XF b)
{
    return xf_sum(xf_sum(xf_dot2_chain(m0, u, m1, v), xf_prod(m2, a)), xf_prod(m3, b));
}
/* The same sums with ONE alignment: every product is scaled to the largest exponent of
 * the lot and the additions run left to right on the scaled mantissas.  A scaling by a
 * power of two is exact and commutes with rounding, so each partial sum is the chain's
 * partial sum at another scale -- bit for bit -- as long as no scaled term leaves the
 * normal range: guarded (exponents within 900 of the largest, no zero / denormal /
 * non-finite product field), the chain otherwise.  An Xrange value does not depend on its
 * mantissa / exponent split (when leading terms cancel exactly the split differs from the
 * chain's, the value does not).  One shift per term instead of up to three. */
FSB_HD double xf_scaled(double p, int f, int shift)      /* p 2^shift, field stays in [1, 2046] */
{
    (void)f;
    return mk64(hi32(p) + (shift << 20), lo32(p));
}
FSB_HD XF xf_dot2(double m0, XF u, double m1, XF v)
{
#if FSB_XF_FUSED_DOT
    const double p0 = m0 * u.m, p1 = m1 * v.m;
    const int f0 = expfield(p0), f1 = expfield(p1);
    const int E0 = u.e + f0, E1 = v.e + f1;
    const int Emax = imax(E0, E1), Emin = imin(E0, E1);
    if ((unsigned)(f0 - 1) < 2046u && (unsigned)(f1 - 1) < 2046u && Emax - Emin <= 900) {
        const double t0 = xf_scaled(p0, f0, (1023 - f0) - (Emax - E0));
        const double t1 = xf_scaled(p1, f1, (1023 - f1) - (Emax - E1));
        return mkXF(t0 + t1, Emax - 1023);
    }
#endif
    return xf_dot2_chain(m0, u, m1, v);
}
FSB_HD XF xf_dot4(double m0, XF u, double m1, XF v, double m2, XF a,
                                      double m3, XF b)
{
#if FSB_XF_FUSED_DOT
    const double p0 = m0 * u.m, p1 = m1 * v.m, p2 = m2 * a.m, p3 = m3 * b.m;
    const int f0 = expfield(p0), f1 = expfield(p1), f2 = expfield(p2), f3 = expfield(p3);
    const int E0 = u.e + f0, E1 = v.e + f1, E2 = a.e + f2, E3 = b.e + f3;
    const int Emax = imax(imax(E0, E1), imax(E2, E3)), Emin = imin(imin(E0, E1), imin(E2, E3));
    if ((unsigned)(f0 - 1) < 2046u && (unsigned)(f1 - 1) < 2046u && (unsigned)(f2 - 1) < 2046u
        && (unsigned)(f3 - 1) < 2046u && Emax - Emin <= 900) {
        const double t0 = xf_scaled(p0, f0, (1023 - f0) - (Emax - E0));
        const double t1 = xf_scaled(p1, f1, (1023 - f1) - (Emax - E1));
        const double t2 = xf_scaled(p2, f2, (1023 - f2) - (Emax - E2));
        const double t3 = xf_scaled(p3, f3, (1023 - f3) - (Emax - E3));
        return mkXF(((t0 + t1) + t2) + t3, Emax - 1023);
    }
#endif
    return xf_dot4_chain(m0, u, m1, v, m2, a, m3, b);
}
FSB_HD XF xf_clean(XF x)      /* zero results carry exponent 0 */
{
    return mkXF(x.m, (x.m == 0.) ? 0 : x.e);
}
/* to_standard of a value whose mantissa is below 4 in magnitude */
FSB_HD double to_std_small(XF x)
{
    if (x.e < -1200) return mk64(hi32(x.m) & (int)0x80000000, 0);
    return to_std(x);
}

/* ---- fused exact form of a burning-ship iteration for a tiny (x, y) ---------
 * Flavour 1 ("Burning ship", burning_ship.py:539-553, 629-667) when the
 * perturbation is far below the reference point: |x| < 2^-60 |X| and
 * |y| < 2^-60 |Y|, X and Y normal doubles.  Then X + x, 2 X + x ... round to X,
 * 2 X ..., |X Y + (x y + x Y + y X)| keeps the sign of X Y (diffabs and its two
 * derivatives are linear: dX = 0, dx = sgn(X Y)), and x y is absorbed by x Y in
 * the first addition of the chain.  What is left is evaluated with the SAME
 * products and the SAME left-to-right aligned additions as the chain of
 * numba_xr operators the reference runs (about eighty of them per iteration),
 * skipping only the normalisations in between -- an Xrange value does not
 * depend on its mantissa / exponent split.  1e-500 frames spend most of their
 * instructions in the handful of tiny iterations that follow each rebase. */
FSB_HD XF xf_mul(XF u, XF v)              /* u * v as a normalised term */
{
    const double p = u.m * v.m;
    const int fld = expfield(p);
    return mkXF(xshift(p, 1023 - fld), (fld == 0) ? XR_ZERO_E : u.e + v.e + (fld - 1023));
}
/* aligned sum of two Xrange reals, zero operands passing through (numba_xr.py:716-733) */
FSB_HD XF xf_add(XF u, XF v)
{
    const int ue = (u.m == 0.) ? XR_ZERO_E : u.e, ve = (v.m == 0.) ? XR_ZERO_E : v.e;
    const int e = imax(ue, ve);
    return mkXF(xshift(u.m, ue - e) + xshift(v.m, ve - e), e);
}
FSB_HD XF xf_neg(XF u) { return mkXF(-u.m, u.e); }
FSB_HD XF xf_dbl(XF u) { return mkXF(u.m, u.e + 1); }      /* 2. * u */
/* true exponent of an Xrange real (very negative for 0) */
FSB_HD int xf_expo(XF u) { return (u.m == 0.) ? XR_ZERO_E : u.e + expfield(u.m) - 1023; }

FSB_HD bool bs_tiny_f1_ok(double rx, double ry, XF x, XF y)
{
    const int fx = expfield(rx), fy = expfield(ry);
    return fx >= 64 && fx < 1600 && fy >= 64 && fy < 1600
           && xf_expo(x) + 60 <= fx - 1023 && xf_expo(y) + 60 <= fy - 1023;
}
/* (x, y) <- one iteration; (a, b) = the pixel's c, (rx, ry) the reference point */
FSB_HD void bs_tiny_f1_zn(XF &x, XF &y, double rx, double ry, XF a, XF b)
{
    /* nx = x * (x + 2. * rx) - y * (y + 2. * ry) + a */
    const XF nx = xf_add(xf_add(xf_prod(2. * rx, x), xf_neg(xf_prod(2. * ry, y))), a);
    /* ny = 2. * diffabs(rx * ry, x * y + x * ry + y * rx) - b */
    XF s = xf_add(xf_prod(ry, x), xf_prod(rx, y));
    if ((rx < 0.) != (ry < 0.)) s = xf_neg(s);          /* sign of X Y */
    const XF ny = xf_add(xf_dbl(s), xf_neg(b));
    x = xf_clean(nx); y = xf_clean(ny);
}
/* the four derivatives; (ra, rb, rc, rd) = the reference's dX/da, dX/db, dY/da, dY/db */
FSB_HD void bs_tiny_f1_hessian(XF x, XF y, XF &dxa, XF &dxb, XF &dya, XF &dyb, double rx,
                               double ry, XF ra, XF rb, XF rc, XF rd)
{
    const bool neg = (rx < 0.) != (ry < 0.);           /* sign of X Y */
    /* ndxa = 2. * ((rx + x) * dxa + rdxa * x) - 2. * ((ry + y) * dya + rdya * y) */
    const XF ndxa = xf_add(xf_dbl(xf_add(xf_prod(rx, dxa), xf_mul(ra, x))),
                           xf_neg(xf_dbl(xf_add(xf_prod(ry, dya), xf_mul(rc, y)))));
    const XF ndxb = xf_add(xf_dbl(xf_add(xf_prod(rx, dxb), xf_mul(rb, x))),
                           xf_neg(xf_dbl(xf_add(xf_prod(ry, dyb), xf_mul(rd, y)))));
    /* dxa_ = dxa * y + x * dya + dxa * ry + x * rdya + dya * rx + y * rdxa ;
     * ndya = 2. * (0. * dXa + (+-1.) * dxa_) */
    XF da = xf_add(xf_mul(dxa, y), xf_mul(x, dya));
    da = xf_add(da, xf_prod(ry, dxa));
    da = xf_add(da, xf_mul(x, rc));
    da = xf_add(da, xf_prod(rx, dya));
    da = xf_add(da, xf_mul(y, ra));
    XF db = xf_add(xf_mul(dxb, y), xf_mul(x, dyb));
    db = xf_add(db, xf_prod(ry, dxb));
    db = xf_add(db, xf_mul(x, rd));
    db = xf_add(db, xf_prod(rx, dyb));
    db = xf_add(db, xf_mul(y, rb));
    if (neg) { da = xf_neg(da); db = xf_neg(db); }
    dxa = xf_clean(ndxa); dxb = xf_clean(ndxb);
    dya = xf_clean(xf_dbl(da)); dyb = xf_clean(xf_dbl(db));
}

/* ======================================================================== */
/* BLA lookup without the square root.
 *
 * `|z| < r` (perturbation.py:2158) is decided on squares: with
 *   s = fl(fl((a up)^2) + fl((b up)^2))   -- the quantity hypot_rn takes the root of
 *   q = fl((r up)^2)
 * the rounded hypot is below r whenever s is below q by at least one unit of
 * the HIGH word (a relative margin of 2^-21, far above the handful of ulp the
 * roundings can move either side), and not below it whenever s is above q by
 * that much.  High words closer than 2 (three cases in a million) are decided
 * by the exact hypot.  Same decisions as `cabs_rn(z) < r`, bit for bit. */
struct AbsSq { double s, up; C z; };
FSB_HD AbsSq abs_sq(C z)
{
    AbsSq q;
    q.z = z;
    const double a = fabs(z.re), b = fabs(z.im);
    const int ea = imax(expfield(a), expfield(b));
    q.up = 1.;
    if (ea > 1023 + 500) q.up = 0x1p-600;
    else if (ea < 1023 - 500) q.up = 0x1p600;
    const double au = mul_rn(a, q.up), bu = mul_rn(b, q.up);
    q.s = add_rn(mul_rn(au, au), mul_rn(bu, bu));
    return q;
}
FSB_HD bool abs_lt(const AbsSq &q, double r)
{
    const double ru = mul_rn(r, q.up);
    const double r2 = mul_rn(ru, ru);
    const int d = hi32(r2) - hi32(q.s);
    /* non-negative, non-NaN operands order like their high words */
    if (r == r && q.s == q.s && r >= 0.) {
        if (d >= 2) return true;
        if (d <= -2) return false;
    }
    return cabs_rn(q.z) < r;
}

/* perturbation.py:2116-2170 with the lookup order of ref_bla_get (lowest
 * stored stage first) and the square-free comparison above. */
FSB_HD int ref_bla_get2(const double *__restrict__ r_bla, int stages_bla, C zn, int w,
                        int first_invalid, int &index_out)
{
    const int it = w >> 3;
    const int invalid_step = first_invalid - w;
    if (invalid_step <= 8 || stages_bla < 4) return 0;
    const int base = 2 * it - 1;
    const double r3 = ldg_(r_bla + base + 1);
    if (!(fabs(zn.re) < r3 && fabs(zn.im) < r3)) return 0;
    const AbsSq q = abs_sq(zn);
    if (!abs_lt(q, r3)) return 0;
    int stages = stages_bla - 1;
    if (it != 0) stages = imin(stages, 3 + (ffs_(it) - 1));
    stages = imin(stages, 31 - clz_(invalid_step - 1));   /* largest stg with 2^stg < invalid_step */
    for (int stg = stages; stg > 3; stg--) {
        const int ib = base + (1 << (stg - 3));
        if (abs_lt(q, ldg_(r_bla + ib))) { index_out = ib; return 1 << stg; }
    }
    index_out = base + 1;
    return 8;
}

#if defined(__CUDA_ARCH__) && defined(FSB_NOINLINE_EXACT)
#define FSB_COLD_FN static __device__ __noinline__
#else
#define FSB_COLD_FN static inline
#endif
/* the rare exact decisions, out of line */
FSB_COLD_FN bool abs_lt_exact(C z, double r) { return cabs_rn(z) < r; }
FSB_COLD_FN int ref_bla_get2_cold(const double *r_bla, int stages_bla, C zn, int w, int first_invalid,
                                  int *index_out)
{
    int ib = 0;
    const int step = ref_bla_get2(r_bla, stages_bla, zn, w, first_invalid, ib);
    *index_out = ib;
    return step;
}

/* Table entry of the integer form of that comparison: the high word of
 * fl(fl(r up)^2), or -2 when `|z| < r` can never hold (r = 0, negative or NaN:
 * the difference with any high word of s, which is >= 0, is then <= -2).  One int32 per BLA node and scale: the lookup then
 * costs an integer load, a subtraction and two compares per stage. */
#define FSB_R2HI_NEVER (-2)
FSB_HD int bla_r2hi(double r, double up)
{
    if (!(r > 0.)) return FSB_R2HI_NEVER;
    const double ru = mul_rn(r, up);
    return hi32(mul_rn(ru, ru));
}

/* high word of r for the component test of ref_bla_get4_ (0: never) */
FSB_HD int bla_rhi(double r)
{
    if (!(r > 0.)) return 0;                 /* r <= 0 or NaN */
    return hi32(r);                          /* +inf: above every finite |z| */
}

/* perturbation.py:2116-2170 with the lookup order of ref_bla_get (lowest
 * stored stage first) and the square-free comparison above, on the integer
 * tables.  The per-component pre-test `|re|, |im| < r3` of ref_bla_get is
 * implied by |z| < r3 and left out. */
FSB_HD int ref_bla_get3_(const FrameDev &f, C zn, int w, int &index_out);
FSB_HD int ref_bla_get4_(const FrameDev &f, C zn, int w, int &index_out);
FSB_HD int ref_bla_get3(const FrameDev &f, C zn, int w, int &index_out)
{
#if FSB_BLA_LOOKUP4
    return ref_bla_get4_(f, zn, w, index_out);
#endif
#if defined(FSB_DEBUG_BLA) && !defined(__CUDA_ARCH__)
    int i2 = -1, i3 = -1;
    const int s2 = ref_bla_get2(f.r_bla, f.stages_bla, zn, w, f.first_invalid_i, i2);
    const int s3 = ref_bla_get3_(f, zn, w, i3);
    if (s2 != s3 || (s2 && i2 != i3)) printf("BLA mismatch w %d z (%g, %g): %d/%d vs %d/%d\n", w, zn.re, zn.im, s2, i2, s3, i3);
    index_out = i3;
    return s3;
#else
    return ref_bla_get3_(f, zn, w, index_out);
#endif
}
#if defined(FSB_DEBUG_WALK) && !defined(__CUDA_ARCH__)
static long long g_walk_tests = 0, g_walk_lookups = 0, g_walk_pass3 = 0;
#endif
FSB_HD int ref_bla_get3_(const FrameDev &f, C zn, int w, int &index_out)
{
#if defined(FSB_DEBUG_WALK) && !defined(__CUDA_ARCH__)
    g_walk_lookups++;
#endif
    const int it = w >> 3;
    const int invalid_step = f.first_invalid_i - w;
    if (invalid_step <= 8 || f.stages_bla < 4) return 0;
    const int base = 2 * it - 1;
    const double a = fabs(zn.re), b = fabs(zn.im);
    const int ea = imax(expfield(a), expfield(b));
    if (ea > 1023 + 500) {
        int ib = 0;
        const int step = ref_bla_get2_cold(f.r_bla, f.stages_bla, zn, w, f.first_invalid_i, &ib);
        index_out = ib;
        return step;
    }
    const bool tiny = ea < 1023 - 500;
    const int *__restrict__ tab = tiny ? f.r2hi_up : f.r2hi;
    const double up = tiny ? 0x1p600 : 1.;
    const double au = mul_rn(a, up), bu = mul_rn(b, up);
    /* s >= 0 or NaN (sign stripped, clamped: above every entry, no overflow below) */
    const int hs = imin(hi32(add_rn(mul_rn(au, au), mul_rn(bu, bu))) & 0x7fffffff, 0x7ff80000);
    const int d3 = ldg_(tab + base + 1) - hs;
    if (d3 < 2 && (d3 <= -2 || !abs_lt_exact(zn, ldg_(f.r_bla + base + 1)))) return 0;
    int stages = f.stages_bla - 1;
    if (it != 0) stages = imin(stages, 3 + (ffs_(it) - 1));
    stages = imin(stages, 31 - clz_(invalid_step - 1));   /* largest stg with 2^stg < invalid_step */
#if defined(FSB_DEBUG_WALK) && !defined(__CUDA_ARCH__)
    g_walk_pass3++;
#endif
    for (int stg = stages; stg > 3; stg--) {
#if defined(FSB_DEBUG_WALK) && !defined(__CUDA_ARCH__)
        g_walk_tests++;
#endif
        const int ib = base + (1 << (stg - 3));
        const int d = ldg_(tab + ib) - hs;
        if (d >= 2 || (d > -2 && abs_lt_exact(zn, ldg_(f.r_bla + ib)))) { index_out = ib; return 1 << stg; }
    }
    index_out = base + 1;
    return 8;
}

/* The same lookup deciding most radius tests on the larger component alone.
 * With m = max(|re|, |im|):  m <= |z| <= sqrt(2) m, and the rounded hypot is never
 * below its larger argument.  On sign-stripped high words (hm of m, hr of r):
 *   hm > hr                 =>  m > r                  =>  not (|z| < r)
 *   hm + 0xA0000 < hr       =>  1.4545 m < r           =>  |z| < r, rounded or not
 * (adding 0xA0000 to a high word multiplies the value by at least 2 / 1.375 = 1.4545;
 * sqrt(2) m (1 + a few ulp) stays below that).  What falls in between -- a band of
 * ratio 1.45 -- takes the square test of ref_bla_get3_, its sum of squares formed once,
 * on demand.  Same decisions, bit for bit; rhi[] is 0 where |z| < r can never hold. */
FSB_HD int ref_bla_get4_(const FrameDev &f, C zn, int w, int &index_out)
{
    const int it = w >> 3;
    const int invalid_step = f.first_invalid_i - w;
    if (invalid_step <= 8 || f.stages_bla < 4) return 0;
    const int base = 2 * it - 1;
    const int hm = imax(hi32(zn.re) & 0x7fffffff, hi32(zn.im) & 0x7fffffff);
    if (hm > 0x7fe00000 - 0xA0000) {          /* huge, inf or NaN: the exact walk */
        int ib = 0;
        const int step = ref_bla_get2_cold(f.r_bla, f.stages_bla, zn, w, f.first_invalid_i, &ib);
        index_out = ib;
        return step;
    }
    const int hm_pass = hm + 0xA0000;
    /* the square test, prepared on first use */
    int hs = -1;
    const int *__restrict__ tab = f.r2hi;
    auto square_lt = [&](int ib) -> bool {
        if (hs < 0) {
            const double a = fabs(zn.re), b = fabs(zn.im);
            const bool tiny = imax(expfield(a), expfield(b)) < 1023 - 500;
            tab = tiny ? f.r2hi_up : f.r2hi;
            const double up = tiny ? 0x1p600 : 1.;
            const double au = mul_rn(a, up), bu = mul_rn(b, up);
            hs = imin(hi32(add_rn(mul_rn(au, au), mul_rn(bu, bu))) & 0x7fffffff, 0x7ff80000);
        }
        const int d = ldg_(tab + ib) - hs;
        return d >= 2 || (d > -2 && abs_lt_exact(zn, ldg_(f.r_bla + ib)));
    };
    const int h3 = ldg_(f.rhi + base + 1);
    if (hm > h3) return 0;
    if (!(hm_pass < h3) && !square_lt(base + 1)) return 0;
    int stages = f.stages_bla - 1;
    if (it != 0) stages = imin(stages, 3 + (ffs_(it) - 1));
    stages = imin(stages, 31 - clz_(invalid_step - 1));   /* largest stg with 2^stg < invalid_step */
    for (int stg = stages; stg > 3; stg--) {
        const int ib = base + (1 << (stg - 3));
        const int hr = ldg_(f.rhi + ib);
        if (hm > hr) continue;
        if (hm_pass < hr || square_lt(ib)) { index_out = ib; return 1 << stg; }
    }
    index_out = base + 1;
    return 8;
}

/* ======================================================================== */
/* Lane state machine of the persistent holomorphic kernel (k_perturb_m2_v2).
 *
 * The reference's per-pixel loop (perturbation.py:1076-1398) as an event-driven
 * machine.  A lane's life alternates between
 *   - the HOT iteration (`m2_hot_iter`): one full perturbation iteration on
 *     plain doubles with the cheapest sufficient tests -- no table lookup
 *     other than the orbit record, no flag, no counter;
 *   - the EVENT step (`lane_step`): everything else, entered when a pre-test
 *     of the hot iteration fires: exact stop tests, both rebase kinds, BLA
 *     lookups and steps, the first iteration after a rebase (sticky
 *     `bool_dyn_rebase`: reference derivative = 0), exact Xrange iterations,
 *     pixel epilogue and the initialisation of the lane's next pixel.
 * The pre-tests are necessary conditions evaluated on the high words
 * (integer pipe): `w >= wlim` (max_iter, reference about to diverge),
 * exponent of Z + z against the escape radius, |Z + z| <= |z| per component
 * for the dynamic rebase, and |z| per component against the stage-3 BLA
 * radius of the next index (carried by the orbit record, 0 where no lookup
 * takes place); whatever fires, the event step re-evaluates the reference's
 * tests exactly and in the reference's order, so the outputs do not depend on
 * the pre-tests.  n_iter is not stored: n_iter = nbase + w between events. */
enum : unsigned {
    LF_EV = 1u,        /* something to do in the event section                     */
    LF_ITER = 2u,      /* an iteration was performed: its stop / rebase tests are pending */
    LF_DYN = 4u,       /* bool_dyn_rebase (sticky, perturbation.py:1116,1317)      */
    LF_SLOW = 8u,      /* Xrange frames: state in Xrange form (not on the fp64 lane) */
    LF_NEED = 16u,     /* pixel finished, waits for the next one                   */
    LF_INIT = 32u,     /* a new pixel was assigned                                 */
    LF_DEAD = 64u,     /* no pixel (none left, or waiting for the whole warp)      */
    LF_BAD = 128u,     /* Xrange frames: the range guard of the hot loop failed    */
    LF_CAREFUL = 256u  /* Xrange frames: replaying from the checkpoint, one guarded
                        * iteration per event step, up to the failing one          */
};

/* Build knobs of the hot loop.  Measured on B200 (round 2, config 2 / config 3, ms per 4K
 * frame; default = 11.20 / 24.94):
 *   FSB_H3_DIRECT=1      pre-test word table indexed by w (zeros off the multiples of 8):
 *                        one unpredicated 4-byte load instead of predicate + shift +
 *                        predicated load.  Four instructions fewer per iteration on paper;
 *                        ptxas then breaks the two-register-set allocation of the loop
 *                        (10-16 moves per trip, tools/sass_hotloop.py): 11.59 / 25.19.
 *   FSB_HOT_RECOMPUTE=1  the loop carries only `ev | bad` to the vote and the flags are
 *                        recomputed after it: 11.52 / 24.82 (two SELs saved, three moves added).
 *   FSB_EVENT_PRETEST    the hot loop's necessary condition before every tree lookup of the
 *                        event section (a chain of BLA steps ends in a failed lookup):
 *                        11.28 / 24.79 -- inside the run-to-run noise.
 * A `mov` through volatile asm does not pin esc_hi in a register either: ptxas
 * rematerialises it from the constant bank all the same.  All off. */
#ifndef FSB_H3_DIRECT
#define FSB_H3_DIRECT 0
#endif
/* FSB_SHORT_TRIP=1: events that are only a BLA pre-test (three visits of lane_step in four)
 * handled next to the hot loop (lane_bla_short): same pixels and counters, 63 % fewer visits of
 * lane_step -- and slower (config 2 / 3: 10.88 / 23.64 ms against 10.73 / 23.10): the flag and
 * arming traffic it avoids is not what an event costs, the lookup and the step are, and a
 * second inlined copy of them is more code.  Off. */
#ifndef FSB_SHORT_TRIP
#define FSB_SHORT_TRIP 0
#endif

/* FSB_ZZ2 (default build only; the -fmad=false build keeps the literal operation order):
 * the hot loop carries C = 2 (Zn[w] + z) from one iteration to the next instead of Zn[w]:
 * 2 Z + z = C - z and 2 (Z + z) = C, two FP64 additions fewer per iteration (16 FP64
 * instructions instead of 18).  The orbit table then holds 2 Zn[w+1]; the pre-tests compare
 * DOUBLED high words (x + x drops the sign bit; the factor 2 of C is a constant offset folded
 * into the same instruction, into the escape bound and into the radius table: h3_word,
 * esc_word), and the escape bound takes max(a, b) -- the `a | b` shortcut of the plain form
 * needs both exponent fields below 0x400, and C passes 2 all the time.
 * Measured (B200, 4K frames, ms; plain form 11.20 / 24.93 on configs 2 / 3):
 *   1  C - z, one iteration per trip: 10.73 / 23.46  -- the default
 *   2  2 Z + z formed from the kept 2 Zn[w] exactly as in the plain form (same pixels as
 *      the plain form, two more registers, two iterations per trip): 11.37 / 24.19
 *   1 with two iterations per trip (FSB_V2_UNROLL=2): 10.88 / 23.65; 1 with the w-indexed
 *     pre-test table (FSB_H3_DIRECT=1): 11.53 / 24.37 (8 moves per trip), the two together:
 *     10.99 / 24.36 (39 instructions per iteration against 42, and still slower: the extra
 *     load per iteration costs more than the four instructions it saves)
 * With 1 the full-size parity rates of configs 2-5 are unchanged (100 % / 99.9992 % / 100 % /
 * 100 % identical stop_iter against the oracle); on the 2 304 pixels of the `p_M2_flake`
 * case 3 pixels differ by one iteration instead of 1 (the other rounding of 2 Z + z). */
#ifndef FSB_ZZ2
#ifdef FSB_STRICT
#define FSB_ZZ2 0
#else
#define FSB_ZZ2 1
#endif
#endif
#if defined(FSB_STRICT) && FSB_ZZ2
#error "FSB_ZZ2 re-associates the iteration: default build only"
#endif
#ifndef FSB_STAGE_TMA          /* orbit staging through shared memory by 1-D bulk copies (experiment, fsb_kernels.cuh) */
#define FSB_STAGE_TMA 0
#endif
#ifndef FSB_STAGE_WIN          /* records per staged window */
#define FSB_STAGE_WIN 64
#endif
#ifndef FSB_HOT_RECOMPUTE
#define FSB_HOT_RECOMPUTE 0
#endif
#if FSB_ZZ2
#define FSB_ZSCALE 2.
#else
#define FSB_ZSCALE 1.
#endif
/* the pre-test words as the hot loop compares them: high word hi of a radius, and the
 * escape bound */
FSB_HD unsigned h3_word(unsigned hi)
{
#if FSB_ZZ2
    return hi == 0u ? 0u : (hi >= 0x7fe00000u ? 0xffffffffu : 2u * hi + 0x200000u);
#else
    return hi;
#endif
}
FSB_HD unsigned esc_word(unsigned esc_hi)
{
#if FSB_ZZ2
    return esc_hi >= 0x7fe00000u ? 0xffffffffu : 2u * (esc_hi + 0x100000u);
#else
    return esc_hi;
#endif
}
/* slot of index w in the pre-test word table */
FSB_HD int h3_slot(int w) { return FSB_H3_DIRECT ? w : (w >> 3); }
#ifdef FSB_STRICT
#define FSB_TSCALE 1.
#else
#define FSB_TSCALE 2.
#endif
/* Xrange frames: the hot loop returns to the event section at least this often
 * (bounds the replay after a failed range guard) */
#ifndef FSB_XR_STRETCH
#define FSB_XR_STRETCH 128   /* measured on config 3: 32 / 64 / 128 / 256 -> 24.16 / 23.50 / 23.08 / 23.33 ms */
#endif

/* Rare paths of the event section.  Measured: making them real calls on the
 * device (__noinline__, the lane passed through a copy) puts the lane in local
 * memory -- 2.4 GB of DRAM writes per 4K frame, config 2 12.0 -> 15.9 ms -- so
 * they stay inlined unless FSB_OUTLINE_COLD is defined. */
#if defined(__CUDA_ARCH__) && defined(FSB_OUTLINE_COLD)
#define FSB_COLD static __device__ __noinline__
#define FSB_COLD_CALL(s, call) do { LaneM2 t_ = (s); call; (s) = t_; } while (0)
#else
#define FSB_COLD FSB_HD
#define FSB_COLD_CALL(s, call) do { LaneM2 &t_ = (s); call; } while (0)
#endif

struct LaneM2 {        /* the part of a lane that lives in registers */
    double zr, zi;     /* delta z: the value (fp64 lane) or the Xrange mantissa (LF_SLOW) */
    double dr, di;     /* delta z' likewise                                            */
    double cr, ci;     /* c, standard                                                  */
    double Zr, Zi;     /* Zn[w] (event section only)                                   */
    int w, wlim, winc; /* reference index; the hot loop leaves when w >= wlim; 0 for parked lanes */
    unsigned flags;
};
/* state at the entry of the hot loop (Xrange frames) */
struct LaneCk { double zr, zi, dr, di; int w; };
/* the part only the event section touches: shared memory on the device (the
 * register budget decides how many warps hide the latencies of the event
 * section: 90 -> 7x registers per thread) */
struct LaneCold {
    XC c_xr;
    int ze, de;        /* LF_SLOW: exponents of (zr, zi) and (dr, di)                   */
    int nbase;         /* n_iter - w                                                   */
    int ipt;
    unsigned p_skip, p_bla, p_reb, p_slow;
    int c_tiny;
    LaneCk ck;
};

/* one full iteration on doubles (mandelbrot_M2.py:607-622); (er, ei) is the
 * table's derivative entry FSB_TSCALE * dZndc[w] (0 after a rebase).
 * Default build: FMA chains, 2 (Z + z) formed as (2 Z + z) + z and the
 * doubling of dZndc folded into the table -- 18 FP64 instructions; the
 * reference's loops are numba fastmath (contracted and re-associated by LLVM),
 * so no particular rounding sequence is "the" reference.  -fmad=false build:
 * the literal operation order, bit-exact with the oracle. */
template <bool DZNDC>
FSB_HD void m2_iter_fp64(double zr, double zi, double dr, double di, double Zr, double Zi,
                         double er, double ei, double cr, double ci, double &nzr, double &nzi,
                         double &ndr, double &ndi)
{
#ifdef FSB_STRICT
    const C z = mkC(zr, zi), ref = mkC(Zr, Zi);
    if (DZNDC) {
        const C nd = p_iter_deriv(z, mkC(dr, di), ref, mkC(er, ei));
        ndr = nd.re; ndi = nd.im;
    }
    const C nz = p_iter_zn(z, ref, mkC(cr, ci));
    nzr = nz.re; nzi = nz.im;
#else
    const double tr = fma(2., Zr, zr), ti = fma(2., Zi, zi);
    if (DZNDC) {
        const double sr = tr + zr, si = ti + zi;
        ndr = fma(sr, dr, fma(-si, di, fma(er, zr, -(ei * zi))));
        ndi = fma(sr, di, fma(si, dr, fma(er, zi, ei * zr)));
    }
    nzr = fma(zr, tr, fma(-zi, ti, cr));
    nzi = fma(zr, ti, fma(zi, tr, ci));
#endif
}

FSB_HD void lane_park(LaneM2 &s, unsigned flags)
{
    s.zr = s.zi = s.dr = s.di = s.cr = s.ci = 0.;
    s.Zr = s.Zi = 0.;
    s.w = 0; s.winc = 0; s.wlim = 0x7fffffff;
    s.flags = flags;
}

/* pixel epilogue, perturbation.py:1374-1398.  Counters of the launch (executed
 * iterations, BLA steps, rebases, sum of stop_iter, fp64-lane iterations) go
 * to the caller's slots cnt[k * cstride]: one private slot set per thread. */
template <bool XR, bool DZNDC>
FSB_COLD void lane_finish(const FrameDev &f, LaneM2 &s, LaneCold &k, int stop, double *Z, int *U,
                        signed char *stop_reason, int *stop_iter, unsigned long long *cnt,
                        int cstride)
{
    const int w = s.w, ipt = k.ipt;
    U[ipt] = w;
    C zn, dz = mkC(0., 0.);
    const C Zw = ldC(f.Zn, w);
    if (XR && !(s.flags & LF_SLOW)) {
        zn = mkC(s.zr, s.zi) + Zw;
        if (DZNDC) {
            const C rd = ldC(f.dZndc_std, w);
            if (rd.re == rd.re && rd.im == rd.im) dz = mkC(s.dr, s.di) + rd;
            else dz = to_std(to_xr(mkC(s.dr, s.di)) + mkXC(ldC(f.dZndc, w), ldg_(f.dZndc_e + w)));
        }
    } else if (XR) {
        zn = to_std(mkXC(mkC(s.zr, s.zi), k.ze)) + Zw;
        if (DZNDC) dz = to_std(mkXC(mkC(s.dr, s.di), k.de) + mkXC(ldC(f.dZndc, w), ldg_(f.dZndc_e + w)));
    } else {
        zn = mkC(s.zr, s.zi) + Zw;
        if (DZNDC) dz = mkC(s.dr, s.di) + ldC(f.dZndc, w);
    }
    stC(Z, 0, f.zstride, ipt, zn);
    if (DZNDC) stC(Z, 1, f.zstride, ipt, dz);
    const int n_iter = k.nbase + w;
    stop_reason[ipt] = (signed char)stop;
    stop_iter[ipt] = n_iter;
    const unsigned p_exec = (unsigned)n_iter - k.p_skip;
    cnt[0] += p_exec;
    cnt[cstride] += k.p_bla;
    cnt[2 * cstride] += k.p_reb;
    cnt[3 * cstride] += (unsigned long long)n_iter;
    if (XR) cnt[4 * cstride] += p_exec - k.p_slow;
    lane_park(s, LF_NEED | LF_EV);
}

/* Xrange state <-> fp64 lane */
template <bool DZNDC> FSB_HD void lane_to_slow(LaneM2 &s, LaneCold &k)
{
    const XC zx = to_xr(mkC(s.zr, s.zi));
    s.zr = zx.m.re; s.zi = zx.m.im; k.ze = zx.e;
    if (DZNDC) {
        const XC dx = to_xr(mkC(s.dr, s.di));
        s.dr = dx.m.re; s.di = dx.m.im; k.de = dx.e;
    }
    s.flags = (s.flags | LF_SLOW) & ~LF_CAREFUL;
}
/* back to the fp64 lane when every live component is in the safe range;
 * zstd = to_std of the Xrange z */
template <bool DZNDC> FSB_HD void lane_try_fast(LaneM2 &s, LaneCold &k, C zstd)
{
    if (!in_fast_range(zstd)) return;
    C dstd = mkC(0., 0.);
    if (DZNDC) {
        dstd = to_std(mkXC(mkC(s.dr, s.di), k.de));
        if (!in_fast_range(dstd)) return;
    }
    s.zr = zstd.re; s.zi = zstd.im; k.ze = 0;
    if (DZNDC) { s.dr = dstd.re; s.di = dstd.im; k.de = 0; }
    s.flags &= ~LF_SLOW;
}

/* a new pixel, perturbation.py:1026-1031, 2214-2230 */
template <bool XR>
FSB_COLD void lane_init(const FrameDev &f, LaneM2 &s, LaneCold &k, const C *c_pix)
{
    const C pix = ldC(c_pix, k.ipt);
    const double x1 = f.lin_mat[0] * pix.re + f.lin_mat[1] * pix.im;
    const double y1 = f.lin_mat[2] * pix.re + f.lin_mat[3] * pix.im;
    k.c_xr = (mkXF(f.lin_scale, f.lin_scale_e) * mkC(x1, y1))
             + mkXC(mkC(f.drift[0], f.drift[1]), f.drift_e[0]);
    const C c = to_std(k.c_xr);
    s.cr = c.re; s.ci = c.im;
    /* |c| < 2^-1600: lets the fp64 lane of an Xrange frame take BLA steps */
    k.c_tiny = XR && (k.c_xr.e + cexp_field(k.c_xr.m) - 1023 < -1600);
    s.zr = s.zi = s.dr = s.di = 0.;
    k.ze = k.de = 0;
    s.w = 0; s.winc = 1; k.nbase = 0;
    const C Z0 = ldC(f.Zn, 0);
    s.Zr = Z0.re; s.Zi = Z0.im;
    k.p_skip = k.p_bla = k.p_reb = k.p_slow = 0;
    s.flags = LF_EV | LF_DYN | (XR ? LF_SLOW : 0u);
}

/* rebase: reference diverging (:1283-1313, `rebase`) or dynamic glitch
 * (:1317-1372); zn = the standard delta z, ZZ = zn + Zn[w] */
template <bool XR, bool DZNDC>
FSB_COLD void lane_rebase(const FrameDev &f, LaneM2 &s, LaneCold &k, C zn, C ZZ, bool rebase)
{
    const C *Zn = f.Zn;
    const bool has_xr = XR && f.n_xr_i > 0;
    const bool slow = XR && (s.flags & LF_SLOW);
    const C ref_next = mkC(s.Zr, s.Zi);
#define L_DZNDC_X(i) mkXC(ldC(f.dZndc, (i)), ldg_(f.dZndc_e + (i)))
#define L_REF_X(k) mkXC(ldC(f.ref_xr, (k)), ldg_(f.ref_xr_e + (k)))
#define L_LOAD_Z() do { const C Zw_ = ldC(Zn, s.w); s.Zr = Zw_.re; s.Zi = Zw_.im; } while (0)
    bool do_rebase = true, fast_rebase = false;
    XC ZZ_xr = mkXC(mkC(0., 0.), 0);
    if (XR && !rebase) {
        if (!slow && in_fast_range(ZZ)) {
            /* same comparison on the same correctly rounded values */
            do_rebase = norm2(ZZ) <= norm2(zn);
            fast_rebase = true;
        } else {
            if (!slow) lane_to_slow<DZNDC>(s, k);
            int knext = -1;
            if (has_xr && s.w != 0 && fabs(ref_next.re) < 1.e-300 && fabs(ref_next.im) < 1.e-300)
                knext = xr_find(f.ref_index_xr, f.n_xr_i, s.w);
            const XC zx = mkXC(mkC(s.zr, s.zi), k.ze);
            ZZ_xr = (knext >= 0) ? (zx + L_REF_X(knext)) : (zx + ref_next);
            do_rebase = xr_le(abs2(ZZ_xr), abs2(zx));
        }
    }
    if (do_rebase) {
        if (XR && !(s.flags & LF_SLOW) && (fast_rebase || rebase)) {
            /* rebase in plain fp64; leave the fp64 lane if a result falls
             * out of the safe range (exact conversion) */
            C nd = mkC(s.dr, s.di);
            if (DZNDC) nd = nd + ldC(f.dZndc_std, s.w);
            if (in_fast_range(ZZ) && (!DZNDC || in_fast_range(nd))) {
                s.zr = ZZ.re; s.zi = ZZ.im;
                if (DZNDC) { s.dr = nd.re; s.di = nd.im; }
            } else {
                if (DZNDC) {
                    const XC dx = to_xr(mkC(s.dr, s.di)) + L_DZNDC_X(s.w);
                    s.dr = dx.m.re; s.di = dx.m.im; k.de = dx.e;
                }
                const XC zx = to_xr(ZZ);
                s.zr = zx.m.re; s.zi = zx.m.im; k.ze = zx.e;
                s.flags = (s.flags | LF_SLOW) & ~LF_CAREFUL;
            }
        } else if (XR) {
            const XC zx = rebase ? to_xr(ZZ) : ZZ_xr;
            const C zstd = rebase ? ZZ : to_std(ZZ_xr);
            s.zr = zx.m.re; s.zi = zx.m.im; k.ze = zx.e;
            if (DZNDC) {
                const XC dx = mkXC(mkC(s.dr, s.di), k.de) + L_DZNDC_X(s.w);
                s.dr = dx.m.re; s.di = dx.m.im; k.de = dx.e;
            }
            lane_try_fast<DZNDC>(s, k, zstd);
        } else {
            s.zr = ZZ.re; s.zi = ZZ.im;
            if (DZNDC) {
                const C nd = mkC(s.dr, s.di) + ldC(f.dZndc, s.w);
                s.dr = nd.re; s.di = nd.im;
            }
        }
        k.nbase += s.w;
        s.w = 0;
        L_LOAD_Z();
        k.p_reb++;
    }

#undef L_DZNDC_X
#undef L_REF_X
#undef L_LOAD_Z
}

struct RebIO { double zr, zi, dr, di, Zr, Zi; int w, ze, de, nbase; unsigned flags, p_reb; };
#if defined(__CUDA_ARCH__) && defined(FSB_OUTLINE_REBASE)
#define FSB_VALUE_CALL2 static __device__ __noinline__
#else
#define FSB_VALUE_CALL2 FSB_HD
#endif
template <bool XR, bool DZNDC>
FSB_VALUE_CALL2 RebIO lane_rebase_v(const FrameDev &f, RebIO io, C zn, C ZZ, bool rebase)
{
    LaneM2 s;
    LaneCold k;
    s.zr = io.zr; s.zi = io.zi; s.dr = io.dr; s.di = io.di; s.Zr = io.Zr; s.Zi = io.Zi;
    s.cr = s.ci = 0.; s.w = io.w; s.wlim = 0; s.winc = 1; s.flags = io.flags;
    k.ze = io.ze; k.de = io.de; k.nbase = io.nbase; k.p_reb = io.p_reb;
    lane_rebase<XR, DZNDC>(f, s, k, zn, ZZ, rebase);
    io.zr = s.zr; io.zi = s.zi; io.dr = s.dr; io.di = s.di; io.Zr = s.Zr; io.Zi = s.Zi;
    io.w = s.w; io.ze = k.ze; io.de = k.de; io.nbase = k.nbase; io.flags = s.flags; io.p_reb = k.p_reb;
    return io;
}

/* one iteration by the operator chain of numba_xr.py (perturbation.py:1158-1209
 * in Xrange arithmetic): whatever the fused forms of lane_step do not cover.
 * Rare, and large once inlined.  FSB_OUTLINE_GENERIC / FSB_OUTLINE_REBASE make
 * this and the Xrange rebase real calls on the device, with everything passed
 * and returned BY VALUE so that the lane never has its address taken (no local
 * memory, 0 bytes of stack).  Measured on config 3: 24.9 ms inlined, 25.1 ms with
 * the generic iteration outlined, 26.5 ms with both -- the instruction cache is
 * not what bounds this kernel (stall_no_instruction 0.65) -- so both stay inlined. */
struct XC2 { XC z, d; };
#if defined(__CUDA_ARCH__) && defined(FSB_OUTLINE_GENERIC)
#define FSB_VALUE_CALL static __device__ __noinline__
#else
#define FSB_VALUE_CALL FSB_HD
#endif
template <bool DZNDC>
FSB_VALUE_CALL XC2 slow_generic_iter(XC zx, XC dx, XC ref_x, XC ref_d, XC c_xr)
{
    XC2 r;
    r.d = dx;
    if (DZNDC) r.d = p_iter_deriv(zx, dx, ref_x, ref_d);
    r.z = p_iter_zn(zx, ref_x, c_xr);
    return r;
}
template <bool DZNDC>
FSB_HD void lane_slow_generic(const FrameDev &f, LaneM2 &s, LaneCold &k, int kx, bool dyn)
{
    const C ref_zn = mkC(s.Zr, s.Zi);
    const XC ref_x = (kx >= 0) ? mkXC(ldC(f.ref_xr, kx), ldg_(f.ref_xr_e + kx)) : to_xr(ref_zn);
    XC ref_d = mkXC(mkC(0., 0.), 0);
    if (DZNDC && !dyn) ref_d = mkXC(ldC(f.dZndc, s.w), ldg_(f.dZndc_e + s.w));
    const XC2 r = slow_generic_iter<DZNDC>(mkXC(mkC(s.zr, s.zi), k.ze), mkXC(mkC(s.dr, s.di), k.de),
                                           ref_x, ref_d, k.c_xr);
    if (DZNDC) { s.dr = r.d.m.re; s.di = r.d.m.im; k.de = r.d.e; }
    s.zr = r.z.m.re; s.zi = r.z.m.im; k.ze = r.z.e;
}

/* The event section of one lane: runs until the lane is armed for the hot
 * loop (LF_EV cleared, wlim set) or its pixel has ended (LF_NEED). */
template <bool XR, bool DZNDC, bool BLA>
FSB_HD void lane_step(const FrameDev &f, LaneM2 &s, const C *c_pix, double *Z, int *U,
                      signed char *stop_reason, int *stop_iter, unsigned long long *cnt,
                      int cstride, LaneCold &k)
{
    const C *Zn = f.Zn;
    const bool has_xr = XR && f.n_xr_i > 0;
#define L_DZNDC_X(i) mkXC(ldC(f.dZndc, (i)), ldg_(f.dZndc_e + (i)))
#define L_REF_X(k) mkXC(ldC(f.ref_xr, (k)), ldg_(f.ref_xr_e + (k)))
#define L_LOAD_Z() do { const C Zw_ = ldC(Zn, s.w); s.Zr = Zw_.re; s.Zi = Zw_.im; } while (0)

    if (s.flags & LF_INIT) {
        FSB_COLD_CALL(s, (lane_init<XR>(f, t_, k, c_pix)));
    } else if (XR && (s.flags & LF_BAD)) {
        /* the range guard failed somewhere after the checkpoint: back to it, then
         * one guarded iteration at a time until the failing one is reached */
        s.zr = k.ck.zr; s.zi = k.ck.zi; s.dr = k.ck.dr; s.di = k.ck.di; s.w = k.ck.w;
        L_LOAD_Z();
        s.flags = (s.flags & ~LF_BAD) | LF_CAREFUL;
    }

    /* zc: the standard value of delta z, kept along (a lane enters on the fp64
     * lane, or with z = 0) */
    C zc = mkC(s.zr, s.zi);
    for (;;) {
        if (s.flags & LF_ITER) {
            s.flags &= ~LF_ITER;
            /* ---- stop tests of the iteration just done, :1218-1279 ---- */
            if (k.nbase + s.w >= f.max_iter_i) {
                FSB_COLD_CALL(s, (lane_finish<XR, DZNDC>(f, t_, k, 0, Z, U, stop_reason, stop_iter, cnt, cstride)));
                return;
            }
            const C zn = zc;
            const C ref_next = mkC(s.Zr, s.Zi);
            const C ZZ = zn + ref_next;
            if (norm2(ZZ) > f.Mdiv_sq) {
                FSB_COLD_CALL(s, (lane_finish<XR, DZNDC>(f, t_, k, 1, Z, U, stop_reason, stop_iter, cnt, cstride)));
                return;
            }
            /* ---- rebase: reference diverging (:1283-1313) or dynamic glitch
             * (:1317-1372); only the dynamic test assigns the sticky flag ---- */
            const bool rebase = (s.w >= f.ref_div_m1_i);
            bool go = rebase;
            if (!rebase) {
                const bool dyn = (fabs(ZZ.re) <= fabs(zn.re)) && (fabs(ZZ.im) <= fabs(zn.im));
                s.flags = dyn ? (s.flags | LF_DYN) : (s.flags & ~LF_DYN);
                go = dyn;
            }
            if (go) {
#ifdef FSB_OUTLINE_REBASE
                if (XR) {       /* Xrange frames: a real call, state by value (see slow_generic_iter) */
                    RebIO io;
                    io.zr = s.zr; io.zi = s.zi; io.dr = s.dr; io.di = s.di; io.Zr = s.Zr; io.Zi = s.Zi;
                    io.w = s.w; io.ze = k.ze; io.de = k.de; io.nbase = k.nbase; io.flags = s.flags;
                    io.p_reb = k.p_reb;
                    io = lane_rebase_v<XR, DZNDC>(f, io, zn, ZZ, rebase);
                    s.zr = io.zr; s.zi = io.zi; s.dr = io.dr; s.di = io.di; s.Zr = io.Zr; s.Zi = io.Zi;
                    s.w = io.w; k.ze = io.ze; k.de = io.de; k.nbase = io.nbase; s.flags = io.flags;
                    k.p_reb = io.p_reb;
                } else
#endif
                FSB_COLD_CALL(s, (lane_rebase<XR, DZNDC>(f, t_, k, zn, ZZ, rebase)));
                zc = (XR && (s.flags & LF_SLOW)) ? to_std_small(mkXC(mkC(s.zr, s.zi), k.ze)) : mkC(s.zr, s.zi);
            }
        }

        /* ---- BLA steps, perturbation.py:1121-1154: no stop test in between ---- */
        if (BLA) {
            while ((s.w & 7) == 0) {
                const bool slow = XR && (s.flags & LF_SLOW);
                const C zn = zc;
                int ib = 0;
#if defined(FSB_EVENT_PRETEST) && !FSB_ZZ2
                {   /* the hot loop's necessary condition first: most chains of steps end here */
                    const unsigned h3 = ldg_(f.h3 + h3_slot(s.w));
                    const unsigned c = (unsigned)hi32(zn.re) & 0x7fffffffu, d = (unsigned)hi32(zn.im) & 0x7fffffffu;
                    if (!((c <= h3) & (d <= h3))) break;
                }
#endif
#ifdef FSB_BLA_LOOKUP2
                const int step = ref_bla_get2(f.r_bla, f.stages_bla, zn, s.w, f.first_invalid_i, ib);
#else
                const int step = ref_bla_get3(f, zn, s.w, ib);
#endif
                if (step == 0) break;
                const C *M = reinterpret_cast<const C *>(f.M_bla);
                const C A = ldC(M, 2 * ib), B = ldC(M, 2 * ib + 1);
                k.p_skip += (unsigned)step;
                s.w += step;                       /* n_iter = nbase + w moves with it */
                L_LOAD_Z();
                k.p_bla++;
                if (!XR) {
                    const C nz = A * zn + B * mkC(s.cr, s.ci);
                    s.zr = nz.re; s.zi = nz.im; zc = nz;
                    if (DZNDC) { const C nd = A * mkC(s.dr, s.di); s.dr = nd.re; s.di = nd.im; }
                    continue;
                }
                if (!slow && k.c_tiny) {
                    /* fp64 form of the step.  With every component of A z a normal
                     * double >= 2^-460 and |B c| <= 2^1025 |c| < 2^-575, the B c term is
                     * far below half an ulp of the sums it enters: fl(A z) IS the
                     * correctly rounded Xrange result.  Out-of-range results or a
                     * non-finite B: exact path below. */
                    const C nz = A * zn;
                    C nd = mkC(s.dr, s.di);
                    if (DZNDC) nd = A * nd;
#ifdef FSB_SC_GUARDS
                    if (in_fast_range(nz) && (!DZNDC || in_fast_range(nd))
                        && expfield(B.re) != 0x7ff && expfield(B.im) != 0x7ff) {
#else
                    if (in_fast_range(nz) & (!DZNDC || in_fast_range(nd))
                        & (imax(expfield(B.re), expfield(B.im)) != 0x7ff)) {
#endif
                        s.zr = nz.re; s.zi = nz.im; zc = nz;
                        if (DZNDC) { s.dr = nd.re; s.di = nd.im; }
                        continue;
                    }
                }
                if (!slow) lane_to_slow<DZNDC>(s, k);      /* B * c needs the exact c */
                const XC zx = xr_lin(A, mkXC(mkC(s.zr, s.zi), k.ze), B, k.c_xr);
                s.zr = zx.m.re; s.zi = zx.m.im; k.ze = zx.e;
                if (DZNDC) {
                    const XC dx = xr_mulc(A, mkXC(mkC(s.dr, s.di), k.de));
                    s.dr = dx.m.re; s.di = dx.m.im; k.de = dx.e;
                }
                zc = to_std_small(zx);
                lane_try_fast<DZNDC>(s, k, zc);
            }
        }

        if (!(s.flags & (LF_DYN | LF_SLOW | LF_CAREFUL))) break;      /* -> arm the hot loop */

        /* ---- one full iteration here: reference derivative 0 after a rebase
         * (sticky flag), exact Xrange arithmetic, or a guarded replay, :1158-1209 ---- */
        const bool dyn = (s.flags & LF_DYN) != 0;
        if (XR && (s.flags & LF_SLOW)) {
            const C ref_zn = mkC(s.Zr, s.Zi);
            int kx = -1;
            if (has_xr && s.w != 0 && fabs(ref_zn.re) < 1.e-300 && fabs(ref_zn.im) < 1.e-300)
                kx = xr_find(f.ref_index_xr, f.n_xr_i, s.w);
            XC zx = mkXC(mkC(s.zr, s.zi), k.ze);
            /* Fused exact forms of the two model formulas for a tiny z (the
             * iterations that follow a rebase onto a sub-1e-138 value, before the first
             * BLA node applies -- 6 % of the iterations of a 1e-1000 frame but, through
             * the operator chain, a third of its instructions):
             *   form 1  |z| < 2^-60 of either part of Z (both normal): z + 2 Z and Z + z
             *           round to 2 Z and Z, so z' = (2 Z) z + c, z'' = 2 (Z dz + Zd z);
             *   form 2  Z = 0 (index 0): z' = z z + c, z'' = 2 (z dz + Zd z).
             * Same products, same aligned additions and same roundings as the chain
             * of numba_xr operators (only the normalisations in between are skipped:
             * an Xrange value does not depend on its mantissa / exponent split). */
            int form = 0;
#ifndef FSB_NO_FUSED_ITER
            if (kx < 0) {
                const int fzr = expfield(ref_zn.re), fzi = expfield(ref_zn.im);
                const int Ez = k.ze + imax(expfield(s.zr), expfield(s.zi)) - 1023;
                if (ref_zn.re == 0. && ref_zn.im == 0.) form = 2;
                else if (fzr >= 64 && fzi >= 64 && fzr < 1600 && fzi < 1600
                         && fzr - fzi <= 500 && fzi - fzr <= 500
                         && Ez + 60 <= imin(fzr, fzi) - 1023) form = 1;
            }
#endif
            if (form != 0) {
                const C zm = zx.m;
                const C Az = (form == 1) ? mkC(2. * ref_zn.re, 2. * ref_zn.im) : zm;
                const C Ad = (form == 1) ? ref_zn : zm;
                if (DZNDC) {
                    const XC D = dyn ? mkXC(mkC(0., 0.), 0) : L_DZNDC_X(s.w);
                    XC dx = xr_sum2(Ad * mkC(s.dr, s.di), (form == 1) ? k.de : k.ze + k.de,
                                    D.m * zm, D.e + k.ze);
                    dx.e += 1;                                   /* 2. * ( ... ) */
                    s.dr = dx.m.re; s.di = dx.m.im; k.de = dx.e;
                }
                zx = xr_sum2(Az * zm, (form == 1) ? k.ze : 2 * k.ze, k.c_xr.m, k.c_xr.e);
                s.zr = zx.m.re; s.zi = zx.m.im; k.ze = zx.e;
            } else {
                FSB_COLD_CALL(s, (lane_slow_generic<DZNDC>(f, t_, k, kx, dyn)));
                zx = mkXC(mkC(s.zr, s.zi), k.ze);
            }
            k.p_slow++;
            zc = to_std_small(zx);
            lane_try_fast<DZNDC>(s, k, zc);
        } else {
            double er = 0., ei = 0.;
            if (DZNDC && !dyn) {
                const C rd = ldC(XR ? f.dZndc_std : f.dZndc, s.w);
                er = mul_rn(FSB_TSCALE, rd.re); ei = mul_rn(FSB_TSCALE, rd.im);
            }
            double nzr, nzi, ndr = 0., ndi = 0.;
            m2_iter_fp64<DZNDC>(s.zr, s.zi, s.dr, s.di, s.Zr, s.Zi, er, ei, s.cr, s.ci,
                                nzr, nzi, ndr, ndi);
            if (XR && !(in_fast_range(mkC(nzr, nzi)) && (!DZNDC || in_fast_range(mkC(ndr, ndi))))) {
                /* out of the safe range: redo this iteration in Xrange arithmetic */
                lane_to_slow<DZNDC>(s, k);
                continue;
            }
            s.zr = nzr; s.zi = nzi; zc = mkC(nzr, nzi);
            if (DZNDC) { s.dr = ndr; s.di = ndi; }
        }
        s.w += 1;
        L_LOAD_Z();
        s.flags |= LF_ITER;
        /* (a replay always meets its failing iteration; this bound is a safety net) */
        if (XR && (s.flags & LF_CAREFUL) && s.w > k.ck.w + FSB_XR_STRETCH) s.flags &= ~LF_CAREFUL;
    }

    /* ---- arm the hot loop ---- */
    int wl = imin(f.ref_div_m1_i, f.max_iter_i - k.nbase);
#if FSB_STAGE_TMA
    wl = imin(wl, s.w + FSB_STAGE_WIN);      /* the staged loop reads one window of the orbit */
#endif
    if (XR) {
        wl = imin(wl, s.w + FSB_XR_STRETCH);
        k.ck.zr = s.zr; k.ck.zi = s.zi; k.ck.dr = s.dr; k.ck.di = s.di; k.ck.w = s.w;
    }
    s.wlim = wl;
    s.flags &= ~LF_EV;
#if defined(__CUDA_ARCH__) && defined(FSB_PREFETCH_BLA)
    /* the tree nodes the next lookup (at the next multiple of 8) reads first: its stage-3
     * entry of the integer radius table and that node's (A, B) */
    if (BLA && f.stages_bla >= 4) {
        const int wn = (s.w + 8) & ~7;
        if (f.first_invalid_i - wn > 8) {
            const int ib = 2 * (wn >> 3);
            asm volatile("prefetch.global.L1 [%0];" :: "l"(f.r2hi + ib));
            asm volatile("prefetch.global.L1 [%0];" :: "l"(f.M_bla + 4 * ib));
#if FSB_PREFETCH_BLA > 1
            asm volatile("prefetch.global.L1 [%0];" :: "l"(f.r_bla + ib));
#endif
        }
    }
#endif
#undef L_DZNDC_X
#undef L_REF_X
#undef L_LOAD_Z
}

/* The hot iteration with its pre-tests.  `ev`: the lane must visit the event
 * section for the tests of this iteration; `bad` (Xrange frames): the range
 * guard failed -- the state is then garbage and the lane goes back to its
 * checkpoint.  (Zr, Zi) = Zn[w], carried from the previous record; the orbit
 * record of index w is {Zn[w+1] = (t0, t1), FSB_TSCALE dZndc[w] = (t2, t3)}. */
/* the pre-tests on the state AFTER an iteration: (Zr, Zi) = Zn[w] */
template <bool XR, bool DZNDC, bool BLA>
FSB_HD void m2_hot_flags(const LaneM2 &s, double Zr, double Zi, const unsigned *__restrict__ h3tab,
                         unsigned esc_hi, bool &ev, bool &bad)
{
    const double ZZr = s.zr + Zr, ZZi = s.zi + Zi;
    const unsigned a = (unsigned)hi32(ZZr) & 0x7fffffffu, b = (unsigned)hi32(ZZi) & 0x7fffffffu;
    const unsigned c = (unsigned)hi32(s.zr) & 0x7fffffffu, d = (unsigned)hi32(s.zi) & 0x7fffffffu;
    /* |Z + z| <= |z| and |z| < r3 per component can only hold if they hold for the
     * sign-stripped high words; a | b bounds both exponent fields from above */
    ev = (s.w >= s.wlim) | ((a | b) >= esc_hi) | ((a <= c) & (b <= d));
    if (BLA) {
        /* the tree is looked up at multiples of 8: h3tab[w] = high word of the stage-3
         * radius of index w there, 0 elsewhere and where no lookup takes place */
#if FSB_H3_DIRECT
        const unsigned h3 = ldg_(h3tab + s.w);
#else
        unsigned h3 = 0u;
        if ((s.w & 7) == 0) h3 = ldg_(h3tab + (s.w >> 3));
#endif
        ev = ev | ((c <= h3) & (d <= h3));
    }
    bad = false;
    if (XR) {
        bad = !(in_fast_range(s.zr) & in_fast_range(s.zi));
        if (DZNDC) bad = bad | !(in_fast_range(s.dr) & in_fast_range(s.di));
    }
}
template <bool XR, bool DZNDC, bool BLA>
FSB_HD void m2_hot_iter(LaneM2 &s, double Zr, double Zi, double t2, double t3)
{
    double nzr, nzi, ndr = 0., ndi = 0.;
    m2_iter_fp64<DZNDC>(s.zr, s.zi, s.dr, s.di, Zr, Zi, t2, t3, s.cr, s.ci, nzr, nzi, ndr, ndi);
    s.zr = nzr; s.zi = nzi;
    if (DZNDC) { s.dr = ndr; s.di = ndi; }
    s.w += s.winc;
}
#if FSB_ZZ2
/* The same two functions with C = 2 (Zn[w] + z) carried.  hot_enter / hot_exit convert
 * between the carried value and Zn[w] at the two ends of the loop. */
FSB_HD void m2_hot_enter(const LaneM2 &s, double &Cr, double &Ci)
{
    Cr = 2. * (s.Zr + s.zr); Ci = 2. * (s.Zi + s.zi);
}
/* (Z2r, Z2i) = 2 Zn[w] (FSB_ZZ2 == 2: kept from the previous record so that 2 Z + z is
 * formed exactly as in the plain form; FSB_ZZ2 == 1: unused, 2 Z + z = C - z) */
template <bool XR, bool DZNDC, bool BLA>
FSB_HD void m2_hot_iter_c(LaneM2 &s, double Cr, double Ci, double Z2r, double Z2i, double t2, double t3)
{
    const double zr = s.zr, zi = s.zi, dr = s.dr, di = s.di;
#if FSB_ZZ2 == 2
    const double tr = Z2r + zr, ti = Z2i + zi;            /* fl(2 Z + z) */
#else
    const double tr = Cr - zr, ti = Ci - zi;              /* 2 Z + z */
    (void)Z2r; (void)Z2i;
#endif
    if (DZNDC) {
        s.dr = fma(Cr, dr, fma(-Ci, di, fma(t2, zr, -(t3 * zi))));
        s.di = fma(Cr, di, fma(Ci, dr, fma(t2, zi, t3 * zr)));
    }
    s.zr = fma(zr, tr, fma(-zi, ti, s.cr));
    s.zi = fma(zr, ti, fma(zi, tr, s.ci));
    s.w += s.winc;
}
/* (t0, t1) = 2 Zn[w] of the new index: next carried value and the pre-tests */
template <bool XR, bool DZNDC, bool BLA>
FSB_HD void m2_hot_flags_c(const LaneM2 &s, double t0, double t1, const unsigned *__restrict__ h3tab,
                           unsigned esc2, bool &ev, bool &bad, double &Cr, double &Ci)
{
    Cr = fma(2., s.zr, t0); Ci = fma(2., s.zi, t1);
    const unsigned hr = (unsigned)hi32(Cr), hi_ = (unsigned)hi32(Ci);
    const unsigned zr_ = (unsigned)hi32(s.zr), zi_ = (unsigned)hi32(s.zi);
    const unsigned a = hr + hr, b = hi_ + hi_;                        /* 2 |hi(2 (Z + z))| */
    const unsigned c = zr_ + zr_ + 0x200000u, d = zi_ + zi_ + 0x200000u;   /* 2 |hi(2 z)| */
    /* (the a | b shortcut of the plain form needs both fields below 0x400: C = 2 (Z + z)
     * passes 2 all the time) */
    ev = (s.w >= s.wlim) | ((a > b ? a : b) >= esc2) | ((a <= c) & (b <= d));
#if defined(FSB_DEBUG_EV) && !defined(__CUDA_ARCH__)
    { static int n_ = 0; if (n_++ < 12) fprintf(stderr, "ev w %d wlim %d a %08x b %08x c %08x d %08x esc %08x | %d %d %d  C (%g, %g) z (%g, %g)\n", s.w, s.wlim, a, b, c, d, esc2, s.w >= s.wlim, (a | b) >= esc2, (a <= c) & (b <= d), Cr, Ci, s.zr, s.zi); }
#endif
    if (BLA) {
#if FSB_H3_DIRECT
        const unsigned h3 = ldg_(h3tab + s.w);
#else
        unsigned h3 = 0u;
        if ((s.w & 7) == 0) h3 = ldg_(h3tab + (s.w >> 3));
#endif
        ev = ev | ((c <= h3) & (d <= h3));
    }
    bad = false;
    if (XR) {
        bad = !(in_fast_range(s.zr) & in_fast_range(s.zi));
        if (DZNDC) bad = bad | !(in_fast_range(s.dr) & in_fast_range(s.di));
    }
}
#endif

/* ---- the short trip ----------------------------------------------------------
 * Three event visits in four are a lane at a multiple of 8 whose BLA pre-test fired
 * and nothing else: no stop test can fire (checked here with the exact tests of the
 * iteration just done, not their pre-tests), the lane is on the fp64 lane.  Those
 * lanes take their BLA steps right here -- lookup, (A, B), `A z [+ B c]`, guard -- and go
 * back to the hot loop without the flag / arming / cold-state traffic of lane_step.
 * Anything else (another test true, a guard that fails, a pixel whose c is not tiny in
 * an Xrange frame) leaves the lane to lane_step, state untouched or consistent at a
 * multiple of 8 (`false`: the caller sets LF_EV without LF_ITER -- the tests of the
 * iteration are known to be false -- and lane_step resumes at its BLA loop). */
template <bool XR>
FSB_HD bool m2_trip_is_short(const FrameDev &f, const LaneM2 &s, double Zr, double Zi)
{
    const double ZZr = s.zr + Zr, ZZi = s.zi + Zi;
    const bool stop = (s.w >= s.wlim) | (ZZr * ZZr + ZZi * ZZi > f.Mdiv_sq);
    const bool dyn = (fabs(ZZr) <= fabs(s.zr)) & (fabs(ZZi) <= fabs(s.zi));
    return !(stop | dyn);
}
template <bool XR, bool DZNDC>
FSB_HD bool lane_bla_short(const FrameDev &f, LaneM2 &s, LaneCold &k)
{
    if (XR && !k.c_tiny) return false;
    while ((s.w & 7) == 0) {
        const C zn = mkC(s.zr, s.zi);
        int ib = 0;
        const int step = ref_bla_get3(f, zn, s.w, ib);
        if (step == 0) break;
        const C *M = reinterpret_cast<const C *>(f.M_bla);
        const C A = ldC(M, 2 * ib), B = ldC(M, 2 * ib + 1);
        C nz, nd = mkC(s.dr, s.di);
        if (!XR) {
            nz = A * zn + B * mkC(s.cr, s.ci);
            if (DZNDC) nd = A * nd;
        } else {
            nz = A * zn;                       /* B c is below half an ulp: see lane_step */
            if (DZNDC) nd = A * nd;
            if (!(in_fast_range(nz) & (!DZNDC || in_fast_range(nd))
                  & (imax(expfield(B.re), expfield(B.im)) != 0x7ff))) return false;
        }
        s.zr = nz.re; s.zi = nz.im;
        if (DZNDC) { s.dr = nd.re; s.di = nd.im; }
        k.p_skip += (unsigned)step;
        k.p_bla++;
        s.w += step;
        const C Zw = ldC(f.Zn, s.w);
        s.Zr = Zw.re; s.Zi = Zw.im;
    }
    if (XR) {                                  /* new checkpoint, as when lane_step arms the loop */
        s.wlim = imin(imin(f.ref_div_m1_i, f.max_iter_i - k.nbase), s.w + FSB_XR_STRETCH);
        k.ck.zr = s.zr; k.ck.zi = s.zi; k.ck.dr = s.dr; k.ck.di = s.di; k.ck.w = s.w;
    }
    return true;
}

/* exponent-field bound of the escape pre-test: both parts of Z + z below 2^k
 * imply |Z + z|^2 < 2^(2k+1) <= Mdiv_sq */
inline unsigned esc_hi_of(double Mdiv_sq)
{
    if (!(Mdiv_sq > 0.)) return 0u;                     /* every iteration takes the exact test */
    if (Mdiv_sq > 1.7e308) return 0x7ff00000u;
    int e = 0;
    frexp(Mdiv_sq, &e);                                 /* Mdiv_sq in [2^(e-1), 2^e) */
    int k = (e - 2) / 2;                                /* 2k + 1 <= e - 1 */
    if (e - 2 < 0) k = -((2 - e + 1) / 2);
    if (k < -1000) return 0u;
    return (unsigned)(k + 1023) << 20;
}

} /* namespace fsb */
