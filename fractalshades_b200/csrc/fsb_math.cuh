/*
 * fsb_math.cuh -- scalar arithmetic shared by the device kernels and the host
 * table builders of libfsb200: complex128 with numba's lowering, and the
 * Xrange "extended range" scalars (fp64 mantissa + int32 base-2 exponent).
 *
 * Xrange semantics follow the reference's numba overloads
 * (src/fractalshades/numpy_utils/numba_xr.py): lazy renormalisation when any
 * part's |exponent| > 100 (:390-411), _frexp to [1,2) by bit surgery
 * (:674-685), _exp2_shift on the exponent field clamped at 0 (:696-706),
 * co-exponent alignment (:708-756), to_standard (:802-829).
 *
 * Operators that must give the same bits in the default (FMA-contracting) and
 * the -fmad=false build use the explicit _rn intrinsics.
 */
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#define FSB_HD __host__ __device__ __forceinline__
/* Real Xrange operators are real calls on the device: the burning-ship Xrange
 * code chains some eighty of them per iteration, and inlined they put the
 * kernel far beyond the instruction cache (it then waits on instruction fetch
 * two cycles out of three). */
#ifdef __CUDA_ARCH__
#define FSB_XF_OP static __device__ __noinline__
#else
#define FSB_XF_OP static inline
#endif

namespace fsb {

/* ---- bit access ---------------------------------------------------------- */
FSB_HD int hi32(double x)
{
#ifdef __CUDA_ARCH__
    return __double2hiint(x);
#else
    int64_t b; memcpy(&b, &x, 8); return (int)(b >> 32);
#endif
}
FSB_HD int lo32(double x)
{
#ifdef __CUDA_ARCH__
    return __double2loint(x);
#else
    int64_t b; memcpy(&b, &x, 8); return (int)(b & 0xffffffffLL);
#endif
}
FSB_HD double mk64(int hi, int lo)
{
#ifdef __CUDA_ARCH__
    return __hiloint2double(hi, lo);
#else
    int64_t b = ((int64_t)hi << 32) | (int64_t)(uint32_t)lo;
    double x; memcpy(&x, &b, 8); return x;
#endif
}
FSB_HD int expfield(double m) { return (hi32(m) >> 20) & 0x7ff; }

/* individually rounded mul / add, immune to FMA contraction */
FSB_HD double mul_rn(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    volatile double r = a * b; return r;
#endif
}
FSB_HD double add_rn(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}

/* numba_xr.py:390-411 */
FSB_HD bool need_renorm(double m)
{
    return (unsigned)(expfield(m) - (1023 - 100)) > 200u;
}
/* numba_xr.py:674-685 */
FSB_HD void xr_frexp(double m, double &nm, int &ne)
{
    int hi = hi32(m);
    ne = ((hi >> 20) & 0x7ff) - 1023;
    nm = mk64((hi & (int)0x800fffff) | 0x3ff00000, lo32(m));
}
/* numba_xr.py:687-694 */
FSB_HD void normalize_real(double m, int e, double &nm, int &ne)
{
    if (m == 0.) { nm = m; ne = 0; return; }
    int k;
    xr_frexp(m, nm, k);
    ne = e + k;
}
/* numba_xr.py:696-706 ; shift <= 0 in every use */
FSB_HD double exp2_shift(double m, int shift)
{
    int hi = hi32(m);
    if (shift < -4096) shift = -4096;
    int e = ((hi >> 20) & 0x7ff) + shift;
    if (e < 0) e = 0;
    return mk64((hi & (int)0x800fffff) | (e << 20), lo32(m));
}
FSB_HD double pymax(double a, double b) { return (b > a) ? b : a; }
FSB_HD double pymin(double a, double b) { return (b < a) ? b : a; }

/* exact 2**e (np.ldexp(1., e)) */
FSB_HD double ldexp1(int e)
{
    if (e > 1023) return mk64(0x7ff00000, 0);
    if (e >= -1022) return mk64((e + 1023) << 20, 0);
    if (e >= -1074) {
        int s = e + 1074;
        return (s >= 32) ? mk64(1 << (s - 32), 0) : mk64(0, (int)(1u << s));
    }
    return 0.;
}

/* ---- complex128, numba lowering ------------------------------------------ */
struct C { double re, im; };
FSB_HD C mkC(double r, double i) { C c; c.re = r; c.im = i; return c; }
FSB_HD C operator+(C a, C b) { return mkC(a.re + b.re, a.im + b.im); }
FSB_HD C operator-(C a, C b) { return mkC(a.re - b.re, a.im - b.im); }
FSB_HD C operator*(C a, C b)
{
    return mkC(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
FSB_HD C operator*(double s, C a) { return mkC(s * a.re, s * a.im); }
FSB_HD C operator+(C a, double b) { return mkC(a.re + b, a.im); }
FSB_HD bool is0(C a) { return a.re == 0. && a.im == 0.; }
FSB_HD double norm2(C a) { return a.re * a.re + a.im * a.im; }
/* contraction-proof complex product / sum for the table builders */
FSB_HD C cmul_rn(C a, C b)
{
    return mkC(add_rn(mul_rn(a.re, b.re), -mul_rn(a.im, b.im)),
               add_rn(mul_rn(a.re, b.im), mul_rn(a.im, b.re)));
}
FSB_HD C cadd_rn(C a, C b) { return mkC(add_rn(a.re, b.re), add_rn(a.im, b.im)); }
FSB_HD C scale2_rn(C a) { return mkC(mul_rn(2., a.re), mul_rn(2., a.im)); }
FSB_HD double norm2_rn(C a) { return add_rn(mul_rn(a.re, a.re), mul_rn(a.im, a.im)); }

/* |x + iy| : same definition as the oracle's fso_hypot (oracle/fs_oracle.cpp):
 * power-of-two pre-scaling then sqrt(a*a + b*b), each operation rounded once. */
FSB_HD double hypot_rn(double x, double y)
{
    double a = fabs(x), b = fabs(y);
    if (a < b) { double t = a; a = b; b = t; }
    if (!(a == a) || !(b == b)) return mk64(0x7ff80000, 0);
    if (a == 0.) return 0.;
    if (a > 1.7976931348623157e308) return a;
    int ea = expfield(a);
    double up = 1., down = 1.;
    if (ea > 1023 + 500) { up = 0x1p-600; down = 0x1p600; }
    else if (ea < 1023 - 500) { up = 0x1p600; down = 0x1p-600; }
    a = mul_rn(a, up);
    b = mul_rn(b, up);
    double s = add_rn(mul_rn(a, a), mul_rn(b, b));
    return mul_rn(sqrt(s), down);
}
FSB_HD double cabs_rn(C z) { return hypot_rn(z.re, z.im); }

/* ---- Xrange scalars ------------------------------------------------------- */
struct XF { double m; int e; };
struct XC { C m; int e; };
FSB_HD XF mkXF(double m, int e) { XF x; x.m = m; x.e = e; return x; }
FSB_HD XC mkXC(C m, int e) { XC x; x.m = m; x.e = e; return x; }

/* numba_xr.py:716-733 */
FSB_HD void coexp_f(double m0, int e0, double m1, int e1, double &o0, double &o1, int &oe)
{
    o0 = m0; o1 = m1;
    int d = e0 - e1;
    if (m0 == 0.) oe = e1;
    else if (m1 == 0.) oe = e0;
    else if (e1 > e0) { o0 = exp2_shift(m0, d); oe = e1; }
    else if (e0 > e1) { o1 = exp2_shift(m1, -d); oe = e0; }
    else oe = e0;
}
/* numba_xr.py:735-754 */
FSB_HD void coexp_c(C m0, int e0, C m1, int e1, C &o0, C &o1, int &oe)
{
    o0 = m0; o1 = m1;
    int d = e0 - e1;
    if (is0(m0)) oe = e1;
    else if (is0(m1)) oe = e0;
    else if (e1 > e0) { o0 = mkC(exp2_shift(m0.re, d), exp2_shift(m0.im, d)); oe = e1; }
    else if (e0 > e1) { o1 = mkC(exp2_shift(m1.re, -d), exp2_shift(m1.im, -d)); oe = e0; }
    else oe = e0;
}
/* numba_xr.py:650-672 */
FSB_HD XF normalize(double m, int e)
{
    XF r;
    normalize_real(m, e, r.m, r.e);
    return r;
}
FSB_HD XC normalize(C m, int e)
{
    double nre, nim;
    int ere, eim;
    normalize_real(m.re, e, nre, ere);
    normalize_real(m.im, e, nim, eim);
    XC r;
    coexp_f(nre, ere, nim, eim, r.m.re, r.m.im, r.e);
    return r;
}
FSB_HD bool need_renorm(C m) { return need_renorm(m.re) || need_renorm(m.im); }
FSB_HD XF to_xr(double v) { return normalize(v, 0); }
FSB_HD XC to_xr(C v) { return normalize(v, 0); }

/* numba_xr.py:802-829 */
/* m * 2^e: exact exponent-field arithmetic while the mantissa and the result
 * are normal doubles (the common case, a dozen instructions); everything else
 * -- zeros, denormal results, overflow -- goes to one out-of-line ldexp so that
 * the many call sites stay small (instruction-cache footprint). */
#ifdef __CUDA_ARCH__
static __device__ __forceinline__ double ldexp_slow(double m, int e) { return ldexp(m, e); }
#endif
FSB_XF_OP double to_std(XF x)
{
#ifdef __CUDA_ARCH__
    const int hi = hi32(x.m);
    const int fld = (hi >> 20) & 0x7ff;
    const int nf = fld + x.e;
    if (fld != 0 && fld != 0x7ff && x.e > -8192 && x.e < 8192) {
        if (nf > 0 && nf < 0x7ff) return mk64(hi + (x.e << 20), lo32(x.m));
        if (nf < -53) return mk64(hi & (int)0x80000000, 0);    /* below half the smallest denormal */
    }
    if (x.m == 0.) return x.m;
    return ldexp_slow(x.m, x.e);
#else
    return ldexp(x.m, x.e);
#endif
}
FSB_HD C to_std(XC x)
{
    XC n = normalize(x.m, x.e);
    double s = ldexp1(n.e);
    return mkC(n.m.re * s, n.m.im * s);
}

/* add / sub, numba_xr.py:318-388 */
FSB_XF_OP XF operator+(XF a, XF b)
{
    double x, y; int e;
    coexp_f(a.m, a.e, b.m, b.e, x, y, e);
    return mkXF(x + y, e);
}
FSB_XF_OP XF operator-(XF a, XF b)
{
    double x, y; int e;
    coexp_f(a.m, a.e, b.m, b.e, x, y, e);
    return mkXF(x - y, e);
}
FSB_HD XF as_operand(double v) { return need_renorm(v) ? normalize(v, 0) : mkXF(v, 0); }
FSB_HD XC as_operand(C v) { return need_renorm(v) ? normalize(v, 0) : mkXC(v, 0); }
FSB_HD XF operator+(XF a, double b) { return a + as_operand(b); }
FSB_HD XF operator-(XF a, double b) { return a - as_operand(b); }
FSB_HD XF operator-(XF a) { return mkXF(-a.m, a.e); }
FSB_HD XC operator+(XC a, XC b)
{
    C x, y; int e;
    coexp_c(a.m, a.e, b.m, b.e, x, y, e);
    return mkXC(x + y, e);
}
FSB_HD XC operator+(XC a, C b) { return a + as_operand(b); }
FSB_HD XC operator+(XC a, XF b)
{
    C x, y; int e;
    coexp_c(a.m, a.e, mkC(b.m, 0.), b.e, x, y, e);
    return mkXC(x + y, e);
}

/* mul, numba_xr.py:416-444 */
FSB_XF_OP XF xr_pack(double m, int e) { return need_renorm(m) ? normalize(m, e) : mkXF(m, e); }
FSB_HD XC xr_pack(C m, int e) { return need_renorm(m) ? normalize(m, e) : mkXC(m, e); }
FSB_HD XF operator*(XF a, XF b) { return xr_pack(a.m * b.m, a.e + b.e); }
FSB_HD XF operator*(XF a, double b) { return xr_pack(a.m * b, a.e); }
FSB_HD XF operator*(double a, XF b) { return xr_pack(a * b.m, b.e); }
FSB_HD XC operator*(XC a, XC b) { return xr_pack(a.m * b.m, a.e + b.e); }
FSB_HD XC operator*(double a, XC b) { return xr_pack(a * b.m, b.e); }
FSB_HD XC operator*(C a, XC b) { return xr_pack(a * b.m, b.e); }
FSB_HD XC operator*(XF a, C b) { return xr_pack(a.m * b, a.e); }

/* compare, numba_xr.py:476-510 */
FSB_HD bool xr_le(XF a, XF b)
{
    double x, y; int e;
    coexp_f(a.m, a.e, b.m, b.e, x, y, e);
    return x <= y;
}
FSB_HD bool xr_lt(XF a, double b)
{
    double x, y; int e;
    coexp_f(a.m, a.e, b, 0, x, y, e);
    return x < y;
}
FSB_HD bool operator>=(XF a, double b)
{
    double x, y; int e;
    coexp_f(a.m, a.e, b, 0, x, y, e);
    return x >= y;
}
FSB_HD bool operator<=(XF a, double b)
{
    double x, y; int e;
    coexp_f(a.m, a.e, b, 0, x, y, e);
    return x <= y;
}
/* numba_xr.py:531-561 */
FSB_HD XF abs2(XC a) { return mkXF(a.m.re * a.m.re + a.m.im * a.m.im, a.e + a.e); }
FSB_HD double fabs_(double x) { return fabs(x); }
FSB_HD XF fabs_(XF x) { return mkXF(fabs(x.m), x.e); }
FSB_HD double sgn(double x) { return (x < 0.) ? -1. : 1.; }
FSB_HD double sgn_(double x) { return sgn(x); }
FSB_HD double sgn_(XF x) { return (x.m < 0.) ? -1. : 1.; }


/* ---- projections (projection.py) ------------------------------------------
 * exp / sin / cos with a fixed operation sequence.  The reference evaluates
 * the Expmap projection with numba's complex exp = libm exp, cos, sin
 * (cmathimpl.exp_impl: r = exp(x); (r cos y, r sin y)).  libm results are not
 * portable bit for bit (glibc on the host, CUDA's libdevice here: both < 1
 * ulp but not identical), so the pixel -> c mapping uses the classic
 * table-free algorithms below (the fdlibm ones: Cody-Waite reduction, a
 * rational / polynomial kernel, error < 1 ulp) with every operation rounded
 * individually -- no FMA contraction in either build -- which makes the
 * projected pixel a pure function of its input on every platform: the CPU
 * oracle restates the same sequence and agrees bit for bit. */
FSB_HD double sub_rn(double a, double b) { return add_rn(a, -b); }
FSB_HD double div_rn(double a, double b)
{
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    volatile double r = a / b; return r;
#endif
}
FSB_HD double det_exp(double x)
{
    const double ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10,
                 invln2 = 1.44269504088896338700e+00,
                 P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03,
                 P3 = 6.61375632143793436117e-05, P4 = -1.65339022054652515390e-06,
                 P5 = 4.13813679705723846039e-08;
    if (!(x == x)) return x;
    if (x > 7.09782712893383973096e+02) return mk64(0x7ff00000, 0);
    if (x < -7.45133219101941108420e+02) return 0.;
    const double ax = fabs(x);
    double hi = x, lo = 0.;
    int k = 0;
    if (ax > 0.34657359027997264) {                 /* 0.5 ln 2 */
        if (ax < 1.0397207708399179) {              /* 1.5 ln 2 */
            k = (x < 0.) ? -1 : 1;
            hi = (x < 0.) ? add_rn(x, ln2HI) : sub_rn(x, ln2HI);
            lo = (x < 0.) ? -ln2LO : ln2LO;
        } else {
            k = (int)add_rn(mul_rn(invln2, x), (x < 0.) ? -0.5 : 0.5);
            const double t = (double)k;
            hi = sub_rn(x, mul_rn(t, ln2HI));
            lo = mul_rn(t, ln2LO);
        }
        x = sub_rn(hi, lo);
    } else if (ax < 3.725290298461914e-09) {        /* 2^-28 */
        return add_rn(1., x);
    }
    const double t = mul_rn(x, x);
    double q = add_rn(P4, mul_rn(t, P5));
    q = add_rn(P3, mul_rn(t, q));
    q = add_rn(P2, mul_rn(t, q));
    q = add_rn(P1, mul_rn(t, q));
    const double c = sub_rn(x, mul_rn(t, q));
    if (k == 0)
        return sub_rn(1., sub_rn(div_rn(mul_rn(x, c), sub_rn(c, 2.)), x));
    const double y = sub_rn(1., sub_rn(sub_rn(lo, div_rn(mul_rn(x, c), sub_rn(2., c))), hi));
    const int k1 = k / 2;
    return mul_rn(mul_rn(y, ldexp1(k1)), ldexp1(k - k1));   /* 2^k in two exact factors */
}
/* x = n pi/2 + (y0 + y1), |y0| <= pi/4 ; |x| < 2^20 pi/2 */
FSB_HD int det_rem_pio2(double x, double &y0, double &y1)
{
    const double invpio2 = 6.36619772367581382433e-01,
                 pio2_1 = 1.57079632673412561417e+00, pio2_1t = 6.07710050650619224932e-11,
                 pio2_2 = 6.07710050630396597660e-11, pio2_2t = 2.02226624879595063154e-21,
                 pio2_3 = 2.02226624871116645580e-21, pio2_3t = 8.47842766036889956997e-32;
    const double ax = fabs(x);
    if (ax <= 0.78539816339744830962) { y0 = x; y1 = 0.; return 0; }
    const int n = (int)add_rn(mul_rn(ax, invpio2), 0.5);
    const double fn = (double)n;
    double r = sub_rn(ax, mul_rn(fn, pio2_1));
    double w = mul_rn(fn, pio2_1t);
    const int j = expfield(ax);
    y0 = sub_rn(r, w);
    if (j - expfield(y0) > 16) {
        double t = r;
        w = mul_rn(fn, pio2_2);
        r = sub_rn(t, w);
        w = sub_rn(mul_rn(fn, pio2_2t), sub_rn(sub_rn(t, r), w));
        y0 = sub_rn(r, w);
        if (j - expfield(y0) > 49) {
            t = r;
            w = mul_rn(fn, pio2_3);
            r = sub_rn(t, w);
            w = sub_rn(mul_rn(fn, pio2_3t), sub_rn(sub_rn(t, r), w));
            y0 = sub_rn(r, w);
        }
    }
    y1 = sub_rn(sub_rn(r, y0), w);
    if (x < 0.) { y0 = -y0; y1 = -y1; return -n; }
    return n;
}
FSB_HD double det_ksin(double x, double y)
{
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    if (fabs(x) < 7.450580596923828e-09) return x;           /* 2^-27 */
    const double z = mul_rn(x, x), v = mul_rn(z, x);
    double r = add_rn(S5, mul_rn(z, S6));
    r = add_rn(S4, mul_rn(z, r));
    r = add_rn(S3, mul_rn(z, r));
    r = add_rn(S2, mul_rn(z, r));
    return sub_rn(x, sub_rn(sub_rn(mul_rn(z, sub_rn(mul_rn(0.5, y), mul_rn(v, r))), y),
                            mul_rn(v, S1)));
}
FSB_HD double det_kcos(double x, double y)
{
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const double ax = fabs(x);
    if (ax < 7.450580596923828e-09) return 1.;
    const double z = mul_rn(x, x);
    double r = add_rn(C5, mul_rn(z, C6));
    r = add_rn(C4, mul_rn(z, r));
    r = add_rn(C3, mul_rn(z, r));
    r = add_rn(C2, mul_rn(z, r));
    r = mul_rn(z, add_rn(C1, mul_rn(z, r)));
    const double zr_xy = sub_rn(mul_rn(z, r), mul_rn(x, y));
    if (ax < 0.3) return sub_rn(1., sub_rn(mul_rn(0.5, z), zr_xy));
    const double qx = (ax > 0.78125) ? 0.28125 : mk64(hi32(ax) - 0x00200000, 0);   /* ~ |x| / 4 */
    const double hz = sub_rn(mul_rn(0.5, z), qx);
    return sub_rn(sub_rn(1., qx), sub_rn(hz, zr_xy));
}
FSB_HD void det_sincos(double x, double &s, double &c)
{
    if (!(fabs(x) < 1.6e6)) { s = c = mk64(0x7ff80000, 0); return; }   /* out of the supported range */
    double y0, y1;
    const int n = det_rem_pio2(x, y0, y1);
    const double ks = det_ksin(y0, y1), kc = det_kcos(y0, y1);
    switch (n & 3) {
    case 0: s = ks; c = kc; break;
    case 1: s = kc; c = -ks; break;
    case 2: s = -ks; c = -kc; break;
    default: s = -kc; c = ks; break;
    }
}

/* projection.py:363-373 (Expmap.make_f_impl): exp(hmoy + pix_to_ht * pix), with
 * numba's complex product and complex exp (cmathimpl.exp_impl) */
FSB_HD C proj_expmap(C pix, double hmoy, C k)
{
    const C ht = cmul_rn(k, pix);
    const double r = det_exp(add_rn(hmoy, ht.re));
    double s, c;
    det_sincos(add_rn(0., ht.im), s, c);        /* float + complex: (hmoy + re, 0 + im) */
    return mkC(mul_rn(r, c), mul_rn(r, s));
}
/* projection.py:455-471 (Expmap.make_dzndc_modifier): exp(Re(pix_to_ht * pix) + hshift) */
FSB_HD double modifier_expmap(C pix, C k, double hshift)
{
    const double h = add_rn(mul_rn(k.re, pix.re), -mul_rn(k.im, pix.im));
    return det_exp(add_rn(h, hshift));
}
/* projection.py:205-219 (Cartesian.make_dzndc_modifier): |pix + 1e-6| * expmap_seam */
FSB_HD double modifier_seam(C pix, double seam)
{
    return mul_rn(hypot_rn(add_rn(pix.re, 1.e-6), pix.im), seam);
}
/* complex128 *= float64 as numba lowers it: the float is cast to complex first */
FSB_HD C cmul_real_numba(C z, double m)
{
    return mkC(add_rn(mul_rn(z.re, m), -mul_rn(z.im, 0.)), add_rn(mul_rn(z.re, 0.), mul_rn(z.im, m)));
}

} /* namespace fsb */
