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

} /* namespace fsb */
