# -*- coding: utf-8 -*-
"""
Host-side helpers for "extended range" scalars: value = mantissa * 2**exp,
returned as plain (mantissa, exp) pairs.  They reproduce what the reference
builds with `Xrange_array` for the few per-frame scalars of the hot path
(numpy_utils/xrange.py:8-49 mpc_to_Xrange / mpf_to_Xrange, and the
`x_Xr + 1j * y_Xr` sum of mpmath_utils/FP_loop.pyx:424-441).
"""
import numpy as np
import mpmath


def mpf_to_xr(mpf):
    """ numpy_utils/xrange.py:46-49 """
    m, e = mpmath.frexp(mpf)
    return float(m), int(e)


def mpc_to_xr(mpc):
    """ numpy_utils/xrange.py:8-43 (literal, including the exp==0 quirk) """
    mpc = mpmath.mpc(mpc)
    mx, ex = mpmath.frexp(mpc.real)
    my, ey = mpmath.frexp(mpc.imag)
    mx, my, ex, ey = float(mx), float(my), int(ex), int(ey)
    if ex > ey:
        case = 1
    elif ex < ey:
        case = 2
    else:
        case = 3
    if ex == 0:
        case = 2
    if ey == 0:
        case = 1
    if case == 1:
        m = complex(mx, float(np.ldexp(my, ey - ex)))
        e = ex
    elif case == 2:
        m = complex(float(np.ldexp(mx, ex - ey)), my)
        e = ey
    else:
        m = complex(mx, my)
        e = ex
    return m, e


def _exp2(k):
    """ Xrange_array._exp2 (xrange.py:1013-1025): 2**k by exponent-field
    construction, flushed to 0 below the normal range """
    f = 1023 + int(k)
    if f <= 0:
        return 0.0
    return float(np.ldexp(1.0, int(k)))


def xr_complex_from_parts(mx, ex, my, ey):
    """ (mx * 2**ex) + 1j * (my * 2**ey) as the reference's Xrange_array sum
    builds it (xrange.py:797-911, vector branch of _cplx_coexp_ufunc) """
    m0 = complex(mx, 0.0)
    m1 = complex(0.0, my)
    m0_null = (m0 == 0)
    m1_null = (m1 == 0)
    d = ex - ey
    if (ey > ex) and not m1_null:
        m0 = m0 * _exp2(d)
    if (ex > ey) and not m0_null:
        m1 = m1 * _exp2(-d)
    e = max(ex, ey)
    if m0_null:
        e = ey
    if m1_null:
        e = ex
    return m0 + m1, int(e)


def xr_to_float(m, e):
    return float(np.ldexp(m, e))
