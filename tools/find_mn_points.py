# -*- coding: utf-8 -*-
"""
Boundary points of the power-N Mandelbrot sets for the Perturbation_mandelbrot_N
parity cases: high-precision bisection along a ray between a bounded point and
an escaping one, with the native MPFR orbit as membership test.
    python tools/find_mn_points.py      -> prints the strings used in tests/cases.py
"""
import ctypes, os, sys
import mpmath
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
from fractalshades_b200 import _native

# exponent -> (bounded point, escaping point, digits, iterations of the test)
START = {
    3: (("0.1", "0.5"), ("0.3", "0.9"), 60, 400),
    4: (("-0.6", "0.4"), ("-1.0", "0.9"), 60, 400),
    5: (("0.2", "0.5"), ("0.5", "0.9"), 400, 8000),
}


def escapes(lib, nexp, x, y, prec, digits, n):
    orbit = np.zeros(2 * (n + 1))
    buf = (_native.OrbitXr * 4096)()
    cnt = ctypes.c_int64(0)
    i = lib.fsb_orbit_mandelbrot(orbit.ctypes.data, n, nexp, 0, 1e3,
                                 mpmath.nstr(x, digits).encode(), mpmath.nstr(y, digits).encode(),
                                 prec, buf, 4096, ctypes.byref(cnt))
    assert i >= 0
    return i <= n


def main():
    lib = _native.load_orbit_lib()
    for nexp, (pa, pb, digits, n) in START.items():
        mpmath.mp.dps = digits
        prec = mpmath.mp.prec
        a = (mpmath.mpf(pa[0]), mpmath.mpf(pa[1]))
        b = (mpmath.mpf(pb[0]), mpmath.mpf(pb[1]))
        assert not escapes(lib, nexp, a[0], a[1], prec, digits, n), nexp
        assert escapes(lib, nexp, b[0], b[1], prec, digits, n), nexp
        while max(abs(a[0] - b[0]), abs(a[1] - b[1])) > mpmath.mpf(10) ** (-(digits - 20)):
            m = ((a[0] + b[0]) / 2, (a[1] + b[1]) / 2)
            if escapes(lib, nexp, m[0], m[1], prec, digits, n):
                b = m
            else:
                a = m
        print(f'    {nexp}: ("{mpmath.nstr(a[0], digits - 25)}",\n        "{mpmath.nstr(a[1], digits - 25)}"),')


if __name__ == "__main__":
    main()
