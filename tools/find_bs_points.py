# -*- coding: utf-8 -*-
"""
Deep boundary points of the burning-ship family for the Xrange parity cases:
high-precision bisection between a point that stays bounded for N iterations
and one that escapes, with the native MPFR orbit as membership test.
    python tools/find_bs_points.py            -> prints the strings used in tests/cases.py
"""
import ctypes, os, sys
import mpmath
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
from fractalshades_b200 import _native

START = {   # (inside, outside) in double precision, from tests/cases.py _BS_PTS
    2: ("-1.3604879916723847", "0.0015808658322309468"),
    3: ("-1.4477399868839198", "-0.6048320439477123"),
    4: ("-1.760370697674034", "0.011733974791909326"),
    5: ("-1.758364745737221", "0.024352431136909887"),
}
N = 6000
DIGITS = 420


def escapes(lib, flavor, x, y, prec):
    orbit = np.zeros(2 * (N + 1))
    buf = (_native.OrbitXr * 4096)()
    cnt = ctypes.c_int64(0)
    i = lib.fsb_orbit_burning_ship(orbit.ctypes.data, N, flavor, 0, 1e3,
                                   mpmath.nstr(x, DIGITS).encode(), mpmath.nstr(y, DIGITS).encode(),
                                   prec, buf, 4096, ctypes.byref(cnt))
    assert i >= 0
    return i <= N


def main():
    lib = _native.load_orbit_lib()
    mpmath.mp.dps = DIGITS
    prec = mpmath.mp.prec
    for flavor, (sx, sy) in START.items():
        x0, y0 = mpmath.mpf(sx), mpmath.mpf(sy)
        e0 = escapes(lib, flavor, x0, y0, prec)
        # a partner with the other status along +x / -x / +y
        partner = None
        for step in ("1e-10", "1e-8", "1e-6", "1e-4", "1e-2"):
            for dxs, dys in ((1, 0), (-1, 0), (0, 1), (0, -1)):
                x1, y1 = x0 + dxs * mpmath.mpf(step), y0 + dys * mpmath.mpf(step)
                if escapes(lib, flavor, x1, y1, prec) != e0:
                    partner = (x1, y1)
                    break
            if partner:
                break
        assert partner, flavor
        a, b = ((x0, y0), partner) if not e0 else (partner, (x0, y0))   # a bounded, b escapes
        while max(abs(a[0] - b[0]), abs(a[1] - b[1])) > mpmath.mpf(10) ** (-(DIGITS - 40)):
            m = ((a[0] + b[0]) / 2, (a[1] + b[1]) / 2)
            if escapes(lib, flavor, m[0], m[1], prec):
                b = m
            else:
                a = m
        print(f'    {flavor}: ("{mpmath.nstr(a[0], DIGITS - 50)}",\n        "{mpmath.nstr(a[1], DIGITS - 50)}"),')


if __name__ == "__main__":
    main()
