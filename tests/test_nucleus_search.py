# -*- coding: utf-8 -*-
"""
Native period / nucleus search (include/fsb200_orbit.h) against the
reference's own known-answer vectors (tests/test_FP_loop.py:166-272 of the
reference, copied as data into tests/golden/fp_loop_kat.json by the session
that built this repo) and against a plain-mpmath restatement on a small case.
Host only (MPFR), no GPU.
"""
import ctypes
import json
import os

import mpmath
import pytest

from fractalshades_b200 import _native, settings
import fractalshades_b200.models as fsm

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "fp_loop_kat.json")))


def test_ball_method_known_period():
    b = KAT["ball"]
    lib = _native.load_orbit_lib()
    order = lib.fsb_ball_method_mandelbrot(
        b["x"].encode(), b["y"].encode(), b["precision_bits"], b["radius"].encode(),
        b["maxiter"], b["M_divergence"])
    assert order == b["order"] == 4252


def test_newton_known_nucleus():
    n = KAT["newton"]
    prec = n["precision_bits"]
    old = mpmath.mp.prec
    mpmath.mp.prec = prec
    try:
        eps_pix = mpmath.mpf(val=(2, -n["pix_bits"]))
        c = mpmath.mpc(mpmath.mpf(n["x_start"]), mpmath.mpf(n["y_start"]))
        ok, val = fsm.Perturbation_mandelbrot.find_nucleus(c, n["order"], eps_pix, 80)
        assert ok
        with mpmath.workprec(prec - 10):
            assert mpmath.almosteq(mpmath.mpf(n["x_nucleus"]), val.real)
            assert mpmath.almosteq(mpmath.mpf(n["y_nucleus"]), val.imag)
        ok2, val2 = fsm.Perturbation_mandelbrot.find_any_nucleus(c, n["order"], eps_pix, 80)
        assert ok2
        with mpmath.workprec(prec - 10):
            assert mpmath.almosteq(val.real, val2.real)
    finally:
        mpmath.mp.prec = old


def _mp_find_nucleus(c, order, max_newton):
    """ the Python statement quoted in the reference's docstring
    (FP_loop.pyx:935-960), plain mpmath """
    c_loop = c
    for _ in range(max_newton):
        zr = mpmath.mpc(0); dz = mpmath.mpc(0); h = mpmath.mpc(1); dh = mpmath.mpc(0)
        for i in range(1, order + 1):
            dz = 2 * dz * zr + 1
            zr = zr * zr + c_loop
            if i < order and order % i == 0:
                h *= zr
                dh += dz / zr
        f = zr / h
        df = (dz * h - zr * dh) / (h * h)
        cc = c_loop - f / df
        done = mpmath.almosteq(cc, c_loop)
        c_loop = cc
        if done:
            break
    return c_loop


def test_small_case_matches_mpmath_restatement():
    old = mpmath.mp.prec
    mpmath.mp.prec = 200
    try:
        c = mpmath.mpc("-1.7548", "0.0001")       # near the period-3 nucleus
        lib = _native.load_orbit_lib()
        order = lib.fsb_ball_method_mandelbrot(b"-1.7548", b"0.0001", 200, b"0.01", 1000, 1e5)
        assert order == 3
        ok, val = fsm.Perturbation_mandelbrot.find_nucleus(c, 3, mpmath.mpf("1e-30"), 80)
        assert ok
        ref = _mp_find_nucleus(c, 3, 80)
        with mpmath.workprec(180):
            assert mpmath.almosteq(val.real, ref.real)
            assert abs(val.imag) < mpmath.mpf("1e-50") and abs(ref.imag) < mpmath.mpf("1e-50")
        # escaping seed: no period
        assert lib.fsb_ball_method_mandelbrot(b"1.0", b"1.0", 200, b"1e-3", 1000, 1e5) == -1
        # bad number string
        assert lib.fsb_ball_method_mandelbrot(b"abc", b"1.0", 200, b"1e-3", 10, 1e5) == -3
    finally:
        mpmath.mp.prec = old


def test_get_FP_orbit_uses_the_nucleus_as_periodic_reference(tmp_path):
    """ default reference behaviour (settings.no_newton = False): periodic
    reference orbit of length `order` (perturbation.py:651-768, 808-854) """
    old = settings.no_newton
    settings.no_newton = False
    try:
        f = fsm.Perturbation_mandelbrot(str(tmp_path))
        f.zoom(precision=40, x="-1.7548776662466927600495088963585286918946066177727931",
               y="1.e-12", dx="1.e-8", nx=200, xy_ratio=1.0, theta_deg=0.)
        f.max_iter = 5000
        f.M_divergence = 1.e3
        f.get_FP_orbit()
        FP = f.FP_params
        assert FP["order"] == 3
        assert FP["ref_orbit_len"] == 3 and f.Zn_path.shape[0] == 3
        assert abs(complex(FP["ref_point"]) - (-1.7548776662466927 + 0j)) < 1e-12
    finally:
        settings.no_newton = old


# ---------------------------------------------------------------------------
# Burning-ship family and z^N + c (include/fsb200_orbit.h, second half).
# Known answers of the reference's own tests (tests/test_FP_loop.py:311-366 of the
# reference: two period-3 seeds of the burning ship), and a plain-mpmath
# restatement of the two algorithms (the Python given in FP_loop.pyx's docstrings
# :2243-2266, :2612-2632 with the Jacobian in place of the complex derivative) for
# every flavour / a few exponents.  Test infrastructure only.
BS_KAT = [("-1.7545128115395", "0.0015894811966473", "0.01"),
          ("0.88410156557344", "1.5218981991448", "0.01")]


def _bs_iter(flavor, x, y, a, b):
    ab = mpmath.fabs
    if flavor == 1: return x * x - y * y + a, 2 * ab(x * y) - b
    if flavor == 2: return x * x - y * y + a, 2 * x * ab(y) - b
    if flavor == 3: return x * x - y * ab(y) + a, 2 * x * y - b
    if flavor == 4: return ab(x * x - y * y) + a, 2 * x * y - b
    return ab(x * x - y * y) + a, 2 * ab(x * y) - b


def _sgn(v):
    return 1 if v >= 0 else -1


def _bs_jac(flavor, x, y, J):
    """ d(x', y')/d(a, b) from d(x, y)/d(a, b) = J = (xx, xy, yx, yy) """
    xx, xy, yx, yy = J
    sd = _sgn(x * x - y * y) if flavor in (4, 5) else 1
    Y = mpmath.fabs(y) if flavor == 3 else y
    n_xx, n_xy = 2 * sd * (x * xx - Y * yx), 2 * sd * (x * xy - Y * yy)
    if flavor in (1, 5):
        ax, ay, sx, sy = mpmath.fabs(x), mpmath.fabs(y), _sgn(x), _sgn(y)
        n_yx, n_yy = 2 * (ax * sy * yx + sx * xx * ay), 2 * (ax * sy * yy + sx * xy * ay)
    elif flavor == 2:
        ay, sy = mpmath.fabs(y), _sgn(y)
        n_yx, n_yy = 2 * (x * sy * yx + xx * ay), 2 * (x * sy * yy + xy * ay)
    else:
        n_yx, n_yy = 2 * (x * yx + xx * y), 2 * (x * yy + xy * y)
    return n_xx + 1, n_xy, n_yx, n_yy - 1


def _solve(J, e, f):
    a, b, c, d = J
    det = a * d - c * b
    return (d * e - b * f) / det, (a * f - c * e) / det


def _mp_bs_ball(flavor, a, b, px, maxiter, M):
    x = y = mpmath.mpf(0)
    J = (mpmath.mpf(0),) * 4
    for i in range(1, maxiter + 1):
        J = _bs_jac(flavor, x, y, J)
        x, y = _bs_iter(flavor, x, y, a, b)
        rx, ry = _solve(J, x, y)
        if mpmath.hypot(x, y) > M:
            return -1
        if mpmath.hypot(rx / px, ry / px) < 1:
            return i
    return -1


def _mp_bs_newton(flavor, a, b, order, max_newton):
    for _ in range(max_newton):
        x = y = mpmath.mpf(0)
        J = (mpmath.mpf(0),) * 4
        for i in range(order):
            J = _bs_jac(flavor, x, y, J)
            x, y = _bs_iter(flavor, x, y, a, b)
        da, db = _solve(J, x, y)
        a, b = a - da, b - db
        if mpmath.hypot(da, db) <= mpmath.mpf(2) ** (6 - mpmath.mp.prec):
            break
    return a, b


def test_burning_ship_known_periods_and_newton():
    lib = _native.load_orbit_lib()
    for (xs, ys, px) in BS_KAT:
        order = lib.fsb_ball_method_burning_ship(1, xs.encode(), ys.encode(), 53, px.encode(),
                                                 100, 1000.)
        assert order == 3                       # reference: tests/test_FP_loop.py:325,339
    old = mpmath.mp.prec
    mpmath.mp.prec = 53
    try:
        f = fsm.Perturbation_burning_ship(os.path.join(HERE, "_tmp_unused"))
        for (xs, ys, px) in BS_KAT:
            c = mpmath.mpc(mpmath.mpf(xs), mpmath.mpf(ys))
            ok, val = f.find_any_nucleus(c, 3, mpmath.mpf("0.012"), 40)
            assert ok
            assert abs(val - c) < 0.02
            x, y = mpmath.mpf(0), mpmath.mpf(0)
            for _ in range(3):                   # a genuine period-3 point
                x, y = _bs_iter(1, x, y, val.real, val.imag)
            assert mpmath.hypot(x, y) < 1e-13
        with pytest.raises(NotImplementedError):
            f.find_nucleus(c, 3, 0.01)
    finally:
        mpmath.mp.prec = old


# (flavour, x, y): points next to a mini-ship of each flavour (tests/cases.py)
_BS_SEEDS = {1: ("-1.7505941429008662", "0.0214423020148999"),
             2: ("-1.3604879916723847", "0.0015808658322309468"),
             3: ("-1.4477399868839198", "-0.6048320439477123"),
             4: ("-1.760370697674034", "0.011733974791909326"),
             5: ("-1.758364745737221", "0.024352431136909887")}


@pytest.mark.parametrize("flavor", [1, 2, 3, 4, 5])
def test_burning_ship_search_matches_mpmath_restatement(flavor):
    lib = _native.load_orbit_lib()
    xs, ys = _BS_SEEDS[flavor]
    old = mpmath.mp.prec
    mpmath.mp.prec = 200
    try:
        a, b = mpmath.mpf(xs), mpmath.mpf(ys)
        found = None
        for px in ("1e-3", "1e-2", "5e-2"):
            got = lib.fsb_ball_method_burning_ship(flavor, xs.encode(), ys.encode(), 200,
                                                   px.encode(), 3000, 1000.)
            want = _mp_bs_ball(flavor, a, b, mpmath.mpf(px), 3000, 1000.)
            assert got == want, (flavor, px, got, want)
            if got > 0 and found is None:
                found = (got, px)
        assert found is not None, "no period found around the seed"
        order, px = found
        f = fsm.Perturbation_burning_ship(os.path.join(HERE, "_tmp_unused"),
                                          flavor=fsm.BS_flavor_list[flavor - 1])
        ok, val = f.find_any_nucleus(mpmath.mpc(a, b), order, mpmath.mpf(px), 80)
        wa, wb = _mp_bs_newton(flavor, a, b, order, 80)
        x, y = mpmath.mpf(0), mpmath.mpf(0)
        for _ in range(order):
            x, y = _bs_iter(flavor, x, y, wa, wb)
        if mpmath.hypot(x, y) < mpmath.mpf(px):          # the restatement converged
            assert ok
            assert mpmath.hypot(val.real - wa, val.imag - wb) < mpmath.mpf(2) ** -170
        else:
            assert not ok
    finally:
        mpmath.mp.prec = old


@pytest.mark.parametrize("exponent,xs,ys", [(3, "-0.1245", "1.0832"), (4, "-1.0107", "0.3721"),
                                            (5, "0.6713", "0.5172")])
def test_power_n_search_matches_mpmath_restatement(exponent, xs, ys):
    lib = _native.load_orbit_lib()
    old = mpmath.mp.prec
    mpmath.mp.prec = 160
    try:
        c = mpmath.mpc(mpmath.mpf(xs), mpmath.mpf(ys))

        def ball(px):
            z = dz = mpmath.mpc(0)
            for i in range(1, 2001):
                dz = exponent * dz * z ** (exponent - 1) + 1
                z = z ** exponent + c
                if abs(z) > 1000.:
                    return -1
                if abs(z / dz / px) < 1:
                    return i
            return -1
        found = None
        for px in ("1e-3", "1e-2", "1e-1"):
            got = lib.fsb_ball_method_mandelbrot_n(exponent, xs.encode(), ys.encode(), 160,
                                                   px.encode(), 2000, 1000.)
            assert got == ball(mpmath.mpf(px)), (exponent, px)
            if got > 0 and found is None:
                found = (got, px)
        if found is None:
            pytest.skip("no period around this seed")
        order, px = found
        f = fsm.Perturbation_mandelbrot_N(os.path.join(HERE, "_tmp_unused"), exponent)
        ok, val = f.find_any_nucleus(c, order, mpmath.mpf(px), 80)
        cl = c
        for _ in range(80):
            z = dz = mpmath.mpc(0)
            for i in range(order):
                dz = exponent * dz * z ** (exponent - 1) + 1
                z = z ** exponent + cl
            step = z / dz
            cl = cl - step
            if abs(step) <= mpmath.mpf(2) ** (6 - 160):
                break
        z = mpmath.mpc(0)
        for i in range(order):
            z = z ** exponent + cl
        if abs(z) < mpmath.mpf(px):
            assert ok and abs(val - cl) < mpmath.mpf(2) ** -130
        else:
            assert not ok
        with pytest.raises(NotImplementedError):
            f.find_nucleus(c, order, 0.01)
    finally:
        mpmath.mp.prec = old
