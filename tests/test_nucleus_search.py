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
