# -*- coding: utf-8 -*-
"""
Models on the hot path, with the reference's class names, constructor and
`calc_std_div` signatures (models/mandelbrot_M2.py:61-71,432-444;
models/burning_ship.py:230-238,971-981), field codes and stop codes.

The plugin contract of the reference is kept (`calc_std_div` returns the dict
of three factories `set_state / initialize / iterate`, utils.py:278-339,
core.py:1918-1949); since numba closures cannot cross the C ABI, `initialize()`
and `iterate()` return a `KernelSpec` naming the device-kernel variant.
"""
import ctypes

import mpmath
import numpy as np

from . import settings
from . import _native
from .core import Fractal, KernelSpec, calc_options
from .perturbation import PerturbationFractal

BS_flavor_list = ("Burning ship", "Perpendicular burning ship", "Shark fin",
                  "Celtic", "Buffalo")          # models/burning_ship.py:63-69


def get_flavor_int(flavor):
    if flavor not in BS_flavor_list:
        raise ValueError(f"unknown burning-ship flavor {flavor!r}")
    return BS_flavor_list.index(flavor) + 1


def _set_potential(self):
    self.potential_kind = "infinity"
    self.potential_d = 2
    self.potential_a_d = 1.
    self.potential_M_cutoff = 1000.


class Mandelbrot(Fractal):
    """ Standard power-2 Mandelbrot (models/mandelbrot_M2.py:17-176) """

    def __init__(self, directory: str):
        super().__init__(directory)
        _set_potential(self)
        self.holomorphic = True

    @calc_options
    def calc_std_div(self, *, calc_name: str = "base_calc", subset=None,
                     max_iter: int = 10000, M_divergence: float = 1000.,
                     epsilon_stationnary: float = 0.01,
                     calc_d2zndc2: bool = False, calc_orbit: bool = False,
                     backshift: int = 0):
        complex_codes = ["zn", "dzndz", "dzndc"]
        if calc_d2zndc2:
            complex_codes += ["d2zndc2"]
        if calc_orbit:
            complex_codes += ["zn_orbit"]
        int_codes = []
        stop_codes = ["max_iter", "divergence", "stationnary"]

        def set_state():
            def impl(instance):
                instance.codes = (complex_codes, int_codes, stop_codes)
                instance.complex_type = np.complex128
                instance.potential_M = M_divergence
                instance.backshift = backshift if calc_orbit else None
            return impl

        spec = KernelSpec(kind="std_M2", model=_native.FSB_MODEL_M2, flavor=0,
                          max_iter=max_iter, M_divergence=M_divergence,
                          epsilon_stationnary=epsilon_stationnary,
                          calc_d2zndc2=calc_d2zndc2, calc_orbit=calc_orbit,
                          backshift=backshift)
        return {"set_state": set_state, "initialize": lambda: spec,
                "iterate": lambda: spec}


class Mandelbrot_N(Fractal):
    """ Standard power-N Mandelbrot z -> z^N + c (models/mandelbrot_Mn.py:20-350).
    z^(N-1) is evaluated as a product chain (see k_std_mn): identical to the
    reference for N = 3, a few ulp from its C-library polar form otherwise. """

    def __init__(self, directory: str, exponent: int):
        super().__init__(directory)
        if int(exponent) != exponent or not (2 <= exponent <= 32):
            raise ValueError("exponent shall be an integer in [2, 32]")
        self.exponent = int(exponent)
        self.potential_kind = "infinity"
        self.potential_d = self.exponent
        self.potential_a_d = 1.
        self.holomorphic = True

    @calc_options
    def calc_std_div(self, *, calc_name: str, subset=None, max_iter: int,
                     M_divergence: float, epsilon_stationnary: float,
                     calc_d2zndc2: bool = False, calc_orbit: bool = False,
                     backshift: int = 0):
        """ models/mandelbrot_Mn.py:63-145.  calc_orbit: the reference back-shifts the
        stored orbit point with `zn ** N` = the C library's polar-form complex power;
        here the power is the same product chain as in the loop (a few ulp apart,
        tolerance-tested against the reference's fixtures) """
        complex_codes = ["zn", "dzndz", "dzndc"]
        if calc_d2zndc2:
            complex_codes += ["d2zndc2"]
        if calc_orbit:
            complex_codes += ["zn_orbit"]
        int_codes = []
        stop_codes = ["max_iter", "divergence", "stationnary"]

        def set_state():
            def impl(instance):
                instance.codes = (complex_codes, int_codes, stop_codes)
                instance.complex_type = np.complex128
                instance.potential_M = M_divergence
                instance.backshift = backshift if calc_orbit else None
            return impl

        spec = KernelSpec(kind="std_M2", model=_native.FSB_MODEL_M2, flavor=0,
                          nexp=self.exponent, max_iter=max_iter,
                          M_divergence=M_divergence,
                          epsilon_stationnary=epsilon_stationnary,
                          calc_d2zndc2=calc_d2zndc2, calc_orbit=calc_orbit,
                          backshift=backshift)
        return {"set_state": set_state, "initialize": lambda: spec,
                "iterate": lambda: spec}


class Burning_ship(Fractal):
    """ Standard Burning-ship family (models/burning_ship.py:125-429) """

    def __init__(self, directory: str, flavor: str = "Burning ship"):
        super().__init__(directory)
        self.flavor = flavor
        get_flavor_int(flavor)
        _set_potential(self)
        self.holomorphic = False

    @calc_options
    def calc_std_div(self, *, calc_name: str, subset, max_iter: int,
                     M_divergence: float, calc_orbit: bool = False,
                     backshift: int = 0):
        complex_codes = ["xn", "yn", "dxnda", "dxndb", "dynda", "dyndb"]
        if calc_orbit:
            complex_codes += ["xn_orbit", "yn_orbit"]
        int_codes = []
        stop_codes = ["max_iter", "divergence", "stationnary"]

        def set_state():
            def impl(instance):
                instance.codes = (complex_codes, int_codes, stop_codes)
                instance.complex_type = np.float64
                instance.potential_M = M_divergence
                instance.backshift = backshift if calc_orbit else None
            return impl

        spec = KernelSpec(kind="std_BS", model=_native.FSB_MODEL_BS,
                          flavor=get_flavor_int(self.flavor), max_iter=max_iter,
                          M_divergence=M_divergence, calc_orbit=calc_orbit,
                          backshift=backshift)
        return {"set_state": set_state, "initialize": lambda: spec,
                "iterate": lambda: spec}


class Perturbation_mandelbrot(PerturbationFractal):
    """ Arbitrary-precision power-2 Mandelbrot
    (models/mandelbrot_M2.py:359-627) """

    def __init__(self, directory: str):
        super().__init__(directory)
        _set_potential(self)
        self.critical_pt = 0.
        self.FP_code = "zn"
        self.holomorphic = True

    def FP_loop(self, NP_orbit, c0):
        """ models/mandelbrot_M2.py:408-430 -> native MPFR orbit """
        return self._native_orbit(NP_orbit, c0, flavor=None, exponent=2)

    @staticmethod
    def _ball_method(c, px, maxiter, M_divergence):
        """ models/mandelbrot_M2.py:634-650 -> fsb_ball_method_mandelbrot """
        lib = _native.load_orbit_lib()
        order = lib.fsb_ball_method_mandelbrot(
            str(c.real).encode("utf8"), str(c.imag).encode("utf8"),
            mpmath.mp.prec, str(px).encode("utf8"), int(maxiter),
            float(M_divergence))
        if order < -1:
            raise RuntimeError(f"fsb_ball_method_mandelbrot failed ({order})")
        return None if order == -1 else int(order)

    @staticmethod
    def _newton(c, order, eps_pixel, max_newton, eps_cv, any_nucleus):
        if order is None:
            raise ValueError("order shall be defined for Newton method")
        seed_prec = mpmath.mp.prec
        if max_newton is None:
            max_newton = 80
        if eps_cv is None:
            eps_cv = mpmath.mpf(val=(2, -seed_prec))
        cap = int(seed_prec * 0.31) + 64
        bx = ctypes.create_string_buffer(cap)
        by = ctypes.create_string_buffer(cap)
        lib = _native.load_orbit_lib()
        rc = lib.fsb_find_nucleus_mandelbrot(
            str(c.real).encode("utf8"), str(c.imag).encode("utf8"), seed_prec,
            int(order), int(max_newton), str(eps_cv).encode("utf8"),
            str(eps_pixel).encode("utf8"), int(any_nucleus), bx, by, cap)
        if rc < 0:
            raise RuntimeError(f"fsb_find_nucleus_mandelbrot failed ({rc})")
        if rc == 0:
            return False, mpmath.mpc("nan", "nan")
        return True, mpmath.mpc(mpmath.mpf(bx.value.decode()),
                                mpmath.mpf(by.value.decode()))

    @staticmethod
    def find_nucleus(c, order, eps_pixel, max_newton=None, eps_cv=None):
        """ models/mandelbrot_M2.py:654-685 : Newton descent on z_order(c) with
        the roots of the divisors of `order` divided out """
        return Perturbation_mandelbrot._newton(c, order, eps_pixel, max_newton,
                                               eps_cv, False)

    @staticmethod
    def find_any_nucleus(c, order, eps_pixel, max_newton=None, eps_cv=None):
        """ models/mandelbrot_M2.py:688-714 : plain Newton descent """
        return Perturbation_mandelbrot._newton(c, order, eps_pixel, max_newton,
                                               eps_cv, True)

    @calc_options
    def calc_std_div(self, *, calc_name: str, subset, max_iter: int,
                     M_divergence: float, epsilon_stationnary: float,
                     BLA_eps: float = 1e-6, interior_detect: bool = False,
                     calc_dzndc: bool = True, calc_orbit: bool = False,
                     backshift: int = 0):
        complex_codes = ["zn"]
        if interior_detect:
            complex_codes += ["dzndz"]
        if calc_dzndc:
            complex_codes += ["dzndc"]
        if calc_orbit:
            complex_codes += ["zn_orbit"]
        int_codes = ["ref_cycle_iter"]
        stop_codes = ["max_iter", "divergence", "stationnary"]
        BLA_activated = ((BLA_eps is not None)
                         and bool(self.dx < settings.newton_zoom_level))

        def set_state():
            def impl(instance):
                instance.complex_type = np.complex128
                instance.potential_M = M_divergence
                instance.codes = (complex_codes, int_codes, stop_codes)
                instance.calc_dZndz = interior_detect
                instance.calc_dZndc = calc_dzndc
            return impl

        spec = KernelSpec(kind="perturb_M2", max_iter=max_iter,
                          M_divergence=M_divergence,
                          epsilon_stationnary=epsilon_stationnary,
                          BLA_eps=BLA_eps, bla_activated=BLA_activated,
                          calc_dzndc=calc_dzndc, calc_dzndz=interior_detect,
                          calc_orbit=calc_orbit, backshift=backshift)
        return {"set_state": set_state, "initialize": lambda: spec,
                "iterate": lambda: spec}


class Perturbation_mandelbrot_N(PerturbationFractal):
    """ Arbitrary-precision power-N Mandelbrot z -> z^N + c
    (models/mandelbrot_Mn.py:387-742): same perturbation loop, model formulas
    as full binomial expansions.  Reference point: ball method + Newton descent
    (fsb_ball_method_mandelbrot_n / fsb_find_any_nucleus_mandelbrot_n), as the
    reference's model (mandelbrot_Mn.py:744-810). """

    def __init__(self, directory: str, exponent: int):
        super().__init__(directory)
        if int(exponent) != exponent or not (2 <= exponent <= 32):
            raise ValueError("exponent shall be an integer in [2, 32]")
        self.exponent = int(exponent)
        self.potential_kind = "infinity"
        self.potential_d = self.exponent
        self.potential_a_d = 1.
        self.potential_M_cutoff = 1000.
        self.critical_pt = 0.
        self.FP_code = "zn"
        self.holomorphic = True

    def FP_loop(self, NP_orbit, c0):
        """ models/mandelbrot_Mn.py:438-460 -> native MPFR orbit """
        return self._native_orbit(NP_orbit, c0, flavor=None, exponent=self.exponent)

    def _ball_method(self, c, px, maxiter, M_divergence):
        """ models/mandelbrot_Mn.py:744-762 -> fsb_ball_method_mandelbrot_n """
        order = _native.load_orbit_lib().fsb_ball_method_mandelbrot_n(
            self.exponent, str(c.real).encode("utf8"), str(c.imag).encode("utf8"),
            mpmath.mp.prec, str(px).encode("utf8"), int(maxiter), float(M_divergence))
        if order < -1:
            raise RuntimeError(f"fsb_ball_method_mandelbrot_n failed ({order})")
        return None if order == -1 else int(order)

    def find_nucleus(self, c, order, eps_pixel, max_newton=None, eps_cv=None):
        """ models/mandelbrot_Mn.py:765-779 """
        raise NotImplementedError("Divide by undesired roots technique not implemented, "
                                  "Use 'find_any_nucleus'")

    def find_any_nucleus(self, c, order, eps_pixel, max_newton=None, eps_cv=None):
        """ models/mandelbrot_Mn.py:782-810 -> fsb_find_any_nucleus_mandelbrot_n """
        return _native.newton_call("fsb_find_any_nucleus_mandelbrot_n", self.exponent, c, order,
                                   eps_pixel, max_newton, eps_cv)

    @calc_options
    def calc_std_div(self, *, calc_name: str, subset, max_iter: int,
                     M_divergence: float, epsilon_stationnary: float,
                     BLA_eps: float = 1e-6, interior_detect: bool = False,
                     calc_dzndc: bool = True, calc_orbit: bool = False,
                     backshift: int = 0):
        """ models/mandelbrot_Mn.py:463-606.  calc_orbit: `zn ** N` of the back-shift is
        a product chain here, the C library's polar-form power in the reference
        (tolerance-tested against its fixtures) """
        complex_codes = ["zn"]
        if interior_detect:
            complex_codes += ["dzndz"]
        if calc_dzndc:
            complex_codes += ["dzndc"]
        if calc_orbit:
            complex_codes += ["zn_orbit"]
        int_codes = ["ref_cycle_iter"]
        stop_codes = ["max_iter", "divergence", "stationnary"]
        BLA_activated = ((BLA_eps is not None)
                         and bool(self.dx < settings.newton_zoom_level))
        nexp = self.exponent

        def set_state():
            def impl(instance):
                instance.complex_type = np.complex128
                instance.potential_M = M_divergence
                instance.codes = (complex_codes, int_codes, stop_codes)
                instance.calc_dZndz = interior_detect
                instance.calc_dZndc = calc_dzndc
                instance.backshift = backshift if calc_orbit else None
            return impl

        spec = KernelSpec(kind="perturb_M2", nexp=nexp, max_iter=max_iter,
                          M_divergence=M_divergence,
                          epsilon_stationnary=epsilon_stationnary,
                          BLA_eps=BLA_eps, bla_activated=BLA_activated,
                          calc_dzndc=calc_dzndc, calc_dzndz=interior_detect,
                          calc_orbit=calc_orbit, backshift=backshift)
        return {"set_state": set_state, "initialize": lambda: spec,
                "iterate": lambda: spec}


class Perturbation_burning_ship(PerturbationFractal):
    """ Arbitrary-precision Burning-ship family
    (models/burning_ship.py:863-1130) """

    def __init__(self, directory: str, flavor: str = "Burning ship"):
        super().__init__(directory)
        self.flavor = flavor
        get_flavor_int(flavor)
        _set_potential(self)
        self.critical_pt = 0.
        self.FP_code = ["xn", "yn"]
        self.holomorphic = False

    def FP_loop(self, NP_orbit, c0):
        """ models/burning_ship.py:937-968 """
        return self._native_orbit(NP_orbit, c0,
                                  flavor=get_flavor_int(self.flavor))

    def _ball_method(self, c, px, maxiter, M_divergence):
        """ models/burning_ship.py:1167-1186 -> fsb_ball_method_burning_ship """
        order = _native.load_orbit_lib().fsb_ball_method_burning_ship(
            get_flavor_int(self.flavor), str(c.real).encode("utf8"), str(c.imag).encode("utf8"),
            mpmath.mp.prec, str(px).encode("utf8"), int(maxiter), float(M_divergence))
        if order < -1:
            raise RuntimeError(f"fsb_ball_method_burning_ship failed ({order})")
        return None if order == -1 else int(order)

    @staticmethod
    def find_nucleus(c, order, eps_pixel, max_newton=None, eps_cv=None):
        """ models/burning_ship.py:1189-1200 """
        raise NotImplementedError("Divide by undesired roots technique "
                                  "not implemented for burning ship")

    def find_any_nucleus(self, c, order, eps_pixel, max_newton=None, eps_cv=None):
        """ models/burning_ship.py:1203-1232 -> fsb_find_any_nucleus_burning_ship """
        return _native.newton_call("fsb_find_any_nucleus_burning_ship",
                                   get_flavor_int(self.flavor), c, order, eps_pixel,
                                   max_newton, eps_cv)

    @calc_options
    def calc_std_div(self, *, calc_name: str, subset, max_iter: int,
                     M_divergence: float, BLA_eps: float = 1e-6,
                     calc_hessian: bool = True, calc_orbit: bool = False,
                     backshift: int = 0):
        complex_codes = ["xn", "yn"]
        if calc_hessian:
            complex_codes += ["dxnda", "dxndb", "dynda", "dyndb"]
        if calc_orbit:
            complex_codes += ["xn_orbit", "yn_orbit"]
        int_codes = ["ref_cycle_iter"]
        stop_codes = ["max_iter", "divergence"]
        BLA_activated = ((BLA_eps is not None)
                         and bool(self.dx < settings.newton_zoom_level))

        def set_state():
            def impl(instance):
                instance.complex_type = np.float64
                instance.potential_M = M_divergence
                instance.codes = (complex_codes, int_codes, stop_codes)
                instance.backshift = backshift if calc_orbit else None
            return impl

        spec = KernelSpec(kind="perturb_BS", flavor=get_flavor_int(self.flavor),
                          max_iter=max_iter, M_divergence=M_divergence,
                          BLA_eps=BLA_eps, bla_activated=BLA_activated,
                          calc_hessian=calc_hessian, calc_orbit=calc_orbit,
                          backshift=backshift)
        return {"set_state": set_state, "initialize": lambda: spec,
                "iterate": lambda: spec}
