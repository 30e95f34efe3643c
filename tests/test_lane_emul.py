# -*- coding: utf-8 -*-
"""
The lane state machine of the event-driven pixel kernel (k_perturb_m2_v2:
`lane_step` + `m2_hot_iter`, fractalshades_b200/csrc/fsb_lane.cuh) checked on
the CPU.  The functions are `__host__ __device__`; tests/emul/ drives them one
pixel at a time (test infrastructure, see tests/emul_lib.py):

  * compiled with FSB_STRICT (= the -fmad=false build) the machine must equal
    the oracle BIT FOR BIT on every case it handles -- stop_reason, stop_iter,
    U and Z: the hot-loop pre-tests, the n_iter = nbase + w bookkeeping, the
    sticky rebase flag, the Xrange <-> fp64 lane changes ... are then all
    proven equivalent to the reference's loop (perturbation.py:1076-1398);
  * with the default build's FMA formulas (libm's exact fma) it must meet the
    north-star tolerance against the oracle.
"""
import numpy as np
import pytest

import parity_common as pc
import emul_lib as el
import oracle_lib as ol
from cases import CASES

M2_CASES = [n for n, c in CASES.items() if c["kind"] == "perturb_M2"]
_CACHE = {}


def _case(name):
    if name not in _CACHE:
        f, case, t = pc.host_tables(name)
        if not el.supported(t):
            _CACHE[name] = None
        else:
            pc.oracle_fill_tables(t)
            c_pix = pc.all_c_pix(f)
            _CACHE[name] = (t, c_pix, ol.perturb(t, c_pix, 0, True))
    return _CACHE[name]


def test_some_cases_are_handled():
    handled = [n for n in M2_CASES if _case(n) is not None]
    # fp64 with and without BLA, Xrange with and without BLA, projections
    assert {"p_M2_E20", "p_M2_deep250", "p_M2_deep1000_xr", "p_M2_deep400_xr_nobla",
            "p_M2_expmap_E55_step"} <= set(handled)


@pytest.mark.parametrize("name", M2_CASES)
def test_strict_machine_bit_exact(name):
    got = _case(name)
    if got is None:
        pytest.skip("variant kept on the general kernel (interior detection, periodic "
                    "reference, calc_orbit, power N)")
    t, c_pix, (Zo, Uo, sro, sio, cnt) = got
    Z, U, sr, si, c = el.perturb(t, c_pix, strict=True)
    assert np.array_equal(sr, sro) and np.array_equal(si, sio) and np.array_equal(U, Uo)
    assert pc.same_bits(Z, Zo)
    # the counters are those of the reference's loop too
    assert c["sum_stop_iter"] == int(sio.sum(dtype=np.int64))
    assert c["n_iter_exec"] + 0 == int(cnt[0])
    assert c["n_bla_steps"] == int(cnt[1]) and c["n_rebase"] == int(cnt[2])
    # every executed iteration ran either in the hot loop or in the event section
    assert c["hot_iterations"] <= c["n_iter_exec"] + c["event_visits"]


@pytest.mark.parametrize("name", M2_CASES)
def test_fma_machine_tolerance(name):
    got = _case(name)
    if got is None:
        pytest.skip("variant kept on the general kernel")
    t, c_pix, (Zo, Uo, sro, sio, cnt) = got
    Z, U, sr, si, c = el.perturb(t, c_pix, strict=False)
    same = (si == sio)[0] & (sr == sro)[0]
    from test_oracle_golden import FAST_FLOOR, NU_FLOOR
    assert same.mean() >= FAST_FLOOR.get(name, 0.999), (name, same.mean())
    esc = same & (sro[0] == 1)
    frac = pc.nu_within("perturb_M2", float(t["M_divergence"]), Z, si, Zo, sio, esc)
    if frac is not None:
        assert frac >= NU_FLOOR.get(name, 0.995), (name, frac)
