# -*- coding: utf-8 -*-
"""
GPU post-processing kernel against a numpy restatement of the
reference's post-processing classes:
  Continuous_iter_pp   postproc.py:352-406 + Continuous_iter_pp_infinity :1001-1009
  DEM_pp               postproc.py:684-731
  DEM_normal_pp        postproc.py:572-628 (kind "potential"), unskew core.py:3147-3158
and the fused call (pixel kernels + post-processing) against the stand-alone one.
"""
import tempfile

import numpy as np
import pytest

import fractalshades_b200.models as fsm
from fractalshades_b200 import postproc as fpp

pytestmark = pytest.mark.gpu


def np_cont_iter(zn, stop_iter, d, a_d, M, floor_iter=0):
    k = np.abs(a_d) ** (1. / (d - 1.))
    with np.errstate(all="ignore"):
        nu_frac = -(np.log(np.log(np.abs(zn * k)) / np.log(M * k)) / np.log(d))
        nu_div, nu_mod = np.divmod(-nu_frac, 1.)
        nu_frac = -nu_mod
        n = stop_iter - nu_div.astype(stop_iter.dtype)
    return (n - floor_iter) + nu_frac


def np_dem(zn, deriv, holomorphic, px_snap=None):
    abs_zn = np.abs(zn)
    if holomorphic:
        abs_d = np.abs(deriv)
    else:
        dXdA, dXdB, dYdA, dYdB = deriv
        Q = np.hypot(dXdA + dYdB, dXdB - dYdA)
        R = np.hypot(dXdA - dYdB, dXdB + dYdA)
        abs_d = 0.5 * (Q + R)
    with np.errstate(all="ignore"):
        val = abs_zn * np.log(abs_zn) / abs_d
    if px_snap is not None:
        val = np.where(val < px_snap, 0., val)
    return val


def np_normal(zn, deriv, holomorphic, skew=None):
    with np.errstate(all="ignore"):
        if holomorphic:
            normal = zn / deriv
        else:
            dXdA, dXdB, dYdA, dYdB = deriv
            normal = ((dXdA * zn.real + dYdA * zn.imag)
                      + 1j * (dXdB * zn.real + dYdB * zn.imag))
        if skew is not None:
            nx, ny = normal.real.copy(), normal.imag.copy()
            ux = skew[0, 0] * nx + skew[1, 0] * ny
            uy = skew[0, 1] * nx + skew[1, 1] * ny
            normal = ux + 1j * uy
        return normal / np.abs(normal)


def _close(a, b, mask, rtol):
    a, b = np.asarray(a, np.float64)[mask], np.asarray(b, np.float64)[mask]
    ok = np.isfinite(b)
    assert ok.mean() > 0.99
    np.testing.assert_allclose(a[ok], b[ok], rtol=rtol, atol=rtol)


def _m2():
    f = fsm.Perturbation_mandelbrot(tempfile.mkdtemp())
    f.zoom(precision=30, x="-1.74928893611435556407228", y="0.", dx="5.e-20", nx=300,
           xy_ratio=1.5, theta_deg=0.)
    f.calc_std_div(calc_name="c", subset=None, max_iter=50000, M_divergence=1e3,
                   epsilon_stationnary=1e-3, BLA_eps=1e-6, interior_detect=False,
                   calc_dzndc=True)
    return f


def _bs():
    f = fsm.Perturbation_burning_ship(tempfile.mkdtemp(), flavor="Burning ship")
    f.zoom(precision=50, x="0.533551593577038561769721161491702555962775680136595415306315189524970818968817900068355227861158570104764433694",
           y="1.26175074578870311547721223871955368990255513054155186351034363459852900933566891849764050954410207620093433856",
           dx="7.5e-30", nx=260, xy_ratio=1.3, theta_deg=12.0, has_skew=True,
           skew_00=0.9, skew_01=-0.35, skew_10=0.2, skew_11=1.05)
    f.calc_std_div(calc_name="c", subset=None, max_iter=30000, M_divergence=1e3,
                   BLA_eps=1e-6, calc_hessian=True)
    return f


def _raw(f):
    f.calc_raw("c")
    Z = np.array(f.get_data_memmap("c", "Z", mode="r"))
    si = np.array(f.get_data_memmap("c", "stop_iter", mode="r"))[0]
    sr = np.array(f.get_data_memmap("c", "stop_reason", mode="r"))[0]
    return Z, si, sr


@pytest.mark.parametrize("model", ["m2", "bs"])
def test_postproc_kernel_matches_numpy_restatement(model):
    f = _m2() if model == "m2" else _bs()
    Z, si, sr = _raw(f)
    esc = sr == 1
    assert esc.mean() > 0.3
    holo = model == "m2"
    if holo:
        zn, deriv = Z[0], Z[1]
    else:
        zn, deriv = Z[0] + 1j * Z[1], (Z[2], Z[3], Z[4], Z[5])
    skew = getattr(f, "skew", None)
    ref = {"cont_iter": np_cont_iter(zn, si, 2, 1., 1e3, floor_iter=7),
           "DEM": np_dem(zn, deriv, holo, px_snap=None)}
    nrm = np_normal(zn, deriv, holo, None if skew is None else np.asarray(skew))
    ref["normal_x"], ref["normal_y"] = nrm.real, nrm.imag
    out64 = fpp.fields_from_raw(f, "c", Z, si, floor_iter=7, dtype=np.float64)
    for k in fpp.FIELDS:
        _close(out64[k], ref[k], esc, 1e-12)
    out32 = fpp.fields_from_raw(f, "c", Z, si, floor_iter=7, dtype=np.float32)
    for k in fpp.FIELDS:
        assert out32[k].dtype == np.float32
        a, b = out32[k][esc], ref[k].astype(np.float32)[esc]
        ok = np.isfinite(b)
        # one rounding to float32 on each side: equal, or neighbours when the
        # fp64 values straddle a float32 rounding boundary
        assert np.mean(a[ok] == b[ok]) > 0.999
        np.testing.assert_allclose(a[ok], b[ok], rtol=2e-7, atol=1e-7)
    # px_snap
    snap = float(np.nanmedian(ref["DEM"][esc]))
    o = fpp.fields_from_raw(f, "c", Z, si, fields=("DEM",), px_snap=snap, dtype=np.float64)
    _close(o["DEM"], np_dem(zn, deriv, holo, px_snap=snap), esc, 1e-12)


@pytest.mark.parametrize("model", ["m2", "bs"])
def test_fused_frame_fields_equal_standalone(model):
    f = _m2() if model == "m2" else _bs()
    Z, si, sr = _raw(f)
    alone = fpp.fields_from_raw(f, "c", Z, si)
    fused, stats = fpp.frame_fields(f, "c", want_stop_iter=True)
    assert np.array_equal(fused["stop_reason"], sr) and np.array_equal(fused["stop_iter"], si)
    esc = sr == 1
    for k in fpp.FIELDS:
        assert fused[k].dtype == np.float32
        assert np.array_equal(fused[k][esc], alone[k][esc], equal_nan=True), k
    assert stats["sum_stop_iter"] == int(si.sum(dtype=np.int64))
    img = fpp.to_image(f, fused["cont_iter"])
    assert img.shape == (f.ny, f.nx)
    (ix, ixx, iy, iyy) = list(f.chunk_slices())[1]
    off = sum((a[1] - a[0]) * (a[3] - a[2]) for a in list(f.chunk_slices())[:1])
    assert np.array_equal(img[iy:iyy, ix:ixx].ravel(),
                          fused["cont_iter"][off:off + (ixx - ix) * (iyy - iy)], equal_nan=True)


def test_postproc_errors():
    f = fsm.Perturbation_mandelbrot(tempfile.mkdtemp())
    f.zoom(precision=30, x="-1.74928893611435556407228", y="0.", dx="5.e-20", nx=64,
           xy_ratio=1.0, theta_deg=0.)
    f.calc_std_div(calc_name="c", subset=None, max_iter=2000, M_divergence=1e3,
                   epsilon_stationnary=1e-3, BLA_eps=1e-6, interior_detect=False,
                   calc_dzndc=False)
    with pytest.raises(ValueError):
        fpp.frame_fields(f, "c", fields=("DEM",))
    out, _ = fpp.frame_fields(f, "c", fields=("cont_iter",))
    assert set(out) == {"cont_iter", "stop_reason"}


# ---------------------------------------------------------------------------
# k_postproc pinned to the LIVE reference: fixtures generated by
# tools/gen_golden_pp.py drive the reference's own Continuous_iter_pp, DEM_pp and
# DEM_normal_pp (through Fractal.postproc, postproc_dtype float64) on the raw
# arrays of its own calculation; the same raw arrays go through k_postproc.
import json
import os

import parity_common as pc
from cases import CASES

PP_GOLDEN = ["std_M2_cfg1", "p_M2_E20", "p_M2_deep250", "p_BS_f1_E30_skew", "std_BS_f1",
             "std_M2_seahorse_orbit", "p_M2_divref_orbit", "p_BS_f2_E12"]


@pytest.mark.parametrize("name", PP_GOLDEN)
def test_postproc_kernel_matches_reference_fixture(name):
    g = np.load(os.path.join(pc.GOLDEN, f"pp_{name}.npz"))
    meta = json.loads(str(g["meta"]))
    f, case = pc.make_fractal(name)
    f.calc_std_div(calc_name="c", subset=None, **case["calc"])
    codes = list(f._calc_data["c"]["state"].codes[0])
    saved = meta["codes"]
    # the fixture holds the reference's SAVED rows; k_postproc indexes the
    # calculation's rows: put them back in place
    Zs = np.asarray(g["Z"])
    Z = np.zeros((len(codes), Zs.shape[1]), Zs.dtype)
    for i, c in enumerate(saved):
        Z[codes.index(c)] = Zs[i]
    si = np.asarray(g["stop_iter"])
    esc = np.asarray(g["stop_reason"])[0] == 1
    assert esc.sum() > 100
    out64 = fpp.fields_from_raw(f, "c", Z, si, fields=("cont_iter", "DEM", "normal"),
                                dtype=np.float64)
    out32 = fpp.fields_from_raw(f, "c", Z, si, fields=("cont_iter", "DEM", "normal"),
                                dtype=np.float32)
    rates = {}
    for key in ("cont_iter", "DEM", "normal_x", "normal_y"):
        ref = np.asarray(g[key], np.float64)
        ok = esc & np.isfinite(ref)
        assert ok.sum() > 100, key
        # float64 outputs: the same formula evaluated by another libm (log, hypot)
        np.testing.assert_allclose(out64[key][ok], ref[ok], rtol=1e-11, atol=1e-11)
        # float32 outputs (the reference's default postproc_dtype): equal to the
        # reference's float64 value rounded to float32, or its float32 neighbour
        r32 = ref[ok].astype(np.float32)
        same = out32[key][ok] == r32
        nb = np.abs(out32[key][ok].astype(np.float64) - r32) <= np.spacing(np.abs(r32)).astype(np.float64)
        rates[key] = float(same.mean())
        assert same.mean() >= 0.999, (key, same.mean())
        assert nb.all(), key
    # ---- Fieldlines_pp (postproc.py:409-531) and the Blinn coefficients
    # (colors/layers.py:865-903), same fixtures ----
    fl = fpp.Fieldlines_pp(**meta["fieldlines"])
    lighting = fpp.Blinn_lighting(0.2, (1., 1., 1.))
    for ls in meta["lights"]:
        lighting.add_light_source(**ls)
    c_pix = np.asarray(g["c_pix"])
    for dt, tol_fl, tol_lam, tol_sp in ((np.float64, 1e-9, 1e-12, 1e-9), (np.float32, 2e-6, 2e-7, 2e-6)):
        ext = fpp.fields_from_raw(f, "c", Z, si, fields=(), dtype=dt, c_pix=c_pix, fieldlines=fl,
                                  lighting=lighting, max_slope=meta["max_slope"])
        ref = np.asarray(g["fieldlines"], np.float64)
        ok = esc & np.isfinite(ref)
        assert ok.sum() > 100
        # atan2 / sin / hypot of another libm, amplified by the angle doublings of the
        # continued orbit: absolute tolerance on a value of magnitude <= 1
        err = np.abs(ext["fieldlines"][ok].astype(np.float64) - ref[ok])
        assert err.max() <= tol_fl, ("fieldlines", dt, err.max())
        rates[f"fieldlines_maxerr_{np.dtype(dt).name}"] = float(err.max())
        sh = np.asarray(g["shade"], np.float64)
        assert ext["shade"].shape == sh.shape
        for r in range(sh.shape[0]):
            okr = esc & np.isfinite(sh[r])
            tol = tol_lam if r % 2 == 0 else tol_sp
            e2 = np.abs(ext["shade"][r][okr].astype(np.float64) - sh[r][okr])
            assert e2.max() <= tol, ("shade row", r, dt, e2.max())
        assert np.all(ext["shade"][3][esc] == 0.)          # second light: k_specular = 0
    print("\nPP_PARITY", name, json.dumps(rates))
    f._release_indep_args(f._calc_data["c"]["cycle_indep_args"])


@pytest.mark.parametrize("model", ["m2", "bs"])
def test_fused_fieldlines_and_shade_equal_standalone(model):
    """ fsb_frame_run_grid_pp_ext against fsb_postproc_ext_run on the raw arrays """
    f = _m2() if model == "m2" else _bs()
    Z, si, sr = _raw(f)
    fl = fpp.Fieldlines_pp(n_iter=5, swirl=0.5, endpoint_k=0.8)
    lighting = fpp.Blinn_lighting(0.2, (1., 1., 1.), ls0=dict(
        k_diffuse=1.8, k_specular=2., shininess=40., polar_angle=50., azimuth_angle=20.))
    c_pix = np.concatenate([np.ravel(f.chunk_pixel_pos(cs, False, None)) for cs in f.chunk_slices()])
    alone = fpp.fields_from_raw(f, "c", Z, si, c_pix=c_pix, fieldlines=fl, lighting=lighting)
    fused, _ = fpp.frame_fields(f, "c", fieldlines=fl, lighting=lighting)
    esc = sr == 1
    assert np.array_equal(fused["stop_reason"], sr)
    for k in fpp.FIELDS + ("fieldlines",):
        assert np.array_equal(fused[k][esc], alone[k][esc], equal_nan=True), k
    assert fused["shade"].shape == (2, Z.shape[1])
    assert np.array_equal(fused["shade"][:, esc], alone["shade"][:, esc], equal_nan=True)
    a = fused["fieldlines"][esc]
    assert np.isfinite(a).mean() > 0.99 and np.nanstd(a) > 0.05       # a field, not a constant
    lam = fused["shade"][0][esc]
    assert np.nanmin(lam) >= 0. and np.nanmax(lam) <= 1.0000001 and np.nanstd(lam) > 0.01
