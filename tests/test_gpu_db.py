# -*- coding: utf-8 -*-
"""
Database writer (fractalshades_b200/db.py) against `.db` files written
by the LIVE reference: tools/gen_golden_expdb.py runs the reference's own
Fractal_plotter.save_db (core.py:812-889) -> save_expdb_by_steps (:891-952) on two
exponential maps (holomorphic and burning ship) and keeps the (n_posts, ny, nx)
float32 memmap; here the same database is produced on the GPU: stepped driver
(set_exp_zoom_step + reset_bla_tree per step), fused pixel + post-processing call,
projection derivative (Expmap.df / dfBS, projection.py:375-453) applied to DEM and
normals.

Tolerance: the reference's loops are numba fastmath and chaotic points near the
boundary differ between two compilations of the same formulas (tests/
test_oracle_golden.py); the assertions are on the fraction of pixels that agree.
"""
import json
import os
import tempfile

import numpy as np
import pytest
from numpy.lib.format import open_memmap

import parity_common as pc
from fractalshades_b200 import db as fdb
from fractalshades_b200 import settings

DB_GOLDEN = ["p_M2_expmap_E55_horiz", "p_BS_f1_expmap_E30"]


def _fractal(name, chunk):
    settings.chunk_size = chunk
    f, case = pc.make_fractal(name)
    f.calc_std_div(calc_name="c", subset=None, **case["calc"])
    return f


@pytest.fixture
def chunk_guard():
    old, old_newton = settings.chunk_size, settings.no_newton
    settings.no_newton = True
    yield
    settings.chunk_size, settings.no_newton = old, old_newton


@pytest.mark.gpu
@pytest.mark.parametrize("name", DB_GOLDEN)
def test_expdb_matches_reference_db(name, chunk_guard):
    g = np.load(os.path.join(pc.GOLDEN, f"expdb_{name}.npz"))
    meta = json.loads(str(g["meta"]))
    ref = np.asarray(g["db"])
    f = _fractal(name, meta["chunk_size"])
    assert (f.nx, f.ny) == (meta["nx"], meta["ny"])
    w = fdb.Db_writer(f, "c", fields=("cont_iter", "DEM", "normal"))
    assert w.postnames == meta["posts"]
    path = w.save_db()
    assert os.path.relpath(path, f.directory) == meta["relpath"]
    got = np.array(open_memmap(path, mode="r"))
    status = np.array(open_memmap(w.status_path(path), mode="r"))
    assert got.shape == ref.shape and got.dtype == ref.dtype == np.float32
    assert status.tolist() == np.asarray(g["status"]).tolist() and status.all()
    assert w.n_steps == f.chunks_count          # one step per tile column here
    info = open(path + ".info").read()
    assert [l.strip() for l in info.split("*fields description*")[1].split()] == meta["posts"]
    ok = np.isfinite(ref).all(axis=0) & np.isfinite(got).all(axis=0)
    assert ok.mean() > 0.95
    rates = {}
    d = np.abs(got[0] - ref[0])[ok]
    rates["cont_iter"] = float((d <= 1e-3 * np.maximum(1., np.abs(ref[0][ok]) * 1e-3)).mean())
    with np.errstate(all="ignore"):
        rates["DEM"] = float((np.abs(got[1] - ref[1])[ok] <= 1e-3 * np.abs(ref[1][ok]) + 1e-30).mean())
    rates["normal"] = float(((np.abs(got[2] - ref[2])[ok] <= 2e-3)
                             & (np.abs(got[3] - ref[3])[ok] <= 2e-3)).mean())
    print("\nEXPDB_PARITY", name, json.dumps(rates))
    assert rates["cont_iter"] >= 0.99 and rates["DEM"] >= 0.98 and rates["normal"] >= 0.98
    f._release_indep_args(f._calc_data["c"]["cycle_indep_args"])


@pytest.mark.gpu
def test_db_recovery_and_postdb(chunk_guard):
    name = "p_M2_expmap_E55_horiz"
    f = _fractal(name, 50)
    w = fdb.Db_writer(f, "c", fields=("cont_iter", "DEM"))
    path = w.save_db(relpath="expmap.db")
    full = np.array(open_memmap(path, mode="r"))
    # an interrupted run: two tiles lost, the others kept
    st = open_memmap(w.status_path(path), mode="r+")
    mm = open_memmap(path, mode="r+")
    lost = [list(f.chunk_slices())[k] for k in (2, 5)]
    for cs in lost:
        st[f.chunk_rank(cs)] = 0
        mm[:, cs[2]:cs[3], cs[0]:cs[1]] = -1.
    st.flush(); mm.flush()
    del st, mm
    path2 = w.save_db(relpath="expmap.db", recovery_mode=True)
    assert path2 == path and w.n_steps == 2 and len(w.last_stats) == 2
    again = np.array(open_memmap(path, mode="r"))
    assert np.array_equal(again, full, equal_nan=True)
    # without recovery the database is rebuilt from scratch
    w.save_db(relpath="expmap.db", recovery_mode=False)
    assert w.n_steps == f.chunks_count
    # .postdb: one layer frozen as pixels
    layer = fdb.Grey_layer("cont_iter", func=np.log, probes_z=(2., 9.),
                           colors=[(0., 0., 0.2), (1., 0.8, 0.1), (1., 1., 1.)],
                           mask_color=(0.1, 0.1, 0.1))
    p = w.save_db(postdb_layer=layer)
    assert p.endswith("cont_iter.postdb")
    px = np.array(open_memmap(p, mode="r"))
    assert px.shape == (f.ny, f.nx, 3) and px.dtype == np.uint8
    assert (px == layer.pixels(full[0])).all(axis=-1).mean() > 0.9      # masked points aside
    assert px.std() > 10
    f._release_indep_args(f._calc_data["c"]["cycle_indep_args"])


@pytest.mark.gpu
def test_cartesian_db_single_pass(chunk_guard):
    f = _fractal("p_M2_deep250", 200)
    w = fdb.Db_writer(f, "c", fields=("cont_iter", "DEM", "normal"))
    path = w.save_db()
    got = np.array(open_memmap(path, mode="r"))
    assert got.shape == (4, f.ny, f.nx) and w.n_steps == 1
    from fractalshades_b200 import postproc as fpp
    out, _ = fpp.frame_fields(f, "c")
    assert np.array_equal(got[0], fpp.to_image(f, out["cont_iter"]), equal_nan=True)
    f._release_indep_args(f._calc_data["c"]["cycle_indep_args"])


# ---- host logic (no GPU) ------------------------------------------------------
def test_exp_steps_follow_the_reference_arithmetic(chunk_guard):
    """ core.py:907-921 on the fixture's geometry """
    settings.chunk_size = 50
    f, case = pc.make_fractal("p_M2_expmap_E55_horiz")
    w = fdb.Db_writer(f, "c")
    proj = f.projection
    steps = list(w.exp_steps())
    nh, stp = proj.nh(f), proj.nt(f)
    assert (nh, stp) == (f.nx, f.ny)
    assert [s[0] for s in steps] == list(range(0, nh + 1, stp))
    r, _, hmax_s, hmin_s = steps[3]
    i_max, i_min = min(r + stp, nh), max(r - 50, 0)
    assert hmax_s == (proj.hmax * i_max + proj.hmin * (nh - i_max)) / nh
    assert hmin_s == (proj.hmax * i_min + proj.hmin * (nh - i_min)) / nh
    # every tile ends in exactly one step
    for (ix, ixx, iy, iyy) in f.chunk_slices():
        assert sum(1 for s in steps if s[0] < ixx <= s[0] + s[1]) == 1


def test_grey_layer_pixels():
    layer = fdb.Grey_layer("x", probes_z=(0., 2.))
    a = np.array([[0., 1., 2., 3., 4., -1.]])
    px = layer.pixels(a)
    assert px.shape == (1, 6, 1) and px.dtype == np.uint8
    # triangle wave of period 2 probes: 0, .5, 1, .5, 0, .5
    assert px[0, :, 0].tolist() == [0, 127, 255, 127, 0, 127]
    rgb = fdb.Grey_layer("x", probes_z=(0., 1.), colors=[(0, 0, 0), (1, 0.5, 0)], mask_color=(1., 1., 1.))
    px = rgb.pixels(np.array([[0., 1., 0.5]]), np.array([[False, False, True]]))
    assert px[0, 0].tolist() == [0, 0, 0] and px[0, 1].tolist() == [255, 127, 0] and px[0, 2].tolist() == [255, 255, 255]


def test_projection_df_kinds():
    from fractalshades_b200 import postproc as fpp, projection as prj
    assert fpp.projection_df(prj.Cartesian()) == (0, 0j)
    e = prj.Expmap(0., 10., rotates_df=True, orientation="horizontal")
    assert fpp.projection_df(e) == (2, complex(e.pix_to_ht))
    e.set_exp_zoom_step(4., 2.)
    assert fpp.projection_df(e)[0] == 1
    e2 = prj.Expmap(0., 10., rotates_df=False)
    assert fpp.projection_df(e2)[0] == 3
    e2.set_exp_zoom_step(4., 2.)
    assert fpp.projection_df(e2)[0] == 0


@pytest.mark.gpu
def test_db_of_a_standard_model(chunk_guard):
    """ double-precision models go through the raw seam + the stand-alone post-processing """
    f = _fractal("std_M2_seahorse_orbit", 32)
    from fractalshades_b200 import postproc as fpp
    w = fdb.Db_writer(f, "c", fields=("cont_iter", "DEM", "normal"),
                      fieldlines=fpp.Fieldlines_pp(n_iter=3))
    path = w.save_db(relpath="std.db")
    got = np.array(open_memmap(path, mode="r"))
    assert got.shape == (5, f.ny, f.nx) and w.postnames[-1] == "fieldlines"
    # same fields as the reference's post-processing of the same case (pp fixture)
    g = np.load(os.path.join(pc.GOLDEN, "pp_std_M2_seahorse_orbit.npz"))
    assert f.chunks_count > 1 and w.n_steps == 1
    settings.chunk_size = 200                    # the fixture's tiling: one tile
    f1, _ = pc.make_fractal("std_M2_seahorse_orbit")
    assert f1.chunks_count == 1
    esc = fpp.to_image(f1, np.asarray(g["stop_reason"])[0]) == 1
    ref = fpp.to_image(f1, np.asarray(g["cont_iter"], np.float32))
    # (another tiling: the pixel abscissas of a tile are its own linspace, a last-bit
    # difference that the boundary pixels of this seahorse view amplify: 99.8 %)
    same = got[0][esc] == ref[esc]
    assert np.mean(same) > 0.995
    ref_dem = fpp.to_image(f1, np.asarray(g["DEM"], np.float64))
    rel = np.abs(got[1][esc][same] - ref_dem[esc][same]) / np.abs(ref_dem[esc][same])
    assert np.mean(rel < 1e-4) > 0.99, (np.mean(rel < 1e-4), np.nanmax(rel))
    ref_fl = fpp.to_image(f1, np.asarray(g["fieldlines"], np.float64))
    # (the fixture's field lines use other parameters: only the shapes are comparable)
    assert np.isfinite(got[4][esc]).all() and ref_fl.shape == got[4].shape
