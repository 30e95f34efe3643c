# -*- coding: utf-8 -*-
"""
Convert the reference's `cycle_indep_args` tuple (perturbation.py:539-560) to
the plain "tables dict" exchanged with the oracle and the product.
TEST INFRASTRUCTURE (container only; needs the live reference).
"""
import numpy as np


def _xr(arr):
    a = np.asarray(arr)
    return np.array(a["mantissa"]), np.array(a["exp"], dtype=np.int32)


def _xr1(arr):
    m, e = _xr(arr)
    return m.ravel()[0], int(e.ravel()[0])


def proj_from_reference(projection):
    """ the reference's projection object -> the plain "proj" dict """
    import fractalshades as fs
    p = projection
    d = dict(kind=0, dzndc_modifier=0, hmoy=0., k_re=0., k_im=0., mod_param=0.)
    if isinstance(p, fs.projection.Expmap):
        k = complex(p.pix_to_ht)
        hshift = (p.hmoy - p.exp_step_hmoy) if p.use_step else p.hmoy
        d.update(kind=1, dzndc_modifier=1, hmoy=float(p.hmoy), k_re=k.real,
                 k_im=k.imag, mod_param=float(hshift))
    elif isinstance(p, fs.projection.Cartesian):
        if p.expmap_seam is not None:
            d.update(dzndc_modifier=2, mod_param=float(p.expmap_seam))
    else:
        raise ValueError(type(p))
    return d


def tables_from_reference(f, indep):
    """ f: reference Fractal instance after calc_std_div ; indep: its tuple """
    holomorphic = indep[0]
    xr_detect = bool(f.xr_detect_activated)
    t = {"xr_detect": xr_detect, "max_iter": int(f.max_iter),
         "M_divergence": float(f.M_divergence),
         "calc_orbit": bool(getattr(f, "calc_orbit", False)),
         "backshift": int(getattr(f, "backshift", 0) or 0)}
    BLA_eps = getattr(f, "BLA_eps", None)
    t["BLA_eps"] = BLA_eps
    import fractalshades as fs
    t["bla_activated"] = bool((BLA_eps is not None)
                              and (f.dx < fs.settings.newton_zoom_level))
    if holomorphic:
        (_, _init, _iter, Zn_path, dZndc_path, dZndz_path, has_xr,
         ref_index_xr, ref_xr, ref_div_iter, ref_order, drift_xr, dx_xr,
         _proj, lin_mat, lin_scale_xr, kc, M_bla, r_bla, bla_len, stages_bla,
         _mod, _intr) = indep
        t["kind"] = "perturb_M2"
        t["nexp"] = (int(f.exponent)
                     if type(f).__name__ == "Perturbation_mandelbrot_N" else 0)
        t["epsilon_stationnary"] = float(f.epsilon_stationnary)
        t["calc_dzndc"] = bool(f.calc_dZndc)
        t["calc_dzndz"] = bool(f.calc_dZndz)
        t["Zn_path"] = np.array(Zn_path)
        for name, p in (("dZndc", dZndc_path), ("dZndz", dZndz_path)):
            if p is None:
                t[name] = None
                t[name + "_e"] = None
            elif xr_detect:
                t[name], t[name + "_e"] = _xr(p)
            else:
                t[name], t[name + "_e"] = np.array(p), None
        t["ref_index_xr"] = np.array(ref_index_xr, np.int32) if has_xr else None
        if has_xr:
            t["ref_xr"], t["ref_xr_e"] = _xr(ref_xr)
        else:
            t["ref_xr"], t["ref_xr_e"] = None, None
        t["drift"], t["drift_e"] = _xr1(drift_xr)
    else:
        (_, _init, _iter, Zn_path, dXnda, dXndb, dYnda, dYndb, has_xr,
         ref_index_xr, refx_xr, refy_xr, ref_div_iter, ref_order, driftx_xr,
         drifty_xr, dx_xr, _proj, lin_mat, lin_scale_xr, kc, M_bla, r_bla,
         bla_len, stages_bla, _mod, _intr) = indep
        t["kind"] = "perturb_BS"
        import fractalshades.models.burning_ship as bs
        t["flavor"] = int(bs.get_flavor_int(f.flavor))
        t["calc_hessian"] = bool(f.calc_hessian)
        t["Zn_path"] = np.array(Zn_path)
        for name, p in (("dXnda", dXnda), ("dXndb", dXndb), ("dYnda", dYnda),
                        ("dYndb", dYndb)):
            if p is None:
                t[name] = None
                t[name + "_e"] = None
            elif xr_detect:
                t[name], t[name + "_e"] = _xr(p)
            else:
                t[name], t[name + "_e"] = np.array(p), None
        t["ref_index_xr"] = np.array(ref_index_xr, np.int32) if has_xr else None
        if has_xr:
            t["refx_xr"], t["refx_xr_e"] = _xr(refx_xr)
            t["refy_xr"], t["refy_xr_e"] = _xr(refy_xr)
            # reference stores them with a complex dtype mantissa
            t["refx_xr"] = np.real(t["refx_xr"]).astype(np.float64)
            t["refy_xr"] = np.real(t["refy_xr"]).astype(np.float64)
        else:
            t["refx_xr"] = t["refx_xr_e"] = t["refy_xr"] = t["refy_xr_e"] = None
        t["driftx"], t["driftx_e"] = _xr1(driftx_xr)
        t["drifty"], t["drifty_e"] = _xr1(drifty_xr)
    t["proj"] = proj_from_reference(f.projection)
    t["ref_div_iter"] = int(ref_div_iter)
    t["ref_order"] = int(ref_order)
    t["lin_mat"] = np.array(lin_mat, np.float64)
    t["lin_scale"], t["lin_scale_e"] = _xr1(lin_scale_xr)
    t["dx"], t["dx_e"] = _xr1(dx_xr)
    t["kc"], t["kc_e"] = _xr1(kc)
    if M_bla is None:
        t["M_bla"] = t["r_bla"] = None
        t["bla_len"] = 0
        t["stages_bla"] = 0
    else:
        t["M_bla"] = np.array(M_bla)
        t["r_bla"] = np.array(r_bla)
        t["bla_len"] = int(bla_len)
        t["stages_bla"] = int(stages_bla)
    return t
