# -*- coding: utf-8 -*-
"""
Parity cases shared by tools/gen_golden.py (which runs them through the live
reference) and the test-suite (which re-runs them through the oracle and the
CUDA path).  Views are those of the reference's own tests / examples
(tests/test_perturbation.py, examples/batch_mode/*) at reduced `nx` so that
the oracle finishes in seconds and the fixtures stay small.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fractalshades_b200.views import VIEWS  # noqa: E402

_STD = dict(M_divergence=1e3, epsilon_stationnary=1e-3)

CASES = {
    # ---- standard escape-time loop (BASELINE config 1) ----
    "std_M2_cfg1": dict(
        kind="std_M2", x=-1.0, y=0.0, dx=5.0, nx=64,
        calc=dict(max_iter=5000, M_divergence=1000., epsilon_stationnary=1e-3)),
    "std_M2_seahorse_orbit": dict(
        kind="std_M2", x=-0.746223962861, y=-0.0959468433527, dx=0.00745,
        nx=64, theta_deg=20.,
        calc=dict(max_iter=5000, M_divergence=1000., epsilon_stationnary=1e-3,
                  calc_d2zndc2=True, calc_orbit=True, backshift=3)),
    # ---- perturbation, fp64 range ----
    "p_M2_E20": dict(
        kind="perturb_M2", precision=30, x="-1.74928893611435556407228",
        y="0.", dx="5.e-20", nx=64,
        calc=dict(max_iter=50000, BLA_eps=1e-6, interior_detect=False,
                  calc_dzndc=True, **_STD)),
    # reference's default flow: ball method + Newton -> periodic reference orbit
    "p_M2_E20_newton": dict(
        kind="perturb_M2", precision=30, x="-1.74928893611435556407228",
        y="0.", dx="5.e-20", nx=64, newton=True,
        calc=dict(max_iter=50000, BLA_eps=1e-6, interior_detect=False,
                  calc_dzndc=True, **_STD)),
    "p_M2_int_E11_newton": dict(
        kind="perturb_M2", precision=17, x="-1.74920463345912691e+00",
        y="-2.8684660237361114e-04", dx="5e-12", nx=64, newton=True,
        calc=dict(max_iter=50000, BLA_eps=1e-6, interior_detect=True,
                  calc_dzndc=True, **_STD)),
    "p_M2_E20_nobla_interior": dict(
        kind="perturb_M2", precision=30, x="-1.74928893611435556407228",
        y="0.", dx="5.e-20", nx=48,
        calc=dict(max_iter=20000, BLA_eps=None, interior_detect=True,
                  calc_dzndc=True, **_STD)),
    "p_M2_int_E11": dict(
        kind="perturb_M2", precision=17, x="-1.74920463345912691e+00",
        y="-2.8684660237361114e-04", dx="5e-12", nx=64,
        calc=dict(max_iter=50000, BLA_eps=1e-6, interior_detect=True,
                  calc_dzndc=False, **_STD)),
    "p_M2_divref_orbit": dict(
        kind="perturb_M2", precision=18, x="-1.36768994867991128",
        y="0.00949048853859240532", dx="2.477633848347765e-8", nx=64,
        xy_ratio=1.6, theta_deg=30.,
        calc=dict(max_iter=20000, BLA_eps=1e-6, interior_detect=False,
                  calc_dzndc=True, calc_orbit=True, backshift=2, **_STD)),
    "p_M2_flake": dict(
        kind="perturb_M2", precision=200, x=VIEWS["glitch_dyn"]["x"],
        y=VIEWS["glitch_dyn"]["y"], dx="1.8e-157", nx=64, xy_ratio=16 / 9.,
        calc=dict(max_iter=50000, BLA_eps=1e-6, interior_detect=False,
                  calc_dzndc=True, **_STD)),
    "p_M2_shallow": dict(
        kind="perturb_M2", precision=12, x="-0.75", y="0.1", dx="0.01", nx=48,
        calc=dict(max_iter=2000, BLA_eps=1e-6, interior_detect=False,
                  calc_dzndc=True, **_STD)),
    "p_M2_E213": dict(
        kind="perturb_M2", precision=224, x=VIEWS["M2_E213"]["x"],
        y=VIEWS["M2_E213"]["y"], dx="3.226224123547768e-213", nx=48,
        calc=dict(max_iter=350000, BLA_eps=1e-6, interior_detect=False,
                  calc_dzndc=True, **_STD)),
    "p_M2_deep250": dict(
        kind="perturb_M2", precision=270, x=VIEWS["deep_julia_2608"]["x"][:290],
        y=VIEWS["deep_julia_2608"]["y"][:290], dx="1e-250", nx=48,
        xy_ratio=16 / 9.,
        calc=dict(max_iter=200000, BLA_eps=1e-6, interior_detect=False,
                  calc_dzndc=True, **_STD)),
    # ---- perturbation, Xrange path (dx < 1e-300) ----
    "p_M2_ultradeep_xr": dict(
        kind="perturb_M2", precision=550,
        x=VIEWS["ultradeep_interior_detect"]["x"],
        y=VIEWS["ultradeep_interior_detect"]["y"], dx="5.06722630e-433",
        nx=48,
        calc=dict(max_iter=200000, BLA_eps=1e-6, interior_detect=True,
                  calc_dzndc=True, **_STD)),
    "p_M2_deep1000_xr": dict(
        kind="perturb_M2", precision=1020,
        x=VIEWS["deep_julia_2608"]["x"][:1040],
        y=VIEWS["deep_julia_2608"]["y"][:1040], dx="1e-1000", nx=32,
        xy_ratio=16 / 9.,
        calc=dict(max_iter=1000000, BLA_eps=1e-6, interior_detect=False,
                  calc_dzndc=True, **_STD)),
    "p_M2_deep400_xr_nobla": dict(
        kind="perturb_M2", precision=420,
        x=VIEWS["deep_julia_2608"]["x"][:440],
        y=VIEWS["deep_julia_2608"]["y"][:440], dx="1e-400", nx=24,
        calc=dict(max_iter=250000, BLA_eps=None, interior_detect=False,
                  calc_dzndc=True, **_STD)),
}

# Burning-ship family, standard loop: all five flavours
for _i, _fl in enumerate(("Burning ship", "Perpendicular burning ship",
                          "Shark fin", "Celtic", "Buffalo")):
    CASES[f"std_BS_f{_i + 1}"] = dict(
        kind="std_BS", init=dict(flavor=_fl), x=-0.5, y=-0.5, dx=3.0, nx=48,
        theta_deg=10. * _i,
        calc=dict(max_iter=800, M_divergence=1000.,
                  calc_orbit=(_i == 0), backshift=(2 if _i == 0 else 0)))

_BS = VIEWS["bs_deep_julia_2430"]
# Burning-ship family, perturbation
CASES["p_BS_f1_E30_skew"] = dict(
    kind="perturb_BS", init=dict(flavor="Burning ship"), precision=50,
    x=_BS["x"][:60], y=_BS["y"][:60], dx="1e-30", nx=48, xy_ratio=1.8,
    theta_deg=12.0, skew=_BS["skew"],
    calc=dict(max_iter=30000, M_divergence=1e3, BLA_eps=1e-6,
              calc_hessian=True))
CASES["p_BS_f1_E500_xr"] = dict(
    kind="perturb_BS", init=dict(flavor="Burning ship"), precision=520,
    x=_BS["x"][:540], y=_BS["y"][:540], dx="1e-500", nx=32, xy_ratio=1.8,
    theta_deg=12.0, skew=_BS["skew"],
    calc=dict(max_iter=100000, M_divergence=1e3, BLA_eps=1e-6,
              calc_hessian=True))
# boundary points found by bisection with the standard-loop oracle
_BS_PTS = {
    1: ("-1.7505941429008662", "0.0214423020148999"),
    2: ("-1.3604879916723847", "0.0015808658322309468"),
    3: ("-1.4477399868839198", "-0.6048320439477123"),
    4: ("-1.760370697674034", "0.011733974791909326"),
    5: ("-1.758364745737221", "0.024352431136909887"),
}
for _i, _fl in enumerate(("Perpendicular burning ship", "Shark fin", "Celtic",
                          "Buffalo")):
    CASES[f"p_BS_f{_i + 2}_E12"] = dict(
        kind="perturb_BS", init=dict(flavor=_fl), precision=30,
        x=_BS_PTS[_i + 2][0], y=_BS_PTS[_i + 2][1], dx="1e-12", nx=32,
        calc=dict(max_iter=20000, M_divergence=1e3, BLA_eps=1e-6,
                  calc_hessian=True, calc_orbit=(_i == 0),
                  backshift=(2 if _i == 0 else 0)))
# the reference's default flow for the family: ball method + Newton descent with
# the full Jacobian (FP_loop.pyx:2357-2755) -> periodic reference (order 156 here)
CASES["p_BS_f1_E12_newton"] = dict(
    kind="perturb_BS", init=dict(flavor="Burning ship"), precision=30,
    x=_BS_PTS[1][0], y=_BS_PTS[1][1], dx="1e-12", nx=32, newton=True,
    calc=dict(max_iter=20000, M_divergence=1e3, BLA_eps=1e-6,
              calc_hessian=True))
CASES["p_BS_f1_E12_nohess_nobla"] = dict(
    kind="perturb_BS", init=dict(flavor="Burning ship"), precision=30,
    x=_BS_PTS[1][0], y=_BS_PTS[1][1], dx="1e-12", nx=32,
    calc=dict(max_iter=20000, M_divergence=1e3, BLA_eps=None,
              calc_hessian=False))

# Xrange depth for flavours 2-5: boundary points refined to 370 digits by
# tools/find_bs_points.py (bisection with the native MPFR orbit)
_BS_PTS_DEEP = {
    2: ("-1.360487991749726900215933412585413412285941436379158985277606966498635261716098622422867771682228033068535518935593947357076960560050199605218706060749680904745216564414856079349951577490144456181921615837643147118666940547732646997181518864846452051695204455670188158206358173731996222868673070844837129912878140443866352112601196075982176312956362174529329299904832415",
        "0.0015808658322309468"),
    3: ("-1.447739947356317556357978548237778495646369402685350135930601575961449773077529696363655060114379075313955548508370405603817016044948045329721346977006988758372330218753579780494738912613058261725359625690988419821558257481687967166902943952736767044387784562930685975780898950021368075016650108359290776866492566128621467559846151414854370939045478590115324715970512931",
        "-0.6048320439477123"),
    4: ("-1.760370697674033990629647602248646769272164196525861499300163149203719957897864439338322827659524966514346500657140812071286660636210128949256738905621205132628675321665558581286312562176520968463586795207478191410050785873347963484571032346104172548153073753678276038822146254247271259723143412395751803705107699499051576986032176207405852407046871337707076056660839532",
        "0.011733974791909326"),
    5: ("-1.758364745737221",
        "0.02435243123671457412747096729151485838826829360905491589504265390060499273229258730046411529559729038829435562100392963917617537209550025441909063863372456105101734638469804602015943204033310574818211732186262196166685204744482058449168705825576193112028437495897107990981970209824756389049187037777424330710506891329508576318294380939514132084838644631572261730179073714"),
}
for _i, _fl in enumerate(("Perpendicular burning ship", "Shark fin", "Celtic",
                          "Buffalo")):
    CASES[f"p_BS_f{_i + 2}_E330_xr"] = dict(
        kind="perturb_BS", init=dict(flavor=_fl), precision=345,
        x=_BS_PTS_DEEP[_i + 2][0], y=_BS_PTS_DEEP[_i + 2][1], dx="1e-330", nx=32,
        calc=dict(max_iter=30000, M_divergence=1e3, BLA_eps=1e-6,
                  calc_hessian=True))


# ---- projections: Expmap and the Cartesian expmap seam ----
def make_projection(mod, spec):
    """ Build the case's projection with `mod` = the reference's
    fractalshades.projection or fractalshades_b200.projection """
    if spec is None:
        return mod.Cartesian()
    spec = dict(spec)
    kind = spec.pop("kind")
    spec.pop("step", None)
    return {"expmap": mod.Expmap, "cartesian": mod.Cartesian}[kind](**spec)


import math as _math   # noqa: E402

# examples/projections/P04-deep_expmap.py (vertical, 20 decades)
CASES["p_M2_expmap_E20_vert"] = dict(
    kind="perturb_M2", precision=31, x="-0.18476527944640054234980108927",
    y="1.0532419344392547587734377701", dx="7.603772829116657e-20", nx=200,
    xy_ratio=1.6666,
    proj=dict(kind="expmap", hmin=0.0, hmax=_math.log(1.e20) + 0.3,
              orientation="vertical"),
    calc=dict(max_iter=20000, M_divergence=1000.0, epsilon_stationnary=0.01,
              BLA_eps=1e-6, interior_detect=True, calc_dzndc=True))
# examples/movies/with_DEM/zoom_script_DEM.py: 55 decades, rotates_df=False,
# here as one frame (horizontal) and as one step of the stepped flow
_DEM_X = "-1.929319698524937920226708049698305350754670432084006734339806946"
_DEM_Y = "-0.0000000000000000007592779387989739090287550144163328879329853232537252481600401185"
CASES["p_M2_expmap_E55_horiz"] = dict(
    kind="perturb_M2", precision=70, x=_DEM_X, y=_DEM_Y,
    dx="7.032184999234219e-55", nx=400, xy_ratio=1.0,
    proj=dict(kind="expmap", hmin=0.0, hmax=127.5, rotates_df=False,
              orientation="horizontal"),
    calc=dict(max_iter=20000, M_divergence=1000.0, epsilon_stationnary=0.001,
              BLA_eps=1e-6, interior_detect=False, calc_dzndc=True))
CASES["p_M2_expmap_E55_step"] = dict(
    kind="perturb_M2", precision=70, x=_DEM_X, y=_DEM_Y,
    dx="7.032184999234219e-55", nx=400, xy_ratio=1.0,
    proj=dict(kind="expmap", hmin=0.0, hmax=127.5, rotates_df=False,
              orientation="horizontal", step=(20.0, 10.0)),
    calc=dict(max_iter=20000, M_divergence=1000.0, epsilon_stationnary=0.001,
              BLA_eps=1e-6, interior_detect=False, calc_dzndc=True))
# the Cartesian end of the same movie: derivative seam
CASES["p_M2_seam_E55"] = dict(
    kind="perturb_M2", precision=70, x=_DEM_X, y=_DEM_Y,
    dx="1.4064369998468438e-54", nx=64, xy_ratio=1.0,
    proj=dict(kind="cartesian", expmap_seam=1.0),
    calc=dict(max_iter=20000, M_divergence=1000.0, epsilon_stationnary=0.001,
              BLA_eps=1e-6, interior_detect=False, calc_dzndc=True))
# burning ship: the modifier scales the four Jacobian rows
CASES["p_BS_f1_expmap_E30"] = dict(
    kind="perturb_BS", init=dict(flavor="Burning ship"), precision=50,
    x=_BS["x"][:60], y=_BS["y"][:60], dx="1e-30", nx=160, xy_ratio=1.0,
    theta_deg=12.0, skew=_BS["skew"],
    proj=dict(kind="expmap", hmin=0.0, hmax=30.0, orientation="horizontal"),
    calc=dict(max_iter=30000, M_divergence=1e3, BLA_eps=1e-6,
              calc_hessian=True))
# standard loops: examples/projections/P01-feigenbaum_expmap.py view
CASES["std_M2_expmap"] = dict(
    kind="std_M2", x=-1.40115519, y=0.0, dx=1e-06, nx=200, xy_ratio=1.0,
    proj=dict(kind="expmap", hmin=0.0, hmax=_math.log(1.e7) + 0.3),
    calc=dict(max_iter=20000, M_divergence=1000., epsilon_stationnary=0.01))
CASES["std_BS_f1_expmap"] = dict(
    kind="std_BS", init=dict(flavor="Burning ship"), x=-1.75, y=-0.03, dx=1e-3,
    nx=120, xy_ratio=1.0,
    proj=dict(kind="expmap", hmin=0.0, hmax=7.0, orientation="vertical"),
    calc=dict(max_iter=800, M_divergence=1000.))


# ---- Perturbation_mandelbrot_N (models/mandelbrot_Mn.py:387-742): boundary
# points found by tools/find_mn_points.py.  kind "perturb_M2" (same loop, same
# tables); init["exponent"] selects the power-N class.
_MN_PTS = {
    3: ("0.21480616146988719579945680880778255",
        "0.72961232293977439159891361761556509"),
    4: ("-0.65487208557069640038326065370157834",
        "0.46859010696337050047907581712697293"),
    5: ("0.351110383579925899931787250019077539000760490700596205803325571051937185272628815179798649674656372116452474161818304145015560421660991822330770457612782251125566412916330544429869243478142030281842380541534604809596919425096614823236551004761306001432908634047555106974113284791152784436351615197467046056965649546391424679397423614881773389893555234761885894384434092268284",
        "0.701480511439901199909049666692103385334347320934128274404434094735916247030171753573064866232875162821936632215757738860020747228881322429774360610150376334834088550555107392573158991304189373709123174055379473079462559233462153097648734673015074668577211512063406809298817713054870379248468820263289394742620866061855232905863231486509031186524740313015847859179245456357711"),
}
CASES["p_M3_E20"] = dict(
    kind="perturb_M2", init=dict(exponent=3), precision=40, x=_MN_PTS[3][0],
    y=_MN_PTS[3][1], dx="3.e-20", nx=64, xy_ratio=1.25, theta_deg=25.,
    calc=dict(max_iter=20000, BLA_eps=1e-6, interior_detect=False,
              calc_dzndc=True, **_STD))
# the reference's default flow for z^3 + c: ball method + Newton (FP_loop.pyx:631-758,
# 1159-1340) -> periodic reference
CASES["p_M3_E20_newton"] = dict(
    kind="perturb_M2", init=dict(exponent=3), precision=40, x=_MN_PTS[3][0],
    y=_MN_PTS[3][1], dx="3.e-20", nx=64, xy_ratio=1.25, theta_deg=25., newton=True,
    calc=dict(max_iter=20000, BLA_eps=1e-6, interior_detect=False,
              calc_dzndc=True, **_STD))
# calc_orbit for z^N + c: the back-shift runs zn_iterate = zn ** N + c
CASES["p_M3_E20_orbit"] = dict(
    kind="perturb_M2", init=dict(exponent=3), precision=40, x=_MN_PTS[3][0],
    y=_MN_PTS[3][1], dx="3.e-20", nx=48, xy_ratio=1.25, theta_deg=25.,
    calc=dict(max_iter=20000, BLA_eps=1e-6, interior_detect=False,
              calc_dzndc=True, calc_orbit=True, backshift=3, **_STD))
CASES["p_M4_E18_interior"] = dict(
    kind="perturb_M2", init=dict(exponent=4), precision=40, x=_MN_PTS[4][0],
    y=_MN_PTS[4][1], dx="2.e-18", nx=48,
    calc=dict(max_iter=20000, BLA_eps=1e-6, interior_detect=True,
              calc_dzndc=True, **_STD))
CASES["p_M3_E20_nobla"] = dict(
    kind="perturb_M2", init=dict(exponent=3), precision=40, x=_MN_PTS[3][0],
    y=_MN_PTS[3][1], dx="3.e-20", nx=32,
    calc=dict(max_iter=8000, BLA_eps=None, interior_detect=True,
              calc_dzndc=False, **_STD))
# Xrange depth: the reference cannot compile the derivative closures there
# (`k * ref_zn_pk`, mandelbrot_Mn.py:712, is int64 x Xrange, which numba_xr's
# mul overload rejects), so the reference-pinned case has no derivative
CASES["p_M5_E340_xr"] = dict(
    kind="perturb_M2", init=dict(exponent=5), precision=360, x=_MN_PTS[5][0],
    y=_MN_PTS[5][1], dx="1.e-340", nx=32,
    calc=dict(max_iter=30000, BLA_eps=1e-6, interior_detect=False,
              calc_dzndc=False, **_STD))
CASES["p_M2n_E20"] = dict(        # exponent 2 through the binomial forms
    kind="perturb_M2", init=dict(exponent=2), precision=30,
    x="-1.74928893611435556407228", y="0.", dx="5.e-20", nx=48,
    calc=dict(max_iter=50000, BLA_eps=1e-6, interior_detect=True,
              calc_dzndc=True, **_STD))
# same model on a boundary point of iteration depth ~5000: every pixel differs,
# and the view is chaotic at the fp64 resolution (low fastmath floor)
CASES["p_M3_E20_chaotic"] = dict(
    kind="perturb_M2", init=dict(exponent=3), precision=40,
    x="0.21459138481025131313267883316524207",
    y="0.72918276962050262626535766633048414", dx="3.e-20", nx=48,
    calc=dict(max_iter=20000, BLA_eps=1e-6, interior_detect=True,
              calc_dzndc=True, **_STD))

# ---- Mandelbrot_N, standard loop (models/mandelbrot_Mn.py:20-350) ----
CASES["std_M3"] = dict(        # z^2 is a product in numba: bit-defined
    kind="std_M2", init=dict(exponent=3), x=0.0, y=0.0, dx=3.0, nx=64,
    calc=dict(max_iter=2000, M_divergence=1000., epsilon_stationnary=1e-3))
CASES["std_M6_d2"] = dict(     # examples/interactive_standard/S03: exponent 6
    kind="std_M2", init=dict(exponent=6), x=0.0, y=0.0, dx=2.6, nx=64,
    theta_deg=15.,
    calc=dict(max_iter=2000, M_divergence=1000., epsilon_stationnary=1e-3,
              calc_d2zndc2=True))
CASES["std_M3_orbit"] = dict(  # calc_orbit: back-shift with zn ** 3 + c
    kind="std_M2", init=dict(exponent=3), x=0.1, y=0.05, dx=2.5, nx=48, theta_deg=10.,
    calc=dict(max_iter=2000, M_divergence=1000., epsilon_stationnary=1e-3,
              calc_orbit=True, backshift=3))
CASES["std_M4_zoom"] = dict(
    kind="std_M2", init=dict(exponent=4), x=-0.6548, y=0.4686, dx=2e-3, nx=64,
    calc=dict(max_iter=3000, M_divergence=1000., epsilon_stationnary=1e-3))
