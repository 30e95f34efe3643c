# -*- coding: utf-8 -*-
"""
gen_golden.py -- generate the committed fixtures tests/golden/*.npz by running
the live reference (imported from /root/reference/src through
tools/ref_harness.py).  Container-only; the fixtures travel, the reference
does not.

Every case of tests/cases.py is run in two reference modes:

  fast    the reference exactly as shipped (numba fastmath=True loops): the
          arrays a user of the reference gets.  Integer outputs are the
          exact-parity targets, Z is a tolerance target.
  strict  the same sources compiled with fastmath=False (FS_REF_STRICT=1, see
          ref_harness.load_reference): an IEEE-strict sequence which the C
          oracle must reproduce BIT FOR BIT, Z included.  This is what pins
          the restatement.

usage:  python tools/gen_golden.py [case ...]      (both modes, subprocesses)
"""
import hashlib
import json
import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(REPO, "tests"))
GOLDEN = os.path.join(REPO, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def sample_idx(n, k=96):
    if n <= 0:
        return np.zeros(0, np.int64)
    idx = np.unique(np.concatenate([
        np.arange(min(n, 24)), np.linspace(0, n - 1, min(n, k)).astype(np.int64),
        np.arange(max(0, n - 8), n)]))
    return idx.astype(np.int64)


def run_one(name, mode):
    import ref_harness as rh
    import ref_tables as rt
    import oracle_lib as ol
    from cases import CASES
    case = CASES[name]
    t0 = time.time()
    r = rh.run_case(case)
    t_ref = time.time() - t0
    f = r["fractal"]
    out = {"Z": r["Z"], "U": r["U"], "stop_reason": r["stop_reason"],
           "stop_iter": r["stop_iter"], "c_pix_sha": sha(r["c_pix"]),
           "xy_ratio": float(f.xy_ratio),
           "nx": r["nx"], "ny": r["ny"], "lin_mat": r["lin_mat"]}
    import numba
    meta = {"case": name, "mode": mode, "numba": numba.__version__,
            "numpy": np.__version__, "ref_seconds": round(t_ref, 3)}
    if case["kind"].startswith("std"):
        import ref_tables as rt2
        pj = rt2.proj_from_reference(f.projection)
        r["c_pix_raw"] = r["c_pix"]
        r["c_pix"] = ol.project(pj, r["c_pix"], False)
        if case["kind"] == "std_M2" and "exponent" in case.get("init", {}):
            Z, U, sr, si = ol.std_mn(f.exponent, r["c_pix"], complex(f.x, f.y), f.dx,
                                     f.lin_mat, **case["calc"])
        elif case["kind"] == "std_M2":
            Z, U, sr, si = ol.std_m2(r["c_pix"], complex(f.x, f.y), f.dx,
                                     f.lin_mat, **case["calc"])
        else:
            import fractalshades.models.burning_ship as bs
            Z, U, sr, si = ol.std_bs(bs.get_flavor_int(f.flavor), r["c_pix"],
                                     complex(f.x, f.y), f.dx, f.lin_mat,
                                     **case["calc"])
    else:
        t = rt.tables_from_reference(f, r["indep"])
        scal = {k: (v if not isinstance(v, complex) else [v.real, v.imag])
                for k, v in t.items()
                if isinstance(v, (int, float, complex, bool, str, type(None)))}
        meta["scalars"] = scal
        # entries past the escape index of the reference point are
        # uninitialised memory in the reference (np.empty): hash the valid part
        n_valid = min(len(t["Zn_path"]), t["ref_div_iter"] + 1)
        out["Zn_sha"] = sha(t["Zn_path"][:n_valid])
        out["Zn_valid"] = n_valid
        out["L"] = len(t["Zn_path"])
        if t.get("ref_index_xr") is not None:
            for k in ("ref_index_xr", "ref_xr", "ref_xr_e", "refx_xr",
                      "refx_xr_e", "refy_xr", "refy_xr_e"):
                if t.get(k) is not None:
                    out[k] = t[k]
        L = len(t["Zn_path"])
        ip = sample_idx(L)
        out["samp_path"] = ip
        for k in ("dZndc", "dXnda", "dXndb", "dYnda", "dYndb"):
            if t.get(k) is not None:
                out["samp_" + k] = t[k][ip]
                if t.get(k + "_e") is not None:
                    out["samp_" + k + "_e"] = t[k + "_e"][ip]
        if t.get("dZndz") is not None:
            iz = sample_idx(L + 1)
            out["samp_pathz"] = iz
            out["samp_dZndz"] = t["dZndz"][iz]
            if t.get("dZndz_e") is not None:
                out["samp_dZndz_e"] = t["dZndz_e"][iz]
        if t.get("M_bla") is not None:
            ib = sample_idx(t["bla_len"], 160)
            out["samp_bla"] = ib
            w = len(t["M_bla"]) // t["bla_len"]
            out["samp_M_bla"] = t["M_bla"].reshape(t["bla_len"], w)[ib]
            out["samp_r_bla"] = t["r_bla"][ib]
        Z, U, sr, si, cnt = ol.perturb(t, r["c_pix"])
        meta["oracle_counters"] = [int(x) for x in cnt]
    # report
    eq_i = np.mean(si == r["stop_iter"])
    eq_r = np.mean(sr == r["stop_reason"])
    eq_u = np.mean(U == r["U"]) if r["U"].size else 1.0
    zbit = float(np.mean((Z == r["Z"]) | (np.isnan(Z) & np.isnan(r["Z"]))))
    meta["oracle_vs_ref"] = {"stop_iter_exact": float(eq_i),
                             "stop_reason_exact": float(eq_r),
                             "U_exact": float(eq_u), "Z_bit_exact": zbit}
    print(f"[{mode:6s}] {name:28s} n={r['stop_iter'].size:6d} "
          f"sum_iter={int(r['stop_iter'].sum()):12d} ref {t_ref:6.1f}s | oracle: "
          f"stop_iter {eq_i * 100:8.4f}% reason {eq_r * 100:8.4f}% U "
          f"{eq_u * 100:8.4f}% Z bit-exact {zbit * 100:8.4f}%", flush=True)
    out["meta"] = json.dumps(meta, default=str)
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, f"{name}.{mode}.npz"), **out)


def main(argv):
    if len(argv) >= 2 and argv[0] == "--worker":
        mode = argv[1]
        for name in argv[2:]:
            run_one(name, mode)
        return
    from cases import CASES
    names = argv if argv else list(CASES)
    for mode in ("strict", "fast"):
        env = dict(os.environ)
        env["FS_REF_STRICT"] = "1" if mode == "strict" else "0"
        env["PYTHONWARNINGS"] = "ignore"
        p = subprocess.run([sys.executable, os.path.abspath(__file__),
                            "--worker", mode] + names, env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           text=True)
        for line in p.stdout.splitlines():
            if line.startswith("[") or "Error" in line or "Traceback" in line \
                    or line.startswith("  File") or "error" in line.lower():
                print(line)
        if p.returncode != 0:
            print(p.stdout[-4000:])
            raise SystemExit(p.returncode)


if __name__ == "__main__":
    main(sys.argv[1:])
