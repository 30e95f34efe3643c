# -*- coding: utf-8 -*-
""" Quick GPU-vs-oracle sweep over all parity cases (run under gpurun). """
import sys, os, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
import numpy as np
from fractalshades_b200 import build
build.build_orbit()
import parity_common as pc
from cases import CASES

names = sys.argv[1:] or list(CASES)
for name in names:
    t0 = time.time()
    Zo, Uo, sro, sio, ex = pc.run_oracle(name)
    t1 = time.time()
    for strict in (True, False):
        try:
            Z, U, sr, si, gx = pc.run_gpu_case(name, strict)
        except Exception as e:
            print(f"{name:28s} strict={strict} FAILED: {e}")
            continue
        m = (si == sio)[0] & (sr == sro)[0]
        with np.errstate(all="ignore"):
            rel = np.abs(Z[:, m] - Zo[:, m]) / np.maximum(np.abs(Zo[:, m]), 1e-300)
        rel = rel[np.isfinite(rel)]
        st = gx.get("stats") or {}
        msg = (f"{name:28s} strict={int(strict)} n={si.size:5d} iter_exact {np.mean(si == sio) * 100:8.4f}% "
               f"reason {np.mean(sr == sro) * 100:8.4f}% U {pc.frac_same(U, Uo) * 100:8.4f}% "
               f"Z bits {pc.frac_same(Z, Zo) * 100:8.4f}% maxrel {rel.max() if rel.size else 0:.2e} "
               f"kernel {st.get('kernel_ms', 0):8.3f} ms")
        if "bla" in gx and "tables" in ex and ex["tables"].get("M_bla") is not None:
            M, r, n, stg = gx["bla"]
            msg += f" | BLA M {pc.frac_same(M, ex['tables']['M_bla']) * 100:.3f}% r {pc.frac_same(r, ex['tables']['r_bla']) * 100:.3f}%"
        if "dzndc" in gx and "tables" in ex:
            d, de = gx["dzndc"]
            t = ex["tables"]
            ref = t["dZndc"] if t["kind"] == "perturb_M2" else np.stack([t[k] for k in ("dXnda", "dXndb", "dYnda", "dYndb")])
            msg += f" dZndc {pc.frac_same(d, ref) * 100:.3f}%"
        print(msg, flush=True)
    print(f"   oracle {t1 - t0:.2f}s")
