# -*- coding: utf-8 -*-
"""
Pixel grids of the live reference (Fractal.chunk_pixel_pos, core.py:1767-1830)
with jitter and supersampling -> tests/golden/pixel_grid.npz (sha-256 of every
grid + a few sampled values).  Container-only (needs /root/reference).
"""
import hashlib, json, os, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh

GRIDS = [  # (nx, xy_ratio, jitter, supersampling)
    (50, 1.3, False, None), (50, 1.3, 0.7, None), (50, 1.3, False, 3), (50, 1.3, 0.7, 3),
    (450, 1.0, 1.0, 2), (333, 16 / 9., 0.25, None),
]


def main():
    fs = rh.load_reference()
    import fractalshades.models as fsm
    out = {}
    meta = []
    for k, (nx, ratio, jitter, ss) in enumerate(GRIDS):
        f = fsm.Mandelbrot(tempfile.mkdtemp())
        f.zoom(x=-0.5, y=0.1, dx=2.5, nx=nx, xy_ratio=ratio, theta_deg=0.,
               projection=fs.projection.Cartesian())
        f.complex_type = np.complex128      # normally set by calc_std_div (float_type)
        shas, samples = [], []
        for cs in f.chunk_slices():
            p = np.ascontiguousarray(f.chunk_pixel_pos(cs, jitter, ss))
            shas.append(hashlib.sha256(p.tobytes()).hexdigest())
            flat = p.ravel()
            idx = np.linspace(0, flat.size - 1, 16).astype(np.int64)
            samples.append(flat[idx])
        out[f"samples_{k}"] = np.concatenate(samples)
        meta.append(dict(nx=nx, xy_ratio=ratio, jitter=jitter, supersampling=ss, sha=shas))
    out["meta"] = json.dumps(meta)
    np.savez_compressed(os.path.join(os.path.dirname(HERE), "tests", "golden", "pixel_grid.npz"), **out)
    print("wrote", len(GRIDS), "grids")


if __name__ == "__main__":
    main()
